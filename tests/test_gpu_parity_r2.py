"""GPU parity, round 2: the holes the round-1 review named.
 * BASELINE config 4's modulus chain (BFV N = 2^15, the reference's default 128-bit modulus):
   rotation bit-exact against the reference's own kernels.
 * BFV Method II with |P| = 3: the reference's digits have size 2 (contextpool.hpp:29); multiply +
   relinearize + rotation against the reference kernels fed with the reference context's own tables.
 * all-zero / impulse / all-(p-1) ciphertexts through apply_galois (CKKS Method I and II, BFV):
   equal to the reference kernels except for the one documented difference -- the reference's
   coefficient-domain permute negates without a zero check (p - 0 = p, switchkey.cu:1692-1695) and
   its NTT can carry that p to the output; this engine permutes NTT words and never negates, so
   it stores the canonical 0 where the reference stores p.
 * rotate_rows(shift = 0) is the identity (ckks/operator.cuh:1123, bfv/operator.cuh:591-595)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O, ref as R
from tests.common import PARAMS, ciphertext, eval_key, oracle_ctx, residues
from tests.gpu_common import gpu_ctx, ref_gpu, to_dev, to_host

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.have_gpu(), reason="oracle/_ref/libref_gpu.so not built")


def _api():
    from heongpu_b200 import api
    return api


# ------------------------------------------------------------------ BFV default chain (config 4) --
@needs_ref
def test_bfv_default_modulus_rotation_vs_reference_kernels():
    import bench
    api = _api()
    primes, n = bench.BFV_32768_MODULUS, 1 << 15
    Q = len(primes) - 1
    ctx = api.HEContext(15, q_values=primes[:Q], p_values=primes[Q:], plain_modulus=786433, device=0)
    t = R.tables_for_refgpu(15, primes, Q, 1, scheme="BFV")
    rg = R.RefGpu(15, primes, Q, 1, t)
    op = api.HEArithmeticOperator(ctx)
    a = residues(201, primes[:Q], n, (2, 2))
    for i, steps in enumerate((1, -1, 64, -128)):
        key = to_dev(residues(210 + i, primes, n, (Q, 2)))
        elt = api.lib.heon_steps_to_galois_elt(steps, n, 3)
        assert elt == bench.galois_elt(steps, n, 3)
        A = api.Ciphertext(ctx, to_dev(a))
        A.in_ntt_domain_ = False
        out = api.Ciphertext(ctx, torch.zeros(2, 2, Q, n, dtype=torch.int64, device="cuda"))
        op.apply_galois_bfv(A, out, api.Galoiskey(ctx, {elt: key}), elt)
        for bi in range(2):
            ro = torch.zeros(2, Q, n, dtype=torch.int64, device="cuda")
            R.bfv_apply_galois(rg, to_dev(a[bi]), ro, key, elt)
            torch.cuda.synchronize()
            assert torch.equal(out.data[bi], ro), f"default-chain rotation by {steps} differs from the reference kernels"


# ------------------------------------------------------------------ BFV Method II, |P| = 3 --------
@needs_ref
@pytest.mark.parametrize("log_n,qb,pb,t", [(13, [59, 59, 59, 59, 59], [60, 60, 60], 786433),
                                           (12, [40, 40, 40], [41, 41, 41], 1032193)])
def test_bfv_method2_three_special_primes_vs_reference_kernels(log_n, qb, pb, t):
    api = _api()
    n, Q, K = 1 << log_n, len(qb), len(pb)
    primes = O.generate_primes(n, qb + pb)
    ob = O.BfvOracle(log_n, primes, Q, K, t)
    ctx = api.HEContext(log_n, q_values=primes[:Q], p_values=primes[Q:], plain_modulus=t, device=0)
    d = (Q + 1) // 2
    assert ctx.digits(0) == d == ob.digits()
    tab = R.tables_for_refgpu(log_n, primes, Q, K, scheme="BFV", plain_modulus=t)
    assert tab["method2"][0]["d"] == d
    rb, rg = R.RefBfv(ob), R.RefGpu(log_n, primes, Q, K, tab)
    a = residues(221, primes[:Q], n, (1, 2))
    b = residues(222, primes[:Q], n, (1, 2))
    key = residues(223, primes, n, (d, 2))
    dkey = to_dev(key)
    op = api.HEArithmeticOperator(ctx)
    A, B = api.Ciphertext(ctx, to_dev(a)), api.Ciphertext(ctx, to_dev(b))
    Cc = api.Ciphertext(ctx, torch.zeros(1, 3, Q, n, dtype=torch.int64, device="cuda"))
    op.multiply_bfv(A, B, Cc)
    rc = torch.zeros(3, Q, n, dtype=torch.int64, device="cuda")
    rb.multiply(to_dev(a[0]), to_dev(b[0]), rc)
    torch.cuda.synchronize()
    assert torch.equal(Cc.data[0], rc)
    mul_host = to_host(Cc.data)[0].copy()
    op.relinearize_inplace_bfv(Cc, api.Relinkey(ctx, dkey))
    R.bfv_relinearize(rg, rc, dkey)
    torch.cuda.synchronize()
    assert torch.equal(Cc.data[0, :2], rc[:2]), "BFV Method-II relinearize (digit size 2) differs from the reference kernels"
    assert np.array_equal(to_host(Cc.data)[0, :2], ob.relinearize(mul_host, key)[:2])
    elt = api.lib.heon_steps_to_galois_elt(3, n, 3)
    out = api.Ciphertext(ctx, torch.zeros(1, 2, Q, n, dtype=torch.int64, device="cuda"))
    op.apply_galois_bfv(A, out, api.Galoiskey(ctx, {elt: dkey}), elt)
    ro = torch.zeros(2, Q, n, dtype=torch.int64, device="cuda")
    R.bfv_apply_galois(rg, to_dev(a[0]), ro, dkey, elt)
    torch.cuda.synchronize()
    assert torch.equal(out.data[0], ro)


# ------------------------------------------------------------------ edge ciphertexts through apply_galois
def _edge_ciphertexts(primes, L, n):
    pr = np.array(primes[:L], dtype=np.uint64)
    e = np.zeros((5, 2, L, n), dtype=np.uint64)
    # 0: all zero
    e[1] = (pr - 1)[None, :, None]  # all p-1
    e[2, :, :, 0] = 1  # impulse at 0 (both components)
    e[3, 1, :, n - 1] = pr - 1  # c1 = -X^(n-1), c0 = 0
    e[4, 0] = residues(231, primes[:L], n)  # random c0, zero c1: key switch contributes nothing
    return e


def _assert_equal_up_to_zero_vs_p(ours, theirs, primes, L, what):
    """Equal, except where this engine stores the canonical 0 and the reference stores p
    (its unchecked negation p - 0)."""
    ours, theirs = to_host(ours), to_host(theirs)
    diff = ours != theirs
    if not diff.any():
        return 0
    pr = np.array(primes[:L], dtype=np.uint64).reshape(1, L, 1)
    pfull = np.broadcast_to(pr, ours.shape)
    ok = (ours[diff] == 0) & (theirs[diff] == pfull[diff])
    assert ok.all(), f"{what}: {int((~ok).sum())} words differ beyond the documented 0-vs-p case"
    return int(diff.sum())


@needs_ref
@pytest.mark.parametrize("name,depth", [("n12_I", 0), ("n13_II", 0), ("n13_II", 2), ("n16_I_small", 0), ("n16_II_small", 1)])
def test_ckks_apply_galois_edge_ciphertexts_vs_reference_kernels(name, depth):
    api = _api()
    ctx, oc, rg = gpu_ctx(name), oracle_ctx(name), ref_gpu(name)
    n, L = oc.n, oc.Q - depth
    e = _edge_ciphertexts(oc.primes, L, n)
    key = to_dev(eval_key(232, oc.primes, oc.digits(0), n))
    op = api.HEArithmeticOperator(ctx)
    seen = 0
    for elt in (5, 2 * n - 1):
        gk = api.Galoiskey(ctx, {elt: key})
        A = api.Ciphertext(ctx, to_dev(e), depth=depth)
        out = api.Ciphertext(ctx, torch.zeros(e.shape[0], 2, L, n, dtype=torch.int64, device="cuda"), depth=depth)
        op.apply_galois(A, out, gk, elt)
        for i in range(e.shape[0]):
            ro = torch.zeros(2, L, n, dtype=torch.int64, device="cuda")
            rg.apply_galois(to_dev(e[i]), ro, key, elt, depth)
            torch.cuda.synchronize()
            seen += _assert_equal_up_to_zero_vs_p(out.data[i], ro, oc.primes, L, f"{name} depth {depth} elt {elt} vector {i}")
        # this engine's output is canonical everywhere
        pr = np.array(oc.primes[:L], dtype=np.uint64).reshape(1, 1, L, 1)
        assert (to_host(out.data) < pr).all()
    print(f"{name} depth {depth}: {seen} words where the reference stores p for 0")


@needs_ref
@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II", "bfv_n15_II"])
def test_bfv_apply_galois_edge_ciphertexts_vs_reference_kernels(name):
    """BFV automorphisms run in the coefficient domain through the same permute-with-negation as the
    reference's divide_round_lastq_permute_bfv_kernel: bit-identical, the stored p included."""
    api = _api()
    from tests import test_gpu_bfv as TB
    ob, oc = TB.bfv_oracle(name)
    ctx = TB.bfv_gpu_ctx(name)
    rb, rg = TB._ref_handles(name)
    n, Q = ob.n, ob.Q
    e = _edge_ciphertexts(ob.primes, Q, n)
    key = to_dev(residues(233, ob.primes, n, (ob.digits(), 2)))
    op = api.HEArithmeticOperator(ctx)
    for elt in (3, 2 * n - 1):
        A = api.Ciphertext(ctx, to_dev(e))
        out = api.Ciphertext(ctx, torch.zeros(e.shape[0], 2, Q, n, dtype=torch.int64, device="cuda"))
        op.apply_galois_bfv(A, out, api.Galoiskey(ctx, {elt: key}), elt)
        for i in range(e.shape[0]):
            ro = torch.zeros(2, Q, n, dtype=torch.int64, device="cuda")
            R.bfv_apply_galois(rg, to_dev(e[i]), ro, key, elt)
            torch.cuda.synchronize()
            assert torch.equal(out.data[i], ro), f"{name} elt {elt} vector {i}"


# ------------------------------------------------------------------ shift 0 -----------------------
def test_rotate_rows_shift_zero_is_identity():
    api = _api()
    ctx, oc = gpu_ctx("n12_I"), oracle_ctx("n12_I")
    n, L = oc.n, oc.Q
    a = ciphertext(241, oc.primes, L, n, 2, 2)
    key = to_dev(eval_key(242, oc.primes, oc.digits(0), n))
    op = api.HEArithmeticOperator(ctx)
    gk = api.Galoiskey(ctx, {2 * n - 1: key, 5: key})  # the conjugation key is present: it must NOT be used
    A = api.Ciphertext(ctx, to_dev(a))
    out = api.Ciphertext(ctx, torch.zeros(2, 2, L, n, dtype=torch.int64, device="cuda"))
    op.rotate_rows(A, out, gk, 0)
    assert np.array_equal(to_host(out.data), a)
    hoist = torch.zeros(2, 2, 2, L, n, dtype=torch.int64, device="cuda")
    op.rotate_rows_hoisted(A, hoist, gk, [0, 1])
    assert np.array_equal(to_host(hoist[0]), a)
    one = api.Ciphertext(ctx, torch.zeros(2, 2, L, n, dtype=torch.int64, device="cuda"))
    op.rotate_rows(A, one, gk, 1)
    assert torch.equal(hoist[1], one.data)
    # BFV
    from tests import test_gpu_bfv as TB
    ob, _ = TB.bfv_oracle("bfv_n12_I")
    bctx = TB.bfv_gpu_ctx("bfv_n12_I")
    b = residues(243, ob.primes[: ob.Q], ob.n, (1, 2))
    bkey = to_dev(residues(244, ob.primes, ob.n, (ob.digits(), 2)))
    bop = api.HEArithmeticOperator(bctx)
    Bc = api.Ciphertext(bctx, to_dev(b))
    Bo = api.Ciphertext(bctx, torch.zeros(1, 2, ob.Q, ob.n, dtype=torch.int64, device="cuda"))
    bop.rotate_rows_bfv(Bc, Bo, api.Galoiskey(bctx, {2 * ob.n - 1: bkey}), 0)
    assert np.array_equal(to_host(Bo.data), b)


# ------------------------- BSGS mat-vec with double hoisting (multiply_matrix_v2) --
def _bsgs_case(api, name, depth, seed, rot_n2, rot_n1, diags_bsgs):
    ctx, oc, rg = gpu_ctx(name), oracle_ctx(name), ref_gpu(name)
    L, n, K = oc.Q - depth, oc.n, oc.K
    pql = L + K
    a = ciphertext(seed, oc.primes, L, n, 2, 1)
    baby, giant, sizes, terms = api.HEArithmeticOperator.bsgs_plan(n, 5, diags_bsgs, rot_n1, rot_n2)
    keys = {}
    for i, e in enumerate(sorted(set(baby + giant) - {0})):
        keys[e] = to_dev(eval_key(seed + 10 + i, oc.primes, oc.digits(0), n))
    pq_primes = list(oc.primes[:L]) + list(oc.primes[oc.Q:])
    matrix = to_dev(residues(seed + 5, pq_primes, n, (len(terms),)))
    assert matrix.shape == (len(terms), pql, n)
    return ctx, oc, rg, L, n, a, keys, matrix, (baby, giant, sizes, terms)


@needs_ref
@pytest.mark.parametrize("name,depth", [("n13_II", 0), ("n13_II", 2), ("n15_II", 1), ("C3_II", 3)])
def test_bsgs_matvec_bit_exact_vs_reference_kernels(name, depth):
    """heon_ckks_multiply_matrix against the replay of multiply_matrix_v2 (ckks/operator.cu:2898-3390) on the
    reference's own kernels: baby steps {0,1,2,3}, giant steps {0,4,8} with ragged groups, then rescale."""
    from heongpu_b200 import api
    rot_n2 = [3, 0, 1, 2]  # unsorted on purpose: the reference sorts it (:2926)
    rot_n1 = [0, 4, 8]
    diags_bsgs = [[0, 1, 3], [4, 5, 6, 7], [9, 11]]
    ctx, oc, rg, L, n, a, keys, matrix, plan = _bsgs_case(api, name, depth, 300, rot_n2, rot_n1, diags_bsgs)
    baby, giant, sizes, terms = plan
    op = api.HEArithmeticOperator(ctx)
    A = api.Ciphertext(ctx, to_dev(a), depth=depth)
    out = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"), depth=depth)
    gk = api.Galoiskey(ctx, keys)
    op.multiply_matrix(A, out, matrix, diags_bsgs, rot_n1, rot_n2, gk, rescale=False)
    ro = torch.zeros(2, L, n, dtype=torch.int64, device="cuda")
    rg.bsgs_matvec(to_dev(a[0]), ro, matrix, baby, [keys.get(e) for e in baby], giant, [keys.get(e) for e in giant],
                   sizes, terms, depth)
    torch.cuda.synchronize()
    assert torch.equal(out.data[0], ro), "BSGS mat-vec differs from the reference kernels"
    assert torch.equal(A.data[0], to_dev(a[0])), "input ciphertext modified"
    op.rescale_inplace(out)
    rg.rescale(ro, depth)
    torch.cuda.synchronize()
    assert torch.equal(out.words()[0], ro.reshape(-1)[: 2 * (L - 1) * n].reshape(2, L - 1, n))


@needs_ref
def test_bsgs_matvec_single_group_and_missing_key():
    from heongpu_b200 import api
    rot_n2, rot_n1, diags_bsgs = [1, 2], [16], [[17, 18]]  # no zero rotation anywhere
    ctx, oc, rg, L, n, a, keys, matrix, plan = _bsgs_case(api, "n13_II", 1, 320, rot_n2, rot_n1, diags_bsgs)
    baby, giant, sizes, terms = plan
    op = api.HEArithmeticOperator(ctx)
    A = api.Ciphertext(ctx, to_dev(a), depth=1)
    out = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"), depth=1)
    op.multiply_matrix(A, out, matrix, diags_bsgs, rot_n1, rot_n2, api.Galoiskey(ctx, keys), rescale=False)
    ro = torch.zeros(2, L, n, dtype=torch.int64, device="cuda")
    rg.bsgs_matvec(to_dev(a[0]), ro, matrix, baby, [keys.get(e) for e in baby], giant, [keys.get(e) for e in giant],
                   sizes, terms, 1)
    torch.cuda.synchronize()
    assert torch.equal(out.data[0], ro)
    some = next(iter(keys))
    fewer = {e: k for e, k in keys.items() if e != some}
    with pytest.raises(api.HeonLogicError):
        op.multiply_matrix(A, out, matrix, diags_bsgs, rot_n1, rot_n2, api.Galoiskey(ctx, fewer))
    with pytest.raises(api.HeonError):  # Method I context: the reference's path is Method II only
        c1 = gpu_ctx("n13_I")
        op1 = api.HEArithmeticOperator(c1)
        o1 = oracle_ctx("n13_I")
        a1 = api.Ciphertext(c1, to_dev(ciphertext(1, o1.primes, o1.Q, o1.n, 2, 1)), depth=0)
        k1 = {e: to_dev(eval_key(2, o1.primes, o1.digits(0), o1.n)) for e in set(baby + giant)}
        m1 = to_dev(residues(3, list(o1.primes), o1.n, (len(terms),)))
        op1.multiply_matrix(a1, a1, m1, diags_bsgs, rot_n1, rot_n2, api.Galoiskey(c1, k1))


@pytest.mark.parametrize("name,depth,count", [("n12_II", 0, 1), ("n13_II", 1, 5), ("mixed", 0, 9), ("n13_I", 2, 16)])
def test_multiply_plain_accumulate_vs_integer_arithmetic(name, depth, count):
    """cipherplain_multiply_accumulate_kernel (multiplication.cu:374-403): sum of canonical products, canonical."""
    from heongpu_b200 import api
    ctx, oc = gpu_ctx(name), oracle_ctx(name)
    L, n = oc.Q - depth, oc.n
    cts = residues(400, oc.primes[:L], n, (count, 2))
    pts = residues(401, oc.primes[:L], n, (count,))
    out = torch.zeros(2, L, n, dtype=torch.int64, device="cuda")
    api.HEArithmeticOperator(ctx).multiply_plain_accumulate(to_dev(cts), to_dev(pts), out, depth)
    want = np.zeros((2, L, n), dtype=object)
    p = np.array([int(x) for x in oc.primes[:L]], dtype=object).reshape(1, L, 1)
    for i in range(count):
        want = (want + cts[i].astype(object) * pts[i].astype(object)[None]) % p
    assert np.array_equal(to_host(out), want.astype(np.uint64))


# ------------------------------------------------------------------ HOST-resident operands ---------
@pytest.mark.parametrize("name,chunk,rescale", [("n13_II", 1, False), ("n13_II", 2, True), ("n13_II", 3, False),
                                                ("n12_I", 2, True), ("n16_II_small", 2, False)])
def test_host_operand_pipeline_equals_device_path(name, chunk, rescale):
    """heon_ckks_multiply_relinearize_host (storage_type::HOST operands, storagemanager.cuh:113-167) returns the words
    of the device-resident multiply + relinearize_inplace (+ rescale_inplace): on one caller stream, and with calls
    issued alternately on two caller streams (they overlap inside the library's pipeline; ragged last chunk)."""
    api = _api()
    ctx, oc = gpu_ctx(name), oracle_ctx(name)
    batch, L, n = 5, oc.Q, oc.n
    a = ciphertext(301, oc.primes, L, n, 2, batch)
    b = ciphertext(302, oc.primes, L, n, 2, batch)
    key = to_dev(eval_key(303, oc.primes, oc.digits(0), n))
    op = api.HEArithmeticOperator(ctx)
    rk = api.Relinkey(ctx, key)
    A, B = api.Ciphertext(ctx, to_dev(a)), api.Ciphertext(ctx, to_dev(b))
    Cc = api.Ciphertext(ctx, torch.zeros(batch, 3, L, n, dtype=torch.int64, device="cuda"))
    op.multiply(A, B, Cc)
    op.relinearize_inplace(Cc, rk)
    if rescale:
        op.rescale_inplace(Cc)
    torch.cuda.synchronize()
    want = Cc.words().cpu()[:, :2]
    Lout = L - 1 if rescale else L
    ha, hb = torch.from_numpy(a.astype(np.int64)).pin_memory(), torch.from_numpy(b.astype(np.int64)).pin_memory()
    res = [torch.zeros(batch, 2, Lout, n, dtype=torch.int64).pin_memory() for _ in range(4)]
    op.multiply_relinearize_host(ha, hb, res[0], rk, depth=0, rescale=rescale, chunk=chunk)
    torch.cuda.synchronize()
    assert torch.equal(res[0], want), "host-operand path differs from the device path"
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for i in range(1, 4):
        with torch.cuda.stream(streams[i & 1]):
            op.multiply_relinearize_host(ha, hb, res[i], rk, depth=0, rescale=rescale, chunk=chunk)
    torch.cuda.synchronize()
    for i in range(1, 4):
        assert torch.equal(res[i], want), f"overlapped call {i} differs from the device path"

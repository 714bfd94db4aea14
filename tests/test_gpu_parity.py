"""GPU parity tests: the CUDA path (through the C ABI) against
 (a) the CPU oracle on seeded inputs at sizes it finishes in seconds,
 (b) the reference's OWN CUDA kernels (oracle/_ref/libref_gpu.so) at the
     BASELINE sizes -- bit-exact memcmp of whole buffers,
 (c) size-independent properties (INTT(NTT(x)) = x, linearity).
Integer work: the bar is bit-exact everywhere."""
import numpy as np
import pytest
import torch

from oracle import ref as R
from tests.common import PARAMS, ciphertext, eval_key, oracle_ctx, residues
from tests.gpu_common import gpu_ctx, ref_gpu, to_dev, to_host

pytestmark = pytest.mark.gpu

SMALL = ["n12_I", "n12_II", "n13_I", "n13_II", "mixed"]
needs_ref = pytest.mark.skipif(not R.have_gpu(), reason="oracle/_ref/libref_gpu.so not built")


def _api():
    from heongpu_b200 import api
    return api


# ---------------------------------------------------------------- NTT -----
@pytest.mark.parametrize("name", ["n12_I", "n13_II", "n14_C2", "n15_II", "n16_I_small"])
def test_ntt_matches_oracle_and_roundtrips(name):
    ctx, oc = gpu_ctx(name), oracle_ctx(name)
    order = ctx.level_primes(0)
    x = residues(1, [oc.primes[i] for i in order], oc.n, (2,))  # 2 polys per prime
    d = to_dev(x)
    ctx.ntt(d, order)
    torch.cuda.synchronize()
    want = oc.ntt(x, order)
    assert np.array_equal(to_host(d), want)
    ctx.ntt(d, order, inverse=True)
    assert np.array_equal(to_host(d), x)
    # inverse of arbitrary canonical data vs oracle
    d2 = to_dev(x)
    ctx.ntt(d2, order, inverse=True)
    assert np.array_equal(to_host(d2), oc.ntt(x, order, inverse=True))


@pytest.mark.parametrize("name", ["n12_I", "n16_I_small"])
def test_ntt_edge_vectors(name):
    ctx, oc = gpu_ctx(name), oracle_ctx(name)
    order = list(range(oc.Qp))
    pr = np.array(oc.primes, dtype=np.uint64)
    n = oc.n
    vecs = np.zeros((4, oc.Qp, n), dtype=np.uint64)
    vecs[1] = (pr - 1)[:, None]  # all p-1
    vecs[2, :, 0] = 1  # impulse
    vecs[3, :, n - 1] = pr - 1  # impulse at the far end
    for inverse in (False, True):
        d = to_dev(vecs)
        ctx.ntt(d, order, inverse=inverse)
        assert np.array_equal(to_host(d), oc.ntt(vecs, order, inverse=inverse))


def test_ntt_out_of_place_and_poly_ordered():
    ctx, oc = gpu_ctx("n13_II"), oracle_ctx("n13_II")
    n = oc.n
    x = residues(3, oc.primes[:4], n)
    src, dst = to_dev(x), torch.zeros(4, n, dtype=torch.int64, device="cuda")
    ctx.ntt(src, [0, 1, 2, 3], out=dst)
    assert np.array_equal(to_host(src), x)
    assert np.array_equal(to_host(dst), oc.ntt(x, [0, 1, 2, 3]))
    # Poly_Ordered: only polys 1 and 3, both with prime 2
    y = residues(4, [oc.primes[2]] * 4, n)
    d = to_dev(y)
    ctx.ntt_poly_ordered(d, [1 * n, 3 * n], 2, inverse=True)
    want = y.copy()
    want[[1, 3]] = oc.ntt(y[[1, 3]], [2], inverse=True)
    assert np.array_equal(to_host(d), want)


@needs_ref
@pytest.mark.parametrize("name", ["C3_I", "C3_II"])
def test_ntt_bit_exact_vs_reference_kernels_full_size(name):
    ctx, oc, rg = gpu_ctx(name), oracle_ctx(name), ref_gpu(name)
    order = ctx.level_primes(0)
    x = residues(5, [oc.primes[i] for i in order], oc.n, (2,))
    ours, theirs = to_dev(x), to_dev(x)
    ctx.ntt(ours, order)
    rg.ntt_level(theirs, 0)
    torch.cuda.synchronize()
    assert torch.equal(ours, theirs)
    ctx.ntt(ours, order, inverse=True)
    rg.ntt_level(theirs, 0, inverse=True)
    assert torch.equal(ours, theirs)
    assert np.array_equal(to_host(ours), x)


# ---------------------------------------------------- operators vs oracle --
def _mk(name, depth, batch, seed):
    oc = oracle_ctx(name)
    L = oc.Q - depth
    a = ciphertext(seed, oc.primes, L, oc.n, 2, batch)
    b = ciphertext(seed + 1, oc.primes, L, oc.n, 2, batch)
    key = eval_key(seed + 2, oc.primes, oc.digits(0), oc.n)
    return oc, L, a, b, key


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("depth", [0, 1])
def test_multiply_relinearize_rescale_vs_oracle(name, depth):
    api = _api()
    ctx = gpu_ctx(name)
    batch = 2
    oc, L, a, b, key = _mk(name, depth, batch, 10)
    op = api.HEArithmeticOperator(ctx)
    A = api.Ciphertext(ctx, to_dev(a), depth=depth)
    B = api.Ciphertext(ctx, to_dev(b), depth=depth)
    Cc = api.Ciphertext(ctx, torch.zeros(batch, 3, L, oc.n, dtype=torch.int64, device="cuda"), depth=depth)
    rk = api.Relinkey(ctx, to_dev(key))
    op.multiply(A, B, Cc)
    got_mul = to_host(Cc.data).copy()
    op.relinearize_inplace(Cc, rk)
    got_rel = to_host(Cc.data).copy()
    if L >= 2:
        op.rescale_inplace(Cc)
        got_res = to_host(Cc.words()).copy()
    for bi in range(batch):
        m = oc.multiply(a[bi], b[bi], depth)
        assert np.array_equal(got_mul[bi], m), "multiply"
        r = oc.relinearize(m, key, depth)
        assert np.array_equal(got_rel[bi], r), "relinearize (all three components)"
        if L >= 2:
            s = oc.rescale(r[:2], depth)
            assert np.array_equal(got_res[bi], s), "rescale"


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("depth", [0, 1])
def test_apply_galois_vs_oracle(name, depth):
    api = _api()
    ctx = gpu_ctx(name)
    batch = 2
    oc, L, a, _, key = _mk(name, depth, batch, 20)
    op = api.HEArithmeticOperator(ctx)
    for shift in (1, -3):
        elt = api.lib.heon_steps_to_galois_elt(shift, oc.n, 5)
        gk = api.Galoiskey(ctx, {elt: to_dev(key)})
        A = api.Ciphertext(ctx, to_dev(a), depth=depth)
        out = api.Ciphertext(ctx, torch.zeros(batch, 2, L, oc.n, dtype=torch.int64, device="cuda"), depth=depth)
        op.rotate_rows(A, out, gk, shift)
        got = to_host(out.data)
        assert np.array_equal(to_host(A.data), a), "input must stay untouched"
        for bi in range(batch):
            assert np.array_equal(got[bi], oc.apply_galois(a[bi], key, elt, depth))


@pytest.mark.parametrize("name", ["n12_I", "n13_II"])
def test_add_sub_negate_mod_drop_vs_oracle(name):
    api = _api()
    ctx = gpu_ctx(name)
    oc, L, a, b, _ = _mk(name, 0, 3, 30)
    op = api.HEArithmeticOperator(ctx)
    A, B = api.Ciphertext(ctx, to_dev(a)), api.Ciphertext(ctx, to_dev(b))
    out = api.Ciphertext(ctx, torch.zeros_like(A.data))
    op.add(A, B, out)
    add = to_host(out.data).copy()
    op.sub(A, B, out)
    sub = to_host(out.data).copy()
    op.negate(A, out)
    neg = to_host(out.data).copy()
    op.mod_drop_inplace(A)
    drop = to_host(A.words()).copy()
    for bi in range(3):
        assert np.array_equal(add[bi], oc.addsub(a[bi], b[bi], 0))
        assert np.array_equal(sub[bi], oc.addsub(a[bi], b[bi], 1))
        assert np.array_equal(neg[bi], oc.addsub(a[bi], a[bi], 2))
        assert np.array_equal(drop[bi], oc.mod_drop(a[bi]))


def test_edge_ciphertexts_zero_and_max():
    """all-zero and all p-1 ciphertexts through multiply+relinearize+rescale."""
    api = _api()
    for name in ("n12_I", "n12_II"):
        ctx, oc = gpu_ctx(name), oracle_ctx(name)
        L, n = oc.Q, oc.n
        pr = np.array(oc.primes[:L], dtype=np.uint64)
        key = eval_key(41, oc.primes, oc.digits(0), n)
        for fill in ("zero", "max"):
            a = np.zeros((1, 2, L, n), dtype=np.uint64)
            if fill == "max":
                a[:] = (pr - 1)[None, None, :, None]
            op = api.HEArithmeticOperator(ctx)
            A, B = api.Ciphertext(ctx, to_dev(a)), api.Ciphertext(ctx, to_dev(a))
            Cc = api.Ciphertext(ctx, torch.zeros(1, 3, L, n, dtype=torch.int64, device="cuda"))
            op.multiply(A, B, Cc)
            op.relinearize_inplace(Cc, api.Relinkey(ctx, to_dev(key)))
            got = to_host(Cc.data)[0]
            want = oc.relinearize(oc.multiply(a[0], a[0]), key)
            assert np.array_equal(got, want), (name, fill)


def test_operator_state_errors():
    api = _api()
    ctx, oc = gpu_ctx("n12_I"), oracle_ctx("n12_I")
    a = ciphertext(50, oc.primes, oc.Q, oc.n, 2, 1)
    op = api.HEArithmeticOperator(ctx)
    A = api.Ciphertext(ctx, to_dev(a), relinearization_required=True)
    B = api.Ciphertext(ctx, to_dev(a))
    out = api.Ciphertext(ctx, torch.zeros(1, 3, oc.Q, oc.n, dtype=torch.int64, device="cuda"))
    with pytest.raises(api.HeonError):
        op.multiply(A, B, out)
    with pytest.raises(api.HeonError):
        op.relinearize_inplace(B, api.Relinkey(ctx, to_dev(a)))
    D = api.Ciphertext(ctx, to_dev(a[:, :, :2]), depth=1)
    with pytest.raises(api.HeonError):
        op.multiply(B, D, out)


# ------------------------------- full size vs the reference's CUDA kernels --
@needs_ref
@pytest.mark.parametrize("name,depth", [("C3_I", 0), ("C3_I", 5), ("C3_II", 0), ("C3_II", 4), ("n14_C2", 0)])
def test_mul_relin_rescale_bit_exact_vs_reference_kernels(name, depth):
    api = _api()
    ctx, oc, rg = gpu_ctx(name), oracle_ctx(name), ref_gpu(name)
    batch = 2
    L, n = oc.Q - depth, oc.n
    a = ciphertext(60, oc.primes, L, n, 2, batch)
    b = ciphertext(61, oc.primes, L, n, 2, batch)
    key = to_dev(eval_key(62, oc.primes, oc.digits(0), n))
    op = api.HEArithmeticOperator(ctx)
    A, B = api.Ciphertext(ctx, to_dev(a), depth=depth), api.Ciphertext(ctx, to_dev(b), depth=depth)
    Cc = api.Ciphertext(ctx, torch.zeros(batch, 3, L, n, dtype=torch.int64, device="cuda"), depth=depth)
    op.multiply(A, B, Cc)
    op.relinearize_inplace(Cc, api.Relinkey(ctx, key))
    ours_relin = Cc.data.clone()
    op.rescale_inplace(Cc)
    ours_rescale = Cc.words().clone()
    for bi in range(batch):
        ra, rb = to_dev(a[bi]), to_dev(b[bi])
        rc = torch.zeros(3, L, n, dtype=torch.int64, device="cuda")
        rg.multiply(ra, rb, rc, depth)
        rg.relinearize(rc, key, depth)
        torch.cuda.synchronize()
        assert torch.equal(ours_relin[bi], rc), "relinearize differs from the reference kernels"
        rg.rescale(rc, depth)
        torch.cuda.synchronize()
        theirs = rc.reshape(-1)[: 2 * (L - 1) * n].reshape(2, L - 1, n)
        assert torch.equal(ours_rescale[bi], theirs), "rescale differs from the reference kernels"


@needs_ref
@pytest.mark.parametrize("name,depth", [("C3_I", 0), ("C3_II", 0), ("C3_II", 3), ("n15_II", 1)])
def test_rotate_bit_exact_vs_reference_kernels(name, depth):
    api = _api()
    ctx, oc, rg = gpu_ctx(name), oracle_ctx(name), ref_gpu(name)
    L, n = oc.Q - depth, oc.n
    a = ciphertext(70, oc.primes, L, n, 2, 1)
    key = to_dev(eval_key(71, oc.primes, oc.digits(0), n))
    elt = api.lib.heon_steps_to_galois_elt(7, n, 5)
    op = api.HEArithmeticOperator(ctx)
    A = api.Ciphertext(ctx, to_dev(a), depth=depth)
    out = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"), depth=depth)
    op.apply_galois(A, out, api.Galoiskey(ctx, {elt: key}), elt)
    ra, ro = to_dev(a[0]), torch.zeros(2, L, n, dtype=torch.int64, device="cuda")
    rg.apply_galois(ra, ro, key, elt, depth)
    torch.cuda.synchronize()
    assert torch.equal(out.data[0], ro)




# ------------------------------------------- alternate code paths (opt-in) --
def _ctx_with_env(name, **env):
    """A fresh context created under the given HEON_* switches (read once, at context creation)."""
    import os
    api = _api()
    log_n, qb, pb = PARAMS[name]
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return api.HEContext(log_n, qb, pb, device=0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("env", [{"HEON_NTT_PIPE": 1}, {"HEON_NTT_FUSED": 1}, {"HEON_NTT_FP64": 0}, {"HEON_NTT_TMA": 0},
                                 {"HEON_ROW_TILE": 16}, {"HEON_ROW_TILE": 4}, {"HEON_COL_THREADS": 256},
                                 {"HEON_COL_THREADS": 128}, {"HEON_NTT_PERSISTENT": 1}, {"HEON_COL_TMA": 0}, {"HEON_COL_TMA_TILES": 1},
                                 {"HEON_COL_TMA_TILES": 2}, {"HEON_COL_TMA_TILES": 8}, {"HEON_COL_TMA_TILES": 16},
                                 {"HEON_COL_TMA_TILES": 4, "HEON_COL_TMA_BUFS": 3},
                                 {"HEON_COL_TMA_TILES": 16, "HEON_COL_TMA_BUFS": 3}, {"HEON_ROW_WALK": 0}, {"HEON_ROW_WALK": 2},
                                 {"HEON_ROW_WALK": 4}, {"HEON_ROW_WALK": 3}, {"HEON_ROW_WALK": 6}])
def test_alternate_ntt_paths_agree(env):
    """The opt-in transforms (warp-specialised pipelined kernel, ticket-ordered fused kernel), the
    integer-only butterflies and the LSU row pass must give the default path's words."""
    name = "n16_II_small"
    ctx, alt, oc = gpu_ctx(name), _ctx_with_env(name, **env), oracle_ctx(name)
    order = ctx.level_primes(0)
    x = residues(130, [oc.primes[i] for i in order], oc.n, (5,))  # 5 polys per prime (a ragged walk of 4 + 1)
    a, b = to_dev(x), to_dev(x)
    ctx.ntt(a, order)
    alt.ntt(b, order)
    torch.cuda.synchronize()
    assert torch.equal(a, b), env
    ctx.ntt(a, order, inverse=True)
    alt.ntt(b, order, inverse=True)
    assert torch.equal(a, b) and np.array_equal(to_host(a), x), env


@pytest.mark.parametrize("name,env", [("n16_II_small", {"HEON_NTT_PIPE": 1}), ("n13_II", {"HEON_GALOIS_NTT": 0}),
                                      ("n16_I_small", {"HEON_GALOIS_NTT": 0}), ("n13_II", {"HEON_NTT_FP64": 0}),
                                      ("n16_II_small", {"HEON_SKIP_OWN": 0}), ("n16_II_small", {"HEON_MAC_BY": 8}),
                                      # round 2: the fused row-pass + inner-product kernel against the separate kernels,
                                      # its 4-row tiling, and the mod-up fused into the column-pass load
                                      ("n16_II_small", {"HEON_ROW_MAC": 0}), ("n16_I_small", {"HEON_ROW_MAC": 0}),
                                      ("n13_II", {"HEON_ROW_MAC": 0}), ("n16_II_small", {"HEON_ROW_MAC_ROWS": 4}),
                                      ("n16_II_small", {"HEON_MODUP_FUSED": 1}), ("n13_II", {"HEON_MODUP_FUSED": 1}),
                                      ("n16_II_small", {"HEON_MODUP_FUSED": 1, "HEON_SKIP_OWN": 0}),
                                      ("n15_II", {"HEON_MODUP_FUSED": 1, "HEON_ROW_MAC": 0}),
                                      ("n16_II_small", {"HEON_COL_TMA": 0}), ("n16_I_small", {"HEON_COL_TMA": 0}),
                                      ("n16_II_small", {"HEON_COL_TMA_TILES": 8}), ("n16_I_small", {"HEON_COL_TMA_TILES": 4}),
                                      ("n16_I_small", {"HEON_COL_TMA_TILES": 16, "HEON_COL_TMA_BUFS": 3}),
                                      ("n16_II_small", {"HEON_MODUP_DOUBLES": 0}), ("n13_II", {"HEON_MODUP_DOUBLES": 0}),
                                      ("n15_II", {"HEON_MODUP_DOUBLES": 0, "HEON_ROW_MAC": 0}),
                                      # the Method-II mod-down with its last row pass and final combination in separate kernels
                                      ("n16_II_small", {"HEON_ROW_FINAL": 0}), ("n13_II", {"HEON_ROW_FINAL": 0}),
                                      ("n15_II", {"HEON_ROW_FINAL": 0, "HEON_NTT_FP64": 0}),
                                      ("n16_II_small", {"HEON_ROW_FINAL": 3}), ("n13_II", {"HEON_ROW_FINAL": 8}),
                                      ("n16_I_small", {"HEON_ROW_FINAL": 0}), ("n12_I", {"HEON_ROW_FINAL": 0}),
                                      ("n16_I_small", {"HEON_ROW_FINAL": 3}),
                                      # Method-II mod-up fused with the column pass (N = 2^16) against the separate kernels,
                                      # and its integer conversion / integer butterfly branches
                                      ("n16_II_small", {"HEON_MODUP_COL": 1}), ("n16_II_small", {"HEON_MODUP_COL": 2}),
                                      ("n16_II_small", {"HEON_NTT_FP64": 0}),
                                      ("n16_II_small", {"HEON_MODUP_COL": 2, "HEON_NTT_FP64": 0}),
                                      # the fast mod-up with four coefficients per thread (default: two)
                                      ("n16_II_small", {"HEON_MODUP_CW": 4}), ("n16_II_small", {"HEON_MODUP_CW": 4, "HEON_SKIP_OWN": 0})])
def test_alternate_operator_paths_agree(name, env):
    """multiply + relinearize + rotation through the alternate paths equal the default path."""
    api = _api()
    ctx, alt, oc = gpu_ctx(name), _ctx_with_env(name, **env), oracle_ctx(name)
    batch, L, n = 2, oc.Q, oc.n
    a = ciphertext(131, oc.primes, L, n, 2, batch)
    b = ciphertext(132, oc.primes, L, n, 2, batch)
    key = to_dev(eval_key(133, oc.primes, oc.digits(0), n))
    elt = api.lib.heon_steps_to_galois_elt(3, n, 5)
    res = []
    for c in (ctx, alt):
        op = api.HEArithmeticOperator(c)
        A, B = api.Ciphertext(c, to_dev(a)), api.Ciphertext(c, to_dev(b))
        Cc = api.Ciphertext(c, torch.zeros(batch, 3, L, n, dtype=torch.int64, device="cuda"))
        op.multiply(A, B, Cc)
        op.relinearize_inplace(Cc, api.Relinkey(c, key))
        R_ = api.Ciphertext(c, torch.zeros(batch, 2, L, n, dtype=torch.int64, device="cuda"))
        op.apply_galois(A, R_, api.Galoiskey(c, {elt: key}), elt)
        torch.cuda.synchronize()
        res.append((Cc.data.clone(), R_.data.clone()))
    assert torch.equal(res[0][0], res[1][0]), ("relinearize", env)
    assert torch.equal(res[0][1], res[1][1]), ("apply_galois", env)


@pytest.mark.parametrize("name,walk", [("n13_II", 8), ("n12_I", 8), ("n13_II", 5), ("n16_II_small", 8)])
def test_row_final_long_walks(name, walk):
    """k_row_final with walks of eight polynomials per CTA and a ragged last group (the default at the BASELINE batch
    sizes; the small test batches otherwise walk two): relinearize equals the separate-kernel path and the oracle."""
    api = _api()
    ctx, alt, oc = _ctx_with_env(name, HEON_ROW_FINAL=walk), _ctx_with_env(name, HEON_ROW_FINAL=0), oracle_ctx(name)
    batch, L, n = 5, oc.Q, oc.n
    a = ciphertext(141, oc.primes, L, n, 3, batch)
    key_h = eval_key(143, oc.primes, oc.digits(0), n)
    key = to_dev(key_h)
    res = []
    for c in (ctx, alt):
        op = api.HEArithmeticOperator(c)
        Cc = api.Ciphertext(c, to_dev(a))
        Cc.cipher_size_, Cc.relinearization_required_ = 3, True
        op.relinearize_inplace(Cc, api.Relinkey(c, key))
        torch.cuda.synchronize()
        res.append(Cc.data.clone())
    assert torch.equal(res[0][:, :2], res[1][:, :2]), (name, walk)
    if n <= 8192:
        want = oc.relinearize(a[4], key_h)[:2]
        assert np.array_equal(to_host(res[0])[4, :2], want), "differs from the oracle"

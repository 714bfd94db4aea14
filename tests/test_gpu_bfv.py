"""GPU parity for the BFV path: BEHZ multiply and relinearize (Method I / II)
against the CPU oracle and, bit for bit, against the reference's own kernels."""
import functools

import numpy as np
import pytest
import torch

from oracle import oracle as O, ref as R
from tests.common import residues
from tests.gpu_common import to_dev, to_host
from tests.test_bfv_decrypt_level import BFV_PARAMS, bfv_gpu_ctx, bfv_oracle

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.have_gpu(), reason="oracle/_ref/libref_gpu.so not built")

BFV_PARAMS.update({
    "bfv_n14": (14, [54, 54, 54, 54, 55, 55, 55], [55], 786433),  # test_bfv_multiplication.cpp N=16384 shape
    "bfv_n15_II": (15, [59, 59, 59, 59, 59, 59], [60, 60], 786433),
})


def _inputs(name, batch):
    ob, oc = bfv_oracle(name)
    a = residues(101, ob.primes[: ob.Q], ob.n, (batch, 2))
    b = residues(102, ob.primes[: ob.Q], ob.n, (batch, 2))
    key = residues(103, ob.primes, ob.n, (ob.digits(), 2))
    return ob, oc, a, b, key


@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II", "bfv_n13_I"])
def test_bfv_multiply_relinearize_vs_oracle(name):
    from heongpu_b200 import api
    ob, oc, a, b, key = _inputs(name, 2)
    ctx = bfv_gpu_ctx(name)
    assert ctx.bsk_primes == ob.bsk_primes
    op = api.HEArithmeticOperator(ctx)
    A, B = api.Ciphertext(ctx, to_dev(a)), api.Ciphertext(ctx, to_dev(b))
    Cc = api.Ciphertext(ctx, torch.zeros(2, 3, ob.Q, ob.n, dtype=torch.int64, device="cuda"))
    op.multiply_bfv(A, B, Cc)
    mul = to_host(Cc.data).copy()
    op.relinearize_inplace_bfv(Cc, api.Relinkey(ctx, to_dev(key)))
    rel = to_host(Cc.data).copy()
    for bi in range(2):
        m = ob.multiply(a[bi], b[bi])
        assert np.array_equal(mul[bi], m), "BEHZ multiply"
        r = ob.relinearize(m, key)
        assert np.array_equal(rel[bi][:2], r[:2]), "relinearize"


@functools.lru_cache(maxsize=2)
def _ref_handles(name):
    ob, oc = bfv_oracle(name)
    t = R.tables_for_refgpu(ob.n_power, ob.primes, ob.Q, ob.K, scheme="BFV", plain_modulus=ob.t)
    return R.RefBfv(ob), R.RefGpu(ob.n_power, ob.primes, ob.Q, ob.K, t)


@needs_ref
@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II", "bfv_n14", "bfv_n15_II"])
def test_bfv_bit_exact_vs_reference_kernels(name):
    from heongpu_b200 import api
    ob, oc, a, b, key = _inputs(name, 1)
    ctx = bfv_gpu_ctx(name)
    rb, rg = _ref_handles(name)
    op = api.HEArithmeticOperator(ctx)
    dkey = to_dev(key)
    A, B = api.Ciphertext(ctx, to_dev(a)), api.Ciphertext(ctx, to_dev(b))
    Cc = api.Ciphertext(ctx, torch.zeros(1, 3, ob.Q, ob.n, dtype=torch.int64, device="cuda"))
    op.multiply_bfv(A, B, Cc)
    ours_mul = Cc.data.clone()
    op.relinearize_inplace_bfv(Cc, api.Relinkey(ctx, dkey))
    ra, rb_, rc = to_dev(a[0]), to_dev(b[0]), torch.zeros(3, ob.Q, ob.n, dtype=torch.int64, device="cuda")
    rb.multiply(ra, rb_, rc)
    torch.cuda.synchronize()
    assert torch.equal(ours_mul[0], rc), "BEHZ multiply differs from the reference kernels"
    R.bfv_relinearize(rg, rc, dkey)
    torch.cuda.synchronize()
    assert torch.equal(Cc.data[0, :2], rc[:2]), "BFV relinearize differs from the reference kernels"


@needs_ref
@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II", "bfv_n15_II"])
def test_bfv_rotate_vs_oracle_and_reference_kernels(name):
    from heongpu_b200 import api
    ob, oc, a, b, key = _inputs(name, 2)
    ctx = bfv_gpu_ctx(name)
    rb, rg = _ref_handles(name)
    op = api.HEArithmeticOperator(ctx)
    dkey = to_dev(key)
    for elt in (api.lib.heon_steps_to_galois_elt(1, ob.n, 3), api.lib.heon_steps_to_galois_elt(-2, ob.n, 3), 2 * ob.n - 1):
        gk = api.Galoiskey(ctx, {elt: dkey})
        A = api.Ciphertext(ctx, to_dev(a))
        out = api.Ciphertext(ctx, torch.zeros(2, 2, ob.Q, ob.n, dtype=torch.int64, device="cuda"))
        op.apply_galois_bfv(A, out, gk, elt)
        got = to_host(out.data)
        if ob.n <= 8192:
            for bi in range(2):
                assert np.array_equal(got[bi], ob.apply_galois(a[bi], key, elt))
        ra, ro = to_dev(a[0]), torch.zeros(2, ob.Q, ob.n, dtype=torch.int64, device="cuda")
        R.bfv_apply_galois(rg, ra, ro, dkey, elt)
        torch.cuda.synchronize()
        assert torch.equal(out.data[0], ro), "BFV rotation differs from the reference kernels"

"""CPU: every context table of the product equals, word for word, what the reference's OWN
context classes build -- HEContextImpl<BFV>::generate / HEContextImpl<CKKS>::generate
(src/lib/host/bfv/context.cu, src/lib/host/ckks/context.cu) compiled unmodified into
oracle/_ref/libref_ctx.so (oracle/ref_ctx_harness.cu).  This pins, bit-wise:
  * the BEHZ multiplication tables (bfv/context.cu:500-700, 990-1290),
  * the UN-LEVELLED BFV Method-II tables base_change_matrix_D_to_Qtilda / Mi_inv_D_to_Qtilda /
    prod_D_to_Qtilda / I_j / I_location (contextpool.cpp:160-191, 242-264, 361-394), whose digits
    have size m = 2 whatever |P| is, and whose generator multiplies by an UNREDUCED prime,
  * the CKKS rescale tables built inline in ckks/context.cu:342-368,
on same-size, mixed-size and default-modulus chains."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.have_ctx(), reason="oracle/_ref/libref_ctx.so not built (needs /root/reference)")

COMMON = {"ntt": 2, "intt": 3, "n_inverse": 4, "last_q_modinv": 5, "half": 6, "half_mod": 7, "factor": 8}
II = {"ii_base_change": 12, "ii_mi_inv": 13, "ii_prod": 14, "ii_i_j": 15, "ii_i_location": 16}
BFV = {"bfv_base_change_bsk": 20, "bfv_inv_punct_q": 21, "bfv_base_change_mtilde": 22, "bfv_inv_mtilde_mod_bsk": 23,
       "bfv_prod_q_mod_bsk": 24, "bfv_inv_prod_q_mod_bsk": 25, "bfv_base_change_q": 26, "bfv_base_change_msk": 27,
       "bfv_inv_punct_b": 28, "bfv_prod_b_mod_q": 29, "bfv_scalars": 30}

BFV_CASES = {
    # name: (log_n, q_bits, p_bits, t)
    "bfv_I": (12, [36, 36], [37], 1032193),  # test_bfv_multiplication params (BASELINE config 1 shape)
    "bfv_II_k2": (12, [40, 40, 40, 40], [41, 41], 786433),  # mixed sizes: digit primes below / above the targets
    "bfv_II_k3": (13, [59, 59, 59, 59, 59, 59], [60, 60, 60], 786433),  # |P| = 3: reference digits still have size 2
    "bfv_II_mixed": (12, [50, 30, 30, 45, 30], [60, 60], 65537),
}


def _primes(log_n, qb, pb):
    return O.generate_primes(1 << log_n, qb + pb)


@pytest.mark.parametrize("name", list(BFV_CASES))
def test_bfv_context_tables_equal_reference_context(name):
    from heongpu_b200 import api
    log_n, qb, pb, t = BFV_CASES[name]
    pr = _primes(log_n, qb, pb)
    Q, K = len(qb), len(pb)
    ctx = api.HEContext(log_n, q_values=pr[:Q], p_values=pr[Q:], plain_modulus=t, device=-1)
    rc = R.RefContext("BFV", 1 << log_n, pr[:Q], pr[Q:], plain_modulus=t)
    assert np.array_equal(ctx.table("modulus"), rc.table(0))  # Q' chain followed by the Bsk primes
    for nm, code in {**COMMON, **BFV}.items():
        got, want = ctx.table(nm), rc.table(code)
        if nm in ("ntt", "intt", "n_inverse"):
            want_q = want
            got = got[: len(want_q)]  # ours lists Q' then Bsk; the reference keeps the Q' tables separately
        assert np.array_equal(got, want), nm
    # merged q||Bsk transform tables of the BEHZ multiply
    n = 1 << log_n
    mods = ctx.table("modulus").reshape(-1, 3)
    Qp, bsk = Q + K, len(mods) - (Q + K)
    sel = list(range(Q)) + list(range(Qp, Qp + bsk))
    assert np.array_equal(mods[sel].reshape(-1), rc.table(40))
    assert np.array_equal(ctx.table("ntt").reshape(-1, n)[sel].reshape(-1), rc.table(41))
    assert np.array_equal(ctx.table("intt").reshape(-1, n)[sel].reshape(-1), rc.table(42))
    assert np.array_equal(ctx.table("n_inverse")[sel], rc.table(43))
    # plaintext-operand constants: ours = [coeff_div_plainmod[Q], upper_halfincrement[Q], Q mod t, upper_threshold]
    plain = rc.table(31)
    ours = ctx.table("bfv_plain")
    assert np.array_equal(ours[:Q], plain[2:2 + Q]) and np.array_equal(ours[Q:2 * Q], plain[2 + Q:2 + 2 * Q])
    assert ours[2 * Q] == plain[0] and ours[2 * Q + 1] == plain[1]
    if K > 1:
        m, l, lt, d = (int(v) for v in rc.table(44))
        assert (l, lt) == (Q, Q + K)
        assert ctx.digits(0) == d
        for nm, code in II.items():
            assert np.array_equal(ctx.table(nm, 0), rc.table(code)), nm


def test_bfv_default_modulus_chain_equals_reference():
    """BASELINE config 4: BFV N = 2^15 with the reference's default 128-bit modulus (defaultmodulus.cpp:34-51)."""
    from heongpu_b200 import api
    import bench
    rc = R.RefContext("BFV", 1 << 15, plain_modulus=786433, default_p_size=1)
    mods = rc.table(0).reshape(-1, 3)
    pr = [int(v) for v in mods[:, 0]]
    assert pr[: len(bench.BFV_32768_MODULUS)] == bench.BFV_32768_MODULUS
    Qp = len(bench.BFV_32768_MODULUS)
    ctx = api.HEContext(15, q_values=pr[: Qp - 1], p_values=pr[Qp - 1:Qp], plain_modulus=786433, device=-1)
    assert np.array_equal(ctx.table("modulus"), rc.table(0))
    for nm, code in {**COMMON, **BFV}.items():
        got, want = ctx.table(nm), rc.table(code)
        if nm in ("ntt", "intt", "n_inverse"):
            got = got[: len(want)]
        assert np.array_equal(got, want), nm


@pytest.mark.parametrize("name,log_n,qb,pb", [
    ("n12_I", 12, [40, 30, 30], [40]),
    ("n13_II", 13, [50, 40, 40, 40, 40], [50, 50, 50]),
    ("mixed_II", 12, [40, 40, 40, 40, 40], [41, 41]),  # digit primes above some targets: unreduced-operand Barrett
    ("mixed_I", 12, [60, 30, 30, 30], [60]),
])
def test_ckks_context_tables_equal_reference_context(name, log_n, qb, pb):
    from heongpu_b200 import api
    pr = _primes(log_n, qb, pb)
    Q, K = len(qb), len(pb)
    ctx = api.HEContext(log_n, q_values=pr[:Q], p_values=pr[Q:], device=-1)
    rc = R.RefContext("CKKS", 1 << log_n, pr[:Q], pr[Q:])
    assert np.array_equal(ctx.table("modulus"), rc.table(0))
    names = dict(COMMON)
    names.update({"rescaled_last_q_modinv": 9, "rescaled_half_mod": 10, "rescaled_half": 11})
    for nm, code in names.items():
        assert np.array_equal(ctx.table(nm), rc.table(code)), nm
    if K > 1:
        for depth in range(Q):
            for nm, code in II.items():
                assert np.array_equal(ctx.table(nm, depth), rc.table(code, depth)), (nm, depth)

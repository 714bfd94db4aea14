"""CPU: pin the oracle's (and the product's) table generators and CPU NTT
against the reference's OWN host code (oracle/_ref/libref_host.so, compiled
from /root/reference by `make -C oracle ref`)."""
import numpy as np
import pytest

from oracle import oracle as O, ref as R
from tests.common import PARAMS, residues

pytestmark = pytest.mark.skipif(not R.have_host(), reason="oracle/_ref/libref_host.so not built")

SETS = ["n12_I", "n12_II", "n13_II", "n14_C2", "mixed"]


@pytest.mark.parametrize("name", SETS + ["C3_I", "C3_II"])
def test_primes_and_modulus_records(name):
    log_n, qb, pb = PARAMS[name]
    a = O.generate_primes(1 << log_n, qb + pb)
    b = R.generate_primes(1 << log_n, qb + pb)
    assert a == b
    for p in set(a):
        assert O.make_mod(p) == R.modulus(p)


@pytest.mark.parametrize("name", SETS)
def test_tables(name):
    log_n, qb, pb = PARAMS[name]
    Q, K = len(qb), len(pb)
    pr = O.generate_primes(1 << log_n, qb + pb)
    for x, y in zip(O.ntt_tables(pr, log_n), R.ntt_tables(pr, log_n)):
        assert np.array_equal(x, y)
    for x, y in zip(O.moddown_tables(pr, Q, K), R.moddown_tables(pr, Q, K)):
        assert np.array_equal(x, y)
    if K > 1:
        for depth in range(Q):
            mo, mr = O.method2_tables(pr, Q, K, depth), R.method2_tables(1 << log_n, pr, Q, K, depth)
            assert mo["d"] == mr["d"]
            for k in ("base_change", "mi_inv", "prod", "I_j", "I_location"):
                assert np.array_equal(mo[k], mr[k]), (k, depth)


def test_psi_headline_sets():
    for name in ("C3_I", "C3_II"):
        log_n, qb, pb = PARAMS[name]
        pr = O.generate_primes(1 << log_n, qb + pb)
        sub = [pr[0], pr[1], pr[-1]]
        po, fo, io, no = O.ntt_tables(sub, log_n)
        pr_, fr, ir, nr = R.ntt_tables(sub, log_n)
        assert np.array_equal(po, pr_) and np.array_equal(fo, fr) and np.array_equal(io, ir) and np.array_equal(no, nr)


@pytest.mark.parametrize("name", ["n12_I", "n13_II", "mixed"])
def test_cpu_ntt_matches_nttcpu(name):
    from tests.common import oracle_ctx
    oc = oracle_ctx(name)
    psi, *_ = O.ntt_tables(oc.primes, oc.n_power)
    for i in (0, oc.Qp - 1):
        a = residues(7 + i, [oc.primes[i]], oc.n)[0]
        f = oc.ntt(a, [i])
        g = R.ntt_cpu(a, oc.n_power, oc.primes[i], int(psi[i]))
        assert np.array_equal(f, g)
        assert np.array_equal(oc.ntt(f, [i], inverse=True), a)
        assert np.array_equal(R.ntt_cpu(g, oc.n_power, oc.primes[i], int(psi[i]), inverse=True), a)


def test_barrett_mult_matches_reference():
    rng = np.random.default_rng(5)
    for p in O.generate_primes(4096, [30, 45, 60]) + [2305843009213554689]:
        for _ in range(200):
            a, b = int(rng.integers(0, p)), int(rng.integers(0, p))
            assert O.lib().oracle_mult(a, b, p) == R.host().ref_mult(a, b, p) == (a * b) % p

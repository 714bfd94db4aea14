"""CPU: the C-ABI library loads, exports every symbol include/heon_b200.h
declares, and its host-side tables equal the oracle's (no compute calls)."""
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from tests.common import PARAMS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from heongpu_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "heon_b200.h")).read()
    declared = set(re.findall(r"\b(heon_[a-z0-9_]+)\s*\(", header))
    declared -= {"heon_context_s"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.heon_version()


@pytest.mark.parametrize("name", ["n12_I", "n12_II", "n13_II", "n14_C2", "mixed"])
def test_host_tables_equal_oracle(name):
    from heongpu_b200 import api
    log_n, qb, pb = PARAMS[name]
    Q, K = len(qb), len(pb)
    ctx = api.HEContext(log_n, qb, pb, device=-1)
    pr = O.generate_primes(1 << log_n, qb + pb)
    assert ctx.primes == pr
    mods = ctx.table("modulus").reshape(-1, 3)
    for i, p in enumerate(pr):
        assert tuple(int(v) for v in mods[i]) == O.make_mod(p)
    psi, fwd, inv, ninv = O.ntt_tables(pr, log_n)
    assert np.array_equal(psi, ctx.table("psi"))
    assert np.array_equal(fwd, ctx.table("ntt"))
    assert np.array_equal(inv, ctx.table("intt"))
    assert np.array_equal(ninv, ctx.table("n_inverse"))
    for a, nm in zip(O.moddown_tables(pr, Q, K), ["last_q_modinv", "half", "half_mod", "factor"]):
        assert np.array_equal(a, ctx.table(nm)), nm
    for a, nm in zip(O.rescale_tables(pr, Q), ["rescaled_last_q_modinv", "rescaled_half_mod", "rescaled_half"]):
        assert np.array_equal(a, ctx.table(nm)), nm
    if K > 1:
        for depth in range(Q):
            m = O.method2_tables(pr, Q, K, depth)
            for k, nm in [("base_change", "ii_base_change"), ("mi_inv", "ii_mi_inv"), ("prod", "ii_prod"),
                          ("I_j", "ii_i_j"), ("I_location", "ii_i_location")]:
                assert np.array_equal(m[k].astype(np.uint64), ctx.table(nm, depth)), (k, depth)


def test_explicit_prime_values_and_errors():
    from heongpu_b200 import api
    pr = O.generate_primes(4096, [40, 30, 30, 40])
    ctx = api.HEContext(12, q_values=pr[:3], p_values=pr[3:], device=-1)
    assert ctx.primes == pr and ctx.keyswitch_method == 1
    with pytest.raises(api.HeonError):
        api.HEContext(12, q_values=[pr[0], 1000003], p_values=[pr[3]], device=-1)  # not 1 mod 2N
    with pytest.raises(api.HeonError):
        api.HEContext(11, [40, 30], [40], device=-1)  # ring too small (MIN_POLY_DEGREE 4096)
    with pytest.raises(api.HeonError):
        api.HEContext(12, [40, 30], [], device=-1)  # P cannot be empty
    with pytest.raises(api.HeonError):
        api.HEContext(12, [61, 30], [40], device=-1)  # invalid modulus bit size


def test_host_only_context_refuses_compute():
    import ctypes as C
    from heongpu_b200 import api
    ctx = api.HEContext(12, [40, 30], [40], device=-1)
    rc = api.lib.heon_ckks_rescale(ctx._h, C.c_void_p(8), 0, 0, 1, None)
    assert rc != 0 and b"host-only" in api.lib.heon_last_error()


def test_steps_to_galois_elt():
    from heongpu_b200 import api
    n = 4096
    assert api.lib.heon_steps_to_galois_elt(0, n, 5) == 2 * n - 1
    assert api.lib.heon_steps_to_galois_elt(1, n, 5) == 5
    assert api.lib.heon_steps_to_galois_elt(3, n, 5) == pow(5, 3, 2 * n)
    assert api.lib.heon_steps_to_galois_elt(-1, n, 5) == pow(5, n // 2 - 1, 2 * n)
    assert api.lib.heon_steps_to_galois_elt(n // 2, n, 5) == 0


@pytest.mark.parametrize("log_n,qb,pb,t", [(12, [36, 36], [37], 1032193), (13, [40, 40, 40, 40], [45, 45], 786433)])
def test_bfv_plain_constants_equal_big_integer_arithmetic(log_n, qb, pb, t):
    """coeeff_div_plainmod_ = floor(Q/t) mod q_i (the reference divides the big integer with GMP,
    bfv/context.cu:953-984); the engine uses -(Q mod t) * t^-1 mod q_i.  Checked against Python integers."""
    from heongpu_b200 import api
    ctx = api.HEContext(log_n, qb, pb, device=-1, plain_modulus=t)
    Q = len(qb)
    tab = [int(v) for v in ctx.table("bfv_plain")]
    primes = ctx.primes[:Q]
    bigQ = 1
    for q in primes:
        bigQ *= q
    assert tab[:Q] == [(bigQ // t) % q for q in primes]
    assert tab[Q:2 * Q] == [q - t for q in primes]
    assert tab[2 * Q] == bigQ % t and tab[2 * Q + 1] == (t + 1) >> 1


def test_bench_galois_elt_matches_the_abi():
    """bench.py's reference arm computes Galois elements itself (so that arm does not touch this
    library): same values as heon_steps_to_galois_elt (keygeneration.cu:684-727)."""
    import importlib.util
    from heongpu_b200 import _lib
    lib = _lib.load()
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for n in (4096, 32768, 65536):
        for order in (3, 5):
            for steps in bench.ROT_STEPS + [3, -5, 100]:
                assert bench.galois_elt(steps, n, order) == lib.heon_steps_to_galois_elt(steps, n, order), (n, order, steps)

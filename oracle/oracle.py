"""ctypes/numpy wrapper of oracle/heon_oracle.c (the CPU restatement of the
reference hot path).  TEST INFRASTRUCTURE ONLY -- see the header of
heon_oracle.c for what each function restates and how the oracle is pinned."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libheon_oracle.so")
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int)


def build(force=False):
    src = os.path.join(_HERE, "heon_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_mult.restype = C.c_uint64
        L.oracle_mult.argtypes = [C.c_uint64] * 3
        L.oracle_minimal_root.restype = C.c_uint64
        L.oracle_minimal_root.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_ctx_create.restype = C.c_void_p
        L.oracle_ctx_create.argtypes = [C.c_int, u64p, C.c_int, C.c_int]
        L.oracle_ctx_destroy.argtypes = [C.c_void_p]
        L.oracle_digits.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(i32p)


def make_mod(p):
    out = (C.c_uint64 * 3)()
    lib().oracle_make_mod(C.c_uint64(p), out)
    return tuple(int(v) for v in out)


def generate_primes(n, bits):
    b = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    rc = lib().oracle_generate_primes(C.c_uint64(n), b, len(bits), _p(out))
    if rc:
        raise RuntimeError("failed to find enough qualifying primes")
    return [int(v) for v in out]


def ntt_tables(primes, n_power):
    n = 1 << n_power
    pr = np.array(primes, dtype=np.uint64)
    psi = np.zeros(len(primes), dtype=np.uint64)
    fwd = np.zeros(len(primes) * n, dtype=np.uint64)
    inv = np.zeros(len(primes) * n, dtype=np.uint64)
    ninv = np.zeros(len(primes), dtype=np.uint64)
    lib().oracle_ntt_tables(_p(pr), len(primes), n_power, _p(psi), _p(fwd), _p(inv), _p(ninv))
    return psi, fwd, inv, ninv


def moddown_tables(primes, Q, K):
    Qp = Q + K
    pr = np.array(primes, dtype=np.uint64)
    lqm = np.zeros(K * Qp, dtype=np.uint64)
    half = np.zeros(K, dtype=np.uint64)
    hm = np.zeros(K * Qp, dtype=np.uint64)
    fac = np.zeros(K * Q, dtype=np.uint64)
    w = lib().oracle_moddown_tables(_p(pr), Qp, K, Q, _p(lqm), _p(half), _p(hm), _p(fac))
    return lqm[:w].copy(), half, hm[:w].copy(), fac


def rescale_tables(primes, Q):
    pr = np.array(primes, dtype=np.uint64)
    a = np.zeros(Q * Q, dtype=np.uint64)
    b = np.zeros(Q * Q, dtype=np.uint64)
    h = np.zeros(max(Q - 1, 1), dtype=np.uint64)
    w = lib().oracle_rescale_tables(_p(pr), Q, _p(a), _p(b), _p(h))
    return a[:w].copy(), b[:w].copy(), h[: Q - 1].copy()


def method2_tables(primes, Q, K, depth, digit_m=None):
    """digit_m: digit size (default |P|, the CKKS rule; the reference's BFV context always uses 2)."""
    Qp = Q + K
    m = K if digit_m is None else digit_m
    pr = np.array(primes, dtype=np.uint64)
    bc = np.zeros(Q * Qp * (m + 1), dtype=np.uint64)
    mi = np.zeros(Q, dtype=np.uint64)
    prod = np.zeros(Q * Qp, dtype=np.uint64)
    ij = np.zeros(Q, dtype=np.int32)
    il = np.zeros(Q, dtype=np.int32)
    cnt = np.zeros(3, dtype=np.int32)
    d = lib().oracle_method2_tables_m(_p(pr), Qp, K, depth, m, _p(bc), _p(mi), _p(prod), _ip(ij), _ip(il), _ip(cnt))
    return dict(d=d, base_change=bc[: cnt[0]].copy(), mi_inv=mi[: cnt[1]].copy(), prod=prod[: cnt[2]].copy(),
                I_j=ij[:d].copy(), I_location=il[:d].copy())


class OracleContext:
    """Bundle of tables + operator-level restatements (one ciphertext at a time)."""

    def __init__(self, n_power, primes, Q, K):
        self.n_power, self.n = n_power, 1 << n_power
        self.Q, self.K, self.Qp = Q, K, Q + K
        self.primes = [int(p) for p in primes]
        self.method = 1 if K == 1 else 2
        pr = np.array(self.primes, dtype=np.uint64)
        self._h = C.c_void_p(lib().oracle_ctx_create(n_power, _p(pr), Q, K))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_ctx_destroy(self._h)
            self._h = None

    def level_primes(self, depth=0):
        L = self.Q - depth
        return list(range(L)) + [self.Q + j for j in range(self.K)]

    def digits(self, depth=0):
        return lib().oracle_digits(self._h, depth)

    def ntt(self, data, order, inverse=False):
        """data [..., N] uint64 (copied); poly z uses prime order[z % len(order)]."""
        a = np.ascontiguousarray(data, dtype=np.uint64).copy()
        o = np.array(order, dtype=np.int32)
        lib().oracle_ntt_batch(self._h, _p(a), C.c_longlong(a.size // self.n), _ip(o), len(o), int(inverse))
        return a

    def multiply(self, a, b, depth=0):
        L = self.Q - depth
        out = np.zeros((3, L, self.n), dtype=np.uint64)
        lib().oracle_cross_multiply(self._h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out), depth)
        return out

    def addsub(self, a, b, op, depth=0):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        out = np.zeros_like(a)
        lib().oracle_addsub(self._h, _p(a), _p(b), _p(out), a.shape[0], depth, op)
        return out

    def relinearize(self, ct3, key, depth=0):
        ct = np.ascontiguousarray(ct3, dtype=np.uint64).copy()
        lib().oracle_relinearize(self._h, _p(ct), _p(np.ascontiguousarray(key)), depth)
        return ct  # [3][L][N]: c0', c1', INTT(c2)

    def rescale(self, ct2, depth=0):
        L = self.Q - depth
        ct = np.ascontiguousarray(ct2, dtype=np.uint64).copy()
        lib().oracle_rescale(self._h, _p(ct), depth)
        return ct.reshape(-1)[: 2 * (L - 1) * self.n].reshape(2, L - 1, self.n).copy()

    def apply_galois(self, ct2, key, galois_elt, depth=0):
        ct = np.ascontiguousarray(ct2, dtype=np.uint64)
        out = np.zeros_like(ct)
        lib().oracle_apply_galois(self._h, _p(ct), _p(out), _p(np.ascontiguousarray(key)), int(galois_elt), depth)
        return out

    def keyswitch(self, ct2, key, depth=0):
        ct = np.ascontiguousarray(ct2, dtype=np.uint64)
        out = np.zeros_like(ct)
        lib().oracle_keyswitch(self._h, _p(ct), _p(out), _p(np.ascontiguousarray(key)), depth)
        return out

    def conjugate(self, ct2, key, depth=0):
        """conjugate_ckks_method_I/II (ckks/operator.cu:2027-2311): apply_galois with galois_elt_zero = 2N-1."""
        return self.apply_galois(ct2, key, 2 * self.n - 1, depth)

    def plain(self, ct, pt, op, depth=0):
        """op 0 multiply_plain, 1 add_plain, 2 sub_plain; ct [comps][L][N], pt [L][N]."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        pt = np.ascontiguousarray(pt, dtype=np.uint64)
        out = np.zeros_like(ct)
        lib().oracle_plain(self._h, _p(ct), _p(pt), _p(out), ct.shape[0], depth, op)
        return out

    def mod_drop(self, ct, depth=0):
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        comps, L = ct.shape[0], self.Q - depth
        out = np.zeros((comps, L - 1, self.n), dtype=np.uint64)
        lib().oracle_mod_drop(self._h, _p(ct), _p(out), comps, depth)
        return out

    def modup(self, coef, depth=0):
        L = self.Q - depth
        d = self.digits(depth)
        out = np.zeros((d, L + self.K, self.n), dtype=np.uint64)
        lib().oracle_modup(self._h, _p(np.ascontiguousarray(coef)), _p(out), depth)
        return out

    def keyswitch_core(self, coef, key, depth=0):
        L = self.Q - depth
        acc = np.zeros((2, L + self.K, self.n), dtype=np.uint64)
        lib().oracle_keyswitch_core(self._h, _p(np.ascontiguousarray(coef)), _p(np.ascontiguousarray(key)), _p(acc), depth)
        return acc


class BfvOracle:
    """BFV restatement: BEHZ multiply and un-levelled relinearize (heon_oracle.c, BFV section)."""

    def __init__(self, n_power, primes, Q, K, plain_modulus):
        L = lib()
        L.oracle_bfv_create.restype = C.c_void_p
        L.oracle_bfv_create.argtypes = [C.c_int, u64p, C.c_int, C.c_int, C.c_uint64]
        L.oracle_bfv_destroy.argtypes = [C.c_void_p]
        L.oracle_bfv_bsk_count.argtypes = [C.c_void_p]
        self.n_power, self.n, self.Q, self.K, self.Qp = n_power, 1 << n_power, Q, K, Q + K
        self.primes = [int(p) for p in primes]
        self.t = int(plain_modulus)
        self.method = 1 if K == 1 else 2
        pr = np.array(self.primes, dtype=np.uint64)
        self._h = C.c_void_p(L.oracle_bfv_create(n_power, _p(pr), Q, K, C.c_uint64(self.t)))
        self.bsk = L.oracle_bfv_bsk_count(self._h)
        b = np.zeros(self.bsk, dtype=np.uint64)
        L.oracle_bfv_bsk_primes(self._h, _p(b))
        self.bsk_primes = [int(v) for v in b]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_bfv_destroy(self._h)
            self._h = None

    def digits(self):
        # Method II digits have size 2 in the reference's BFV context whatever |P| is (contextpool.hpp:29)
        return self.Q if self.K == 1 else -(-self.Q // 2)

    def table(self, which):
        out = np.zeros(64 * 64, dtype=np.uint64)
        n = lib().oracle_bfv_table(self._h, which, _p(out))
        return out[:n].copy()

    def multiply(self, a, b):
        out = np.zeros((3, self.Q, self.n), dtype=np.uint64)
        lib().oracle_bfv_multiply(self._h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
        return out

    def relinearize(self, ct3, key):
        ct = np.ascontiguousarray(ct3, dtype=np.uint64).copy()
        lib().oracle_bfv_relinearize(self._h, _p(ct), _p(np.ascontiguousarray(key)))
        return ct

    def plain(self, ct, pt, op):
        """op 0 multiply_plain_bfv, 1 add_plain_bfv, 2 sub_plain_bfv; ct [comps][Q][N] (coefficient domain), pt [N] < t."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        pt = np.ascontiguousarray(pt, dtype=np.uint64)
        out = np.zeros_like(ct)
        lib().oracle_bfv_plain(self._h, _p(ct), _p(pt), _p(out), ct.shape[0], op)
        return out

    def apply_galois(self, ct2, key, galois_elt):
        ct = np.ascontiguousarray(ct2, dtype=np.uint64)
        out = np.zeros_like(ct)
        lib().oracle_bfv_apply_galois(self._h, _p(ct), _p(out), _p(np.ascontiguousarray(key)), int(galois_elt))
        return out


# ---------------------------------------------------------------------------------------------
# TFHE gate bootstrapping (restatement of small_ntt.cu + bootstrapping.cu; heon_oracle.c)
# ---------------------------------------------------------------------------------------------
class TfheOracle:
    P = 1152921504606877697
    n, N, k, l, bg_bit, ks_base_bit, ks_length = 512, 1024, 1, 2, 10, 2, 8

    def __init__(self):
        self.L = lib()
        self.fwd, self.inv = np.zeros(1024, dtype=np.uint64), np.zeros(1024, dtype=np.uint64)
        self.L.oracle_tfhe_table(_p(self.fwd), 0)
        self.L.oracle_tfhe_table(_p(self.inv), 1)

    def ntt(self, polys, inverse=False):
        out = np.ascontiguousarray(polys, dtype=np.uint64).copy()
        flat = out.reshape(-1, 1024)
        for row in flat:
            self.L.oracle_tfhe_ntt(_p(row), _p(self.inv if inverse else self.fwd), int(inverse))
        return out

    def gate_linear(self, gate, a1, b1, a2=None, b2=None):
        a1, b1 = np.ascontiguousarray(a1, dtype=np.int32), np.ascontiguousarray(b1, dtype=np.int32)
        a2 = np.ascontiguousarray(a2 if a2 is not None else a1, dtype=np.int32)
        b2 = np.ascontiguousarray(b2 if b2 is not None else b1, dtype=np.int32)
        oa, ob = np.zeros_like(a1), np.zeros_like(b1)
        rc = self.L.oracle_tfhe_gate_linear(gate, _ip(a1), _ip(b1), _ip(a2), _ip(b2), _ip(oa), _ip(ob), a1.shape[1], a1.shape[0])
        assert rc == 0
        return oa, ob

    def bootstrap(self, a, b, bk):
        a, b = np.ascontiguousarray(a, dtype=np.int32), np.ascontiguousarray(b, dtype=np.int32)
        bk = np.ascontiguousarray(bk, dtype=np.uint64)
        oa, ob = np.zeros((a.shape[0], 1024), dtype=np.int32), np.zeros(a.shape[0], dtype=np.int32)
        self.L.oracle_tfhe_bootstrap(_ip(a), _ip(b), _ip(oa), _ip(ob), _p(bk.reshape(-1)), a.shape[1], a.shape[0])
        return oa, ob

    def keyswitch(self, a, b, ks_a, ks_b):
        a, b = np.ascontiguousarray(a, dtype=np.int32), np.ascontiguousarray(b, dtype=np.int32)
        ks_a, ks_b = np.ascontiguousarray(ks_a, dtype=np.int32), np.ascontiguousarray(ks_b, dtype=np.int32)
        oa, ob = np.zeros((a.shape[0], self.n), dtype=np.int32), np.zeros(a.shape[0], dtype=np.int32)
        self.L.oracle_tfhe_keyswitch(_ip(a), _ip(b), _ip(oa), _ip(ob), _ip(ks_a.reshape(-1)), _ip(ks_b.reshape(-1)),
                                     self.ks_base_bit, self.ks_length, self.n, self.k * self.N, a.shape[0])
        return oa, ob

    # plain host keys / encryption for CPU-side checks (not the product's generator)
    def keygen(self, rng, boot_steps=None):
        n = boot_steps or self.n
        lwe = rng.integers(0, 2, n).astype(np.int32)
        tlwe = rng.integers(0, 2, self.N).astype(np.int32)
        return lwe, tlwe

    def phase(self, a, b, lwe):
        return (b.astype(np.int64) - (a.astype(np.int64) * lwe.astype(np.int64)).sum(axis=1)).astype(np.int64) & 0xFFFFFFFF

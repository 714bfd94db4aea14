"""BFV decrypt-level tests (shape of test/test_bfv_multiplication.cpp and
test_bfv_relinearization.cpp): encrypt two plaintext polynomials, BEHZ-multiply,
relinearize (Method I and II), decrypt, compare with the negacyclic product
modulo the plain modulus.  Runs on the oracle (CPU) and on the CUDA path (GPU)."""
import functools

import numpy as np
import pytest

from oracle import oracle as O
from tests.common import splitmix64

BFV_PARAMS = {
    "bfv_n12_I": (12, [40, 40], [40], 1032193),  # test_bfv_multiplication.cpp:12-19
    "bfv_n12_II": (12, [40, 40], [40, 40], 1032193),  # test_bfv_relinearization.cpp:425-517
    "bfv_n13_I": (13, [54, 54, 54], [55], 786433),
}


@functools.lru_cache(maxsize=None)
def bfv_oracle(name):
    log_n, qb, pb, t = BFV_PARAMS[name]
    primes = O.generate_primes(1 << log_n, qb + pb)
    return O.BfvOracle(log_n, primes, len(qb), len(pb), t), O.OracleContext(log_n, primes, len(qb), len(pb))


def _small(seed, n, bound):
    return (splitmix64(seed, n).astype(np.int64) % (2 * bound + 1)) - bound


def _rns(poly, primes):
    return np.stack([(poly % p).astype(np.uint64) for p in primes])


def _mul(a, b, p):
    return ((a.astype(object) * b.astype(object)) % p).astype(np.uint64)


class Bfv:
    def __init__(self, name):
        self.ob, self.oc = bfv_oracle(name)
        ob = self.ob
        self.n, self.Q, self.K, self.Qp, self.t = ob.n, ob.Q, ob.K, ob.Qp, ob.t
        self.primes = ob.primes
        self.s = _small(11, self.n, 1)
        self.s_ntt = self.oc.ntt(_rns(self.s, self.primes), list(range(self.Qp)))
        self.Qprod = 1
        for p in self.primes[: self.Q]:
            self.Qprod *= p
        self.Pprod = 1
        for p in self.primes[self.Q:]:
            self.Pprod *= p

    def polymul_q(self, a_coef, b_ntt, L):
        """a (coefficient RNS) * b (NTT RNS) -> coefficient RNS over the first L primes"""
        a_ntt = self.oc.ntt(a_coef, list(range(L)))
        prod = np.stack([_mul(a_ntt[i], b_ntt[i], self.primes[i]) for i in range(L)])
        return self.oc.ntt(prod, list(range(L)), inverse=True)

    def encrypt(self, m, seed):
        Q = self.Q
        pl = self.primes[:Q]
        delta = self.Qprod // self.t
        a = np.stack([splitmix64(seed + i, self.n) % np.uint64(p) for i, p in enumerate(pl)])
        e = _small(seed + 50, self.n, 4)
        as_ = self.polymul_q(a, self.s_ntt, Q)
        c0 = np.zeros((Q, self.n), dtype=np.uint64)
        for i, p in enumerate(pl):
            dm = (np.array([(int(v) * (delta % p)) % p for v in m], dtype=object) + (e % p).astype(object)) % p
            c0[i] = ((dm - as_[i].astype(object)) % p).astype(np.uint64)
        return np.stack([c0, a])  # coefficient domain

    def relin_key(self, seed):
        d = self.ob.digits()
        key = np.zeros((d, 2, self.Qp, self.n), dtype=np.uint64)
        s2 = np.stack([_mul(self.s_ntt[y], self.s_ntt[y], p) for y, p in enumerate(self.primes)])
        allp = list(range(self.Qp))
        for i in range(d):
            a = np.stack([splitmix64(seed + 100 * i + y, self.n) % np.uint64(p) for y, p in enumerate(self.primes)])
            e = self.oc.ntt(_rns(_small(seed + 100 * i + 55, self.n, 4), self.primes), allp)
            for y, p in enumerate(self.primes):
                v = (-(a[y].astype(object) * self.s_ntt[y].astype(object) + e[y].astype(object))) % p
                in_digit = y < self.Q and ((y == i) if self.K == 1 else (y // self.K == i))
                if in_digit:
                    v = (v + (self.Pprod % p) * s2[y].astype(object)) % p
                key[i, 0, y] = v.astype(np.uint64)
                key[i, 1, y] = a[y]
        return key

    def decrypt(self, ct, count):
        Q = self.Q
        pl = self.primes[:Q]
        x = ct[0].astype(object)
        cs = self.polymul_q(ct[1], self.s_ntt, Q)
        if ct.shape[0] == 3:
            s2 = np.stack([_mul(self.s_ntt[y], self.s_ntt[y], p) for y, p in enumerate(pl)])
            cs2 = self.polymul_q(ct[2], s2, Q)
        out = []
        for j in range(count):
            v = 0
            for i, p in enumerate(pl):
                r = (int(x[i, j]) + int(cs[i, j]) + (int(cs2[i, j]) if ct.shape[0] == 3 else 0)) % p
                Mi = self.Qprod // p
                v += r * Mi * pow(Mi, -1, p)
            v %= self.Qprod
            out.append(((v * self.t + self.Qprod // 2) // self.Qprod) % self.t)
        return out


def _negacyclic_mod_t(a, b, n, t, count):
    out = []
    a = [int(v) for v in a]
    b = [int(v) for v in b]
    for k in range(count):
        acc = 0
        for i in range(n):
            j = k - i
            acc += a[i] * b[j] if j >= 0 else -a[i] * b[j + n]
        out.append(acc % t)
    return out


def _run(name, backend):
    sc = Bfv(name)
    n, t = sc.n, sc.t
    m1 = (splitmix64(1, n) % np.uint64(t)).astype(np.int64)
    m2 = (splitmix64(2, n) % np.uint64(t)).astype(np.int64)
    ct1, ct2 = sc.encrypt(m1, 1000), sc.encrypt(m2, 2000)
    count = 16
    assert sc.decrypt(ct1, count) == [int(v) for v in m1[:count]]
    want = _negacyclic_mod_t(m1, m2, n, t, count)
    rk = sc.relin_key(3000)
    mul, rel = backend(sc, ct1, ct2, rk)
    assert sc.decrypt(mul, count) == want, "BEHZ multiply does not decrypt to the plaintext product"
    assert sc.decrypt(rel[:2], count) == want, "relinearized ciphertext does not decrypt to the plaintext product"


def _oracle_backend(sc, ct1, ct2, rk):
    mul = sc.ob.multiply(ct1, ct2)
    return mul, sc.ob.relinearize(mul, rk)


@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II"])
def test_bfv_decrypt_level_oracle(name):
    _run(name, _oracle_backend)


@functools.lru_cache(maxsize=None)
def bfv_gpu_ctx(name):
    from heongpu_b200 import api
    log_n, qb, pb, t = BFV_PARAMS[name]
    return api.HEContext(log_n, qb, pb, device=0, plain_modulus=t)


def _gpu_backend_for(name):
    def backend(sc, ct1, ct2, rk):
        import torch
        from heongpu_b200 import api
        from tests.gpu_common import to_dev, to_host
        ctx = bfv_gpu_ctx(name)
        op = api.HEArithmeticOperator(ctx)
        A, B = api.Ciphertext(ctx, to_dev(ct1)), api.Ciphertext(ctx, to_dev(ct2))
        Cc = api.Ciphertext(ctx, torch.zeros(1, 3, sc.Q, sc.n, dtype=torch.int64, device="cuda"))
        op.multiply_bfv(A, B, Cc)
        mul = to_host(Cc.data)[0].copy()
        op.relinearize_inplace_bfv(Cc, api.Relinkey(ctx, to_dev(rk)))
        return mul, to_host(Cc.data)[0].copy()
    return backend


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II", "bfv_n13_I"])
def test_bfv_decrypt_level_gpu(name):
    _run(name, _gpu_backend_for(name))


# ---- BFV keyswitch (switchkey_method_I/II, bfv/operator.cu:975-1372): re-encryption under a new secret ----
def _switch_key(sc, s_new_ntt, seed):
    """key that takes a ciphertext part under sc.s to the secret s_new: rk0 = -(a*s_new + e) + [own limb] P*s."""
    d = sc.ob.digits()
    key = np.zeros((d, 2, sc.Qp, sc.n), dtype=np.uint64)
    allp = list(range(sc.Qp))
    for i in range(d):
        a = np.stack([splitmix64(seed + 100 * i + y, sc.n) % np.uint64(p) for y, p in enumerate(sc.primes)])
        e = sc.oc.ntt(_rns(_small(seed + 100 * i + 55, sc.n, 4), sc.primes), allp)
        for y, p in enumerate(sc.primes):
            v = (-(a[y].astype(object) * s_new_ntt[y].astype(object) + e[y].astype(object))) % p
            in_digit = y < sc.Q and ((y == i) if sc.K == 1 else (y // sc.K == i))
            if in_digit:
                v = (v + (sc.Pprod % p) * sc.s_ntt[y].astype(object)) % p
            key[i, 0, y] = v.astype(np.uint64)
            key[i, 1, y] = a[y]
    return key


def _run_keyswitch(name, backend):
    sc = Bfv(name)
    n, t = sc.n, sc.t
    m = (splitmix64(7, n) % np.uint64(t)).astype(np.int64)
    ct = sc.encrypt(m, 5000)
    s_new = _small(4242, n, 1)
    s_new_ntt = sc.oc.ntt(_rns(s_new, sc.primes), list(range(sc.Qp)))
    out = backend(sc, ct, _switch_key(sc, s_new_ntt, 6000))
    sc.s_ntt = s_new_ntt  # decrypt under the NEW secret
    assert sc.decrypt(out, 32) == [int(v) for v in m[:32]], "key-switched ciphertext does not decrypt under the new secret"


@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II"])
def test_bfv_keyswitch_decrypts_oracle(name):
    _run_keyswitch(name, lambda sc, ct, key: sc.ob.apply_galois(ct, key, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II", "bfv_n13_I"])
def test_bfv_keyswitch_decrypts_gpu(name):
    def backend(sc, ct, key):
        import torch
        from heongpu_b200 import api
        from tests.gpu_common import to_dev, to_host
        ctx = bfv_gpu_ctx(name)
        op = api.HEArithmeticOperator(ctx)
        A = api.Ciphertext(ctx, to_dev(ct))
        out = api.Ciphertext(ctx, torch.zeros(1, 2, sc.Q, sc.n, dtype=torch.int64, device="cuda"))
        op.keyswitch_bfv(A, out, api.Switchkey(ctx, to_dev(key)))
        got = to_host(out.data)[0].copy()
        assert np.array_equal(got, sc.ob.apply_galois(ct, key, 1)), "BFV keyswitch differs from the oracle"
        return got
    _run_keyswitch(name, backend)


# ---- BFV plaintext operands (add_plain_bfv / sub_plain_bfv / multiply_plain_bfv) ----
def _run_plain(name, backend):
    sc = Bfv(name)
    n, t = sc.n, sc.t
    m1 = (splitmix64(21, n) % np.uint64(t)).astype(np.int64)
    m2 = (splitmix64(22, n) % np.uint64(t)).astype(np.int64)
    ct = sc.encrypt(m1, 7000)
    pt = m2.astype(np.uint64)
    add, sub, mul = backend(sc, ct, pt)
    count = 24
    assert sc.decrypt(add, count) == [int((a + b) % t) for a, b in zip(m1[:count], m2[:count])], "add_plain"
    assert sc.decrypt(sub, count) == [int((a - b) % t) for a, b in zip(m1[:count], m2[:count])], "sub_plain"
    assert sc.decrypt(mul, count) == _negacyclic_mod_t(m1, m2, n, t, count), "multiply_plain"


@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II"])
def test_bfv_plain_ops_decrypt_oracle(name):
    _run_plain(name, lambda sc, ct, pt: (sc.ob.plain(ct, pt, 1), sc.ob.plain(ct, pt, 2), sc.ob.plain(ct, pt, 0)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bfv_n12_I", "bfv_n12_II", "bfv_n13_I"])
def test_bfv_plain_ops_gpu(name):
    def backend(sc, ct, pt):
        import torch
        from heongpu_b200 import api
        from tests.gpu_common import to_dev, to_host
        ctx = bfv_gpu_ctx(name)
        op = api.HEArithmeticOperator(ctx)
        batch = 2
        cts = np.stack([ct, ct])
        A = api.Ciphertext(ctx, to_dev(cts))
        P = api.Plaintext(ctx, to_dev(pt))
        res = []
        for fn, code in ((op.add_plain_bfv, 1), (op.sub_plain_bfv, 2), (op.multiply_plain_bfv, 0)):
            out = api.Ciphertext(ctx, torch.zeros(batch, 2, sc.Q, sc.n, dtype=torch.int64, device="cuda"))
            fn(A, P, out)
            got = to_host(out.data).copy()
            want = sc.ob.plain(ct, pt, code)
            for bi in range(batch):
                assert np.array_equal(got[bi], want), f"BFV plain op {code} differs from the oracle"
            res.append(got[0])
        assert np.array_equal(to_host(A.data), cts), "input must stay untouched"
        return res
    _run_plain(name, backend)

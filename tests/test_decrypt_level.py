"""Decrypt-level tests (the shape of the reference's own tests,
test/test_ckks_relinearization.cpp:36-830, test_ckks_rotation_method_*.cpp):
encrypt -> multiply -> relinearize (-> rescale) / rotate -> decrypt -> compare
with the plaintext computation.  Guards against "oracle and CUDA path wrong
together": keys here are real RLWE keys built from a secret, not uniform words.

Key structure follows the reference key generators
(src/lib/kernel/keygeneration.cu:145-185 relinkey_gen_kernel, :584-629
relinkey_gen_II_kernel): rk1 = a, rk0 = -(a*s + e) + [limb in digit]*P*s_target.
The CPU variant runs the oracle; the gpu variant runs the CUDA path."""
import numpy as np
import pytest

from tests.common import PARAMS, oracle_ctx, splitmix64


def _small_poly(seed, n, bound):
    r = splitmix64(seed, n).astype(np.int64)
    return (r % (2 * bound + 1)) - bound


def _to_rns(poly, primes):
    return np.stack([(poly % p).astype(np.uint64) for p in primes])


def _mulmod(a, b, p):
    return ((a.astype(object) * b.astype(object)) % p).astype(np.uint64)


class Scheme:
    """Minimal symmetric RLWE machinery over the oracle's NTT (test-only)."""

    def __init__(self, name):
        self.oc = oracle_ctx(name)
        oc = self.oc
        self.n, self.Q, self.K, self.Qp = oc.n, oc.Q, oc.K, oc.Qp
        self.primes = oc.primes
        self.allp = list(range(oc.Qp))
        self.s = _small_poly(1001, self.n, 1)
        self.s_ntt = self.ntt(_to_rns(self.s, self.primes))
        self.Pprod = 1
        for p in self.primes[self.Q:]:
            self.Pprod *= p

    def ntt(self, x, order=None, inverse=False):
        return self.oc.ntt(x, order if order is not None else self.allp[: x.shape[-2]], inverse)

    def mul(self, a, b, plist):
        return np.stack([_mulmod(a[i], b[i], p) for i, p in enumerate(plist)])

    def add(self, a, b, plist):
        pr = np.array(plist, dtype=np.uint64)[:, None]
        s = a + b
        return np.where(s >= pr, s - pr, s)

    def neg(self, a, plist):
        pr = np.array(plist, dtype=np.uint64)[:, None]
        return np.where(a == 0, a, pr - a)

    def encrypt(self, m, seed, L):
        pl = self.primes[:L]
        a = np.stack([splitmix64(seed + i, self.n) % np.uint64(p) for i, p in enumerate(pl)])
        e = self.ntt(_to_rns(_small_poly(seed + 77, self.n, 4), pl))
        m_ntt = self.ntt(_to_rns(m, pl))
        c0 = self.add(self.add(self.neg(self.mul(a, self.s_ntt[:L], pl), pl), m_ntt, pl), e, pl)
        return np.stack([c0, a])

    def switch_key(self, target_ntt, source_s_ntt, seed):
        """[d][2][Qp][N]: switches a ciphertext part under `target` to the key `source_s`."""
        oc = self.oc
        d = oc.digits(0)
        key = np.zeros((d, 2, self.Qp, self.n), dtype=np.uint64)
        for i in range(d):
            a = np.stack([splitmix64(seed + 100 * i + y, self.n) % np.uint64(p) for y, p in enumerate(self.primes)])
            e = self.ntt(_to_rns(_small_poly(seed + 100 * i + 55, self.n, 4), self.primes))
            rk0 = self.neg(self.add(self.mul(a, source_s_ntt, self.primes), e, self.primes), self.primes)
            for y in range(self.Q):
                in_digit = (y == i) if oc.method == 1 else (y // self.K == i)
                if in_digit:
                    fac = self.Pprod % self.primes[y]
                    t = _mulmod(target_ntt[y], np.full(self.n, fac, dtype=np.uint64), self.primes[y])
                    s_ = rk0[y] + t
                    rk0[y] = np.where(s_ >= self.primes[y], s_ - np.uint64(self.primes[y]), s_)
            key[i, 0], key[i, 1] = rk0, a
        return key

    def decrypt_coeffs(self, ct2, L, count=48):
        """centered integer value of the first `count` coefficients of c0 + c1*s mod Q_L."""
        pl = self.primes[:L]
        d = self.add(ct2[0], self.mul(ct2[1], self.s_ntt[:L], pl), pl)
        coef = self.ntt(d, list(range(L)), inverse=True)
        Qprod = 1
        for p in pl:
            Qprod *= p
        out = []
        for j in range(count):
            v = 0
            for i, p in enumerate(pl):
                Mi = Qprod // p
                v += int(coef[i, j]) * Mi * pow(Mi, -1, p)
            v %= Qprod
            out.append(v - Qprod if v > Qprod // 2 else v)
        return out, Qprod


def _negacyclic_product(a, b, n, count):
    """first `count` coefficients of a*b mod X^n+1 over the integers"""
    a = [int(v) for v in a]
    b = [int(v) for v in b]
    out = []
    for k in range(count):
        acc = 0
        for i in range(n):
            j = k - i
            if j >= 0:
                acc += a[i] * b[j]
            else:
                acc -= a[i] * b[j + n]
        out.append(acc)
    return out


def _run(name, backend):
    sc = Scheme(name)
    oc = sc.oc
    n, L = sc.n, sc.Q
    m1, m2 = _small_poly(2001, n, 1 << 14), _small_poly(2002, n, 1 << 14)
    ct1, ct2 = sc.encrypt(m1, 3001, L), sc.encrypt(m2, 4001, L)
    s2_ntt = sc.mul(sc.s_ntt, sc.s_ntt, sc.primes)
    rk = sc.switch_key(s2_ntt, sc.s_ntt, 5001)
    count = 24
    want = _negacyclic_product(m1, m2, n, count)

    res = backend(oc, ct1, ct2, rk)
    got, Qprod = sc.decrypt_coeffs(res["relin"], L, count)
    noise = max(abs(g - w) for g, w in zip(got, want))
    assert noise < 1 << 48, f"mul+relin noise too large: 2^{noise.bit_length()}"
    assert Qprod.bit_length() > 90

    # rescale divides the plaintext by q_{L-1} (rounded)
    got_r, _ = sc.decrypt_coeffs(res["rescale"], L - 1, count)
    ql = sc.primes[L - 1]
    for g, w in zip(got_r, want):
        assert abs(g * ql - w) < (1 << 48) + ql * 64

    # rotation by galois element g: decrypts to m1(X^g)
    g = res["galois_elt"]
    got_g, _ = sc.decrypt_coeffs(res["rot"], L, n)  # all coefficients
    exp = [0] * n
    for i in range(n):
        raw = (i * g) % (2 * n)
        exp[raw % n] = -int(m1[i]) if raw >= n else int(m1[i])
    noise_g = max(abs(a - b) for a, b in zip(got_g[:count * 8], exp[:count * 8]))
    assert noise_g < 1 << 40, f"rotation noise too large: 2^{noise_g.bit_length()}"


def _galois_key(sc, g):
    n = sc.n
    ginv = pow(g, -1, 2 * n)
    s_perm = np.zeros(n, dtype=np.int64)
    for i in range(n):
        raw = (i * ginv) % (2 * n)
        s_perm[raw % n] = -sc.s[i] if raw >= n else sc.s[i]
    sp_ntt = sc.ntt(_to_rns(s_perm, sc.primes))
    # encrypts P*s under the key s(X^{g^-1}); the permutation after the key switch maps it back
    return sc.switch_key(sc.s_ntt, sp_ntt, 9001)


def _oracle_backend(oc, ct1, ct2, rk):
    sc = Scheme.__new__(Scheme)
    m = oc.multiply(ct1, ct2)
    r = oc.relinearize(m, rk)
    out = {"relin": r[:2], "rescale": oc.rescale(r[:2])}
    sc2 = Scheme(_name_of(oc))
    g = 5
    out["galois_elt"] = g
    out["rot"] = oc.apply_galois(ct1, _galois_key(sc2, g), g)
    return out


def _name_of(oc):
    for k, (log_n, qb, pb) in PARAMS.items():
        if log_n == oc.n_power and len(qb) == oc.Q and len(pb) == oc.K and oracle_ctx(k) is oc:
            return k
    raise KeyError


@pytest.mark.parametrize("name", ["n12_I", "n12_II"])
def test_decrypt_level_oracle(name):
    _run(name, _oracle_backend)


def _gpu_backend(oc, ct1, ct2, rk):
    import torch
    from heongpu_b200 import api
    from tests.gpu_common import gpu_ctx, to_dev, to_host
    name = _name_of(oc)
    ctx = gpu_ctx(name)
    op = api.HEArithmeticOperator(ctx)
    L, n = oc.Q, oc.n
    A, B = api.Ciphertext(ctx, to_dev(ct1)), api.Ciphertext(ctx, to_dev(ct2))
    Cc = api.Ciphertext(ctx, torch.zeros(1, 3, L, n, dtype=torch.int64, device="cuda"))
    op.multiply(A, B, Cc)
    op.relinearize_inplace(Cc, api.Relinkey(ctx, to_dev(rk)))
    out = {"relin": to_host(Cc.words())[0].copy()}
    op.rescale_inplace(Cc)
    out["rescale"] = to_host(Cc.words())[0].copy()
    sc2 = Scheme(name)
    g = api.lib.heon_steps_to_galois_elt(1, n, 5)
    gk = api.Galoiskey(ctx, {g: to_dev(_galois_key(sc2, g))})
    R_ = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"))
    op.rotate_rows(A, R_, gk, 1)
    out["galois_elt"] = g
    out["rot"] = to_host(R_.data)[0].copy()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["n12_I", "n12_II", "n13_II"])
def test_decrypt_level_gpu(name):
    _run(name, _gpu_backend)

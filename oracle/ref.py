"""ctypes wrappers of the reference's own code compiled into oracle/_ref/
(libref_host.so: host table generators + CPU NTT; libref_gpu.so: the CUDA
kernels + launch replay).  TEST / BASELINE INFRASTRUCTURE ONLY.  The .so files
are built in the build container by `make -C oracle ref` (needs
/root/reference) and travel to the GPU box as prebuilt files."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(_HERE, "_ref", "libref_host.so")
GPU_SO = os.path.join(_HERE, "_ref", "libref_gpu.so")
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int)


def have_host():
    return os.path.exists(HOST_SO)


def have_gpu():
    return os.path.exists(GPU_SO)


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(i32p)


_host = None


def host():
    global _host
    if _host is None:
        L = C.CDLL(HOST_SO)
        L.ref_mult.restype = C.c_uint64
        L.ref_mult.argtypes = [C.c_uint64] * 3
        _host = L
    return _host


def modulus(p):
    out = (C.c_uint64 * 3)()
    host().ref_modulus(C.c_uint64(p), out)
    return tuple(int(v) for v in out)


def generate_primes(n, bits):
    b = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    if host().ref_generate_primes(n, b, len(bits), _p(out)):
        raise RuntimeError("reference generate_primes failed")
    return [int(v) for v in out]


def ntt_tables(primes, n_power):
    n = 1 << n_power
    pr = np.array(primes, dtype=np.uint64)
    psi = np.zeros(len(primes), dtype=np.uint64)
    fwd = np.zeros(len(primes) * n, dtype=np.uint64)
    inv = np.zeros(len(primes) * n, dtype=np.uint64)
    ninv = np.zeros(len(primes), dtype=np.uint64)
    host().ref_ntt_tables(_p(pr), len(primes), n_power, _p(psi), _p(fwd), _p(inv), _p(ninv))
    return psi, fwd, inv, ninv


def moddown_tables(primes, Q, K):
    Qp = Q + K
    pr = np.array(primes, dtype=np.uint64)
    lqm = np.zeros(K * Qp, dtype=np.uint64)
    half = np.zeros(K, dtype=np.uint64)
    hm = np.zeros(K * Qp, dtype=np.uint64)
    fac = np.zeros(K * Q, dtype=np.uint64)
    w = host().ref_moddown_tables(_p(pr), Qp, K, Q, _p(lqm), _p(half), _p(hm), _p(fac))
    return lqm[:w].copy(), half, hm[:w].copy(), fac


def method2_tables(n, primes, Q, K, depth):
    Qp = Q + K
    pr = np.array(primes, dtype=np.uint64)
    bc = np.zeros(Q * Qp * (K + 1), dtype=np.uint64)
    mi = np.zeros(Q, dtype=np.uint64)
    prod = np.zeros(Q * Qp, dtype=np.uint64)
    ij = np.zeros(Q, dtype=np.int32)
    il = np.zeros(Q, dtype=np.int32)
    cnt = np.zeros(3, dtype=np.int32)
    d = host().ref_method2_tables(n, _p(pr), Qp, K, depth, _p(bc), _p(mi), _p(prod), _ip(ij), _ip(il), _ip(cnt))
    return dict(d=d, base_change=bc[: cnt[0]].copy(), mi_inv=mi[: cnt[1]].copy(), prod=prod[: cnt[2]].copy(),
                I_j=ij[:d].copy(), I_location=il[:d].copy())


def ntt_cpu(a, n_power, p, psi, inverse=False):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    host().ref_ntt_cpu(_p(a), n_power, C.c_uint64(p), C.c_uint64(psi), int(inverse))
    return a


class RefGpu:
    """Reference CUDA kernels + operator.cu launch replay on torch device buffers.
    Tables are passed in (from the oracle or from libref_host)."""

    def __init__(self, n_power, primes, Q, K, tables):
        L = C.CDLL(GPU_SO)
        L.refgpu_create.restype = C.c_void_p
        self.L = L
        self.n_power, self.n, self.Q, self.K, self.Qp = n_power, 1 << n_power, Q, K, Q + K
        pr = np.array(primes, dtype=np.uint64)
        t = tables
        self._h = C.c_void_p(L.refgpu_create(
            n_power, Q, K, _p(pr), _p(t["fwd"]), _p(t["inv"]), _p(t["ninv"]),
            _p(t["last_q_modinv"]), len(t["last_q_modinv"]), _p(t["half"]), _p(t["half_mod"]),
            _p(t["r_modinv"]), _p(t["r_half_mod"]), len(t["r_modinv"]), _p(t["r_half"])))
        if K > 1:
            for depth in range(Q):
                m = t["method2"][depth]
                L.refgpu_set_method2(self._h, depth, _p(m["base_change"]), len(m["base_change"]),
                                     _p(m["mi_inv"]), len(m["mi_inv"]), _p(m["prod"]), len(m["prod"]),
                                     _ip(m["I_j"]), _ip(m["I_location"]), int(m["d"]))

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.refgpu_destroy(self._h)
            self._h = None

    @staticmethod
    def _s(stream):
        import torch
        s = stream if stream is not None else torch.cuda.current_stream()
        return C.c_void_p(s.cuda_stream)

    def _chk(self, rc):
        if rc:
            raise RuntimeError(f"reference kernel launch failed: cuda error {rc}")

    def ntt(self, data, mod_count, inverse=False, stream=None):
        self._chk(self.L.refgpu_ntt(self._h, C.c_void_p(data.data_ptr()), data.numel() // self.n, mod_count, int(inverse), self._s(stream)))

    def ntt_level(self, data, depth, inverse=False, stream=None):
        self._chk(self.L.refgpu_ntt_level(self._h, C.c_void_p(data.data_ptr()), data.numel() // self.n, depth, int(inverse), self._s(stream)))

    def multiply(self, a, b, out, depth=0, stream=None):
        self._chk(self.L.refgpu_multiply(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), depth, self._s(stream)))

    def addsub(self, a, b, out, comps, depth, op, stream=None):
        self._chk(self.L.refgpu_addsub(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), comps, depth, op, self._s(stream)))

    def relinearize(self, ct, key, depth=0, stream=None):
        self._chk(self.L.refgpu_relinearize(self._h, C.c_void_p(ct.data_ptr()), C.c_void_p(key.data_ptr()), depth, self._s(stream)))

    def rescale(self, ct, depth=0, stream=None):
        self._chk(self.L.refgpu_rescale(self._h, C.c_void_p(ct.data_ptr()), depth, self._s(stream)))

    def apply_galois(self, a, out, key, galois_elt, depth=0, stream=None):
        self._chk(self.L.refgpu_apply_galois(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(key.data_ptr()), int(galois_elt), depth, self._s(stream)))

    def bsgs_matvec(self, ct, out, matrix, baby_elts, baby_keys, giant_elts, giant_keys, group_sizes, term_baby, depth=0,
                    stream=None):
        """multiply_matrix_v2, one matrix, without the trailing rescale (ckks/operator.cu:2898-3383)."""
        vp = C.c_void_p
        kp = lambda k: k.data_ptr() if k is not None else None
        self._chk(self.L.refgpu_bsgs_matvec(
            self._h, vp(ct.data_ptr()), vp(out.data_ptr()), vp(matrix.data_ptr()), (C.c_int * len(baby_elts))(*baby_elts),
            (vp * len(baby_keys))(*[kp(k) for k in baby_keys]), len(baby_elts), (C.c_int * len(giant_elts))(*giant_elts),
            (vp * len(giant_keys))(*[kp(k) for k in giant_keys]), (C.c_int * len(group_sizes))(*group_sizes),
            (C.c_int * len(term_baby))(*term_baby), len(giant_elts), depth, self._s(stream)))


def tables_for_refgpu(n_power, primes, Q, K, use_ref_host=None, scheme="CKKS", plain_modulus=786433):
    """Build the table bundle RefGpu needs, from the reference's own host code
    when libref_host.so is present, else from the CPU oracle (identical values,
    see tests/test_oracle_vs_ref_host.py).  scheme="BFV": the Method-II tables are the UN-LEVELLED
    ones of the reference's BFV context (digit size 2, contextpool.cpp:160-191 ff.), read from the
    reference's own HEContextImpl<BFV> (libref_ctx.so) when it was built."""
    from . import oracle as O
    if use_ref_host is None:
        use_ref_host = have_host()
    src = __import__(__name__, fromlist=["x"]) if use_ref_host else O
    psi, fwd, inv, ninv = src.ntt_tables(primes, n_power)
    lqm, half, hm, fac = src.moddown_tables(primes, Q, K)
    r_modinv, r_half_mod, r_half = O.rescale_tables(primes, Q)  # built inline in ckks/context.cu:342-368
    t = dict(psi=psi, fwd=fwd, inv=inv, ninv=ninv, last_q_modinv=lqm, half=half, half_mod=hm,
             r_modinv=r_modinv if len(r_modinv) else np.zeros(1, dtype=np.uint64),
             r_half_mod=r_half_mod if len(r_half_mod) else np.zeros(1, dtype=np.uint64),
             r_half=r_half if len(r_half) else np.zeros(1, dtype=np.uint64))
    if K > 1 and scheme == "BFV":
        if use_ref_host and have_ctx():
            rc = RefContext("BFV", 1 << n_power, primes[:Q], primes[Q:], plain_modulus=plain_modulus)
            m = dict(base_change=rc.table(12), mi_inv=rc.table(13), prod=rc.table(14),
                     I_j=rc.table(15).astype(np.int32), I_location=rc.table(16).astype(np.int32))
            m["d"] = len(m["I_j"])
        else:
            m = O.method2_tables(primes, Q, K, 0, digit_m=2)
        t["method2"] = [m] * Q
    elif K > 1:
        if use_ref_host:
            t["method2"] = [method2_tables(1 << n_power, primes, Q, K, d) for d in range(Q)]
        else:
            t["method2"] = [O.method2_tables(primes, Q, K, d) for d in range(Q)]
    return t


class RefBfv:
    """The reference's fast_convertion / fast_floor / cross_multiplication kernels and GPU-NTT,
    replaying multiply_bfv (bfv/operator.cu:336-430).  BEHZ tables come from the oracle (the
    reference builds them inside HEContextImpl<BFV>, which cannot be compiled offline); the
    decrypt-level test pins those tables semantically."""

    def __init__(self, ob):
        from . import oracle as O
        L = C.CDLL(GPU_SO)
        L.refgpu_bfv_create.restype = C.c_void_p
        self.L, self.ob = L, ob
        Q, m, n_power = ob.Q, ob.bsk, ob.n_power
        merged = ob.primes[:Q] + ob.bsk_primes
        psi, fwd, inv, ninv = O.ntt_tables(merged, n_power)
        t = [np.ascontiguousarray(ob.table(w)) for w in range(20, 31)]
        sc = t[10]
        q = np.array(ob.primes[:Q], dtype=np.uint64)
        b = np.array(ob.bsk_primes, dtype=np.uint64)
        L.refgpu_bfv_create.argtypes = [C.c_int, C.c_int, C.c_int, u64p, u64p, C.c_uint64] + [u64p] * 13 + [C.c_uint64, C.c_uint64]
        self._h = C.c_void_p(L.refgpu_bfv_create(n_power, Q, m, _p(q), _p(b), C.c_uint64(ob.t), _p(fwd), _p(inv), _p(ninv),
                                                 *[_p(x) for x in t[:10]], C.c_uint64(int(sc[0])), C.c_uint64(int(sc[1]))))

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.refgpu_bfv_destroy(self._h)
            self._h = None

    def multiply(self, a, b, out, stream=None):
        rc = self.L.refgpu_bfv_multiply(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()),
                                        C.c_void_p(out.data_ptr()), RefGpu._s(stream))
        if rc:
            raise RuntimeError(f"reference kernel launch failed: cuda error {rc}")


def bfv_relinearize(refgpu, ct, key, stream=None):
    """relinearize_seal_method_inplace / _external_product_method2_inplace on a RefGpu handle."""
    refgpu._chk(refgpu.L.refgpu_bfv_relinearize(refgpu._h, C.c_void_p(ct.data_ptr()), C.c_void_p(key.data_ptr()),
                                                 RefGpu._s(stream)))


def bfv_apply_galois(refgpu, a, out, key, galois_elt, stream=None):
    """apply_galois_method_I / _II of the BFV operator on a RefGpu handle."""
    refgpu._chk(refgpu.L.refgpu_bfv_apply_galois(refgpu._h, C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(key.data_ptr()), int(galois_elt), RefGpu._s(stream)))


# ---------------------------------------------------------------------------------------------
# libref_ctx.so: the reference's OWN HEContextImpl<BFV> / HEContextImpl<CKKS> (bfv/context.cu,
# ckks/context.cu compiled unmodified on host-only infrastructure shims, oracle/ref_ctx_harness.cu)
# ---------------------------------------------------------------------------------------------
CTX_SO = os.path.join(_HERE, "_ref", "libref_ctx.so")


def have_ctx():
    return os.path.exists(CTX_SO)


class RefContext:
    """Tables of a reference context, addressed by the HEON_TBL_* codes of include/heon_b200.h
    (BFV extras: 40..43 merged q||Bsk modulus / NTT / INTT / n^-1, 44 {m, l, l_tilda, d})."""

    def __init__(self, scheme, n, q_values=None, p_values=None, plain_modulus=None, default_p_size=None):
        L = C.CDLL(CTX_SO)
        for f in ("refctx_bfv_create", "refctx_bfv_create_default", "refctx_ckks_create"):
            getattr(L, f).restype = C.c_void_p
        L.refctx_bfv_table.restype = C.c_longlong
        L.refctx_ckks_table.restype = C.c_longlong
        self.L, self.scheme = L, scheme
        if default_p_size is not None:
            h = L.refctx_bfv_create_default(n, default_p_size, int(plain_modulus))
        else:
            q = np.array(q_values, dtype=np.uint64)
            p = np.array(p_values, dtype=np.uint64)
            if scheme == "BFV":
                h = L.refctx_bfv_create(n, _p(q), len(q), _p(p), len(p), int(plain_modulus))
            else:
                h = L.refctx_ckks_create(n, _p(q), len(q), _p(p), len(p))
        if not h:
            raise RuntimeError("reference context construction failed")
        self._h = C.c_void_p(h)

    def __del__(self):
        if getattr(self, "_h", None):
            (self.L.refctx_bfv_destroy if self.scheme == "BFV" else self.L.refctx_ckks_destroy)(self._h)
            self._h = None

    def table(self, which, depth=0):
        if self.scheme == "BFV":
            call = lambda out, cap: self.L.refctx_bfv_table(self._h, which, out, C.c_longlong(cap))
        else:
            call = lambda out, cap: self.L.refctx_ckks_table(self._h, which, depth, out, C.c_longlong(cap))
        n = call(None, 0)
        if n < 0:
            raise KeyError(which)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        call(_p(out), n)
        return out[:n]


# ---------------------------------------------------------------------------------------------
# TFHE: the reference's small_ntt.cu + bootstrapping.cu kernels and the launch replay of
# src/lib/host/tfhe/operator.cu (oracle/ref_tfhe_harness.cu)
# ---------------------------------------------------------------------------------------------
TFHE_SO = os.path.join(_HERE, "_ref", "libref_tfhe.so")


def have_tfhe():
    return os.path.exists(TFHE_SO)


class RefTfhe:
    def __init__(self):
        L = C.CDLL(TFHE_SO)
        L.reftfhe_create.restype = C.c_void_p
        self.L = L
        self._h = C.c_void_p(L.reftfhe_create())

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.reftfhe_destroy(self._h)
            self._h = None

    @staticmethod
    def _s():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _chk(rc):
        if rc:
            raise RuntimeError(f"reference kernel launch failed: {rc}")

    def gate_linear(self, gate, a1, b1, a2, b2, oa, ob, n, shape):
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._chk(self.L.reftfhe_gate_linear(self._h, gate, p(a1), p(b1), p(a2), p(b2), p(oa), p(ob), n, shape, self._s()))

    def bootstrap(self, in_a, in_b, out_a, out_b, bk, shape):
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.L.reftfhe_bootstrap(self._h, p(in_a), p(in_b), p(out_a), p(out_b), p(bk), shape, self._s()))

    def keyswitch(self, in_a, in_b, out_a, out_b, ks_a, ks_b, shape):
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.L.reftfhe_keyswitch(self._h, p(in_a), p(in_b), p(out_a), p(out_b), p(ks_a), p(ks_b), shape, self._s()))

    def ntt(self, data, inverse=False):
        self._chk(self.L.reftfhe_ntt(self._h, C.c_void_p(data.data_ptr()), data.numel() // 1024, int(inverse), self._s()))

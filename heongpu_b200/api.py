"""Host-side mirror of the reference's operator interface for the hot path.

Names, argument meaning and error behaviour follow the reference classes
(src/include/heongpu/host/ckks/{context,ciphertext,evaluationkey,operator}.cuh):
``HEContext`` / ``Ciphertext`` / ``Relinkey`` / ``Galoiskey`` /
``HEArithmeticOperator.{add,sub,multiply,relinearize_inplace,rescale_inplace,
mod_drop_inplace,rotate_rows,apply_galois}``.  Additive extension: a
ciphertext object may carry a batch of B independent ciphertexts
(``data`` of shape [B, components, L, N]); the reference is the B = 1 case.

PyTorch is used only for device memory and streams.  Every operator calls the
C ABI of ``libheon_b200.so``; nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import build_library  # noqa: F401

lib = _lib.load()

TBL = dict(
    modulus=0, psi=1, ntt=2, intt=3, n_inverse=4, last_q_modinv=5, half=6, half_mod=7, factor=8,
    rescaled_last_q_modinv=9, rescaled_half_mod=10, rescaled_half=11,
    ii_base_change=12, ii_mi_inv=13, ii_prod=14, ii_i_j=15, ii_i_location=16,
    bfv_base_change_bsk=20, bfv_inv_punct_q=21, bfv_base_change_mtilde=22, bfv_inv_mtilde_mod_bsk=23,
    bfv_prod_q_mod_bsk=24, bfv_inv_prod_q_mod_bsk=25, bfv_base_change_q=26, bfv_base_change_msk=27,
    bfv_inv_punct_b=28, bfv_prod_b_mod_q=29, bfv_scalars=30, bfv_plain=31,
)

_EXC = {-1: ValueError, -2: RuntimeError, -3: RuntimeError, -4: RuntimeError}


class HeonError(RuntimeError):
    pass


class HeonInvalidArgument(HeonError, ValueError):  # std::invalid_argument
    pass


class HeonLogicError(HeonError):  # std::logic_error
    pass


def _check(status):
    if status == 0:
        return
    msg = lib.heon_last_error().decode()
    if status == -1:
        raise HeonInvalidArgument(msg)
    if status == -2:
        raise HeonLogicError(msg)
    raise HeonError(msg)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class HEContext:
    """HEContext<Scheme::CKKS> (reference: src/lib/host/ckks/context.cu:26-539)."""

    def __init__(self, log_n, q_bits=None, p_bits=None, q_values=None, p_values=None, device=0,
                 plain_modulus=None):
        h = C.c_void_p()
        self.scheme = "BFV" if plain_modulus else "CKKS"
        if plain_modulus and q_values is not None:  # set_coeff_modulus_values / default values (bfv/context.cu:223-300)
            q = (C.c_uint64 * len(q_values))(*q_values)
            p = (C.c_uint64 * len(p_values))(*p_values)
            _check(lib.heon_bfv_context_create_values(device, log_n, q, len(q_values), p, len(p_values),
                                                      int(plain_modulus), C.byref(h)))
        elif plain_modulus:  # HEContext<Scheme::BFV> (reference: src/lib/host/bfv/context.cu)
            q = (C.c_int * len(q_bits))(*q_bits)
            p = (C.c_int * len(p_bits))(*p_bits)
            _check(lib.heon_bfv_context_create(device, log_n, q, len(q_bits), p, len(p_bits), int(plain_modulus), C.byref(h)))
        elif q_values is not None:
            q = (C.c_uint64 * len(q_values))(*q_values)
            p = (C.c_uint64 * len(p_values))(*p_values)
            _check(lib.heon_ckks_context_create_values(device, log_n, q, len(q_values), p, len(p_values), C.byref(h)))
        else:
            q = (C.c_int * len(q_bits))(*q_bits)
            p = (C.c_int * len(p_bits))(*p_bits)
            _check(lib.heon_ckks_context_create(device, log_n, q, len(q_bits), p, len(p_bits), C.byref(h)))
        self._h = h
        info = _lib.heon_info()
        _check(lib.heon_context_info(h, C.byref(info)))
        self.n, self.n_power = info.n, info.log_n
        self.Q_size, self.P_size = info.q_size, info.p_size
        self.Q_prime_size = info.q_size + info.p_size
        self.keyswitch_method = info.keyswitch_method
        self.device = info.device
        allp = [int(v) for v in self.table("modulus").reshape(-1, 3)[:, 0]]
        self.primes = allp[: self.Q_prime_size]
        self.bsk_primes = allp[self.Q_prime_size:]  # BFV auxiliary base (empty for CKKS)
        self.plain_modulus = int(plain_modulus) if plain_modulus else None

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.heon_context_destroy(self._h)
            self._h = None

    def table(self, name, depth=0):
        cnt = C.c_size_t()
        _check(lib.heon_context_table(self._h, TBL[name], depth, None, 0, C.byref(cnt)))
        out = np.zeros(cnt.value, dtype=np.uint64)
        _check(lib.heon_context_table(self._h, TBL[name], depth, out.ctypes.data_as(_lib.u64p), cnt.value, C.byref(cnt)))
        return out

    def digits(self, depth=0):
        L = self.Q_size - depth
        if self.keyswitch_method == 1:
            return L
        # Method II digit size: |P| for CKKS, always 2 for BFV (contextpool.hpp:29, contextpool.cpp:78-104)
        return -(-L // (self.P_size if self.scheme == "CKKS" else 2))

    # ---- NTT (gpuntt::GPU_NTT / GPU_INTT / *_Modulus_Ordered) ----
    def ntt(self, data, prime_index=None, inverse=False, out=None, stream=None):
        """data: uint64/int64 cuda tensor [..., N]; poly z uses prime_index[z % len]."""
        n_polys = data.numel() // self.n
        out = data if out is None else out
        if prime_index is None:
            raise ValueError("prime_index required")
        arr = (C.c_int * len(prime_index))(*prime_index)
        _check(lib.heon_ntt(self._h, _ptr(data), _ptr(out), n_polys, arr, len(prime_index), int(inverse), _stream(stream)))
        return out

    def ntt_poly_ordered(self, base, offsets, prime_index, inverse=False, stream=None):
        arr = (C.c_longlong * len(offsets))(*offsets)
        _check(lib.heon_ntt_poly_ordered(self._h, _ptr(base), arr, len(offsets), prime_index, int(inverse), _stream(stream)))

    def level_primes(self, depth=0):
        L = self.Q_size - depth
        return list(range(L)) + [self.Q_size + j for j in range(self.P_size)]


class Ciphertext:
    """Ciphertext<Scheme::CKKS>: flat [cipher_size][L][N] words, NTT domain
    (reference: src/lib/host/ckks/ciphertext.cu:20-31), optionally batched."""

    def __init__(self, context, data, depth=0, cipher_size=None, scale=1.0,
                 relinearization_required=False, rescale_required=False):
        self.context = context
        if data.dim() == 3:
            data = data.unsqueeze(0)
        self.data = data  # [B, comps_allocated, L_allocated, N] int64 view of uint64 words
        self.depth_ = depth
        self.cipher_size_ = cipher_size if cipher_size is not None else data.shape[1]
        self.scale_ = scale
        self.relinearization_required_ = relinearization_required
        self.rescale_required_ = rescale_required
        self.in_ntt_domain_ = True

    @property
    def batch(self):
        return self.data.shape[0]

    @property
    def stride(self):
        return self.data.stride(0)

    def level_count(self):
        return self.context.Q_size - self.depth_

    def words(self):
        """The live [B, cipher_size, L, N] words (the buffer may be larger after in-place ops)."""
        L, N = self.level_count(), self.context.n
        flat = self.data.reshape(self.batch, -1)[:, : self.cipher_size_ * L * N]
        return flat.reshape(self.batch, self.cipher_size_, L, N)


class _EvalKey:
    def __init__(self, context, data):
        self.context = context
        self.data = data


class Relinkey(_EvalKey):
    """Relinkey<Scheme::CKKS>: [digit][2][Q'_0][N] NTT-domain words
    (reference: src/lib/kernel/keygeneration.cu:180-183)."""


class Switchkey(_EvalKey):
    """Switchkey<Scheme::CKKS>: same layout as the Relinkey, re-encrypts under another secret key
    (reference: src/lib/host/ckks/evaluationkey.cu)."""


class Plaintext:
    """Plaintext<Scheme::CKKS>: [L][N] words in the NTT domain (reference: ckks/plaintext.cu), optionally batched."""

    def __init__(self, context, data, depth=0, scale=1.0):
        self.context = context
        if data.dim() <= 2:  # CKKS [L][N] / BFV [N] -> a batch of one
            data = data.unsqueeze(0)
        self.data = data
        self.depth_ = depth
        self.scale_ = scale

    @property
    def stride(self):
        return self.data.stride(0) if self.data.shape[0] > 1 else 0


class Galoiskey:
    """Galoiskey<Scheme::CKKS>: map galois_elt -> key of the Relinkey layout
    (reference: src/lib/host/ckks/evaluationkey.cu, device_location_)."""

    group_order_ = 5

    def __init__(self, context, keys, conjugate_key=None):
        self.context = context
        self.device_location_ = dict(keys)
        self.galois_elt_zero = 2 * context.n - 1
        self.zero_device_location_ = conjugate_key  # Galoiskey::c_data()


class HEArithmeticOperator:
    """HEArithmeticOperator<Scheme::CKKS>, hot-path subset
    (reference: src/include/heongpu/host/ckks/operator.cuh:95-1600)."""

    def __init__(self, context):
        self.context_ = context

    # -- element-wise --
    def _binary(self, fn, a, b, out):
        if a.depth_ != b.depth_:
            raise HeonLogicError("Ciphertexts leveled are not equal")
        c = self.context_
        comps = max(a.cipher_size_, b.cipher_size_)
        if a.cipher_size_ != b.cipher_size_:
            raise HeonInvalidArgument("Ciphertexts should have the same size")
        _check(fn(c._h, _ptr(a.data), a.stride, _ptr(b.data), b.stride, _ptr(out.data), out.stride,
                  comps, a.depth_, a.batch, _stream()))
        out.depth_, out.cipher_size_, out.scale_ = a.depth_, comps, a.scale_
        out.relinearization_required_ = a.relinearization_required_
        out.rescale_required_ = a.rescale_required_
        return out

    def add(self, a, b, out):
        return self._binary(lib.heon_add, a, b, out)

    def sub(self, a, b, out):
        return self._binary(lib.heon_sub, a, b, out)

    def negate(self, a, out):
        c = self.context_
        _check(lib.heon_negate(c._h, _ptr(a.data), a.stride, _ptr(out.data), out.stride, a.cipher_size_, a.depth_, a.batch, _stream()))
        out.depth_, out.cipher_size_ = a.depth_, a.cipher_size_
        return out

    # -- plaintext operands (operator.cuh:197-620,718-884) --
    def _plain(self, fn, ct, pt, out, scale):
        if ct.depth_ != pt.depth_:
            raise HeonLogicError("Ciphertexts leveled are not equal")
        c = self.context_
        _check(fn(c._h, _ptr(ct.data), ct.stride, _ptr(pt.data), pt.stride, _ptr(out.data), out.stride,
                  ct.cipher_size_, ct.depth_, ct.batch, _stream()))
        out.depth_, out.cipher_size_, out.scale_ = ct.depth_, ct.cipher_size_, scale
        out.relinearization_required_ = ct.relinearization_required_
        out.rescale_required_ = ct.rescale_required_
        return out

    def multiply_plain(self, ct, pt, out):
        out = self._plain(lib.heon_ckks_multiply_plain, ct, pt, out, ct.scale_ * pt.scale_)
        out.rescale_required_ = True
        return out

    def add_plain(self, ct, pt, out):
        return self._plain(lib.heon_ckks_add_plain, ct, pt, out, ct.scale_)

    def sub_plain(self, ct, pt, out):
        return self._plain(lib.heon_ckks_sub_plain, ct, pt, out, ct.scale_)

    # -- keyswitch / conjugate (operator.cuh:1282-1420) --
    def keyswitch(self, ct, out, switch_key):
        c = self.context_
        _check(lib.heon_ckks_keyswitch(c._h, _ptr(ct.data), ct.stride, _ptr(out.data), out.stride,
                                       _ptr(switch_key.data), ct.depth_, ct.batch, _stream()))
        out.depth_, out.cipher_size_, out.scale_ = ct.depth_, 2, ct.scale_
        return out

    def conjugate(self, ct, out, galois_key):
        c = self.context_
        if galois_key.zero_device_location_ is None:
            raise HeonLogicError("Conjugation key not present!")
        _check(lib.heon_ckks_conjugate(c._h, _ptr(ct.data), ct.stride, _ptr(out.data), out.stride,
                                       _ptr(galois_key.zero_device_location_), ct.depth_, ct.batch, _stream()))
        out.depth_, out.cipher_size_, out.scale_ = ct.depth_, 2, ct.scale_
        return out

    # -- multiply / relinearize / rescale (operator.cuh:631-707,1053-1094,1423-1445) --
    def multiply(self, a, b, out):
        if a.relinearization_required_ or b.relinearization_required_:
            raise HeonInvalidArgument("Ciphertexts can not be multiplied because of the non-linear part! Please use relinearization operation!")
        if a.rescale_required_ or b.rescale_required_:
            raise HeonInvalidArgument("Ciphertexts can not be multiplied because of the noise! Please use rescale operation to get rid of additional noise!")
        if a.depth_ != b.depth_:
            raise HeonLogicError("Ciphertexts leveled are not equal")
        c = self.context_
        _check(lib.heon_ckks_multiply(c._h, _ptr(a.data), a.stride, _ptr(b.data), b.stride, _ptr(out.data), out.stride,
                                      a.depth_, a.batch, _stream()))
        out.depth_, out.cipher_size_ = a.depth_, 3
        out.scale_ = a.scale_ * b.scale_
        out.relinearization_required_ = True
        out.rescale_required_ = True
        return out

    def relinearize_inplace(self, ct, relin_key):
        if not ct.relinearization_required_:
            raise HeonInvalidArgument("Ciphertexts can not use relinearization, since no non-linear part!")
        c = self.context_
        _check(lib.heon_ckks_relinearize(c._h, _ptr(ct.data), ct.stride, _ptr(relin_key.data), ct.depth_, ct.batch, _stream()))
        ct.cipher_size_ = 2
        ct.relinearization_required_ = False
        return ct

    def multiply_relinearize_host(self, h_a, h_b, h_out, relin_key, depth=0, rescale=False, chunk=0):
        """multiply + relinearize_inplace (+ rescale_inplace) on HOST-resident ciphertext batches
        (ExecutionOptions::set_storage_type(HOST), storagemanager.cuh:113-167): h_a, h_b [B, 2, L, N] and h_out
        [B, 2, L', N] are (pinned) host tensors; copies and compute are pipelined inside the library."""
        c = self.context_
        assert not h_a.is_cuda and not h_b.is_cuda and not h_out.is_cuda
        _check(lib.heon_ckks_multiply_relinearize_host(c._h, C.c_void_p(h_a.data_ptr()), C.c_void_p(h_b.data_ptr()),
                                                       C.c_void_p(h_out.data_ptr()), _ptr(relin_key.data), depth, int(rescale),
                                                       h_a.shape[0], chunk, _stream()))
        return h_out

    def rescale_inplace(self, ct):
        c = self.context_
        _check(lib.heon_ckks_rescale(c._h, _ptr(ct.data), ct.stride, ct.depth_, ct.batch, _stream()))
        ct.scale_ = ct.scale_ / float(c.primes[c.Q_size - ct.depth_ - 1])
        ct.depth_ += 1
        ct.rescale_required_ = False
        return ct

    # -- BFV (src/lib/host/bfv/operator.cu:336-430, 505-671); ciphertexts in the coefficient domain --
    def multiply_bfv(self, a, b, out):
        c = self.context_
        _check(lib.heon_bfv_multiply(c._h, _ptr(a.data), a.stride, _ptr(b.data), b.stride, _ptr(out.data), out.stride,
                                     a.batch, _stream()))
        out.cipher_size_, out.relinearization_required_, out.in_ntt_domain_ = 3, True, False
        return out

    def relinearize_inplace_bfv(self, ct, relin_key):
        c = self.context_
        _check(lib.heon_bfv_relinearize(c._h, _ptr(ct.data), ct.stride, _ptr(relin_key.data), ct.batch, _stream()))
        ct.cipher_size_, ct.relinearization_required_ = 2, False
        return ct

    def apply_galois_bfv(self, ct, out, galois_key, galois_elt):
        c = self.context_
        if galois_elt not in galois_key.device_location_:
            raise HeonLogicError("Galois key not present!")
        key = galois_key.device_location_[galois_elt]
        _check(lib.heon_bfv_apply_galois(c._h, _ptr(ct.data), ct.stride, _ptr(out.data), out.stride, _ptr(key),
                                         galois_elt, ct.batch, _stream()))
        out.cipher_size_ = 2
        return out

    def add_plain_bfv(self, ct, pt, out):
        c = self.context_
        _check(lib.heon_bfv_add_plain(c._h, _ptr(ct.data), ct.stride, _ptr(pt.data), pt.stride, _ptr(out.data), out.stride,
                                      ct.cipher_size_, ct.batch, _stream()))
        out.cipher_size_ = ct.cipher_size_
        return out

    def sub_plain_bfv(self, ct, pt, out):
        c = self.context_
        _check(lib.heon_bfv_sub_plain(c._h, _ptr(ct.data), ct.stride, _ptr(pt.data), pt.stride, _ptr(out.data), out.stride,
                                      ct.cipher_size_, ct.batch, _stream()))
        out.cipher_size_ = ct.cipher_size_
        return out

    def multiply_plain_bfv(self, ct, pt, out):
        c = self.context_
        _check(lib.heon_bfv_multiply_plain(c._h, _ptr(ct.data), ct.stride, _ptr(pt.data), pt.stride, _ptr(out.data),
                                           out.stride, ct.batch, _stream()))
        out.cipher_size_ = 2
        return out

    def keyswitch_bfv(self, ct, out, switch_key):
        c = self.context_
        _check(lib.heon_bfv_keyswitch(c._h, _ptr(ct.data), ct.stride, _ptr(out.data), out.stride, _ptr(switch_key.data),
                                      ct.batch, _stream()))
        out.cipher_size_ = 2
        return out

    def rotate_rows_bfv(self, ct, out, galois_key, shift):
        if shift == 0:  # the reference returns the input unchanged (bfv/operator.cuh:591-595)
            return self._identity(ct, out)
        return self.apply_galois_bfv(ct, out, galois_key, lib.heon_steps_to_galois_elt(shift, self.context_.n, 3))

    @staticmethod
    def _identity(ct, out):
        if out is not ct:
            out.data.copy_(ct.data)
            for k in ("depth_", "cipher_size_", "scale_", "relinearization_required_", "rescale_required_", "in_ntt_domain_"):
                if hasattr(ct, k):
                    setattr(out, k, getattr(ct, k))
        return out

    def rotate_columns_bfv(self, ct, out, galois_key):
        return self.apply_galois_bfv(ct, out, galois_key, 2 * self.context_.n - 1)

    def mod_drop_inplace(self, ct):
        c = self.context_
        _check(lib.heon_ckks_mod_drop_inplace(c._h, _ptr(ct.data), ct.stride, ct.cipher_size_, ct.depth_, ct.batch, _stream()))
        ct.depth_ += 1
        return ct

    # -- rotations (operator.cuh:1105-1270) --
    def apply_galois(self, ct, out, galois_key, galois_elt):
        c = self.context_
        if galois_elt not in galois_key.device_location_:
            raise HeonLogicError("Galois key not present!")
        key = galois_key.device_location_[galois_elt]
        _check(lib.heon_ckks_apply_galois(c._h, _ptr(ct.data), ct.stride, _ptr(out.data), out.stride, _ptr(key),
                                          galois_elt, ct.depth_, ct.batch, _stream()))
        out.depth_, out.cipher_size_, out.scale_ = ct.depth_, 2, ct.scale_
        out.relinearization_required_ = ct.relinearization_required_
        out.rescale_required_ = ct.rescale_required_
        return out

    def rotate_rows(self, ct, out, galois_key, shift):
        if shift == 0:  # the reference returns the input unchanged (ckks/operator.cuh:1123)
            return self._identity(ct, out)
        elt = lib.heon_steps_to_galois_elt(shift, self.context_.n, galois_key.group_order_)
        return self.apply_galois(ct, out, galois_key, elt)

    def rotate_rows_hoisted(self, ct, out_data, galois_key, shifts):
        """The baby-step loop of the reference's BSGS product (fast_single_hoisting_rotation_ckks_method_I/II,
        ckks/operator.cu:4674-5446): every shift of `shifts` applied to the same ciphertext(s), written to
        out_data[r] ([R, B, 2, L, N]); mod-up and the forward NTTs are shared by all rotations."""
        c = self.context_
        if 0 in shifts:  # shift 0 is the identity (ckks/operator.cuh:1123), not the conjugation element 2N-1
            nz = [i for i, s in enumerate(shifts) if s != 0]
            if nz:
                sub = torch.empty((len(nz),) + tuple(out_data.shape[1:]), dtype=out_data.dtype, device=out_data.device)
                self.rotate_rows_hoisted(ct, sub, galois_key, [shifts[i] for i in nz])
                out_data[nz] = sub
            for i, s in enumerate(shifts):
                if s == 0:
                    out_data[i].copy_(ct.data)
            return out_data
        elts = [lib.heon_steps_to_galois_elt(s, c.n, galois_key.group_order_) for s in shifts]
        for e in elts:
            if e not in galois_key.device_location_:
                raise HeonLogicError("Galois key not present!")
        keys = (C.c_void_p * len(elts))(*[galois_key.device_location_[e].data_ptr() for e in elts])
        earr = (C.c_uint32 * len(elts))(*elts)
        _check(lib.heon_ckks_rotate_hoisted(c._h, _ptr(ct.data), ct.stride, _ptr(out_data), out_data.stride(1),
                                            out_data.stride(0), keys, earr, len(elts), ct.depth_, ct.batch, _stream()))
        return out_data


    def multiply_plain_accumulate(self, cts, pts, out, depth=0):
        """out = sum_i cts[i] * pts[i]: the giant-step inner sum of the single-hoisting BSGS product
        (cipherplain_multiply_accumulate_kernel, multiplication.cu:374-403; ckks/operator.cu:2843-2853).
        cts: [count][2][L][N] (e.g. rotate_rows_hoisted's output), pts: [count][L][N], out: [2][L][N]."""
        c = self.context_
        _check(lib.heon_ckks_multiply_plain_accumulate(c._h, _ptr(cts), _ptr(pts), _ptr(out), cts.shape[0], depth, _stream()))
        return out

    @staticmethod
    def bsgs_plan(n, group_order, diags_bsgs, rot_n1, rot_n2):
        """Resolves one matrix's BSGS description the way multiply_matrix_v2 does (ckks/operator.cu:2926,
        3176-3193): rot_n2 = the baby-step shifts (sorted first), rot_n1[j] = giant-step shift of group j,
        diags_bsgs[j][k] - rot_n1[j] = the baby-step shift term k of group j reads."""
        baby = sorted(rot_n2)
        elt = lambda s: 0 if s == 0 else lib.heon_steps_to_galois_elt(s, n, group_order)
        terms = []
        for j, group in enumerate(diags_bsgs):
            for dg in group:
                terms.append(baby.index(dg - rot_n1[j]))
        return ([elt(s) for s in baby], [elt(s) for s in rot_n1], [len(g) for g in diags_bsgs], terms)

    def multiply_matrix(self, ct, out, matrix, diags_bsgs, rot_n1, rot_n2, galois_key, rescale=True):
        """One matrix of HEOperator<CKKS>::multiply_matrix_v2 (ckks/operator.cu:2898-3390): BSGS diagonal
        matrix-vector product with double hoisting in PQ_l, then rescale.  `matrix`: [terms][L+K][N] device
        words, the diagonals encoded over PQ_l in the NTT domain, in group order."""
        c = self.context_
        if ct.batch != 1:
            raise HeonError("multiply_matrix takes one ciphertext")
        baby, giant, sizes, terms = self.bsgs_plan(c.n, galois_key.group_order_, diags_bsgs, rot_n1, rot_n2)
        for e in baby + giant:
            if e and e not in galois_key.device_location_:
                raise HeonLogicError("Galois key not present!")
        kp = lambda e: galois_key.device_location_[e].data_ptr() if e else None
        bk = (C.c_void_p * len(baby))(*[kp(e) for e in baby])
        gk = (C.c_void_p * len(giant))(*[kp(e) for e in giant])
        _check(lib.heon_ckks_multiply_matrix(
            c._h, _ptr(ct.data), _ptr(out.data), _ptr(matrix), (C.c_uint32 * len(baby))(*baby), bk, len(baby),
            (C.c_uint32 * len(giant))(*giant), gk, (C.c_int * len(sizes))(*sizes), (C.c_int * len(terms))(*terms),
            len(giant), ct.depth_, _stream()))
        out.depth_, out.cipher_size_ = ct.depth_, 2
        # the reference multiplies the scale by prime_vector_[current_decomp_count] (:3384) before rescaling
        out.scale_ = ct.scale_ * float(c.primes[c.Q_size - ct.depth_])
        out.rescale_required_, out.relinearization_required_ = True, False
        if rescale:
            self.rescale_inplace(out)
        return out


# ---------------------------------------------------------------------------------------------
# client side: key generation, encryption, decryption, encoding
# ---------------------------------------------------------------------------------------------
class Secretkey:
    """Secretkey<S>: [Q'][N] NTT-domain words (src/lib/host/ckks/secretkey.cu)."""

    def __init__(self, context, hamming_weight=None):
        self.context = context
        self.hamming_weight_ = hamming_weight if hamming_weight is not None else context.n // 2
        self.data = None
        self.secret_key_generated_ = False


class Publickey:
    """Publickey<S>: [2][Q'][N] NTT-domain words (src/lib/host/ckks/publickey.cu)."""

    def __init__(self, context):
        self.context = context
        self.data = None
        self.public_key_generated_ = False


class HEKeyGenerator:
    """HEKeyGenerator<S> (src/lib/host/ckks/keygenerator.cu:28-1200, bfv twin): keys in the reference
    layouts from a caller-supplied seed (reproducible)."""

    def __init__(self, context, seed=0x48454F4E):
        self.context_ = context
        self.seed_ = seed
        self._n = 0

    def _next(self):
        self._n += 1
        return (self.seed_ * 0x9E3779B97F4A7C15 + self._n) & 0xFFFFFFFFFFFFFFFF

    def _buf(self, *shape):
        return torch.zeros(*shape, dtype=torch.int64, device="cuda")

    def generate_secret_key(self, sk):
        if sk.secret_key_generated_:
            raise HeonLogicError("Secretkey is already generated!")
        c = self.context_
        sk.data = self._buf(c.Q_prime_size, c.n)
        _check(lib.heon_keygen_secret(c._h, self._next(), sk.hamming_weight_, _ptr(sk.data), _stream()))
        sk.secret_key_generated_ = True
        return sk

    def generate_public_key(self, pk, sk):
        if not sk.secret_key_generated_:
            raise HeonLogicError("Secretkey is not generated!")
        c = self.context_
        pk.data = self._buf(2, c.Q_prime_size, c.n)
        _check(lib.heon_keygen_public(c._h, _ptr(sk.data), self._next(), _ptr(pk.data), _stream()))
        pk.public_key_generated_ = True
        return pk

    def generate_relin_key(self, sk):
        c = self.context_
        key = self._buf(c.digits(0), 2, c.Q_prime_size, c.n)
        _check(lib.heon_keygen_relin(c._h, _ptr(sk.data), self._next(), _ptr(key), _stream()))
        return Relinkey(c, key)

    def generate_galois_key(self, sk, shifts=None, galois_elts=None, group_order=None):
        """Keys for rotate_rows by `shifts` (default +-2^i, i < 8: MAX_SHIFT, evaluationkey.cu:306-345) plus the
        conjugation / column-rotation key (galois element 2N-1)."""
        c = self.context_
        order = group_order or (5 if c.scheme == "CKKS" else 3)
        if galois_elts is None:
            if shifts is None:
                shifts = [s * (1 << i) for i in range(8) for s in (1, -1)]
            galois_elts = [lib.heon_steps_to_galois_elt(s, c.n, order) for s in shifts]
        keys = {}
        for e in list(galois_elts) + [2 * c.n - 1]:
            if e in keys:
                continue
            k = self._buf(c.digits(0), 2, c.Q_prime_size, c.n)
            _check(lib.heon_keygen_galois(c._h, _ptr(sk.data), e, self._next(), _ptr(k), _stream()))
            keys[e] = k
        gk = Galoiskey(c, keys, conjugate_key=keys[2 * c.n - 1])
        gk.group_order_ = order
        return gk

    def generate_switch_key(self, new_sk, old_sk):
        c = self.context_
        key = self._buf(c.digits(0), 2, c.Q_prime_size, c.n)
        _check(lib.heon_keygen_switch(c._h, _ptr(new_sk.data), _ptr(old_sk.data), self._next(), _ptr(key), _stream()))
        return Switchkey(c, key)


class HEEncoder:
    """HEEncoder<CKKS> / HEEncoder<BFV> (src/lib/host/{ckks,bfv}/encoder.cu)."""

    def __init__(self, context):
        self.context_ = context
        self.slot_count_ = context.n // 2 if context.scheme == "CKKS" else context.n

    def encode(self, message, scale=None, depth=0):
        c = self.context_
        if c.scheme == "CKKS":
            z = np.ascontiguousarray(np.asarray(message, dtype=np.complex128))
            L = c.Q_size - depth
            data = torch.zeros(L, c.n, dtype=torch.int64, device="cuda")
            _check(lib.heon_ckks_encode(c._h, z.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), len(z), float(scale),
                                        depth, _ptr(data), _stream()))
            return Plaintext(c, data, depth=depth, scale=float(scale))
        m = np.ascontiguousarray(np.asarray(message, dtype=np.int64) % c.plain_modulus).astype(np.uint64)
        data = torch.zeros(c.n, dtype=torch.int64, device="cuda")
        _check(lib.heon_bfv_encode(c._h, m.ctypes.data_as(_lib.u64p), len(m), _ptr(data), _stream()))
        return Plaintext(c, data)

    def decode(self, pt, count=None):
        c = self.context_
        if c.scheme == "CKKS":
            count = self.slot_count_ if count is None else count
            out = np.zeros(count, dtype=np.complex128)
            _check(lib.heon_ckks_decode(c._h, _ptr(pt.data), pt.depth_, float(pt.scale_),
                                        out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), count, _stream()))
            return out
        count = self.slot_count_ if count is None else count
        out = np.zeros(count, dtype=np.uint64)
        _check(lib.heon_bfv_decode(c._h, _ptr(pt.data), out.ctypes.data_as(_lib.u64p), count, _stream()))
        return out


class HEEncryptor:
    """HEEncryptor<S>(context, public_key) (src/lib/host/{ckks,bfv}/encryptor.cu)."""

    def __init__(self, context, public_key, seed=0x454E43):
        self.context_, self.public_key_, self.seed_, self._n = context, public_key, seed, 0

    def encrypt(self, pt):
        c = self.context_
        self._n += 1
        data = torch.zeros(1, 2, c.Q_size, c.n, dtype=torch.int64, device="cuda")
        _check(lib.heon_encrypt(c._h, _ptr(self.public_key_.data), _ptr(pt.data), (self.seed_ << 20) + self._n, _ptr(data),
                                _stream()))
        ct = Ciphertext(c, data, depth=0, scale=getattr(pt, "scale_", 1.0))
        ct.in_ntt_domain_ = c.scheme == "CKKS"
        return ct


class HEDecryptor:
    """HEDecryptor<S>(context, secret_key) (src/lib/host/{ckks,bfv}/decryptor.cu)."""

    def __init__(self, context, secret_key):
        self.context_, self.secret_key_ = context, secret_key

    def decrypt(self, ct, index=0):
        c = self.context_
        words = ct.words()[index]
        if c.scheme == "CKKS":
            L = ct.level_count()
            data = torch.zeros(L, c.n, dtype=torch.int64, device="cuda")
            _check(lib.heon_ckks_decrypt(c._h, _ptr(self.secret_key_.data), _ptr(words.contiguous()), ct.cipher_size_, ct.depth_,
                                         _ptr(data), _stream()))
            return Plaintext(c, data, depth=ct.depth_, scale=ct.scale_)
        data = torch.zeros(c.n, dtype=torch.int64, device="cuda")
        _check(lib.heon_bfv_decrypt(c._h, _ptr(self.secret_key_.data), _ptr(words.contiguous()), ct.cipher_size_, _ptr(data),
                                    _stream()))
        return Plaintext(c, data)

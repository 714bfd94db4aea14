"""Host-side mirror of the reference's TFHE interface (Scheme::TFHE): ``HEContext`` / ``Secretkey`` /
``Bootstrappingkey`` / ``Ciphertext`` / ``HEKeyGenerator`` / ``HEEncryptor`` / ``HEDecryptor`` /
``HELogicOperator.{NAND, AND, NOR, OR, XNOR, XOR, NOT, MUX}``
(reference: src/include/heongpu/host/tfhe/*.cuh, src/lib/host/tfhe/*.cu).

PyTorch is used only for device memory and streams; every method calls the C ABI of ``libheon_b200.so``
(``heon_tfhe_*``, include/heon_b200.h)."""
import ctypes as C

import numpy as np
import torch

from .api import HeonError, _check, _ptr, _stream, lib

GATES = dict(NAND=0, AND=1, NOR=2, OR=3, XNOR=4, XOR=5, ANDNY=6, NOT=7, MUX=8)


def encode_to_torus32(mu, m_size):
    """HELogicOperator<TFHE>::encode_to_torus32 (tfhe/operator.cu:316-322)."""
    interval = ((1 << 63) // m_size) * 2
    v = ((mu * interval) >> 32) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


class HEContext:
    """HEContextImpl<Scheme::TFHE> (src/lib/host/tfhe/context.cu:23-56): the reference's single parameter set."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib.heon_tfhe_create(device, C.byref(h)))
        self._h = h
        p = (C.c_int * 7)()
        _check(lib.heon_tfhe_params(h, p))
        self.n_, self.N_, self.k_, self.bk_l_, self.bk_bg_bit_, self.ks_base_bit_, self.ks_length_ = list(p)
        self.device = device

    def __del__(self):
        if getattr(self, "_h", None):
            lib.heon_tfhe_destroy(self._h)
            self._h = None


class Secretkey:
    def __init__(self, context):
        self.context = context
        self.lwe_key_device_location_ = None
        self.tlwe_key_device_location_ = None
        self.secret_key_generated_ = False


class Bootstrappingkey:
    def __init__(self, context):
        self.context = context
        self.boot_key_device_location_ = None
        self.switch_key_device_location_a_ = None
        self.switch_key_device_location_b_ = None
        self.boot_key_generated_ = False


class Ciphertext:
    """Ciphertext<Scheme::TFHE>: `shape_` LWE samples of `n_` words."""

    def __init__(self, context, a=None, b=None):
        self.context = context
        self.a_device_location_, self.b_device_location_ = a, b
        self.n_ = a.shape[1] if a is not None else context.n_
        self.shape_ = a.shape[0] if a is not None else 0
        self.ciphertext_generated_ = a is not None


def _i32(*shape):
    return torch.empty(shape, dtype=torch.int32, device="cuda")


class HEKeyGenerator:
    def __init__(self, context, seed=0x7F4E):
        self.context_, self.seed_ = context, seed

    def generate_secret_key(self, sk):
        if sk.secret_key_generated_:
            raise HeonError("Secretkey is already generated!")
        c = self.context_
        sk.lwe_key_device_location_, sk.tlwe_key_device_location_ = _i32(c.n_), _i32(c.k_ * c.N_)
        _check(lib.heon_tfhe_keygen_secret(c._h, self.seed_, _ptr(sk.lwe_key_device_location_),
                                           _ptr(sk.tlwe_key_device_location_), _stream()))
        sk.secret_key_generated_ = True
        return sk

    def generate_bootstrapping_key(self, bk, sk):
        if not sk.secret_key_generated_:
            raise HeonError("Secretkey is not generated!")
        if bk.boot_key_generated_:
            raise HeonError("Bootkey is already generated!")
        c = self.context_
        rows = c.k_ * c.N_ * c.ks_length_ * ((1 << c.ks_base_bit_) - 1)
        bk.boot_key_device_location_ = torch.empty(c.n_, c.k_ + 1, c.bk_l_, c.k_ + 1, c.N_, dtype=torch.int64, device="cuda")
        bk.switch_key_device_location_a_, bk.switch_key_device_location_b_ = _i32(rows, c.n_), _i32(rows)
        _check(lib.heon_tfhe_keygen_boot(c._h, _ptr(sk.lwe_key_device_location_), _ptr(sk.tlwe_key_device_location_),
                                         self.seed_ + 1, _ptr(bk.boot_key_device_location_),
                                         _ptr(bk.switch_key_device_location_a_), _ptr(bk.switch_key_device_location_b_),
                                         _stream()))
        bk.boot_key_generated_ = True
        return bk


class HEEncryptor:
    def __init__(self, context, secret_key, seed=0xE4C):
        if not secret_key.secret_key_generated_:
            raise HeonError("Secretkey was not generated!")
        self.context_, self.sk_, self.seed_ = context, secret_key, seed

    def encrypt(self, bits):
        """encrypt(ciphertext, std::vector<bool>): true -> +1/8, false -> -1/8 (tfhe/encryptor.cuh)."""
        c = self.context_
        mu = encode_to_torus32(1, 8)
        msg = torch.tensor([mu if b else -mu for b in bits], dtype=torch.int32, device="cuda")
        a, b = _i32(len(bits), c.n_), _i32(len(bits))
        self.seed_ += 1
        _check(lib.heon_tfhe_encrypt(c._h, _ptr(self.sk_.lwe_key_device_location_), _ptr(msg), self.seed_, _ptr(a), _ptr(b),
                                     len(bits), _stream()))
        return Ciphertext(c, a, b)


class HEDecryptor:
    def __init__(self, context, secret_key):
        self.context_, self.sk_ = context, secret_key

    def phase(self, ct):
        out = _i32(ct.shape_)
        _check(lib.heon_tfhe_phase(self.context_._h, _ptr(self.sk_.lwe_key_device_location_), _ptr(ct.a_device_location_),
                                   _ptr(ct.b_device_location_), _ptr(out), ct.n_, ct.shape_, _stream()))
        return out.cpu().numpy()

    def decrypt(self, ct):
        return [bool(v > 0) for v in self.phase(ct)]


class HELogicOperator:
    """HELogicOperator<Scheme::TFHE> (src/include/heongpu/host/tfhe/operator.cuh:29-812)."""

    def __init__(self, context):
        self.context_ = context

    def _gate(self, gate, in1, in2, boot_key, control=None):
        c = self.context_
        for x in (in1, in2, control):
            if x is not None and x.shape_ != in1.shape_:
                raise HeonError("Both ciphertexts size should be equal!")
        out = Ciphertext(c, _i32(in1.shape_, c.n_), _i32(in1.shape_))
        p = lambda t: _ptr(t) if t is not None else None
        bk = boot_key
        _check(lib.heon_tfhe_gate(
            c._h, GATES[gate], _ptr(in1.a_device_location_), _ptr(in1.b_device_location_),
            p(in2.a_device_location_ if in2 else None), p(in2.b_device_location_ if in2 else None),
            p(control.a_device_location_ if control else None), p(control.b_device_location_ if control else None),
            _ptr(out.a_device_location_), _ptr(out.b_device_location_), p(bk.boot_key_device_location_ if bk else None),
            p(bk.switch_key_device_location_a_ if bk else None), p(bk.switch_key_device_location_b_ if bk else None),
            in1.shape_, _stream()))
        return out

    def NAND(self, a, b, bk): return self._gate("NAND", a, b, bk)
    def AND(self, a, b, bk): return self._gate("AND", a, b, bk)
    def NOR(self, a, b, bk): return self._gate("NOR", a, b, bk)
    def OR(self, a, b, bk): return self._gate("OR", a, b, bk)
    def XNOR(self, a, b, bk): return self._gate("XNOR", a, b, bk)
    def XOR(self, a, b, bk): return self._gate("XOR", a, b, bk)
    def NOT(self, a): return self._gate("NOT", a, None, None)
    def MUX(self, in1, in2, control, bk): return self._gate("MUX", in1, in2, bk, control)

    # the three steps of a gate (protected in the reference)
    def gate_linear(self, gate, in1, in2, n=None):
        c = self.context_
        n = n or in1.n_
        out = Ciphertext(c, _i32(in1.shape_, n), _i32(in1.shape_))
        _check(lib.heon_tfhe_gate_linear(c._h, GATES[gate], _ptr(in1.a_device_location_), _ptr(in1.b_device_location_),
                                         _ptr(in2.a_device_location_) if in2 else None,
                                         _ptr(in2.b_device_location_) if in2 else None, _ptr(out.a_device_location_),
                                         _ptr(out.b_device_location_), n, in1.shape_, _stream()))
        return out

    def bootstrapping(self, ct, boot_key):
        c = self.context_
        out = Ciphertext(c, _i32(ct.shape_, c.k_ * c.N_), _i32(ct.shape_))
        _check(lib.heon_tfhe_bootstrap(c._h, _ptr(ct.a_device_location_), _ptr(ct.b_device_location_), _ptr(out.a_device_location_),
                                       _ptr(out.b_device_location_), _ptr(boot_key.boot_key_device_location_), ct.shape_,
                                       _stream()))
        return out

    def key_switching(self, ct, boot_key):
        c = self.context_
        out = Ciphertext(c, _i32(ct.shape_, c.n_), _i32(ct.shape_))
        _check(lib.heon_tfhe_keyswitch(c._h, _ptr(ct.a_device_location_), _ptr(ct.b_device_location_), _ptr(out.a_device_location_),
                                       _ptr(out.b_device_location_), _ptr(boot_key.switch_key_device_location_a_),
                                       _ptr(boot_key.switch_key_device_location_b_), ct.shape_, _stream()))
        return out

    def ntt(self, data, inverse=False):
        _check(lib.heon_tfhe_ntt(self.context_._h, _ptr(data), data.numel() // 1024, int(inverse), _stream()))
        return data

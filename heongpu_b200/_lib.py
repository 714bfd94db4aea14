"""ctypes binding of the C ABI declared in include/heon_b200.h."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEON_B200_LIB") or os.path.join(_HERE, "lib", "libheon_b200.so")

u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int)
i64p = C.POINTER(C.c_longlong)
vp = C.c_void_p
ll = C.c_longlong
ci = C.c_int


class heon_info(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("scheme", "n", "log_n", "q_size", "p_size", "keyswitch_method", "device")]


# name -> (restype, argtypes); mirrors include/heon_b200.h one to one
SIGNATURES = {
    "heon_last_error": (C.c_char_p, []),
    "heon_version": (C.c_char_p, []),
    "heon_ckks_context_create": (ci, [ci, ci, i32p, ci, i32p, ci, C.POINTER(vp)]),
    "heon_ckks_context_create_values": (ci, [ci, ci, u64p, ci, u64p, ci, C.POINTER(vp)]),
    "heon_bfv_context_create": (ci, [ci, ci, i32p, ci, i32p, ci, C.c_uint64, C.POINTER(vp)]),
    "heon_bfv_context_create_values": (ci, [ci, ci, u64p, ci, u64p, ci, C.c_uint64, C.POINTER(vp)]),
    "heon_bfv_multiply": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, vp]),
    "heon_bfv_relinearize": (ci, [vp, vp, ll, vp, ci, vp]),
    "heon_bfv_apply_galois": (ci, [vp, vp, ll, vp, ll, vp, C.c_uint32, ci, vp]),
    "heon_bfv_add_plain": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, vp]),
    "heon_bfv_sub_plain": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, vp]),
    "heon_bfv_multiply_plain": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, vp]),
    "heon_bfv_keyswitch": (ci, [vp, vp, ll, vp, ll, vp, ci, vp]),
    "heon_context_destroy": (None, [vp]),
    "heon_context_info": (ci, [vp, C.POINTER(heon_info)]),
    "heon_context_table": (ci, [vp, ci, ci, u64p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "heon_steps_to_galois_elt": (ci, [ci, ci, ci]),
    "heon_ntt": (ci, [vp, vp, vp, ll, i32p, ci, ci, vp]),
    "heon_ntt_poly_ordered": (ci, [vp, vp, i64p, ci, ci, ci, vp]),
    "heon_add": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, ci, vp]),
    "heon_sub": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, ci, vp]),
    "heon_negate": (ci, [vp, vp, ll, vp, ll, ci, ci, ci, vp]),
    "heon_ckks_multiply": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, vp]),
    "heon_ckks_relinearize": (ci, [vp, vp, ll, vp, ci, ci, vp]),
    "heon_ckks_multiply_plain": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, ci, vp]),
    "heon_ckks_add_plain": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, ci, vp]),
    "heon_ckks_sub_plain": (ci, [vp, vp, ll, vp, ll, vp, ll, ci, ci, ci, vp]),
    "heon_ckks_rotate_hoisted": (ci, [vp, vp, ll, vp, ll, ll, C.POINTER(vp), C.POINTER(C.c_uint32), ci, ci, ci, vp]),
    "heon_ckks_multiply_matrix": (ci, [vp, vp, vp, vp, C.POINTER(C.c_uint32), C.POINTER(vp), ci, C.POINTER(C.c_uint32),
                                       C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), ci, ci, vp]),
    "heon_ckks_multiply_plain_accumulate": (ci, [vp, vp, vp, vp, ci, ci, vp]),
    "heon_ckks_keyswitch": (ci, [vp, vp, ll, vp, ll, vp, ci, ci, vp]),
    "heon_ckks_conjugate": (ci, [vp, vp, ll, vp, ll, vp, ci, ci, vp]),
    "heon_ckks_rescale": (ci, [vp, vp, ll, ci, ci, vp]),
    "heon_ckks_mod_drop_inplace": (ci, [vp, vp, ll, ci, ci, ci, vp]),
    "heon_ckks_mod_drop": (ci, [vp, vp, ll, vp, ll, ci, ci, vp]),
    "heon_ckks_apply_galois": (ci, [vp, vp, ll, vp, ll, vp, C.c_uint32, ci, ci, vp]),
    "heon_ckks_multiply_relinearize_host": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "heon_keygen_secret": (ci, [vp, C.c_uint64, ci, vp, vp]),
    "heon_keygen_public": (ci, [vp, vp, C.c_uint64, vp, vp]),
    "heon_keygen_relin": (ci, [vp, vp, C.c_uint64, vp, vp]),
    "heon_keygen_galois": (ci, [vp, vp, C.c_uint32, C.c_uint64, vp, vp]),
    "heon_keygen_switch": (ci, [vp, vp, vp, C.c_uint64, vp, vp]),
    "heon_encrypt": (ci, [vp, vp, vp, C.c_uint64, vp, vp]),
    "heon_ckks_decrypt": (ci, [vp, vp, vp, ci, ci, vp, vp]),
    "heon_bfv_decrypt": (ci, [vp, vp, vp, ci, vp, vp]),
    "heon_bfv_noise_budget": (ci, [vp, vp, vp, ci, i32p, vp]),
    "heon_ckks_encode": (ci, [vp, C.POINTER(C.c_double), ci, C.c_double, ci, vp, vp]),
    "heon_ckks_decode": (ci, [vp, vp, ci, C.c_double, C.POINTER(C.c_double), ci, vp]),
    "heon_bfv_encode": (ci, [vp, u64p, ci, vp, vp]),
    "heon_bfv_decode": (ci, [vp, vp, u64p, ci, vp]),
    "heon_compress_bound": (C.c_size_t, [C.c_size_t]),
    "heon_compress": (ci, [vp, C.c_size_t, vp, C.POINTER(C.c_size_t)]),
    "heon_decompress": (ci, [vp, C.c_size_t, vp, C.POINTER(C.c_size_t)]),
    "heon_tfhe_create": (ci, [ci, C.POINTER(vp)]),
    "heon_tfhe_destroy": (None, [vp]),
    "heon_tfhe_params": (ci, [vp, C.POINTER(ci)]),
    "heon_tfhe_gate": (ci, [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, vp]),
    "heon_tfhe_gate_linear": (ci, [vp, ci, vp, vp, vp, vp, vp, vp, ci, ci, vp]),
    "heon_tfhe_bootstrap": (ci, [vp, vp, vp, vp, vp, vp, ci, vp]),
    "heon_tfhe_keyswitch": (ci, [vp, vp, vp, vp, vp, vp, vp, ci, vp]),
    "heon_tfhe_keygen_secret": (ci, [vp, C.c_uint64, vp, vp, vp]),
    "heon_tfhe_keygen_boot": (ci, [vp, vp, vp, C.c_uint64, vp, vp, vp, vp]),
    "heon_tfhe_encrypt": (ci, [vp, vp, vp, C.c_uint64, vp, vp, ci, vp]),
    "heon_tfhe_phase": (ci, [vp, vp, vp, vp, vp, ci, ci, vp]),
    "heon_tfhe_ntt": (ci, [vp, vp, ci, ci, vp]),
    "heon_profile_begin": (ci, []),
    "heon_profile_end": (ci, [C.POINTER(C.c_double), i64p, ci]),
    "heon_profile_class_name": (C.c_char_p, [ci]),
    "heon_kernel_launches": (ll, [ci]),
}


def build_library(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call([os.path.join(_HERE, "build.sh")])
    return LIB_PATH


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run heongpu_b200/build.sh (or __graft_entry__.build()). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib

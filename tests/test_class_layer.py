"""The heongpu:: C++ class layer compiles with a plain host compiler (like the
reference's public headers) and, on a GPU, reproduces the C ABI word for word
with the reference's exception behaviour."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(name="class_layer_test"):
    exe = os.path.join(tempfile.mkdtemp(), name)
    lib = os.path.join(ROOT, "heongpu_b200", "lib")
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "heongpu_b200", "include"), "-I/usr/local/cuda/include",
        os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
        "-L", lib, "-lheon_b200", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{lib}", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def test_class_layer_compiles_and_links():
    exe = _build()
    out = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_class_layer_matches_c_abi_on_gpu():
    exe = _build()
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


def test_serialization_test_compiles_and_links():
    _build("serialization_test")


@pytest.mark.gpu
def test_serialization_round_trips_on_gpu():
    """save / load, heongpu::serializer (zlib) and file round trips of every object, then operators on the
    reloaded objects (example/basic/13_bfv_serialization.cpp, 14_ckks_serialization.cpp)."""
    exe = _build("serialization_test")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "round trips OK" in out.stdout, out.stdout + out.stderr


def test_bsgs_test_compiles_and_links():
    _build("bsgs_test")


@pytest.mark.gpu
def test_bsgs_matrix_products_decrypt_correctly_on_gpu():
    """multiply_matrix (single hoisting), multiply_matrix_less_memory and multiply_matrix_v2 (double hoisting in
    PQ_l) of the class layer against the plain diagonal product (ckks/operator.cu:2803-3496)."""
    exe = _build("bsgs_test")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "BSGS OK" in out.stdout, out.stdout + out.stderr


def test_host_keys_test_compiles_and_links():
    _build("host_keys_test")


@pytest.mark.gpu
def test_evaluation_keys_in_host_memory_give_identical_words_on_gpu():
    """Relinkey / Galoiskey / Switchkey store_in_host, store_in_device, key generation with storage_type::HOST and
    save() of a host-stored key (src/include/heongpu/host/{ckks,bfv}/evaluationkey.cuh)."""
    exe = _build("host_keys_test")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "host keys OK" in out.stdout, out.stdout + out.stderr

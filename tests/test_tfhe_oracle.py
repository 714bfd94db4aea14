"""CPU: the TFHE oracle (oracle/heon_oracle.c, restatement of small_ntt.cu + bootstrapping.cu) against
 (a) tests/golden/tfhe_golden.json -- SHA-256 of the outputs of the reference's own kernels captured on a B200
     (tests/golden/make_tfhe_golden.py), and
 (b) the plaintext truth table: a NAND gate evaluated end to end on the CPU with keys built here decrypts correctly
     (guards against oracle and reference being wrong together)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.tfhe_common import N_LWE, N_RING, P, digest, golden_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tfhe_golden.json")


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/tfhe_golden.json not captured yet")
    return json.load(open(GOLD))


@pytest.fixture(scope="module")
def gin():
    return golden_inputs()


def test_ntt_matches_reference_kernels(gold, gin):
    t = O.TfheOracle()
    f = t.ntt(gin["ntt_in"])
    assert digest(f)["sha256"] == gold["ntt_fwd"]["sha256"]
    b = t.ntt(f, inverse=True)
    assert digest(b)["sha256"] == gold["ntt_inv"]["sha256"]
    assert np.array_equal(b, gin["ntt_in"])


def test_gate_linear_parts_match_reference_kernels(gold, gin):
    t = O.TfheOracle()
    for gate in range(8):
        oa, ob = t.gate_linear(gate, gin["a1"], gin["b1"], gin["a2"], gin["b2"])
        assert digest(oa)["sha256"] == gold[f"gate{gate}_a"]["sha256"], gate
        assert digest(ob)["sha256"] == gold[f"gate{gate}_b"]["sha256"], gate


def test_blind_rotation_matches_reference_kernels(gold, gin):
    t = O.TfheOracle()
    oa, ob = t.bootstrap(gin["boot_a"], gin["boot_b"], gin["bk"])
    assert digest(oa)["sha256"] == gold["boot_a"]["sha256"]
    assert digest(ob)["sha256"] == gold["boot_b"]["sha256"]


def test_key_switch_matches_reference_kernel(gold, gin):
    t = O.TfheOracle()
    oa, ob = t.keyswitch(gin["ks_in_a"], gin["ks_in_b"], gin["ks_a"], gin["ks_b"])
    assert digest(oa)["sha256"] == gold["ks_a"]["sha256"]
    assert digest(ob)["sha256"] == gold["ks_b"]["sha256"]


def _torus(x):
    return np.int32(np.uint32(int(round((x - np.trunc(x)) * 2 ** 32)) & 0xFFFFFFFF))


def test_nand_gate_on_the_cpu_oracle_decrypts_correctly():
    """Keys built with numpy in the reference's layout (keygeneration.cu:1079-1440), one NAND per truth-table row."""
    t = O.TfheOracle()
    rng = np.random.default_rng(42)
    lwe = rng.integers(0, 2, N_LWE).astype(np.int64)
    tlwe = rng.integers(0, 2, N_RING).astype(np.int64)
    s_ntt = t.ntt(tlwe.astype(np.uint64)[None])[0]
    # bootstrapping key: TGSW(s_i), rows (y, z): (a, b = a*s + e) + s_i * 2^(32 - 10(z+1)) on component y
    bk = np.zeros((N_LWE, 2, 2, 2, N_RING), dtype=np.uint64)
    a_all = rng.integers(-2 ** 31, 2 ** 31, (N_LWE, 2, 2, N_RING), dtype=np.int64)
    a_res = np.where(a_all < 0, a_all + P, a_all).astype(np.uint64)
    a_ntt = t.ntt(a_res)
    prod = np.zeros_like(a_ntt)
    sN = [int(v) for v in s_ntt]
    flat_in, flat_out = a_ntt.reshape(-1, N_RING), prod.reshape(-1, N_RING)
    for r in range(flat_in.shape[0]):
        flat_out[r] = np.array([(int(x) * s) % P for x, s in zip(flat_in[r], sN)], dtype=np.uint64)
    prod = t.ntt(prod, inverse=True)
    prod_c = np.where(prod >= (P >> 1), prod.astype(np.int64) - np.int64(P), prod.astype(np.int64))
    for i in range(N_LWE):
        for y in range(2):
            for z in range(2):
                msg = (int(lwe[i]) << (32 - 10 * (z + 1))) & 0xFFFFFFFF
                a = a_all[i, y, z].copy()
                b = (prod_c[i, y, z] + rng.normal(0, 9e-9 * 0.8 * 2 ** 32, N_RING).round().astype(np.int64))
                if y == 0:
                    a[0] += msg
                else:
                    b[0] += msg
                for comp, vec in ((0, a), (1, b)):
                    v32 = (vec & 0xFFFFFFFF).astype(np.uint32).astype(np.int32).astype(np.int64)
                    bk[i, y, z, comp] = np.where(v32 < 0, v32 + P, v32).astype(np.uint64)
    bk = t.ntt(bk)
    # key-switch key
    rows = N_RING * 8 * 3
    ks_a = rng.integers(-2 ** 31, 2 ** 31, (rows, N_LWE), dtype=np.int64)
    dot = (ks_a * lwe[None, :]).sum(axis=1)
    idx = np.arange(rows)
    v, i2, i = idx % 3, (idx // 3) % 8, idx // 24
    msg = tlwe[i] * ((v + 1) << (32 - 2 * (i2 + 1)))
    noise = rng.normal(0, (1 / 32768) * 0.8 * 2 ** 32, rows).round().astype(np.int64)
    ks_b = ((dot + msg + noise) & 0xFFFFFFFF).astype(np.uint32).astype(np.int32)
    ks_a = (ks_a & 0xFFFFFFFF).astype(np.uint32).astype(np.int32)
    # the four input pairs
    mu = 1 << 29
    bits1, bits2 = np.array([0, 0, 1, 1]), np.array([0, 1, 0, 1])

    def enc(bits):
        a = rng.integers(-2 ** 31, 2 ** 31, (len(bits), N_LWE), dtype=np.int64)
        e = rng.normal(0, (1 / 32768) * 0.8 * 2 ** 32, len(bits)).round().astype(np.int64)
        b = (a * lwe[None, :]).sum(axis=1) + np.where(bits == 1, mu, -mu) + e
        return (a & 0xFFFFFFFF).astype(np.uint32).astype(np.int32), (b & 0xFFFFFFFF).astype(np.uint32).astype(np.int32)

    a1, b1 = enc(bits1)
    a2, b2 = enc(bits2)
    la, lb = t.gate_linear(0, a1, b1, a2, b2)
    ea, eb = t.bootstrap(la, lb, bk)
    oa, ob = t.keyswitch(ea, eb, ks_a, ks_b)
    phase = (ob.astype(np.int64) - (oa.astype(np.int64) * lwe[None, :]).sum(axis=1)) & 0xFFFFFFFF
    phase = np.where(phase >= 2 ** 31, phase - 2 ** 32, phase)
    assert list(phase > 0) == list(~((bits1 & bits2).astype(bool)))
    assert np.all(np.abs(np.abs(phase) - mu) < 2 ** 27), "the bootstrapped phase sits at +-1/8 with small noise"

"""Seeded TFHE inputs shared by the golden generator (GPU box) and the CPU oracle tests: numpy Generator streams are
portable, so both sides rebuild identical buffers from the seeds."""
import hashlib

import numpy as np

P = 1152921504606877697
N_LWE, N_RING = 512, 1024


def digest(a):
    a = np.ascontiguousarray(a)
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "shape": list(a.shape), "dtype": str(a.dtype),
            "head": [int(v) for v in a.reshape(-1)[:6]]}


def i32(rng, *shape):
    return rng.integers(-2 ** 31, 2 ** 31, shape, dtype=np.int64).astype(np.int32)


def golden_inputs():
    rng = np.random.default_rng(20261017)
    g = {}
    g["ntt_in"] = rng.integers(0, P, (3, N_RING), dtype=np.uint64)
    g["ntt_in"][1] = P - 1
    g["a1"], g["b1"] = i32(rng, 3, N_LWE), i32(rng, 3)
    g["a2"], g["b2"] = i32(rng, 3, N_LWE), i32(rng, 3)
    g["boot_a"], g["boot_b"] = i32(rng, 2, N_LWE), i32(rng, 2)
    g["boot_a"][0, :5] = 0          # steps with rotation 0
    g["boot_a"][1, 7] = -2 ** 31    # rotation by N
    g["bk"] = rng.integers(0, P, (N_LWE, 2, 2, 2, N_RING), dtype=np.uint64)
    g["ks_in_a"], g["ks_in_b"] = i32(rng, 2, N_RING), i32(rng, 2)
    g["ks_a"], g["ks_b"] = i32(rng, N_RING * 8 * 3, N_LWE), i32(rng, N_RING * 8 * 3)
    return g

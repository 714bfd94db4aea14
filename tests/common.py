"""Shared fixtures: deterministic inputs and parameter sets.

Inputs are uniform canonical residues from splitmix64 with seed
0x48454F4E00000000 + buffer index."""
import functools

import numpy as np

SEED0 = 0x48454F4E00000000
M64 = (1 << 64) - 1

# name -> (log_n, q_bits, p_bits)
PARAMS = {
    "n12_I": (12, [40, 30, 30], [40]),  # test_ckks_multiplication.cpp:44
    "n12_II": (12, [40, 30, 30, 30], [40, 40]),
    "n13_I": (13, [40, 30, 30, 30, 30], [40]),
    "n13_II": (13, [50, 40, 40, 40, 40], [50, 50, 50]),
    "n14_C2": (14, [50, 40, 40, 40], [48]),  # BASELINE config 2 (logq ~ 218)
    "n15_II": (15, [59, 50, 50, 50, 50, 50, 50], [59, 59]),
    "n16_I_small": (16, [59, 45, 45, 45, 45], [59]),
    "n16_II_small": (16, [60, 50, 50, 50, 50, 50, 50], [60, 60, 60]),
    "C3_I": (16, [59] + [45] * 36, [59]),  # test_ckks_multiplication.cpp:356-360
    "C3_II": (16, [60] + [50] * 30, [60, 60, 60]),  # 1_ckks_regular_bootstrapping.cpp:18-21
    "mixed": (12, [60, 30, 30, 30], [60]),
}


def splitmix64(seed, count):
    """count 64-bit outputs of splitmix64 started at `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & M64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def residues(buf_index, primes_per_limb, n, lead=()):
    """Uniform canonical residues, shape lead + (len(primes_per_limb), n)."""
    limbs = len(primes_per_limb)
    total = int(np.prod(lead, dtype=np.int64)) if lead else 1
    raw = splitmix64(SEED0 + buf_index, total * limbs * n).reshape(total, limbs, n)
    p = np.array(primes_per_limb, dtype=np.uint64).reshape(1, limbs, 1)
    return (raw % p).reshape(tuple(lead) + (limbs, n))


def ciphertext(buf_index, primes, L, n, comps=2, batch=None):
    lead = (comps,) if batch is None else (batch, comps)
    return residues(buf_index, primes[:L], n, lead)


def eval_key(buf_index, primes, d, n):
    """[d][2][Q'_0][N] uniform key words (rk1 is uniform in the reference too,
    keygeneration.cu:145-185; throughput and parity do not depend on rk0's structure)."""
    return residues(buf_index, primes, n, (d, 2))


@functools.lru_cache(maxsize=None)
def oracle_ctx(name):
    from oracle import oracle as O
    log_n, qb, pb = PARAMS[name]
    primes = O.generate_primes(1 << log_n, qb + pb)
    return O.OracleContext(log_n, primes, len(qb), len(pb))

"""Generate golden vectors from the REFERENCE's own CUDA kernels (oracle/_ref/libref_gpu.so).

Run on a GPU box:  python tests/golden/make_golden.py gpurun_out/golden_ref_kernels.json
then copy the JSON to tests/golden/.  Inputs are the seeded splitmix64 residues of
tests/common.py; for every operator output we store the SHA-256 of the little-endian
words plus the first 8 words.  tests/test_golden.py replays the same operators on the
CPU oracle and compares -- this pins the oracle to the reference kernels without a GPU."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def digest(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "shape": list(a.shape),
            "head": [int(v) for v in a.reshape(-1)[:8]]}


CKKS_SETS = ["n12_I", "n12_II", "n13_II", "mixed"]
BFV_SETS = ["bfv_n12_I", "bfv_n12_II", "bfv_n13_I"]


def ckks_inputs(name, depth):
    from tests.common import ciphertext, eval_key, oracle_ctx
    oc = oracle_ctx(name)
    L = oc.Q - depth
    a = ciphertext(200, oc.primes, L, oc.n)
    b = ciphertext(201, oc.primes, L, oc.n)
    key = eval_key(202, oc.primes, oc.digits(0), oc.n)
    return oc, L, a, b, key


def bfv_inputs(name):
    from tests.common import residues
    from tests.test_bfv_decrypt_level import bfv_oracle
    ob, oc = bfv_oracle(name)
    a = residues(210, ob.primes[: ob.Q], ob.n, (2,))
    b = residues(211, ob.primes[: ob.Q], ob.n, (2,))
    key = residues(212, ob.primes, ob.n, (ob.digits(), 2))
    return ob, a, b, key


def main(out_path):
    import torch
    from oracle import ref as R
    from tests.gpu_common import to_dev, to_host
    gold = {"source": "reference CUDA kernels (oracle/_ref/libref_gpu.so, sm_100a build of the unmodified sources)",
            "gpu": torch.cuda.get_device_name(0), "cases": {}}
    for name in CKKS_SETS:
        for depth in (0, 1):
            oc, L, a, b, key = ckks_inputs(name, depth)
            t = R.tables_for_refgpu(oc.n_power, oc.primes, oc.Q, oc.K)
            rg = R.RefGpu(oc.n_power, oc.primes, oc.Q, oc.K, t)
            dkey = to_dev(key)
            rc = torch.zeros(3, L, oc.n, dtype=torch.int64, device="cuda")
            rg.multiply(to_dev(a), to_dev(b), rc, depth)
            case = {"multiply": digest(to_host(rc))}
            rg.relinearize(rc, dkey, depth)
            case["relinearize"] = digest(to_host(rc))
            if L >= 2:
                rg.rescale(rc, depth)
                case["rescale"] = digest(to_host(rc).reshape(-1)[: 2 * (L - 1) * oc.n])
            ro = torch.zeros(2, L, oc.n, dtype=torch.int64, device="cuda")
            rg.apply_galois(to_dev(a), ro, dkey, 5, depth)
            case["apply_galois_5"] = digest(to_host(ro))
            x = to_dev(a)
            rg.ntt_level(x, depth, inverse=True) if False else None
            gold["cases"][f"{name}/depth{depth}"] = case
            del rg
    for name in BFV_SETS:
        ob, a, b, key = bfv_inputs(name)
        t = R.tables_for_refgpu(ob.n_power, ob.primes, ob.Q, ob.K)
        rg, rb = R.RefGpu(ob.n_power, ob.primes, ob.Q, ob.K, t), R.RefBfv(ob)
        dkey = to_dev(key)
        rc = torch.zeros(3, ob.Q, ob.n, dtype=torch.int64, device="cuda")
        rb.multiply(to_dev(a), to_dev(b), rc)
        case = {"multiply": digest(to_host(rc))}
        R.bfv_relinearize(rg, rc, dkey)
        case["relinearize"] = digest(to_host(rc)[:2])
        ro = torch.zeros(2, ob.Q, ob.n, dtype=torch.int64, device="cuda")
        R.bfv_apply_galois(rg, to_dev(a), ro, dkey, 3)
        case["apply_galois_3"] = digest(to_host(ro))
        gold["cases"][name] = case
    torch.cuda.synchronize()
    with open(out_path, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", out_path, len(gold["cases"]), "cases")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_ref_kernels.json"))

"""Golden vectors for TFHE from the REFERENCE's own kernels (oracle/_ref/libref_tfhe.so: small_ntt.cu +
bootstrapping.cu compiled unmodified, launch replay of src/lib/host/tfhe/operator.cu).

Run on a GPU box:  python tests/golden/make_tfhe_golden.py gpurun_out/tfhe_golden.json
then copy the JSON to tests/golden/.  Inputs come from tests/tfhe_common.py (seeded numpy streams); for every
output the SHA-256 of the raw little-endian words is stored.  tests/test_tfhe_oracle.py replays the same
operators on the CPU oracle and compares -- this pins the oracle to the reference kernels without a GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ref as R  # noqa: E402
from tests.tfhe_common import digest, golden_inputs  # noqa: E402


def main(out_path):
    rt = R.RefTfhe()
    g = golden_inputs()
    dev = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda()
    host = lambda t, dt: t.cpu().numpy().view(dt) if dt == np.uint64 else t.cpu().numpy()
    out = {}
    x = dev(g["ntt_in"].copy())
    rt.ntt(x)
    out["ntt_fwd"] = digest(host(x, np.uint64))
    rt.ntt(x, inverse=True)
    out["ntt_inv"] = digest(host(x, np.uint64))
    a1, b1, a2, b2 = (dev(g[k]) for k in ("a1", "b1", "a2", "b2"))
    for gate in range(8):
        oa, ob = torch.zeros_like(a1), torch.zeros_like(b1)
        rt.gate_linear(gate, a1, b1, a2, b2, oa, ob, a1.shape[1], a1.shape[0])
        out[f"gate{gate}_a"], out[f"gate{gate}_b"] = digest(host(oa, np.int32)), digest(host(ob, np.int32))
    ba, bb, bk = dev(g["boot_a"]), dev(g["boot_b"]), dev(g["bk"])
    oa = torch.zeros(2, 1024, dtype=torch.int32, device="cuda")
    ob = torch.zeros(2, dtype=torch.int32, device="cuda")
    rt.bootstrap(ba, bb, oa, ob, bk, 2)
    out["boot_a"], out["boot_b"] = digest(host(oa, np.int32)), digest(host(ob, np.int32))
    ka, kb = dev(g["ks_in_a"]), dev(g["ks_in_b"])
    oa = torch.zeros(2, 512, dtype=torch.int32, device="cuda")
    ob = torch.zeros(2, dtype=torch.int32, device="cuda")
    rt.keyswitch(ka, kb, oa, ob, dev(g["ks_a"]), dev(g["ks_b"]), 2)
    torch.cuda.synchronize()
    out["ks_a"], out["ks_b"] = digest(host(oa, np.int32)), digest(host(ob, np.int32))
    out["_source"] = "reference kernels (small_ntt.cu, bootstrapping.cu @ /root/reference) on " + torch.cuda.get_device_name(0)
    json.dump(out, open(out_path, "w"), indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tfhe_golden.json")

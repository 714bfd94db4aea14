"""Pinned-memory PCIe ceiling of this box: H2D alone, D2H alone and both directions at once, per GPU,
with 1..N GPUs driven concurrently (one process per GPU under torchrun, or a single process).

    python tools/pcie_ceiling.py                       # GPU 0 alone
    python -m torch.distributed.run --nproc-per-node N tools/pcie_ceiling.py

Prints one JSON line per run (rank 0): per-rank GB/s and the aggregate.  The e2e number of bench.py
moves 97.5 MB per op across this link (C3_II: two input ciphertexts in, one result out); its ceiling
is  min(h2d_both / 65.0 MB, d2h_both / 32.5 MB)  ops/s per GPU."""
import json
import os
import sys
import time

import torch

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

MB = 512
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=12):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * (MB << 20) / dt / 1e9


res = {}
for name, (a, b) in {"h2d_alone": (True, False), "d2h_alone": (False, True), "both": (True, True)}.items():
    run(a, b, 3)
    res[name] = run(a, b)
vals = torch.tensor([res["h2d_alone"], res["d2h_alone"], res["both"]], device="cuda", dtype=torch.float64)
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    per = [[float(x) for x in v.cpu()] for v in allv]
    both = [p[2] for p in per]
    out = {"n_gpus": world, "buffer_mb": MB,
           "per_gpu_gbs": [{"h2d_alone": p[0], "d2h_alone": p[1], "each_direction_when_both": p[2]} for p in per],
           "aggregate_each_direction_when_both_gbs": sum(both),
           "c3_ii_e2e_ceiling_ops_per_s_per_gpu": min(both) * 1e9 / 65.0e6,
           # whole box: every GPU limited by its own link (65 MB in per op) AND all of them by the host's aggregate
           # pinned-copy bandwidth (65 MB in + 32.5 MB out per op over the sum of both directions)
           "c3_ii_e2e_ceiling_ops_per_s_box": min(sum(both) * 1e9 / 65.0e6, 2 * sum(both) * 1e9 / 97.5e6),
           "note": "both: H2D and D2H streams run concurrently, GB/s quoted per direction"}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python
"""bench.py -- CKKS N=2^16 multiply+relinearize throughput on B200 (BASELINE.json metric).

A "step" = multiply + relinearize_inplace on a batch of B independent
ciphertext pairs (B ops).  `value` = ops/s with inputs resident in HBM;
`e2e` = the same through the public API with pinned HOST buffers (H2D of both
inputs and D2H of the result inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3_II|C3_I|n14_C2|M4_bfv_rot]
                    [--batch B] [--impl ours|reference]

Workloads: C3_II (default, the configuration the BASELINE metric is quoted on) and C3_I are BASELINE
config 3; n14_C2 is config 2 (CKKS N=2^14 L=4, batch 1024, multiply+relinearize+rescale); M4_bfv_rot
is config 4 (BFV N=2^15, default 128-bit modulus, rotate_rows over the steps +-2^0..2^7).

Multi-GPU: one process per GPU (torchrun), ciphertext batches are sharded, keys
and tables replicated, no collective on the data path (weak scaling).

--impl reference times the reference's OWN CUDA kernels (oracle/_ref/libref_gpu.so,
compiled unmodified from the reference sources; the reference has no CPU path)
driven by the launch replay in oracle/ref_gpu_harness.cu, sequentially over
the same batch on one stream, as the reference's operator would.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (params key, description)
    "C3_II": "CKKS N=2^16 {60,50x30}/{60,60,60} L=31 K=3 (Method II, bootstrapping params) mul+relin depth 0",
    "C3_I": "CKKS N=2^16 {59,45x36}/{59} L=37 K=1 (Method I) mul+relin depth 0",
    "n14_C2": "CKKS N=2^14 {50,40,40,40}/{48} L=4 K=1 (logq~218) mul+relin+rescale depth 0 (BASELINE config 2)",
    "M4_bfv_rot": "BFV N=2^15 default 128-bit modulus (14+1 primes, defaultmodulus.cpp:34-51) rotate_rows sweep over steps +-2^0..2^7 (BASELINE config 4)",
}
DEFAULT_BATCH = {"C3_II": 16, "C3_I": 8, "n14_C2": 1024, "M4_bfv_rot": 64}
# src/lib/util/defaultmodulus.cpp:34-51 (N = 32768, 128-bit security): the last prime is P
BFV_32768_MODULUS = [0x2000000002b0001, 0x2000000003a0001, 0x2000000005b0001, 0x200000000640001, 0x400000000270001,
                     0x400000000350001, 0x400000000360001, 0x4000000004d0001, 0x400000000570001, 0x400000000660001,
                     0x4000000008a0001, 0x400000000920001, 0x400000000980001, 0x400000000990001, 0x400000000a40001]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_per_launch_pair(limb_polys):
    """DRAM bytes (read+write) of one forward-NTT launch pair, from the committed ncu --set full
    capture (profiles/r1_traffic.json), scaled to the limb-polynomials one launch pair processes."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path))["fwd_ntt_dram_bytes_per_limb_poly"] * limb_polys


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank, world, local = env_rank_world()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


from heongpu_b200.sharding import barrier, env_rank_world  # noqa: E402
from heongpu_b200 import sharding as _sh  # noqa: E402


def max_over_ranks(x, world):
    return _sh.max_over_ranks(x, world, device="cuda")


def timed(fn, steps, world):
    """barrier + sync, K steps between CUDA events on the current stream, sync + barrier; max over ranks."""
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


def make_inputs(name, batch, rank, primes=None):
    """Synthetic ciphertexts and key.  `primes`: the modulus chain of the engine under test (our arm
    passes its own context's primes; only the reference arm asks the oracle for the chain)."""
    from tests.common import PARAMS, SEED0  # parameter sets + the documented seed only
    log_n, qb, pb = PARAMS[name]
    n, Q, K = 1 << log_n, len(qb), len(pb)
    if primes is None:
        from oracle import oracle as O  # reference arm
        primes = O.generate_primes(n, qb + pb)
    # uniform canonical residues generated ON DEVICE from a counter-based mix of the documented seed
    def dev_residues(buf, shape_lead, plist):
        g = torch.Generator(device="cuda")
        g.manual_seed((SEED0 + buf + 1000 * rank) & 0x7FFFFFFFFFFFFFFF)
        p = torch.tensor(plist, dtype=torch.int64, device="cuda").view(*([1] * len(shape_lead)), len(plist), 1)
        r = torch.randint(0, 1 << 62, (*shape_lead, len(plist), n), dtype=torch.int64, device="cuda", generator=g)
        return r % p
    a = dev_residues(1, (batch, 2), primes[:Q])
    b = dev_residues(2, (batch, 2), primes[:Q])
    d = Q if K == 1 else -(-Q // K)
    key = dev_residues(3, (d, 2), primes)
    return dict(log_n=log_n, n=n, Q=Q, K=K, d=d, primes=primes, qb=qb, pb=pb, a=a, b=b, key=key)


def algorithmic_bytes_per_op(inp):
    # mul+relin = (2*d*Q' + 12*L) * N * 8 bytes
    L, Qp = inp["Q"], inp["Q"] + inp["K"]
    return (2 * inp["d"] * Qp + 12 * L) * inp["n"] * 8


def run_ours(args, rank, world, local):
    from heongpu_b200 import api
    from tests.common import PARAMS
    log_n, qb, pb = PARAMS[args.workload]
    ctx = api.HEContext(log_n, qb, pb, device=local)
    inp = make_inputs(args.workload, args.batch, rank, primes=ctx.primes)
    B, L, n = args.batch, inp["Q"], inp["n"]
    op = api.HEArithmeticOperator(ctx)
    A, Bc = api.Ciphertext(ctx, inp["a"]), api.Ciphertext(ctx, inp["b"])
    out = torch.zeros(B, 3, L, n, dtype=torch.int64, device="cuda")
    rk = api.Relinkey(ctx, inp["key"])

    with_rescale = args.workload == "n14_C2"

    def step():
        Cc = api.Ciphertext(ctx, out)
        op.multiply(A, Bc, Cc)
        op.relinearize_inplace(Cc, rk)
        if with_rescale:
            op.rescale_inplace(Cc)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    api.lib.heon_kernel_launches(1)
    ms = timed(step, args.steps, world)
    launches = api.lib.heon_kernel_launches(0)
    clocks = sampler.stop() if rank == 0 else None
    value = B * args.steps * world / (ms * 1e-3)

    # ---- e2e: pinned host buffers; every step copies both inputs H2D and the result D2H.
    # The copies of neighbouring steps overlap the compute (three streams, two device
    # buffer sets) -- all of it inside the timed region.
    ha, hb = inp["a"].cpu().pin_memory(), inp["b"].cpu().pin_memory()
    hres = [torch.empty(B, 2, L, n, dtype=torch.int64).pin_memory() for _ in range(2)]
    da = [torch.empty_like(inp["a"]) for _ in range(2)]
    db = [torch.empty_like(inp["b"]) for _ in range(2)]
    outs = [out, torch.zeros_like(out)]
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def run_e2e(steps):
        ev_in = [None, None]
        ev_cmp = [None, None]
        ev_out = [None, None]
        main = torch.cuda.current_stream()
        for st in (s_in, s_cmp, s_out):
            st.wait_stream(main)
        for i in range(steps):
            k = i & 1
            with torch.cuda.stream(s_in):
                if ev_cmp[k] is not None:
                    s_in.wait_event(ev_cmp[k])  # inputs of step i-2 consumed
                da[k].copy_(ha, non_blocking=True)
                db[k].copy_(hb, non_blocking=True)
                ev_in[k] = s_in.record_event()
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[k])
                if ev_out[k] is not None:
                    s_cmp.wait_event(ev_out[k])  # result buffer of step i-2 drained
                Cc = api.Ciphertext(ctx, outs[k])
                op.multiply(api.Ciphertext(ctx, da[k]), api.Ciphertext(ctx, db[k]), Cc)
                op.relinearize_inplace(Cc, rk)
                if with_rescale:
                    op.rescale_inplace(Cc)
                ev_cmp[k] = s_cmp.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[k])
                hres[k].copy_(outs[k][:, :2], non_blocking=True)
                ev_out[k] = s_out.record_event()
        for st in (s_in, s_cmp, s_out):
            main.wait_stream(st)

    run_e2e(max(2, args.warmup))
    e2e_steps = max(4, args.steps)
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(e2e_steps)
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms_e2e = max_over_ranks(e0.elapsed_time(e1), world)
    e2e_value = B * e2e_steps * world / (ms_e2e * 1e-3)

    # ---- per-kernel CUDA-event pass (separate, instrumented run of the same steps) ----
    kernels, roof = [], None
    if rank == 0:
        peak, peak_src = peaks()
        api.lib.heon_profile_begin()
        psteps = max(2, min(args.steps, 5))
        for _ in range(psteps):
            step()
        msv = (C.c_double * 16)()
        cnt = (C.c_longlong * 16)()
        ncls = api.lib.heon_profile_end(msv, cnt, 16)
        Qp = inp["Q"] + inp["K"]
        K = inp["K"]
        polys = {  # limb-polynomials transformed per op by each NTT class
            # Method I: d*Q' mod-up digits + 2L corrections; Method II: the digits' own limbs are not
            # transformed (they are the input's NTT words), the mod-down corrections add 2L
            "fwd": inp["d"] * Qp + 2 * L - (0 if K == 1 else L),
            # INTT: c2 (L) + the 2K special limbs of the accumulator (NTT-domain mod-down)
            "inv": L + 2 * K}
        tot = sum(msv[i] for i in range(ncls))
        for i in range(ncls):
            if cnt[i] == 0:
                continue
            name = api.lib.heon_profile_class_name(i).decode()
            per_op_ms = msv[i] / (psteps * B)
            alg = None
            if name.startswith("ntt_fwd"):
                alg = polys["fwd"] * n * 16 / 2  # each pass carries half of the transform's 16 B/coeff
            elif name.startswith("ntt_inv"):
                alg = polys["inv"] * n * 16 / 2
            elif name == "keyswitch_mac":
                alg = (3 * inp["d"] * Qp + 2 * Qp) * n * 8
            elif name == "cross_multiply":
                alg = 7 * L * n * 8
            kernels.append({"kernel": name, "launches_per_step": cnt[i] / psteps, "ms_per_op": per_op_ms,
                            "share": msv[i] / tot if tot else None,
                            "alg_gbs": (alg / (per_op_ms * 1e-3) / 1e9) if alg else None})
        # dominant unit: the forward NTT (column pass + row pass), 16*N bytes per limb-polynomial
        fwd_ms = sum(k["ms_per_op"] for k in kernels if k["kernel"].startswith("ntt_fwd"))
        fwd_launch = sum(k["launches_per_step"] for k in kernels if k["kernel"].startswith("ntt_fwd"))
        if fwd_ms > 0:
            achieved = polys["fwd"] * n * 16 / (fwd_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "forward NTT (ntt_fwd_col_pass + ntt_fwd_row_pass)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "traffic": traffic_per_launch_pair(polys["fwd"] * B),
                    "launches_per_step": fwd_launch,
                    "note": "algorithmic bytes = 16*N per limb-polynomial (SURVEY 8(d)); exact 64-bit modmul makes this kernel arithmetic-bound (FP64 pipe, DESIGN.md 4.4): ceiling ~0.5-0.6 of the HBM roofline"}
    res = dict(value=value, ms=ms, launches=launches, clocks=clocks, e2e=e2e_value,
               h2d=int(ha.numel() * 8 * 2), d2h=int(hres[0].numel() * 8), kernels=kernels, roof=roof, inp=inp)
    return res


# ---------------------------------------------------------------------------------------------
# BASELINE config 4: BFV N=2^15 Galois-rotate key-switch sweep (all power-of-2 steps)
# ---------------------------------------------------------------------------------------------
ROT_STEPS = [s * (1 << i) for i in range(8) for s in (1, -1)]  # +-2^0..2^7 (MAX_SHIFT = 8, bfv/evaluationkey.cu:306-345)


def galois_elt(steps, n, group_order):
    """steps_to_galois_elt (src/lib/kernel/keygeneration.cu:684-727)"""
    m = 2 * n
    pos = abs(steps)
    s_ = (n >> 1) - pos if steps < 0 else pos
    return pow(group_order, s_, m)


def rot_inputs(batch, rank):
    from tests.common import SEED0
    primes, n = BFV_32768_MODULUS, 1 << 15
    Q = len(primes) - 1

    def dev_residues(buf, lead, plist):
        g = torch.Generator(device="cuda")
        g.manual_seed((SEED0 + buf + 1000 * rank) & 0x7FFFFFFFFFFFFFFF)
        p = torch.tensor(plist, dtype=torch.int64, device="cuda").view(*([1] * len(lead)), len(plist), 1)
        r = torch.randint(0, 1 << 62, (*lead, len(plist), n), dtype=torch.int64, device="cuda", generator=g)
        return r % p
    a = dev_residues(1, (batch, 2), primes[:Q])
    keys = [dev_residues(10 + i, (Q, 2), primes) for i in range(len(ROT_STEPS))]  # d = Q digits (Method I)
    return dict(n=n, Q=Q, primes=primes, a=a, keys=keys)


def run_rot(args, rank, world, local, reference):
    inp = rot_inputs(args.batch, rank)
    B, Q, n = args.batch, inp["Q"], inp["n"]
    out = torch.zeros(B, 2, Q, n, dtype=torch.int64, device="cuda")
    if reference:
        from oracle import ref as R
        if not R.have_gpu():
            return None
        t = R.tables_for_refgpu(15, inp["primes"], Q, 1)
        rg = R.RefGpu(15, inp["primes"], Q, 1, t)
        elts = [galois_elt(s, n, 3) for s in ROT_STEPS]

        def step():
            for e, k in zip(elts, inp["keys"]):
                for i in range(B):  # the reference has no batch dimension
                    R.bfv_apply_galois(rg, inp["a"][i], out[i], k, e)
        launches = None
    else:
        from heongpu_b200 import api
        ctx = api.HEContext(15, q_values=inp["primes"][:Q], p_values=inp["primes"][Q:], plain_modulus=786433, device=local)
        op = api.HEArithmeticOperator(ctx)
        A = api.Ciphertext(ctx, inp["a"])
        A.in_ntt_domain_ = False
        elts = [api.lib.heon_steps_to_galois_elt(s, n, 3) for s in ROT_STEPS]
        gk = api.Galoiskey(ctx, dict(zip(elts, inp["keys"])))
        O_ = api.Ciphertext(ctx, out)

        def step():
            for s in ROT_STEPS:
                op.rotate_rows_bfv(A, O_, gk, s)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if not reference:
        api.lib.heon_kernel_launches(1)
    ms = timed(step, args.steps, world)
    launches = None if reference else int(api.lib.heon_kernel_launches(0))
    clocks = sampler.stop() if rank == 0 else None
    ops = B * len(ROT_STEPS)
    res = dict(value=ops * args.steps * world / (ms * 1e-3), ms=ms, launches=launches, clocks=clocks)
    if reference:
        return res
    # e2e: inputs from pinned host memory, the sweep, the last rotation's result back to the host
    ha = inp["a"].cpu().pin_memory()
    hres = torch.empty(B, 2, Q, n, dtype=torch.int64).pin_memory()

    def e2e_step():
        inp["a"].copy_(ha, non_blocking=True)
        step()
        hres.copy_(out, non_blocking=True)
    e2e_step()
    ms_e = timed(e2e_step, max(2, args.steps // 2), world)
    res.update(e2e=ops * max(2, args.steps // 2) * world / (ms_e * 1e-3), h2d=int(ha.numel() * 8), d2h=int(hres.numel() * 8))
    # per-kernel CUDA-event pass
    kernels = []
    if rank == 0:
        api.lib.heon_profile_begin()
        step()
        msv, cnt = (C.c_double * 16)(), (C.c_longlong * 16)()
        ncls = api.lib.heon_profile_end(msv, cnt, 16)
        tot = sum(msv[i] for i in range(ncls))
        Qp = Q + 1
        fwd_polys = Q * Qp  # per op: mod-up fused into the forward NTT of d*Q' limb-polynomials
        for i in range(ncls):
            if cnt[i]:
                name = api.lib.heon_profile_class_name(i).decode()
                kernels.append({"kernel": name, "launches_per_step": int(cnt[i]), "ms_per_op": msv[i] / ops,
                                "share": msv[i] / tot if tot else None})
        fwd_ms = sum(k["ms_per_op"] for k in kernels if k["kernel"].startswith("ntt_fwd"))
        peak, peak_src = peaks()
        if fwd_ms > 0:
            achieved = fwd_polys * n * 16 / (fwd_ms * 1e-3) / 1e9
            res["roof"] = {"bound": "hbm", "kernel": "forward NTT (mod-up fused)", "achieved": achieved, "peak": peak,
                           "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                           "note": "58/59-bit primes: integer butterflies (the FP64 path needs p < 2^50)"}
    res["kernels"] = kernels
    return res


def cpu_baseline(inp, workload):
    """The CPU oracle (port of the reference algorithm) on a bounded sample: one
    multiply+relinearize of the same workload with all host threads (OpenMP)."""
    from oracle import oracle as O
    from tests.common import ciphertext, eval_key
    L, n = inp["Q"], inp["n"]
    oc = O.OracleContext(inp["log_n"], inp["primes"], inp["Q"], inp["K"])
    a = ciphertext(1, inp["primes"], L, n)
    b = ciphertext(2, inp["primes"], L, n)
    key = eval_key(3, inp["primes"], inp["d"], n)
    oc.relinearize(oc.multiply(a, b), key)  # warm-up (page faults, OpenMP pool)
    t0, ops = time.time(), 0
    while ops < 64 and (time.time() - t0 < 12.0 or ops < 2):  # a bounded sample: ~12 s of CPU work
        m = oc.multiply(a, b)
        oc.relinearize(m, key)
        ops += 1
    dt = time.time() - t0
    return {"value": ops / dt, "unit": "ops/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{ops} multiply+relinearize of {workload} on the CPU oracle (OpenMP on all host threads, {dt:.1f} s)"}


def run_reference(args, rank, world, local):
    from oracle import ref as R
    if not R.have_gpu():
        return None
    inp = make_inputs(args.workload, args.batch, rank)
    B, L, n = args.batch, inp["Q"], inp["n"]
    t = R.tables_for_refgpu(inp["log_n"], inp["primes"], inp["Q"], inp["K"])
    rg = R.RefGpu(inp["log_n"], inp["primes"], inp["Q"], inp["K"], t)
    out = torch.zeros(B, 3, L, n, dtype=torch.int64, device="cuda")

    def step():
        for i in range(B):  # the reference has no batch dimension: B sequential ops on one stream
            rg.multiply(inp["a"][i], inp["b"][i], out[i], 0)
            rg.relinearize(out[i], inp["key"], 0)
            if args.workload == "n14_C2":
                rg.rescale(out[i], 0)

    for _ in range(args.warmup):
        step()
    ms = timed(step, args.steps, world)
    return dict(value=B * args.steps * world / (ms * 1e-3), ms=ms, inp=inp)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3_II", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.batch <= 0:
        args.batch = DEFAULT_BATCH[args.workload]
    if args.impl == "reference":
        # the reference is single-GPU: under torchrun rank 0 alone runs it (no process group is
        # created, so the other ranks can leave at once without a collective to wait for)
        rank, world, local = env_rank_world()
        if rank != 0:
            return
        torch.cuda.set_device(local)
        rank, world = 0, 1
    else:
        rank, world, local = dist_setup(args.gpus)

    config = {"workload": WORKLOADS[args.workload], "workload_key": args.workload,
              "batch_per_gpu": args.batch, "ops_per_step": args.batch * world,
              "l2_policy": "inputs larger than L2 (batch of ciphertext pairs + evaluation key >> 126 MB)",
              "parallelism": f"batch sharded over {world} GPU(s), keys replicated, no collective"}
    base = {"metric": "CKKS N=2^16 mul+relin ops/sec", "unit": "ops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": config}
    if args.workload == "n14_C2":
        base["metric"] = "CKKS N=2^14 mul+relin ops/sec"

    if args.workload == "M4_bfv_rot":
        base["metric"] = "BFV N=2^15 rotate_rows (Galois key-switch) ops/sec"
        config["ops_per_step"] = args.batch * world * len(ROT_STEPS)
        if args.impl == "reference":
            if world > 1 and rank != 0:
                return
            r = run_rot(args, 0, 1, local, True)
            if r is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_gpu.so not built"}))
                return
            line = dict(base)
            line.update({"impl": "reference", "value": r["value"], "n_gpus": 1, "ms_per_step": r["ms"] / args.steps,
                         "cpu_baseline": {"value": r["value"], "unit": "ops/s", "cores": 0, "kind": "reference",
                                          "sample": "the reference's own CUDA kernels (sm_100a build) on one B200"},
                         "e2e": {"value": r["value"], "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            print(json.dumps(line))
            return
        r = run_rot(args, rank, world, local, False)
        if rank != 0:
            return
        line = dict(base)
        line.update({"value": r["value"], "ms_per_step": r["ms"] / args.steps, "gpu_launches": r["launches"],
                     "clocks": r["clocks"], "roofline": r.get("roof"), "kernels": r["kernels"],
                     "e2e": {"value": r["e2e"], "unit": "ops/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                     "cpu_baseline": {"value": None, "unit": "ops/s", "cores": 0, "kind": "port",
                                      "sample": "not timed for this workload (see the default workload's line)"}})
        print(json.dumps(line))
        return

    if args.impl == "reference":
        if world > 1 and rank != 0:
            return  # the reference is single-GPU: rank 0 alone runs it
        r = run_reference(args, 0, 1, local)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)"}))
            return
        line = dict(base)
        line.update({"impl": "reference", "value": r["value"], "n_gpus": 1, "ms_per_step": r["ms"] / args.steps,
                     "cpu_baseline": {"value": r["value"], "unit": "ops/s", "cores": 0, "kind": "reference",
                                      "sample": "the reference's own CUDA kernels (sm_100a build) on one B200; it has no CPU path, core count moot"},
                     "e2e": {"value": r["value"], "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line))
        return

    r = run_ours(args, rank, world, local)
    if rank != 0:
        return
    line = dict(base)
    line.update({"value": r["value"], "ms_per_step": r["ms"] / args.steps, "gpu_launches": int(r["launches"]),
                 "clocks": r["clocks"],
                 "e2e": {"value": r["e2e"], "unit": "ops/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                 "roofline": r["roof"], "kernels": r["kernels"]})
    peak, peak_src = peaks()
    ab = algorithmic_bytes_per_op(r["inp"])
    line["roofline_op"] = {"bound": "hbm", "achieved": ab * r["value"] / world / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": ab * r["value"] / world / 1e9 / peak, "alg_bytes_per_op": ab, "peak_source": peak_src}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(r["inp"], args.workload)
        except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
            line["cpu_baseline"] = {"value": None, "unit": "ops/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()

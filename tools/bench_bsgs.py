"""BSGS diagonal matrix-vector product (multiply_matrix_v2, ckks/operator.cu:2898-3390) at the bootstrapping
parameter set C3-II: this engine against the replay of the reference's own kernels, same plan, same buffers.
n1 baby steps x n2 giant steps with full groups (the shape of one CoeffToSlot factor).  Prints one JSON line.
Usage (GPU box):  python tools/bench_bsgs.py [--n1 8 --n2 8 --depth 0 --iters 10]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heongpu_b200 import api  # noqa: E402
from oracle import ref as R  # noqa: E402
from tests.common import PARAMS, oracle_ctx  # noqa: E402
from tests.gpu_common import gpu_ctx, ref_gpu  # noqa: E402


def rand_limbs(primes, n, lead, gen):
    out = torch.empty(tuple(lead) + (len(primes), n), dtype=torch.int64, device="cuda")
    for i, p in enumerate(primes):
        out[..., i, :] = torch.randint(0, int(p), tuple(lead) + (n,), generator=gen, device="cuda", dtype=torch.int64)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="C3_II")
    ap.add_argument("--n1", type=int, default=8)
    ap.add_argument("--n2", type=int, default=8)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    ctx, oc = gpu_ctx(a.name), oracle_ctx(a.name)
    rg = ref_gpu(a.name) if R.have_gpu() else None
    L, n, K = oc.Q - a.depth, oc.n, oc.K
    pql = L + K
    gen = torch.Generator(device="cuda").manual_seed(1)
    rot_n2 = list(range(a.n1))
    rot_n1 = [a.n1 * j for j in range(a.n2)]
    diags = [[g + b for b in rot_n2] for g in rot_n1]
    baby, giant, sizes, terms = api.HEArithmeticOperator.bsgs_plan(n, 5, diags, rot_n1, rot_n2)
    d0 = oc.digits(0)
    keys = {e: rand_limbs(oc.primes, n, (d0, 2), gen) for e in sorted(set(baby + giant) - {0})}
    matrix = rand_limbs(list(oc.primes[:L]) + list(oc.primes[oc.Q:]), n, (len(terms),), gen)
    ct = rand_limbs(oc.primes[:L], n, (1, 2), gen)
    op = api.HEArithmeticOperator(ctx)
    A = api.Ciphertext(ctx, ct, depth=a.depth)
    out = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"), depth=a.depth)
    gk = api.Galoiskey(ctx, keys)

    def ours():
        op.multiply_matrix(A, out, matrix, diags, rot_n1, rot_n2, gk, rescale=False)

    ro = torch.zeros(2, L, n, dtype=torch.int64, device="cuda")

    def theirs():
        rg.bsgs_matvec(ct[0], ro, matrix, baby, [keys.get(e) for e in baby], giant, [keys.get(e) for e in giant], sizes, terms,
                       a.depth)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    res = {"workload": f"{a.name} multiply_matrix_v2 (double hoisting), n1={a.n1} baby x n2={a.n2} giant steps, "
                       f"{len(terms)} diagonals, depth {a.depth}", "ms_ours": timed(ours)}
    api.lib.heon_profile_begin()
    ours()
    msv, cnt = (C.c_double * 16)(), (C.c_longlong * 16)()
    ncls = api.lib.heon_profile_end(msv, cnt, 16)
    res["kernels_ours"] = {api.lib.heon_profile_class_name(i).decode(): {"ms": msv[i], "launches": int(cnt[i])}
                           for i in range(ncls) if cnt[i]}
    if rg is not None:
        res["ms_reference_kernels"] = timed(theirs)
        res["speedup"] = res["ms_reference_kernels"] / res["ms_ours"]
        torch.cuda.synchronize()
        res["bit_exact"] = bool(torch.equal(out.data[0], ro))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

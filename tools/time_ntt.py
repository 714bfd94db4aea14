"""CUDA-event timing of the batched NTT (us per limb-polynomial)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heongpu_b200 import api
from tests.common import PARAMS

name = sys.argv[1] if len(sys.argv) > 1 else "n16_II_small"
ppp = int(sys.argv[2]) if len(sys.argv) > 2 else 37
log_n, qb, pb = PARAMS[name]
ctx = api.HEContext(log_n, qb, pb, device=0)
order = ctx.level_primes(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
p = torch.tensor([ctx.primes[i] for i in order], dtype=torch.int64, device="cuda").view(1, -1, 1)
x = torch.randint(0, 1 << 62, (ppp, len(order), ctx.n), dtype=torch.int64, device="cuda", generator=g) % p
x0 = x.clone()
polys = ppp * len(order)
for inverse in (False, True):
    for _ in range(3):
        ctx.ntt(x, order, inverse=inverse)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ctx.ntt(x, order, inverse=inverse)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps / polys
    print(f"{name} variant={os.environ.get('HEON_NTT_VARIANT','default')} {'INTT' if inverse else 'NTT '} {us:.3f} us/poly  ({polys*ctx.n*16/us/1e3:.0f} GB/s alg)")
# correctness: fwd then inv restores
x.copy_(x0); ctx.ntt(x, order); ctx.ntt(x, order, inverse=True)
print("roundtrip ok:", bool(torch.equal(x, x0)))

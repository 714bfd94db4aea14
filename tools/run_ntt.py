"""Small driver for ncu captures: forward+inverse NTT over many limb-polynomials."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heongpu_b200 import api
from tests.common import PARAMS

name = sys.argv[1] if len(sys.argv) > 1 else "n16_II_small"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
polys_per_prime = int(sys.argv[3]) if len(sys.argv) > 3 else 37
log_n, qb, pb = PARAMS[name]
ctx = api.HEContext(log_n, qb, pb, device=0)
order = ctx.level_primes(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
p = torch.tensor([ctx.primes[i] for i in order], dtype=torch.int64, device="cuda").view(1, -1, 1)
x = torch.randint(0, 1 << 62, (polys_per_prime, len(order), ctx.n), dtype=torch.int64, device="cuda", generator=g) % p
for _ in range(reps):
    ctx.ntt(x, order)
    ctx.ntt(x, order, inverse=True)
torch.cuda.synchronize()
print("done", x.shape)

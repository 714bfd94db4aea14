"""CUDA-event timing: R hoisted rotations of one ciphertext vs R stand-alone rotations (C3_II / C3_I)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heongpu_b200 import api
from tests.common import PARAMS

name = sys.argv[1] if len(sys.argv) > 1 else "C3_II"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
log_n, qb, pb = PARAMS[name]
ctx = api.HEContext(log_n, qb, pb, device=0)
L, n, Qp = ctx.Q_size, ctx.n, ctx.Q_prime_size
g = torch.Generator(device="cuda"); g.manual_seed(1)
def res(lead, plist):
    p = torch.tensor(plist, dtype=torch.int64, device="cuda").view(*([1] * len(lead)), len(plist), 1)
    return torch.randint(0, 1 << 62, (*lead, len(plist), n), dtype=torch.int64, device="cuda", generator=g) % p
a = res((B, 2), ctx.primes[:L])
shifts = [1 << i for i in range(R)]
elts = [api.lib.heon_steps_to_galois_elt(s, n, 5) for s in shifts]
gk = api.Galoiskey(ctx, {e: res((ctx.digits(0), 2), ctx.primes) for e in elts})
op = api.HEArithmeticOperator(ctx)
A = api.Ciphertext(ctx, a)
outs = torch.zeros(R, B, 2, L, n, dtype=torch.int64, device="cuda")
single = api.Ciphertext(ctx, torch.zeros(B, 2, L, n, dtype=torch.int64, device="cuda"))
def t(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
th = t(lambda: op.rotate_rows_hoisted(A, outs, gk, shifts))
ts = t(lambda: [op.rotate_rows(A, single, gk, s) for s in shifts])
print(f"{name}: {R} rotations x {B} ciphertexts: hoisted {th:.3f} ms ({R*B/th*1e3:.0f} rot/s), stand-alone {ts:.3f} ms ({R*B/ts*1e3:.0f} rot/s), speed-up {ts/th:.2f}x")

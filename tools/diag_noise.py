import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from heongpu_b200 import api
for (log_n,qb,pb) in [(12,[40,30,30],[40]),(12,[40,30,30],[50])]:
    ctx = api.HEContext(log_n, qb, pb, device=0)
    kg = api.HEKeyGenerator(ctx, seed=1234)
    sk = kg.generate_secret_key(api.Secretkey(ctx)); pk = kg.generate_public_key(api.Publickey(ctx), sk)
    enc, cry, dec = api.HEEncoder(ctx), api.HEEncryptor(ctx, pk), api.HEDecryptor(ctx, sk)
    op = api.HEArithmeticOperator(ctx)
    swk = kg.generate_switch_key(sk, sk)
    gk = kg.generate_galois_key(sk, shifts=[1,2])
    m = np.random.default_rng(2).uniform(0,1,ctx.n//2)
    scale=2.0**30
    c1=cry.encrypt(enc.encode(m,scale))
    def mk(): return api.Ciphertext(ctx, torch.zeros(1, 2, ctx.Q_size, ctx.n, dtype=torch.int64, device="cuda"))
    d0=enc.decode(dec.decrypt(c1))
    out=mk(); op.keyswitch(c1,out,swk); out.scale_=scale
    e=(enc.decode(dec.decrypt(out))-d0)*scale
    print(qb,pb,'keyswitch: std',np.abs(e).std(),'max',np.abs(e).max())
    for sh in (1,2):
        out=mk(); op.rotate_rows(c1,out,gk,sh)
        e=(enc.decode(dec.decrypt(out))-np.roll(d0,-sh))*scale
        print('   rotate',sh,': std',np.abs(e).std(),'max',np.abs(e).max(),'argmax',np.abs(e).argmax(), 'median', np.median(np.abs(e)))
    out=mk(); op.conjugate(c1,out,gk)
    e=(enc.decode(dec.decrypt(out))-np.conj(d0))*scale
    print('   conj: std',np.abs(e).std(),'max',np.abs(e).max())

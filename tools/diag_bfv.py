import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from heongpu_b200 import api
for (log_n,qb,pb,t) in [(12,[36,36],[37],1032193),(14,[54,54,54,54,55,55,55],[55],786433)]:
    ctx = api.HEContext(log_n, qb, pb, device=0, plain_modulus=t)
    kg = api.HEKeyGenerator(ctx, seed=5)
    sk = kg.generate_secret_key(api.Secretkey(ctx)); pk = kg.generate_public_key(api.Publickey(ctx), sk)
    enc, cry, dec = api.HEEncoder(ctx), api.HEEncryptor(ctx, pk), api.HEDecryptor(ctx, sk)
    op = api.HEArithmeticOperator(ctx)
    rng=np.random.default_rng(1); n=ctx.n; bad=0; badfresh=0
    for it in range(150):
        m1=rng.integers(0,t,n); m2=rng.integers(0,t,n)
        c1=cry.encrypt(enc.encode(m1)); c2=cry.encrypt(enc.encode(m2))
        f=enc.decode(dec.decrypt(c1))
        if not np.array_equal(f,m1.astype(np.uint64)): badfresh+=1; print('fresh mismatch', it, np.nonzero(f!=m1.astype(np.uint64))[0][:5])
        out=api.Ciphertext(ctx, torch.zeros(1,2,ctx.Q_size,n,dtype=torch.int64,device='cuda')); out.in_ntt_domain_=False
        op.add(c1,c2,out)
        g=enc.decode(dec.decrypt(out))
        w=((m1+m2)%t).astype(np.uint64)
        if not np.array_equal(g,w):
            bad+=1; idx=np.nonzero(g!=w)[0]; print('add mismatch', it, idx[:5], g[idx[:3]], w[idx[:3]], m1[idx[:3]], m2[idx[:3]])
    print(qb,pb,'fresh bad',badfresh,'add bad',bad,'of 150')

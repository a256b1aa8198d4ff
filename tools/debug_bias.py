import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from support import *
from oracle import restatement as O
from cmfrec_b200 import _lib
dt=np.dtype(np.float32); R=ref(dt); L=_lib.load(dt)
for k in (3,16):
    scale_lam=False
    m,n=600,380
    ixA,ixB,X=synth_coo(m,n,8000,dt,seed=100+k); X=(X-X.mean()).astype(dt)
    csr=csr_csc(L,dt,ixA,ixB,X,m,n)
    rng=np.random.default_rng(k)
    A0=(rng.normal(size=(m,k))*0.1).astype(dt); B0=(rng.normal(size=(n,k))*0.1).astype(dt)
    bA0=(rng.normal(size=m)*0.3).astype(dt); bB0=(rng.normal(size=n)*0.3).astype(dt)
    lam,lam_bias=(1.5,2.5)
    A_b=np.concatenate([A0,np.ones((m,1),dt)],1); B_b=np.concatenate([B0,np.ones((n,1),dt)],1)
    Xcsc=(csr[5]-bA0[csr[4]]).astype(dt)
    B1=B_b.copy(); ref_optimizeA(R,dt,B1,A_b.copy(),csr[3],csr[4],Xcsc,lam=lam,lam_last=lam_bias,scale_lam=scale_lam,use_cg=True,max_cg_steps=3)
    B2=B_b.copy(); O.optimizeA(dt,B2,A_b.copy(),csr[3],csr[4],Xcsc,lam=lam,lam_last=lam_bias,scale_lam=scale_lam,use_cg=True,max_cg_steps=3)
    with AlsSession(L, dt, csr[:3], csr[3:], m, n, k, implicit=False, user_bias=True, item_bias=True, lam_A=lam, lam_B=lam, lam_biasA=lam_bias, lam_biasB=lam_bias, scale_lam=scale_lam) as s:
        s.set_factors(A0, bA0, B0, bB0)
        s.half_sweep(0, 1, 0)
        _, _, Bg, bBg = s.get_factors(with_bias=True)
    G=np.concatenate([Bg,bBg[:,None]],1)
    d=np.abs(G-B1).max(axis=1); dB=np.abs(G[:,k]-B1[:,k]); dO=np.abs(B2-B1).max(axis=1)
    deg=np.diff(csr[3].astype(np.int64))
    worst=np.argsort(-dB)[:8]
    print('k',k,'scale',np.abs(B1).max())
    for w in worst: print(' row',w,'deg',deg[w],'dbias %.3e dall %.3e oracle-ref %.3e'%(dB[w],d[w],dO[w]),'gpu',G[w,k],'ref',B1[w,k],'orc',B2[w,k])

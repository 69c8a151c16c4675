"""numpy emulation behind the arithmetic choice of csrc/nempc_tc.cuh: forward second-order rows through a 128-wide tanh MLP with the
hidden-to-hidden products done in f32, in the two-term f16 split (two accumulators / one), and in plain TF32; max error of value,
Jacobian and per-output Hessians against float64.  Development tool: imports oracle/ as the checker."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle.mlp_np import MLP
rng=np.random.default_rng(0)
def split(x, scaled=True):
    x=x.astype(np.float32)
    h=x.astype(np.float16)
    r=(x-h.astype(np.float32))
    if scaled: l=(r*np.float32(2048)).astype(np.float16)
    else: l=r.astype(np.float16)
    return h.astype(np.float32), l.astype(np.float32)
def mm(A,W,mode):
    # A (rows,K) f32, W (K,N) f32
    if mode=='f32': return (A.astype(np.float32)@W.astype(np.float32))
    if mode=='split2':   # two accumulators, scaled lo
        a1,a2=split(A); w1,w2=split(W)
        main=(a1.astype(np.float64)@w1).astype(np.float32)
        corr=(a2.astype(np.float64)@w1 + a1.astype(np.float64)@w2).astype(np.float32)
        return main+corr*np.float32(1/2048)
    if mode=='split1':  # single accumulator unscaled lo
        a1,a2=split(A,False); w1,w2=split(W,False)
        return (a1.astype(np.float64)@w1+a2.astype(np.float64)@w1+a1.astype(np.float64)@w2).astype(np.float32)
    if mode=='tf32':
        def t(x): 
            v=x.astype(np.float32).view(np.uint32)&np.uint32(0xffffe000); return v.view(np.float32)
        return (t(A).astype(np.float64)@t(W)).astype(np.float32)
def chain(mlp, z, mode, dt=np.float32):
    # forward second order: returns f (x), J (x,d), Hs (x,d,d)
    d=mlp.d; Ws=mlp.weights; nh=len(Ws)-1
    pairs=[(c,c2) for c in range(d) for c2 in range(c+1)]
    W0,b0=Ws[0]
    a=(z@W0+b0).astype(dt); h=np.tanh(a).astype(dt); s1=1-h*h; s2=-2*h*s1
    P=h[None,:]; T=(s1[None,:]*W0).astype(dt)   # (d,h)
    S=np.stack([s2*W0[c]*W0[c2] for c,c2 in pairs]).astype(dt)
    for l in range(1,nh):
        W,b=Ws[l]
        rows=np.concatenate([P,T,S],0)
        out=mm(rows,W,mode) if mode!='f64' else rows@W
        a=out[0]+b; Tn=out[1:1+d]; Sn=out[1+d:]
        h=np.tanh(a).astype(dt); s1=1-h*h; s2=-2*h*s1
        P=h[None,:]; T=(s1*Tn).astype(dt)
        S=np.stack([s2*Tn[c]*Tn[c2]+s1*Sn[i] for i,(c,c2) in enumerate(pairs)]).astype(dt)
    W,b=Ws[nh]
    rows=np.concatenate([P,T,S],0)
    out=mm(rows,W,mode) if mode!='f64' else rows@W
    f=out[0]+b; J=out[1:1+d].T; Hs=np.zeros((mlp.x_dim,d,d),out.dtype)
    for i,(c,c2) in enumerate(pairs): Hs[:,c,c2]=out[1+d+i]; Hs[:,c2,c]=out[1+d+i]
    return f,J,Hs
for dims,x,u,wscale in [([5,128,128,128,4],4,1,1.0),([5,128,128,128,4],4,1,3.0),([3,128,128,2],2,1,1.0)]:
    mlp=MLP.glorot(dims,x,u,seed=1)
    mlp.weights=[(W*wscale,b) for W,b in mlp.weights]
    errs={m:[0,0,0] for m in ['f32','split2','split1','tf32']}
    for trial in range(20):
        z=rng.uniform(-1,1,x+u)
        ref=mlp.blocks(z) if False else None
        f64=chain(MLP(mlp.weights,x,u),z,'f64',np.float64)
        if trial==0:
            fb,Jb,Hb=[v[0] for v in mlp.blocks(z[None,:])[:3]] if hasattr(mlp,'blocks') else (None,)*3
            print('oracle check', np.abs(fb-f64[0]).max(), np.abs(Jb-f64[1]).max(), np.abs(np.asarray(Hb)-f64[2]).max())
        for m in errs:
            r=chain(mlp,z,m)
            for k in range(3):
                e=np.abs(r[k]-f64[k]).max()/max(1e-30,np.abs(f64[k]).max())
                errs[m][k]=max(errs[m][k],e)
    print(dims,wscale,{m:['%.1e'%v for v in e] for m,e in errs.items()})

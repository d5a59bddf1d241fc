import os, sys, numpy as np
sys.path.insert(0, "/root/repo/reservoircomputing.jl_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import rcb200 as R
from oracle import esn_oracle as O
def serr(a,b):
    num=np.max(np.abs(a.astype(np.float64)-b.astype(np.float64)),axis=0); den=np.maximum(np.max(np.abs(b.astype(np.float64)),axis=0),1e-30); return float(np.max(num/den))
for N,d,mods,f64 in [(300,3,"nlat2",True),(300,3,"nlat2_pad",False),(1024,1,"none",True),(65,3,"pad",True),(65,3,"none",True),(300,3,"pad",True)]:
    rng=np.random.default_rng(N+d)
    chain=dict(nlat2=(R.NLAT2,),nlat2_pad=(R.NLAT2,R.Pad(1.0)),none=(),pad=(R.Pad(0.5),))[mods]
    esn=R.ESN(d,N,d,init_reservoir=R.rand_sparse(radius=0.9,sparsity=min(1.0,6/N)),leak_coefficient=0.7,state_modifiers=chain)
    ps,st=R.setup(rng,esn)
    T=400; tt=np.arange(T+1)[None,:]
    series=(np.sin(0.07*tt*(1+np.arange(d))[:,None])*0.6).astype(np.float32)
    data,tgt=np.asfortranarray(series[:,:T]),np.asfortranarray(series[:,1:])
    obj=R.RidgeRegression(1e-3) if f64 else R.RidgeRegression(np.float32,1e-3)
    ps2,st2=R.train(esn,data,tgt,ps,st,objective=obj,washout=50)
    u0=tgt[:,-1]
    os.environ.pop("RCB_K4_MODE",None)
    Y,_=R.predict(esn,40,ps2,st2,initialdata=u0); k1=R.default_device().last_kernel()
    os.environ["RCB_K4_MODE"]="generic"
    Yg,_=R.predict(esn,40,ps2,st2,initialdata=u0); k2=R.default_device().last_kernel()
    from test_gpu_parity import oracle_layers
    layers=oracle_layers(R,esn,ps)
    Yo,_=O.predict_autoregressive(layers,ps2.readout.weight,40,u0,[st2.reservoir.carry[0]])
    print(N,d,mods,f64,k1,k2,"auto-gen",serr(Y,Yg),"auto-oracle",serr(Y,Yo),"gen-oracle",serr(Yg,Yo), "|W|",np.linalg.norm(ps2.readout.weight), flush=True)

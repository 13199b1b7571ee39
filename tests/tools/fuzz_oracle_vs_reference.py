"""Randomised bit-exactness check of the numpy oracle (oracle/np_oracle.py) against the LIVE reference
(/root/reference, build container only): grid type, shape, batch, random land, NaN / inf on land, batched grid
variables, fp32 / fp64, Gaussian / Taper.  Values AND dtypes must agree exactly (1200 cases clean in round 1).

    python tests/tools/fuzz_oracle_vs_reference.py [seed] [cases]
"""
import sys, warnings
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from oracle import fixtures, np_oracle, ref_loader
GRIDS = fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad=0; N=int(sys.argv[2]) if len(sys.argv)>2 else 200
warnings.simplefilter("ignore")
for k in range(N):
    g=GRIDS[rng.integers(len(GRIDS))]
    ny,nx=int(rng.integers(6,50)),int(rng.integers(6,70))
    if g.startswith("TRIPOLAR") and nx%2: nx+=1
    nb=int(rng.integers(1,4))
    fields,gv=fixtures.fixture(g,(ny,nx)); gv={k_:v.copy() for k_,v in gv.items()}
    masks=[k_ for k_ in gv if "mask" in k_]
    if masks and rng.random()<0.6:
        land=rng.random((ny,nx))<rng.uniform(0,0.4)
        for m in masks: gv[m]=gv[m]*(~land)
    fb=tuple(np.stack([f*(1+0.1*b)+0.05*rng.standard_normal((ny,nx)) for b in range(nb)]) for f in fields)
    if "wet_mask" in gv and rng.random()<0.7:
        junk=[np.nan,np.inf,-np.inf][int(rng.integers(3))]
        for f in fb: f[:,gv["wet_mask"]==0]=junk
    if nb>1 and not g.startswith("MOM5") and rng.random()<0.3:
        for k_ in list(gv):
            if "mask" in k_:
                lvl=np.stack([gv[k_]*(rng.random((ny,nx))>0.1*b) for b in range(nb)])
                if g.startswith("TRIPOLAR"): lvl[:,0,:]=0
                gv[k_]=lvl
            elif "kappa" not in k_:
                gv[k_]=np.stack([gv[k_]*(1.0+0.05*b) for b in range(nb)])
    if rng.random()<0.3:
        dt=np.float32; fb=tuple(f.astype(dt) for f in fb); gv={k_:v.astype(dt) for k_,v in gv.items()}
    dxm=1.0
    if g in fixtures.VECTOR_GRIDS:
        kx,ky=("dxT","dyT") if g=="VECTOR_C_GRID" else ("DXU","DYU"); dxm=float(min(gv[kx].min(),gv[ky].min()))
    fa=dict(filter_scale=float(rng.uniform(3,9))*dxm, dx_min=dxm, filter_shape=["GAUSSIAN","TAPER"][int(rng.integers(2))])
    try:
        a=np_oracle.apply_filter(g,gv,fb,**fa); b=ref_loader.ref_filter(g,gv,fb,**fa)[0]
        a=a if isinstance(a,tuple) else (a,); b=b if isinstance(b,tuple) else (b,)
        ok=all(x.dtype==y.dtype and np.array_equal(x,y,equal_nan=True) for x,y in zip(a,b))
        la=np_oracle.laplacian(g,gv,*fb); lb=ref_loader.ref_laplacian(g,gv,*fb)
        la=la if isinstance(la,tuple) else (la,); lb=lb if isinstance(lb,tuple) else (lb,)
        ok=ok and all(np.array_equal(x,y,equal_nan=True) for x,y in zip(la,lb))
    except Exception as e:
        ok=False; print("EXC",g,type(e).__name__,e)
    if not ok:
        bad+=1; print("FAIL",k,g,ny,nx,nb,fb[0].dtype,fa)
print(N-bad,"/",N)
sys.exit(1 if bad else 0)

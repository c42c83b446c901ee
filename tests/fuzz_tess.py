"""Randomised check of the host tess() driver's ghost rounds (test infrastructure): a block that needed several rounds (the Delaunay
engine inserts only the new ghosts, Delaunay3::add) must equal the block a single round at the final margin gives -- same
particles in the same order, same tets as a set (near-degenerate lattices: same tet count and total volume).
    python tests/fuzz_tess.py [seed] [seconds]
Round 1: 2 324 cases, 16 107 multi-round blocks over 4 seeds (600 s each), no difference."""
import os
import sys
import time

import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tess2_b200 import host_tess
from tess2_b200.harness import particles
def tet_set(t): return set(map(tuple, np.sort(np.asarray(t)[:, :4], axis=1)))
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0); t0=time.time(); it=0; multi_rounds=0; bad=0
while time.time()-t0 < (float(sys.argv[2]) if len(sys.argv) > 2 else 60.0):
    it+=1
    n=int(rng.integers(200,20000)); ext=float(rng.choice([1.0,31.0,1000.0]))
    dom=([0,0,0],[ext]*3)
    kind=rng.choice(["clustered","uniform","lattice"])
    if kind=="clustered": p=particles.clustered_particles(n,*dom,seed=int(rng.integers(1<<30)),n_clumps=int(rng.integers(1,10)))
    elif kind=="uniform": p=particles.uniform_particles(n,*dom,seed=int(rng.integers(1<<30)))
    else:
        k=max(3,int(round(n**(1/3)))); g=np.stack(np.meshgrid(*[np.linspace(0.05,0.95,k)]*3,indexing="ij"),-1).reshape(-1,3)
        p=np.unique((np.clip(g+rng.normal(0,1e-3,g.shape),0.001,0.999)*ext).astype(np.float32),axis=0)
    nb=int(rng.choice([2,4,8,16]))
    bounds,owner=host_tess.kdtree_blocks(p,*dom,nb)
    m0=float(rng.choice([0.0, 0.02*ext, 0.001*ext]))
    blocks=host_tess.tess(p,owner,bounds,*dom,margin0=m0,max_rounds=int(rng.choice([2,3,4])),max_growth=float(rng.choice([2.5,20.0])))
    for b in blocks:
        if b["rounds"]<2: continue
        multi_rounds+=1
        one=host_tess.tess(p,owner,bounds,*dom,margin0=float(b["margin"]),max_rounds=1,gids=[b["gid"]])[0]
        ok = np.array_equal(one["global_ids"],b["global_ids"]) and (kind=="lattice" or tet_set(one["tets"])==tet_set(b["tets"]))
        if kind=="lattice":   # near-degenerate: any valid triangulation; compare total volume of tets instead
            def vol(bb):
                P=bb["particles"].astype(np.float64); T=bb["tets"][:,:4]
                a,b_,c,d=[P[T[:,i]] for i in range(4)]
                return np.abs(np.einsum("ij,ij->i",np.cross(a-d,b_-d),c-d)).sum()/6
            ok = ok and abs(vol(one)-vol(b))<1e-6*vol(one) and len(one["tets"])==len(b["tets"])
        if not ok:
            bad+=1; print("DIFF",dict(n=len(p),kind=str(kind),nb=nb,m0=m0,gid=b["gid"],rounds=b["rounds"]),flush=True)
print("cases",it,"multi-round blocks",multi_rounds,"bad",bad)

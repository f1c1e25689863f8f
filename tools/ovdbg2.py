"""Candidate counts / exact-path fallbacks of the overlap selection in a bench-like run (25 origins)."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/statdyn-analysis_b200')
import __graft_entry__; __graft_entry__.build()
import torch, numpy as np
from sdanalysis_b200.engine import DynamicsEngine
from sdanalysis_b200.molecules import Trimer
from sdanalysis_b200.synthetic import SyntheticTrimer, box_for
n=1_000_000; K=int(sys.argv[1]) if len(sys.argv)>1 else 25
dev=torch.device('cuda',0)
eng=DynamicsEngine(n, box_for(n), 2.9, mol_sites=Trimer().positions, device=dev, ring_depth=4)
walk=SyntheticTrimer(n, seed=4, sigma_r=0.01, sigma_theta=0.02, device=dev)
ts=0
keys=list(range(K))
for k in range(K):
    pos,quat=walk.frame()
    eng.step(ts,pos,quat,keys[:k+1],want_sums=False)
    walk.advance(1); ts+=1
for t in range(30):
    pos,quat=walk.frame()
    r=eng.step(ts,pos,quat,keys)
    tab=r.table()
    f,c=eng.overlap_stats(K)
    print(t, 'fails',f,'cands min/mean/max',int(c.min()),int(c.mean()),int(c.max()), 'ov', np.round(tab['overlap'][:3],4).tolist(), flush=True)
    walk.advance(1); ts+=1

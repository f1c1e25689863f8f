import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/statdyn-analysis_b200')
import __graft_entry__; __graft_entry__.build()
import torch, numpy as np
from sdanalysis_b200.engine import DynamicsEngine
from sdanalysis_b200.molecules import Trimer
from sdanalysis_b200.synthetic import SyntheticTrimer, box_for
n=1_000_000; K=8
dev=torch.device('cuda',0)
eng=DynamicsEngine(n, box_for(n), 2.9, mol_sites=Trimer().positions, device=dev, ring_depth=4)
walk=SyntheticTrimer(n, seed=4, sigma_r=0.01, sigma_theta=0.02, device=dev)
ts=0
for t in range(60):
    pos,quat=walk.frame()
    act=list(range(min(K, t//5+1)))
    r=eng.step(ts,pos,quat,act)
    tab=r.table()
    f,c=eng.overlap_stats(len(act))
    if t%3==0 or f: print(t, 'fails',f,'cands',c.tolist(), 'ov', np.round(tab['overlap'],4).tolist())
    walk.advance(1); ts+=1

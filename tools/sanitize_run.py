"""Small pass over every kernel family for compute-sanitizer (memcheck): dynamics step with sums, overlap,
struct and trackers over a few frames and origins, neighbour search, order parameter, Voronoi counts, RDF."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/statdyn-analysis_b200")
import __graft_entry__; __graft_entry__.build()
import numpy as np
import torch
from sdanalysis_b200 import order
from sdanalysis_b200.dynamics import _util
from sdanalysis_b200.engine import DynamicsEngine
from sdanalysis_b200.molecules import Trimer
from sdanalysis_b200.synthetic import SyntheticTrimer
from sdanalysis_b200.util import create_freud_box

dev = torch.device("cuda", 0)
n = 5003  # not a multiple of 4 or 256: partial tiles and the bounds-checked paths
walk = SyntheticTrimer(n, seed=1, sigma_r=0.05, sigma_theta=0.1, device=dev)
eng = DynamicsEngine(n, walk.box, 2.9, mol_sites=Trimer().positions, device=dev, compute_isf=True)
table = None
for t in range(12):
    pos, quat = walk.frame()
    active = list(range(min(5, t // 2 + 1)))
    table = eng.step(walk.timestep, pos, quat, active).table()
    walk.advance(3)
print("dynamics ok", {k: float(np.asarray(v)[0]) for k, v in table.items() if k in ("msd", "overlap", "struct")})
print("summary ok", sorted(eng.summary(0)))
pos, quat = walk.frame()
res = order.neighbour_analysis(walk.box, pos, quat, 3.5, 8, want=("nlist", "rsq", "counts", "order"))
print("neighbours ok", res["counts"].mean())
print("voronoi ok", order.compute_voronoi_neighs(walk.box, pos).mean())
print("wave number ok", _util.calculate_wave_number(create_freud_box(walk.box, is_2D=True), pos[:1500].cpu().numpy()))
torch.cuda.synchronize()

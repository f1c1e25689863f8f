"""Device time of the auxiliary kernels: Voronoi neighbour counts and the RDF pair histogram."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/statdyn-analysis_b200")
import __graft_entry__; __graft_entry__.build()
import torch
from sdanalysis_b200 import order
from sdanalysis_b200.dynamics import _util
from sdanalysis_b200.synthetic import SyntheticTrimer
from sdanalysis_b200.util import create_freud_box
dev = torch.device("cuda", 0)
for n in (32768, 1_000_000):
    walk = SyntheticTrimer(n, seed=3, device=dev)
    pos, quat = walk.frame()
    order.compute_voronoi_neighs(walk.box, pos)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        c = order.compute_voronoi_neighs(walk.box, pos)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"voronoi N={n}: {1e3*dt:.3f} ms per frame (incl. neighbour search and D2H), mean count {c.mean():.4f}", flush=True)
for n in (1250, 32768):
    walk = SyntheticTrimer(n, seed=3, device=dev)
    pos, quat = walk.frame()
    box = create_freud_box(walk.box, is_2D=True)
    _util.calculate_wave_number(box, pos)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    k = _util.calculate_wave_number(box, pos)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"calculate_wave_number N={n}: {1e3*dt:.3f} ms ({n*(n-1):.3e} pairs), k = {k:.4f}", flush=True)

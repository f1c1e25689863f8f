"""Device time of the order path (BASELINE configs[2] functions) at several N: one cell-list build
producing list + r^2 + counts + order parameter (k = 8, r = 3.5), inputs resident in HBM."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/statdyn-analysis_b200")
import __graft_entry__; __graft_entry__.build()
import torch
from sdanalysis_b200 import order
from sdanalysis_b200.synthetic import SyntheticTrimer
dev = torch.device("cuda", 0)
for n in [int(a) for a in sys.argv[1:]] or [32768, 1_000_000, 4_000_000]:
    walk = SyntheticTrimer(n, seed=3, device=dev)
    pos, quat = walk.frame()
    want = ("nlist", "rsq", "counts", "order")
    for _ in range(3):
        order.neighbour_analysis(walk.box, pos, quat, 3.5, 8, want=want, device=dev, to_host=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        order.neighbour_analysis(walk.box, pos, quat, 3.5, 8, want=want, device=dev, to_host=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"N={n}: {ms:.4f} ms per frame  {n/ms*1e3:.3e} molecule-frames/s  algorithmic 68 B/molecule -> {68*n/ms/1e6:.1f} GB/s", flush=True)

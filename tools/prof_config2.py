"""cProfile of the host loop on BASELINE configs[1] (32k molecules, exponential schedule)."""
import sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/statdyn-analysis_b200")
import __graft_entry__; __graft_entry__.build()
import torch
from sdanalysis_b200.read import process_frames
from sdanalysis_b200.synthetic import synthetic_frames, exponential_schedule, linear_schedule
which = sys.argv[1] if len(sys.argv) > 1 else "exp"
sched = exponential_schedule(1000, 10, 10, 100) if which == "exp" else linear_schedule(1000, 10, 99)
frames = list(synthetic_frames(sched, 32768, seed=2))
rows = []
process_frames(iter(frames[:50]), 2.9, sink=rows.append)  # warm
rows = []
torch.cuda.synchronize(); t0 = time.perf_counter()
process_frames(iter(frames), 2.9, sink=rows.append)
torch.cuda.synchronize(); print("wall", time.perf_counter() - t0, "frames", len(frames), "rows", len(rows))
pr = cProfile.Profile(); pr.enable()
rows = []
process_frames(iter(frames), 2.9, sink=rows.append)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)

"""Throughput of the other BASELINE.json configurations through the public host API
(process_frames = the loop of read.process_file; order.* per frame).  Not a bench line: numbers for
DESIGN.md.  usage: python tools/config_runs.py [config2exp config2lin config3 ...]"""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/statdyn-analysis_b200")
import __graft_entry__; __graft_entry__.build()
import numpy as np
import torch
from sdanalysis_b200.read import process_frames
from sdanalysis_b200.synthetic import synthetic_frames, linear_schedule, exponential_schedule, SyntheticTrimer
from sdanalysis_b200 import order

def run_dynamics(name, schedule, n, seed):
    frames = list(synthetic_frames(schedule, n, seed=seed))
    pairs = sum(len(ix) for ix, _ in frames)
    rows = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng = process_frames(iter(frames), 2.9, sink=rows.append)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: N={n} frames={len(frames)} pairs={pairs} rows={len(rows)} wall={dt:.3f}s "
          f"updates/s={n*pairs/dt:.3e} ms/frame={1e3*dt/len(frames):.3f}", flush=True)

which = sys.argv[1:] or ["config2exp", "config2lin", "config3"]
if "config2exp" in which:
    run_dynamics("config2 S_exp (host frames)", exponential_schedule(1000, 10, 10, 100), 32768, 2)
if "config2lin" in which:
    run_dynamics("config2 S_lin (host frames)", linear_schedule(1000, 10, 99), 32768, 2)
if "config1" in which:
    run_dynamics("config1-like (1250 molecules)", exponential_schedule(10000, 100, 1000, 500), 1250, 1)
if "config3" in which:
    n = 32768
    walk = SyntheticTrimer(n, seed=3)
    frames = []
    for _ in range(50):
        frames.append(walk.frame()); walk.advance(1)
    box = walk.box
    order.compute_neighbours(box, frames[0][0]);  # warm
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for pos, quat in frames:
        nl = order.compute_neighbours(box, pos, 3.5, 8)
        oo = order.orientational_order(box, pos, quat, 3.5, 8)
        nn = order.num_neighbours(box, pos, 3.5)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"config3 order path: N={n} frames={len(frames)} wall={dt:.3f}s ms/frame={1e3*dt/len(frames):.3f} "
          f"molecule-frames/s={n*len(frames)/dt:.3e}", flush=True)

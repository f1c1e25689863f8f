#!/usr/bin/env python
"""Benchmark of the per-frame dynamics hot path (BASELINE.json: molecule*origin updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3], the configuration the target is quoted on): synthetic 2D Trimer
liquid, 1,000,000 molecules, 200 time-origins, dense (linear) schedule in its steady state -- every
frame updates all 200 origins (Dynamics.compute_all + Relaxations.add each).  A *step* is one frame:
200 x 1e6 molecule*origin updates.  With N GPUs the 200 origins are sharded round-robin over the
ranks and each frame is broadcast from rank 0 over NCCL (strong scaling: total work is fixed).

* ``value``  frames already resident in HBM (rank 0) when the timed region starts.
* ``e2e``    the same loop through the public host API with frames in pinned HOST memory: the H2D
             copy of every frame and the D2H read of its result rows are inside the timed region.
* ``roofline``      dominant kernel (``dyn_step_kernel``) timed with CUDA events on its stream.
* ``cpu_baseline``  the CPU oracle (a numpy restatement of the reference; the reference itself cannot
             run here: freud/rowan/gsd are not installed) on a bounded sample, one core.
* ``--impl reference``  the CPU arm alone: origin-parallel oracle processes on all host cores.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for _p in (REPO, REPO / "statdyn-analysis_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

METRIC = "molecule*origin updates/sec (dynamics+relaxation)"
UNIT = "updates/s"
N_MOLS = 1_000_000
N_ORIGINS = 200
WAVE_NUMBER = 2.9
PIPELINE = int(os.environ.get("SDB_BENCH_PIPELINE", "2"))  # frames in flight before the oldest result is read
SIGMA_R, SIGMA_T = 0.01, 0.02
WORKLOAD = (
    "configs[3]: synthetic 2D Trimer liquid, 1M molecules, 200 time-origins, dense linear schedule "
    "(steady state: all origins active every frame), Dynamics.compute_all + Relaxations.add"
)


def _peaks():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernel, updates):
    """DRAM bytes per launch of ``kernel`` from the committed ncu capture, scaled to ``updates`` per launch."""
    path = REPO / "profiles" / "ncu_traffic.json"
    try:
        rec = json.loads(path.read_text())[kernel]
        return (rec["dram_bytes_read"] + rec["dram_bytes_write"]) * updates / rec["updates_in_launch"]
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, sm_max, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                sm_max = max(sm_max, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": sm_max or None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy restatement of the reference), timed on host cores
# ------------------------------------------------------------------------------------------------
def _oracle_worker(n_mols, seed, n_frames, barrier, out_q, widx):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import dynamics as odyn
    from sdanalysis_b200.synthetic import SyntheticTrimer

    walk = SyntheticTrimer(n_mols, seed=seed, sigma_r=SIGMA_R, sigma_theta=SIGMA_T)
    pos, quat = walk.frame()
    dyn = odyn.Dynamics(0, walk.box, pos, quat, "trimer", WAVE_NUMBER)
    relax = odyn.Relaxations(0, walk.box, pos, quat, "trimer", wave_number=WAVE_NUMBER)
    frames = []
    for _ in range(n_frames):
        walk.advance(1)
        frames.append((walk.timestep,) + walk.frame())
    times = []
    for ts, pos, quat in frames:
        barrier.wait()
        t0 = time.perf_counter()
        dyn.compute_all(ts, pos, quat)
        relax.add(ts, pos, quat)
        barrier.wait()
        times.append(time.perf_counter() - t0)
    out_q.put((widx, times))


def cpu_arm(n_mols, workers, steps, warmup):
    """Origin-parallel oracle: ``workers`` processes, one origin each; a step = one frame for every
    worker's origin.  Returns (updates/s, ms per step)."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(workers)
    q = ctx.Queue()
    procs = [ctx.Process(target=_oracle_worker, args=(n_mols, 100 + w, steps + warmup, barrier, q, w)) for w in range(workers)]
    for p in procs:
        p.start()
    results = [q.get() for _ in procs]
    for p in procs:
        p.join()
    per_step = [max(r[1][i] for r in results) for i in range(warmup, warmup + steps)]
    total = sum(per_step)
    return workers * n_mols * steps / total, 1e3 * total / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    n_mols = args.molecules
    # bounded sample: every step is one frame of `workers` origins (of the 200) at full N
    steps, warmup = min(args.steps, 6), min(args.warmup, 1)
    value, ms = cpu_arm(n_mols, workers, steps, warmup)
    sample = (
        f"{steps} timed frames x {workers} of {N_ORIGINS} origins at N={n_mols} "
        f"(one oracle process per origin, {workers} processes); per-origin cost is independent of the other origins"
    )
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "molecules": n_mols, "origins": N_ORIGINS, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of the reference (numpy restatement; freud/rowan/gsd are not installable here)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    __graft_entry__.build()

    from sdanalysis_b200 import _lib
    from sdanalysis_b200.distributed import ShardedDynamics, partition
    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import SyntheticTrimer

    n, n_origins = args.molecules, args.origins
    steps, warmup = args.steps, max(args.warmup, 3)
    all_keys = list(range(n_origins))
    local_keys = partition(all_keys, rank, world)

    from sdanalysis_b200.synthetic import box_for

    def make_engine():
        return DynamicsEngine(n, box_for(n), WAVE_NUMBER, mol_sites=Trimer().positions, device=device, ring_depth=PIPELINE + 2)

    # after the timed region a few more steps run with per-kernel-group CUDA events (sdb_profile_*) for the
    # roofline: with the events in place the library serialises the kernels it otherwise overlaps
    prof_steps = max(2, min(steps, 8))
    n_frames = warmup + steps + prof_steps

    def phase(host_frames):
        """Create the origins (untimed), then W warm-up + K timed steps.  Returns timing info."""
        engine = make_engine()
        sharded = None
        if world > 1:
            def step_fn(ts, pos, quat, keys, rows_out=0):
                return engine.step(ts, pos, quat, keys, to_host=False, rows_out=rows_out)

            sharded = ShardedDynamics(n, step_fn, max_local=len(local_keys) + 1, device=device, batch=n_frames + 1, direct_rows=True)
        # origin creation: origin k starts at frame k (linear schedule, keyframe_interval = 1)
        timestep = 0
        # frames: a random walk generated on rank 0's GPU
        # (several GPUs: every rank runs the same seeded walk on its own device, so that the frames of the
        # device-resident phase can be held sharded by rows -- rank r keeps rows slice_bounds(r) of each frame)
        walk = SyntheticTrimer(n, seed=4, sigma_r=SIGMA_R, sigma_theta=SIGMA_T, device=device)
        sharded_resident = (
            world > 1 and not host_frames and sharded.exchange == "multicast" and bool(getattr(sharded, "_mc_ptr", 0))
            and n % (4 * world) == 0 and os.environ.get("SDB_BENCH_RESIDENT", "sharded") == "sharded"
        )
        for k in range(n_origins):
            active = all_keys[: k + 1]
            if world > 1:
                h = sharded.post_frame(*(walk.frame() if rank == 0 else (None, None)))
                sharded.step(timestep, active, h, gather=False)
            else:
                pos, quat = walk.frame()
                engine.step(timestep, pos, quat, active, want_sums=False)
            walk.advance(args.age_stride)
            timestep += args.age_stride
        torch.cuda.synchronize()
        # the frames of the measured steps
        frames = []
        shared_host = host_frames and world > 1
        if shared_host:
            # e2e with several GPUs: the frames sit in page-locked HOST memory that every rank can read
            # (a shared mapping, like a memory-mapped trajectory file); each rank uploads its 1/R slice
            # over its own PCIe link inside the timed region and multicasts it to all ranks over NVLink
            path = f"/dev/shm/sdb_bench_{os.environ.get('MASTER_PORT', '0')}.bin"
            per_frame = 7 * n
            if rank == 0:
                mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(n_frames, per_frame))
                for i in range(n_frames):
                    pos, quat = walk.frame()
                    walk.advance(1)
                    mm[i, : 3 * n] = pos.cpu().numpy().reshape(-1)
                    mm[i, 3 * n :] = quat.cpu().numpy().reshape(-1)
                mm.flush()
                del mm
            dist.barrier()
            mm = np.load(path, mmap_mode="r+")
            shared = torch.from_numpy(mm)
            rc = torch.cuda.cudart().cudaHostRegister(shared.data_ptr(), shared.numel() * 4, 0)
            assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
            frames = [(shared[i, : 3 * n].view(n, 3), shared[i, 3 * n :].view(n, 4)) for i in range(n_frames)]
        elif sharded_resident:
            lo, hi = sharded.slice_bounds()
            for _ in range(n_frames):
                pos, quat = walk.frame()
                walk.advance(1)
                frames.append((pos[lo:hi].clone(), quat[lo:hi].clone()))
        elif rank == 0:
            for _ in range(n_frames):
                pos, quat = walk.frame()
                walk.advance(1)
                if host_frames:
                    frames.append((pos.cpu().pin_memory(), quat.cpu().pin_memory()))
                else:
                    frames.append((pos.clone(), quat.clone()))
        else:
            frames = [(None, None)] * n_frames
        torch.cuda.synchronize()

        results = []

        posted = {}

        def one_step(i, ts):
            if world > 1:
                # software pipeline: the exchange of frame i + 1 is enqueued (side stream) before the
                # kernels of frame i, so it overlaps them
                post = sharded.post_frame_sliced if (shared_host or sharded_resident) else sharded.post_frame
                if i not in posted:
                    posted[i] = post(*frames[i])
                if i + 1 < n_frames:
                    posted[i + 1] = post(*frames[i + 1])
                # rows stay on the device; one gather per batch of frames (and at the end of the region)
                sharded.step(ts, all_keys, posted.pop(i), gather="batch")
                return
            pos, quat = frames[i]
            results.append(engine.step(ts, pos, quat, all_keys))
            if len(results) > PIPELINE:
                results.pop(0).table()

        def drain():
            last = None
            if world > 1:
                flushed = sharded.flush()
                if flushed:
                    keys, timediffs, sums = flushed[-1]
                    last = engine.finalize_table(sums, timediffs)
                return last
            while results:
                last = results.pop(0).table()
            return last

        for i in range(warmup):
            one_step(i, timestep)
            timestep += 1
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        lib = _lib.load()
        launches0 = _lib.launch_count()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        cuprof = os.environ.get("SDB_BENCH_CUPROFILE") == "1" and not host_frames
        if cuprof:  # `ncu --profile-from-start off`: the launch list then holds exactly the timed region
            torch.cuda.cudart().cudaProfilerStart()
        t_start.record()
        host_t0 = time.perf_counter()
        pyprof = None
        if os.environ.get("SDB_BENCH_PYPROFILE") == "1" and rank == 0:
            import cProfile

            pyprof = cProfile.Profile()
            pyprof.enable()
        shard_prof = os.environ.get("SDB_SHARD_PROFILE") == "1"
        step_events, host_marks = [], []
        for i in range(steps):
            one_step(warmup + i, timestep)
            timestep += 1
            if shard_prof:  # diagnostics: device time and host time of every step
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                step_events.append(ev)
                host_marks.append(time.perf_counter())
        if pyprof is not None:
            import pstats

            pyprof.disable()
            pstats.Stats(pyprof, stream=sys.stderr).sort_stats("tottime").print_stats(35)
        host_ms = 1e3 * (time.perf_counter() - host_t0) / steps
        drain_t0 = time.perf_counter()
        last_rows = drain()
        t_end.record()
        torch.cuda.synchronize()
        if shard_prof and len(step_events) > 2:
            dev_ms = sorted(a.elapsed_time(b) for a, b in zip(step_events[:-1], step_events[1:]))
            host_ms_l = sorted(1e3 * (b - a) for a, b in zip(host_marks[:-1], host_marks[1:]))
            print(f"[rank {rank}] host_frames={host_frames} per-step device interval ms: median {dev_ms[len(dev_ms)//2]:.3f} "
                  f"p90 {dev_ms[int(0.9*len(dev_ms))]:.3f} max {dev_ms[-1]:.3f}; host interval: median {host_ms_l[len(host_ms_l)//2]:.3f} "
                  f"p90 {host_ms_l[int(0.9*len(host_ms_l))]:.3f} max {host_ms_l[-1]:.3f}; first-step-to-start {t_start.elapsed_time(step_events[0]):.3f} "
                  f"drain (host) {1e3*(time.perf_counter()-drain_t0):.3f} ms, last step to end {step_events[-1].elapsed_time(t_end):.3f} ms",
                  file=sys.stderr, flush=True)
        if cuprof:
            torch.cuda.cudart().cudaProfilerStop()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        ms = t_start.elapsed_time(t_end)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches = _lib.launch_count() - launches0
        import ctypes

        # ---- kernel timing pass (untimed for `value`): per-kernel-group CUDA events, kernels serialised
        lib.sdb_profile_collect(None, None, None)  # drop anything collected so far
        lib.sdb_profile_enable(1)
        for i in range(prof_steps):
            one_step(warmup + steps + i, timestep)
            timestep += 1
        drain()
        torch.cuda.synchronize()
        group_ms = (ctypes.c_double * 4)()
        n_steps, n_origin_steps = ctypes.c_int64(0), ctypes.c_int64(0)
        lib.sdb_profile_enable(0)
        _lib.check(lib.sdb_profile_collect(group_ms, ctypes.byref(n_steps), ctypes.byref(n_origin_steps)))
        if world > 1:
            dist.barrier()
        count = max(int(n_steps.value), 1)
        kern = {
            "dyn": group_ms[1], "struct": group_ms[2], "overlap": group_ms[0] + group_ms[3], "count": count,
            "origins": int(n_origin_steps.value) // count,
        }
        if sharded is not None and sharded.stalls:
            import time as _t
            st = [a.elapsed_time(b) for a, b in sharded.stalls[-steps:]]
            xt = [a.elapsed_time(b) for a, b in sharded.exchange_times[-steps:]]
            print(f"[rank {rank}] host_frames={host_frames} frame-wait stall per step: mean {sum(st)/len(st):.3f} ms max {max(st):.3f} ms; "
                  f"exchange (side stream) mean {sum(xt)/max(len(xt),1):.3f} ms max {max(xt):.3f}\n", file=sys.stderr, flush=True)
        h2d, d2h = engine.h2d_bytes, engine.d2h_bytes
        if world > 1:  # every rank uploads its slice of the frame (28 B per molecule in total); rows leave through rank 0
            h2d = 28 * n if host_frames else 0
            d2h = n_origins * 16 * 8
        if shared_host:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(shared.data_ptr())
            del frames, shared, mm
            dist.barrier()
            if rank == 0:
                os.unlink(path)
        del engine, sharded
        torch.cuda.empty_cache()
        if os.environ.get("SDB_SHARD_PROFILE") == "1":
            print(f"[rank {rank}] host enqueue time per step {host_ms:.3f} ms", file=sys.stderr, flush=True)
        return {"ms": ms, "clocks": clocks, "launches": launches, "kern": kern, "h2d": h2d, "d2h": d2h, "rows": last_rows,
                "sharded_resident": sharded_resident}

    dev = phase(host_frames=False)
    e2e = phase(host_frames=True)

    updates_per_step = n * n_origins
    value = updates_per_step * steps / (dev["ms"] * 1e-3)
    e2e_value = updates_per_step * steps / (e2e["ms"] * 1e-3)

    # ---- roofline of the dominant kernel (this rank's share) --------------------------------------
    peak, peak_kind = _peaks()
    a_local = dev["kern"]["origins"]
    dyn_ms = dev["kern"]["dyn"] / max(dev["kern"]["count"], 1)
    bytes_per_update = 49.0 + 56.0 / max(a_local, 1)  # SURVEY 8(d), linear (dense) schedule
    algo_bytes = bytes_per_update * a_local * n
    achieved = algo_bytes / (dyn_ms * 1e-3) / 1e9 if dyn_ms > 0 else 0.0
    roofline = {
        "kernel": "dyn_step_kernel (+ dyn_increments_kernel when the per-frame pass is used)",
        "timing": f"CUDA events around the kernel group in {prof_steps} extra steps right after the timed region (same data, "
                  "kernels serialised; in the timed region dyn_step, struct and the overlap resolve of different origin groups overlap)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": _ncu_traffic("dyn_step_kernel", a_local * n),
        "traffic_frac": (
            (_ncu_traffic("dyn_step_kernel", a_local * n) or 0.0) / (dyn_ms * 1e-3) / 1e9 / peak if dyn_ms > 0 else None
        ),  # DRAM bytes actually moved / time / peak: what the kernel really draws from HBM
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/ncu_traffic.json), "
                          "scaled by the updates of this launch",
        "algorithmic_bytes_per_update": bytes_per_update,
        "updates_per_launch": a_local * n,
        "launch_ms": dyn_ms,
        "peak_source": peak_kind,
        "other_kernels_ms_per_step": {
            "struct_kernel": dev["kern"]["struct"] / max(dev["kern"]["count"], 1),
            "overlap_sample_bracket_resolve": dev["kern"]["overlap"] / max(dev["kern"]["count"], 1),
        },
    }
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample (rank 0, N=1 runs only) -----------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample_steps = 4
        v, _ = cpu_arm(n, 1, sample_steps, 1)
        cpu = {
            "value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample_steps} (frame, origin) pairs at N={n}: oracle Dynamics.compute_all + Relaxations.add, 1 process",
        }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev["ms"] / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "molecules": n, "origins": n_origins, "origins_per_gpu": len(local_keys),
            "parallelism": (
                f"origin-sharded x{world}, frame pushed to every rank by multicast stores over NVSwitch "
                + ("(device-resident phase: every rank holds 1/R of the rows of each frame in HBM and multicasts them), "
                   if dev["sharded_resident"] else "(from the reader rank), ")
                + "rows gathered once per batch (NCCL)"
                if world > 1 else "single GPU"
            ),
            "l2": "inputs larger than L2 (>= 25 MB of state per origin and step)",
        },
        "e2e": {
            "value": e2e_value, "unit": UNIT, "ms_per_step": e2e["ms"] / steps,
            "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
        },
        "gpu_launches": dev["launches"],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": dev["clocks"],
        "check": {k: float(dev["rows"][k][0]) for k in ("time", "msd", "alpha", "rot1", "overlap", "struct")}
        if isinstance(dev["rows"], dict) else None,
        # the end-to-end phase replays the same frames from host memory: its rows must be identical
        "e2e_rows_identical": bool(
            isinstance(dev["rows"], dict) and isinstance(e2e["rows"], dict)
            and all(np.array_equal(dev["rows"][k], e2e["rows"][k], equal_nan=True) for k in dev["rows"] if k != "keyframe")
        ),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--molecules", type=int, default=N_MOLS)
    ap.add_argument("--origins", type=int, default=N_ORIGINS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--age-stride", type=int, default=1, help="timesteps between the creation of consecutive origins")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

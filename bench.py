#!/usr/bin/env python
"""Benchmark of the per-frame dynamics hot path (BASELINE.json: molecule*origin updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--regime aged|early|both]

Workload (BASELINE.json configs[3], the configuration the target is quoted on): synthetic 2D Trimer
liquid, 1,000,000 molecules, 200 time-origins, dense (linear) schedule -- every frame updates all 200
origins (Dynamics.compute_all + Relaxations.add each).  A *step* is one frame: 200 x 1e6
molecule*origin updates.  With N GPUs the 200 origins are sharded round-robin over the ranks and each
frame reaches every rank over NVLink (strong scaling: total work is fixed).

Two regimes of that workload are timed (``regimes`` in the JSON line; ``--headline`` picks ``value``):

* ``aged``   (headline) origins created 10 timesteps apart, i.e. aged 0 ... 2000 steps like in the last
             frames of the 2000-frame run: |dr| is of the order of the relaxation distance, `struct` ~ 0.5,
             passage trackers fire continuously.
* ``early``  origins aged 1 ... 225 steps (round 1's only regime): almost nothing is near a threshold.  Its frames
             are 0.1 timestep apart so that the origins stay young while >= 1 s of regions is timed.

* ``value``  frames already resident in HBM when the timed region starts.  The K-step region (barrier +
             synchronize on both sides) is repeated until >= 1 s has been timed; ``value`` is the median.
* ``e2e``    the same loop fed from a GSD trajectory FILE through the package's reader (mmap -> pinned
             staging -> H2D copy per frame) and the public engine API, rows read back to the host.
* ``roofline``      dominant kernel (``dyn_step_kernel``) timed with CUDA events on its stream.
* ``parity``        rows and per-molecule state of two aged origins of THIS run against the CPU oracle.
* ``cpu_baseline``  the CPU oracle (a numpy restatement of the reference; the reference itself cannot
             run here: freud/rowan/gsd are not installed) on the same frames, one core.
* ``configs``       BASELINE.json configs[0] (the bundled-liquid trajectory through ``read.process_file``, the oracle
             on the whole run beside it, every row compared), configs[1] (32 k molecules, log-spaced and linear
             schedules) and configs[2] (neighbours + orientational order + num_neighbours per frame) as
             sub-records (1 GPU only).
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
# Distinct consecutive frames of the measured steps: as many as a default run uses, so that no frame is seen twice
# (a trajectory never runs backwards: walking a short pool back and forth was tried and makes the top-10 %
# thresholds of young origins jump at every reversal, which sends their `overlap` selection to the exact kernel
# for a few frames -- at 25 origins per GPU that doubled the step time of one region in three).  If the memory
# budgets below cap the pool it is walked back and forth after all (4 M-molecule runs, the e2e file).
# The frames of the `early` regime are 0.1 timestep apart, so that its origins stay young for the >= 1 s of timed
# regions (650 frames = 65 timesteps); the `aged` regime ages by one timestep per frame like the real run.
POOL = int(os.environ.get("SDB_BENCH_POOL", "1300"))
FRAME_DT = {"aged": 1.0, "early": 0.1}
POOL_DEVICE_BYTES = float(os.environ.get("SDB_BENCH_POOL_GB", "36")) * 1e9
POOL_FILE_BYTES = float(os.environ.get("SDB_BENCH_FILE_GB", "4")) * 1e9
AGE_STRIDE = {"aged": 10, "early": 1}
WORKLOAD = (
    "configs[3]: synthetic 2D Trimer liquid, 1M molecules, 200 time-origins, dense linear schedule "
    "(all origins active every frame), Dynamics.compute_all + Relaxations.add"
)


def _peaks():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(regime, kernel, updates):
    """DRAM bytes per launch of ``kernel`` from the committed ncu capture of ``regime``, scaled to ``updates``."""
    path = REPO / "profiles" / "ncu_traffic.json"
    try:
        rec = json.loads(path.read_text())[regime][kernel]
        return (rec["dram_bytes_read"] + rec["dram_bytes_write"]) * updates / rec["updates_in_launch"], rec.get("source")
    except (OSError, KeyError, ValueError):
        return None, None


def pingpong(i, p):
    """0, 1, ..., p-1, p-2, ..., 1, 0, 1, ...: consecutive frames of the pool are one walk step apart."""
    if p <= 1:
        return 0
    period = 2 * (p - 1)
    j = i % period
    return j if j < p else period - j


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
    # every step is one frame of `workers` origins (of the 200) at full N: a bounded sample of the workload,
    # per-origin cost being independent of the other origins; --steps / --warmup as given
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
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
    import ctypes

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
    from sdanalysis_b200.gsdio import GSDWriter, open_gsd
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import SyntheticTrimer, box_for

    lib = _lib.load()
    n, n_origins = args.molecules, args.origins
    steps, warmup = args.steps, max(args.warmup, 3)
    all_keys = list(range(n_origins))
    local_keys = partition(all_keys, rank, world)
    peak, peak_kind = _peaks()
    pool_size = max(2, min(POOL, int(POOL_DEVICE_BYTES // (28 * n))))
    # the e2e trajectory file: a RAM-backed directory when it has room (every rank maps the file), else /tmp, and
    # never more than a third of what is free there
    import shutil

    tmpdir, file_budget = args.tmpdir, POOL_FILE_BYTES
    for candidate in ([args.tmpdir] if args.tmpdir else []) + ["/dev/shm", "/tmp"]:
        try:
            free = shutil.disk_usage(candidate).free
        except OSError:
            continue
        if free >= 3 * min(POOL_FILE_BYTES, 64 * 28 * n):
            tmpdir, file_budget = candidate, min(POOL_FILE_BYTES, free / 3)
            break
    else:
        tmpdir = args.tmpdir or "/tmp"
        file_budget = min(POOL_FILE_BYTES, shutil.disk_usage(tmpdir).free / 3)
    file_frames = max(2, min(pool_size, int(file_budget // (28 * n))))
    if world > 1:  # rank 0 decides where the file goes and how long it is
        decided = [tmpdir, file_frames]
        dist.broadcast_object_list(decided, src=0)
        tmpdir, file_frames = decided
    prof_steps = max(2, min(steps, 8))
    want_parity = args.parity == "on" or (args.parity == "auto" and world == 1 and not args.no_cpu_baseline)
    if n_origins < 3:
        want_parity = False

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(value):
        if world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def summary(times):
        ordered = sorted(times)
        return {
            "regions": len(times), "median_ms": statistics.median(times), "min_ms": ordered[0], "max_ms": ordered[-1],
            "p10_ms": ordered[int(0.1 * (len(times) - 1))], "p90_ms": ordered[int(0.9 * (len(times) - 1) + 0.5)],
            "timed_seconds": sum(times) / 1e3,
        }

    def measure(regime, with_parity):
        """Create the origins of ``regime`` (untimed), then time the device-resident loop, the kernel groups and
        the end-to-end loop fed from a GSD file.  Returns a dict of measurements."""
        age_stride = AGE_STRIDE[regime]
        engine = DynamicsEngine(n, box_for(n), WAVE_NUMBER, mol_sites=Trimer().positions, device=device, ring_depth=PIPELINE + 2)
        sharded = None
        if world > 1:
            def step_fn(ts, pos, quat, keys, rows_out=0):
                return engine.step(ts, pos, quat, keys, to_host=False, rows_out=rows_out)

            # (max_local must be the same on every rank: symmetric buffers are sized with it)
            sharded = ShardedDynamics(
                n, step_fn, max_local=(n_origins + world - 1) // world + 1, device=device, batch=max(steps, warmup, prof_steps) + 1,
                direct_rows=os.environ.get("SDB_BENCH_DIRECT_ROWS", "1") == "1",
            )
        # every rank runs the same seeded walk on its own device (the frames of the device-resident phase are
        # held sharded by rows when there are several ranks)
        walk = SyntheticTrimer(n, seed=4, sigma_r=SIGMA_R, sigma_theta=SIGMA_T, device=device)
        box = walk.box
        sparse = {0, 1} if with_parity else set()  # the two origins followed by the oracle see few frames while they age
        oracle_frames = []  # (timestep, pos, quat, origins of `sparse` taking part) in order, host copies on rank 0
        timestep = 0
        for k in range(n_origins):
            last = k == n_origins - 1
            active = [j for j in range(k + 1) if j not in sparse or j == k or last]
            pos, quat = walk.frame()
            if world > 1:
                h = sharded.post_frame(*((pos, quat) if rank == 0 else (None, None)))
                sharded.step(timestep, active, h, gather=False)
            else:
                engine.step(timestep, pos, quat, active, want_sums=False)
            followed = [j for j in active if j in sparse]
            if followed and rank == 0:
                oracle_frames.append((timestep, pos.cpu().numpy(), quat.cpu().numpy(), followed))
            walk.advance(age_stride)
            timestep += age_stride
        torch.cuda.synchronize()
        t_first = timestep  # timestep of measured step 0

        # ---- the frames of the measured steps: a pool walked back and forth ---------------------------------
        sharded_resident = (
            world > 1 and sharded.exchange == "multicast" and bool(getattr(sharded, "_mc_ptr", 0))
            and os.environ.get("SDB_BENCH_RESIDENT", "sharded") == "sharded"
        )
        lo, hi = sharded.slice_bounds() if sharded is not None else (0, n)
        pool_dev, pool_host = [], []
        tag = os.environ.get("MASTER_PORT", str(os.getpid())) if world > 1 else str(os.getpid())
        gsd_path = Path(tmpdir) / f"sdb_bench_{tag}_{regime}.gsd"
        writer = GSDWriter(gsd_path) if (rank == 0 and not args.no_e2e) else None
        for j in range(pool_size):
            pos, quat = walk.frame()
            if sharded_resident:
                pool_dev.append((pos[lo:hi].clone(), quat[lo:hi].clone()))
            elif world == 1 or rank == 0:
                pool_dev.append((pos.clone(), quat.clone()))
            else:
                pool_dev.append((None, None))
            if rank == 0 and ((writer is not None and j < file_frames) or (with_parity and j <= warmup)):
                hp, hq = pos.cpu().numpy(), quat.cpu().numpy()
                if writer is not None and j < file_frames:
                    writer.append(t_first + j, box, hp, hq, dimensions=2)
                if with_parity and j <= warmup:
                    pool_host.append((hp, hq))
            walk.advance(FRAME_DT[regime])
        if writer is not None:
            writer.close()
        torch.cuda.synchronize()
        barrier()

        # ---- stepping ---------------------------------------------------------------------------------------
        hostprof = {} if os.environ.get("SDB_BENCH_HOSTPROF") == "1" else None  # diagnostics: where the host loop waits
        state = {"i": 0}  # index of the next measured step
        results, tables, posted = [], [], {}
        keep_tables = warmup + 1 if with_parity else 0
        trj = {"gsd": None}

        def frame_of(source, i):
            if source == "gsd":
                snap = trj["gsd"][pingpong(i, file_frames)]
                return snap.position, snap.orientation
            return pool_dev[pingpong(i, pool_size)]

        def collect(item):
            i, res = item
            table = res.table()
            if i < keep_tables:
                tables.append((i, table))
            return table

        def one_step(source):
            i = state["i"]
            state["i"] = i + 1
            ts = t_first + i
            if world > 1:
                # software pipeline: the exchange of frame i + 1 is enqueued (side stream) before the kernels of
                # frame i, so it overlaps them
                post = sharded.post_frame_sliced if (source == "gsd" or sharded_resident) else sharded.post_frame
                if i not in posted:
                    posted[i] = post(*frame_of(source, i))
                posted[i + 1] = post(*frame_of(source, i + 1))
                sharded.step(ts, all_keys, posted.pop(i), gather="batch")
                return
            if hostprof is not None:
                h0 = time.perf_counter()
                pos, quat = frame_of(source, i)
                h1 = time.perf_counter()
                results.append((i, engine.step(ts, pos, quat, all_keys)))
                h2 = time.perf_counter()
                if len(results) > PIPELINE:
                    collect(results.pop(0))
                h3 = time.perf_counter()
                rec = hostprof.setdefault(source, [0.0, 0.0, 0.0, 0])
                rec[0] += h1 - h0
                rec[1] += h2 - h1
                rec[2] += h3 - h2
                rec[3] += 1
                return
            pos, quat = frame_of(source, i)
            results.append((i, engine.step(ts, pos, quat, all_keys)))
            if len(results) > PIPELINE:
                collect(results.pop(0))

        def drain():
            last = None
            if world > 1:
                flushed = sharded.flush()
                first_i = state["i"] - len(flushed)
                for off, (keys, timediffs, sums) in enumerate(flushed):
                    if first_i + off < keep_tables or off == len(flushed) - 1:
                        last = engine.finalize_table(sums, timediffs)
                        last["keyframe"] = list(keys)
                        if first_i + off < keep_tables:
                            tables.append((first_i + off, last))
                return last
            while results:
                last = collect(results.pop(0))
            return last

        def block(source, count):
            """`count` steps bracketed by barrier + synchronize on both sides; ms as the max over ranks."""
            barrier()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            h0 = time.perf_counter()
            for _ in range(count):
                one_step(source)
            host_ms = 1e3 * (time.perf_counter() - h0) / count
            last = drain()
            t1.record()
            torch.cuda.synchronize()
            return max_over_ranks(t0.elapsed_time(t1)), host_ms, last

        def timed(source, min_seconds):
            """Repeat the K-step region until >= min_seconds have been timed (at least 3, at most 400 regions)."""
            times, host = [], []
            last = None
            while True:
                ms, host_ms, last = block(source, steps)
                times.append(ms)
                host.append(host_ms)
                if (len(times) >= 3 and sum(times) >= 1e3 * min_seconds) or len(times) >= 400:
                    break
            return times, host, last

        # device-resident phase
        block("dev", warmup)
        snapshots = {}
        if with_parity:
            torch.cuda.synchronize()
            for key in sparse:
                if key in local_keys:
                    org = engine.origins[key]
                    snapshots[key] = (org.delta.clone(), org.flags.clone(), org.status.clone())
        launches0 = _lib.launch_count()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        cuprof = os.environ.get("SDB_BENCH_CUPROFILE") == "1"
        if cuprof:  # `ncu --profile-from-start off`: the capture then holds the timed regions only
            torch.cuda.cudart().cudaProfilerStart()
        dev_times, dev_host, last_table = timed("dev", args.min_seconds)
        if cuprof:
            torch.cuda.cudart().cudaProfilerStop()
        clocks = sampler.stop() if sampler else None
        launches = (_lib.launch_count() - launches0) // len(dev_times)
        # one long region (no per-region synchronisation): what a long run sustains
        n_sustained = min(len(dev_times), 25) * steps
        sustained_ms, _, _ = block("dev", n_sustained)

        # kernel timing pass: per-kernel-group CUDA events, kernels serialised
        lib.sdb_profile_collect(None, None, None)
        lib.sdb_profile_enable(1)
        block("dev", prof_steps)
        group_ms = (ctypes.c_double * 4)()
        n_steps, n_origin_steps = ctypes.c_int64(0), ctypes.c_int64(0)
        lib.sdb_profile_enable(0)
        _lib.check(lib.sdb_profile_collect(group_ms, ctypes.byref(n_steps), ctypes.byref(n_origin_steps)))
        count = max(int(n_steps.value), 1)
        kern = {
            "dyn": group_ms[1] / count, "struct": group_ms[2] / count, "overlap": (group_ms[0] + group_ms[3]) / count,
            "origins": int(n_origin_steps.value) // count,
        }

        # end-to-end phase: the same loop fed from the GSD file
        e2e = None
        if not args.no_e2e:
            posted.clear()
            trj["gsd"] = open_gsd(gsd_path)
            pinned_file = bool(os.environ.get("SDB_BENCH_PIN_FILE", "1") == "1" and trj["gsd"].pin())
            block("gsd", warmup)
            e2e_times, e2e_host, _ = timed("gsd", args.min_seconds)
            h2d = engine.h2d_bytes if world == 1 else 28 * n
            d2h = engine.d2h_bytes if world == 1 else n_origins * 16 * 8
            e2e = {"times": e2e_times, "host_ms": statistics.median(e2e_host), "h2d": h2d, "d2h": d2h, "pinned_file": pinned_file,
                   "pin_error": getattr(trj["gsd"], "pin_error", None)}
            posted.clear()
            torch.cuda.synchronize()
            barrier()
            trj["gsd"].close()
        barrier()
        if rank == 0 and not args.no_e2e:
            try:
                os.unlink(gsd_path)
            except OSError:
                pass

        # ---- parity of two aged origins of this run against the CPU oracle (and the CPU baseline) ------------
        parity = cpu = None
        if with_parity:
            state_np = {}
            for key, (delta, flags, status) in snapshots.items():
                planes = delta[:, :n].cpu().numpy()
                trans = np.zeros((n, 3), dtype=np.float32)
                rot = np.zeros((n, 3), dtype=np.float32)
                trans[:, 0], trans[:, 1], rot[:, 0] = planes[0], planes[1], planes[2]
                st = status[:, :n].cpu().numpy().view(np.uint32)
                fl = flags[:n].cpu().numpy()
                summ = {}
                for t_i, spec in enumerate(engine.trackers):
                    col = st[t_i].astype(np.int64)
                    if spec.last_passage:
                        col[((fl >> spec.bit) & 3) != 2] = 2 ** 32 - 1
                    summ[spec.name] = col
                state_np[key] = (trans, rot, summ)
            if world > 1:
                gathered = [None] * world
                dist.gather_object(state_np, gathered if rank == 0 else None, dst=0)
                if rank == 0:
                    state_np = {}
                    for part in gathered:
                        state_np.update(part)
            if rank == 0:
                parity, cpu = _oracle_parity(np, n, box, oracle_frames, pool_host, t_first, warmup, tables, state_np, sorted(sparse))
        if sharded is not None and sharded.stalls:
            # diagnostics (SDB_SHARD_PROFILE=1): how long the compute stream waited for a frame, and how long the
            # exchange of a frame took on the side stream
            torch.cuda.synchronize()
            stalls = [a.elapsed_time(b) for a, b in sharded.stalls]
            xch = [a.elapsed_time(b) for a, b in sharded.exchange_times]
            half = len(stalls) // 2
            print(f"[rank {rank} {regime}] frame-wait stall per step: device phase mean {sum(stalls[:half]) / max(half, 1):.3f} ms, "
                  f"e2e phase mean {sum(stalls[half:]) / max(len(stalls) - half, 1):.3f} ms (max {max(stalls):.3f}); exchange on the side "
                  f"stream mean {sum(xch) / max(len(xch), 1):.3f} ms (max {max(xch):.3f}) over {len(stalls)} steps", file=sys.stderr, flush=True)
        if hostprof:
            for source, (t_frame, t_step, t_collect, count) in hostprof.items():
                print(f"[hostprof {regime} {source}] per step: frame lookup {1e3*t_frame/count:.3f} ms, engine.step {1e3*t_step/count:.3f} ms, "
                      f"collect (wait for step i-{PIPELINE}) {1e3*t_collect/count:.3f} ms over {count} steps", file=sys.stderr, flush=True)
        exchange = None if sharded is None else sharded.exchange
        exchange_note = "" if sharded is None else sharded.exchange_note
        del engine, sharded
        torch.cuda.empty_cache()
        return {
            "regime": regime, "dev_times": dev_times, "dev_host_ms": statistics.median(dev_host), "sustained_ms": sustained_ms,
            "n_sustained": n_sustained, "clocks": clocks, "launches": launches, "kern": kern, "e2e": e2e,
            "last_table": last_table, "parity": parity, "cpu": cpu, "sharded_resident": sharded_resident,
            "exchange": exchange, "exchange_note": exchange_note,
        }

    regimes = ["aged", "early"] if args.regime == "both" else [args.regime]
    if args.regime == "both" and args.headline == "early":
        regimes = ["early", "aged"]
    measured = {}
    for regime in regimes:
        measured[regime] = measure(regime, with_parity=want_parity and regime == regimes[0])
    head = measured[regimes[0]]
    updates_per_step = n * n_origins

    def regime_record(m):
        s = summary(m["dev_times"])
        a_local = m["kern"]["origins"]
        dyn_ms = m["kern"]["dyn"]
        bytes_per_update = 49.0 + 56.0 / max(a_local, 1)  # SURVEY 8(d), linear (dense) schedule
        achieved = bytes_per_update * a_local * n / (dyn_ms * 1e-3) / 1e9 if dyn_ms > 0 else 0.0
        traffic, source = _ncu_traffic(m["regime"], "dyn_step_kernel", a_local * n)
        step_frac = (49.0 + 56.0 / n_origins) * updates_per_step * steps / (s["median_ms"] * 1e-3) / 1e9 / peak / world
        rec = {
            "value": updates_per_step * steps / (s["median_ms"] * 1e-3),
            "ms_per_step": s["median_ms"] / steps,
            "regions": s,
            "sustained": {
                "steps": m["n_sustained"], "ms_per_step": m["sustained_ms"] / m["n_sustained"],
                "value": updates_per_step * m["n_sustained"] / (m["sustained_ms"] * 1e-3),
                "note": "one region of this many steps, synchronised only at its ends",
            },
            "host_enqueue_ms_per_step": m["dev_host_ms"],
            "roofline": {
                "kernel": "dyn_step_kernel (+ dyn_increments_kernel when the per-frame pass is used)",
                "timing": f"CUDA events around the kernel group in {prof_steps} extra steps right after the timed regions (same "
                          "state, kernels serialised; in the timed regions dyn_step, struct and the overlap resolve overlap)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_frac": (traffic / (dyn_ms * 1e-3) / 1e9 / peak) if (traffic and dyn_ms > 0) else None,
                "traffic_source": source,
                "algorithmic_bytes_per_update": bytes_per_update,
                "updates_per_launch": a_local * n,
                "launch_ms": dyn_ms,
                "peak_source": peak_kind,
                "other_kernels_ms_per_step": {"struct_kernel": m["kern"]["struct"], "overlap_bracket_resolve": m["kern"]["overlap"]},
                "step_algorithmic_frac": step_frac,
                "step_note": "whole step (all kernels" + (", all ranks" if world > 1 else "") + "): algorithmic bytes / time / peak"
                             + (" / ranks" if world > 1 else ""),
            },
            "check": {k: float(m["last_table"][k][0]) for k in ("time", "msd", "alpha", "rot1", "overlap", "struct", "com_struct")}
            if isinstance(m["last_table"], dict) else None,
        }
        if m["e2e"] is not None:
            es = summary(m["e2e"]["times"])
            rec["e2e"] = {
                "value": updates_per_step * steps / (es["median_ms"] * 1e-3), "unit": UNIT, "ms_per_step": es["median_ms"] / steps,
                "h2d_bytes_per_step": m["e2e"]["h2d"], "d2h_bytes_per_step": m["e2e"]["d2h"], "regions": es,
                "host_ms_per_step": m["e2e"]["host_ms"],
                "source": "GSD trajectory file read through sdanalysis_b200.gsdio (mmap) inside the timed region"
                          + (", file mapping page-locked (cudaHostRegister)" if m["e2e"]["pinned_file"] else
                             ", frames copied to the engine's pinned ring by its host threads"),
                "file_frames": file_frames, "pin_error": m["e2e"]["pin_error"],
            }
        return rec

    records = {r: regime_record(m) for r, m in measured.items()}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    head_rec = records[regimes[0]]

    # ---- CPU baseline -----------------------------------------------------------------------------------
    cpu = head["cpu"]
    if cpu is None and world == 1 and not args.no_cpu_baseline:
        sample_steps = 4
        v, _ = cpu_arm(n, 1, sample_steps, 1)
        cpu = {
            "value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample_steps} (frame, origin) pairs at N={n}: oracle Dynamics.compute_all + Relaxations.add, 1 process",
        }

    # ---- the other BASELINE configurations (one GPU) -------------------------------------------------------------
    configs = None
    if world == 1 and not args.no_configs:
        import bench_configs

        configs = bench_configs.run_all(device, peak, cpu_baseline=not args.no_cpu_baseline)

    if world > 1:
        if head["exchange"] == "multicast":
            parallelism = (
                f"origin-sharded x{world}; frame exchange: multicast stores over NVSwitch into every rank's frame ring"
                + (" (every rank holds 1/R of the rows of each frame and multicasts them)" if head["sharded_resident"]
                   else " (from the reader rank)")
                + "; rows written by every rank straight into rank 0's row buffer (peer copy enqueued by the native step)"
            )
        else:
            parallelism = (
                f"origin-sharded x{world}; frame exchange: torch.distributed.broadcast (NCCL) -- {head['exchange_note'] or 'requested'}; "
                "rows gathered once per region (NCCL)"
            )
    else:
        parallelism = "single GPU"

    keep = ("value", "ms_per_step", "regions", "sustained", "host_enqueue_ms_per_step", "e2e", "check")
    roof_keep = ("frac", "traffic_frac", "launch_ms", "achieved", "other_kernels_ms_per_step", "step_algorithmic_frac")
    line = {
        "metric": METRIC, "value": head_rec["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": head_rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "regime": regimes[0], "molecules": n, "origins": n_origins, "origins_per_gpu": len(local_keys),
            "parallelism": parallelism,
            "l2": "inputs larger than L2 (>= 25 MB of state per origin and step)",
            "timing": f"{steps}-step region (barrier + synchronize on both sides) repeated {head_rec['regions']['regions']} times "
                      f"= {head_rec['regions']['timed_seconds']:.2f} s; value = median region",
            "frames": f"{pool_size} consecutive frames resident in HBM (walked back and forth only if a run needs more), "
                      f"{FRAME_DT[regimes[0]]} timestep apart; e2e: a GSD file of the first {file_frames} of them",
        },
        "e2e": head_rec.get("e2e"),
        "gpu_launches": head["launches"],
        "roofline": head_rec["roofline"],
        "cpu_baseline": cpu,
        "clocks": head["clocks"],
        "check": head_rec["check"],
        "parity": head["parity"],
        "regimes": {
            r: {**{k: rec[k] for k in keep if k in rec}, "roofline": {k: rec["roofline"][k] for k in roof_keep}}
            for r, rec in records.items()
        },
        "configs": configs,
    }
    print(json.dumps(line), flush=True)
    parity_failed = head["parity"] is not None and not head["parity"]["ok"]
    if parity_failed:
        print("PARITY FAILURE: " + json.dumps(head["parity"]), file=sys.stderr, flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity_failed:
        raise SystemExit(3)


def _oracle_parity(np, n, box, oracle_frames, pool_host, t_first, warmup, tables, state_np, followed):
    """Replay the frames the two followed origins saw through the CPU oracle; compare the rows of the warm-up
    steps and of the first timed step, and the per-molecule state after the warm-up.  Also the CPU baseline:
    the oracle's time for exactly these (frame, origin) pairs."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import parity as opar

    run = opar.OracleRun(WAVE_NUMBER)
    report = {
        "checked_rows": 0, "max_rel": 0.0, "failed_rows": [], "origins": followed, "molecules": n,
        "delta_translation_mismatch": 0, "delta_rotation_differs": 0, "delta_rotation_beyond_1ulp": 0,
        "passage_mismatch": 0, "passage_excluded_1ulp": 0, "state_checked": [],
    }
    t0 = time.perf_counter()
    for ts, pos, quat, keys in oracle_frames:
        run.frame(ts, box, pos, quat, keys)
    creation_rows = len(run.rows)
    for i, (pos, quat) in enumerate(pool_host):
        run.frame(t_first + i, box, pos, quat, followed)
        if i == warmup - 1:  # the device state was copied after the warm-up steps
            seconds_before = time.perf_counter()
            for key in followed:
                if key in state_np:
                    trans, rot, summ = state_np[key]
                    dyn, relax = run.keyframes[key]
                    res = opar.compare_state(trans, rot, summ, dyn, relax.summary(), run.near[key])
                    for k_, v_ in res.items():
                        report[k_] += v_
                    report["state_checked"].append(key)
            t0 += time.perf_counter() - seconds_before  # (the comparison is not oracle work)
    seconds = time.perf_counter() - t0
    # rows: origin `key` at measured step i has time difference t_first + i - t0(key)
    by_step = dict(tables)
    t0_of = {key: run.keyframes[key][0].timestep for key in followed}
    for row in run.rows[creation_rows:]:
        key = row["keyframe"]
        i = int(row["time"]) + t0_of[key] - t_first
        table = by_step.get(i)
        if table is None:
            continue
        col = list(table["keyframe"]).index(key)
        got = {q: table[q][col] for q in row if q not in ("keyframe", "_tie")}
        max_rel, failed = opar.compare_row(got, row, n)
        report["checked_rows"] += 1
        report["max_rel"] = max(report["max_rel"], max_rel)
        if failed:
            report["failed_rows"].append({"origin": key, "step": i, "quantities": failed})
    report["passage_times_set_origin0"] = {
        name: int((col != 2 ** 32 - 1).sum()) for name, col in run.keyframes[followed[0]][1].summary().items()
    }
    report["ok"] = bool(
        report["checked_rows"] == len(followed) * (warmup + 1) and not report["failed_rows"]
        and report["delta_translation_mismatch"] == 0 and report["delta_rotation_beyond_1ulp"] == 0
        and report["delta_rotation_differs"] <= max(1, int(1e-5 * n * len(report["state_checked"])))
        and report["passage_mismatch"] == 0 and len(report["state_checked"]) == len(followed)
    )
    report["bar"] = (
        "rows rtol 1e-5 (counts exact, struct +-2 rows); dr bit-exact; dtheta <= 1 ulp on <= 1e-5 of the molecules; "
        "passage times bit-exact outside 1 ulp of a cutoff"
    )
    cpu = {
        "value": run.pairs * n / seconds, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"{run.pairs} (frame, origin) pairs at N={n} -- the frames two aged origins of this run saw: oracle "
                  f"Dynamics.compute_all + Relaxations.add, 1 process, {seconds:.1f} s",
    }
    return report, cpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--molecules", type=int, default=N_MOLS)
    ap.add_argument("--origins", type=int, default=N_ORIGINS)
    ap.add_argument("--regime", default="both", choices=["aged", "early", "both"])
    ap.add_argument("--headline", default="aged", choices=["aged", "early"], help="with --regime both: which regime is `value`")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="repeat the K-step region until this much has been timed")
    ap.add_argument("--parity", default="auto", choices=["auto", "on", "off"], help="oracle check of two origins (auto: 1 GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[1] / configs[2] sub-records")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tmpdir", default=os.environ.get("SDB_BENCH_TMP"), help="directory of the e2e trajectory file (default: /dev/shm or /tmp, whichever has room)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

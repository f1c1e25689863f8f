"""Sub-records of bench.py for the BASELINE.json configurations other than configs[3] (one GPU):

* ``0_small``  configs[0], the reference's own CPU-runnable case.  The bundled trajectory blob is missing from the
  reference checkout, so the bundled liquid dump (1250 molecules, ``tests/golden/liquid.npz``) is
  frame 0 of a seeded Brownian trajectory on ``GenerateStepSeries(10000, 100, 1000, 500)`` (1901 frames, 11
  origins, 2361 pairs), written as ``trajectory-Trimer-P13.50-T3.00.gsd`` and analysed by the public
  ``read.process_file``; the oracle runs the WHOLE trajectory beside it and every row is compared.
* ``1_exp`` / ``1_lin``  configs[1]: 32 768 molecules, 100 origins, 1000 frames; the reference's default
  log-spaced schedule ``GenerateStepSeries(1000, 10, 10, 100)`` (2405 (frame, origin) pairs) and the dense
  linear schedule (a new origin every 10 frames, 50 500 pairs).  ``value``: whole run with the frames
  resident in HBM, through ``DynamicsEngine.run`` (the native frame loop); ``e2e``: the public
  ``read.process_file`` on a GSD file of the same trajectory (file -> DataFrame).
* ``2_order``  configs[2]: ``order.compute_neighbours`` + ``orientational_order`` + ``num_neighbours`` for
  every frame of the 32 k trajectory, through the public functions.

Each record carries a roofline fraction from SURVEY 8(d)'s byte model and a CPU-oracle baseline on a bounded
sample of the same frames.
"""

import ctypes
import os
import statistics
import time
from pathlib import Path

import numpy as np

WAVE_NUMBER = 2.9
N32K = 32768
SIGMA_R, SIGMA_T = 0.01, 0.02


def _schedules():
    from sdanalysis_b200.synthetic import exponential_schedule, linear_schedule

    return {
        "1_exp": list(exponential_schedule(1000, 10, 10, 100)),
        "1_lin": list(linear_schedule(1000, 10, 99)),
    }


def _device_frames(schedule, n, device, seed):
    from sdanalysis_b200.synthetic import SyntheticTrimer

    walk = SyntheticTrimer(n, seed=seed, sigma_r=SIGMA_R, sigma_theta=SIGMA_T, device=device)
    frames = []
    for timestep, _ in schedule:
        walk.advance(timestep - walk.timestep)
        frames.append(walk.frame())
    return walk.box, frames


def _dynamics_config(name, schedule, device, peak, cpu_baseline, tmpdir):
    import torch

    from sdanalysis_b200 import _lib
    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.gsdio import GSDWriter
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.read import process_file

    lib = _lib.load()
    n = N32K
    box, frames = _device_frames(schedule, n, device, seed=2)
    pairs = sum(len(ix) for _, ix in schedule)
    updates = pairs * n
    linear = name == "1_lin"
    # 49 + 56/A (dense) or 65 + 28/A (log-spaced) bytes per update, A = active origins of the frame
    algo_bytes = sum(n * ((49 * len(ix) + 56) if linear else (65 * len(ix) + 28)) for _, ix in schedule if ix)

    def one_run(profile=False):
        engine = DynamicsEngine(n, box, WAVE_NUMBER, mol_sites=Trimer().positions, device=device, ring_depth=4)
        batch = [(ts, pos, quat, ix) for (ts, ix), (pos, quat) in zip(schedule, frames)]
        torch.cuda.synchronize()
        if profile:
            lib.sdb_profile_collect(None, None, None)
            lib.sdb_profile_enable(1)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.launch_count()
        t0.record()
        rows = 0
        for first in range(0, len(batch), 256):  # one native call per 256 frames
            for res in engine.run(batch[first : first + 256]):
                rows += len(res.keys)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        out = {"ms": ms, "rows": rows, "launches": _lib.launch_count() - launches0}
        if profile:
            group_ms = (ctypes.c_double * 4)()
            lib.sdb_profile_enable(0)
            _lib.check(lib.sdb_profile_collect(group_ms, None, None))
            out["dyn_ms"], out["struct_ms"], out["overlap_ms"] = group_ms[1], group_ms[2], group_ms[0] + group_ms[3]
        del engine
        return out

    one_run()  # warm-up (allocations, first launches)
    runs = [one_run() for _ in range(5)]
    ms = statistics.median(r["ms"] for r in runs)
    assert runs[0]["rows"] == pairs
    prof = one_run(profile=True)
    record = {
        "workload": f"configs[1] {name[2:]}: {n} molecules, {len(schedule)} frames, {pairs} (frame, origin) pairs, "
                    + ("dense linear schedule (new origin every 10 frames, 100 origins)" if linear else
                       "log-spaced schedule GenerateStepSeries(1000, 10, 10, 100)"),
        "value": updates / (ms * 1e-3), "unit": "updates/s", "ms_per_frame": ms / len(schedule), "ms_per_run": ms,
        "runs_ms": [r["ms"] for r in runs], "best_value": updates / (min(r["ms"] for r in runs) * 1e-3),
        "gpu_launches": runs[0]["launches"],
        "roofline": {
            "bound": "hbm", "kernel": "dyn_step_kernel", "launch_ms_total": prof["dyn_ms"],
            "achieved": algo_bytes / (prof["dyn_ms"] * 1e-3) / 1e9 if prof["dyn_ms"] > 0 else None, "peak": peak, "unit": "GB/s",
            "frac": algo_bytes / (prof["dyn_ms"] * 1e-3) / 1e9 / peak if prof["dyn_ms"] > 0 else None,
            "run_frac": algo_bytes / (ms * 1e-3) / 1e9 / peak,
            "algorithmic_bytes_per_run": algo_bytes,
            "other_kernels_ms_total": {"struct": prof["struct_ms"], "overlap": prof["overlap_ms"]},
            "note": "frac: dominant kernel alone (CUDA events, profiled run); run_frac: whole run including launch gaps",
        },
    }

    # ---- e2e: the public process_file on a GSD file of the same trajectory -----------------------------------
    path = Path(tmpdir) / f"sdb_bench_cfg_{os.getpid()}_{name}.gsd"
    with GSDWriter(path) as writer:
        for (ts, _), (pos, quat) in zip(schedule, frames):
            writer.append(ts, box, pos.cpu().numpy(), quat.cpu().numpy(), dimensions=2)
    kwargs = dict(linear_steps=None, keyframe_interval=10, keyframe_max=99) if linear else dict(
        steps_max=1000, linear_steps=10, keyframe_interval=10, keyframe_max=100
    )
    try:
        process_file(path, WAVE_NUMBER, **kwargs)  # warm-up
        times = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            df = process_file(path, WAVE_NUMBER, **kwargs)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        assert len(df) == pairs, (len(df), pairs)
        sec = statistics.median(times)
        record["e2e"] = {
            "value": updates / sec, "unit": "updates/s", "ms_per_frame": 1e3 * sec / len(schedule), "rows": len(df),
            "h2d_bytes_per_frame": 28 * n, "source": "read.process_file(GSD file) -> pandas.DataFrame, wall clock",
        }
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass

    # ---- CPU oracle on a bounded sample of the same frames ----------------------------------------------------
    if cpu_baseline:
        from oracle import parity as opar

        run = opar.OracleRun(WAVE_NUMBER)
        host = [(ts, ix, pos.cpu().numpy(), quat.cpu().numpy()) for (ts, ix), (pos, quat) in list(zip(schedule, frames))[:40]]
        t0 = time.perf_counter()
        for ts, ix, pos, quat in host:
            run.frame(ts, box, pos, quat, [k for k in ix if k == 0])
        sec = time.perf_counter() - t0
        record["cpu_baseline"] = {
            "value": run.pairs * n / sec, "unit": "updates/s", "cores": 1, "kind": "port",
            "sample": f"origin 0 over the first {len(host)} frames ({run.pairs} pairs) of the same trajectory, 1 process",
        }
    del frames
    torch.cuda.empty_cache()
    return record


def _order_config(device, peak, cpu_baseline):
    import torch

    from sdanalysis_b200 import _lib, order

    n = N32K
    schedule = _schedules()["1_exp"][:400]  # the same trajectory; 400 of its 1001 frames bound the run time
    box, frames = _device_frames(schedule, n, device, seed=2)

    def one_pass(frame_list, to_numpy):
        launches0 = _lib.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ordered = 0.0
        last = None
        for pos, quat in frame_list:
            nlist = order.compute_neighbours(box, pos, 3.5, 8)
            oo = order.orientational_order(box, pos, quat, 3.5, 8)
            nn = order.num_neighbours(box, pos, 3.5)
            last = (nlist, oo, nn)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        del ordered
        return sec, _lib.launch_count() - launches0, last

    one_pass(frames[:20], False)
    times = [one_pass(frames, False) for _ in range(3)]
    sec = statistics.median(t[0] for t in times)
    algo = 68.0 * n * len(frames)  # 28 read + 8 x 4 list + 4 count + 4 order per molecule and frame
    nlist, oo, nn = times[0][2]
    as_np = lambda v: v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)  # noqa: E731
    record = {
        "workload": f"configs[2]: {n} molecules, order.compute_neighbours(r=3.5, k=8) + orientational_order(r=3.5, k=8) + "
                    f"num_neighbours(r=3.5) through the public functions for each of {len(frames)} frames (device-resident frames)",
        "value": n * len(frames) / sec, "unit": "molecule*frames/s", "ms_per_frame": 1e3 * sec / len(frames),
        "gpu_launches_per_frame": times[0][1] / len(frames),
        "roofline": {"bound": "hbm", "achieved": algo / sec / 1e9, "peak": peak, "unit": "GB/s", "frac": algo / sec / 1e9 / peak,
                     "algorithmic_bytes_per_molecule_frame": 68,
                     "note": "latency bound at this size: a frame is 2.2 MB of algorithmic traffic; the search is arithmetic bound"},
        "check": {"mean_order": float(as_np(oo).mean()), "mean_neighbours": float(as_np(nn).astype(np.float64).mean()),
                  "missing_in_list": int((as_np(nlist).astype(np.int64) & 0xFFFFFFFF == 0xFFFFFFFF).sum())},
    }
    host_frames = [(p.cpu().numpy(), q.cpu().numpy()) for p, q in frames[:100]]
    for p, q in host_frames:  # read-only like the views of a memory-mapped trajectory: the three calls share one build
        p.flags.writeable = False
        q.flags.writeable = False
    one_pass(host_frames[:10], True)
    e2e_times = [one_pass(host_frames, True)[0] for _ in range(3)]
    e2e_sec = statistics.median(e2e_times)
    record["e2e"] = {
        "value": n * len(host_frames) / e2e_sec, "unit": "molecule*frames/s", "ms_per_frame": 1e3 * e2e_sec / len(host_frames),
        "h2d_bytes_per_frame": 28 * n, "d2h_bytes_per_frame": (8 * 4 + 8 + 4) * n,
        "source": "read-only numpy frames in (like gsdio's mmap views), numpy results out (three public calls per frame, one build)",
    }
    if cpu_baseline:
        from oracle import order as oorder

        t0 = time.perf_counter()
        count = 2
        for pos, quat in host_frames[:count]:
            oorder.compute_neighbours(box, pos, 3.5, 8)
            oorder.orientational_order(box, pos, quat, 3.5, 8)
            oorder.num_neighbours(box, pos, 3.5)
        cpu_sec = time.perf_counter() - t0
        record["cpu_baseline"] = {
            "value": n * count / cpu_sec, "unit": "molecule*frames/s", "cores": 1, "kind": "port",
            "sample": f"{count} frames of the same trajectory through the oracle's order functions (three neighbour builds each, like the reference)",
        }
    return record


def small_trajectory(seed=1):
    """configs[0]: ``(box, [(timestep, indexes, position, orientation), ...])`` on the host -- the bundled liquid frame
    followed by independent Gaussian steps (sigma 0.01 per component and timestep, 0.02 rad about z)."""
    from sdanalysis_b200.synthetic import exponential_schedule

    golden = np.load(Path(__file__).resolve().parent / "tests" / "golden" / "liquid.npz")
    box = np.asarray(golden["box"], dtype=np.float32)
    xy = np.asarray(golden["position"], dtype=np.float64)[:, :2].copy()
    quat0 = np.asarray(golden["orientation"], dtype=np.float64)
    theta = 2.0 * np.arctan2(quat0[:, 3], quat0[:, 0])
    n = len(xy)
    rng = np.random.default_rng(seed)
    span = np.asarray(box[:2], dtype=np.float64)
    frames, now = [], 0
    for timestep, indexes in exponential_schedule(10000, 100, 1000, 500):
        gap = timestep - now
        if gap > 0:
            xy += rng.normal(0.0, SIGMA_R * np.sqrt(gap), (n, 2))
            theta += rng.normal(0.0, SIGMA_T * np.sqrt(gap), n)
            xy -= span * np.floor(xy / span + 0.5)  # back into [-L/2, L/2)
            now = timestep
        pos = np.zeros((n, 3), dtype=np.float32)
        pos[:, :2] = xy
        quat = np.zeros((n, 4), dtype=np.float32)
        quat[:, 0], quat[:, 3] = np.cos(theta / 2), np.sin(theta / 2)
        frames.append((int(timestep), list(indexes), pos, quat))
    return box, frames


def _small_config(device, peak, cpu_baseline, tmpdir):
    import tempfile

    import torch

    from sdanalysis_b200 import _lib
    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.gsdio import GSDWriter
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.read import process_file

    box, frames = small_trajectory()
    n = len(frames[0][2])
    pairs = sum(len(ix) for _, ix, _, _ in frames)
    updates = pairs * n
    algo_bytes = sum(n * (65 * len(ix) + 28) for _, ix, _, _ in frames if ix)  # SURVEY 8(d), log-spaced schedule
    resident = [(ts, torch.from_numpy(pos).to(device), torch.from_numpy(quat).to(device), ix) for ts, ix, pos, quat in frames]

    def one_run():
        engine = DynamicsEngine(n, box, WAVE_NUMBER, mol_sites=Trimer().positions, device=device, ring_depth=4)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.launch_count()
        t0.record()
        rows = 0
        for first in range(0, len(resident), 256):
            for res in engine.run(resident[first : first + 256]):
                rows += len(res.keys)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1), rows, _lib.launch_count() - launches0

    one_run()
    runs = [one_run() for _ in range(5)]
    ms = statistics.median(r[0] for r in runs)
    assert runs[0][1] == pairs
    record = {
        "workload": f"configs[0]: bundled liquid dump ({n} molecules) + seeded Brownian trajectory, {len(frames)} frames, "
                    f"{pairs} (frame, origin) pairs on GenerateStepSeries(10000, 100, 1000, 500), wave number {WAVE_NUMBER}",
        "value": updates / (ms * 1e-3), "unit": "updates/s", "ms_per_frame": ms / len(frames), "ms_per_run": ms,
        "runs_ms": [r[0] for r in runs], "gpu_launches": runs[0][2],
        "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo_bytes / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_run": algo_bytes,
                     "note": "whole run; at 1250 molecules a frame is 35 KB of state per origin: launch latency, not bandwidth"},
    }
    del resident

    # ---- e2e: the reference's entry point on the trajectory file ----------------------------------------------
    with tempfile.TemporaryDirectory(dir=tmpdir) as scratch:
        path = Path(scratch) / "trajectory-Trimer-P13.50-T3.00.gsd"
        with GSDWriter(path) as writer:
            for ts, _, pos, quat in frames:
                writer.append(ts, box, pos, quat, dimensions=2)
        kwargs = dict(steps_max=10000, linear_steps=100, keyframe_interval=1000, keyframe_max=500)
        process_file(path, WAVE_NUMBER, **kwargs)  # warm-up
        times = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            df = process_file(path, WAVE_NUMBER, **kwargs)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    assert len(df) == pairs, (len(df), pairs)
    if os.environ.get("SDB_BENCH_CONFIG0_ROWS"):  # keep the rows (for a look at deviations off the GPU box)
        df.to_csv(os.environ["SDB_BENCH_CONFIG0_ROWS"], index=False, float_format="%.17g")
    sec = statistics.median(times)
    record["e2e"] = {
        "value": updates / sec, "unit": "updates/s", "ms_per_frame": 1e3 * sec / len(frames), "seconds_per_run": sec, "rows": len(df),
        "h2d_bytes_per_frame": 28 * n, "source": "read.process_file(GSD file) -> pandas.DataFrame, wall clock",
    }

    # ---- the oracle on the whole trajectory: CPU baseline and row-by-row parity ---------------------------------
    if cpu_baseline:
        record.update(small_oracle_leg(box, frames, df.to_dict("records")))
    return record


def small_oracle_leg(box, frames, got_rows):
    """Time the oracle over every (frame, origin) pair of configs[0] and compare ``got_rows`` (process_file's rows,
    in the reference's order: frame by frame, origins in index order) with its rows."""
    from oracle import parity as opar

    n = len(frames[0][2])
    run = opar.OracleRun(WAVE_NUMBER)
    t0 = time.perf_counter()
    for ts, ix, pos, quat in frames:
        run.frame(ts, box, pos, quat, ix)
    sec = time.perf_counter() - t0
    max_rel, max_abs_unit, failed = 0.0, 0.0, []
    if len(got_rows) != len(run.rows):
        failed.append(f"row count {len(got_rows)} != {len(run.rows)}")
    for i, (got, want) in enumerate(zip(got_rows, run.rows)):
        if int(got["keyframe"]) != int(want["keyframe"]):
            failed.append(f"row {i}: keyframe {got['keyframe']} != {want['keyframe']}")
            continue
        rel, bad = opar.compare_row(got, want, n)
        max_rel = max(max_rel, rel)
        max_abs_unit = max([max_abs_unit] + [abs(float(got[k]) - float(want[k])) for k in opar.UNIT_MEAN_KEYS])
        failed.extend(f"row {i} (keyframe {want['keyframe']}, time {want['time']}): {name}" for name in bad)
    return {
        "cpu_baseline": {
            "value": run.pairs * n / sec, "unit": "updates/s", "cores": 1, "kind": "port", "seconds_per_run": sec,
            "sample": f"the whole trajectory: {run.pairs} (frame, origin) pairs, 1 process",
        },
        "parity": {"ok": not failed, "rows_checked": min(len(got_rows), len(run.rows)), "max_rel": max_rel,
                   "max_abs_rot1_rot2": max_abs_unit, "failed": failed[:10],
                   "bar": "15 quantities of every row: rtol 1e-5 (rot1 / rot2, means of unit-magnitude f32 terms that cross "
                          f"zero: or within {opar.UNIT_MEAN_ATOL:.1e} absolute), counts exact, overlap exact unless tied"},
    }


def run_all(device, peak, cpu_baseline=True, tmpdir=None, only=None):
    tmpdir = tmpdir or ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    out = {}
    wanted = lambda name: not only or name in only  # noqa: E731
    if wanted("0_small"):
        try:
            out["0_small"] = _small_config(device, peak, cpu_baseline, tmpdir)
        except Exception as exc:  # a sub-record must not take the headline line down with it
            out["0_small"] = {"error": f"{type(exc).__name__}: {exc}"}
    for name, schedule in _schedules().items():
        if not wanted(name):
            continue
        try:
            out[name] = _dynamics_config(name, schedule, device, peak, cpu_baseline, tmpdir)
        except Exception as exc:  # a sub-record must not take the headline line down with it
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
    if wanted("2_order"):
        try:
            out["2_order"] = _order_config(device, peak, cpu_baseline)
        except Exception as exc:
            out["2_order"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


if __name__ == "__main__":
    import json
    import sys

    REPO = Path(__file__).resolve().parent
    for _p in (REPO, REPO / "statdyn-analysis_b200"):
        if str(_p) not in sys.path:
            sys.path.insert(0, str(_p))
    import torch

    import __graft_entry__

    __graft_entry__.build()
    peak = 6551.7
    try:
        peak = float(json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    names = [a for a in sys.argv[1:] if not a.startswith("--")]  # e.g. `python bench_configs.py 0_small 2_order`
    print(json.dumps(run_all(torch.device("cuda", 0), peak, cpu_baseline="--no-cpu-baseline" not in sys.argv, only=names or None)))

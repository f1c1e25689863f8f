"""The reference arm of ``bench.py`` (``--impl reference``: the CPU oracle on the host cores) runs without a
GPU and prints one JSON line with the keys the driver reads."""

import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run(
        [sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--molecules", "3000", "--steps", "2", "--warmup", "1"],
        capture_output=True, text=True, timeout=600, cwd=str(REPO),
    )
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "updates/s"
    assert line["metric"].startswith("molecule*origin updates/sec")
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("configs[3]") and line["vs_baseline"] is None


def test_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback: without CUDA the GPU arm exits with an error instead of measuring anything."""
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run(
        [sys.executable, str(REPO / "bench.py"), "--steps", "1", "--warmup", "3", "--molecules", "1000", "--origins", "2"],
        capture_output=True, text=True, timeout=600, cwd=str(REPO),
    )
    assert out.returncode != 0
    assert "CUDA" in (out.stderr + out.stdout)


def test_reference_arm_honours_steps_and_warmup():
    out = subprocess.run(
        [sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--molecules", "2000", "--steps", "3", "--warmup", "2"],
        capture_output=True, text=True, timeout=600, cwd=str(REPO),
    )
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["steps"] == 3 and line["warmup"] == 2 and "3 timed frames" in line["cpu_baseline"]["sample"]


def test_frame_pool_index_walks_back_and_forth_one_step_at_a_time():
    sys.path.insert(0, str(REPO))
    import bench

    seq = [bench.pingpong(i, 5) for i in range(20)]
    assert seq[:9] == [0, 1, 2, 3, 4, 3, 2, 1, 0]
    assert all(abs(a - b) == 1 for a, b in zip(seq, seq[1:]))  # consecutive frames are one walk step apart
    assert [bench.pingpong(i, 1300) for i in (0, 1, 1299)] == [0, 1, 1299] and bench.pingpong(7, 1) == 0


def test_bench_parity_report_accepts_the_oracle_and_rejects_a_wrong_row():
    """The parity leg of bench.py (``_oracle_parity``) with the oracle standing in for the device: identical rows and
    state pass; one perturbed quantity or one wrong passage time fails the run.  (On the GPU box the tables and the
    state come from the CUDA path; this keeps the bookkeeping -- which measured step a row belongs to, which
    origins were followed -- from rotting.)"""
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "statdyn-analysis_b200"))
    import numpy as np

    import bench
    from oracle import parity as opar
    from sdanalysis_b200.synthetic import SyntheticTrimer

    n, warmup, followed = 600, 3, [0, 1]
    walk = SyntheticTrimer(n, seed=3, sigma_r=0.15, sigma_theta=0.3)
    box = walk.box
    # creation: origin 0 at t = 0, origin 1 at t = 10, both at the last creation frame t = 50
    oracle_frames = []
    for ts, keys in ((0, [0]), (10, [1]), (50, [0, 1])):
        walk.advance(ts - walk.timestep)
        pos, quat = walk.frame()
        oracle_frames.append((ts, pos, quat, keys))
    t_first = 60
    pool_host = []
    for i in range(warmup + 1):
        walk.advance(t_first + i - walk.timestep)
        pool_host.append(walk.frame())
    # the "device": a second oracle over the same frames, rows as column tables per measured step, state after warm-up
    dev = opar.OracleRun(bench.WAVE_NUMBER)
    for ts, pos, quat, keys in oracle_frames:
        dev.frame(ts, box, pos, quat, keys)
    tables, state = [], {}
    for i, (pos, quat) in enumerate(pool_host):
        first_row = len(dev.rows)
        dev.frame(t_first + i, box, pos, quat, [0, 1, ])
        rows = dev.rows[first_row:]
        table = {q: np.array([r[q] for r in rows]) for q in rows[0] if q not in ("keyframe", "_tie")}
        table["keyframe"] = [r["keyframe"] for r in rows]
        tables.append((i, table))
        if i == warmup - 1:
            for key in followed:
                dyn, relax = dev.keyframes[key]
                state[key] = (dyn.delta_translation.copy(), dyn.delta_rotation.copy(), {k: v.copy() for k, v in relax.summary().items()})
    report, cpu = bench._oracle_parity(np, n, box, oracle_frames, pool_host, t_first, warmup, tables, state, followed)
    assert report["ok"] and report["checked_rows"] == 2 * (warmup + 1) and report["state_checked"] == [0, 1]
    assert cpu["kind"] == "port" and cpu["cores"] == 1 and cpu["value"] > 0
    # a wrong quantity in one row
    tables[1][1]["msd"] = tables[1][1]["msd"] * (1 + 1e-3)
    bad, _ = bench._oracle_parity(np, n, box, oracle_frames, pool_host, t_first, warmup, tables, state, followed)
    assert not bad["ok"] and bad["failed_rows"] and "msd" in bad["failed_rows"][0]["quantities"]
    tables[1][1]["msd"] = tables[1][1]["msd"] / (1 + 1e-3)
    # a wrong passage time of one molecule
    state[1][2]["tau_F"][17] += 1
    bad, _ = bench._oracle_parity(np, n, box, oracle_frames, pool_host, t_first, warmup, tables, state, followed)
    assert not bad["ok"] and bad["passage_mismatch"] == 1


def test_config0_trajectory_and_its_oracle_leg():
    """bench_configs' configs[0] record: the trajectory is the one SURVEY 8(d) describes (bundled liquid frame, 1901
    frames, 11 origins, 2361 pairs on GenerateStepSeries(10000, 100, 1000, 500)); the oracle leg accepts the oracle's
    own rows and reports a perturbed one."""
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "statdyn-analysis_b200"))
    import numpy as np

    import bench_configs
    from oracle import parity as opar

    box, frames = bench_configs.small_trajectory()
    assert len(frames) == 1901 and sum(len(ix) for _, ix, _, _ in frames) == 2361
    assert frames[0][0] == 0 and frames[-1][0] == 10000 and frames[-1][1] == list(range(11))
    golden = np.load(REPO / "tests" / "golden" / "liquid.npz")
    assert np.array_equal(frames[0][2], golden["position"]) and frames[0][2].shape == (1250, 3)
    assert np.allclose(np.abs(frames[0][3]), np.abs(golden["orientation"]), atol=1e-6)  # (q and -q are one rotation)
    half = np.asarray(box[:2]) / 2
    for _, _, pos, quat in frames[::97]:
        assert np.all(np.abs(pos[:, :2]) <= half + 1e-4) and not pos[:, 2].any() and not quat[:, 1:3].any()
        assert np.allclose(np.linalg.norm(quat, axis=1), 1, atol=1e-6)
    # mean squared displacement grows like 2 sigma^2 t (two components), crossing d = pi / (2 k) for a tail of molecules
    step = frames[200][2][:, :2].astype(np.float64) - frames[0][2][:, :2]
    step -= np.asarray(box[:2]) * np.rint(step / np.asarray(box[:2]))
    assert 0.5 < np.mean(np.sum(step ** 2, axis=1)) / (2 * 0.01 ** 2 * frames[200][0]) < 2.0

    sample = frames[:150]
    run = opar.OracleRun(bench_configs.WAVE_NUMBER)
    for ts, ix, pos, quat in sample:
        run.frame(ts, box, pos, quat, ix)
    rows = [dict(r) for r in run.rows]
    leg = bench_configs.small_oracle_leg(box, sample, rows)
    assert leg["parity"]["ok"] and leg["parity"]["rows_checked"] == len(rows) and leg["parity"]["max_rel"] == 0.0
    assert leg["cpu_baseline"]["kind"] == "port" and leg["cpu_baseline"]["value"] > 0
    rows[40]["msr"] = rows[40]["msr"] * (1 + 1e-3)
    bad = bench_configs.small_oracle_leg(box, sample, rows)
    assert not bad["parity"]["ok"] and "msr" in bad["parity"]["failed"][0]
    assert not bench_configs.small_oracle_leg(box, sample, rows[:-1])["parity"]["ok"]

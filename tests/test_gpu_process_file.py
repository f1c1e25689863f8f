"""The driver loop end to end: GSD file -> read.process_file -> rows / passage-time tables, against
the oracle driven like reference ``read/_read.py:128-168`` on the same frames (BASELINE config 1
substitute: the bundled trajectory blob is missing, so a seeded Brownian walk continues the bundled
liquid frame on a ``GenerateStepSeries`` schedule and is written with this repo's GSD writer)."""

import numpy as np
import pandas
import pytest

from _helpers import WAVE_NUMBER, compare_rows, oracle_process

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trajectory(tmp_path_factory, golden_dir):
    from sdanalysis_b200.gsdio import GSDWriter
    from sdanalysis_b200.StepSize import GenerateStepSeries

    g = np.load(golden_dir / "liquid.npz")
    box, pos0, quat0 = g["box"], g["position"], g["orientation"]
    n = len(pos0)
    rng = np.random.default_rng(1)
    xy = pos0[:, :2].astype(np.float64)
    theta = 2 * np.arctan2(quat0[:, 3], quat0[:, 0]).astype(np.float64)
    path = tmp_path_factory.mktemp("traj") / "trajectory-Trimer-P13.50-T3.00.gsd"
    last = 0
    with GSDWriter(path) as dst:
        for step in GenerateStepSeries(600, num_linear=10, gen_steps=100, max_gen=4):
            dt = step - last
            last = step
            if dt:
                xy += rng.normal(0, 0.02 * np.sqrt(dt), xy.shape)
                theta += rng.normal(0, 0.05 * np.sqrt(dt), n)
            L = box[:2].astype(np.float64)
            pos = np.zeros((n, 3), np.float32)
            pos[:, :2] = xy - L * np.floor(xy / L + 0.5)
            quat = np.zeros((n, 4), np.float32)
            quat[:, 0], quat[:, 3] = np.cos(theta / 2), np.sin(theta / 2)
            # three particles per molecule like the bundled files: molecule centres first
            body = np.concatenate([np.arange(n), np.arange(n), np.arange(n)]).astype(np.int32)
            dst.append(step, box, np.concatenate([pos, pos, pos]), np.concatenate([quat, quat, quat]), body=body)
    return path


def test_process_file_exponential(trajectory):
    from sdanalysis_b200 import read

    df = read.process_file(trajectory, wave_number=WAVE_NUMBER, linear_steps=10, keyframe_interval=100, keyframe_max=4)
    frames = list(read.read_gsd_trajectory(trajectory, linear_steps=10, keyframe_interval=100, keyframe_max=4))
    ref_rows, *_ = oracle_process(frames)
    assert len(df) == len(ref_rows)
    assert np.all(df["temperature"] == 3.00) and np.all(df["pressure"] == 13.50)
    assert df["temperature"].dtype == np.float64  # read_test.py:143-151
    rows = {(r["keyframe"], r["time"]): r for r in df.to_dict("records")}
    ordered = [rows[(r["keyframe"], r["time"])] for r in ref_rows]
    compare_rows(ordered, ref_rows, len(frames[0][1].position))
    # read_test.py:41-44: the first origin sees exactly the schedule of a single series
    from sdanalysis_b200.StepSize import GenerateStepSeries

    first = df[df["keyframe"] == 0]["time"].tolist()
    assert first == list(GenerateStepSeries(600, num_linear=10, gen_steps=10 ** 9))


@pytest.mark.parametrize("num_steps", [0, 10, 20, 100])
def test_stopiter_handling(trajectory, num_steps):
    """read_test.py:33-52: steps_max truncates both schedules."""
    from sdanalysis_b200 import read
    from sdanalysis_b200.StepSize import GenerateStepSeries

    df = read.process_file(trajectory, wave_number=WAVE_NUMBER, steps_max=num_steps, linear_steps=10)
    assert np.all(df["time"] == list(GenerateStepSeries(num_steps, num_linear=10)))
    df = read.process_file(trajectory, wave_number=WAVE_NUMBER, linear_steps=None, steps_max=num_steps)
    assert df["time"].max() == num_steps


def test_process_file_writes_tables(trajectory, tmp_path):
    """integration_test.py:20-26 / read_test.py:143-151: both tables, float64 T/P columns."""
    from sdanalysis_b200 import read

    out = tmp_path / "dynamics.parquet"
    assert read.process_file(trajectory, wave_number=WAVE_NUMBER, linear_steps=10, keyframe_interval=100, outfile=out) is None
    dyn = pandas.read_parquet(out)
    mol = pandas.read_parquet(out.with_suffix(".molecular_relaxations.parquet"))
    assert len(dyn) > 50 and dyn["temperature"].dtype == np.float64
    assert list(mol.columns[:2]) == ["keyframe", "molecule"]
    assert set(mol.columns) >= {"tau_D", "tau_F", "tau_L", "tau_T2", "tau_T3", "tau_T4", "temperature", "pressure"}
    assert len(mol) == 1250 * mol["keyframe"].nunique()
    # passage times of the table agree with the oracle
    frames = list(read.read_gsd_trajectory(trajectory, linear_steps=10, keyframe_interval=100))
    _, summaries, near, _ = oracle_process(frames)
    for key, want in summaries.items():
        got = mol[mol["keyframe"] == key]
        for name, col in want.items():
            mask = ~near[key][name]
            assert np.array_equal(got[name].to_numpy()[mask], col[mask]), (key, name)


def test_run_analysis_order(trajectory, tmp_path):
    """reference run_analysis.py:20-42: one CSV row per frame, fraction of particles with order > 0.9."""
    from oracle import order as oorder
    from sdanalysis_b200 import run_analysis
    from sdanalysis_b200.gsdio import open_gsd

    out = tmp_path / "order.csv"
    run_analysis.order(trajectory, out)
    lines = out.read_text().strip().splitlines()
    assert lines[0] == "Timestep, OrientOrder"
    with open_gsd(trajectory) as traj:
        assert len(lines) == len(traj) + 1
        snap = traj[len(traj) - 1]
        step, _, value = lines[-1].split()
        want = oorder.orientational_order(snap.box, snap._position, snap._orientation)
        assert int(step) == snap.step
        assert np.isclose(float(value), np.sum(want > 0.9) / len(want), atol=2.0 / len(want))

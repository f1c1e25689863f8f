"""Parity of the CUDA hot path with the CPU oracle on seeded synthetic Trimer trajectories.

All tests call through the C-ABI library (via the Python host layer).  Bars (BASELINE.json
north_star): relaxation times bit-exact except molecules within 1 ulp of a cutoff, displacement
accumulators bit-exact, floating-point quantities within rtol 1e-5.
"""

import numpy as np
import pytest

from _helpers import WAVE_NUMBER, compare_rows, near_cutoff, oracle_process

pytestmark = pytest.mark.gpu


def _run_engine(frames, n, **kw):
    from sdanalysis_b200.read import process_frames

    rows = []
    engine = process_frames(iter(frames), WAVE_NUMBER, sink=rows.append, **kw)
    return rows, engine


def _frames(schedule, n, seed, sigma_r=0.03, sigma_theta=0.08):
    from sdanalysis_b200.synthetic import synthetic_frames

    return list(synthetic_frames(schedule, n, seed=seed, sigma_r=sigma_r, sigma_theta=sigma_theta))


def _check_state(engine, keyframes, summaries, near, n):
    total_near = 0
    for key, (dyn, _) in keyframes.items():
        trans, rot = engine.delta_arrays(key)
        # translation: every operation is a single IEEE f32 operation in the oracle's order
        assert np.array_equal(trans, dyn.delta_translation), f"delta_translation of origin {key}"
        # rotation: f64 increment evaluated by a different (equally accurate) formula, then one
        # rounding to f32: differences are confined to f64 sums within ~1e-16 of a rounding boundary
        diff = rot[:, 0] != dyn.delta_rotation[:, 0]
        assert diff.mean() <= 1e-5, f"delta_rotation of origin {key}: {diff.sum()} differ"
        if diff.any():
            a, b = rot[diff, 0], dyn.delta_rotation[diff, 0]
            assert np.all(np.abs(a - b) <= np.spacing(np.abs(b)) * 1.01), "rotation differs by more than 1 ulp"
        got = engine.summary(key)
        for name, want in summaries[key].items():
            mask = ~near[key][name]
            total_near += int((~mask).sum())
            assert np.array_equal(got[name][mask], want[mask]), f"{name} of origin {key}"
    return total_near


@pytest.mark.parametrize("n", [1000, 4099])
def test_exponential_schedule_matches_oracle(n):
    from sdanalysis_b200.synthetic import exponential_schedule

    frames = _frames(exponential_schedule(200, 10, 20, 6), n, seed=11)
    ref_rows, summaries, near, keyframes = oracle_process(frames)
    rows, engine = _run_engine(frames, n)
    order = {(r["keyframe"], r["time"]): r for r in rows}
    rows = [order[(r["keyframe"], r["time"])] for r in ref_rows]
    compare_rows(rows, ref_rows, n)
    _check_state(engine, keyframes, summaries, near, n)


def test_linear_schedule_matches_oracle():
    from sdanalysis_b200.synthetic import linear_schedule

    n = 2048
    frames = _frames(linear_schedule(60, 5, 7), n, seed=5, sigma_r=0.08, sigma_theta=0.2)
    ref_rows, summaries, near, keyframes = oracle_process(frames)
    rows, engine = _run_engine(frames, n)
    order = {(r["keyframe"], r["time"]): r for r in rows}
    rows = [order[(r["keyframe"], r["time"])] for r in ref_rows]
    compare_rows(rows, ref_rows, n)
    _check_state(engine, keyframes, summaries, near, n)
    # the relaxation-heavy walk must actually exercise the trackers
    fired = sum(int((s["tau_F"] != 2 ** 32 - 1).sum()) for s in summaries.values())
    last = sum(int((s["tau_L"] != 2 ** 32 - 1).sum()) for s in summaries.values())
    assert fired > n and last > 0


def test_per_object_api_matches_oracle(golden_dir):
    """Dynamics / Relaxations objects (one origin each) on the bundled liquid frame + a walk."""
    from oracle import dynamics as odyn
    from sdanalysis_b200 import dynamics
    from sdanalysis_b200.molecules import Trimer

    g = np.load(golden_dir / "liquid.npz")
    box, pos0, quat0 = g["box"], g["position"], g["orientation"]
    n = len(pos0)
    rng = np.random.default_rng(3)
    theta = 2 * np.arctan2(quat0[:, 3], quat0[:, 0]).astype(np.float64)
    xy = pos0[:, :2].astype(np.float64)
    dyn = dynamics.Dynamics(0, box, pos0, quat0, molecule=Trimer(), wave_number=WAVE_NUMBER)
    relax = dynamics.Relaxations(0, box, pos0, quat0, molecule=Trimer(), wave_number=WAVE_NUMBER)
    ref_dyn = odyn.Dynamics(0, box, pos0, quat0, "trimer", WAVE_NUMBER)
    ref_relax = odyn.Relaxations(0, box, pos0, quat0, "trimer", wave_number=WAVE_NUMBER)
    near = None
    for step in range(1, 9):
        xy += rng.normal(0, 0.15, xy.shape)
        theta += rng.normal(0, 0.3, n)
        pos = np.zeros((n, 3), np.float32)
        L = box[:2].astype(np.float64)
        pos[:, :2] = xy - L * np.floor(xy / L + 0.5)
        quat = np.zeros((n, 4), np.float32)
        quat[:, 0], quat[:, 3] = np.cos(theta / 2), np.sin(theta / 2)
        got = dyn.compute_all(step * 10, pos, quat)
        want = ref_dyn.compute_all(step * 10, pos, quat)
        relax.add(step * 10, pos, quat)
        ref_relax.add(step * 10, pos, quat)
        near = near_cutoff(ref_relax, near)
        got["keyframe"] = want["keyframe"] = 0
        want["_tie"] = False
        compare_rows([got], [want], n)
        # the single-quantity accessors agree with compute_all
        assert dyn.compute_msd() == got["msd"] and dyn.compute_alpha() == got["alpha"]
    assert np.array_equal(dyn.delta_translation, ref_dyn.delta_translation)
    assert np.array_equal(dyn.compute_displacement(), ref_dyn.compute_displacement())
    summary, ref_summary = relax.summary(), ref_relax.summary()
    for name, want in ref_summary.items():
        # bit-exact except molecules that came within 1 ulp of a cutoff of that tracker (north_star)
        mask = ~near[name]
        assert np.array_equal(summary[name].to_numpy()[mask], want[mask]), name
        assert (~mask).sum() <= 2, name
    assert sorted(summary.columns) == sorted(ref_summary)


@pytest.mark.parametrize("sigma_r", [0.05, 8.0])
def test_scattering_function_matches_oracle(sigma_r):
    """Dynamics.compute_isf / compute_all(scattering_function=True) (dynamics.py:233-235,364-366): the
    Bessel-function evaluation of the angular mean (csrc/isf.cu) against the oracle's literal
    (360 x 2) . (2 x N) product; sigma_r = 8 pushes k |dr| past 0.4 * 360, where the kernel evaluates
    the 360 cosines literally."""
    from oracle import dynamics as odyn
    from sdanalysis_b200 import dynamics
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import SyntheticTrimer

    walk = SyntheticTrimer(3000, seed=21, sigma_r=sigma_r, sigma_theta=0.05)
    pos0, quat0 = walk.frame()
    dyn = dynamics.Dynamics(0, walk.box, pos0, quat0, molecule=Trimer(), wave_number=WAVE_NUMBER)
    ref = odyn.Dynamics(0, walk.box, pos0, quat0, "trimer", WAVE_NUMBER)
    for _ in range(3):
        walk.advance(40)
        pos, quat = walk.frame()
        got = dyn.compute_all(walk.timestep, pos, quat, scattering_function=True)
        want = ref.compute_all(walk.timestep, pos, quat, scattering_function=True)
        np.testing.assert_allclose(got["scattering_function"], want["scattering_function"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(dyn.compute_isf(), want["scattering_function"], rtol=1e-6, atol=1e-9)
    assert np.isnan(dyn.compute_all(walk.timestep, pos, quat)["scattering_function"])  # off by default


def test_scattering_function_column_of_process_frames():
    """process_file(..., scattering_function=True) fills the column for every (frame, origin) row."""
    from sdanalysis_b200.synthetic import exponential_schedule

    n = 2000
    frames = _frames(exponential_schedule(60, 5, 10, 4), n, seed=5)
    ref_rows, *_ = oracle_process(frames, scattering_function=True)
    rows, _ = _run_engine(frames, n, scattering_function=True)
    order = {(r["keyframe"], r["time"]): r for r in rows}
    for want in ref_rows:
        got = order[(want["keyframe"], want["time"])]
        np.testing.assert_allclose(got["scattering_function"], want["scattering_function"], rtol=1e-6, atol=1e-9)


def test_passage_times_when_time_moves_backwards():
    """Relaxations.add with a timestep smaller than an earlier one (relaxations.py:215-218: a fired tracker
    is rewritten while ``status > timediff``): the device's literal path (SDB_ORIGIN_LITERAL) against the
    oracle, every molecule, including the frames after time moves forward again."""
    from oracle import dynamics as odyn
    from sdanalysis_b200 import dynamics
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import SyntheticTrimer

    walk = SyntheticTrimer(6000, seed=41, sigma_r=0.12, sigma_theta=0.25)
    pos0, quat0 = walk.frame()
    relax = dynamics.Relaxations(0, walk.box, pos0, quat0, molecule=Trimer(), wave_number=WAVE_NUMBER)
    ref = odyn.Relaxations(0, walk.box, pos0, quat0, "trimer", wave_number=WAVE_NUMBER)
    near = None
    for timestep in (10, 20, 5, 6, 30, 31, 12, 40):
        walk.advance(4)
        pos, quat = walk.frame()
        relax.add(timestep, pos, quat)
        ref.add(timestep, pos, quat)
        near = near_cutoff(ref, near)
    got, want = relax.summary(), ref.summary()
    rewritten = 0
    for name, col in want.items():
        mask = ~near[name]
        assert np.array_equal(got[name].to_numpy()[mask], col[mask]), name
        rewritten += int(np.isin(col, (5, 6, 12)).sum())
    assert rewritten > 100  # the literal rule really rewrote passage times

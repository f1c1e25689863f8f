"""Shared helpers of the parity tests: drive the CPU oracle exactly like reference
``read/_read.py:128-168`` drives ``Dynamics`` / ``Relaxations`` and compare with the CUDA path."""

import numpy as np

from oracle import dynamics as odyn
from oracle.parity import near_cutoff, overlap_tie as _overlap_tie  # noqa: F401

WAVE_NUMBER = 2.9
FLOAT_KEYS = ("mean_displacement", "msd", "mfd", "mean_rotation", "msr", "rot1", "rot2")
RATIO_KEYS = ("alpha", "alpha_rot", "gamma")  # value = ratio - 1: compared on the ratio
RTOL = 1e-5  # BASELINE.json north_star: floating-point dynamics quantities within rtol 1e-5


def oracle_process(frames, wave_number=WAVE_NUMBER, ulps=1, scattering_function=False, only=None):
    """frames: iterable of (indexes, frame).  Returns (rows, summaries, near) where ``near[key]`` is
    a dict tracker-name -> bool mask of molecules that ever came within ``ulps`` ulp of a cutoff of
    that tracker (those are excluded from the bit-exact comparison, as north_star allows: "excluding
    pairs within 1 ulp of the cutoff").  ``only``: restrict the oracle to these origins (the others are
    independent of them: subsampling origins at the sizes where the oracle needs 0.1 - 20 s per pair); a dict
    ``origin -> number of frames`` also bounds how long an origin is followed (None: to the end)."""
    keyframes = {}
    rows = []
    near = {}
    done = {}
    for indexes, frame in frames:
        for index in indexes:
            if only is not None:
                if index not in only:
                    continue
                budget = only[index] if isinstance(only, dict) else None
                if budget is not None and done.get(index, 0) >= budget:
                    continue  # rows of the first `budget` frames of this origin only
                done[index] = done.get(index, 0) + 1
            if index not in keyframes:
                keyframes[index] = (
                    odyn.Dynamics(frame.timestep, frame.box, frame.position, frame.orientation, "trimer", wave_number),
                    odyn.Relaxations(frame.timestep, frame.box, frame.position, frame.orientation, "trimer", wave_number=wave_number),
                )
                near[index] = {name: np.zeros(len(frame.position), dtype=bool) for name in keyframes[index][1].mol_relax}
            dyn, relax = keyframes[index]
            row = dyn.compute_all(frame.timestep, frame.position, frame.orientation, scattering_function=scattering_function)
            relax.add(frame.timestep, frame.position, frame.orientation)
            row["keyframe"] = index
            row["_tie"] = _overlap_tie(dyn)
            rows.append(row)
            near[index] = near_cutoff(relax, near[index], ulps)
    summaries = {index: relax.summary() for index, (_, relax) in keyframes.items()}
    return rows, summaries, near, keyframes


def compare_rows(gpu_rows, ref_rows, n, n_sites=3, struct_slack=2):
    """Assert the parity bar on two row lists (same order)."""
    assert len(gpu_rows) == len(ref_rows)
    for got, want in zip(gpu_rows, ref_rows):
        where = f"keyframe {want['keyframe']} time {want['time']}"
        assert got["keyframe"] == want["keyframe"] and got["time"] == want["time"], where
        for key in FLOAT_KEYS:
            np.testing.assert_allclose(got[key], want[key], rtol=RTOL, atol=1e-12, err_msg=f"{key} {where}")
        for key in RATIO_KEYS:
            if np.isnan(want[key]):
                assert np.isnan(got[key]), f"{key} {where}"
            else:
                np.testing.assert_allclose(got[key] + 1, want[key] + 1, rtol=RTOL, err_msg=f"{key} {where}")
        # counts of f32 threshold tests on bit-identical displacements: exact
        assert got["com_struct"] == want["com_struct"], f"com_struct {where}"
        # struct compares f64 site distances built from f32 sin/cos of two different libraries
        assert abs(got["struct"] - want["struct"]) <= struct_slack / (n * n_sites) + 1e-15, f"struct {where}"
        if not want["_tie"]:
            assert got["overlap"] == want["overlap"], f"overlap {where}"
        else:
            assert 0.0 <= got["overlap"] <= 1.0
        assert np.isnan(got["scattering_function"]) == np.isnan(want["scattering_function"])

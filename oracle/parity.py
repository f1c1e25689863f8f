"""Parity report between rows / per-molecule state of the CUDA path and the CPU oracle.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``): used by ``bench.py``'s cpu_baseline leg and by the tests.

The bar is the one BASELINE.json's north_star states: floating-point dynamics quantities within rtol 1e-5,
relaxation timesteps bit-exact excluding molecules within 1 ulp of a cutoff; on top of that the accumulated
displacement of every molecule must be bit-identical (every operation on that path is a single IEEE f32
operation in the reference's order) and the accumulated rotation may differ by 1 ulp on at most 1e-5 of the
molecules (its f64 increment is evaluated by a different, equally accurate formula before the one rounding).
"""

import numpy as np

from . import dynamics as odyn

RTOL = 1e-5
FLOAT_KEYS = ("mean_displacement", "msd", "mfd", "mean_rotation", "msr", "rot1", "rot2")
# rot1 = <cos t> and rot2 = <2 cos^2 t - 1> are means of terms of magnitude <= 1 that the reference evaluates in f32
# (6e-8 of rounding per term) and that cross zero at long times (rot2 ~ 1e-3 +- 0.03 for 1250 molecules after
# 5000 steps): there the reference's own f32 value is 1e-5 (relative) away from the exact mean, and no relative bar
# is meaningful.  Two values also agree when they differ by less than 2 f32 roundings of a unit-magnitude term.
UNIT_MEAN_KEYS = ("rot1", "rot2")
UNIT_MEAN_ATOL = 2.0 ** -22
RATIO_KEYS = ("alpha", "alpha_rot", "gamma")  # value = ratio - 1: compared on the ratio


def near_cutoff(relax, near=None, ulps=1):
    """Accumulate, per tracker of the oracle ``relax``, the mask of molecules whose current distance lies
    within ``ulps`` ulp (f32) of a cutoff of that tracker."""
    if near is None:
        near = {name: np.zeros(relax._num_elements, dtype=bool) for name in relax.mol_relax}
    disp = np.linalg.norm(relax.motion.delta_translation, axis=1)
    rot = np.linalg.norm(relax.motion.delta_rotation, axis=1)
    for name, func in relax.mol_relax.items():
        value = rot if func.relaxation_type == "rotation" else disp
        cuts = [func.threshold] + ([func._irreversibility] if hasattr(func, "_irreversibility") else [])
        for cut in cuts:
            c = np.float32(cut)
            near[name] |= np.abs(value - c) <= ulps * np.spacing(c)
    return near


def overlap_tie(dyn):
    """True when the top-10% cut of either key array falls inside a run of equal values (numpy's choice
    among equal keys is implementation-defined: SURVEY trap T4)."""
    k = int(dyn.num_particles * 0.1)
    if k == 0:
        return True
    for values in (dyn.compute_displacement(), np.abs(dyn.compute_rotation())):
        s = np.sort(values)
        if s[-k] == s[-k - 1]:
            return True
    return False


class OracleRun:
    """Drives oracle ``Dynamics`` / ``Relaxations`` pairs exactly like reference ``read/_read.py:134-153``."""

    def __init__(self, wave_number):
        self.wave_number = wave_number
        self.keyframes = {}
        self.near = {}
        self.rows = []
        self.pairs = 0

    def frame(self, timestep, box, position, orientation, indexes):
        for index in indexes:
            if index not in self.keyframes:
                self.keyframes[index] = (
                    odyn.Dynamics(timestep, box, position, orientation, "trimer", self.wave_number),
                    odyn.Relaxations(timestep, box, position, orientation, "trimer", wave_number=self.wave_number),
                )
            dyn, relax = self.keyframes[index]
            row = dyn.compute_all(timestep, position, orientation)
            relax.add(timestep, position, orientation)
            row["keyframe"] = index
            row["_tie"] = overlap_tie(dyn)
            self.rows.append(row)
            self.near[index] = near_cutoff(relax, self.near.get(index))
            self.pairs += 1


def compare_row(got, want, n, n_sites=3, struct_slack=2):
    """Return (max relative deviation over the floating-point quantities, list of failed quantities)."""
    failed = []
    max_rel = 0.0
    if int(got["time"]) != int(want["time"]):
        failed.append("time")
    for key in FLOAT_KEYS + RATIO_KEYS:
        g, w = float(got[key]), float(want[key])
        if key in RATIO_KEYS:
            if np.isnan(w):
                if not np.isnan(g):
                    failed.append(key)
                continue
            g, w = g + 1, w + 1
        denom = max(abs(w), 1e-300)
        floor = UNIT_MEAN_ATOL if key in UNIT_MEAN_KEYS else 1e-12
        rel = abs(g - w) / denom if abs(g - w) > floor else 0.0
        max_rel = max(max_rel, rel)
        if not rel <= RTOL:
            failed.append(key)
    if float(got["com_struct"]) != float(want["com_struct"]):
        failed.append("com_struct")
    if abs(float(got["struct"]) - float(want["struct"])) > struct_slack / (n * n_sites) + 1e-15:
        failed.append("struct")
    if not want.get("_tie", False) and float(got["overlap"]) != float(want["overlap"]):
        failed.append("overlap")
    return max_rel, failed


def compare_state(trans, rot, summary, dyn, ref_summary, near):
    """Per-molecule state of one origin: returns a dict of mismatch counts (all zero = parity)."""
    out = {
        "delta_translation_mismatch": int(np.sum(np.any(trans != dyn.delta_translation, axis=1))),
        "delta_rotation_differs": 0, "delta_rotation_beyond_1ulp": 0, "passage_mismatch": 0, "passage_excluded_1ulp": 0,
    }
    diff = rot[:, 0] != dyn.delta_rotation[:, 0]
    out["delta_rotation_differs"] = int(diff.sum())
    if diff.any():
        a, b = rot[diff, 0], dyn.delta_rotation[diff, 0]
        out["delta_rotation_beyond_1ulp"] = int(np.sum(np.abs(a - b) > np.spacing(np.abs(b)) * 1.01))
    for name, want in ref_summary.items():
        mask = ~near[name]
        out["passage_excluded_1ulp"] += int((~mask).sum())
        out["passage_mismatch"] += int(np.sum(np.asarray(summary[name])[mask] != want[mask]))
    return out

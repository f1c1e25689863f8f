"""Generate the committed golden fixtures from the reference checkout.  Run in the build container:

    python tests/golden/make_golden.py

Needs /root/reference (read-only).  Nothing in tests/ or bench.py reads /root/reference at run
time; they read the files this script writes next to itself.

* ``stepsize_golden.json``  -- outputs of the reference's own ``StepSize.py`` (stdlib-only, so it is
  imported and *executed*, not restated): ``generate_steps`` sequences and, for several
  ``GenerateStepSeries`` parameter sets, the full list of ``(step, get_index())`` pairs.
* ``crystal_*.npz`` / ``liquid.npz`` -- box, molecule positions and orientations of the four
  single-frame GSD dumps bundled with the reference tests (``test/data/dump-*.gsd``), decoded with
  this repo's own GSD reader (the ``gsd`` package is not installed).  These are data, not source.
"""

import importlib.util
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REPO / "statdyn-analysis_b200"))


def load_reference_stepsize():
    spec = importlib.util.spec_from_file_location("ref_StepSize", REF / "src/sdanalysis/StepSize.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def stepsize_golden():
    ref = load_reference_stepsize()
    out = {"generate_steps": [], "series": []}
    for total, lin, start in [(100, 10, 0), (1000, 100, 0), (99, 10, 0), (12345, 7, 3), (10, 100, 0), (0, 100, 0), (500, 10, 250)]:
        out["generate_steps"].append(
            {"args": [total, lin, start], "steps": list(ref.generate_steps(total, lin, start))}
        )
    for total, lin, gen, max_gen in [
        (1000, 10, 10, 100),       # BASELINE config 2 S_exp
        (2000, 10, 10, 200),       # BASELINE config 4 S_exp
        (10000, 100, 1000, 500),   # BASELINE config 1 / read_gsd_test.py:24-43
        (100, 10, 200000, 500),
        (1000, 100, 100, 3),
        (777, 9, 50, 5),
        (0, 100, 1000, 500),
        (20, 100, 1_000_000, 500),
    ]:
        series = ref.GenerateStepSeries(total, num_linear=lin, gen_steps=gen, max_gen=max_gen)
        pairs = [[int(step), list(series.get_index())] for step in series]
        out["series"].append({"args": [total, lin, gen, max_gen], "pairs": pairs})
    (HERE / "stepsize_golden.json").write_text(json.dumps(out, separators=(",", ":")))
    print("stepsize_golden.json:", sum(len(s["pairs"]) for s in out["series"]), "steps")


def gsd_fixtures():
    from sdanalysis_b200.gsdio import open_gsd

    for src, dst in [
        ("dump-Trimer-13.50-0.40-p2.gsd", "crystal_p2.npz"),
        ("dump-Trimer-13.50-0.40-p2gg.gsd", "crystal_p2gg.npz"),
        ("dump-Trimer-13.50-0.40-pg.gsd", "crystal_pg.npz"),
        ("dump-Trimer-13.50-3.00.gsd", "liquid.npz"),
    ]:
        with open_gsd(REF / "test/data" / src) as trj:
            snap = trj[0]
            np.savez_compressed(
                HERE / dst,
                box=np.array(snap.box),
                position=np.array(snap.position),
                orientation=np.array(snap.orientation),
                timestep=np.int64(snap.timestep),
                dimensions=np.int64(snap.dimensions),
                num_particles=np.int64(snap.N),
            )
            print(dst, "molecules:", len(snap), "particles:", snap.N, "box:", np.array(snap.box), "step:", snap.timestep)


if __name__ == "__main__":
    stepsize_golden()
    gsd_fixtures()

"""CPU-only tests of the host logic: step schedules, GSD reader/writer, trajectory iterators."""

import json

import numpy as np
import pytest

from sdanalysis_b200 import StepSize
from sdanalysis_b200.frame import HoomdFrame
from sdanalysis_b200.gsdio import GSDWriter, open_gsd
from sdanalysis_b200.read import WriteCache, open_trajectory, read_gsd_trajectory
from sdanalysis_b200.read import _gsd
from sdanalysis_b200.synthetic import SyntheticTrimer, exponential_schedule, linear_schedule
from sdanalysis_b200.util import get_filename_vars


def test_stepsize_matches_reference_golden(golden_dir):
    """Sequences produced by the reference's own StepSize.py (tests/golden/make_golden.py)."""
    gold = json.loads((golden_dir / "stepsize_golden.json").read_text())
    for case in gold["generate_steps"]:
        assert list(StepSize.generate_steps(*case["args"])) == case["steps"]
    for case in gold["series"]:
        total, lin, gen, max_gen = case["args"]
        series = StepSize.GenerateStepSeries(total, num_linear=lin, gen_steps=gen, max_gen=max_gen)
        assert [[int(s), list(series.get_index())] for s in series] == case["pairs"], case["args"]


@pytest.mark.parametrize("total,lin", [(100, 10), (1000, 100), (12345, 7)])
def test_stepsize_properties(total, lin):
    """stepsize_test.py:105-222: strictly increasing, no duplicates, ends on total_steps."""
    steps = list(StepSize.GenerateStepSeries(total, num_linear=lin, gen_steps=50, max_gen=4))
    assert steps == sorted(set(steps)) and steps[0] == 0 and steps[-1] == total


@pytest.fixture()
def gsd_trajectory(tmp_path):
    """read_gsd_test.py:24-43: 10 particles on the GenerateStepSeries(10000, 100, 1000) schedule."""
    path = tmp_path / "test-Trimer-P13.50-T3.00.gsd"
    with GSDWriter(path) as dst:
        for timestep in StepSize.GenerateStepSeries(total_steps=10000, num_linear=100, gen_steps=1000):
            dst.append(timestep, [1.0, 1.0, 0.0, 0.0, 0.0, 0.0], np.zeros((10, 3)), dimensions=2)
    return path


def test_gsd_roundtrip(tmp_path):
    walk = SyntheticTrimer(257, seed=3)
    path = tmp_path / "walk.gsd"
    frames = []
    with GSDWriter(path) as dst:
        for step in (0, 5, 50):
            walk.advance(step - walk.timestep)
            pos, quat = walk.frame()
            frames.append((step, pos, quat))
            dst.append(step, walk.box, pos, quat)
    with open_gsd(path) as trj:
        assert len(trj) == 3
        for snap, (step, pos, quat) in zip(trj, frames):
            assert snap.step == step and snap.dimensions == 2 and len(snap) == 257
            assert np.array_equal(snap.position, pos) and np.array_equal(snap.orientation, quat)
            assert np.array_equal(snap.box, walk.box)
            assert not snap.position.flags.writeable  # zero-copy, read-only views
        assert trj[-1].step == 50
        with pytest.raises(IndexError):
            trj[3]


def test_gsd_bodies_define_molecules(tmp_path):
    """frame.py:136-156: num_mols = max(body) + 1, clamped to the particle count."""
    path = tmp_path / "b.gsd"
    pos = np.arange(27, dtype=np.float32).reshape(9, 3)
    with GSDWriter(path) as dst:
        dst.append(7, [5, 5, 1, 0, 0, 0], pos, body=[0, 1, 2, 0, 0, 1, 1, 2, 2])
        dst.append(8, [5, 5, 1, 0, 0, 0], pos, body=[-1] * 9)
    with open_gsd(path) as trj:
        assert len(trj[0]) == 3 and trj[0].position.shape == (3, 3) and trj[0].orientation.shape == (3, 4)
        assert len(trj[1]) == 9
        assert np.all(trj[0].orientation[:, 0] == 1)  # default orientation (1,0,0,0)


def test_golden_dump_matches_reference_test_expectations(golden_dir):
    """frame_test.py:46-72 contract on data decoded from the bundled dump files."""
    g = np.load(golden_dir / "liquid.npz")
    assert g["position"].dtype == np.float32 and g["position"].shape == (1250, 3)
    assert g["orientation"].dtype == np.float32 and g["orientation"].shape == (1250, 4)
    assert int(g["num_particles"]) == 3750 and int(g["timestep"]) == 20_000_000
    assert not np.any(g["orientation"][:, 1:3])  # planar quaternions
    assert np.allclose(np.linalg.norm(g["orientation"], axis=1), 1, atol=1e-6)


def test_read_trajectory_linear(gsd_trajectory):
    """read_gsd_test.py:51-57."""
    for index, (_, snap) in enumerate(read_gsd_trajectory(gsd_trajectory, steps_max=100, linear_steps=None)):
        assert isinstance(snap, HoomdFrame) and snap.timestep == index


def test_read_trajectory_exp(gsd_trajectory):
    """read_gsd_test.py:60-67."""
    traj = read_gsd_trajectory(gsd_trajectory, steps_max=200, linear_steps=100, keyframe_interval=1000)
    steps = StepSize.GenerateStepSeries(total_steps=200, num_linear=100, gen_steps=1000)
    count = 0
    for index, (_, snap) in zip(steps, traj):
        assert snap.timestep == index
        count += 1
    assert count > 100


def test_linear_sequence_keyframes(gsd_trajectory):
    """read_test.py:49-66: one shared growing index list; keyframe_max + 1 origins (trap T5)."""
    for index, _ in _gsd._gsd_linear_trajectory(gsd_trajectory):
        assert index == [0]
    seen = 0
    for index, frame in read_gsd_trajectory(gsd_trajectory, linear_steps=None, keyframe_interval=10, steps_max=100):
        assert index == list(range(frame.timestep // 10 + 1))
        seen += 1
    assert seen == 101
    longest = max(len(i) for i, _ in read_gsd_trajectory(gsd_trajectory, linear_steps=None, keyframe_interval=1, keyframe_max=4, steps_max=50))
    assert longest == 5


def test_exponential_iterator_skips_missing_steps(tmp_path):
    """_gsd.py:154-182: trajectory frames absent from the schedule and vice versa are skipped."""
    path = tmp_path / "gaps.gsd"
    steps = [0, 1, 2, 3, 5, 7, 8, 9, 10, 15, 20, 30]
    with GSDWriter(path) as dst:
        for s in steps:
            dst.append(s, [1, 1, 0, 0, 0, 0], np.zeros((4, 3)))
    got = [f.timestep for _, f in read_gsd_trajectory(path, steps_max=30, linear_steps=10, keyframe_interval=1000)]
    assert got == [0, 1, 2, 3, 5, 7, 8, 9, 10, 20, 30]


def test_open_trajectory(gsd_trajectory):
    frames = list(open_trajectory(gsd_trajectory, frame_interval=50))
    assert [f.timestep for f in frames][:3] == [0, 50, 100]
    with pytest.raises(ValueError):
        next(open_trajectory("x.xyz"))


def test_write_cache(tmp_path):
    """read_test.py:69-110."""
    out = tmp_path / "cache.parquet"
    cache = WriteCache(out)
    for i in range(9000):
        cache.append({"value": i})
    assert len(cache._cache) == 9000 - 8192 and len(cache) == 9000
    cache.flush()
    assert len(cache._cache) == 0 and out.is_file()
    cache.close()  # parquet: the row groups written by the flushes become a readable file
    import pandas

    assert len(pandas.read_parquet(out)) == 9000
    # appending never re-reads the file while the run is open: one row group per flush
    import pyarrow.parquet as pq

    assert pq.ParquetFile(out).num_row_groups == 2


def test_write_cache_column_blocks_and_append_after_close(tmp_path):
    """Column blocks (one per frame, what process_file feeds) end up as rows in order; a table closed by an
    earlier run is extended by a later one."""
    import pandas

    from sdanalysis_b200.read._read import _write_table, close_tables

    out = tmp_path / "blocks.parquet"
    cache = WriteCache(out)
    for frame in range(30):
        count = 1 + frame % 4
        cache.append_table({"time": np.full(count, frame), "msd": np.arange(count, dtype=float), "keyframe": np.arange(count)})
    assert len(cache) == sum(1 + f % 4 for f in range(30))
    df = cache.to_dataframe()
    assert df["time"].tolist() == [f for f in range(30) for _ in range(1 + f % 4)]
    cache.flush()
    cache.close()
    first = pandas.read_parquet(out)
    assert first.equals(df)
    _write_table(df.iloc[:5], out, "dynamics", append=True)
    close_tables(out)
    assert len(pandas.read_parquet(out)) == len(df) + 5


def test_filename_variables():
    """util_test.py:86-148 (subset relevant to process_file)."""
    v = get_filename_vars("trajectory-Trimer-P13.50-T3.00.gsd")
    # "Trimer" starts with "T": the reference parses it as a temperature that the real one overrides
    assert (v.temperature, v.pressure, v.crystal) == ("3.00", "13.50", None)
    v = get_filename_vars("dump-Trimer-P1.00-T0.40-p2gg-ID3.gsd")
    assert (v.temperature, v.pressure, v.crystal, v.iteration_id) == ("0.40", "1.00", "p2gg", "3")
    assert get_filename_vars("test.gsd") == (None, None, None, None)


def test_schedules_match_survey_counts():
    """(frame, origin) pair counts of the BASELINE configs."""
    pairs = list(exponential_schedule(1000, 10, 10, 100))
    assert len(pairs) == 1001 and sum(len(i) for _, i in pairs) == 2405
    pairs = list(linear_schedule(1000, 10, 99))
    assert len(pairs[-1][1]) == 100 and sum(len(i) for _, i in pairs) == 50500


def test_quaternion_helpers_match_the_rowan_restatement():
    """util.rotate_vectors / quaternion_rotation / quaternion2z / z2quaternion (reference util.py:138-167)
    against oracle/rowan_shim.py, and the reference's own golden vectors (test/util_test.py:235-250)."""
    from oracle import rowan_shim as rowan
    from sdanalysis_b200 import util

    rng = np.random.default_rng(3)
    q = rng.normal(size=(64, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    p = rng.normal(size=(64, 4))
    p /= np.linalg.norm(p, axis=1)[:, None]
    v = rng.normal(size=(64, 3))
    assert np.allclose(util.rotate_vectors(q, v), rowan.rotate(q, v), atol=1e-12)
    assert np.allclose(util.quaternion_rotation(q, p), rowan.intrinsic_distance(q, p), atol=1e-12)
    assert np.allclose(util.quaternion2z(q), rowan.to_euler(q)[:, 0], atol=1e-6)
    theta = rng.uniform(-np.pi, np.pi, 64)
    assert np.allclose(util.quaternion2z(util.z2quaternion(theta)), theta, atol=2e-7)
    golden = [  # test/util_test.py:235-250: the four orientations of the p2gg crystal, one per quadrant
        ([0.982_461_99, 0.0, 0.0, 0.186_462_98], 0.375_121_38),
        ([0.209_398_27, 0.0, 0.0, 0.977_830_41], 2.719_673_4),
        ([0.976_600_05, 0.0, 0.0, -0.215_063_53], -0.433_513_61),
        ([-0.211_795_45, 0.0, 0.0, 0.977_314], -2.714_769_36),
    ]
    for quat, angle in golden:
        assert np.allclose(util.quaternion2z(np.array([quat], dtype=np.float32)), angle)


def test_molecule2particles_reproduces_the_reference_pairing():
    """dynamics/_util.py:42-54: rotated sites are concatenated site-major, translations repeated
    molecule-major (SURVEY trap T2) -- the helper keeps that behaviour, like csrc/struct.cu."""
    from sdanalysis_b200.dynamics._util import molecule2particles
    from sdanalysis_b200.molecules import Trimer

    mol = Trimer()
    n = 4
    position = np.arange(3 * n, dtype=np.float32).reshape(n, 3)
    orientation = np.zeros((n, 4), dtype=np.float32)
    out = molecule2particles(position, orientation, mol.positions)
    assert out.shape == (3 * n, 3)
    sites = mol.positions.astype(np.float32)
    for row in range(3 * n):
        assert np.allclose(out[row], sites[row // n] + position[row // 3], atol=1e-6)
    with pytest.raises(ValueError):
        molecule2particles(position, (1, 0, 0, 0), mol.positions)


def _write_lammpstrj(path, steps, n=7, seed=0):
    rng = np.random.default_rng(seed)
    frames = []
    with open(path, "w") as dst:
        for step in steps:
            pos = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
            order = rng.permutation(n)  # rows in arbitrary order: the id column decides the slot
            dst.write(f"ITEM: TIMESTEP\n{step}\nITEM: NUMBER OF ATOMS\n{n}\n")
            dst.write("ITEM: BOX BOUNDS pp pp pp\n-5.0 5.0\n-6.0 6.0\n-0.5 0.5\n")
            dst.write("ITEM: ATOMS id type x y z\n")
            for j in order:
                dst.write(f"{j + 1} 1 {pos[j, 0]:.6f} {pos[j, 1]:.6f} {pos[j, 2]:.6f}\n")
            frames.append(pos)
    return frames


def test_lammps_trajectory(tmp_path):
    """reference read/_lammps.py:28-95 and frame.py:75-125: rows placed by id, box from the bounds, identity
    orientations, one time-origin, steps_max cut-off; open_trajectory dispatches on the suffix."""
    from sdanalysis_b200.read import open_trajectory, parse_lammpstrj, read_lammps_trajectory

    path = tmp_path / "short.lammpstrj"
    written = _write_lammpstrj(path, [0, 10, 20, 30])
    frames = list(parse_lammpstrj(path))
    assert [f.timestep for f in frames] == [0, 10, 20, 30]
    for frame, pos in zip(frames, written):
        assert frame.position.dtype == np.float32 and frame.position.shape == (7, 3)
        np.testing.assert_allclose(frame.position, pos, atol=1e-6)
        np.testing.assert_array_equal(frame.box, np.array([10, 12, 1, 0, 0, 0], dtype=np.float32))
        assert frame.orientation.shape == (7, 4) and np.all(frame.orientation[:, 0] == 1) and len(frame) == 7
        assert frame.dimensions == 2 and frame.image is None
    pairs = list(read_lammps_trajectory(path, steps_max=20))
    assert [f.timestep for _, f in pairs] == [0, 10, 20] and all(ix == [0] for ix, _ in pairs)
    assert [f.timestep for f in open_trajectory(path)] == [0, 10, 20, 30]
    bad = tmp_path / "bad.lammpstrj"
    bad.write_text("ITEM: TIMESTEP\n0\nITEM: NUMBER OF MOLECULES\n3\n")
    with pytest.raises(RuntimeError):
        list(parse_lammpstrj(bad))


def test_order_frame_cache_recognises_only_frames_nobody_can_write():
    """order.compute_neighbours / orientational_order / num_neighbours share one build per frame (the reference
    builds three times); the frame is recognised by identity, which is only safe for arrays that cannot change:
    torch tensors (version counter) and numpy arrays that are read-only all the way down (mmap views)."""
    import torch

    from sdanalysis_b200 import order

    a = np.zeros((10, 3), dtype=np.float32)
    assert order._identity(a, torch) is None  # writable: never cached
    view = a.view()
    view.flags.writeable = False
    assert not order._immutable(view) and order._identity(view, torch) is None  # its base can still be written
    frozen = np.frombuffer(bytes(120), dtype=np.float32).reshape(10, 3)  # read-only buffer, like an mmap view
    assert order._immutable(frozen) and order._identity(frozen, torch) == order._identity(frozen[:], torch)
    assert order._identity(frozen, torch) != order._identity(frozen[1:], torch)
    t = torch.zeros((10, 3))
    before = order._identity(t, torch)
    assert before == order._identity(t, torch)
    t.add_(1.0)  # an in-place write bumps the version: a different frame
    assert order._identity(t, torch) != before


def test_gsd_views_are_read_only_and_pinning_degrades_gracefully(tmp_path):
    """Frames are read-only views of the file mapping; without a CUDA device pin() reports failure and nothing changes."""
    import torch

    path = tmp_path / "t.gsd"
    pos = np.arange(30, dtype=np.float32).reshape(10, 3)
    with GSDWriter(path) as w:
        for step in range(3):
            w.append(step, [5, 5, 1, 0, 0, 0], pos + step)
    with open_gsd(path) as trj:
        snap = trj[1]
        assert not snap.position.flags.writeable and np.array_equal(snap.position, pos + 1)
        if not torch.cuda.is_available():
            assert trj.pin() is False
        assert np.array_equal(trj[2].position, pos + 2)


def test_process_file_host_loop_without_a_gpu(gsd_trajectory, tmp_path, monkeypatch):
    """The whole host side of ``process_file`` (_read.py:82-195) -- schedule, key lists, native frame loop,
    batch finalisation, table writers -- with a dry engine (all bookkeeping, no CUDA): one row per (frame, active
    keyframe) with ``time = frame.timestep - first timestep of the keyframe``, in frame order."""
    import functools

    import pandas

    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.read import _read

    monkeypatch.setattr(_read, "DynamicsEngine", functools.partial(DynamicsEngine, dry=True))
    for kwargs in (
        dict(steps_max=100, linear_steps=None, keyframe_interval=10, keyframe_max=5),
        dict(steps_max=300, linear_steps=10, keyframe_interval=50, keyframe_max=4),
    ):
        first, want = {}, []
        for indexes, frame in read_gsd_trajectory(gsd_trajectory, **kwargs):
            for key in indexes:
                want.append((key, frame.timestep - first.setdefault(key, frame.timestep)))
        assert len(first) > 3 and len(want) > 50
        df = _read.process_file(gsd_trajectory, wave_number=2.9, mol_relaxations=[], **kwargs)
        assert list(zip(df["keyframe"].tolist(), df["time"].tolist())) == want
        assert {"msd", "alpha", "struct", "overlap", "temperature", "pressure"} <= set(df.columns)
        out = tmp_path / f"out_{kwargs['steps_max']}.h5"
        assert _read.process_file(gsd_trajectory, wave_number=2.9, mol_relaxations=[], outfile=out, **kwargs) is None
        written = sorted(tmp_path.glob(f"out_{kwargs['steps_max']}*"))
        assert written, "no table written"
        for table in written:
            if table.suffix == ".parquet":  # (no PyTables in this image: the dynamics table is a parquet file)
                back = pandas.read_parquet(table)
                assert list(zip(back["keyframe"].tolist(), back["time"].tolist())) == want


def test_closed_forms_from_the_sums_reproduce_compute_all():
    """Host half of ``Dynamics.compute_all`` without a GPU: the SDB_S_* sums (include/sdb200.h) taken in f64 from the
    oracle's own accumulators, pushed through ``finalize_row`` / ``finalize_table``, give the oracle's 15 quantities
    (dynamics.py:183-303), including NaN for alpha / alpha_rot / gamma at zero motion (SURVEY trap T3)."""
    from oracle import parity as opar
    from sdanalysis_b200 import _lib
    from sdanalysis_b200.engine import ALL_QUANTITIES, finalize_row, finalize_table
    from sdanalysis_b200.synthetic import SyntheticTrimer

    n, wave_number = 900, 2.9
    distance = np.pi / (2 * wave_number)
    walk = SyntheticTrimer(n, seed=11, sigma_r=0.08, sigma_theta=0.2)
    run = opar.OracleRun(wave_number)
    sums, timediffs = [], []
    for timestep in (0, 1, 4, 20, 90):
        walk.advance(timestep - walk.timestep)
        pos, quat = walk.frame()
        run.frame(timestep, walk.box, pos, quat, [0])
        dyn, _ = run.keyframes[0]
        want = run.rows[-1]
        dr = np.linalg.norm(dyn.delta_translation, axis=1)  # f32, like compute_displacement
        th = np.linalg.norm(dyn.delta_rotation, axis=1)
        if timestep == 0:
            # rowan's to_euler(q / q) leaves +-2e-16 in some accumulators; their f32 square underflows in the reference's
            # own msr^2, which is how it arrives at NaN.  The device state of a new origin is exactly zero.
            assert th.max() < 1e-15
            th = np.zeros_like(th)
        d2, t2 = dr.astype(np.float64) ** 2, th.astype(np.float64) ** 2
        s = np.zeros(_lib.SDB_NSUM)
        s[_lib.S_DISP], s[_lib.S_DISP2], s[_lib.S_DISP4] = dr.astype(np.float64).sum(), d2.sum(), (d2 * d2).sum()
        s[_lib.S_COMSTRUCT] = np.count_nonzero(dr < distance)
        s[_lib.S_ROT], s[_lib.S_ROT2], s[_lib.S_ROT4] = th.astype(np.float64).sum(), t2.sum(), (t2 * t2).sum()
        c = np.cos(th.astype(np.float64))
        s[_lib.S_COS], s[_lib.S_COS2], s[_lib.S_D2R2] = c.sum(), (2 * c * c - 1).sum(), (d2 * t2).sum()
        s[_lib.S_STRUCT] = round(float(want["struct"]) * n * 3)
        s[_lib.S_OVERLAP] = round(float(want["overlap"]) * int(0.1 * n))
        row = finalize_row(s, timestep, n, distance=distance, overlap_k=int(0.1 * n), n_sites=3)
        assert list(row) == list(ALL_QUANTITIES) and len(row) == 15
        max_rel, failed = opar.compare_row(row, want, n)
        assert not failed and max_rel < 2e-6, (timestep, failed, max_rel)
        if timestep == 0:
            assert all(np.isnan(row[k]) for k in ("alpha", "alpha_rot", "gamma")) and want["_tie"]
        sums.append(s)
        timediffs.append(timestep)
    table = finalize_table(np.array(sums), timediffs, n, distance=distance, overlap_k=int(0.1 * n), n_sites=3)
    for i, s in enumerate(sums):
        row = finalize_row(s, timediffs[i], n, distance=distance, overlap_k=int(0.1 * n), n_sites=3)
        for key in ALL_QUANTITIES:
            assert (np.isnan(row[key]) and np.isnan(table[key][i])) or row[key] == table[key][i], (key, i)
    # without a wave number / overlap / sites the dependent quantities stay NaN; N < 10 divides by zero like the reference
    bare = finalize_row(sums[1], 1, n)
    assert all(np.isnan(bare[k]) for k in ("com_struct", "struct", "overlap", "scattering_function"))
    with pytest.raises(ZeroDivisionError):
        finalize_row(sums[1], 1, 9, overlap_k=0)
    assert np.isnan(finalize_row(sums[1], 1, 9, overlap_k=0, strict=False)["overlap"])


def test_gsd_reader_reports_corruption_as_runtime_error(tmp_path):
    """Like the gsd package: a file that cannot be indexed raises RuntimeError when opened, a frame whose data lie
    outside the file raises RuntimeError when it is read -- which is what the reference's loops skip
    (_read.py:218-224, _gsd.py:100-109)."""
    from sdanalysis_b200 import gsdio
    from sdanalysis_b200.gsdio import open_gsd

    path = tmp_path / "ok.gsd"
    rng = np.random.default_rng(0)
    quat = np.tile(np.array([1, 0, 0, 0], np.float32), (50, 1))
    with GSDWriter(path) as dst:
        for step in range(6):
            dst.append(step, [10, 10, 1, 0, 0, 0], rng.random((50, 3)).astype(np.float32), quat)
    data = path.read_bytes()
    for name, blob in (("empty", b""), ("short", data[:100]), ("zeros", bytes(400)), ("truncated", data[: len(data) - 700])):
        bad = tmp_path / f"{name}.gsd"
        bad.write_bytes(blob)
        with pytest.raises(RuntimeError):
            open_gsd(bad)
    # one chunk of frame 3 pointing past the end of the file: that frame alone is unreadable
    with open_gsd(path) as trj:
        assert len(trj) == 6
    header = gsdio._HEADER.unpack_from(data, 0)
    index_location, allocated = header[1], header[2]
    index = np.frombuffer(data, dtype=gsdio._ENTRY, count=allocated, offset=index_location).copy()
    hit = np.flatnonzero((index["frame"] == 3) & (index["N"] == 50))[0]
    index["location"][hit] = len(data) + 4096
    patched = bytearray(data)
    patched[index_location : index_location + index.nbytes] = index.tobytes()
    broken = tmp_path / "broken.gsd"
    broken.write_bytes(bytes(patched))
    with open_gsd(broken) as trj:
        assert len(trj) == 6 and trj[2].step == 2
        with pytest.raises(RuntimeError):
            trj[3]
    assert [f.timestep for f in open_trajectory(broken)] == [0, 1, 2, 4, 5]
    assert [f.timestep for _, f in read_gsd_trajectory(broken, steps_max=5, linear_steps=None)] == [0, 1, 2, 4, 5]

"""CPU-only checks of the C-ABI library: it loads, exports every symbol ``include/sdb200.h``
declares, and its host-compiled arithmetic helpers agree with the oracle.  No GPU compute."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import rowan_shim as rowan
from oracle.freud_shim import Box

REPO = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()
    from sdanalysis_b200 import _lib

    return _lib.load()


def test_header_symbols_are_exported(lib):
    from sdanalysis_b200 import _lib

    header = (REPO / "include" / "sdb200.h").read_text()
    declared = set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.sdb_version()


def test_struct_layouts_match_header():
    from sdanalysis_b200 import _lib

    assert ctypes.sizeof(_lib.SdbOrigin) == 32 == np.dtype(_lib.ORIGIN_DTYPE).itemsize
    assert ctypes.sizeof(_lib.SdbSegment) == 24 == np.dtype(_lib.SEGMENT_DTYPE).itemsize
    assert ctypes.sizeof(_lib.SdbTracker) == 24
    assert ctypes.sizeof(_lib.SdbDynParams) == 32 + 6 * 24


def test_missing_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device the product path raises instead of computing."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sdanalysis_b200 import dynamics, order

    with pytest.raises(RuntimeError, match="CUDA"):
        dynamics.Dynamics(0, [10, 10, 10], np.zeros((4, 3), np.float32))
    with pytest.raises(RuntimeError, match="CUDA"):
        order.compute_neighbours([10, 10, 1, 0, 0, 0], np.zeros((4, 3), np.float32))


def test_wrap_helper_is_bit_exact(lib):
    rng = np.random.default_rng(0)
    n = 4000
    v = np.zeros((n, 3), np.float32)
    v[:, 0] = rng.uniform(-300, 300, n)
    v[:, 1] = rng.uniform(-300, 300, n)
    v[:200, :2] = rng.normal(0, 1e-5, (200, 2))
    v[200:210] = 0
    ox, oy = ctypes.c_float(), ctypes.c_float()
    for lx, ly, xy in [(62.75029, 108.68669, 0.0), (10.0, 7.0, 0.3), (2236.068, 2236.068, 0.0)]:
        want = Box(lx, ly, xy=xy, is2D=True).wrap(v)
        for i in range(n):
            lib.sdb_test_wrap2d(lx, ly, xy, v[i, 0], v[i, 1], ctypes.byref(ox), ctypes.byref(oy))
            assert np.float32(ox.value) == want[i, 0] and np.float32(oy.value) == want[i, 1], (lx, i)


def test_rotation_increment_helper_matches_oracle(lib):
    rng = np.random.default_rng(1)
    n = 6000
    th0 = rng.uniform(-np.pi, np.pi, n)
    th1 = th0 + np.concatenate([rng.normal(0, 0.05, n // 2), rng.uniform(-np.pi, np.pi, n - n // 2)])
    q0 = np.zeros((n, 4), np.float32)
    q1 = np.zeros((n, 4), np.float32)
    q0[:, 0], q0[:, 3] = np.cos(th0 / 2), np.sin(th0 / 2)
    q1[:, 0], q1[:, 3] = np.cos(th1 / 2), np.sin(th1 / 2)
    want = rowan.to_euler(rowan.divide(q1, q0))[:, 0]
    bad = ctypes.c_int32(0)
    got = np.array(
        [lib.sdb_test_rot_increment(q1[i, 0], q1[i, 3], q0[i, 0], q0[i, 3], ctypes.byref(bad)) for i in range(n)]
    )
    assert np.max(np.abs(got - want)) < 4e-15
    # what matters: the f32 accumulator after adding the increment
    acc = rng.normal(0, 1.0, n).astype(np.float32)
    a = (acc.astype(np.float64) + want).astype(np.float32)
    b = (acc.astype(np.float64) + got).astype(np.float32)
    assert np.sum(a != b) <= 1
    # rowan rejects non-unit quotients with ValueError: the helper flags them
    lib.sdb_test_rot_increment(0.6, 0.9, 0.6, 0.8, ctypes.byref(bad))
    assert bad.value == 1
    assert lib.sdb_test_rot_increment(0.6, 0.8, 0.6, 0.8, ctypes.byref(bad)) == 0.0 and bad.value == 0


def test_wrap_and_rotation_increment_edge_cases(lib):
    """The same two helpers where the rounding is delicate: coordinates within a few ulps of 0, +-L/2, +-L, +-3L/2,
    2L and far outside the box, signed zeros and denormals; quaternion quotients next to the +-pi branch cut, after a
    full turn, below f32 resolution, and with one operand negated (q and -q are one rotation)."""
    ox, oy = ctypes.c_float(), ctypes.c_float()
    for lx, ly, xy in [(62.75029, 108.68669, 0.0), (10.0, 7.0, 0.3), (2236.068, 2236.068, 0.0), (4472.136, 4472.136, 0.0)]:
        specials = [-0.0, 1e-30, -1e-30, 1e-45, -1e-45]
        for base in (0.0, lx / 2, -lx / 2, lx, -lx, 1.5 * lx, -1.5 * lx, 2 * lx, 1e6, -1e6, 3e7):
            for k in range(-3, 4):
                x = np.float32(base)
                for _ in range(abs(k)):
                    x = np.nextafter(x, np.float32(np.inf if k > 0 else -np.inf))
                specials.append(float(x))
        m = len(specials)
        v = np.zeros((2 * m, 3), np.float32)
        v[:m, 0], v[:m, 1] = specials, np.float32(ly) / 3
        v[m:, 1], v[m:, 0] = specials, np.float32(lx) / 3
        want = Box(lx, ly, xy=xy, is2D=True).wrap(v)
        for i in range(2 * m):
            lib.sdb_test_wrap2d(lx, ly, xy, v[i, 0], v[i, 1], ctypes.byref(ox), ctypes.byref(oy))
            assert np.float32(ox.value) == want[i, 0] and np.float32(oy.value) == want[i, 1], (lx, v[i])
    rng = np.random.default_rng(3)
    n = 2000
    th0 = rng.uniform(-np.pi, np.pi, n)
    step = np.concatenate([np.pi + rng.normal(0, 1e-4, n // 4), -np.pi + rng.normal(0, 1e-4, n // 4),
                           rng.normal(0, 1e-7, n // 4), 2 * np.pi + rng.normal(0, 1e-3, n - 3 * (n // 4))])
    q0 = np.zeros((n, 4), np.float32)
    q1 = np.zeros((n, 4), np.float32)
    q0[:, 0], q0[:, 3] = np.cos(th0 / 2), np.sin(th0 / 2)
    q1[:, 0], q1[:, 3] = np.cos((th0 + step) / 2), np.sin((th0 + step) / 2)
    q1[::2] *= -1
    want = rowan.to_euler(rowan.divide(q1, q0))[:, 0]
    bad = ctypes.c_int32(0)
    got = np.array([lib.sdb_test_rot_increment(q1[i, 0], q1[i, 3], q0[i, 0], q0[i, 3], ctypes.byref(bad)) for i in range(n)])
    # the same side of the branch cut as rowan's atan2 -- except for a turn by EXACTLY pi (the quotient's scalar part is an
    # exact zero, e.g. when the f32 components of the two quaternions are swapped bit for bit), where rowan's sign of
    # pi hangs on the sign of a zero product and the helper returns -pi: a documented measure-zero deviation
    exact_half_turn = rowan.divide(q1, q0)[:, 0] == 0
    assert 0 < exact_half_turn.sum() < 10
    assert np.max(np.abs(got - want)[~exact_half_turn]) < 4e-15
    assert np.allclose(np.abs(got[exact_half_turn]), np.pi, rtol=0, atol=4e-15)


def test_threshold_cuts_are_exact():
    """disp2 >= cut  <=>  fl32(sqrt(disp2)) > thr, checked on both sides of every cut."""
    from sdanalysis_b200.engine import _cut_direct, _cut_sqrt

    for thr in [0.0, 0.4, np.pi / (2 * 2.9), 3 * np.pi / (2 * 2.9), 1.0, 1e-3, 123.456]:
        t32 = np.float32(thr)
        gt, ge = _cut_sqrt(thr, True), _cut_sqrt(thr, False)
        assert np.sqrt(gt) > t32 and not np.sqrt(np.nextafter(gt, np.float32(0))) > t32 or gt == 0
        assert np.sqrt(ge) >= t32
        if ge > 0:
            assert not np.sqrt(np.nextafter(ge, np.float32(0))) >= t32
        assert _cut_direct(thr, True) == np.nextafter(t32, np.float32(np.inf)) and _cut_direct(thr, False) == t32


def test_tracker_validation_errors():
    """Error behaviour of Relaxations.set_mol_relax / create_mol_relaxations (relaxations.py:108-116,316-329)."""
    from sdanalysis_b200.engine import build_trackers

    with pytest.raises(ValueError):
        build_trackers([{"threshold": 1.0}])
    with pytest.raises(ValueError):
        build_trackers([{"name": "a"}])
    with pytest.raises(ValueError):
        build_trackers([{"name": "a", "threshold": -1.0}])
    with pytest.raises(ValueError):
        build_trackers([{"name": "a", "threshold": 1.0, "last_passage": True, "last_passage_cutoff": -2.0}])
    specs = build_trackers(
        [{"name": "a", "threshold": 1.0}, {"name": "b", "threshold": 0.5, "last_passage": True}, {"name": "c", "threshold": 1, "type": "rotation"}]
    )
    assert [s.bit for s in specs] == [0, 1, 3] and specs[1].cutoff == 1.5


# ---- the boundary seen from plain C ---------------------------------------------------------------------
@pytest.fixture(scope="module")
def c_client(lib, tmp_path_factory):
    """tests/cabi_client.c compiled as strict C99 against include/sdb200.h and linked with the shared library."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    from sdanalysis_b200 import _lib

    libdir = _lib.LIB_PATH.parent
    exe = tmp_path_factory.mktemp("cabi") / "client"
    subprocess.run(
        ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{REPO / 'include'}",
         str(REPO / "tests" / "cabi_client.c"), "-o", str(exe), f"-L{libdir}", "-lsdb200", f"-Wl,-rpath,{libdir}"],
        check=True, capture_output=True, text=True)

    def run(mode):
        return subprocess.run([str(exe), mode], check=True, capture_output=True, text=True, timeout=60).stdout.splitlines()

    return run


def test_header_is_plain_c_and_layouts_match_the_ctypes_mirrors(c_client):
    from sdanalysis_b200 import _lib

    sizes, fields, consts = {}, {}, {}
    for line in c_client("layout"):
        kind, *rest = line.split()
        if kind == "size":
            sizes[rest[0]] = int(rest[1])
        elif kind == "field":
            fields.setdefault(rest[0], {})[rest[1]] = int(rest[2])
        else:
            consts[rest[0]] = int(rest[1])
    assert set(sizes) == {"SdbTracker", "SdbDynParams", "SdbOrigin", "SdbSegment", "SdbEngineConfig", "SdbFrame",
                          "SdbOriginInfo", "SdbEngineStats"}
    for name, size in sizes.items():
        mirror = getattr(_lib, name)
        assert ctypes.sizeof(mirror) == size, name
        assert {f: getattr(mirror, f).offset for f, *_ in mirror._fields_} == fields[name], name
    assert np.dtype(_lib.ORIGIN_DTYPE).itemsize == sizes["SdbOrigin"]
    assert {f: np.dtype(_lib.ORIGIN_DTYPE).fields[f][1] for f in np.dtype(_lib.ORIGIN_DTYPE).names} == fields["SdbOrigin"]
    assert {f: np.dtype(_lib.SEGMENT_DTYPE).fields[f][1] for f in np.dtype(_lib.SEGMENT_DTYPE).names} == fields["SdbSegment"]
    assert consts["SDB_NSUM"] == _lib.SDB_NSUM and consts["SDB_MAX_TRACKERS"] == _lib.SDB_MAX_TRACKERS
    assert (consts["SDB_MEM_HOST"], consts["SDB_MEM_PINNED"]) == (_lib.SDB_MEM_HOST, _lib.SDB_MEM_PINNED)
    assert consts["SDB_ENGINE_DRY"] == _lib.SDB_ENGINE_DRY and consts["SDB_FRAME_NO_SUMS"] == _lib.SDB_FRAME_NO_SUMS
    assert (consts["SDB_OK"], consts["SDB_EINVAL"], consts["SDB_ECUDA"]) == (0, -1, -2)


def test_c_client_drives_the_frame_loop_through_the_header_alone(c_client):
    """The loop of read/_read.py:128-168 for a linear schedule (new origin every 3 frames, 3 alive), written
    against the header in C: every frame reports ``timestep - origin timestep`` per active origin, oldest first."""
    out = c_client("dry")
    assert out[0].startswith("version sdb200") and "sm_100a" in out[0]
    frames = [line.split() for line in out if line.startswith("frame")]
    assert len(frames) == 12
    first = 0
    for f, words in enumerate(frames):
        newest = f // 3
        first = max(first, newest - 2)
        assert int(words[1]) == f and int(words[3]) == newest - first + 1
        assert [int(w) for w in words[5:]] == [f - 3 * k for k in range(first, newest + 1)]
    assert out[13] == "oldest key 1 timestep 3 updates 9 max_timediff 8 stride 1000"
    assert out[14] == "unknown key rc -1"
    assert out[15] == "stats frames 12 updates 27000 origins 3"


def test_integration_md_ctypes_stub_runs_as_written(lib):
    """The reference-side binding shown in INTEGRATION.md (level 2) is executed verbatim -- only the library path and
    the dry flag are supplied -- so its struct layouts, signatures and loop cannot drift from the library."""
    from types import SimpleNamespace

    from sdanalysis_b200 import _lib

    text = (REPO / "INTEGRATION.md").read_text()
    start = text.index("# sdanalysis/_b200.py")
    code = text[start : text.index("```", start)]
    assert 'C.CDLL("libsdb200.so")' in code
    scope = {}
    exec(compile(code.replace('C.CDLL("libsdb200.so")', f'C.CDLL("{_lib.LIB_PATH}")'), "INTEGRATION.md", "exec"), scope)
    for name, mirror in (("Tracker", _lib.SdbTracker), ("EngineConfig", _lib.SdbEngineConfig), ("Frame", _lib.SdbFrame)):
        assert ctypes.sizeof(scope[name]) == ctypes.sizeof(mirror)
        assert [(f, getattr(scope[name], f).offset) for f, _ in scope[name]._fields_] == [
            (f, getattr(mirror, f).offset) for f, *_ in mirror._fields_]
    n = 64
    cfg = scope["EngineConfig"](n=n, wave_number=2.9, com_cut_lt=0.29, compute_sums=1, overlap_k=n // 10, ring_depth=4,
                                flags=_lib.SDB_ENGINE_DRY)
    cfg.box[0] = cfg.box[1] = 20.0
    cfg.box[2] = 1.0
    rng = np.random.default_rng(0)
    quat = np.zeros((n, 4), np.float32)
    quat[:, 0] = 1
    frames = [([k for k in range(t // 4 + 1)], SimpleNamespace(timestep=t, position=rng.normal(size=(n, 3)), orientation=quat))
              for t in range(10)]
    got = [(key, int(td), sums.shape) for key, td, sums in scope["process_frames"](iter(frames), cfg)]
    assert got == [(k, t - 4 * k, (16,)) for t in range(10) for k in range(t // 4 + 1)]


def test_every_export_is_documented_in_integration_md():
    """INTEGRATION.md's table names the reference interface behind every entry point of include/sdb200.h
    (lists such as ``sdb_engine_create / destroy / step`` share their prefix)."""
    header = (REPO / "include" / "sdb200.h").read_text()
    doc = (REPO / "INTEGRATION.md").read_text()
    named = set(re.findall(r"\bsdb_[a-z0-9_]+", doc))
    for match in re.finditer(r"`(sdb_[a-z0-9]+_)([a-z0-9_]+)((?: / [a-z0-9_]+)+)`", doc):
        named.update(match.group(1) + tail for tail in match.group(3).split(" / ")[1:])
    declared = set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", header))
    assert not declared - named, sorted(declared - named)

"""Host logic of the native engine (csrc/engine.cu) without a GPU: a *dry* engine does all of its
bookkeeping -- origin records, previous-sample pool, segments, descriptors, tickets -- and skips every CUDA
call.  The expectations restate the reference's loop (read/_read.py:128-168, dynamics/_util.py:155-156):
which origins exist, what their time differences are, which previous sample each one steps from."""

import ctypes
import math

import numpy as np
import pytest

from sdanalysis_b200 import _lib
from sdanalysis_b200.engine import DynamicsEngine
from sdanalysis_b200.synthetic import exponential_schedule, linear_schedule

ORIGIN_NEW, ORIGIN_LITERAL, ORIGIN_HISTORY = 1, 2, 4
SEG_NO_ROTATION, SEG_ZERO_STEP = 1, 2


def _engine(n=1000, **kw):
    kw.setdefault("mol_sites", np.array([[0, -1 / 3, 0], [-0.866, 1 / 6, 0], [0.866, 1 / 6, 0]], dtype=np.float32))
    return DynamicsEngine(n, [70.0, 70.0, 1, 0, 0, 0], 2.9, dry=True, **kw)


def _frame(n, seed=0, quat=True):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-30, 30, (n, 3)).astype(np.float32)
    q = np.zeros((n, 4), dtype=np.float32)
    q[:, 0] = 1
    return pos, (q if quat else None)


def _expected_segment_sizes(count, n, chunk=None):
    n_parts = (n + 255) // 256
    blocks = (n_parts + 3) // 4
    floor = 8 if blocks >= 592 else (4 if blocks >= 148 else 2)  # small systems: more, shorter segments
    size = chunk or max(floor, math.ceil(count * blocks / (148 * 16)))
    size = math.ceil(count / math.ceil(count / size))
    return [min(size, count - first) for first in range(0, count, size)]


def test_linear_schedule_bookkeeping():
    n = 5000
    eng = _engine(n)
    pos, quat = _frame(n)
    known = set()
    prev_ptr = None
    for t, active in linear_schedule(45, 3, 13):
        res = eng.step(t, pos, quat, active)
        origins, segments = eng.last_descriptors()
        assert len(origins) == len(active)
        fresh = [k for k in active if k not in known]
        # existing origins first (one group: they all saw the previous frame), origins created now last
        assert res.timediffs == [t - 3 * k for k in active]
        assert sorted(origins["timediff"].tolist(), reverse=True) == origins["timediff"].tolist()
        new_flags = (origins["flags"] & ORIGIN_NEW) != 0
        assert new_flags.tolist() == [False] * (len(active) - len(fresh)) + [True] * len(fresh)
        assert np.all(origins["flags"] & ORIGIN_HISTORY) and not np.any(origins["flags"] & ORIGIN_LITERAL)
        assert len(set(origins["d_delta"].tolist())) == len(active) and np.all(origins["d_flags"] != 0)
        old = len(active) - len(fresh)
        want = _expected_segment_sizes(old, n) if old else []
        got = [(s["first"], s["count"]) for s in segments]
        firsts = np.cumsum([0] + want).tolist()
        assert got[: len(want)] == list(zip(firsts, want))
        if fresh:
            assert got[-1] == (old, len(fresh)) and segments[-1]["d_prev"] == 0
        if old:
            # all existing origins step from the frame of the previous call
            assert len(set(segments["d_prev"][: len(want)].tolist())) == 1
            assert prev_ptr is None or True
        st = eng.stats()
        assert st.pool_frames - st.pool_free == 1  # one previous sample is alive: the frame just processed
        assert st.pool_frames <= 2
        known.update(active)
    assert eng.frames_seen == 45 and eng.updates == n * sum(len(a) for _, a in linear_schedule(45, 3, 13))
    assert list(eng.origins.keys()) == list(range(14))
    assert eng.origins[3].timestep == 9 and eng.origins[3].updates == 45 - 9  # creation counts as an update


def test_exponential_schedule_previous_samples():
    """Every origin steps from the frame IT saw last: origins whose last frames differ must
    be in different segments, the pool holds exactly one frame per distinct last-seen frame."""
    n = 777
    eng = _engine(n)
    pos, quat = _frame(n)
    last_seen = {}
    ptr_of_frame = {}
    for t, active in exponential_schedule(300, 10, 20, 8):
        res = eng.step(t, pos, quat, active)
        origins, segments = eng.last_descriptors()
        assert res.keys == active
        assert sorted(res.timediffs) == sorted(origins["timediff"].tolist())
        # map every origin record back to its key through the time difference and the NEW flag
        by_prev = {}
        for seg in segments:
            members = origins[seg["first"] : seg["first"] + seg["count"]]
            by_prev.setdefault(int(seg["d_prev"]), []).extend(members["timediff"].tolist())
        t0 = {k: t - td for k, td in zip(active, res.timediffs)}
        want_groups = {}
        for k in active:
            want_groups.setdefault(last_seen.get(k), []).append(t - t0[k])
        assert sorted(sorted(v) for v in by_prev.values()) == sorted(sorted(v) for v in want_groups.values())
        for k in active:
            last_seen[k] = t
        st = eng.stats()
        assert st.pool_frames - st.pool_free == len(set(last_seen.values()))
    assert eng.stats().pool_frames <= len(last_seen) + 1


def test_literal_flag_when_time_moves_backwards_and_zero_step():
    n = 300
    eng = _engine(n)
    pos, quat = _frame(n)
    for t in (0, 10, 5, 6, 12):
        eng.step(t, pos, quat, [0])
        origins, segments = eng.last_descriptors()
        literal = bool(origins["flags"][0] & ORIGIN_LITERAL)
        assert literal == (t in (5, 6)), t  # relaxations.py:215-218: status > timediff can hold again
    assert eng.origins[0].max_timediff == 12
    res = eng.step(12, None, None, [0], zero_step=True)
    origins, segments = eng.last_descriptors()
    assert segments["flags"][0] & SEG_ZERO_STEP and res.timediffs == [0 + 12]
    assert eng.origins[0].updates == 5 and eng.frames_seen == 5  # a zero step adds no sample
    with pytest.raises(RuntimeError):
        eng.step(13, None, None, [7], zero_step=True)
    with pytest.raises(OverflowError):
        eng.step(2 ** 32 + 5, pos, quat, [0])
    with pytest.raises(OverflowError):
        eng.step(-1, pos, quat, [0])


def test_orientation_flags_and_shape_errors():
    n = 64
    eng = _engine(n)
    pos, quat = _frame(n)
    eng.step(0, pos, None, [0])
    assert eng.last_descriptors()[1]["flags"][0] & SEG_NO_ROTATION
    eng.step(1, pos, quat, [0, 1])  # previous sample of origin 0 has no orientation
    origins, segments = eng.last_descriptors()
    assert segments["flags"].tolist() == [SEG_NO_ROTATION, 0]
    eng.step(2, pos, quat, [0, 1])
    assert eng.last_descriptors()[1]["flags"].tolist() == [0]
    with pytest.raises(RuntimeError):
        eng.step(3, pos[:-1], quat, [0])
    with pytest.raises(RuntimeError):
        eng.step(3, pos, quat[:, :3], [0])
    with pytest.raises(RuntimeError):
        DynamicsEngine(0, [1, 1, 1], 2.9, dry=True)


def test_predicted_overlap_flag_needs_four_earlier_selections():
    n = 40_000
    eng = _engine(n)
    pos, quat = _frame(n)
    flags = []
    for t in range(8):
        eng.step(t, pos, quat, [0, 1])
        flags.append(eng.stats().last_step_flags)
    assert flags == [0, 0, 0, 0, 1, 1, 1, 1]
    eng.step(8, pos, quat, [0, 1], want_sums=False)  # Relaxations.add only: no selection, history broken
    eng.step(9, pos, quat, [0, 1])
    assert eng.stats().last_step_flags == 0
    eng.step(10, pos, quat, [0, 1, 2])  # a new origin has no history
    assert eng.stats().last_step_flags == 0


def test_increment_sets_for_shared_previous_samples(monkeypatch):
    n = 70_001
    monkeypatch.setenv("SDB_INCREMENTS_MAX_PER_SEGMENT", "64")
    eng = _engine(n, chunk=8)
    pos, quat = _frame(n)
    eng.step(0, pos, quat, list(range(20)))
    assert eng.stats().last_inc_sets == 0  # all new: previous sample == current frame
    eng.step(1, pos, quat, list(range(24)))
    origins, segments = eng.last_descriptors()
    assert [s["count"] for s in segments] == [7, 7, 6, 4]
    assert segments["reserved"].tolist() == [1, 1, 1, 0] and eng.stats().last_inc_sets == 1
    monkeypatch.setenv("SDB_INCREMENTS", "0")
    off = _engine(n, chunk=8)
    off.step(0, pos, quat, list(range(20)))
    off.step(1, pos, quat, list(range(20)))
    assert off.stats().last_inc_sets == 0 and not off.last_descriptors()[1]["reserved"].any()


def test_rows_come_back_in_caller_order_and_tickets_expire():
    n = 100
    eng = _engine(n, ring_depth=3)
    pos, quat = _frame(n)
    eng.step(0, pos, quat, [5])
    eng.step(4, pos, quat, [9, 5])
    res = eng.step(6, pos, quat, [9, 7, 5])  # engine order: 9 and 5 (old, grouped), then the new 7
    assert res.keys == [9, 7, 5] and res.timediffs == [2, 0, 6]
    td = (ctypes.c_int64 * 3)()
    _lib.check(eng.lib.sdb_engine_wait(eng._handle, 2, None, td, None))
    assert list(td) == [2, 0, 6]
    assert res.sums().shape == (3, 16)
    old = [eng.step(7 + i, pos, quat, [5]) for i in range(4)]
    assert eng.lib.sdb_engine_wait(eng._handle, 2, None, td, None) != 0  # three later steps reused its slot
    assert all(r.sums().shape == (1, 16) for r in old)  # results are copied out before their slot is reused


def test_drop_and_recreate_and_arbitrary_keys():
    n = 50
    eng = _engine(n)
    pos, quat = _frame(n)
    eng.step(0, pos, quat, ["a", ("b", 1), 3])
    eng.step(2, pos, quat, ["a", ("b", 1), 3])
    assert list(eng.origins.keys()) == ["a", ("b", 1), 3] and len(eng.origins) == 3
    eng.drop_origin("a")
    assert "a" not in eng.origins and len(eng.origins) == 2
    res = eng.step(5, pos, quat, ["a", 3])
    assert res.timediffs == [0, 5]
    origins, _ = eng.last_descriptors()
    assert [bool(f & ORIGIN_NEW) for f in origins["flags"]] == [False, True]


def test_batched_run_equals_stepping():
    n = 400
    schedule = list(exponential_schedule(120, 5, 10, 6))
    pos, quat = _frame(n)
    a, b = _engine(n), _engine(n)
    single = [a.step(t, pos, quat, active) for t, active in schedule]
    batch = b.run([(t, pos, quat, active) for t, active in schedule])
    assert len(batch) == len(single)
    for x, y in zip(single, batch):
        assert x.keys == y.keys and x.timediffs == y.timediffs
    sa, sb = a.stats(), b.stats()
    assert (sa.frames, sa.updates, sa.pool_frames) == (sb.frames, sb.updates, sb.pool_frames)
    oa, sga = a.last_descriptors()
    ob, sgb = b.last_descriptors()
    assert oa["timediff"].tolist() == ob["timediff"].tolist() and sga["count"].tolist() == sgb["count"].tolist()


def test_set_trackers_changes_the_descriptor_pointers_only_when_trackers_exist():
    n = 128
    eng = DynamicsEngine(n, [30.0, 30.0, 1, 0, 0, 0], None, dry=True)  # no wave number: no trackers, no struct
    pos, quat = _frame(n)
    eng.step(0, pos, quat, [0])
    origins, _ = eng.last_descriptors()
    assert origins["d_flags"][0] == 0 and origins["d_status"][0] == 0
    eng.set_trackers([{"name": "tau_X", "threshold": 0.3}])
    eng.step(1, pos, quat, [0])
    origins, _ = eng.last_descriptors()
    assert origins["d_flags"][0] != 0 and origins["d_status"][0] != 0
    with pytest.raises(ValueError):
        eng.set_trackers([{"name": "bad", "threshold": -1.0}])


def test_host_copy_pool_copies_exactly_what_memcpy_would():
    """The engine's multi-threaded staging copy (pageable -> pinned) on caller buffers: every byte, for sizes below
    and above the threshold at which the worker threads take part, with unaligned ends."""
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for nbytes in (0, 1, 4096, 4 * 2 ** 20 - 3, 28_000_000 + 7):
        src = rng.integers(0, 255, nbytes + 64, dtype=np.uint8)
        dst = np.zeros(nbytes + 64, dtype=np.uint8)
        lib.sdb_test_host_copy(dst.ctypes.data + 3, src.ctypes.data + 5, nbytes)
        assert np.array_equal(dst[3 : 3 + nbytes], src[5 : 5 + nbytes]) and not dst[:3].any() and not dst[3 + nbytes :].any()


def test_large_host_frames_go_through_the_threaded_staging_copy():
    n = 400_000  # 4.8 MB of positions, 6.4 MB of quaternions: above the threshold of the copy pool
    eng = _engine(n)
    pos, quat = _frame(n)
    for t in range(4):
        res = eng.step(t, pos, quat, [0, 1])
    assert res.timediffs == [3, 3] and eng.stats().h2d_bytes == 28 * n

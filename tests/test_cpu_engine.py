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


# ---- randomised runs against a model of the reference's loop ---------------------------------------------
class _LoopModel:
    """What read/_read.py:128-168 keeps per keyframe, restated for arbitrary call sequences: the origin's first
    timestep, the frame it saw last (``TrackedMotion.previous_*``, _util.py:155-156), its largest time difference so
    far (decides the literal passage-time rule of relaxations.py:215-218) and how many frames it has seen."""

    def __init__(self):
        self.t0, self.last_frame, self.max_td, self.updates, self.order = {}, {}, {}, {}, []
        self.frames = 0

    def step(self, frame_id, timestep, keys, zero_step=False):
        out = []
        for k in keys:
            new = k not in self.t0
            if new:
                self.t0[k], self.max_td[k], self.updates[k] = timestep, -1, 0
                self.order.append(k)
            td = timestep - self.t0[k]
            out.append((k, td, new, (not new) and td < self.max_td[k], self.last_frame.get(k)))
        if not zero_step:  # a zero step (rows of the current state once more) leaves the origin's history alone
            for k in keys:
                self.max_td[k] = max(self.max_td[k], timestep - self.t0[k])
                self.updates[k] += 1
                self.last_frame[k] = frame_id
            self.frames += 1
        return out

    def drop(self, k):
        for d in (self.t0, self.last_frame, self.max_td, self.updates):
            d.pop(k, None)
        self.order.remove(k)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("SDB_TEST_SEEDS", "6"))))
def test_random_call_sequences_follow_the_loop_model(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.choice([37, 260, 1500]))
    eng = _engine(n, ring_depth=int(rng.integers(2, 5)), chunk=int(rng.choice([0, 0, 3])))
    model = _LoopModel()
    pos, quat = _frame(n, seed)
    timestep, next_key, frame_id = 0, 0, 0
    state_ptrs = {}
    for op in range(150):
        live = list(model.t0)
        roll = rng.random()
        if live and roll < 0.08:
            k = live[int(rng.integers(len(live)))]
            eng.drop_origin(k)
            model.drop(k)
            state_ptrs.pop(k, None)
            continue
        if live and roll < 0.12:  # a call the engine must refuse without side effects
            with pytest.raises(RuntimeError):
                eng.step(timestep, pos[:-1], quat, live[:1] + [next_key + 1000])
            assert next_key + 1000 not in eng.origins and len(eng.origins) == len(live)
            continue
        if roll < 0.15:  # Relaxations.set_mol_relax mid-run: passage state afresh, the bookkeeping untouched
            count = int(rng.integers(0, 4))
            eng.set_trackers([{"name": f"tau_{i}", "threshold": 0.1 * (i + 1), "last_passage": bool(i % 2)} for i in range(count)])
            for k in live:
                view = eng.origins[k]
                assert (view.timestep, view.updates, view.max_timediff) == (model.t0[k], model.updates[k], model.max_td[k])
            continue
        if live and roll < 0.20:  # a batch of frames in one native call
            batch = []
            timestep = max(timestep, max(model.t0.values()))  # (no frame before an origin's first one)
            for _ in range(int(rng.integers(2, 5))):
                timestep += int(rng.choice([1, 3, 8]))
                keys = [int(k) for k in model.t0 if rng.random() < 0.7] + [next_key]
                next_key += 1
                frame_id += 1
                batch.append((timestep, keys, model.step(frame_id, timestep, keys)))
            got = eng.run([(t, pos, quat, keys) for t, keys, _ in batch])
            for res, (t, keys, want) in zip(got, batch):
                assert res.keys == keys and res.timediffs == [td for _, td, *_ in want]
            st = eng.stats()
            assert st.frames == model.frames and st.pool_frames - st.pool_free == len(set(model.last_frame.values()))
            assert eng.origins.keys() == model.order
            continue
        zero_step = bool(live) and roll < 0.24
        keys = [k for k in live if rng.random() < 0.6]
        if not zero_step:
            for _ in range(int(rng.integers(0, 3)) if live else 1):
                keys.append(next_key)
                next_key += 1
        if not keys:
            continue
        rng.shuffle(keys)
        keys = [int(k) for k in keys]
        if not zero_step:
            timestep = max(0, timestep + int(rng.choice([1, 1, 2, 5, 17, -3])))
            frame_id += 1
        if any(k in model.t0 and timestep < model.t0[k] for k in keys):
            # before an origin's first frame: refused (passage times are unsigned), nothing created, nothing stepped
            fresh = [k for k in keys if k not in model.t0]
            before = eng.stats()
            with pytest.raises(OverflowError):
                eng.step(timestep, pos, quat, keys, zero_step=zero_step)
            after = eng.stats()
            assert (after.frames, after.updates, after.pool_frames, after.pool_free) == (
                before.frames, before.updates, before.pool_frames, before.pool_free)
            assert not any(k in eng.origins for k in fresh) and eng.origins.keys() == model.order
            frame_id -= 0 if zero_step else 1
            continue
        prev_before = {k: eng.origins[k]._info.d_prev for k in keys if k in model.t0}
        res = eng.step(timestep, None if zero_step else pos, None if zero_step else quat, keys, zero_step=zero_step)
        want = model.step(frame_id, timestep, keys, zero_step)
        # rows come back in the caller's order with timediff = frame.timestep - origin.timestep
        assert res.keys == keys and res.timediffs == [td for _, td, *_ in want]
        origins, segments = eng.last_descriptors()
        assert len(origins) == len(keys)
        # the segments tile the origin records in order
        assert segments["first"].tolist() == np.cumsum([0] + segments["count"].tolist())[:-1].tolist()
        assert int(segments["count"].sum()) == len(keys) and np.all(segments["count"] > 0)
        # every record can be matched to one key: same state pointer as ever, same timediff and flags as the model
        by_ptr = {}
        for k in keys:
            view = eng.origins[k]
            seen = state_ptrs.setdefault(k, list(view.ptrs))
            for slot, ptr in enumerate(view.ptrs):  # an origin's state never moves (no trackers: no flag / status pointers)
                seen[slot] = seen[slot] or ptr
                assert ptr in (seen[slot], 0) and (slot > 0 or ptr)
            by_ptr[view.ptrs[0]] = k
        assert len(by_ptr) == len(keys)
        seg_of = np.repeat(np.arange(len(segments)), segments["count"])
        expect = {k: (td, new, literal, last) for k, td, new, literal, last in want}
        for rec, s in zip(origins, seg_of):
            k = by_ptr[int(rec["d_delta"])]
            td, new, literal, last = expect[k]
            assert int(rec["timediff"]) == td
            assert bool(rec["flags"] & ORIGIN_NEW) == new and bool(rec["flags"] & ORIGIN_LITERAL) == literal, (op, k)
            seg = segments[s]
            if zero_step:
                assert seg["flags"] & SEG_ZERO_STEP
            elif new:
                assert seg["d_prev"] == 0  # created now: previous sample == this frame
            else:
                assert int(seg["d_prev"]) == prev_before[k] != 0  # steps from the frame IT saw last
        # origins grouped in one segment share their last-seen frame (the converse need not hold: long groups are split)
        for s in range(len(segments)):
            members = {expect[by_ptr[int(r["d_delta"])]][3] for r in origins[segments["first"][s] : segments["first"][s] + segments["count"][s]]}
            assert len(members) == 1 or zero_step
        # afterwards: the pool holds exactly one sample per distinct last-seen frame, shared by its origins
        st = eng.stats()
        assert st.pool_frames - st.pool_free == len(set(model.last_frame.values()))
        assert st.frames == model.frames
        prev_now = {}
        for k in model.t0:
            view = eng.origins[k]
            assert (view.timestep, view.updates, view.max_timediff) == (model.t0[k], model.updates[k], model.max_td[k])
            prev_now.setdefault(model.last_frame[k], set()).add(view._info.d_prev)
        assert all(len(ptrs) == 1 for ptrs in prev_now.values())
        assert len({next(iter(p)) for p in prev_now.values()}) == len(prev_now)
        assert eng.origins.keys() == model.order


def test_refused_frames_leave_nothing_behind():
    n = 90
    eng = _engine(n)
    pos, quat = _frame(n)
    eng.step(10, pos, quat, [0])
    # a frame before origin 0's first one: refused by the native engine before it creates origin 1
    with pytest.raises(OverflowError):
        eng.step(7, pos, quat, [1, 0])
    assert 1 not in eng.origins and eng.origins.keys() == [0] and eng.stats().frames == 1
    # a frame of the wrong shape: refused before the key gets a record, so its later creation starts at ITS timestep
    with pytest.raises(RuntimeError):
        eng.step(11, pos[:-1], quat, [0, 1])
    assert 1 not in eng.origins
    res = eng.step(15, pos, quat, [0, 1])
    assert res.timediffs == [5, 0] and eng.origins[1].timestep == 15
    # a batch that fails at its third frame: the first two are applied, the keys of the refused frame are forgotten
    with pytest.raises(OverflowError):
        eng.run([(16, pos, quat, [0, 1]), (17, pos, quat, [0, 1, 2]), (12, pos, quat, [0, 1, 3]), (18, pos, quat, [0, 4])])
    assert eng.origins.keys() == [0, 1, 2] and 3 not in eng.origins and 4 not in eng.origins
    assert eng.stats().frames == 4 and eng.origins[2].timestep == 17 and eng.origins[0].updates == 4
    res = eng.step(20, pos, quat, [3, 0])
    assert res.timediffs == [0, 10]
    with pytest.raises(RuntimeError):
        eng.run([(21, pos, quat, [0, 5]), (22, pos[:-1], quat, [0, 6])])  # nothing of this batch reaches the engine
    assert 5 not in eng.origins and 6 not in eng.origins and eng.stats().frames == 5

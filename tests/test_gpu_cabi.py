"""The C-ABI drop-in boundary exercised the way a host WITHOUT torch would use it: ctypes + numpy host buffers
in, rows out (``sdb_engine_create`` / ``sdb_engine_step`` / ``sdb_engine_wait`` on the default stream).  This is
the binding INTEGRATION.md shows for the reference's own loop (read/_read.py:128-168)."""

import ctypes as C

import numpy as np
import pytest

from _helpers import WAVE_NUMBER, compare_rows, oracle_process

pytestmark = pytest.mark.gpu


def _config(n, box, trackers):
    from sdanalysis_b200 import _lib
    from sdanalysis_b200.engine import _cut_direct, _cut_sqrt, build_trackers
    from sdanalysis_b200.molecules import Trimer

    cfg = _lib.SdbEngineConfig()
    cfg.n = n
    for i in range(6):
        cfg.box[i] = float(box[i])
    cfg.wave_number = WAVE_NUMBER
    distance = np.pi / (2 * WAVE_NUMBER)
    cfg.com_cut_lt = float(_cut_sqrt(distance, False))
    specs = build_trackers(trackers)
    cfg.n_trackers = len(specs)
    for i, spec in enumerate(specs):
        cut = _cut_direct if spec.type == "rotation" else _cut_sqrt
        t = cfg.trackers[i]
        t.is_rotation, t.is_last, t.bit = int(spec.type == "rotation"), int(spec.last_passage), spec.bit
        t.cut_gt, t.cut_lt, t.cut_irr = float(cut(spec.threshold, True)), float(cut(spec.threshold, False)), float(cut(spec.cutoff, True))
    cfg.compute_sums = 1
    cfg.overlap_k = int(n * 0.1)
    sites = np.asarray(Trimer().positions, dtype=np.float32)[:, :2]
    cfg.n_sites = len(sites)
    for i, v in enumerate(sites.reshape(-1)):
        cfg.sites_xy[i] = float(v)
    cfg.ring_depth = 3
    cfg.angular_resolution = 360
    return cfg, specs


def test_engine_through_ctypes_with_host_buffers_only():
    from sdanalysis_b200 import _lib
    from sdanalysis_b200.engine import MAX_STATUS, default_tracker_definitions, finalize_row
    from sdanalysis_b200.synthetic import exponential_schedule, synthetic_frames

    lib = _lib.load()  # ctypes.CDLL: no torch involved below
    n = 3001
    frames = list(synthetic_frames(exponential_schedule(120, 8, 15, 5), n, seed=23, sigma_r=0.05, sigma_theta=0.12))
    box = frames[0][1].box
    distance = np.pi / (2 * WAVE_NUMBER)
    cfg, specs = _config(n, box, default_tracker_definitions(distance))
    engine = C.c_void_p()
    assert lib.sdb_engine_create(C.byref(cfg), C.byref(engine)) == 0, lib.sdb_last_error()
    rows = []
    pending = []
    try:
        for indexes, frame in frames:
            keys = (C.c_int64 * len(indexes))(*indexes)
            pos = np.ascontiguousarray(frame.position, dtype=np.float32)
            quat = np.ascontiguousarray(frame.orientation, dtype=np.float32)
            f = _lib.SdbFrame()
            f.timestep = frame.timestep
            f.pos, f.quat, f.where = pos.ctypes.data, quat.ctypes.data, _lib.SDB_MEM_HOST
            f.keys, f.n_keys = C.addressof(keys), len(indexes)
            ticket = lib.sdb_engine_step(engine, C.byref(f), None)  # NULL stream: CUDA's default stream
            assert ticket >= 0, lib.sdb_last_error()
            pending.append((ticket, list(indexes)))
            if len(pending) > 1:  # one frame in flight while the next one is staged
                rows.extend(_collect(lib, engine, pending.pop(0), n, distance, finalize_row))
        while pending:
            rows.extend(_collect(lib, engine, pending.pop(0), n, distance, finalize_row))
        ref_rows, summaries, near, keyframes = oracle_process(frames)
        order = {(r["keyframe"], r["time"]): r for r in rows}
        compare_rows([order[(r["keyframe"], r["time"])] for r in ref_rows], ref_rows, n)
        # passage times through the raw device pointers the engine reports
        cudart = C.CDLL("libcudart.so.12")
        info = _lib.SdbOriginInfo()
        for key, ref in summaries.items():
            assert lib.sdb_engine_origin(engine, key, C.byref(info)) == 0
            status = np.empty((len(specs), info.stride), dtype=np.uint32)
            flags = np.empty(info.stride, dtype=np.uint8)
            assert cudart.cudaMemcpy(C.c_void_p(status.ctypes.data), C.c_void_p(info.d_status), C.c_size_t(status.nbytes), 2) == 0
            assert cudart.cudaMemcpy(C.c_void_p(flags.ctypes.data), C.c_void_p(info.d_flags), C.c_size_t(flags.nbytes), 2) == 0
            for i, spec in enumerate(specs):
                col = status[i, :n].astype(np.int64)
                if spec.last_passage:
                    col[((flags[:n] >> spec.bit) & 3) != 2] = MAX_STATUS
                mask = ~near[key][spec.name]
                assert np.array_equal(col[mask], ref[spec.name][mask]), (key, spec.name)
    finally:
        lib.sdb_engine_destroy(engine)


def _collect(lib, engine, item, n, distance, finalize_row):
    ticket, indexes = item
    sums = np.empty((len(indexes), 16), dtype=np.float64)
    timediffs = np.empty(len(indexes), dtype=np.int64)
    assert lib.sdb_engine_wait(engine, ticket, sums.ctypes.data, timediffs.ctypes.data, None) == 0, lib.sdb_last_error()
    out = []
    for key, td, s in zip(indexes, timediffs.tolist(), sums):
        row = finalize_row(s, td, n, distance=distance, overlap_k=int(n * 0.1), n_sites=3)
        row["keyframe"] = key
        out.append(row)
    return out

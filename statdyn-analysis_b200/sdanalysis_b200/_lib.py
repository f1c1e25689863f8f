"""ctypes binding of ``libsdb200.so`` (C-ABI declared in ``include/sdb200.h``).

There is no CPU fallback: if the shared library is missing or a CUDA device is not available the
functions here raise, loudly.  ``torch`` is used only as the owner of device memory and streams.
"""

import ctypes
from ctypes import POINTER, c_double, c_float, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p
from pathlib import Path

LIB_NAME = "libsdb200.so"
LIB_PATH = Path(__file__).resolve().parent / LIB_NAME

SDB_MAX_TRACKERS = 6
SDB_NSUM = 16
SDB_STEP_PREDICTED = 1
SDB_NO_STATUS = 0xFFFFFFFF
SDB_MAX_K = 16

S_DISP, S_DISP2, S_DISP4, S_COMSTRUCT, S_ROT, S_ROT2, S_COS, S_COS2, S_ROT4, S_D2R2 = range(10)
S_STRUCT, S_OVERLAP, S_BADQUAT = 10, 11, 12

ORIGIN_NEW, ORIGIN_LITERAL, ORIGIN_HISTORY = 1, 2, 4
HISTORY_WORDS = 16
SDB_MAX_INC_SETS = 4
SEG_NO_ROTATION, SEG_ZERO_STEP = 1, 2

# Every symbol include/sdb200.h declares; tests check that the library exports all of them.
EXPORTS = (
    "sdb_dyn_step",
    "sdb_dyn_increments",
    "sdb_dyn_increments_bytes",
    "sdb_dyn_num_parts",
    "sdb_dyn_norms",
    "sdb_pack_frame",
    "sdb_struct",
    "sdb_overlap",
    "sdb_overlap_scratch_bytes",
    "sdb_overlap_ws_bytes",
    "sdb_overlap_prepare",
    "sdb_overlap_resolve",
    "sdb_overlap_stats",
    "sdb_multicast_copy",
    "sdb_frame_step",
    "sdb_rdf_histogram",
    "sdb_voronoi_counts",
    "sdb_isf_origins",
    "sdb_isf",
    "sdb_profile_enable",
    "sdb_profile_collect",
    "sdb_first_passage",
    "sdb_last_passage",
    "sdb_motion_general",
    "sdb_sums_general",
    "sdb_sums_general_parts",
    "sdb_neighbours",
    "sdb_neighbours_ex",
    "sdb_neighbours_scratch_bytes",
    "sdb_orientational_order",
    "sdb_last_error",
    "sdb_version",
    "sdb_launch_count",
    "sdb_test_rot_increment",
    "sdb_test_wrap2d",
    "sdb_test_host_copy",
    "sdb_upload_async",
    "sdb_host_register",
    "sdb_host_unregister",
    "sdb_engine_create",
    "sdb_engine_destroy",
    "sdb_engine_set_trackers",
    "sdb_engine_step",
    "sdb_engine_wait",
    "sdb_engine_device_rows",
    "sdb_engine_run",
    "sdb_engine_drop",
    "sdb_engine_origin",
    "sdb_engine_num_origins",
    "sdb_engine_keys",
    "sdb_engine_stats",
    "sdb_engine_last_descriptors",
)


class SdbTracker(ctypes.Structure):
    _fields_ = [
        ("is_rotation", c_int32),
        ("is_last", c_int32),
        ("bit", c_int32),
        ("cut_gt", c_float),
        ("cut_lt", c_float),
        ("cut_irr", c_float),
    ]


class SdbDynParams(ctypes.Structure):
    _fields_ = [
        ("n", c_int32),
        ("stride", c_int32),
        ("lx", c_float),
        ("ly", c_float),
        ("xy", c_float),
        ("com_cut_lt", c_float),
        ("n_trackers", c_int32),
        ("do_sums", c_int32),
        ("trackers", SdbTracker * SDB_MAX_TRACKERS),
    ]


class SdbOrigin(ctypes.Structure):
    _fields_ = [
        ("d_delta", c_void_p),
        ("d_flags", c_void_p),
        ("d_status", c_void_p),
        ("timediff", c_uint32),
        ("flags", c_uint32),
    ]


class SdbSegment(ctypes.Structure):
    _fields_ = [
        ("d_prev", c_void_p),
        ("first", c_int32),
        ("count", c_int32),
        ("flags", c_uint32),
        ("reserved", c_uint32),
    ]


# ---- native engine (csrc/engine.cu) -------------------------------------------------------------------
SDB_MEM_DEVICE, SDB_MEM_HOST, SDB_MEM_PINNED = 0, 1, 2
SDB_ENGINE_DRY, SDB_ENGINE_NO_PREDICT, SDB_ENGINE_NO_INCREMENTS = 1, 2, 4
SDB_FRAME_NO_SUMS, SDB_FRAME_ZERO_STEP, SDB_FRAME_ROWS_DEVICE = 1, 2, 4


class SdbEngineConfig(ctypes.Structure):
    _fields_ = [
        ("n", c_int32),
        ("box", c_float * 6),
        ("wave_number", c_double),
        ("com_cut_lt", c_float),
        ("n_trackers", c_int32),
        ("trackers", SdbTracker * SDB_MAX_TRACKERS),
        ("compute_sums", c_int32),
        ("overlap_k", c_int32),
        ("n_sites", c_int32),
        ("sites_xy", c_float * 16),
        ("struct_mode", c_int32),
        ("compute_isf", c_int32),
        ("angular_resolution", c_int32),
        ("ring_depth", c_int32),
        ("inc_max_per_segment", c_int32),
        ("chunk", c_int32),
        ("flags", c_uint32),
    ]


class SdbFrame(ctypes.Structure):
    _fields_ = [
        ("timestep", c_int64),
        ("pos", c_void_p),
        ("quat", c_void_p),
        ("where", c_int32),
        ("flags", c_uint32),
        ("keys", c_void_p),
        ("n_keys", c_int32),
        ("reserved", c_int32),
        ("d_rows_out", c_void_p),
    ]


class SdbOriginInfo(ctypes.Structure):
    _fields_ = [
        ("timestep", c_int64),
        ("updates", c_int64),
        ("max_timediff", c_int64),
        ("d_delta", c_void_p),
        ("d_flags", c_void_p),
        ("d_status", c_void_p),
        ("d_prev", c_void_p),
        ("prev_has_orientation", c_int32),
        ("stride", c_int32),
    ]


class SdbEngineStats(ctypes.Structure):
    _fields_ = [
        ("frames", c_int64),
        ("updates", c_int64),
        ("h2d_bytes", c_int64),
        ("d2h_bytes", c_int64),
        ("pool_frames", c_int64),
        ("pool_free", c_int64),
        ("device_bytes", c_int64),
        ("last_segments", c_int32),
        ("last_inc_sets", c_int32),
        ("last_step_flags", c_int32),
        ("capacity", c_int32),
        ("d_ov_ws", c_void_p),
        ("d_inc_ws", c_void_p),
    ]


# numpy views of the two descriptor structs (filled vectorised by the engine)
ORIGIN_DTYPE = [("d_delta", "<u8"), ("d_flags", "<u8"), ("d_status", "<u8"), ("timediff", "<u4"), ("flags", "<u4")]
SEGMENT_DTYPE = [("d_prev", "<u8"), ("first", "<i4"), ("count", "<i4"), ("flags", "<u4"), ("reserved", "<u4")]

_lib = None


class SdbError(RuntimeError):
    """A call into libsdb200.so failed."""


def load():
    """Load the CUDA library once; raise ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root). "
            "sdanalysis_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    P = c_void_p
    sig = {
        "sdb_dyn_step": (c_int32, [POINTER(SdbDynParams), P, P, P, P, c_int32, P, c_int32, P, P, P, c_int32, P, c_int32, P]),
        "sdb_dyn_increments_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
        "sdb_dyn_increments": (c_int32, [POINTER(SdbDynParams), P, P, P, POINTER(c_void_p), POINTER(c_uint32), c_int32, P, P]),
        "sdb_dyn_num_parts": (c_int32, [c_int32]),
        "sdb_dyn_norms": (c_int32, [P, c_int32, c_int32, P, P, P]),
        "sdb_pack_frame": (c_int32, [P, P, c_int32, c_int32, P, P, P]),
        "sdb_struct": (c_int32, [P, c_int32, c_int32, c_int32, POINTER(c_float), c_int32, c_double, c_int32, P, P]),
        "sdb_overlap": (c_int32, [P, c_int32, c_int32, c_int32, c_int32, c_int32, P, P, P]),
        "sdb_overlap_scratch_bytes": (c_size_t, [c_int32, c_int32]),
        "sdb_overlap_ws_bytes": (c_size_t, [c_int32, c_int32]),
        "sdb_overlap_prepare": (c_int32, [POINTER(SdbDynParams), P, P, P, c_int32, P, c_int32, c_int32, c_int32, P, c_int32, P]),
        "sdb_overlap_resolve": (c_int32, [P, c_int32, c_int32, c_int32, c_int32, P, c_int32, P, P]),
        "sdb_overlap_stats": (c_int32, [P, c_int32, c_int32, c_int32, P, P]),
        "sdb_multicast_copy": (c_int32, [P, P, ctypes.c_size_t, c_int32, P]),
        "sdb_frame_step": (
            c_int32,
            [POINTER(SdbDynParams), P, P, P, P, P, c_int32, c_int32, c_int32, c_int32, c_int32, P, P, P, P, c_int32,
             c_int32, POINTER(c_float), c_int32, ctypes.c_double, c_int32, POINTER(c_void_p), POINTER(c_uint32),
             c_int32, P, c_int32, P],
        ),
        "sdb_voronoi_counts": (c_int32, [POINTER(c_float), c_int32, P, P, c_int32, c_float, P, P, P]),
        "sdb_rdf_histogram": (c_int32, [POINTER(c_float), c_int32, c_int32, P, c_float, c_float, c_int32, P, P]),
        "sdb_isf_origins": (c_int32, [P, c_int32, c_int32, c_int32, c_double, c_int32, P, P]),
        "sdb_isf": (c_int32, [P, P, c_int32, c_int32, c_double, c_int32, P, P]),
        "sdb_profile_enable": (None, [c_int32]),
        "sdb_profile_collect": (c_int32, [POINTER(c_double), POINTER(c_int64), POINTER(c_int64)]),
        "sdb_first_passage": (c_int32, [P, c_int32, c_int32, c_double, c_int64, P, P]),
        "sdb_last_passage": (c_int32, [P, c_int32, c_int32, c_double, c_double, c_int64, P, P, P]),
        "sdb_motion_general": (c_int32, [POINTER(c_float), c_int32, c_int32, P, P, P, P, c_int32, P, P, P, P]),
        "sdb_sums_general": (c_int32, [c_int32, P, P, c_float, P, P, P, P, P]),
        "sdb_sums_general_parts": (c_int32, [c_int32]),
        "sdb_neighbours": (c_int32, [POINTER(c_float), c_int32, P, P, c_float, c_int32, P, P, P, P, P, c_size_t, P]),
        "sdb_neighbours_ex": (c_int32, [POINTER(c_float), c_int32, P, P, c_float, c_int32, c_int32, P, P, P, P, P, c_size_t, P]),
        "sdb_neighbours_scratch_bytes": (c_size_t, [POINTER(c_float), c_int32, c_float]),
        "sdb_orientational_order": (c_int32, [P, c_int32, c_int32, P, P, P, P]),
        "sdb_last_error": (ctypes.c_char_p, []),
        "sdb_version": (ctypes.c_char_p, []),
        "sdb_launch_count": (c_uint64, []),
        "sdb_test_rot_increment": (c_double, [c_float, c_float, c_float, c_float, POINTER(c_int32)]),
        "sdb_test_wrap2d": (None, [c_float, c_float, c_float, c_float, c_float, POINTER(c_float), POINTER(c_float)]),
        "sdb_test_host_copy": (None, [P, P, c_size_t]),
        "sdb_upload_async": (c_int32, [P, P, c_size_t, P]),
        "sdb_host_register": (c_int32, [P, c_size_t]),
        "sdb_host_unregister": (c_int32, [P]),
        "sdb_engine_create": (c_int32, [POINTER(SdbEngineConfig), POINTER(c_void_p)]),
        "sdb_engine_destroy": (None, [P]),
        "sdb_engine_set_trackers": (c_int32, [P, POINTER(SdbTracker), c_int32]),
        "sdb_engine_step": (c_int64, [P, POINTER(SdbFrame), P]),
        "sdb_engine_wait": (c_int32, [P, c_int64, P, P, P]),
        "sdb_engine_device_rows": (c_int32, [P, c_int64, POINTER(c_void_p), POINTER(c_int32)]),
        "sdb_engine_run": (c_int32, [P, POINTER(SdbFrame), c_int32, P, P, P, P]),
        "sdb_engine_drop": (c_int32, [P, c_int64]),
        "sdb_engine_origin": (c_int32, [P, c_int64, POINTER(SdbOriginInfo)]),
        "sdb_engine_num_origins": (c_int32, [P]),
        "sdb_engine_keys": (c_int32, [P, POINTER(c_int64), c_int32]),
        "sdb_engine_stats": (c_int32, [P, POINTER(SdbEngineStats)]),
        "sdb_engine_last_descriptors": (c_int32, [P, P, c_int32, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SdbError(f"libsdb200 error {rc}: {load().sdb_last_error().decode()}")


def launch_count():
    return int(load().sdb_launch_count())


_torch = None


def torch():
    """Import torch lazily (the first import can take a minute on a fresh box)."""
    global _torch
    if _torch is None:
        import torch as _t

        _torch = _t
    return _torch


def require_cuda(device=None):
    """Return a torch CUDA device or raise: there is no CPU path."""
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError(
            "sdanalysis_b200 needs a CUDA device (B200, sm_100a); torch.cuda.is_available() is False "
            "and there is no CPU fallback."
        )
    load()
    if device is None:
        return t.device("cuda", t.cuda.current_device())
    return t.device(device)


def ptr(tensor):
    return 0 if tensor is None else tensor.data_ptr()


def stream_ptr(device=None):
    """Raw ``cudaStream_t`` of torch's current stream on ``device`` (the C call, not the Python wrapper:
    ``torch.cuda.current_stream`` costs ~12 us per call)."""
    t = torch()
    if isinstance(device, str):
        device = t.device(device)
    index = device.index if isinstance(device, t.device) else device
    if index is None:
        index = t.cuda.current_device()
    try:
        return t._C._cuda_getCurrentRawStream(index)
    except AttributeError:  # older / newer torch without the private hook
        return t.cuda.current_stream(device).cuda_stream


def current_stream(device):
    """``torch.cuda.current_stream(device)`` without its Python-side argument handling (~12 us -> ~2 us);
    ``device`` is a ``torch.device`` with an index."""
    t = torch()
    try:
        sid, di, dt = t._C._cuda_getCurrentStream(device.index)
        return t.cuda.Stream(stream_id=sid, device_index=di, device_type=dt)
    except Exception:
        return t.cuda.current_stream(device)


class _NullContext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_CONTEXT = _NullContext()


def device_context(device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (the context manager
    costs ~10 us; with one process per GPU the device is always current)."""
    t = torch()
    try:
        if t._C._cuda_getDevice() == device.index:
            return _NULL_CONTEXT
    except Exception:
        pass
    return t.cuda.device(device)


# ---- page-locked host ranges registered by this package (a memory-mapped trajectory file pinned with
# cudaHostRegister): frames inside them are uploaded without the staging copy (SDB_MEM_PINNED)
_pinned_ranges = []


def register_pinned(address, nbytes):
    _pinned_ranges.append((int(address), int(address) + int(nbytes)))


def unregister_pinned(address):
    _pinned_ranges[:] = [r for r in _pinned_ranges if r[0] != int(address)]


def is_pinned(address, nbytes=1):
    for lo, hi in _pinned_ranges:
        if lo <= address and address + nbytes <= hi:
            return True
    return False


def float_array(values):
    arr = (c_float * len(values))(*[float(v) for v in values])
    return arr

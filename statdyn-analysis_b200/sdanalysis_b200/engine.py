"""Batched multi-origin dynamics engine: the B200 replacement for the reference's hot loop.

The reference keeps one ``(Dynamics, Relaxations)`` pair of Python objects per time-origin
("keyframe") and, for every frame, calls ``compute_all`` + ``Relaxations.add`` on each active one
(``read/_read.py:128-168``).  Here all origins of a rank live in HBM as planar arrays and one call to
:meth:`DynamicsEngine.step` advances every active origin with the fused CUDA kernel
(``csrc/dyn_step.cu``) plus the ``struct`` and ``overlap`` kernels.

Since round 2 the engine itself is native code (``csrc/engine.cu``, C-ABI ``sdb_engine_*``): origin
records, the pool of previous samples, segment grouping, descriptor / result rings, staging of host
frames.  This module is the thin Python face of it: argument checking with the reference's error
conventions, threshold digestion, and the closed forms that turn the 16 sums of a (frame, origin)
into the 15 quantities of ``compute_all``.

Per-origin state (N molecules, ``stride`` = N rounded up to 4):

=============  ======================  =========================================================
``delta``      ``float32[3][stride]``  accumulated dx, dy, rotation about z (``_util.py:123-124``)
``flags``      ``uint8[stride]``       fired bits of the first-passage trackers + last-passage state
``status``     ``uint32[T][stride]``   passage times (``relaxations.py:199``), written only
=============  ======================  =========================================================

Previous samples (``TrackedMotion.previous_*``, ``_util.py:155-156``) are shared, reference-counted
planar frames ``float32[4][stride]`` = (x, y, qw, qz): after a frame is processed every origin that
was active points at it, so in the dense (linear) schedule one frame is shared by all origins and in
the exponential schedule at most one frame per origin is alive.
"""

import ctypes
import weakref
from collections import deque

import numpy as np

from . import _lib
from ._lib import (
    ORIGIN_DTYPE,
    S_BADQUAT,
    S_COMSTRUCT,
    S_COS,
    S_COS2,
    S_D2R2,
    S_DISP,
    S_DISP2,
    S_DISP4,
    S_OVERLAP,
    S_ROT,
    S_ROT2,
    S_ROT4,
    S_STRUCT,
    SDB_MAX_TRACKERS,
    SDB_NSUM,
    SEGMENT_DTYPE,
)

F32 = np.float32
MAX_STATUS = 2 ** 32 - 1

ALL_QUANTITIES = (
    "time",
    "mean_displacement",
    "msd",
    "mfd",
    "alpha",
    "scattering_function",
    "com_struct",
    "mean_rotation",
    "msr",
    "rot1",
    "rot2",
    "alpha_rot",
    "gamma",
    "overlap",
    "struct",
)


# ----------------------------------------------------------------------------------------------
# threshold digestion (host): exact cut points for "f32 array <op> python float" comparisons
# ----------------------------------------------------------------------------------------------
def _cut_sqrt(threshold, strict_greater):
    """Smallest f32 ``x`` with ``sqrt32(x) > thr`` (strict_greater) or ``sqrt32(x) >= thr``.

    The reference compares ``np.linalg.norm(delta, axis=1)`` (= ``fl32(sqrt(disp2))``) with a Python
    float, which numpy does in f32 (``relaxations.py:216``, ``dynamics.py:425``).  sqrt is monotone,
    so the comparison is equivalent to a comparison of ``disp2`` with this cut.
    """
    thr = F32(threshold)
    if np.isnan(thr):
        return F32(np.inf)
    ok = (lambda v: np.sqrt(v) > thr) if strict_greater else (lambda v: np.sqrt(v) >= thr)
    x = F32(thr * thr) if thr > 0 else F32(0)
    if not np.isfinite(x):
        return F32(np.inf)
    while x > 0 and ok(x):
        x = np.nextafter(x, F32(0), dtype=F32)
    while not ok(x):
        nx = np.nextafter(x, F32(np.inf), dtype=F32)
        if nx == x:
            return F32(np.inf)
        x = nx
    return x


def _cut_direct(threshold, strict_greater):
    """Cut for quantities compared directly (|dtheta|): ``v > thr``  <=>  ``v >= nextafter(thr)``."""
    thr = F32(threshold)
    return np.nextafter(thr, F32(np.inf), dtype=F32) if strict_greater else thr


def default_tracker_definitions(distance):
    """The six default trackers of ``Relaxations.__init__`` (``relaxations.py:67-81``)."""
    return [
        {"name": "tau_D", "threshold": 3 * distance},
        {"name": "tau_F", "threshold": distance},
        {"name": "tau_L", "threshold": distance, "last_passage": True, "last_passage_cutoff": 3 * distance},
        {"name": "tau_T2", "threshold": np.pi / 2, "type": "rotation"},
        {"name": "tau_T3", "threshold": np.pi / 3, "type": "rotation"},
        {"name": "tau_T4", "threshold": np.pi / 4, "type": "rotation"},
    ]


class TrackerSpec:
    """Validated tracker definition (``Relaxations.set_mol_relax`` + ``create_mol_relaxations``,
    ``relaxations.py:108-133,308-335``)."""

    def __init__(self, item):
        if item.get("name") is None:
            raise ValueError("'name' is a required attribute")
        if item.get("threshold") is None:
            raise ValueError("'threshold' is a required attribute")
        self.name = str(item["name"])
        self.threshold = float(item["threshold"])
        self.last_passage = bool(item.get("last_passage", False))
        self.cutoff = float(item.get("last_passage_cutoff", 3.0 * self.threshold))
        self.type = str(item.get("type", "translation"))
        if self.type not in ("translation", "rotation"):
            raise KeyError(self.type)
        if self.threshold < 0:
            raise ValueError(f"Threshold needs a positive value, got {self.threshold}")
        if self.last_passage and self.cutoff < 0:
            raise ValueError(
                f"When using last passage a positive cutoff value is required, got {self.cutoff}, default is 1.0"
            )


def build_trackers(definitions):
    specs = [TrackerSpec(item) for item in definitions]
    if len(specs) > SDB_MAX_TRACKERS:
        raise NotImplementedError(f"at most {SDB_MAX_TRACKERS} molecular relaxations are supported on the device")
    bit = 0
    for spec in specs:
        spec.bit = bit
        bit += 2 if spec.last_passage else 1
    if bit > 8:
        raise NotImplementedError("tracker state does not fit the per-molecule flag byte")
    return specs


def finalize_row(s, td, n, *, distance=None, overlap_k=None, n_sites=None, strict=True):
    """The 15 quantities of ``Dynamics.compute_all`` from the SDB_S_* sums of one (frame, origin).

    Closed forms of reference ``dynamics.py:183-303``; 0/0 gives NaN exactly like the reference under
    its pinned numpy (SURVEY trap T3).  ``overlap_k`` / ``n_sites`` None leave the quantity NaN.
    """
    nan = float("nan")
    n = float(n)
    msd, mfd = s[S_DISP2] / n, s[S_DISP4] / n
    msr, mfr = s[S_ROT2] / n, s[S_ROT4] / n
    row = dict.fromkeys(ALL_QUANTITIES, nan)
    row["time"] = td
    row["mean_displacement"] = s[S_DISP] / n
    row["msd"] = msd
    row["mfd"] = mfd
    row["alpha"] = mfd / (2 * msd * msd) - 1 if msd > 0 else nan
    if distance is not None:
        row["com_struct"] = s[S_COMSTRUCT] / n
    row["mean_rotation"] = s[S_ROT] / n
    row["msr"] = msr
    row["rot1"] = s[S_COS] / n
    row["rot2"] = s[S_COS2] / n
    row["alpha_rot"] = mfr / (3 * msr * msr) - 1 if msr > 0 else nan
    row["gamma"] = (s[S_D2R2] / n) / (msd * msr) - 1 if msd > 0 and msr > 0 else nan
    if overlap_k is not None:
        if overlap_k == 0:
            if strict:
                raise ZeroDivisionError("division by zero")  # dynamics.py:445 with N < 10
        else:
            row["overlap"] = s[S_OVERLAP] / overlap_k
    if n_sites is not None:
        row["struct"] = s[S_STRUCT] / (n * n_sites)
    return row


def finalize_table(sums, timediffs, n, *, distance=None, overlap_k=None, n_sites=None, strict=True):
    """Vectorised :func:`finalize_row`: ``sums`` is ``(A, SDB_NSUM)``; returns ``{quantity: (A,) array}``
    with the same closed forms and the same NaN conventions (0/0 -> NaN, SURVEY trap T3)."""
    s = np.asarray(sums, dtype=np.float64).reshape(-1, SDB_NSUM)
    a = s.shape[0]
    n = float(n)
    nan = np.full(a, np.nan)
    msd, mfd = s[:, S_DISP2] / n, s[:, S_DISP4] / n
    msr, mfr = s[:, S_ROT2] / n, s[:, S_ROT4] / n
    table = {key: nan for key in ALL_QUANTITIES}
    table["time"] = np.asarray(timediffs, dtype=np.int64)
    table["mean_displacement"] = s[:, S_DISP] / n
    table["msd"] = msd
    table["mfd"] = mfd
    with np.errstate(divide="ignore", invalid="ignore"):
        table["alpha"] = np.where(msd > 0, mfd / (2 * msd * msd) - 1, np.nan)
        table["alpha_rot"] = np.where(msr > 0, mfr / (3 * msr * msr) - 1, np.nan)
        table["gamma"] = np.where((msd > 0) & (msr > 0), (s[:, S_D2R2] / n) / (msd * msr) - 1, np.nan)
    if distance is not None:
        table["com_struct"] = s[:, S_COMSTRUCT] / n
    table["mean_rotation"] = s[:, S_ROT] / n
    table["msr"] = msr
    table["rot1"] = s[:, S_COS] / n
    table["rot2"] = s[:, S_COS2] / n
    if overlap_k is not None:
        if overlap_k == 0:
            if strict and a:
                raise ZeroDivisionError("division by zero")  # dynamics.py:445 with N < 10
        else:
            table["overlap"] = s[:, S_OVERLAP] / overlap_k
    if n_sites is not None:
        table["struct"] = s[:, S_STRUCT] / (n * n_sites)
    return table


# ----------------------------------------------------------------------------------------------
class _CudaArray:
    """Minimal ``__cuda_array_interface__`` carrier: lets torch view memory owned by the native engine."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(v) for v in shape), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None,
        }


class _OriginView:
    """One origin of the native engine: torch views of its device state (no copies)."""

    def __init__(self, engine, key, info):
        self._engine = engine
        self.key = key
        self.timestep = int(info.timestep)
        self.updates = int(info.updates)
        self.max_timediff = int(info.max_timediff)
        self._info = info

    def _view(self, ptr, shape, typestr):
        if not ptr:
            return None
        eng = self._engine
        t = eng.torch.as_tensor(_CudaArray(ptr, shape, typestr), device=eng.device)
        t._sdb_owner = eng  # the engine owns the memory
        return t

    @property
    def delta(self):
        return self._view(self._info.d_delta, (3, self._engine.stride), "<f4")

    @property
    def flags(self):
        return self._view(self._info.d_flags, (self._engine.stride,), "|u1")

    @property
    def status(self):
        n_tr = len(self._engine.trackers)
        return self._view(self._info.d_status, (n_tr, self._engine.stride), "<i4") if n_tr else None

    @property
    def prev_planes(self):
        """(x, y, qw, qz) planes of the previous sample (``TrackedMotion.previous_*``)."""
        return self._view(self._info.d_prev, (4, self._engine.stride), "<f4")

    @property
    def prev_has_orientation(self):
        return bool(self._info.prev_has_orientation)

    @property
    def ptrs(self):
        return (self._info.d_delta or 0, self._info.d_flags or 0, self._info.d_status or 0)


class _Origins:
    """Mapping view of the engine's origins, keyed like ``process_file``'s ``keyframes`` dict."""

    def __init__(self, engine):
        self._engine = engine

    def __getitem__(self, key):
        eng = self._engine
        info = _lib.SdbOriginInfo()
        if key not in eng._ids or eng.lib.sdb_engine_origin(eng._handle, eng._ids[key], ctypes.byref(info)) != 0:
            raise KeyError(key)
        return _OriginView(eng, key, info)

    def __contains__(self, key):
        return key in self._engine._t0

    def __len__(self):
        return int(self._engine.lib.sdb_engine_num_origins(self._engine._handle))

    def keys(self):
        eng = self._engine
        count = len(self)
        buf = (ctypes.c_int64 * max(count, 1))()
        eng.lib.sdb_engine_keys(eng._handle, buf, count)
        return [eng._keys_by_id[int(buf[i])] for i in range(count)]

    def __iter__(self):
        return iter(self.keys())

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class StepResult:
    """Handle to the rows of one engine step; ``rows()`` / ``table()`` wait for the device and finalise."""

    def __init__(self, engine, keys, timestep, ticket, has_rows, has_isf):
        self.engine = engine
        self.keys = keys
        self._timestep = timestep
        self._ticket = ticket
        self._has_rows = has_rows
        self._has_isf = has_isf
        self._sums = None
        self._isf = None
        self._timediffs = None

    @property
    def timediffs(self):
        if self._timediffs is None:
            t0 = self.engine._t0
            self._timediffs = [self._timestep - t0[k] for k in self.keys]
        return self._timediffs

    def _materialize(self):
        """Copy the rows out of the engine's ring (called before the ring slot is reused, or on demand)."""
        if self._ticket is None:
            return
        ticket, self._ticket = self._ticket, None
        eng = self.engine
        n = len(self.keys)
        if not self._has_rows or n == 0:
            self._sums = np.zeros((n, SDB_NSUM))
            return
        sums = np.empty((n, SDB_NSUM), dtype=np.float64)
        isf = np.empty(n, dtype=np.float64) if self._has_isf else None
        _lib.check(
            eng.lib.sdb_engine_wait(eng._handle, ticket, sums.ctypes.data, None, isf.ctypes.data if isf is not None else None)
        )
        self._sums, self._isf = sums, isf

    def sums(self):
        self._materialize()
        return self._sums

    def rows(self, check=True, strict=True):
        """One dict of the 15 ``compute_all`` quantities per active origin, in ``keys`` order.
        ``strict``: raise ZeroDivisionError for N < 10 like ``mobile_overlap`` does in compute_all."""
        rows = self.engine.finalize_rows(self.sums(), self.timediffs, check=check, strict=strict)
        if self._isf is not None:
            for row, value in zip(rows, self._isf):
                row["scattering_function"] = float(value) / self.engine.n
        return rows

    def table(self, check=True, strict=True):
        """The same quantities as columns: ``{quantity: (A,) array}`` plus ``keyframe`` (vectorised over
        the active origins; this is what the ``dynamics`` table of ``process_file`` is assembled from)."""
        table = self.engine.finalize_table(self.sums(), self.timediffs, check=check, strict=strict)
        if self._isf is not None:
            table["scattering_function"] = np.asarray(self._isf, dtype=np.float64) / self.engine.n
        table["keyframe"] = list(self.keys)
        return table


class DynamicsEngine:
    """All time-origins of one GPU.  ``step`` = one frame of ``read/_read.py:128-168`` for every active origin."""

    def __init__(
        self,
        num_mols,
        box,
        wave_number=None,
        *,
        trackers="default",
        compute_sums=True,
        compute_overlap=True,
        compute_struct=True,
        struct_mode="reference",
        mol_sites=None,
        device=None,
        ring_depth=3,
        overlap_fraction=0.1,
        compute_isf=False,
        angular_resolution=360,
        chunk=None,
        dry=False,
    ):
        import os

        self.lib = _lib.load()
        self._dry = bool(dry)  # test hook: bookkeeping only, no CUDA call (CPU test-suite)
        if self._dry:
            self.torch, self.device = None, None
        else:
            self.torch = _lib.torch()
            self.device = _lib.require_cuda(device)
        self.n = int(num_mols)
        if self.n <= 0:
            raise RuntimeError("Position must contain values, has length of 0.")
        self.stride = (self.n + 3) // 4 * 4
        box = np.asarray(box, dtype=F32).reshape(-1)
        self.box = np.zeros(6, dtype=F32)
        self.box[: len(box)] = box
        self.wave_number = wave_number
        self.distance = None if wave_number is None else np.pi / (2 * wave_number)

        if isinstance(trackers, str) and trackers == "default":
            trackers = None if self.distance is None else default_tracker_definitions(self.distance)
        self.compute_sums = bool(compute_sums)
        self.compute_overlap = bool(compute_overlap) and self.compute_sums
        self.overlap_k = int(self.n * overlap_fraction)
        self.mol_sites = None if mol_sites is None else np.asarray(mol_sites, dtype=F32)[:, :2].copy()
        self.compute_struct = (
            bool(compute_struct) and self.compute_sums and self.distance is not None and self.mol_sites is not None
        )
        # intermediate scattering function (Dynamics.compute_isf, off by default like the reference's
        # --scattering-function): one extra kernel per frame, csrc/isf.cu
        self.compute_isf = bool(compute_isf) and self.compute_sums and wave_number is not None
        self.angular_resolution = int(angular_resolution)
        if struct_mode not in ("reference", "physical"):
            raise ValueError("struct_mode must be 'reference' or 'physical'")
        self.struct_mode = 0 if struct_mode == "reference" else 1
        self.trackers = build_trackers(trackers) if trackers else []
        self.ring_depth = max(2, int(ring_depth))

        cfg = _lib.SdbEngineConfig()
        cfg.n = self.n
        for i in range(6):
            cfg.box[i] = float(self.box[i])
        cfg.wave_number = float(wave_number) if wave_number is not None else 0.0
        cfg.com_cut_lt = float(_cut_sqrt(self.distance, False)) if self.distance is not None else -1.0
        self._fill_trackers(cfg.trackers)
        cfg.n_trackers = len(self.trackers)
        cfg.compute_sums = int(self.compute_sums)
        cfg.overlap_k = self.overlap_k if self.compute_overlap else 0
        if self.compute_struct:
            cfg.n_sites = len(self.mol_sites)
            for i, v in enumerate(self.mol_sites.reshape(-1)):
                cfg.sites_xy[i] = float(v)
        cfg.struct_mode = self.struct_mode
        cfg.compute_isf = int(self.compute_isf)
        cfg.angular_resolution = self.angular_resolution
        cfg.ring_depth = self.ring_depth
        # SDB_INCREMENTS=0: every segment recomputes the frame increments itself (no per-frame pass); the pass
        # costs about one phase 1 plus 76 B per molecule of traffic: it pays when several segments share a
        # previous sample and each has few origins to amortise its own phase 1 over (measured: 25 origins in 3
        # segments -11 % on dyn_step, 200 origins in 3 segments +3 %)
        cfg.inc_max_per_segment = int(os.environ.get("SDB_INCREMENTS_MAX_PER_SEGMENT", "40"))
        cfg.chunk = int(chunk or 0)
        flags = 0
        if self._dry:
            flags |= _lib.SDB_ENGINE_DRY
        if os.environ.get("SDB_OV_PREDICT", "1") == "0":  # sampled overlap brackets every frame
            flags |= _lib.SDB_ENGINE_NO_PREDICT
        if os.environ.get("SDB_INCREMENTS", "1") == "0":
            flags |= _lib.SDB_ENGINE_NO_INCREMENTS
        cfg.flags = flags
        self._handle = ctypes.c_void_p()
        if self._dry:
            _lib.check(self.lib.sdb_engine_create(ctypes.byref(cfg), ctypes.byref(self._handle)))
        else:
            with _lib.device_context(self.device):
                _lib.check(self.lib.sdb_engine_create(ctypes.byref(cfg), ctypes.byref(self._handle)))
        self.origins = _Origins(self)
        self._t0 = {}          # key -> timestep of the origin's first frame
        self._ids = {}         # key -> int64 id handed to the native engine
        self._keys_by_id = {}
        self._frame = _lib.SdbFrame()
        self._last_active = None
        self._last_ids = None
        self._next_ticket = 0
        self._outstanding = deque()  # (ticket, weakref to StepResult): materialised before their slot is reused
        self._keepalive = deque()    # host tensors whose upload may still be in flight

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            try:
                self.lib.sdb_engine_destroy(handle)
            except Exception:
                pass

    # ------------------------------------------------------------------------------------------
    def _fill_trackers(self, array):
        for i, spec in enumerate(self.trackers):
            cut = _cut_direct if spec.type == "rotation" else _cut_sqrt
            t = array[i]
            t.is_rotation = int(spec.type == "rotation")
            t.is_last = int(spec.last_passage)
            t.bit = spec.bit
            t.cut_gt = float(cut(spec.threshold, True))
            t.cut_lt = float(cut(spec.threshold, False))
            t.cut_irr = float(cut(spec.cutoff, True))

    def set_trackers(self, definitions):
        """Replace the tracker set; every origin's passage state starts afresh while its accumulated
        motion is kept (``Relaxations.set_mol_relax``, ``relaxations.py:108-133``)."""
        self.trackers = build_trackers(definitions) if definitions else []
        array = (_lib.SdbTracker * SDB_MAX_TRACKERS)()
        self._fill_trackers(array)
        _lib.check(self.lib.sdb_engine_set_trackers(self._handle, array, len(self.trackers)))

    def _flush_plan(self):
        """(kept for callers of round 1: the native engine has no deferred bookkeeping)"""

    def drop_origin(self, key):
        _lib.check(self.lib.sdb_engine_drop(self._handle, self._ids[key]))
        ident = self._ids.pop(key)
        self._keys_by_id.pop(ident, None)
        self._t0.pop(key, None)
        self._last_active = None

    def _key_ids(self, active, timestep):
        """int64 ids of ``active`` for the native engine (cached while the list does not change)."""
        if active == self._last_active:
            return self._last_ids
        ids = self._ids
        for key in active:
            if key not in ids:
                ident = int(key) if isinstance(key, (int, np.integer)) and not isinstance(key, bool) else None
                if ident is None or ident in self._keys_by_id or not -(2 ** 62) < ident < 2 ** 62:
                    ident = 2 ** 62 + len(ids)  # arbitrary hashable keys get ids outside the plain-integer range
                ids[key] = ident
                self._keys_by_id[ident] = key
                self._t0[key] = int(timestep)
        arr = (ctypes.c_int64 * max(len(active), 1))(*[ids[k] for k in active])
        self._last_active, self._last_ids = list(active), arr
        return arr

    def _forget_failed_keys(self, active):
        """After a failed step nothing of it stays: keys the native engine did not create lose their Python-side
        record, origins it created without ever stepping them are dropped."""
        info = _lib.SdbOriginInfo()
        for key in active:
            ident = self._ids.get(key)
            if ident is None:
                continue
            known = self.lib.sdb_engine_origin(self._handle, ident, ctypes.byref(info)) == 0
            if known and info.updates == 0:
                self.lib.sdb_engine_drop(self._handle, ident)
                known = False
            if not known:
                self._ids.pop(key, None)
                self._keys_by_id.pop(ident, None)
                self._t0.pop(key, None)
        self._last_active = None

    def _frame_pointers(self, position, orientation):
        """(pos pointer, quat pointer, SDB_MEM_* kind, objects to keep alive) of a frame in any of the accepted forms."""
        torch, n = self.torch, self.n
        if torch is not None and isinstance(position, torch.Tensor):
            pos, quat = position, orientation
            if tuple(pos.shape) != (n, 3) or (quat is not None and tuple(quat.shape) != (n, 4)):
                raise RuntimeError("frame tensors must have shapes (N,3) and (N,4)")
            if pos.dtype != torch.float32 or not pos.is_contiguous():
                pos = pos.to(torch.float32).contiguous()
            if quat is not None and (quat.dtype != torch.float32 or not quat.is_contiguous()):
                quat = quat.to(torch.float32).contiguous()
            if pos.is_cuda:
                if pos.data_ptr() % 16:
                    pos = pos.clone()
                if quat is not None and quat.data_ptr() % 16:
                    quat = quat.clone()
                return pos.data_ptr(), _lib.ptr(quat), _lib.SDB_MEM_DEVICE, (pos, quat)
            kind = _lib.SDB_MEM_PINNED if pos.is_pinned() and (quat is None or quat.is_pinned()) else _lib.SDB_MEM_HOST
            return pos.data_ptr(), _lib.ptr(quat), kind, (pos, quat)
        position = np.asarray(position)
        if position.shape != (n, 3):
            raise RuntimeError(f"position has shape {position.shape}, expected {(n, 3)}")
        if position.dtype != F32 or not position.flags.c_contiguous:
            position = np.ascontiguousarray(position, dtype=F32)
        quat_ptr = 0
        if orientation is not None:
            orientation = np.asarray(orientation)
            if orientation.shape != (n, 4):
                raise RuntimeError(f"orientation has shape {orientation.shape}, expected {(n, 4)}")
            if orientation.dtype != F32 or not orientation.flags.c_contiguous:
                orientation = np.ascontiguousarray(orientation, dtype=F32)
            quat_ptr = orientation.ctypes.data
        pos_ptr = position.ctypes.data
        if _lib._pinned_ranges and _lib.is_pinned(pos_ptr, position.nbytes) and (
            not quat_ptr or _lib.is_pinned(quat_ptr, orientation.nbytes)
        ):
            # views of a page-locked file mapping (gsdio.GSDTrajectory.pin): uploaded in place
            return pos_ptr, quat_ptr, _lib.SDB_MEM_PINNED, (position, orientation)
        # pageable host memory is copied into the engine's pinned ring inside the call: the (possibly converted)
        # arrays only have to outlive the call itself
        return pos_ptr, quat_ptr, _lib.SDB_MEM_HOST, (position, orientation)

    # ------------------------------------------------------------------------------------------
    def step(self, timestep, position, orientation, active, *, zero_step=False, want_sums=None, to_host=True, rows_out=0):
        """Advance every origin in ``active`` (iterable of hashable keys) with one frame.

        Unknown keys are created at this frame (``Dynamics.from_frame`` + first ``compute_all``,
        ``read/_read.py:135-152``).  Returns a :class:`StepResult` (asynchronous; rows in the order of
        ``active``).  ``to_host=False`` keeps the rows on the device: returns ``(keys, timediffs, rows)``
        with ``rows`` a device tensor view valid for ``ring_depth - 1`` further steps; ``rows_out`` (a raw
        device pointer, possibly peer memory) additionally receives the rows in ``active`` order.
        """
        active = active if isinstance(active, list) else list(active)
        n_active = len(active)
        if n_active == 0:
            return StepResult(self, [], int(timestep), None, False, False)
        timestep = int(timestep)
        do_sums = self.compute_sums if want_sums is None else (bool(want_sums) and self.compute_sums)
        if zero_step:
            for key in active:
                if key not in self._ids:
                    raise RuntimeError("zero_step on an origin that does not exist")
        f = self._frame
        flags = 0
        keep = None
        if zero_step:
            flags |= _lib.SDB_FRAME_ZERO_STEP
            f.pos = f.quat = None
            f.where = _lib.SDB_MEM_DEVICE
        else:  # (a frame that is refused must not leave records of the keys it would have created)
            pos_ptr, quat_ptr, kind, keep = self._frame_pointers(position, orientation)
            f.pos, f.quat, f.where = pos_ptr, quat_ptr or None, kind
        ids = self._key_ids(active, timestep)
        f.timestep = timestep
        f.keys = ctypes.addressof(ids)
        f.n_keys = n_active
        if not do_sums:
            flags |= _lib.SDB_FRAME_NO_SUMS
        if not to_host:
            flags |= _lib.SDB_FRAME_ROWS_DEVICE
        f.flags = flags
        f.d_rows_out = rows_out or None

        # results about to lose their ring slot are copied out first
        out = self._outstanding
        horizon = self._next_ticket - self.ring_depth
        while out and out[0][0] <= horizon:
            _, ref = out.popleft()
            res = ref()
            if res is not None:
                res._materialize()
        stream = 0 if self._dry else _lib.stream_ptr(self.device)
        ticket = int(self.lib.sdb_engine_step(self._handle, ctypes.byref(f), stream))
        if ticket < 0:
            message = self.lib.sdb_last_error().decode()
            self._forget_failed_keys(active)
            if "time difference" in message:
                raise OverflowError(message)
            raise _lib.SdbError(f"libsdb200 error {ticket}: {message}")
        self._next_ticket = ticket + 1
        if keep is not None and f.where == _lib.SDB_MEM_PINNED:
            # pinned host tensors are read by the copy stream after this call returns
            self._keepalive.append(keep)
            if len(self._keepalive) > self.ring_depth + 1:
                self._keepalive.popleft()
        if not to_host:
            timediffs = [timestep - self._t0[k] for k in active]
            rows = None
            if do_sums and not self._dry and not rows_out:  # (with rows_out the rows are already where they belong)
                d_rows = ctypes.c_void_p()
                perm = (ctypes.c_int32 * n_active)()
                _lib.check(self.lib.sdb_engine_device_rows(self._handle, ticket, ctypes.byref(d_rows), perm))
                rows = self.torch.as_tensor(_CudaArray(d_rows.value, (n_active, SDB_NSUM), "<f8"), device=self.device)
                order = list(perm)
                if order != list(range(n_active)):
                    rows = rows[self.torch.tensor(order, device=self.device)]
            return active, timediffs, rows
        result = StepResult(self, active, timestep, ticket, do_sums, do_sums and self.compute_isf)
        out.append((ticket, weakref.ref(result)))
        return result

    def run(self, frames, *, want_sums=True):
        """A batch of frames in ONE native call (``sdb_engine_run``): ``frames`` is a sequence of
        ``(timestep, position, orientation, active)`` with host arrays (numpy, memory-mapped or pinned) or
        device tensors.  Returns one :class:`StepResult`-like table per frame, already on the host.  This is
        the loop of ``read/_read.py:128-168`` with no Python between the frames."""
        frames = list(frames)
        count = len(frames)
        if count == 0:
            return []
        arr = (_lib.SdbFrame * count)()
        keep, key_arrays, actives = [], [], []
        total = 0
        for i, (timestep, position, orientation, active) in enumerate(frames):
            active = list(active)
            try:
                pos_ptr, quat_ptr, kind, alive = self._frame_pointers(position, orientation)
            except Exception:
                for seen in actives:  # nothing has reached the native engine yet
                    self._forget_failed_keys(seen)
                raise
            ids = self._key_ids(active, int(timestep))
            keep.append((alive, position, orientation))
            key_arrays.append(ids)
            actives.append(active)
            f = arr[i]
            f.timestep = int(timestep)
            f.pos, f.quat, f.where = pos_ptr, quat_ptr or None, kind
            f.flags = 0 if (want_sums and self.compute_sums) else _lib.SDB_FRAME_NO_SUMS
            f.keys, f.n_keys = ctypes.addressof(ids), len(active)
            f.d_rows_out = None
            total += len(active)
        for _, ref in self._outstanding:
            res = ref()
            if res is not None:
                res._materialize()
        self._outstanding.clear()
        sums = np.zeros((max(total, 1), SDB_NSUM), dtype=np.float64)
        isf = np.zeros(max(total, 1), dtype=np.float64) if self.compute_isf else None
        stream = 0 if self._dry else _lib.stream_ptr(self.device)
        rc = self.lib.sdb_engine_run(self._handle, arr, count, sums.ctypes.data, None, isf.ctypes.data if isf is not None else None, stream)
        self._next_ticket += count
        if rc != 0:
            message = self.lib.sdb_last_error().decode()
            for active in actives:
                self._forget_failed_keys(active)
            if "time difference" in message:
                raise OverflowError(message)
            raise _lib.SdbError(f"libsdb200 error {rc}: {message}")
        del keep, key_arrays
        results, off = [], 0
        for (timestep, _, _, _), active in zip(frames, actives):
            res = StepResult(self, active, int(timestep), None, True, False)
            res._sums = sums[off : off + len(active)]
            res._isf = isf[off : off + len(active)] if (isf is not None and want_sums) else None
            results.append(res)
            off += len(active)
        return results

    # ------------------------------------------------------------------------------------------
    def finalize_rows(self, sums, timediffs, check=True, strict=True):
        """Host-side closed forms of ``Dynamics.compute_all`` (``dynamics.py:183-303,328-394``)."""
        rows = []
        for s, td in zip(sums, timediffs):
            if check and s[S_BADQUAT] > 0:
                # rowan.to_euler raises ValueError for non-unit quaternions (process_file skips the row)
                raise ValueError("Arguments must be unit quaternions (planar, rotation about z)")
            rows.append(
                finalize_row(
                    s,
                    td,
                    self.n,
                    distance=self.distance,
                    overlap_k=self.overlap_k if self.compute_overlap else None,
                    n_sites=len(self.mol_sites) if self.compute_struct else None,
                    strict=strict,
                )
            )
        return rows

    def finalize_table(self, sums, timediffs, check=True, strict=True):
        """Column form of :meth:`finalize_rows` (one numpy expression per quantity)."""
        sums = np.asarray(sums).reshape(-1, SDB_NSUM)
        if check and sums.shape[0] and (sums[:, S_BADQUAT] > 0).any():
            raise ValueError("Arguments must be unit quaternions (planar, rotation about z)")
        return finalize_table(
            sums,
            timediffs,
            self.n,
            distance=self.distance,
            overlap_k=self.overlap_k if self.compute_overlap else None,
            n_sites=len(self.mol_sites) if self.compute_struct else None,
            strict=strict,
        )

    # ------------------------------------------------------------------------------------------
    def stats(self):
        st = _lib.SdbEngineStats()
        _lib.check(self.lib.sdb_engine_stats(self._handle, ctypes.byref(st)))
        return st

    h2d_bytes = property(lambda self: int(self.stats().h2d_bytes))
    d2h_bytes = property(lambda self: int(self.stats().d2h_bytes))
    frames_seen = property(lambda self: int(self.stats().frames))
    updates = property(lambda self: int(self.stats().updates))
    _inc_ws = property(lambda self: self.stats().d_inc_ws or None)

    def last_descriptors(self):
        """(origin records, segment records) of the last step as numpy struct arrays (test hook)."""
        n_org, n_seg, seg_off = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        size = self.lib.sdb_engine_last_descriptors(self._handle, None, 0, ctypes.byref(n_org), ctypes.byref(n_seg), ctypes.byref(seg_off))
        buf = np.zeros(max(seg_off.value + 24 * n_seg.value, 1), dtype=np.uint8)
        self.lib.sdb_engine_last_descriptors(self._handle, buf.ctypes.data, buf.nbytes, None, None, None)
        del size
        origins = buf[: 32 * n_org.value].view(ORIGIN_DTYPE).copy()
        segments = buf[seg_off.value : seg_off.value + 24 * n_seg.value].view(SEGMENT_DTYPE).copy()
        return origins, segments

    def delta_arrays(self, key):
        """(delta_translation (N,3), delta_rotation (N,3)) of one origin as host arrays."""
        planes = self.origins[key].delta[:, : self.n].cpu().numpy()
        trans = np.zeros((self.n, 3), dtype=F32)
        rot = np.zeros((self.n, 3), dtype=F32)
        trans[:, 0], trans[:, 1], rot[:, 0] = planes[0], planes[1], planes[2]
        return trans, rot

    def norms(self, key):
        """(|dr|, |dtheta|) per molecule, computed on the device."""
        torch = self.torch
        ptr = self.origins[key].ptrs[0]
        with _lib.device_context(self.device):
            out = torch.empty((2, self.n), dtype=torch.float32, device=self.device)
            _lib.check(
                self.lib.sdb_dyn_norms(ptr, self.n, self.stride, out[0].data_ptr(), out[1].data_ptr(), _lib.stream_ptr(self.device))
            )
            host = out.cpu().numpy()
        return host[0], host[1]

    def tracker_state(self, key):
        """Raw passage times (uint32 (T, N)) and the flag byte per molecule of one origin."""
        org = self.origins[key]
        status = org.status[:, : self.n].cpu().numpy().view(np.uint32)
        flags = org.flags[: self.n].cpu().numpy()
        return status, flags

    def summary(self, key):
        """``{tracker name: get_status()}`` like ``Relaxations.summary`` (``relaxations.py:181-184``):
        first-passage trackers give the raw word, last-passage ones mask everything that is not in
        the irreversible state (``relaxations.py:291-294``)."""
        status, flags = self.tracker_state(key)
        out = {}
        for i, spec in enumerate(self.trackers):
            col = status[i].astype(np.int64)
            if spec.last_passage:
                state = (flags >> spec.bit) & 3
                col[state != 2] = MAX_STATUS
            out[spec.name] = col
        return out

    def overlap_stats(self, n_active):
        """(origins that needed the exact kernel, candidates per origin) of the last step (diagnostics)."""
        st = self.stats()
        out = np.zeros(1 + n_active, dtype=np.uint32)
        with _lib.device_context(self.device):
            _lib.check(
                self.lib.sdb_overlap_stats(st.d_ov_ws, st.capacity, self.n, n_active, out.ctypes.data, _lib.stream_ptr(self.device))
            )
        return int(out[0]), out[1:]

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

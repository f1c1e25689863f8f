"""Batched multi-origin dynamics engine: the B200 replacement for the reference's hot loop.

The reference keeps one ``(Dynamics, Relaxations)`` pair of Python objects per time-origin
("keyframe") and, for every frame, calls ``compute_all`` + ``Relaxations.add`` on each active one
(``read/_read.py:128-168``).  Here all origins of a rank live in HBM as planar arrays and one call to
:meth:`DynamicsEngine.step` advances every active origin with the fused CUDA kernel
(``csrc/dyn_step.cu``) plus the ``struct`` and ``overlap`` kernels.

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

import math
from collections import OrderedDict

import numpy as np

from . import _lib
from ._lib import (
    HISTORY_WORDS,
    ORIGIN_DTYPE,
    ORIGIN_HISTORY,
    ORIGIN_LITERAL,
    ORIGIN_NEW,
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
    SEG_NO_ROTATION,
    SEG_ZERO_STEP,
    SEGMENT_DTYPE,
    SdbDynParams,
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
class _PoolFrame:
    __slots__ = ("planes", "refs", "has_orientation", "ptr")

    def __init__(self, planes):
        self.planes = planes
        self.ptr = planes.data_ptr()
        self.refs = 0
        self.has_orientation = True


class _Origin:
    __slots__ = ("key", "timestep", "delta", "flags", "status", "prev", "max_timediff", "updates", "ptrs")

    def __init__(self, key, timestep):
        self.key = key
        self.timestep = int(timestep)
        self.delta = self.flags = self.status = None
        self.prev = None
        self.max_timediff = -1
        self.updates = 0
        self.ptrs = (0, 0, 0)  # raw device pointers of delta / flags / status (cached: data_ptr() is slow)


class _Ring:
    """A small ring of pinned host buffers, each guarded by a CUDA event."""

    def __init__(self, torch, depth, make):
        self.bufs = [make() for _ in range(depth)]
        self.events = [None] * depth
        self.torch = torch
        self.i = 0

    def acquire(self):
        self.i = (self.i + 1) % len(self.bufs)
        ev = self.events[self.i]
        if ev is not None:
            ev.synchronize()
        return self.i, self.bufs[self.i]

    def release(self, i, stream=None):
        ev = self.events[i]
        if ev is None:
            ev = self.events[i] = self.torch.cuda.Event()
        if stream is None:
            ev.record()
        else:
            ev.record(stream)  # with an explicit stream torch skips its (slow) current-stream lookup
        return ev


class StepResult:
    """Handle to the sums of one engine step; ``rows()`` waits for the device and finalises."""

    def __init__(self, engine, keys, timediffs, sums_host, event, isf_host=None):
        self.engine = engine
        self.keys = keys
        self.timediffs = timediffs
        self._sums = sums_host
        self._event = event
        self._isf = isf_host  # per-origin sums of the angular mean of cos(k . dr), or None

    def sums(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
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
    ):
        self.torch = torch = _lib.torch()
        self.device = _lib.require_cuda(device)
        self.lib = _lib.load()
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

        p = SdbDynParams()
        p.n, p.stride = self.n, self.stride
        p.lx, p.ly, p.xy = float(self.box[0]), float(self.box[1]), float(self.box[3])
        p.com_cut_lt = float(_cut_sqrt(self.distance, False)) if self.distance is not None else -1.0
        p.do_sums = int(self.compute_sums)
        self.params = p
        self.origins = OrderedDict()
        self.set_trackers(trackers)

        self._free_frames = []
        self._n_parts = int(self.lib.sdb_dyn_num_parts(self.n))
        self._ring_depth = ring_depth
        self._cap = 0  # capacity (origins) of the per-step scratch
        self._frame_ring = None
        self._dev_frames = None
        self._frame_i = 0
        self._staged_slot = None
        self.h2d_bytes = self.d2h_bytes = 0
        self._plan = None
        # SDB_OV_PREDICT=0 disables the history-predicted overlap brackets (sampled brackets every frame)
        import os

        self._origin_flags = ORIGIN_HISTORY if os.environ.get("SDB_OV_PREDICT", "1") != "0" else 0
        # SDB_INCREMENTS=0: every segment recomputes the frame increments itself (no per-frame pass)
        self._inc_ok = os.environ.get("SDB_INCREMENTS", "1") != "0"
        # the per-frame pass costs about one phase 1 plus 76 B per molecule of traffic: it pays when
        # several segments share a previous sample and each has few origins to amortise its own phase 1
        # over (measured: 25 origins in 3 segments -11 % on dyn_step, 200 origins in 3 segments +3 %)
        self._inc_max_per_segment = int(os.environ.get("SDB_INCREMENTS_MAX_PER_SEGMENT", "40"))
        self._inc_ws = None
        self._sites_c = None if self.mol_sites is None else _lib.float_array(self.mol_sites.reshape(-1))
        self.frames_seen = 0
        self.updates = 0  # molecule x origin updates performed

    # ------------------------------------------------------------------------------------------
    def _ensure_capacity(self, n_active):
        if n_active <= self._cap:
            return
        torch = self.torch
        cap = max(16, 1 << (n_active - 1).bit_length())
        dev = self.device
        self._cap = cap
        n_desc = cap * 32 + (cap + 8) * 24  # origins (32 B) + segments (24 B), generous
        self._desc_ring = _Ring(torch, self._ring_depth, lambda: torch.empty(n_desc, dtype=torch.uint8).pin_memory())
        self._desc_dev = [torch.empty(n_desc, dtype=torch.uint8, device=dev) for _ in range(self._ring_depth)]
        self._sums_ring = _Ring(
            torch, self._ring_depth, lambda: torch.zeros((cap, SDB_NSUM), dtype=torch.float64).pin_memory()
        )
        self._sums_dev = [torch.zeros((cap, SDB_NSUM), dtype=torch.float64, device=dev) for _ in range(self._ring_depth)]
        self._partials = torch.empty((cap, self._n_parts, SDB_NSUM), dtype=torch.float64, device=dev)
        self._ov_ws_ptr = 0
        if self.compute_overlap and self.overlap_k > 0:
            nbytes = int(self.lib.sdb_overlap_ws_bytes(cap, self.n))
            self._ov_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._ov_ws_ptr = self._ov_ws.data_ptr()
        if self.compute_isf:
            self._isf_dev = [torch.zeros(cap, dtype=torch.float64, device=dev) for _ in range(self._ring_depth)]
            self._isf_host = [torch.zeros(cap, dtype=torch.float64).pin_memory() for _ in range(self._ring_depth)]
        # raw pointers / numpy views used every frame (tensor.data_ptr() and .numpy() are slow)
        self._desc_host = [t.numpy() for t in self._desc_ring.bufs]
        self._desc_host_ptr = [t.data_ptr() for t in self._desc_ring.bufs]
        self._desc_dev_ptr = [t.data_ptr() for t in self._desc_dev]
        self._sums_host_ptr = [t.data_ptr() for t in self._sums_ring.bufs]
        self._sums_dev_ptr = [t.data_ptr() for t in self._sums_dev]
        self._partials_ptr = self._partials.data_ptr()

    def _inc_args(self, inc_sets):
        """ctypes arguments (h_prev, h_flags, n_sets, d_ws) of the per-frame increments pass."""
        import ctypes

        if not inc_sets:
            return None, None, 0, 0
        if self._inc_ws is None:
            nbytes = int(self.lib.sdb_dyn_increments_bytes(self.n, self.stride, _lib.SDB_MAX_INC_SETS))
            self._inc_ws = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)
        k = len(inc_sets)
        prev = (ctypes.c_void_p * _lib.SDB_MAX_INC_SETS)(*[p for p, _ in inc_sets])
        flags = (ctypes.c_uint32 * _lib.SDB_MAX_INC_SETS)(*[f for _, f in inc_sets])
        # the workspace is laid out for exactly k sets by both calls
        return prev, flags, k, self._inc_ws.data_ptr()

    def _stage_frame(self, position, orientation):
        """Return device tensors (pos (N,3) f32, quat (N,4) f32 or None).

        Device tensors are used in place.  Host input (numpy arrays or CPU tensors) is copied to a
        ring of device frame buffers on a dedicated copy stream, so the upload of frame t+1 overlaps
        the kernels of frame t; numpy arrays first go through a ring of pinned staging buffers, pinned
        CPU tensors are uploaded directly."""
        torch = self.torch
        if isinstance(position, torch.Tensor) and position.is_cuda:
            pos = position
            quat = orientation
            if pos.dtype != torch.float32 or not pos.is_contiguous():
                pos = pos.to(torch.float32).contiguous()
            if quat is not None and (quat.dtype != torch.float32 or not quat.is_contiguous()):
                quat = quat.to(torch.float32).contiguous()
            if pos.data_ptr() % 16:
                pos = pos.clone()
            if quat is not None and quat.data_ptr() % 16:
                quat = quat.clone()
            self.h2d_bytes = 0
            return pos, quat
        n = self.n
        qoff = (3 * n + 3) // 4 * 4  # quaternions start 16-byte aligned (float4 loads)
        if self._dev_frames is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._dev_frames = [
                torch.empty(qoff + 4 * n, dtype=torch.float32, device=self.device) for _ in range(self._ring_depth)
            ]
            self._dev_consumed = [None] * self._ring_depth
        if isinstance(position, torch.Tensor):
            if tuple(position.shape) != (n, 3) or (orientation is not None and tuple(orientation.shape) != (n, 4)):
                raise RuntimeError("frame tensors must have shapes (N,3) and (N,4)")
            src_pos, src_quat, ring_i = position, orientation, None
        else:
            position = np.asarray(position)
            if position.shape != (n, 3):
                raise RuntimeError(f"position has shape {position.shape}, expected {(n, 3)}")
            if orientation is not None:
                orientation = np.asarray(orientation)
                if orientation.shape != (n, 4):
                    raise RuntimeError(f"orientation has shape {orientation.shape}, expected {(n, 4)}")
            if self._frame_ring is None:
                self._frame_ring = _Ring(
                    torch, self._ring_depth, lambda: torch.zeros(qoff + 4 * n, dtype=torch.float32).pin_memory()
                )
            ring_i, pinned = self._frame_ring.acquire()
            host = pinned.numpy()
            np.copyto(host[: 3 * n].reshape(n, 3), position, casting="same_kind")
            src_pos = pinned[: 3 * n].view(n, 3)
            src_quat = None
            if orientation is not None:
                np.copyto(host[qoff : qoff + 4 * n].reshape(n, 4), orientation, casting="same_kind")
                src_quat = pinned[qoff : qoff + 4 * n].view(n, 4)
        self._frame_i = (self._frame_i + 1) % self._ring_depth
        dev = self._dev_frames[self._frame_i]
        pos = dev[: 3 * n].view(n, 3)
        quat = dev[qoff : qoff + 4 * n].view(n, 4) if src_quat is not None else None
        main = self._main
        with torch.cuda.stream(self._copy_stream):
            consumed = self._dev_consumed[self._frame_i]
            if consumed is not None:
                self._copy_stream.wait_event(consumed)  # kernels that read this buffer are done
            pos.copy_(src_pos, non_blocking=True)
            if quat is not None:
                quat.copy_(src_quat, non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(self._copy_stream)
        if ring_i is not None:
            self._frame_ring.events[ring_i] = uploaded
        main.wait_event(uploaded)
        self._staged_slot = self._frame_i
        self.h2d_bytes = 4 * (7 * n if src_quat is not None else 3 * n)
        return pos, quat

    def _new_pool_frame(self):
        if self._free_frames:
            return self._free_frames.pop()
        return _PoolFrame(self.torch.zeros((4, self.stride), dtype=self.torch.float32, device=self.device))

    def _release(self, frame):
        if frame is None:
            return
        frame.refs -= 1
        if frame.refs == 0:
            self._free_frames.append(frame)

    def _alloc_origin(self, org):
        torch, dev = self.torch, self.device
        # three planes followed by the threshold history of the overlap selection (SDB_ORIGIN_HISTORY)
        buf = torch.empty(3 * self.stride + HISTORY_WORDS, dtype=torch.float32, device=dev)
        buf[3 * self.stride :].zero_()
        org.delta = buf[: 3 * self.stride].view(3, self.stride)
        if self.trackers:
            org.flags = torch.empty(self.stride, dtype=torch.uint8, device=dev)
            org.status = torch.empty((len(self.trackers), self.stride), dtype=torch.int32, device=dev)
        org.ptrs = (org.delta.data_ptr(), _lib.ptr(org.flags), _lib.ptr(org.status))

    def set_trackers(self, definitions):
        """Replace the tracker set; every origin's passage state starts afresh while its accumulated
        motion is kept (``Relaxations.set_mol_relax``, ``relaxations.py:108-133``)."""
        # a cached steady-state plan holds the per-origin records (previous sample, update count,
        # largest time difference) of the frames stepped since it was made: bring them up to date
        # before dropping it, or the next step would group origins by a released previous sample
        if getattr(self, "_plan", None) is not None:
            self._flush_plan()
        self.trackers = build_trackers(definitions) if definitions else []
        p = self.params
        p.n_trackers = len(self.trackers)
        for i, spec in enumerate(self.trackers):
            cut = _cut_direct if spec.type == "rotation" else _cut_sqrt
            t = p.trackers[i]
            t.is_rotation = int(spec.type == "rotation")
            t.is_last = int(spec.last_passage)
            t.bit = spec.bit
            t.cut_gt = float(cut(spec.threshold, True))
            t.cut_lt = float(cut(spec.threshold, False))
            t.cut_irr = float(cut(spec.cutoff, True))
        torch, dev = self.torch, self.device
        for org in self.origins.values():
            if self.trackers:
                org.flags = torch.zeros(self.stride, dtype=torch.uint8, device=dev)
                org.status = torch.full((len(self.trackers), self.stride), -1, dtype=torch.int32, device=dev)
            else:
                org.flags = org.status = None
            org.ptrs = (org.delta.data_ptr(), _lib.ptr(org.flags), _lib.ptr(org.status))
        self._plan = None

    def _release_many(self, frame, count):
        if frame is None:
            return
        frame.refs -= count
        if frame.refs == 0:
            self._free_frames.append(frame)

    def _flush_plan(self):
        """Bring the per-origin records up to date after a run of steady-state steps."""
        plan, self._plan = self._plan, None
        if plan is None or plan["steps"] == 0:
            return
        last = plan["timestep"]
        for org in plan["ordered"]:
            org.prev = plan["prev"]
            org.updates += plan["steps"]
            org.max_timediff = max(org.max_timediff, last - org.timestep)

    def drop_origin(self, key):
        self._flush_plan()
        org = self.origins.pop(key)
        self._release(org.prev)

    # ------------------------------------------------------------------------------------------
    def step(self, timestep, position, orientation, active, *, zero_step=False, want_sums=None, chunk=None, to_host=True):
        """Advance every origin in ``active`` (iterable of hashable keys) with one frame.

        Unknown keys are created at this frame (``Dynamics.from_frame`` + first ``compute_all``,
        ``read/_read.py:135-152``).  Returns a :class:`StepResult` (asynchronous).
        """
        torch, lib = self.torch, self.lib
        active = list(active)
        n_active = len(active)
        if n_active == 0:
            return StepResult(self, [], [], np.zeros((0, SDB_NSUM)), None)
        self._ensure_capacity(n_active)
        do_sums = self.compute_sums if want_sums is None else bool(want_sums)

        with _lib.device_context(self.device):
            self._staged_slot = None
            main = self._main = _lib.current_stream(self.device)
            if zero_step:
                pos = quat = None
            else:
                pos, quat = self._stage_frame(position, orientation)

            # ---- descriptors: origins grouped into segments by previous sample ------------------
            plan = self._plan
            if (
                plan is not None
                and not zero_step
                and plan["active"] == active
                and plan["has_quat"] == (quat is not None)
                and int(timestep) > plan["timestep"]  # time moves forward: no origin needs ORIGIN_LITERAL
                and plan["monotone"]
                and chunk is None
            ):
                # steady state of the dense schedule: same origins, all sharing the previous frame --
                # only the time differences and the previous-sample pointer change
                ordered = plan["ordered"]
                odesc = plan["odesc"]
                odesc["timediff"] = int(timestep) - plan["t0"]
                odesc["flags"] = self._origin_flags
                timediffs = odesc["timediff"].tolist()
                if timediffs and max(timediffs) >= MAX_STATUS:
                    raise OverflowError("time difference outside [0, 2**32-1): passage times are 32-bit on the device")
                sdesc = plan["sdesc"]
                sdesc["d_prev"] = plan["prev"].ptr
                seg_count_max = plan["seg_count_max"]
                n_segments = len(sdesc)
                fast = True
                # every segment shares the previous frame: its increments are computed once per frame
                inc_sets = []
                if n_segments > 1 and self._inc_ok and n_active <= self._inc_max_per_segment * n_segments:
                    inc_sets = [(plan["prev"].ptr, int(sdesc["flags"][0]))]
                    sdesc["reserved"] = 1
                else:
                    sdesc["reserved"] = 0
            else:
                fast = False
                self._flush_plan()
                groups = OrderedDict()
                fresh = []
                for key in active:
                    org = self.origins.get(key)
                    if org is None:
                        org = _Origin(key, timestep)
                        self._alloc_origin(org)
                        self.origins[key] = org
                        fresh.append(org)
                    else:
                        groups.setdefault(id(org.prev), (org.prev, []))[1].append(org)
                if zero_step and fresh:
                    raise RuntimeError("zero_step on an origin that does not exist")

                target_blocks = 148 * 16
                blocks_per_segment = (self._n_parts + 3) // 4
                seg_rows = []  # (previous planes pointer, flags, first, count)
                ordered = []
                for prev, members in groups.values():
                    flags = 0
                    if zero_step:
                        flags |= SEG_ZERO_STEP
                    elif quat is None or prev is None or not prev.has_orientation:
                        flags |= SEG_NO_ROTATION
                    size = chunk or max(8, math.ceil(len(members) * blocks_per_segment / target_blocks))
                    size = math.ceil(len(members) / math.ceil(len(members) / size))  # equal shares
                    for first in range(0, len(members), size):
                        part = members[first : first + size]
                        seg_rows.append((0 if prev is None else prev.ptr, flags, len(ordered), len(part)))
                        ordered.extend(part)
                if fresh:
                    flags = SEG_NO_ROTATION if quat is None else 0
                    seg_rows.append((0, flags, len(ordered), len(fresh)))
                    ordered.extend(fresh)

                odesc = np.zeros(n_active, dtype=ORIGIN_DTYPE)
                t0 = np.fromiter((org.timestep for org in ordered), dtype=np.int64, count=n_active)
                td = int(timestep) - t0
                if td.min() < 0 or td.max() >= MAX_STATUS:
                    raise OverflowError(
                        f"time difference outside [0, 2**32-1): passage times are 32-bit on the device"
                    )
                ptrs = np.array([org.ptrs for org in ordered], dtype=np.uint64)
                odesc["d_delta"], odesc["d_flags"], odesc["d_status"] = ptrs[:, 0], ptrs[:, 1], ptrs[:, 2]
                odesc["timediff"] = td
                odesc["flags"] = [
                    self._origin_flags
                    | (ORIGIN_NEW if org.updates == 0 else (ORIGIN_LITERAL if t < org.max_timediff else 0))
                    for org, t in zip(ordered, td.tolist())
                ]
                timediffs = td.tolist()
                # previous samples shared by several segments: their increments are computed once per frame
                inc_sets, set_of = [], {}
                if self._inc_ok and not zero_step:
                    counts = {}
                    for (p, fl, f, c) in seg_rows:
                        if p:
                            counts[(p, fl)] = counts.get((p, fl), 0) + 1
                    for key, cnt in counts.items():
                        members = sum(c for (p, fl, f, c) in seg_rows if (p, fl) == key)
                        if cnt > 1 and members <= self._inc_max_per_segment * cnt and len(inc_sets) < _lib.SDB_MAX_INC_SETS:
                            set_of[key] = len(inc_sets) + 1
                            inc_sets.append(key)
                sdesc = np.array([(p, f, c, fl, set_of.get((p, fl), 0)) for (p, fl, f, c) in seg_rows], dtype=SEGMENT_DTYPE)
                seg_count_max = max(row[3] for row in seg_rows)
                n_segments = len(seg_rows)

            ri, pinned = self._desc_ring.acquire()
            ob, sb = odesc.nbytes, sdesc.nbytes
            seg_off = (ob + 15) // 16 * 16
            host = self._desc_host[ri]
            host[:ob] = odesc.view(np.uint8)
            host[seg_off : seg_off + sb] = sdesc.view(np.uint8)

            # ---- the frame that becomes everybody's previous sample ---------------------------
            new_prev = None
            if not zero_step:
                new_prev = self._new_pool_frame()
                new_prev.has_orientation = quat is not None

            si, sums_pinned = self._sums_ring.acquire()
            sums_dev = self._sums_dev[si]
            params = self.params
            params.do_sums = int(do_sums)
            use_ov = do_sums and self.compute_overlap and self.overlap_k > 0
            want_host = do_sums and to_host
            # steady state with three or more preceding overlap selections at increasing times: every
            # origin's bracket is predicted from its threshold history, the sample kernels are skipped
            step_flags = 0
            if fast:
                if use_ov and self._origin_flags and plan["ov_steps"] >= 3 and int(timestep) > plan["timestep"]:
                    step_flags |= _lib.SDB_STEP_PREDICTED
                plan["ov_steps"] = plan["ov_steps"] + 1 if (use_ov and int(timestep) > plan["timestep"]) else 0
            # one native call enqueues the whole frame: descriptor upload, overlap sampling/brackets, the
            # fused dynamics kernel, struct, overlap resolve, download of the sums (csrc/frame_step.cu)
            _lib.check(
                lib.sdb_frame_step(
                    params,
                    _lib.ptr(pos),
                    _lib.ptr(quat),
                    0 if new_prev is None else new_prev.ptr,
                    self._desc_host_ptr[ri],
                    self._desc_dev_ptr[ri],
                    seg_off + sb,
                    seg_off,
                    n_active,
                    n_segments,
                    seg_count_max,
                    self._partials_ptr,
                    self._sums_dev_ptr[si],
                    self._sums_host_ptr[si] if want_host else 0,
                    self._ov_ws_ptr if use_ov else 0,
                    self._cap,
                    self.overlap_k if use_ov else 0,
                    self._sites_c if (do_sums and self.compute_struct) else None,
                    len(self.mol_sites) if (do_sums and self.compute_struct) else 0,
                    float(self.distance) if self.distance is not None else 0.0,
                    self.struct_mode,
                    *self._inc_args(inc_sets),
                    step_flags,
                    main.cuda_stream,
                )
            )
            self._desc_ring.release(ri, main)
            isf_host = None
            if do_sums and self.compute_isf:
                _lib.check(
                    lib.sdb_isf_origins(
                        self._desc_dev_ptr[ri], n_active, self.n, self.stride, float(self.wave_number),
                        self.angular_resolution, self._isf_dev[si].data_ptr(), _lib.stream_ptr(self.device),
                    )
                )
                if want_host:
                    self._isf_host[si][:n_active].copy_(self._isf_dev[si][:n_active], non_blocking=True)
                    isf_host = self._isf_host[si].numpy()[:n_active]
            if self._staged_slot is not None:
                consumed = torch.cuda.Event()
                consumed.record(main)
                self._dev_consumed[self._staged_slot] = consumed
            event = None
            if want_host:
                event = self._sums_ring.release(si, main)
                self.d2h_bytes = n_active * SDB_NSUM * 8

            # ---- bookkeeping -------------------------------------------------------------------
            if not zero_step:
                if fast:
                    # the per-origin records are brought up to date lazily (_flush_plan)
                    plan["steps"] += 1
                    plan["timestep"] = int(timestep)
                    new_prev.refs += n_active
                    self._release_many(plan["prev"], n_active)
                    plan["prev"] = new_prev
                else:
                    for org, td in zip(ordered, timediffs):
                        self._release(org.prev)
                        org.prev = new_prev
                        org.updates += 1
                        if td > org.max_timediff:
                            org.max_timediff = td
                    new_prev.refs += n_active
                    # a plan for the next frame: valid if the same origins come again
                    if len({id(o) for o in ordered}) == n_active:
                        plan_odesc = odesc.copy()
                        plan_sdesc_rows = []
                        target_blocks = 148 * 16
                        blocks_per_segment = (self._n_parts + 3) // 4
                        size = max(8, math.ceil(n_active * blocks_per_segment / target_blocks))
                        size = math.ceil(n_active / math.ceil(n_active / size))  # equal shares
                        flags = 0 if quat is not None else SEG_NO_ROTATION
                        for first in range(0, n_active, size):
                            plan_sdesc_rows.append((0, first, min(size, n_active - first), flags, 0))
                        by_key = {org.key: org for org in ordered}
                        plan_order = [by_key[k] for k in active]
                        order_ptrs = np.array([org.ptrs for org in plan_order], dtype=np.uint64)
                        plan_odesc["d_delta"], plan_odesc["d_flags"], plan_odesc["d_status"] = (
                            order_ptrs[:, 0], order_ptrs[:, 1], order_ptrs[:, 2],
                        )
                        self._plan = {
                            "active": list(active),
                            "ordered": plan_order,
                            "odesc": plan_odesc,
                            "t0": np.fromiter((org.timestep for org in plan_order), dtype=np.int64, count=n_active),
                            "sdesc": np.array(plan_sdesc_rows, dtype=SEGMENT_DTYPE),
                            "seg_count_max": max(r[2] for r in plan_sdesc_rows),
                            "prev": new_prev,
                            "has_quat": quat is not None,
                            "timestep": int(timestep),
                            # every origin's largest time difference so far is the current one: from here
                            # on a later timestep is later for every origin (relaxations.py:215-218)
                            "monotone": all(t >= org.max_timediff for org, t in zip(ordered, timediffs)),
                            "steps": 0,
                            "ov_steps": 0,
                        }
                self.frames_seen += 1
                self.updates += n_active * self.n

        keys = [org.key for org in ordered]
        if not to_host:
            # device-resident result (multi-GPU gather path): valid until ring_depth further steps
            return keys, timediffs, (sums_dev[:n_active] if do_sums else None)
        sums_host = sums_pinned.numpy()[:n_active] if do_sums else None
        return StepResult(self, keys, timediffs, sums_host, event, isf_host)

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
    def delta_arrays(self, key):
        """(delta_translation (N,3), delta_rotation (N,3)) of one origin as host arrays."""
        org = self.origins[key]
        planes = org.delta[:, : self.n].cpu().numpy()
        trans = np.zeros((self.n, 3), dtype=F32)
        rot = np.zeros((self.n, 3), dtype=F32)
        trans[:, 0], trans[:, 1], rot[:, 0] = planes[0], planes[1], planes[2]
        return trans, rot

    def norms(self, key):
        """(|dr|, |dtheta|) per molecule, computed on the device."""
        torch = self.torch
        org = self.origins[key]
        with torch.cuda.device(self.device):
            out = torch.empty((2, self.n), dtype=torch.float32, device=self.device)
            _lib.check(
                self.lib.sdb_dyn_norms(
                    org.delta.data_ptr(), self.n, self.stride, out[0].data_ptr(), out[1].data_ptr(),
                    _lib.stream_ptr(self.device),
                )
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
        out = np.zeros(1 + n_active, dtype=np.uint32)
        with self.torch.cuda.device(self.device):
            _lib.check(
                self.lib.sdb_overlap_stats(
                    self._ov_ws.data_ptr(), self._cap, self.n, n_active, out.ctypes.data, _lib.stream_ptr(self.device)
                )
            )
        return int(out[0]), out[1:]

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

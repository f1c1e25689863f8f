"""Device-backed ``TrackedMotion`` (reference ``dynamics/_util.py:89-156``) and helpers.

Two backends hold the accumulated motion in HBM:

* planar  -- 2D box and quaternions about z (all of the reference's data): the fused multi-origin
  engine of ``engine.py`` with a single origin; this is the hot path.
* general -- 3D box and/or arbitrary quaternions: ``sdb_motion_general`` with the reference's own
  (N,3) state layout.  A planar object switches to it if it is ever handed a non-planar quaternion.
"""

import numpy as np

from .. import _lib
from ..engine import DynamicsEngine, finalize_row
from ..util import Box, as_box

F32 = np.float32


def _is_planar(orientation):
    if orientation is None:
        return True
    q = np.asarray(orientation)
    return not np.any(q[:, 1:3])


def create_wave_vector(wave_number: float, angular_resolution: int):
    """``angular_resolution`` wave vectors of length ``wave_number`` (reference ``_util.py:24-39``)."""
    angles = np.linspace(0, 2 * np.pi, num=angular_resolution, endpoint=False).reshape((-1, 1))
    return np.concatenate([np.cos(angles), np.sin(angles)], axis=1) * wave_number


class _Rdf:
    """What ``_static_structure_factor`` reads from a ``freud.density.RDF``: ``R``, ``RDF``, ``box``."""

    def __init__(self, R, RDF, box, counts):
        self.R, self.RDF, self.box, self.counts = R, RDF, box, counts


def radial_distribution(box: Box, positions, rmax: float, dr: float, device=None) -> _Rdf:
    """``freud.density.RDF(rmax=rmax, dr=dr).compute(box, positions)`` (reference call
    ``dynamics/_util.py:78-79``): the pair-distance histogram comes from ``sdb_rdf_histogram``
    (``csrc/rdf.cu``, all ordered pairs, f32 minimum image exactly as ``Box.wrap``); shell volumes and
    density normalisation are 200 bins of f32 arithmetic on the host."""
    torch = _lib.torch()
    device = _lib.require_cuda(device)
    lib = _lib.load()
    rmax32, dr32 = F32(rmax), F32(dr)
    nbins = int(np.floor(rmax32 / dr32))
    if nbins <= 0 or nbins > 1024:
        raise ValueError(f"rmax / dr gives {nbins} bins; between 1 and 1024 are supported")
    if isinstance(positions, torch.Tensor):
        pos = positions.to(device=device, dtype=torch.float32).contiguous()
    else:
        pos = torch.as_tensor(np.ascontiguousarray(positions, dtype=F32), device=device)
    n = pos.shape[0]
    counts = torch.empty(nbins, dtype=torch.int64, device=device)
    with _lib.device_context(device):
        _lib.check(
            lib.sdb_rdf_histogram(
                _lib.float_array(box.to_array()), int(box.is2D), n, pos.data_ptr(), float(rmax32), float(dr32), nbins,
                counts.data_ptr(), _lib.stream_ptr(device),
            )
        )
    counts = counts.cpu().numpy().astype(np.uint64)
    i = np.arange(nbins, dtype=F32)
    r0, r1 = i * dr32, (i + F32(1.0)) * dr32
    R = (F32(2.0) / F32(3.0) * (r1 * r1 * r1 - r0 * r0 * r0) / (r1 * r1 - r0 * r0)).astype(F32)
    if box.is2D:
        vol = F32(np.pi) * (r1 * r1 - r0 * r0)
    else:
        vol = F32(4.0) / F32(3.0) * F32(np.pi) * (r1 * r1 * r1 - r0 * r0 * r0)
    lx, ly, lz = (float(F32(v)) for v in (box.Lx, box.Ly, box.Lz))
    ndens = F32(n) / F32(lx * ly if box.is2D else lx * ly * lz)
    rdf = (counts.astype(F32) / F32(n) / vol.astype(F32) / ndens).astype(F32)
    return _Rdf(R, rdf, box, counts)


def _static_structure_factor(rdf, wave_number: float, num_particles: int):
    """Reference ``dynamics/_util.py:57-63``.  ``wave_number`` arrives as a numpy float64 scalar; with the
    reference's pinned numpy (< 1.19) its product with the f32 array ``rdf.R`` is f32, which is kept here."""
    dr = rdf.R[1] - rdf.R[0]
    integral = dr * np.sum((rdf.RDF - F32(1)) * rdf.R * np.sin(F32(wave_number) * rdf.R))
    lx, ly, lz = (float(F32(v)) for v in (rdf.box.Lx, rdf.box.Ly, rdf.box.Lz))
    density = num_particles / (lx * ly if rdf.box.is2D else lx * ly * lz)
    return 1 + 4 * np.pi * density / float(wave_number) * float(integral)


def calculate_wave_number(box: Box, positions: np.ndarray, device=None):
    """Wave number of the maximum of the static structure factor (reference ``dynamics/_util.py:66-86``;
    used by ``Relaxations`` when no wave number is given, ``relaxations.py:61-62``).  Like the reference
    says: not recommended -- pass the wave number explicitly."""
    box = as_box(box)
    lx, ly, lz = (float(F32(v)) for v in (box.Lx, box.Ly, box.Lz))
    rmax = min(lx / 2.2, ly / 2.2)
    if not box.is2D:
        rmax = min(rmax, lz / 2.2)
    dr = rmax / 200
    rdf = radial_distribution(box, positions, rmax=rmax, dr=dr, device=device)
    x = np.linspace(0.5, 20, 200)
    ssf = [_static_structure_factor(rdf, value, len(positions)) for value in x]
    return x[np.argmax(ssf)]


def molecule2particles(position: np.ndarray, orientation: np.ndarray, mol_vector: np.ndarray) -> np.ndarray:
    """Reference ``dynamics/_util.py:42-54`` verbatim in behaviour, including the pairing of SURVEY trap
    T2 (rotated sites are site-major, translations molecule-major).  Host-side helper: the hot path
    evaluates the same rows inside ``csrc/struct.cu`` without materialising them."""
    from ..util import rotate_vectors

    if isinstance(orientation, tuple):
        raise ValueError("Expecting numpy array found tuple", orientation)
    if np.allclose(orientation, 0.0):
        orientation[:, 0] = 1.0
    rotated = np.concatenate([rotate_vectors(orientation, pos) for pos in mol_vector.astype(np.float32)])
    translated = np.repeat(position, mol_vector.shape[0], axis=0)
    return rotated + translated


class _GeneralState:
    """(N,3) accumulators on the device, advanced by ``sdb_motion_general``."""

    def __init__(self, box: Box, position, orientation, device=None):
        self.torch = torch = _lib.torch()
        self.device = _lib.require_cuda(device)
        self.lib = _lib.load()
        self.box = box
        self.n = len(position)
        self._box_c = _lib.float_array(box.to_array())
        dev = self.device
        self.dtrans = torch.zeros((self.n, 3), dtype=torch.float32, device=dev)
        self.drot = torch.zeros((self.n, 3), dtype=torch.float32, device=dev)
        self.prev_pos = self._upload(position, 3)
        self.prev_quat = torch.zeros((self.n, 4), dtype=torch.float32, device=dev)
        self.has_orientation = orientation is not None
        if orientation is not None:
            self.prev_quat.copy_(self._upload(orientation, 4))
        self.bad = torch.zeros(1, dtype=torch.int32, device=dev)

    def _upload(self, array, width):
        array = np.ascontiguousarray(array, dtype=F32)
        if array.shape != (self.n, width):
            raise RuntimeError(f"array has shape {array.shape}, expected {(self.n, width)}")
        return self.torch.from_numpy(array).to(self.device)

    def add(self, position, orientation):
        torch = self.torch
        with torch.cuda.device(self.device):
            pos = self._upload(position, 3)
            quat = None if orientation is None else self._upload(orientation, 4)
            rotate = int(self.has_orientation and quat is not None)
            self.bad.zero_()
            _lib.check(
                self.lib.sdb_motion_general(
                    self._box_c, int(self.box.is2D), self.n, pos.data_ptr(), _lib.ptr(quat),
                    self.prev_pos.data_ptr(), self.prev_quat.data_ptr(), rotate, self.dtrans.data_ptr(),
                    self.drot.data_ptr(), self.bad.data_ptr(), _lib.stream_ptr(self.device),
                )
            )
            self.has_orientation = quat is not None
            if int(self.bad.item()):
                raise ValueError("Arguments must be unit quaternions")

    def sums(self, com_threshold, want_norms=False):
        torch = self.torch
        with torch.cuda.device(self.device):
            parts = int(self.lib.sdb_sums_general_parts(self.n))
            partials = torch.empty((parts, _lib.SDB_NSUM), dtype=torch.float64, device=self.device)
            sums = torch.zeros(_lib.SDB_NSUM, dtype=torch.float64, device=self.device)
            norms = torch.empty((2, self.n), dtype=torch.float32, device=self.device)
            _lib.check(
                self.lib.sdb_sums_general(
                    self.n, self.dtrans.data_ptr(), self.drot.data_ptr(),
                    float(com_threshold) if com_threshold is not None else -1.0,
                    norms[0].data_ptr(), norms[1].data_ptr(), partials.data_ptr(), sums.data_ptr(),
                    _lib.stream_ptr(self.device),
                )
            )
        return sums, norms


class TrackedMotion:
    """Keep track of the motion of particles through periodic boundaries and full rotations.

    Same constructor and ``add`` as the reference (``_util.py:108-156``); ``box`` may be a
    :class:`sdanalysis_b200.util.Box` or a box array.  The accumulators live on the GPU;
    ``delta_translation`` / ``delta_rotation`` copy them back as (N,3) arrays.
    """

    def __init__(self, box, position, orientation, *, timestep=0, engine_options=None, device=None):
        self.box = as_box(box)
        position = np.asarray(position)
        if orientation is not None:
            orientation = np.asarray(orientation)
            if orientation.shape[0] == 0:
                raise RuntimeError("Orientation must contain values, has length of 0.")
        if position.shape[0] == 0:
            raise RuntimeError("Position must contain values, has length of 0.")
        self.num_particles = position.shape[0]
        self.timestep = int(timestep)
        self._options = dict(engine_options or {})
        self._device = device
        self.previous_position = position
        self.previous_orientation = orientation
        self._engine = None
        self._general = None
        if self.box.is2D and _is_planar(orientation):
            opts = dict(trackers=None, compute_sums=True)
            opts.update(self._options)
            self._engine = DynamicsEngine(self.num_particles, self.box.to_array(), device=device, **opts)
            # creating the origin initialises its state with the first sample as "previous"
            self._engine.step(self.timestep, position, orientation, [0], want_sums=False)
        else:
            self._general = _GeneralState(self.box, position, orientation, device)

    # -- backend switch ---------------------------------------------------------------------------
    def _to_general(self):
        """Move a planar state into the general layout (first non-planar quaternion seen)."""
        eng = self._engine
        org = eng.origins[0]
        trans, rot = eng.delta_arrays(0)
        prev = org.prev_planes[:, : self.num_particles].cpu().numpy()
        pos = np.zeros((self.num_particles, 3), dtype=F32)
        pos[:, 0], pos[:, 1] = prev[0], prev[1]
        quat = None
        if org.prev_has_orientation:
            quat = np.zeros((self.num_particles, 4), dtype=F32)
            quat[:, 0], quat[:, 3] = prev[2], prev[3]
        gen = _GeneralState(self.box, pos, quat, self._device)
        gen.dtrans.copy_(gen.torch.from_numpy(trans))
        gen.drot.copy_(gen.torch.from_numpy(rot))
        self._general, self._engine = gen, None

    @property
    def is_planar(self):
        return self._engine is not None

    # -- reference API ---------------------------------------------------------------------------
    def add(self, position, orientation, *, timestep=None, want_sums=False):
        position = np.asarray(position)
        if orientation is not None:
            orientation = np.asarray(orientation)
        if self._engine is not None and not _is_planar(orientation):
            self._to_general()
        result = None
        if self._engine is not None:
            if not want_sums and orientation is not None and self.previous_orientation is not None:
                # rowan.to_euler rejects quotients that are not unit quaternions (reference _util.py:151-153
                # raises ValueError out of TrackedMotion.add / Relaxations.add).  With sums the device counts
                # them (SDB_S_BADQUAT, checked when the row is read); without sums nothing is read back, so
                # the per-object API validates here, before any state changes.
                n_cur = np.linalg.norm(np.asarray(orientation, dtype=np.float64), axis=1)
                n_prev = np.linalg.norm(np.asarray(self.previous_orientation, dtype=np.float64), axis=1)
                with np.errstate(divide="ignore", invalid="ignore"):
                    if not np.allclose(n_cur / n_prev, 1.0):
                        raise ValueError("Arguments must be unit quaternions")
            ts = self.timestep if timestep is None else timestep
            result = self._engine.step(ts, position, orientation, [0], want_sums=want_sums)
        else:
            self._general.add(position, orientation)
        self.previous_position = position
        self.previous_orientation = orientation
        return result

    @property
    def delta_translation(self):
        if self._engine is not None:
            return self._engine.delta_arrays(0)[0]
        return self._general.dtrans.cpu().numpy()

    @property
    def delta_rotation(self):
        if self._engine is not None:
            return self._engine.delta_arrays(0)[1]
        return self._general.drot.cpu().numpy()

    # -- device reductions used by Dynamics / Relaxations -----------------------------------------
    def norms(self):
        """(|dr|, |dtheta|) per molecule as f32 host arrays, evaluated on the device."""
        if self._engine is not None:
            return self._engine.norms(0)
        _, norms = self._general.sums(None)
        host = norms.cpu().numpy()
        return host[0], host[1]

    def isf(self, wave_number, angular_resolution=360):
        """Mean over ``angular_resolution`` wave vectors and over the molecules of ``cos(k . dr)``
        (``Dynamics.compute_isf``, reference ``dynamics.py:233-235``), evaluated on the device."""
        torch, lib = _lib.torch(), _lib.load()
        if self._engine is not None:
            eng = self._engine
            delta = eng.origins[0].delta
            device, n = eng.device, eng.n
            args = (delta.data_ptr(), delta.data_ptr() + 4 * eng.stride, 1)
        else:
            gen = self._general
            device, n = gen.device, gen.n
            args = (gen.dtrans.data_ptr(), gen.dtrans.data_ptr() + 4, 3)
        with torch.cuda.device(device):
            out = torch.zeros(1, dtype=torch.float64, device=device)
            _lib.check(
                lib.sdb_isf(*args, n, float(wave_number), int(angular_resolution), out.data_ptr(), _lib.stream_ptr(device))
            )
            return float(out.item()) / n

    def current_row(self, timediff=0):
        """The ``compute_all`` quantities of the current state without adding a sample."""
        if self._engine is not None:
            res = self._engine.step(self.timestep + timediff, None, None, [0], zero_step=True, want_sums=True)
            return res.rows(strict=False)[0]
        eng_opts = self._options
        wave_number = eng_opts.get("wave_number")
        distance = None if wave_number is None else np.pi / (2 * wave_number)
        sums, norms = self._general.sums(distance)
        s = sums.cpu().numpy()
        k = int(self.num_particles * 0.1)
        overlap_k = None
        if k > 0:
            s[_lib.S_OVERLAP] = _overlap_from_norms(self._general, norms, k)
            overlap_k = k
        return finalize_row(s, timediff, self.num_particles, distance=distance, overlap_k=overlap_k, n_sites=None)


def _overlap_from_norms(state, norms, k):
    """``mobile_overlap`` on device-resident norms (general layout): sdb_overlap with key_mode 1."""
    torch, lib = state.torch, state.lib
    n = state.n
    stride = (n + 3) // 4 * 4
    with torch.cuda.device(state.device):
        planes = torch.zeros((3, stride), dtype=torch.float32, device=state.device)
        planes[0, :n] = norms[0]
        planes[2, :n] = norms[1]
        return overlap_count(planes, n, stride, k, state.device)


def overlap_count(planes, n, stride, k, device):
    """|top-k(plane 0) & top-k(|plane 2|)| with the exact device selection (``csrc/overlap.cu``)."""
    torch, lib = _lib.torch(), _lib.load()
    desc = np.zeros(1, dtype=_lib.ORIGIN_DTYPE)
    desc[0] = (planes.data_ptr(), 0, 0, 0, 0)
    d_desc = torch.from_numpy(desc.view(np.uint8).copy()).to(device)
    scratch = torch.empty(int(lib.sdb_overlap_scratch_bytes(1, n)), dtype=torch.uint8, device=device)
    sums = torch.zeros((1, _lib.SDB_NSUM), dtype=torch.float64, device=device)
    _lib.check(
        lib.sdb_overlap(d_desc.data_ptr(), 1, n, stride, 1, k, scratch.data_ptr(), sums.data_ptr(), _lib.stream_ptr(device))
    )
    return float(sums[0, _lib.S_OVERLAP].item())

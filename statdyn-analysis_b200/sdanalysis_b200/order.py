"""Neighbour lists and orientational order, drop-in for the hot-path part of reference ``order.py``.

``compute_neighbours`` / ``orientational_order`` / ``num_neighbours`` / ``relative_distances`` /
``relative_orientations`` keep the reference signatures (``order.py:186-309``).  The search is a GPU
cell list (``csrc/neighbours.cu``) replacing ``freud.locality.NearestNeighbors``; one call of
:func:`neighbour_analysis` produces everything from a single build, where the reference rebuilds the
neighbour structure for each function.  Voronoi counts and the sklearn classifier factories of the
reference (``order.py:89-171,312-315``) are outside the accelerated path.
"""

import warnings

import numpy as np

from . import _lib

F32 = np.float32


def _box6(box):
    box = np.asarray(box, dtype=F32).reshape(-1)
    out = np.zeros(6, dtype=F32)
    out[: len(box)] = box
    return out


def _to_device(array, width, torch, device):
    if isinstance(array, torch.Tensor):
        t = array.to(device=device, dtype=torch.float32)
        return t.contiguous()
    array = np.ascontiguousarray(array, dtype=F32)
    if array.ndim != 2 or array.shape[1] != width:
        raise ValueError(f"expected an (N, {width}) array, got {array.shape}")
    return torch.from_numpy(array).to(device)


# ---- one cell-list build per frame, shared by the public functions -------------------------------------------
# The reference rebuilds freud's neighbour structure in compute_neighbours, orientational_order and
# num_neighbours (order.py:207,258-261,279-281): a frame that wants all three pays three builds.  Here the
# functions share ONE search of capacity 9 (its first 8 columns are the k = 8 list; the count saturates at 9 like
# num_neighbours') as long as they are called with the same frame.  "The same frame" must be cheap and safe to
# recognise: a torch tensor (same storage, shape and in-place version counter) or a read-only numpy array (same
# address and shape, e.g. a view of a memory-mapped trajectory).  Writable numpy arrays are never cached.
_workspaces = {}
_analysis_cache = {"key": None, "value": None}


def _immutable(array):
    """A numpy array nobody can write through: read-only, and so is every array it is a view of."""
    while isinstance(array, np.ndarray):
        if array.flags.writeable:
            return False
        array = array.base
    return True


def _identity(array, torch):
    """Hashable identity of a frame array's *contents*, or None when they cannot be recognised cheaply.  Valid
    only while the cache holds a reference to the array (its memory cannot be reused for something else)."""
    if isinstance(array, torch.Tensor):
        return ("t", array.data_ptr(), tuple(array.shape), array.dtype, array._version, str(array.device))
    if isinstance(array, np.ndarray) and _immutable(array):
        return ("n", array.ctypes.data, array.shape, array.dtype.str, array.strides)
    return None


def _workspace(lib, torch, device, box6, n, max_radius):
    key = (device, n, box6.tobytes(), float(max_radius))
    ws = _workspaces.get(key)
    if ws is None:
        box_c = _lib.float_array(box6)
        nbytes = int(lib.sdb_neighbours_scratch_bytes(box_c, n, float(max_radius)))
        if nbytes == 0:
            raise ValueError("invalid box or radius")
        if len(_workspaces) > 8:
            _workspaces.clear()
        ws = _workspaces[key] = (box_c, torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes)
    return ws


def neighbour_analysis(
    box, position, orientation=None, max_radius=3.5, max_neighbours=8, *, want=("nlist",), device=None, to_host=None,
    count_capacity=None,
):
    """One cell-list build, any of ``nlist`` (N,k) uint32, ``rsq`` (N,k) f32, ``counts`` (N,) uint32,
    ``order`` (N,) f64.  Missing neighbours are 2**32-1 in ``nlist`` and -1 in ``rsq``; ``counts`` saturates at
    ``count_capacity`` (default: ``max_neighbours``).  Results are device tensors (int32 views of the uint32
    data) when the frame was given as torch tensors or ``to_host=False``, numpy arrays otherwise."""
    torch = _lib.torch()
    device = _lib.require_cuda(device)
    lib = _lib.load()
    k = int(max_neighbours)
    if not 0 < k <= _lib.SDB_MAX_K:
        raise ValueError(f"max_neighbours must be in 1..{_lib.SDB_MAX_K}")
    cap = k if count_capacity is None else max(int(count_capacity), k)
    if cap > _lib.SDB_MAX_K:
        raise ValueError(f"count capacity must be <= {_lib.SDB_MAX_K}")
    if to_host is None:
        to_host = not isinstance(position, torch.Tensor)
    with _lib.device_context(device):
        pos = _to_device(position, 3, torch, device)
        n = pos.shape[0]
        if n == 0:
            raise RuntimeError("Position must contain values, has length of 0.")
        quat = None
        if "order" in want:
            if orientation is None:
                raise ValueError("orientational order needs orientations")
            quat = _to_device(orientation, 4, torch, device)
        box_c, scratch, nbytes = _workspace(lib, torch, device, _box6(box), n, max_radius)
        out = {}
        if "nlist" in want:
            out["nlist"] = torch.empty((n, k), dtype=torch.int32, device=device)
        if "rsq" in want:
            out["rsq"] = torch.empty((n, k), dtype=torch.float32, device=device)
        if "counts" in want:
            out["counts"] = torch.empty(n, dtype=torch.int32, device=device)
        if "order" in want:
            out["order"] = torch.empty(n, dtype=torch.float64, device=device)
        _lib.check(
            lib.sdb_neighbours_ex(
                box_c, n, pos.data_ptr(), _lib.ptr(quat), float(max_radius), cap, k,
                _lib.ptr(out.get("nlist")), _lib.ptr(out.get("rsq")), _lib.ptr(out.get("counts")),
                _lib.ptr(out.get("order")), scratch.data_ptr(), nbytes, _lib.stream_ptr(device),
            )
        )
        if not to_host:
            return out
        host = {key: val.cpu().numpy() for key, val in out.items()}
    for key in ("nlist", "counts"):
        if key in host:
            host[key] = host[key].view(np.uint32)
    return host


def _frame_analysis(box, position, orientation, max_radius, max_neighbours, need):
    """The shared build: list / r^2 of width ``max_neighbours``, counts saturating at max(9, max_neighbours),
    order parameter when orientations are known (computed from the cached list when they arrive later).
    Cached for the frame it was computed for (see above)."""
    torch = _lib.torch()
    lib = _lib.load()
    k = int(max_neighbours)
    cap = max(k, 9)
    ident = _identity(position, torch)
    key = None if ident is None else (ident, _box6(box).tobytes(), float(max_radius), cap, k)
    cached = _analysis_cache["value"] if (key is not None and _analysis_cache["key"] == key) else None
    if cached is not None:
        if need != "order":
            return cached
        quat_ident = _identity(orientation, torch)
        if "order" in cached and quat_ident is not None and cached["quat_ident"] == quat_ident:
            return cached
        if orientation is None:
            raise ValueError("orientational order needs orientations")
        # the list of this frame is known: the order parameter is one more small kernel, not another build
        device = cached["nlist"].device
        with _lib.device_context(device):
            quat = _to_device(orientation, 4, torch, device)
            n = cached["nlist"].shape[0]
            if quat.shape[0] != n:
                raise ValueError("orientation and position have different lengths")
            order = torch.empty(n, dtype=torch.float64, device=device)
            _lib.check(lib.sdb_orientational_order(cached["nlist"].data_ptr(), n, k, quat.data_ptr(), order.data_ptr(), 0,
                                                   _lib.stream_ptr(device)))
        cached["order"] = order
        cached["quat_ident"] = quat_ident
        cached["refs"] = (cached["refs"][0], orientation)
        cached["host"].pop("order", None)
        return cached
    if need == "order" and orientation is None:
        raise ValueError("orientational order needs orientations")
    want = ("nlist", "rsq", "counts") + (("order",) if orientation is not None else ())
    result = neighbour_analysis(box, position, orientation, max_radius, k, want=want, to_host=False, count_capacity=cap)
    result["quat_ident"] = _identity(orientation, torch) if orientation is not None else None
    result["host"] = {}
    result["refs"] = (position, orientation)  # keeps the memory the identities refer to from being reused
    if key is not None:
        _analysis_cache["key"], _analysis_cache["value"] = key, result
    return result


def _deliver(result, name, position):
    """Device tensor for torch callers, numpy array (cached per frame) for numpy callers."""
    torch = _lib.torch()
    if isinstance(position, torch.Tensor):
        return result[name]
    host = result["host"].get(name)
    if host is None:
        host = result[name].cpu().numpy()
        if name in ("nlist", "counts"):
            host = host.view(np.uint32)
        result["host"][name] = host
    return host


class NeighbourQuery:
    """What ``setup_neighbours`` returns (stands in for ``freud.locality.NearestNeighbors`` after
    ``compute``): ``getNeighborList()`` (N,k) uint32, ``getRsqList()`` (N,k) f32 with -1 for a missing
    neighbour, ``nlist.neighbor_counts`` (N,) uint32 -- all from one device cell-list build."""

    def __init__(self, result):
        self._result = result
        self.nlist = self

    def getNeighborList(self):
        return self._result["nlist"]

    def getRsqList(self):
        return self._result["rsq"]

    @property
    def r_sq_list(self):
        return self._result["rsq"]

    @property
    def neighbor_counts(self):
        return self._result["counts"]


def setup_neighbours(box, position, max_radius: float = 3.5, max_neighbours: int = 8, is_2D: bool = True):
    """Reference ``order.py:174-183``: k nearest neighbours within ``max_radius`` (strict cut)."""
    if not is_2D:
        raise NotImplementedError("the cell-list kernel covers the reference's 2D boxes only")
    return NeighbourQuery(
        neighbour_analysis(box, position, None, max_radius, max_neighbours, want=("nlist", "rsq", "counts"), to_host=True)
    )


def compute_neighbours(box, position, max_radius: float = 3.5, max_neighbours: int = 8) -> np.ndarray:
    """Indices of the nearest neighbours of each molecule, (N, max_neighbours) uint32, sorted by
    distance, ``2**32 - 1`` marking a missing neighbour (reference ``order.py:186-207``).  Torch tensors in:
    an int32 device tensor (the same bits) out, without synchronising."""
    return _deliver(_frame_analysis(box, position, None, max_radius, max_neighbours, "nlist"), "nlist", position)


def _list_on_device(neighbourlist, orientation, torch, device):
    neighbourlist = np.asarray(neighbourlist)
    if len(neighbourlist) != len(orientation):
        warnings.warn(
            "Having different lengths for the neighbourlist and orientation arguments can lead to unusual "
            f"outcomes. Found {len(neighbourlist)} and {len(orientation)}"
        )
    n = len(orientation)
    # order.py:56-59 -- entries that are not valid molecule indices mean "missing"
    nl = np.where((neighbourlist >= 0) & (neighbourlist < n), neighbourlist, 2 ** 32 - 1).astype(np.uint32)
    d_nl = torch.from_numpy(np.ascontiguousarray(nl).view(np.int32)).to(device)
    return d_nl, _to_device(orientation, 4, torch, device)


def _neighbour_relative_angle(neighbourlist: np.ndarray, orientation: np.ndarray) -> np.ndarray:
    """Relative angle to every listed neighbour in [0, 2 pi], 0 for missing ones
    (reference ``order.py:28-65``)."""
    torch = _lib.torch()
    device = _lib.require_cuda()
    with torch.cuda.device(device):
        d_nl, quat = _list_on_device(neighbourlist, orientation, torch, device)
        n, k = d_nl.shape
        angles = torch.empty((n, k), dtype=torch.float64, device=device)
        _lib.check(
            _lib.load().sdb_orientational_order(
                d_nl.data_ptr(), n, k, quat.data_ptr(), 0, angles.data_ptr(), _lib.stream_ptr(device)
            )
        )
        return angles.cpu().numpy()


def _orientational_order(neighbourlist: np.ndarray, orientation: np.ndarray) -> np.ndarray:
    """mean_k cos^2(relative angle), in [0, 1] (reference ``order.py:68-86``)."""
    torch = _lib.torch()
    device = _lib.require_cuda()
    with torch.cuda.device(device):
        d_nl, quat = _list_on_device(neighbourlist, orientation, torch, device)
        n, k = d_nl.shape
        order = torch.empty(n, dtype=torch.float64, device=device)
        _lib.check(
            _lib.load().sdb_orientational_order(
                d_nl.data_ptr(), n, k, quat.data_ptr(), order.data_ptr(), 0, _lib.stream_ptr(device)
            )
        )
        return order.cpu().numpy()


def relative_orientations(box, position, orientation, max_radius: float = 3.5, max_neighbours: int = 8) -> np.ndarray:
    """Reference ``order.py:210-231``."""
    neighbours = compute_neighbours(box, position, max_radius, max_neighbours)
    return _neighbour_relative_angle(neighbours, orientation)


def orientational_order(box, position, orientation, max_radius: float = 3.5, max_neighbours: int = 8) -> np.ndarray:
    """Orientational order parameter of every molecule (reference ``order.py:234-262``)."""
    return _deliver(_frame_analysis(box, position, orientation, max_radius, max_neighbours, "order"), "order", position)


def num_neighbours(box, position, max_radius: float = 3.5) -> np.ndarray:
    """Number of neighbours within ``max_radius``; the search has capacity 9, so it saturates at 9 like the
    reference (``order.py:265-281``, SURVEY trap T8)."""
    torch = _lib.torch()
    ident = _identity(position, torch)
    key, cached = _analysis_cache["key"], _analysis_cache["value"]
    if ident is not None and key is not None and key[:4] == (ident, _box6(box).tobytes(), float(max_radius), 9):
        return _deliver(cached, "counts", position)  # any list width <= 9 of this frame has the count
    return _deliver(_frame_analysis(box, position, None, max_radius, 8, "counts"), "counts", position)


def relative_distances(box, position, max_radius: float = 3.5, max_neighbours: int = 8) -> np.ndarray:
    """Distance to each neighbour, 0 where missing (reference ``order.py:284-309``)."""
    rsq = neighbour_analysis(box, position, None, max_radius, max_neighbours, want=("rsq",), to_host=True)["rsq"]
    distances = rsq.astype(np.float64)
    distances[distances < 0] = 0
    return np.sqrt(distances)


def compute_voronoi_neighs(box, position) -> np.ndarray:
    """Number of Voronoi neighbours of every molecule (reference ``order.py:312-315``, which asks
    ``freud.voronoi.Voronoi(box, buff=5).computeNeighbors(position)`` for ``nlist.neighbor_counts``).

    The 16 nearest neighbours within ~4 mean spacings come from the cell-list kernel; ``csrc/voronoi.cu``
    clips every molecule's cell by their bisectors in order of distance until no farther molecule can cut
    it, and redoes the few molecules whose list ends too early against all molecules."""
    torch = _lib.torch()
    device = _lib.require_cuda()
    lib = _lib.load()
    box6 = _box6(box)
    if box6[3] != 0:
        raise NotImplementedError("Voronoi neighbour counts cover orthorhombic 2D boxes only")
    with torch.cuda.device(device):
        pos = _to_device(position, 3, torch, device)
        n = pos.shape[0]
        if n == 0:
            raise RuntimeError("Position must contain values, has length of 0.")
        k = _lib.SDB_MAX_K
        spacing = float(np.sqrt(float(box6[0]) * float(box6[1]) / n))
        rmax = min(4.0 * spacing, 0.49 * float(min(box6[0], box6[1])))
        found = neighbour_analysis(box6, pos, None, rmax, k, want=("nlist",), device=device, to_host=False)  # (own build: k = 16)
        counts = torch.empty(n, dtype=torch.int32, device=device)
        scratch = torch.empty(n + 1, dtype=torch.int32, device=device)
        _lib.check(
            lib.sdb_voronoi_counts(
                _lib.float_array(box6), n, pos.data_ptr(), found["nlist"].data_ptr(), k, float(rmax),
                counts.data_ptr(), scratch.data_ptr(), _lib.stream_ptr(device),
            )
        )
        out = counts.cpu().numpy().view(np.uint32)
    if (out == 2 ** 32 - 1).any():
        raise RuntimeError("a Voronoi cell reaches the molecule's own periodic images (too few molecules for the box)")
    return out


def create_orient_ordering(threshold: float):
    """Reference ``order.py:126-147``."""

    def compute_orient_ordering(snap) -> np.ndarray:
        return orientational_order(snap.box, snap.position, snap.orientation) > threshold

    return compute_orient_ordering


def create_neigh_ordering(neighbours: int):
    """Reference ``order.py:150-171``: a molecule is ordered when it has exactly ``neighbours`` Voronoi
    neighbours."""

    def compute_neigh_ordering(snap) -> np.ndarray:
        return compute_voronoi_neighs(snap.box, snap.position) == neighbours

    return compute_neigh_ordering


def create_ml_ordering(model):
    """Reference ``order.py:89-123``: classify the relative orientations of the 8 nearest neighbours
    within 3.5 with a pickled scikit-learn model (loaded with joblib like the reference)."""
    import joblib

    ml_model = joblib.load(model)

    def compute_ml_order(snap) -> np.ndarray:
        orientations = relative_orientations(snap.box, snap.position, snap.orientation, 3.5, 8)
        return ml_model.predict(orientations)

    return compute_ml_order

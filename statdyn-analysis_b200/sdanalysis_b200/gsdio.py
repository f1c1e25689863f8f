"""Minimal GSD (HOOMD schema) trajectory reader and writer.

The reference reads trajectories with the third-party ``gsd`` package (``gsd.hoomd.open``,
reference ``read/_gsd.py:100,128``; ``read/_read.py:218``).  That package is a C extension which is
not part of this build, so the on-disk layout is decoded directly here.  Frames are returned as
zero-copy, read-only numpy views over an ``mmap`` of the file, ready to be staged into pinned
memory by the driver loop.

File layout (GSD file-layer v1.0 / v2.0):

* 256-byte header ``<QQQQQII 64s 64s 80x``: magic ``0x65DF65DF65DF65DF``, index_location,
  index_allocated_entries, namelist_location, namelist_allocated_entries, schema_version,
  gsd_version, application, schema.
* index entries, 32 bytes ``<QQqIHBB``: frame, N, location, M, name id, type, flags.  An entry
  with ``location == 0`` terminates the list.
* namelist: v1 = fixed 64-byte zero-padded strings; v2 = packed zero-terminated strings.
* chunk data: ``N x M`` elements of the given type, row-major.
"""

import mmap
import os
import struct
from pathlib import Path

import numpy as np

MAGIC = 0x65DF65DF65DF65DF
_HEADER = struct.Struct("<QQQQQII64s64s80x")
_ENTRY = np.dtype(
    [("frame", "<u8"), ("N", "<u8"), ("location", "<i8"), ("M", "<u4"), ("id", "<u2"), ("type", "u1"), ("flags", "u1")]
)
_TYPES = {
    1: np.dtype("u1"),
    2: np.dtype("<u2"),
    3: np.dtype("<u4"),
    4: np.dtype("<u8"),
    5: np.dtype("i1"),
    6: np.dtype("<i2"),
    7: np.dtype("<i4"),
    8: np.dtype("<i8"),
    9: np.dtype("<f4"),
    10: np.dtype("<f8"),
}
_TYPE_CODES = {v: k for k, v in _TYPES.items()}
_NAME_SIZE = 64


class Snapshot:
    """One trajectory frame with the ``gsd.hoomd.Snapshot`` fields the hot path consumes.

    Mirrors the contract of reference ``frame.py:128-205`` (``HoomdFrame``): ``position`` and
    ``orientation`` are the first ``num_mols`` rows, ``num_mols = max(body) + 1`` clamped to the
    particle count (``frame.py:136-156``).
    """

    __slots__ = ("step", "dimensions", "box", "N", "_position", "_orientation", "body", "image", "num_mols")

    def __init__(self, step, dimensions, box, N, position, orientation, body=None, image=None):
        self.step = int(step)
        self.dimensions = int(dimensions)
        self.box = box
        self.N = int(N)
        self._position = position
        self._orientation = orientation
        self.body = body
        self.image = image
        self.num_mols = self._num_bodies()

    def _num_bodies(self):
        num_mols = self.N
        if self.body is not None and len(self.body) > 0:
            num_mols = int(self.body.max()) + 1
        if num_mols > self.N or num_mols <= 0:
            num_mols = self.N
        if num_mols > len(self._position):
            raise ValueError("There are more molecules than position datapoints.")
        return num_mols

    @property
    def position(self):
        return self._position[: self.num_mols]

    @property
    def orientation(self):
        return self._orientation[: self.num_mols]

    @property
    def timestep(self):
        return self.step

    def __len__(self):
        return self.num_mols


class GSDTrajectory:
    """Read-only random access to the frames of a GSD file."""

    def __init__(self, filename):
        self.filename = Path(filename)
        self._fh = open(self.filename, "rb")
        try:
            if os.fstat(self._fh.fileno()).st_size < _HEADER.size:
                raise RuntimeError(f"{filename}: not a GSD file (too short)")
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
            try:
                self._read_index(filename)
            except (ValueError, IndexError, struct.error, UnicodeDecodeError) as exc:
                # index or name list outside the file, name ids outside the name list: like the gsd package
                # (RuntimeError for a file it cannot index), so callers catch one exception type
                raise RuntimeError(f"{filename}: corrupt GSD file ({exc})") from exc
        except Exception:
            self._fh.close()
            raise

    def _read_index(self, filename):
        (magic, iloc, ialloc, nloc, nalloc, self.schema_version, self.gsd_version, app, schema) = _HEADER.unpack_from(
            self._mm, 0
        )
        if magic != MAGIC:
            raise RuntimeError(f"{filename}: not a GSD file (bad magic)")
        self.application = app.split(b"\0")[0].decode()
        self.schema = schema.split(b"\0")[0].decode()
        raw_names = bytes(self._mm[nloc : nloc + nalloc * _NAME_SIZE])
        if (self.gsd_version >> 16) >= 2:
            names = raw_names.split(b"\0")
            names = names[: names.index(b"")] if b"" in names else names
        else:
            names = [raw_names[i : i + _NAME_SIZE].split(b"\0")[0] for i in range(0, len(raw_names), _NAME_SIZE)]
            names = [n for n in names if n]
        self._names = [n.decode() for n in names]
        index = np.frombuffer(self._mm, dtype=_ENTRY, count=ialloc, offset=iloc)
        used = np.flatnonzero(index["location"] == 0)
        index = index[: used[0]] if len(used) else index
        self._frames = {}
        names = self._names
        # column-wise .tolist(): per-element access to a structured array costs microseconds per field
        for frame, name_id, loc, n, m, typ in zip(
            index["frame"].tolist(), index["id"].tolist(), index["location"].tolist(), index["N"].tolist(),
            index["M"].tolist(), index["type"].tolist(),
        ):
            self._frames.setdefault(frame, {})[names[name_id]] = (loc, n, m, typ)
        self._nframes = (max(self._frames) + 1) if self._frames else 0

    def __len__(self):
        return self._nframes

    def pin(self):
        """Page-lock the file mapping (``cudaHostRegister``, read-only) so that frames are uploaded to the GPU
        straight from the page cache, without a staging copy.  Returns True on success; on failure (no CUDA,
        read-only registration unsupported, locked-memory limit) nothing changes and frames take the staged path."""
        if getattr(self, "_pinned_addr", None):
            return True
        try:
            from . import _lib

            torch = _lib.torch()
            if not torch.cuda.is_available():
                return False
            lib = _lib.load()
            view = np.frombuffer(self._mm, dtype=np.uint8)
            addr, size = view.ctypes.data, view.nbytes
            if size > int(float(os.environ.get("SDB_PIN_MAX_GB", "16")) * 2 ** 30):
                self.pin_error = "file larger than SDB_PIN_MAX_GB"
                return False
            if lib.sdb_host_register(addr, size) != 0:
                self.pin_error = lib.sdb_last_error().decode()
                # the driver pins pages through a writable mapping only: map the file read-write when its
                # permissions allow (nothing is ever written through it; the views handed out stay read-only)
                del view
                try:
                    fh = open(self.filename, "r+b")
                    mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_WRITE)
                except (OSError, ValueError) as exc:
                    self.pin_error += f"; no writable mapping: {exc}"
                    return False
                view = np.frombuffer(mm, dtype=np.uint8)
                addr, size = view.ctypes.data, view.nbytes
                if lib.sdb_host_register(addr, size) != 0:
                    self.pin_error += "; writable mapping: " + lib.sdb_last_error().decode()
                    del view
                    mm.close()
                    fh.close()
                    return False
                self._fh_ro, self._mm_ro = self._fh, self._mm  # views handed out earlier keep the old mapping alive
                self._fh, self._mm = fh, mm
                self.pin_error = None
            self._pinned_addr = addr
            _lib.register_pinned(addr, size)
            return True
        except Exception as exc:
            self.pin_error = repr(exc)
            return False

    def _unpin(self):
        addr = getattr(self, "_pinned_addr", None)
        if addr:
            from . import _lib

            _lib.torch().cuda.synchronize()  # uploads reading the mapping have finished
            _lib.load().sdb_host_unregister(addr)
            _lib.unregister_pinned(addr)
            self._pinned_addr = None

    def close(self):
        self._unpin()
        # numpy views keep the mmap alive; closing is best effort
        try:
            self._mm.close()
        except BufferError:
            pass
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def chunk(self, frame, name):
        """Return the chunk as a read-only array view, falling back to frame 0, else ``None``."""
        for f in (frame, 0):
            entry = self._frames.get(f, {}).get(name)
            if entry is not None:
                loc, n, m, typ = entry
                dtype = _TYPES.get(typ)
                if dtype is None or loc + n * m * dtype.itemsize > len(self._mm):
                    raise RuntimeError(f"corrupt chunk {name} in frame {frame}")
                arr = np.frombuffer(self._mm, dtype=dtype, count=n * m, offset=loc)
                if arr.flags.writeable:  # (mapping made writable for pinning: the views stay read-only)
                    arr.flags.writeable = False
                return arr.reshape(n, m) if m > 1 else arr
        return None

    def __getitem__(self, idx):
        if idx < 0:
            idx += self._nframes
        if not 0 <= idx < self._nframes:
            raise IndexError(idx)
        step = self.chunk(idx, "configuration/step")
        dims = self.chunk(idx, "configuration/dimensions")
        box = self.chunk(idx, "configuration/box")
        n = self.chunk(idx, "particles/N")
        step = 0 if step is None else int(step[0])
        dims = 3 if dims is None else int(dims[0])
        box = np.array([1, 1, 1, 0, 0, 0], dtype=np.float32) if box is None else box
        n = 0 if n is None else int(n[0])

        def per_particle(name, default, width):
            arr = self.chunk(idx, name)
            if arr is None or len(arr) != n:
                arr = np.tile(np.asarray(default), (n, 1)) if width > 1 else np.full(n, default)
            return arr

        position = per_particle("particles/position", np.zeros(3, np.float32), 3)
        orientation = per_particle("particles/orientation", np.array([1, 0, 0, 0], np.float32), 4)
        body = self.chunk(idx, "particles/body")
        if body is not None and len(body) != n:
            body = None
        image = self.chunk(idx, "particles/image")
        return Snapshot(step, dims, box, n, position, orientation, body, image)

    def __iter__(self):
        for i in range(self._nframes):
            yield self[i]


def open_gsd(filename):
    return GSDTrajectory(filename)


class GSDWriter:
    """Append-only writer producing GSD v1.0 files with the HOOMD schema chunks the reader uses."""

    def __init__(self, filename, application="sdanalysis_b200", schema="hoomd"):
        self._fh = open(filename, "wb")
        self._names = []
        self._entries = []
        self._frame = 0
        self._application = application
        self._schema = schema
        self._fh.write(b"\0" * _HEADER.size)

    def _chunk(self, name, array, width):
        array = np.ascontiguousarray(array)
        if name not in self._names:
            self._names.append(name)
        n = array.shape[0] if array.ndim else 1
        self._entries.append(
            (self._frame, n, self._fh.tell(), width, self._names.index(name), _TYPE_CODES[array.dtype.newbyteorder("<")], 0)
        )
        self._fh.write(array.tobytes())

    def append(self, step, box, position, orientation=None, dimensions=2, body=None, image=None):
        position = np.asarray(position, dtype=np.float32).reshape(-1, 3)
        self._chunk("configuration/step", np.array([step], dtype=np.uint64), 1)
        self._chunk("configuration/dimensions", np.array([dimensions], dtype=np.uint8), 1)
        self._chunk("configuration/box", np.asarray(box, dtype=np.float32).reshape(6), 1)
        self._chunk("particles/N", np.array([len(position)], dtype=np.uint32), 1)
        self._chunk("particles/position", position, 3)
        if orientation is not None:
            self._chunk("particles/orientation", np.asarray(orientation, dtype=np.float32).reshape(-1, 4), 4)
        if body is not None:
            self._chunk("particles/body", np.asarray(body, dtype=np.int32), 1)
        if image is not None:
            self._chunk("particles/image", np.asarray(image, dtype=np.int32).reshape(-1, 3), 3)
        self._frame += 1

    def close(self):
        if self._fh is None:
            return
        nloc = self._fh.tell()
        nalloc = max(128, len(self._names))
        names = b"".join(n.encode().ljust(_NAME_SIZE, b"\0") for n in self._names)
        self._fh.write(names.ljust(nalloc * _NAME_SIZE, b"\0"))
        iloc = self._fh.tell()
        index = np.zeros(len(self._entries) + 1, dtype=_ENTRY)
        for i, e in enumerate(self._entries):
            index[i] = e
        self._fh.write(index.tobytes())
        self._fh.seek(0)
        self._fh.write(
            _HEADER.pack(MAGIC, iloc, len(index), nloc, nalloc, 0x10004, 0x10000, self._application.encode(), self._schema.encode())
        )
        self._fh.close()
        self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

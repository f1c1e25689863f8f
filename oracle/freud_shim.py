"""Restatement of the two ``freud`` (pin ``freud-analysis>=1.0,<1.3``, reference ``setup.py:40``)
facilities on the hot path: ``freud.box.Box.wrap`` and ``freud.locality.NearestNeighbors`` with
``strict_cut=True``.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

freud is C++/Cython and is not installed here; the published algorithm is restated [recalled].
Reference call sites: ``Box(...)`` ``util.py:178-193``; ``box.wrap`` ``dynamics/_util.py:149``;
``NearestNeighbors(rmax, k, strict_cut=True).compute / getNeighborList / nlist.neighbor_counts /
r_sq_list`` ``order.py:181-183,207,281,303-305``.

All arithmetic is IEEE binary32 with one rounding per operation (no fused multiply-add): freud's
wheels are built for baseline x86-64, which has no FMA.
"""

import numpy as np

F32 = np.float32
UINT_MAX = np.uint32(2 ** 32 - 1)


class Box:
    """freud 1.x ``Box``: lengths, tilt factors, 2D flag; ``lo = -L/2``."""

    def __init__(self, Lx, Ly, Lz=0.0, xy=0.0, xz=0.0, yz=0.0, is2D=False):
        if is2D:
            Lz = 0.0
        self.L = np.array([Lx, Ly, Lz], dtype=F32)
        self.xy, self.xz, self.yz = F32(xy), F32(xz), F32(yz)
        self.is2D = bool(is2D)
        self.hi = self.L / F32(2.0)
        self.lo = -self.hi

    Lx = property(lambda self: float(self.L[0]))
    Ly = property(lambda self: float(self.L[1]))
    Lz = property(lambda self: float(self.L[2]))

    @property
    def volume(self):
        return self.Lx * self.Ly if self.is2D else self.Lx * self.Ly * self.Lz

    def make_fraction(self, v):
        """``Box::makeFraction``: (v - lo - tilt terms) / L, z forced to 0 in 2D."""
        x, y, z = v[:, 0], v[:, 1], v[:, 2]
        dx = x - self.lo[0]
        dy = y - self.lo[1]
        dz = z - self.lo[2]
        dx = dx - ((self.xz - self.yz * self.xy) * z + self.xy * y)
        dy = dy - self.yz * z
        with np.errstate(divide="ignore", invalid="ignore"):
            f = np.stack([dx / self.L[0], dy / self.L[1], dz / self.L[2]], axis=1)
        if self.is2D:
            f[:, 2] = 0
        return f.astype(F32, copy=False)

    def make_coordinates(self, f):
        """``Box::makeCoordinates``: lo + f*L, then shear back; z forced to 0 in 2D."""
        v = self.lo[np.newaxis, :] + f * self.L[np.newaxis, :]
        v = v.astype(F32, copy=False)
        v[:, 0] = v[:, 0] + (self.xy * v[:, 1] + self.xz * v[:, 2])
        v[:, 1] = v[:, 1] + self.yz * v[:, 2]
        if self.is2D:
            v[:, 2] = 0
        return v

    def wrap(self, vecs):
        """``Box::wrap``: fraction -> fmodf(.,1) (+1 if negative) -> coordinates, all in f32.

        The python binding first converts its argument to a contiguous f32 copy, so f64 input is
        rounded to f32 before any arithmetic.  The result lies in [-L/2, L/2] (the upper edge is
        reachable when a tiny negative fraction rounds to 1 after the +1).
        """
        v = np.array(vecs, dtype=F32, order="C", copy=True)
        single = v.ndim == 1
        v = v.reshape(-1, 3)
        with np.errstate(invalid="ignore"):
            f = self.make_fraction(v)
            f = np.fmod(f, F32(1.0))
            f = np.where(f < 0, f + F32(1.0), f).astype(F32)
            out = self.make_coordinates(f)
        return out[0] if single else out


class NeighborListView:
    def __init__(self, counts):
        self.neighbor_counts = counts


class NearestNeighbors:
    """freud 1.x ``NearestNeighbors(rmax, n_neigh, strict_cut=True)`` on one point set.

    For every point: all other points with ``rsq < rmax*rmax`` (strict, f32; ``rsq`` is the f32 dot
    product of ``box.wrap(r_j - r_i)``), sorted ascending by ``(rsq, index)``, truncated to
    ``n_neigh`` and padded with ``UINT_MAX`` (``r_sq_list`` padded with -1).  The padding values are
    confirmed by reference ``order.py:200-203,306-308``; the tie order is this oracle's definition.
    """

    def __init__(self, rmax, n_neigh, strict_cut=True):
        if not strict_cut:
            raise NotImplementedError("reference only uses strict_cut=True (order.py:181)")
        self.rmax = F32(rmax)
        self.k = int(n_neigh)

    def compute(self, box, points, chunk=2048):
        pts = np.ascontiguousarray(points, dtype=F32).reshape(-1, 3)
        n = len(pts)
        k = self.k
        rmaxsq = F32(self.rmax * self.rmax)
        nlist = np.full((n, k), UINT_MAX, dtype=np.uint32)
        rsq_out = np.full((n, k), F32(-1.0), dtype=F32)
        counts = np.zeros(n, dtype=np.uint32)
        cand = self._candidates(box, pts)
        for i0 in range(0, n, chunk):
            i1 = min(n, i0 + chunk)
            for i in range(i0, i1):
                js = cand(i)
                if len(js) == 0:
                    continue
                d = box.wrap(pts[js] - pts[i])
                # dot(rij, rij): ((x*x + y*y) + z*z), one f32 rounding per operation
                rsq = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
                keep = rsq < rmaxsq
                js, rsq = js[keep], rsq[keep]
                order = np.lexsort((js, rsq))[:k]
                m = len(order)
                nlist[i, :m] = js[order]
                rsq_out[i, :m] = rsq[order]
                counts[i] = m
        self._nlist, self._rsq = nlist, rsq_out
        self.nlist = NeighborListView(counts)
        return self

    def _candidates(self, box, pts):
        """Superset of the true neighbours of each point (exact filter is applied by the caller)."""
        n = len(pts)
        ortho = box.xy == 0 and box.xz == 0 and box.yz == 0
        if n > 512 and ortho:
            try:
                from scipy.spatial import cKDTree

                dims = 2 if box.is2D else 3
                L = box.L[:dims].astype(np.float64)
                p = np.mod(pts[:, :dims].astype(np.float64) + L / 2, L)
                p = np.where(p >= L, 0.0, p)
                tree = cKDTree(p, boxsize=L)
                lists = tree.query_ball_point(p, float(self.rmax) * (1 + 1e-4) + 1e-4)

                def cand(i):
                    js = np.asarray(lists[i], dtype=np.int64)
                    return js[js != i]

                return cand
            except ImportError:
                pass
        everyone = np.arange(n, dtype=np.int64)

        def cand_all(i):
            return everyone[everyone != i]

        return cand_all

    def getNeighborList(self):
        return self._nlist

    @property
    def r_sq_list(self):
        return self._rsq


class RDF:
    """freud 1.x ``freud.density.RDF(rmax, dr)`` on one point set (reference call site
    ``dynamics/_util.py:78-79``; consumers ``_util.py:57-63``: ``.R``, ``.RDF``, ``.box``) [recalled].

    ``nbins = int(floorf(rmax / dr))``; bin ``i`` spans ``[i dr, (i+1) dr)``; ``R[i]`` is the centre of
    mass of that shell, ``2/3 (r1^3 - r0^3) / (r1^2 - r0^2)``; shell volume ``pi (r1^2 - r0^2)`` in 2D,
    ``4/3 pi (r1^3 - r0^3)`` in 3D.  Every ordered pair ``i != j`` with ``rsq < rmax^2``
    (``rsq`` = f32 dot product of ``box.wrap(r_j - r_i)``) adds one count to bin
    ``floorf(sqrtf(rsq) / dr)``.  ``RDF[i] = counts[i] / N / volume[i] / (N / box.volume)``, all f32.
    """

    def __init__(self, rmax, dr):
        self.rmax, self.dr = F32(rmax), F32(dr)
        self.nbins = int(np.floor(self.rmax / self.dr))
        i = np.arange(self.nbins, dtype=F32)
        r0 = i * self.dr
        r1 = (i + F32(1.0)) * self.dr
        self.R = (F32(2.0) / F32(3.0) * (r1 * r1 * r1 - r0 * r0 * r0) / (r1 * r1 - r0 * r0)).astype(F32)
        self._r0, self._r1 = r0.astype(F32), r1.astype(F32)

    def pair_counts(self, box, points, chunk=256):
        pts = np.ascontiguousarray(points, dtype=F32).reshape(-1, 3)
        n = len(pts)
        counts = np.zeros(self.nbins, dtype=np.uint64)
        rmaxsq = F32(self.rmax * self.rmax)
        dr_inv = F32(1.0) / self.dr
        for i0 in range(0, n, chunk):
            i1 = min(n, i0 + chunk)
            d = (pts[np.newaxis, :, :] - pts[i0:i1, np.newaxis, :]).reshape(-1, 3)
            w = box.wrap(d)
            rsq = (w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1]) + w[:, 2] * w[:, 2]
            keep = rsq < rmaxsq
            keep.reshape(i1 - i0, n)[np.arange(i1 - i0), np.arange(i0, i1)] = False  # i != j
            bins = np.floor(np.sqrt(rsq[keep]) * dr_inv).astype(np.int64)
            bins = bins[bins < self.nbins]
            counts += np.bincount(bins, minlength=self.nbins).astype(np.uint64)
        return counts

    def compute(self, box, points):
        pts = np.ascontiguousarray(points, dtype=F32).reshape(-1, 3)
        self.box = box
        self.counts = self.pair_counts(box, pts)
        self.RDF = normalise_rdf(self.counts, len(pts), box, self._r0, self._r1)
        return self


def normalise_rdf(counts, n, box, r0, r1):
    """``RDF::reduceRDF`` [recalled]: f32 throughout."""
    if box.is2D:
        vol = F32(np.pi) * (r1 * r1 - r0 * r0)
    else:
        vol = F32(4.0) / F32(3.0) * F32(np.pi) * (r1 * r1 * r1 - r0 * r0 * r0)
    ndens = F32(n) / F32(box.volume)
    avg = counts.astype(F32) / F32(n)
    return (avg / vol.astype(F32) / ndens).astype(F32)


class Voronoi:
    """freud 1.x ``freud.voronoi.Voronoi(box, buff).computeNeighbors(positions)`` in 2D (reference call
    site ``order.py:312-315``; only ``nlist.neighbor_counts`` is consumed) [recalled]: the points are
    extended by the periodic images that lie within ``buff`` of the box, ``scipy.spatial.Voronoi`` (qhull)
    is run on the extended set, and every finite ridge that touches a real point makes the two generating
    points (images mapped back to their originals) neighbours of each other; a pair is counted once."""

    def __init__(self, box, buff=0.1):
        self.box, self.buff = box, float(buff)

    def computeNeighbors(self, positions):
        from scipy.spatial import Voronoi as QhullVoronoi

        pts = np.asarray(positions, dtype=np.float64)[:, :2]
        n = len(pts)
        L = np.array([self.box.Lx, self.box.Ly], dtype=np.float64)
        buff = min(self.buff, 0.5 * float(L.min()))
        expanded, ids = [pts], [np.arange(n)]
        for sx in (-1, 0, 1):
            for sy in (-1, 0, 1):
                if sx == 0 and sy == 0:
                    continue
                img = pts + np.array([sx, sy]) * L
                keep = np.all((img > -L / 2 - buff) & (img < L / 2 + buff), axis=1)
                expanded.append(img[keep])
                ids.append(np.nonzero(keep)[0])
        expanded, ids = np.concatenate(expanded), np.concatenate(ids)
        vor = QhullVoronoi(expanded)
        pairs = set()
        self.ridge_lengths = {}
        for (p, q), verts in zip(vor.ridge_points, vor.ridge_vertices):
            if (p >= n and q >= n) or -1 in verts:
                continue
            i, j = int(ids[p]), int(ids[q])
            if i == j:
                continue
            key = (min(i, j), max(i, j))
            pairs.add(key)
            length = float(np.linalg.norm(vor.vertices[verts[0]] - vor.vertices[verts[1]]))
            self.ridge_lengths[key] = min(length, self.ridge_lengths.get(key, np.inf))
        counts = np.zeros(n, dtype=np.uint32)
        for i, j in pairs:
            counts[i] += 1
            counts[j] += 1
        self.nlist = NeighborListView(counts)
        self.pairs = pairs
        return self

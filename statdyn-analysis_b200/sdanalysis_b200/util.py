"""Small host-side helpers mirroring the pieces of reference ``sdanalysis/util.py`` the hot path uses.

``Box`` stands in for ``freud.box.Box`` as a plain parameter holder (the arithmetic of ``Box.wrap`` lives
in the CUDA kernels); ``create_freud_box`` keeps the reference's name and argument meaning
(``util.py:178-193``) so call sites read the same.
"""

from pathlib import Path
from typing import NamedTuple, Optional

import numpy as np


class Box:
    """Simulation cell: lengths, tilt factors and the 2D flag (``freud.box.Box`` constructor order)."""

    def __init__(self, Lx, Ly, Lz=0.0, xy=0.0, xz=0.0, yz=0.0, is2D=False):
        self.is2D = bool(is2D)
        self.Lx, self.Ly = float(Lx), float(Ly)
        self.Lz = 0.0 if self.is2D else float(Lz)
        self.xy = float(xy)
        self.xz = 0.0 if self.is2D else float(xz)
        self.yz = 0.0 if self.is2D else float(yz)

    def to_array(self):
        return np.array([self.Lx, self.Ly, self.Lz, self.xy, self.xz, self.yz], dtype=np.float32)

    @property
    def volume(self):
        return self.Lx * self.Ly if self.is2D else self.Lx * self.Ly * self.Lz

    def __repr__(self):
        return f"Box(Lx={self.Lx}, Ly={self.Ly}, Lz={self.Lz}, xy={self.xy}, xz={self.xz}, yz={self.yz}, is2D={self.is2D})"


def create_freud_box(box, is_2D=True) -> Box:
    """Array of box values -> :class:`Box` (reference ``util.py:178-193``)."""
    Lx, Ly, Lz = box[:3]
    xy = xz = yz = 0
    if len(box) == 6:
        xy, xz, yz = box[3:6]
    if is_2D:
        return Box(Lx=Lx, Ly=Ly, xy=xy, is2D=True)
    return Box(Lx=Lx, Ly=Ly, Lz=Lz, xy=xy, xz=xz, yz=yz)


def as_box(box, is_2D=None) -> Box:
    if isinstance(box, Box):
        return box
    return create_freud_box(np.asarray(box, dtype=np.float64), True if is_2D is None else is_2D)


def zero_quaternion(num: int) -> np.ndarray:
    """Identity quaternions, float64 like the reference (``util.py:27-30``)."""
    q = np.zeros((num, 4))
    q[:, 0] = 1
    return q


def z2quaternion(theta) -> np.ndarray:
    """Rotation about z -> quaternion (w, 0, 0, z), float32 (reference ``util.py:150-157``)."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    q = np.zeros(theta.shape + (4,), dtype=np.float64)
    q[..., 0] = np.cos(theta / 2)
    q[..., 3] = np.sin(theta / 2)
    return q.astype(np.float32)


def _qmul(a, b):
    """Hamilton product of quaternion arrays (w, x, y, z), broadcasting."""
    aw, ax, ay, az = np.moveaxis(np.asarray(a), -1, 0)
    bw, bx, by, bz = np.moveaxis(np.asarray(b), -1, 0)
    return np.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        axis=-1,
    )


def _qconj(q):
    q = np.array(q, copy=True)
    q[..., 1:] *= -1
    return q


def rotate_vectors(quaternion, vector):
    """``rowan.rotate``: q (0, v) q* (reference ``util.py:142-143``).  Host-side helper for a handful of
    vectors; the per-site rotations of ``struct`` run in ``csrc/struct.cu``."""
    quaternion = np.asarray(quaternion)
    vector = np.asarray(vector)
    v4 = np.concatenate([np.zeros(vector.shape[:-1] + (1,), dtype=vector.dtype), vector], axis=-1)
    return _qmul(quaternion, _qmul(v4, _qconj(quaternion)))[..., 1:]


def quaternion_rotation(initial, final):
    """``rowan.geometry.intrinsic_distance``: |log(initial* final)| in [0, pi] (``util.py:138-139``)."""
    p = _qmul(_qconj(np.asarray(initial, dtype=np.float64)), np.asarray(final, dtype=np.float64))
    norm = np.linalg.norm(p, axis=-1)
    vec = np.linalg.norm(p[..., 1:], axis=-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ang = np.where(vec > 0, np.arccos(np.clip(p[..., 0] / norm, -1.0, 1.0)), 0.0)
        return np.sqrt(np.log(norm) ** 2 + ang ** 2)


def quaternion_angle(quaternion) -> np.ndarray:
    """``rowan.geometry.angle``: 2 acos(w) (``util.py:146-147``)."""
    return 2 * np.arccos(np.asarray(quaternion)[..., 0])


def quaternion2z(quaternion: np.ndarray) -> np.ndarray:
    """z-angle of ``rowan.to_euler`` (convention zyx, intrinsic), float32 (``util.py:160-167``)."""
    q = np.asarray(quaternion, dtype=np.float64)
    w, x, y, z = np.moveaxis(q, -1, 0)
    s = 1.0 / np.sqrt(w * w + x * x + y * y + z * z)  # rowan.to_matrix scales by 1 / |q|
    m10 = 2 * s * (x * y + z * w)
    m00 = 1 - 2 * s * (y * y + z * z)
    return np.arctan2(np.clip(m10, -1, 1), np.clip(m00, -1, 1)).astype(np.float32)


def orientation2positions(mol, position: np.ndarray, orientation: np.ndarray) -> np.ndarray:
    """Site positions, site-major (reference ``util.py:170-175``)."""
    return np.tile(position, (mol.num_particles, 1)) + np.concatenate(
        [rotate_vectors(orientation, pos) for pos in mol.positions]
    )


class Variables(NamedTuple):
    """Simulation variables encoded in a file name (reference ``util.py:33-88``)."""

    temperature: Optional[str]
    pressure: Optional[str]
    crystal: Optional[str]
    iteration_id: Optional[str]

    @classmethod
    def from_filename(cls, fname) -> "Variables":
        parts = Path(fname).stem.split("-")
        if parts and parts[0] in ("dump", "trajectory", "thermo"):
            parts = parts[1:]
        if len(parts) < 3:
            return cls(None, None, None, None)
        found = {"temperature": None, "pressure": None, "crystal": None, "iteration_id": None}
        for item in parts:
            if item.startswith("P"):
                found["pressure"] = item[1:]
            elif item.startswith("T"):
                found["temperature"] = item[1:]
            elif item.startswith("ID"):
                found["iteration_id"] = item[2:]
            else:
                found["crystal"] = item
        return cls(**found)


def get_filename_vars(fname) -> Variables:
    return Variables.from_filename(fname)

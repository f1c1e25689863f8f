"""``Dynamics``: per-origin dynamic quantities, drop-in for reference ``dynamics/dynamics.py``.

Same constructor, methods, result keys and error behaviour as the reference class; the arithmetic
runs in the CUDA kernels behind ``libsdb200.so``.  Every ``compute_*`` returns a host scalar (or
array) like the reference; the batched, asynchronous path used by ``read.process_file`` is
``engine.DynamicsEngine``.
"""

import logging
from typing import Dict, Optional, Union

import numpy as np

from .. import _lib
from ..engine import ALL_QUANTITIES
from ..util import create_freud_box
from ._util import TrackedMotion, overlap_count

logger = logging.getLogger(__name__)


class Dynamics:
    """Compute dynamic properties of a simulation relative to one reference configuration.

    Args mirror reference ``dynamics.py:67-76``: ``timestep``, ``box`` (3 or 6 values), ``position``
    (N,3), ``orientation`` (N,4) quaternions (w,x,y,z) or None, ``molecule`` (site table for
    ``struct``; None means a 2D box and no ``struct``), ``wave_number``.
    """

    _all_quantities = list(ALL_QUANTITIES)

    def __init__(
        self,
        timestep: int,
        box: np.ndarray,
        position: np.ndarray,
        orientation: Optional[np.ndarray] = None,
        molecule=None,
        wave_number: Optional[float] = None,
        angular_resolution=360,
        struct_mode: str = "reference",
    ) -> None:
        position = np.asarray(position)
        if position.shape[0] == 0:
            raise RuntimeError("Position must contain values, has length of 0.")
        if molecule is None:
            is2d = True
            self.mol_vector = None
        else:
            is2d = molecule.dimensions == 2
            self.mol_vector = molecule.positions
        self.timestep = timestep
        self.num_particles = position.shape[0]
        self.wave_number = wave_number
        self.angular_resolution = int(angular_resolution)
        options = dict(wave_number=wave_number, mol_sites=self.mol_vector, struct_mode=struct_mode)
        self.motion = TrackedMotion(
            create_freud_box(np.asarray(box, dtype=np.float64), is_2D=is2d),
            position,
            orientation,
            timestep=timestep,
            engine_options=options,
        )
        self._row = None  # cached quantities of the current state
        if wave_number is not None:
            angles = np.linspace(0, 2 * np.pi, num=angular_resolution, endpoint=False).reshape((-1, 1))
            self.wave_vector = np.concatenate([np.cos(angles), np.sin(angles)], axis=1) * wave_number

    @classmethod
    def from_frame(cls, frame, molecule=None, wave_number: Optional[float] = None) -> "Dynamics":
        """Reference ``dynamics.py:114-134``."""
        return cls(
            frame.timestep, frame.box, frame.position, frame.orientation, molecule=molecule, wave_number=wave_number
        )

    @property
    def delta_translation(self):
        return self.motion.delta_translation

    @property
    def delta_rotation(self):
        return self.motion.delta_rotation

    @property
    def distance(self) -> Optional[float]:
        if self.wave_number is None:
            return None
        return np.pi / (2 * self.wave_number)

    def add(self, position: np.ndarray, orientation: Optional[np.ndarray] = None):
        """Add the next sample (reference ``dynamics.py:144-160``)."""
        self._row = None
        self.motion.add(position, orientation)

    def add_frame(self, frame):
        self.add(frame.position, frame.orientation)

    # ---- per-molecule arrays --------------------------------------------------------------------
    def compute_displacement(self) -> np.ndarray:
        return self.motion.norms()[0]

    def compute_displacement2(self) -> np.ndarray:
        return np.square(self.delta_translation).sum(axis=1)

    def compute_rotation(self) -> np.ndarray:
        return self.motion.norms()[1]

    def compute_rotation2(self) -> np.ndarray:
        return np.square(self.delta_rotation).sum(axis=1)

    # ---- aggregates (all from one device reduction of the current state) ---------------------------
    def _current(self):
        if self._row is None:
            self._row = self.motion.current_row()
        return self._row

    def compute_msd(self) -> float:
        return self._current()["msd"]

    def compute_mfd(self) -> float:
        return self._current()["mfd"]

    def compute_alpha(self) -> float:
        return self._current()["alpha"]

    def compute_time_delta(self, timestep: int) -> int:
        return timestep - self.timestep

    def compute_mean_rotation(self) -> float:
        return self._current()["mean_rotation"]

    def compute_msr(self) -> float:
        return self._current()["msr"]

    def compute_isf(self) -> float:
        """Intermediate scattering function (reference ``dynamics.py:233-235``), ``csrc/isf.cu``."""
        return self.motion.isf(self.wave_number, self.angular_resolution)

    def compute_rotational_relax1(self) -> float:
        return self._current()["rot1"]

    def compute_rotational_relax2(self) -> float:
        return self._current()["rot2"]

    def compute_alpha_rot(self) -> float:
        return self._current()["alpha_rot"]

    def compute_gamma(self) -> float:
        return self._current()["gamma"]

    def compute_struct_relax(self) -> float:
        if self.distance is None:
            raise ValueError("The wave number is required for the structural relaxation.")
        if not self.motion.is_planar:
            raise NotImplementedError("struct is only implemented for planar (2D, rotation about z) motion")
        return self._current()["struct"]

    def compute_all(
        self,
        timestep: int,
        position: np.ndarray,
        orientation: np.ndarray = None,
        scattering_function: bool = False,
    ) -> Dict[str, Union[int, float]]:
        """Add a sample and return all 15 quantities (reference ``dynamics.py:328-394``)."""
        self._row = None
        result = self.motion.add(position, orientation, timestep=timestep, want_sums=True)
        if result is not None:
            row = result.rows()[0]
        else:
            row = self.motion.current_row()
        row["time"] = self.compute_time_delta(timestep)
        if self.mol_vector is None or self.motion.previous_orientation is None:
            row["struct"] = np.nan
        if scattering_function and self.wave_number is not None:
            row["scattering_function"] = self.compute_isf()
        self._row = row
        return dict(row)

    def __len__(self) -> int:
        return self.num_particles

    def get_molid(self):
        return np.arange(self.num_particles)


def structural_relax(displacement: np.ndarray, dist: float = 0.3) -> float:
    """Fraction of ``displacement < dist`` (reference ``dynamics.py:410-428``), reduced on the device."""
    torch = _lib.torch()
    device = _lib.require_cuda()
    d = torch.from_numpy(np.ascontiguousarray(displacement)).to(device)
    if d.numel() == 0:
        return float("nan")
    # comparison in the array's dtype with the scalar rounded to it, like numpy with a python float
    thr = torch.tensor(dist, dtype=d.dtype, device=device)
    return float((d < thr).double().mean().item())


def mobile_overlap(displacement: np.ndarray, rotation: np.ndarray, fraction: float = 0.1) -> float:
    """Overlap of the most mobile translators and rotators (reference ``dynamics.py:431-445``).

    Exact device selection (``csrc/overlap.cu``); ties are broken towards the larger index.
    Displacements must be non-negative (they are norms).
    """
    displacement = np.ascontiguousarray(displacement, dtype=np.float32)
    rotation = np.ascontiguousarray(rotation, dtype=np.float32)
    n = len(displacement)
    k = int(n * fraction)
    if k == 0:
        raise ZeroDivisionError("division by zero")
    if np.any(displacement < 0):
        raise ValueError("displacement must be non-negative")
    torch = _lib.torch()
    device = _lib.require_cuda()
    stride = (n + 3) // 4 * 4
    planes = torch.zeros((3, stride), dtype=torch.float32, device=device)
    planes[0, :n] = torch.from_numpy(displacement).to(device)
    planes[2, :n] = torch.from_numpy(rotation).to(device)
    return overlap_count(planes, n, stride, k, device) / k

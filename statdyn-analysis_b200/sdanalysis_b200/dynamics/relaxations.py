"""Per-molecule first/last-passage relaxation times, drop-in for reference
``dynamics/relaxations.py``.

``Relaxations`` drives the fused kernel (one flag byte + write-only passage times per molecule);
``MolecularRelaxation`` / ``LastMolecularRelaxation`` can also be driven directly with distance
arrays like in the reference tests (``test/dynamics_test.py:191-265``) -- they then run the
stand-alone branch-free kernels of ``csrc/passage.cu`` on the reference's own state layout.
"""

import logging
from enum import Enum, auto
from typing import Dict, List, Optional

import numpy as np
import pandas

from .. import _lib
from ..engine import MAX_STATUS, TrackerSpec, default_tracker_definitions
from ..util import create_freud_box
from ._util import TrackedMotion, calculate_wave_number

logger = logging.getLogger(__name__)


class RelaxationType(Enum):
    rotation = auto()
    translation = auto()


class MolecularRelaxation:
    """First-passage time of every molecule past ``threshold`` (reference ``relaxations.py:187-221``)."""

    _max_value = MAX_STATUS
    relaxation_type = RelaxationType.translation

    def __init__(self, num_elements: int, threshold: float, relaxation_type: Optional[str] = None) -> None:
        self.num_elements = num_elements
        self.threshold = threshold
        self._max_value = MAX_STATUS
        if relaxation_type is not None:
            self.relaxation_type = RelaxationType[relaxation_type]
        self._torch = torch = _lib.torch()
        self._device = _lib.require_cuda()
        self._lib = _lib.load()
        self._d_status = torch.full((num_elements,), MAX_STATUS, dtype=torch.int64, device=self._device)

    def _upload(self, distance):
        distance = np.asarray(distance)
        if distance.shape != (self.num_elements,):
            raise RuntimeError(
                "Current state and initial state have different shapes. "
                f"current: {distance.shape}, initial: {(self.num_elements,)}"
            )
        if distance.dtype != np.float32:
            distance = distance.astype(np.float64)
        return self._torch.from_numpy(np.ascontiguousarray(distance)).to(self._device)

    def add(self, timediff: int, distance: np.ndarray) -> None:
        d = self._upload(distance)
        with self._torch.cuda.device(self._device):
            _lib.check(
                self._lib.sdb_first_passage(
                    d.data_ptr(), int(d.dtype == self._torch.float64), self.num_elements, float(self.threshold),
                    int(timediff), self._d_status.data_ptr(), _lib.stream_ptr(self._device),
                )
            )

    @property
    def _status(self):
        return self._d_status.cpu().numpy()

    def get_status(self):
        return self._status


class LastMolecularRelaxation(MolecularRelaxation):
    """Last-passage time with an irreversibility distance (reference ``relaxations.py:224-294``)."""

    _is_irreversible = 2

    def __init__(
        self, num_elements: int, threshold: float, irreversibility: float = 1.0, relaxation_type: Optional[str] = None
    ) -> None:
        super().__init__(num_elements, threshold, relaxation_type)
        self._irreversibility = irreversibility
        self._d_state = self._torch.zeros(num_elements, dtype=self._torch.uint8, device=self._device)

    def add(self, timediff: int, distance: np.ndarray) -> None:
        d = self._upload(distance)
        with self._torch.cuda.device(self._device):
            _lib.check(
                self._lib.sdb_last_passage(
                    d.data_ptr(), int(d.dtype == self._torch.float64), self.num_elements, float(self.threshold),
                    float(self._irreversibility), int(timediff), self._d_status.data_ptr(),
                    self._d_state.data_ptr(), _lib.stream_ptr(self._device),
                )
            )

    @property
    def _state(self):
        return self._d_state.cpu().numpy()

    def get_status(self):
        status = self._status
        status[self._state != 2] = self._max_value
        return status


class _FusedTracker:
    """View of one tracker of a ``Relaxations`` object whose state lives in the fused engine."""

    _max_value = MAX_STATUS
    _is_irreversible = 2

    def __init__(self, owner, index, spec: TrackerSpec):
        self._owner, self._index, self._spec = owner, index, spec
        self.num_elements = owner._num_elements
        self.threshold = spec.threshold
        self.relaxation_type = RelaxationType[spec.type]
        if spec.last_passage:
            self._irreversibility = spec.cutoff

    @property
    def _status(self):
        status, _ = self._owner._tracker_state()
        return status[self._index].astype(np.int64)

    @property
    def _state(self):
        _, flags = self._owner._tracker_state()
        return ((flags >> self._spec.bit) & 3).astype(np.uint8)

    def get_status(self):
        status = self._status
        if self._spec.last_passage:
            status[self._state != 2] = MAX_STATUS
        return status

    def add(self, timediff, distance):
        raise RuntimeError("this tracker is advanced by its Relaxations object (fused on the device)")


class StructRelaxations(MolecularRelaxation):
    """First passage of every *site* of every molecule, reported as the per-molecule mean
    (reference ``relaxations.py:297-305``; unused by the reference's driver)."""

    def __init__(self, num_elements: int, threshold: float, molecule) -> None:
        self.molecule = molecule
        super().__init__(num_elements * self.molecule.num_particles, threshold)

    def get_status(self):
        return self._status.reshape((-1, self.molecule.num_particles)).mean(axis=1)


def create_mol_relaxations(
    num_elements: int,
    threshold: float,
    last_passage: bool = False,
    last_passage_cutoff: Optional[float] = None,
    relaxation_type: Optional[str] = None,
) -> MolecularRelaxation:
    """Reference ``relaxations.py:308-335``."""
    if threshold is None or threshold < 0:
        raise ValueError(f"Threshold needs a positive value, got {threshold}")
    if last_passage:
        if last_passage_cutoff is None:
            last_passage_cutoff = 3 * threshold
        elif last_passage_cutoff < 0:
            raise ValueError(
                f"When using last passage a positive cutoff value is required, got {last_passage_cutoff}, default is 1.0"
            )
        return LastMolecularRelaxation(num_elements, threshold, last_passage_cutoff, relaxation_type)
    return MolecularRelaxation(num_elements, threshold, relaxation_type)


class Relaxations:
    """Molecular relaxation times of one time-origin (reference ``relaxations.py:40-184``)."""

    def __init__(
        self,
        timestep: int,
        box: np.ndarray,
        position: np.ndarray,
        orientation: np.ndarray,
        molecule=None,
        is2D: Optional[bool] = None,
        wave_number: Optional[float] = None,
    ) -> None:
        self.init_time = timestep
        is2D = False if molecule is None else molecule.dimensions == 2  # relaxations.py:52-55 (trap T7)
        position = np.asarray(position)
        self._num_elements = position.shape[0]
        if wave_number is None:  # relaxations.py:61-62
            wave_number = calculate_wave_number(create_freud_box(np.asarray(box, dtype=np.float64), is_2D=is2D), position)
        self.wave_number = wave_number
        self._definition = default_tracker_definitions(self.distance)
        self.motion = TrackedMotion(
            create_freud_box(np.asarray(box, dtype=np.float64), is_2D=is2D),
            position,
            orientation,
            timestep=timestep,
            engine_options=dict(wave_number=wave_number, trackers=self._definition, compute_sums=False),
        )
        self.mol_relax: Dict[str, MolecularRelaxation] = {}
        self.set_mol_relax(self._definition)

    @property
    def distance(self):
        return np.pi / (2 * self.wave_number)

    @classmethod
    def from_frame(cls, frame, molecule=None, wave_number: Optional[float] = None) -> "Relaxations":
        return cls(
            frame.timestep, frame.box, frame.position, frame.orientation, molecule=molecule, wave_number=wave_number
        )

    def set_mol_relax(self, definition: List[Dict]) -> None:
        """Replace the trackers; their passage state starts afresh (reference ``relaxations.py:108-133``)."""
        specs = [TrackerSpec(item) for item in definition]
        self._definition = list(definition)
        self._cache = None
        self.mol_relax = {}
        if self.motion.is_planar:
            self.motion._engine.set_trackers(definition)
            for i, spec in enumerate(self.motion._engine.trackers):
                self.mol_relax[spec.name] = _FusedTracker(self, i, spec)
        else:
            for spec in specs:
                self.mol_relax[spec.name] = create_mol_relaxations(
                    self._num_elements, spec.threshold, spec.last_passage, spec.cutoff, spec.type
                )

    def get_timediff(self, timestep: int):
        return timestep - self.init_time

    def _tracker_state(self):
        if self._cache is None:
            self._cache = self.motion._engine.tracker_state(0)
        return self._cache

    def add(self, timestep: int, position: np.ndarray, orientation: np.ndarray) -> None:
        """Add a sample and advance every tracker (reference ``relaxations.py:135-162``)."""
        self._cache = None
        was_planar = self.motion.is_planar
        self.motion.add(position, orientation, timestep=timestep)
        if was_planar and self.motion.is_planar:
            return  # trackers advanced inside the fused kernel
        if was_planar:
            raise NotImplementedError("switching a Relaxations object from planar to general motion")
        displacement, rotation = self.motion.norms()
        for func in self.mol_relax.values():
            if func.relaxation_type is RelaxationType.rotation:
                func.add(self.get_timediff(timestep), rotation)
            elif func.relaxation_type is RelaxationType.translation:
                func.add(self.get_timediff(timestep), displacement)
            else:
                raise RuntimeError("Invalid relaxation type")

    def add_frame(self, frame):
        self.add(frame.timestep, frame.position, frame.orientation)

    def summary(self) -> pandas.DataFrame:
        return pandas.DataFrame({key: func.get_status() for key, func in self.mol_relax.items()})

"""CPU restatement of ``sdanalysis.dynamics`` (reference ``src/sdanalysis/dynamics/``).
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``): never imported by the product path.

Each function cites the reference lines it follows.  Semantics are those of the numpy the
reference pins (``numpy>=1.14,<1.19``, ``setup.py:29``), restated explicitly where numpy 2.x behaves
differently (SURVEY F3 / parity trap T3).
"""

import numpy as np

from . import rowan_shim as rowan
from .freud_shim import Box

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

# molecules.py:136-152 -- Trimer(): three sites at (0,0,0), (-/+ d sin60, d cos60, 0), d = 1,
# recentred on their mean by Molecule.__attrs_post_init__ (molecules.py:37-39).
_rad = np.deg2rad(120)
TRIMER_SITES = np.array(
    [[0, 0, 0], [-np.sin(_rad / 2), np.cos(_rad / 2), 0], [np.sin(_rad / 2), np.cos(_rad / 2), 0]]
)
TRIMER_SITES = TRIMER_SITES - TRIMER_SITES.mean(axis=0)


def create_freud_box(box, is_2D=True):
    """util.py:178-193."""
    Lx, Ly, Lz = box[:3]
    xy = xz = yz = 0
    if len(box) == 6:
        xy, xz, yz = box[3:6]
    if is_2D:
        return Box(Lx=Lx, Ly=Ly, xy=xy, is2D=True)
    return Box(Lx=Lx, Ly=Ly, Lz=Lz, xy=xy, xz=xz, yz=yz)


def zero_quaternion(num):
    """util.py:27-30 -- identity quaternions, float64."""
    q = np.zeros((num, 4))
    q[:, 0] = 1
    return q


class TrackedMotion:
    """dynamics/_util.py:89-156."""

    def __init__(self, box, position, orientation):
        self.box = box
        self.previous_position = position
        self.delta_translation = np.zeros_like(position)
        self.delta_rotation = np.zeros_like(position)
        if orientation is not None:
            if orientation.shape[0] == 0:
                raise RuntimeError("Orientation must contain values, has length of 0.")
            self.previous_orientation = orientation
        else:
            self.previous_orientation = None

    def add(self, position, orientation):
        # _util.py:149 -- in-place subtract of freud's f32 minimum-image vector
        self.delta_translation -= self.box.wrap(self.previous_position - position)
        if self.previous_orientation is not None and orientation is not None:
            # _util.py:150-153 -- f64 Euler angles added in place onto the accumulator: numpy
            # evaluates the sum in f64 and rounds once to the accumulator's dtype.
            euler = rowan.to_euler(rowan.divide(orientation, self.previous_orientation))
            np.add(self.delta_rotation, euler, out=self.delta_rotation, casting="same_kind")
        self.previous_position = position
        self.previous_orientation = orientation


def molecule2particles(position, orientation, mol_vector):
    """dynamics/_util.py:42-54, including the site-major / molecule-major mismatch (trap T2)."""
    if np.allclose(orientation, 0.0):
        orientation[:, 0] = 1.0
    rotated = np.concatenate(
        [rowan.rotate(orientation, site) for site in mol_vector.astype(F32)]
    )
    translated = np.repeat(position, mol_vector.shape[0], axis=0)
    return rotated + translated


def _ratio_minus_one(numerator_fn):
    """dynamics.py:195-211 / 261-303: with ``np.seterr(invalid="raise")`` a 0/0 raises, the except
    branch recomputes with the error ignored and returns the NaN untouched (``nan_to_num(copy=False)``
    on a scalar does not modify it under the pinned numpy).  x/0 with x != 0 raises
    FloatingPointError out of the except branch, as in the reference."""
    try:
        with np.errstate(divide="raise", invalid="raise", over="raise"):
            return numerator_fn()
    except FloatingPointError:
        with np.errstate(divide="raise", invalid="ignore", over="raise"):
            return numerator_fn()


def structural_relax(displacement, dist=0.3):
    """dynamics.py:410-428."""
    with np.errstate(invalid="ignore"):
        return np.mean(displacement < dist)


def mobile_overlap(displacement, rotation, fraction=0.1):
    """dynamics.py:431-445 (k = 0 raises ZeroDivisionError, trap T4)."""
    k = int(len(displacement) * fraction)
    trans_order = np.argsort(displacement)[-k:]
    rot_order = np.argsort(np.abs(rotation))[-k:]
    return len(np.intersect1d(trans_order, rot_order)) / k


class Dynamics:
    """dynamics/dynamics.py:27-407."""

    _all_quantities = list(ALL_QUANTITIES)

    def __init__(
        self,
        timestep,
        box,
        position,
        orientation=None,
        molecule=None,
        wave_number=None,
        angular_resolution=360,
    ):
        if position.shape[0] == 0:
            raise RuntimeError("Position must contain values, has length of 0.")
        # dynamics.py:92-101; ``molecule`` is either None or the string "trimer" (only Trimer is
        # ever passed on the path, read/_read.py:138-139).
        if molecule is None:
            is2d, self.mol_vector = True, None
        else:
            is2d, self.mol_vector = True, TRIMER_SITES
        self.motion = TrackedMotion(create_freud_box(box, is_2D=is2d), position, orientation)
        self.timestep = timestep
        self.num_particles = position.shape[0]
        self.wave_number = wave_number
        if wave_number is not None:
            angles = np.linspace(0, 2 * np.pi, num=angular_resolution, endpoint=False).reshape(-1, 1)
            self.wave_vector = np.concatenate([np.cos(angles), np.sin(angles)], axis=1) * wave_number

    delta_translation = property(lambda self: self.motion.delta_translation)
    delta_rotation = property(lambda self: self.motion.delta_rotation)

    @property
    def distance(self):
        return None if self.wave_number is None else np.pi / (2 * self.wave_number)

    def add(self, position, orientation=None):
        self.motion.add(position, orientation)

    # dynamics.py:179-231
    def compute_displacement(self):
        return np.linalg.norm(self.delta_translation, axis=1)

    def compute_displacement2(self):
        return np.square(self.delta_translation).sum(axis=1)

    def compute_msd(self):
        return self.compute_displacement2().mean()

    def compute_mfd(self):
        return np.square(self.compute_displacement2()).mean()

    def compute_alpha(self):
        d2 = self.compute_displacement2()
        return _ratio_minus_one(lambda: np.square(d2).mean() / (2 * np.square(d2.mean())) - 1)

    def compute_time_delta(self, timestep):
        return timestep - self.timestep

    def compute_rotation(self):
        return np.linalg.norm(self.delta_rotation, axis=1)

    def compute_rotation2(self):
        return np.square(self.delta_rotation).sum(axis=1)

    def compute_mean_rotation(self):
        return self.compute_rotation().mean()

    def compute_msr(self):
        return self.compute_rotation2().mean()

    # dynamics.py:233-303
    def compute_isf(self):
        return np.cos(np.dot(self.wave_vector, self.delta_translation[:, :2].T)).mean()

    def compute_rotational_relax1(self):
        return np.cos(self.compute_rotation()).mean()

    def compute_rotational_relax2(self):
        return np.mean(2 * np.square(np.cos(self.compute_rotation())) - 1)

    def compute_alpha_rot(self):
        t2 = self.compute_rotation2()
        return _ratio_minus_one(lambda: np.square(t2).mean() / (3 * np.square(t2.mean())) - 1)

    def compute_gamma(self):
        t2 = self.compute_rotation2()
        d2 = self.compute_displacement2()
        return _ratio_minus_one(lambda: (d2 * t2).mean() / (d2.mean() * t2.mean()) - 1)

    def compute_struct_relax(self):
        """dynamics.py:305-326, including the axis-order trap T1 (z-angle fed as gamma)."""
        if self.distance is None:
            raise ValueError("The wave number is required for the structural relaxation.")
        dr = self.delta_rotation
        quat_rot = rowan.from_euler(dr[:, 2], dr[:, 1], dr[:, 0])
        final_pos = molecule2particles(self.delta_translation, quat_rot, self.mol_vector)
        init_pos = molecule2particles(
            np.zeros_like(self.delta_translation), zero_quaternion(self.num_particles), self.mol_vector
        )
        with np.errstate(invalid="ignore"):
            return np.mean(np.linalg.norm(final_pos - init_pos, axis=1) < self.distance)

    def compute_all(self, timestep, position, orientation=None, scattering_function=False):
        """dynamics.py:328-394."""
        self.add(position, orientation)
        out = {key: np.nan for key in ALL_QUANTITIES}
        out["time"] = self.compute_time_delta(timestep)
        out["mean_displacement"] = np.linalg.norm(self.delta_translation, axis=1).mean()
        out["msd"] = self.compute_msd()
        out["mfd"] = self.compute_mfd()
        out["alpha"] = self.compute_alpha()
        if scattering_function and self.wave_number is not None:
            out["scattering_function"] = self.compute_isf()
        if self.distance is not None:
            out["com_struct"] = structural_relax(self.compute_displacement(), dist=self.distance)
        out["mean_rotation"] = self.compute_mean_rotation()
        out["msr"] = self.compute_msr()
        out["rot1"] = self.compute_rotational_relax1()
        out["rot2"] = self.compute_rotational_relax2()
        out["alpha_rot"] = self.compute_alpha_rot()
        out["gamma"] = self.compute_gamma()
        out["overlap"] = mobile_overlap(self.compute_displacement(), self.compute_rotation())
        if (
            self.distance is not None
            and self.mol_vector is not None
            and self.motion.previous_orientation is not None
        ):
            out["struct"] = self.compute_struct_relax()
        return out

    def __len__(self):
        return self.num_particles


class MolecularRelaxation:
    """dynamics/relaxations.py:187-221."""

    _max_value = MAX_STATUS
    relaxation_type = "translation"

    def __init__(self, num_elements, threshold, relaxation_type=None):
        self.num_elements = num_elements
        self.threshold = threshold
        self._status = np.full(num_elements, MAX_STATUS, dtype=int)
        if relaxation_type is not None:
            if relaxation_type not in ("rotation", "translation"):
                raise KeyError(relaxation_type)
            self.relaxation_type = relaxation_type

    def add(self, timediff, distance):
        if distance.shape != self._status.shape:
            raise RuntimeError("Current state and initial state have different shapes.")
        with np.errstate(invalid="ignore"):
            # python-float threshold against an f32 array compares in f32 (value-based casting in
            # the pinned numpy, weak scalars in numpy 2 -- same outcome).
            moved = np.greater(distance, self.threshold)
            moveable = np.greater(self._status, timediff)
            self._status[np.logical_and(moved, moveable)] = timediff

    def get_status(self):
        return self._status


class LastMolecularRelaxation(MolecularRelaxation):
    """dynamics/relaxations.py:224-294: three masked passes in the reference's order."""

    _is_irreversible = 2

    def __init__(self, num_elements, threshold, irreversibility=1.0, relaxation_type=None):
        super().__init__(num_elements, threshold, relaxation_type)
        self._state = np.zeros(num_elements, dtype=np.uint8)
        self._irreversibility = irreversibility

    def add(self, timediff, distance):
        if distance.shape != self._status.shape:
            raise RuntimeError("Current state and initial state have different shapes.")
        with np.errstate(invalid="ignore"):
            passed = (distance > self.threshold) & (self._state == 0)
            self._state[passed] = 1
            self._status[passed] = timediff
            below = (distance < self.threshold) & (self._state == 1)
            self._state[below] = 0
            irreversible = (distance > self._irreversibility) & (self._state == 1)
            self._state[irreversible] = 2

    def get_status(self):
        status = np.copy(self._status)
        status[self._state != 2] = MAX_STATUS
        return status


def create_mol_relaxations(
    num_elements, threshold, last_passage=False, last_passage_cutoff=None, relaxation_type=None
):
    """dynamics/relaxations.py:308-335."""
    if threshold is None or threshold < 0:
        raise ValueError(f"Threshold needs a positive value, got {threshold}")
    if last_passage:
        if last_passage_cutoff is None:
            last_passage_cutoff = 3 * threshold
        elif last_passage_cutoff < 0:
            raise ValueError("When using last passage a positive cutoff value is required")
        return LastMolecularRelaxation(num_elements, threshold, last_passage_cutoff, relaxation_type)
    return MolecularRelaxation(num_elements, threshold, relaxation_type)


def _static_structure_factor(rdf, wave_number, num_particles):
    """dynamics/_util.py:57-63.  ``wave_number`` is a numpy float64 scalar (an element of ``np.linspace``): under
    the pinned numpy (< 1.19, value-based casting) its product with the f32 array ``rdf.R`` is f32."""
    F32 = np.float32
    dr = rdf.R[1] - rdf.R[0]
    integral = dr * np.sum((rdf.RDF - F32(1)) * rdf.R * np.sin(F32(wave_number) * rdf.R))
    density = num_particles / rdf.box.volume
    return 1 + 4 * np.pi * density / float(wave_number) * float(integral)


def calculate_wave_number(box, positions):
    """dynamics/_util.py:66-86: position of the maximum of the static structure factor on
    ``linspace(0.5, 20, 200)``, from ``freud.density.RDF(rmax=min(L)/2.2, dr=rmax/200)``."""
    from .freud_shim import RDF

    rmax = min(box.Lx / 2.2, box.Ly / 2.2)
    if not box.is2D:
        rmax = min(rmax, box.Lz / 2.2)
    dr = rmax / 200
    rdf = RDF(dr=dr, rmax=rmax)
    rdf.compute(box, positions)
    x = np.linspace(0.5, 20, 200)
    ssf = [_static_structure_factor(rdf, value, len(positions)) for value in x]
    return x[np.argmax(ssf)]


class Relaxations:
    """dynamics/relaxations.py:40-184."""

    def __init__(self, timestep, box, position, orientation, molecule=None, is2D=None, wave_number=None):
        self.init_time = timestep
        is2D = molecule is not None  # relaxations.py:52-55: None -> 3D box (trap T7)
        self.motion = TrackedMotion(create_freud_box(box, is_2D=is2D), position, orientation)
        self._num_elements = position.shape[0]
        if wave_number is None:  # relaxations.py:61-62
            wave_number = calculate_wave_number(self.motion.box, position)
        self.wave_number = wave_number
        d = self.distance
        self.set_mol_relax(
            [
                {"name": "tau_D", "threshold": 3 * d},
                {"name": "tau_F", "threshold": d},
                {"name": "tau_L", "threshold": d, "last_passage": True, "last_passage_cutoff": 3 * d},
                {"name": "tau_T2", "threshold": np.pi / 2, "type": "rotation"},
                {"name": "tau_T3", "threshold": np.pi / 3, "type": "rotation"},
                {"name": "tau_T4", "threshold": np.pi / 4, "type": "rotation"},
            ]
        )

    @property
    def distance(self):
        return np.pi / (2 * self.wave_number)

    def set_mol_relax(self, definition):
        self.mol_relax = {}
        for item in definition:
            if item.get("name") is None:
                raise ValueError("'name' is a required attribute")
            if item.get("threshold") is None:
                raise ValueError("'threshold' is a required attribute")
            threshold = float(item["threshold"])
            self.mol_relax[str(item["name"])] = create_mol_relaxations(
                self._num_elements,
                threshold=threshold,
                last_passage=bool(item.get("last_passage", False)),
                last_passage_cutoff=float(item.get("last_passage_cutoff", 3.0 * threshold)),
                relaxation_type=str(item.get("type", "translation")),
            )

    def get_timediff(self, timestep):
        return timestep - self.init_time

    def add(self, timestep, position, orientation):
        self.motion.add(position, orientation)
        displacement = np.linalg.norm(self.motion.delta_translation, axis=1)
        rotation = np.linalg.norm(self.motion.delta_rotation, axis=1)
        for func in self.mol_relax.values():
            if func.relaxation_type == "rotation":
                func.add(self.get_timediff(timestep), rotation)
            else:
                func.add(self.get_timediff(timestep), displacement)

    def summary(self):
        return {key: func.get_status() for key, func in self.mol_relax.items()}

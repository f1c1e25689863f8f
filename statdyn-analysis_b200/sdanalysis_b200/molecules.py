"""Molecule definitions (site tables) used by the hot path: mirrors reference ``molecules.py``.

Only the geometry matters here: ``positions`` (sites relative to the centre of mass) feed the
``struct`` kernel and ``dimensions`` selects the 2D box (reference ``dynamics.py:92-101``).
"""

import numpy as np


class Molecule:
    """Generic molecule; sites are recentred on their mean (reference ``molecules.py:19-61``)."""

    def __init__(self, dimensions=3, particles=None, positions=None, radii=None, rigid=False, moment_inertia_scale=1):
        self.dimensions = dimensions
        self.particles = list(particles) if particles is not None else ["A"]
        positions = np.zeros((1, 3)) if positions is None else np.array(positions, dtype=np.float64)
        self.positions = positions - positions.mean(axis=0)
        self._radii = dict(radii) if radii else {}
        self._radii.setdefault("A", 1.0)
        self.rigid = rigid
        self.moment_inertia_scale = moment_inertia_scale

    @property
    def num_particles(self):
        return len(self.particles)

    def get_types(self):
        return sorted(self._radii)

    def get_radii(self):
        return np.array([self._radii[p] for p in self.particles])

    def __str__(self):
        return type(self).__name__

    def __eq__(self, other):
        return isinstance(self, type(other))

    __hash__ = object.__hash__


class Disc(Molecule):
    def __init__(self):
        super().__init__(dimensions=2, radii={"A": 0.5})


class Sphere(Molecule):
    def __init__(self):
        super().__init__(radii={"A": 0.5})


class Trimer(Molecule):
    """Three discs: a central 'A' and two 'B' at ``distance``, ``angle`` degrees apart
    (reference ``molecules.py:82-168``)."""

    def __init__(self, radius=0.637556, distance=1.0, angle=120, moment_inertia_scale=1.0):
        for name, value in (("radius", radius), ("distance", distance), ("angle", angle)):
            if not isinstance(value, (float, int)):
                raise ValueError(f"The parameter '{name}' needs to be specified as a float, got, {type(value)}")
        self.radius, self.distance, self.angle = radius, distance, angle
        half = np.deg2rad(angle) / 2
        sites = [[0, 0, 0], [-distance * np.sin(half), distance * np.cos(half), 0], [distance * np.sin(half), distance * np.cos(half), 0]]
        super().__init__(
            dimensions=2,
            particles=["A", "B", "B"],
            positions=sites,
            radii={"A": 1.0, "B": radius},
            rigid=True,
            moment_inertia_scale=moment_inertia_scale,
        )

    @property
    def rad_angle(self):
        return np.radians(self.angle)

    def __repr__(self):
        return f"Trimer(radius={self.radius}, distance={self.distance}, angle={self.angle}, moment_inertia_scale={self.moment_inertia_scale})"

    def __eq__(self, other):
        return super().__eq__(other) and (self.radius, self.distance, self.angle) == (other.radius, other.distance, other.angle)

    __hash__ = object.__hash__


class Dimer(Molecule):
    def __init__(self, radius=0.637556, distance=1.0, moment_inertia_scale=1):
        self.radius, self.distance = radius, distance
        super().__init__(
            dimensions=2,
            particles=["A", "B"],
            positions=[[0, 0, 0], [0, distance, 0]],
            radii={"A": 1.0, "B": radius},
            rigid=True,
            moment_inertia_scale=moment_inertia_scale,
        )

    def __eq__(self, other):
        return super().__eq__(other) and (self.radius, self.distance) == (other.radius, other.distance)

    __hash__ = object.__hash__

"""Frame adapters (reference ``frame.py``): the input contract of the hot path.

``position`` (N,3) float32, ``orientation`` (N,4) float32 (w,x,y,z), ``box`` [Lx,Ly,Lz,xy,xz,yz],
integer ``timestep`` (reference ``frame.py:128-205``, ``test/frame_test.py:46-72``).
"""

import numpy as np

from .util import create_freud_box


class Frame:
    """Minimal frame protocol; concrete frames provide the attributes below."""

    position: np.ndarray
    orientation: np.ndarray
    box: np.ndarray
    timestep: int
    dimensions: int

    @property
    def x_position(self):
        return self.position[:, 0]

    @property
    def y_position(self):
        return self.position[:, 1]

    @property
    def z_position(self):
        return self.position[:, 2]

    def freud_box(self):
        return create_freud_box(self.box, is_2D=self.dimensions == 2)


class HoomdFrame(Frame):
    """View of a GSD snapshot restricted to the molecule centres (first ``num_mols`` rows)."""

    def __init__(self, snapshot):
        self.frame = snapshot
        self._num_mols = snapshot.num_mols

    num_mols = property(lambda self: self._num_mols)
    position = property(lambda self: self.frame.position)
    orientation = property(lambda self: self.frame.orientation)
    image = property(lambda self: self.frame.image)
    timestep = property(lambda self: self.frame.step)
    box = property(lambda self: self.frame.box)
    dimensions = property(lambda self: self.frame.dimensions)

    def __len__(self):
        return self._num_mols


class LammpsFrame(Frame):
    """A frame of a ``.lammpstrj`` dump: a dict of per-atom float32 columns plus ``box`` (three lengths)
    and ``timestep`` (reference ``frame.py:75-125``): no orientations (identity quaternions), 2D."""

    def __init__(self, frame):
        self.frame = frame

    @property
    def position(self):
        return np.array([self.frame["x"], self.frame["y"], self.frame["z"]], dtype=np.float32).T

    @property
    def image(self):
        return None

    @property
    def orientation(self):
        orient = np.zeros((len(self), 4), dtype=np.float32)
        orient[:, 0] = 1
        return orient

    @property
    def timestep(self):
        return self.frame["timestep"]

    @property
    def box(self):
        box = np.zeros(6, dtype=np.float32)
        for index, value in enumerate(self.frame["box"]):
            if index > 5:
                break
            box[index] = value
        return box

    @property
    def dimensions(self):
        return 2

    def __len__(self):
        return len(self.frame["x"])


class ArrayFrame(Frame):
    """A frame held in plain arrays (synthetic trajectories, tests)."""

    def __init__(self, timestep, box, position, orientation=None, dimensions=2):
        self.timestep = int(timestep)
        self.box = np.asarray(box, dtype=np.float32)
        self.position = position
        self.orientation = orientation
        self.dimensions = dimensions
        self.image = None

    def __len__(self):
        return len(self.position)

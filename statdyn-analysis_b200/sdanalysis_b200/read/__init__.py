"""Drop-in for ``sdanalysis.read`` (reference ``read/__init__.py``)."""

from ._gsd import read_gsd_trajectory
from ._lammps import parse_lammpstrj, read_lammps_trajectory
from ._read import WriteCache, open_trajectory, process_file, process_frames

__all__ = [
    "process_file", "process_frames", "open_trajectory", "read_gsd_trajectory", "read_lammps_trajectory",
    "parse_lammpstrj", "WriteCache",
]

"""Drop-in for ``sdanalysis.dynamics`` (reference ``dynamics/__init__.py:12-22``)."""

from ._util import TrackedMotion
from .dynamics import Dynamics
from .relaxations import LastMolecularRelaxation, MolecularRelaxation, Relaxations

__all__ = ["Dynamics", "TrackedMotion", "LastMolecularRelaxation", "MolecularRelaxation", "Relaxations"]

"""Run an analysis on a trajectory (reference ``run_analysis.py:20-42``)."""

import logging

import numpy as np

from .gsdio import open_gsd
from .order import orientational_order

logger = logging.getLogger(__name__)


def order(infile, outfile):
    """Orientational order of each frame of a trajectory: writes ``Timestep, OrientOrder`` rows, the
    second column being the fraction of particles with an order parameter above 0.9.  Like the reference
    it works on every particle record of the snapshot (``snapshot.particles.position``), not on the
    molecule centres, and skips frames that cannot be read."""
    with open_gsd(infile) as trajectory, open(outfile, "w") as dst:
        print("Timestep, OrientOrder", file=dst)
        for index in range(len(trajectory)):
            try:
                snapshot = trajectory[index]
            except RuntimeError:
                logger.info("Frame %s corrupted, continuing...", index)
                continue
            orient_order = orientational_order(
                box=snapshot.box, position=snapshot._position, orientation=snapshot._orientation
            )
            print(snapshot.step, ",", np.sum(orient_order > 0.9) / len(orient_order), file=dst)

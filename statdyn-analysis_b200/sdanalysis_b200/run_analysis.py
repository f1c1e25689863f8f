"""Per-frame order analysis of a whole trajectory (host loop around ``order.orientational_order``;
counterpart of reference ``run_analysis.py:20-42``)."""

import csv
import logging

import numpy as np

from .gsdio import open_gsd
from .order import orientational_order

logger = logging.getLogger(__name__)

ORDERED_ABOVE = 0.9  # a particle counts as orientationally ordered above this value


def _ordered_fractions(trajectory):
    """``(timestep, fraction of ordered particles)`` of every readable frame.  Like the reference this
    looks at every particle record of a snapshot (``snapshot.particles``), not only at molecule centres."""
    for number in range(len(trajectory)):
        try:
            snapshot = trajectory[number]
        except RuntimeError:
            logger.info("Frame %s corrupted, continuing...", number)
            continue
        values = orientational_order(box=snapshot.box, position=snapshot._position, orientation=snapshot._orientation)
        yield snapshot.step, np.count_nonzero(values > ORDERED_ABOVE) / len(values)


def order(infile, outfile):
    """Write ``Timestep, OrientOrder`` rows -- the fraction of particles whose orientational order
    parameter exceeds 0.9 -- for each frame of the GSD trajectory ``infile``."""
    with open_gsd(infile) as trajectory, open(outfile, "w", newline="") as dst:
        writer = csv.writer(dst, delimiter=" ", lineterminator="\n")
        dst.write("Timestep, OrientOrder\n")
        for step, fraction in _ordered_fractions(trajectory):
            writer.writerow([step, ",", fraction])  # "step , fraction": the reference's print(step, ",", value)

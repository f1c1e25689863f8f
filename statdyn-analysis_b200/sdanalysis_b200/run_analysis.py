"""Per-frame order analysis of a whole trajectory (host loop around ``order.orientational_order``;
counterpart of reference ``run_analysis.py:20-42``)."""

import csv
import logging

import numpy as np

from .gsdio import open_gsd
from .order import orientational_order

logger = logging.getLogger(__name__)

ORDERED_ABOVE = 0.9  # a particle counts as orientationally ordered above this value


def _ordered_fractions(trajectory, frames=None, order_fn=orientational_order):
    """``(frame number, timestep, fraction of ordered particles)`` of every readable frame in ``frames`` (default:
    all).  Like the reference this looks at every particle record of a snapshot (``snapshot.particles``), not only
    at molecule centres."""
    for number in (range(len(trajectory)) if frames is None else frames):
        try:
            snapshot = trajectory[number]
        except RuntimeError:
            logger.info("Frame %s corrupted, continuing...", number)
            continue
        values = order_fn(box=snapshot.box, position=snapshot._position, orientation=snapshot._orientation)
        yield number, snapshot.step, np.count_nonzero(np.asarray(values) > ORDERED_ABOVE) / len(values)


def _write(outfile, rows):
    with open(outfile, "w", newline="") as dst:
        writer = csv.writer(dst, delimiter=" ", lineterminator="\n")
        dst.write("Timestep, OrientOrder\n")
        for step, fraction in rows:
            writer.writerow([step, ",", fraction])  # "step , fraction": the reference's print(step, ",", value)


def order(infile, outfile, *, group=None, order_fn=orientational_order):
    """Write ``Timestep, OrientOrder`` rows -- the fraction of particles whose orientational order
    parameter exceeds 0.9 -- for each frame of the GSD trajectory ``infile``.

    Under ``torch.distributed`` (one process per GPU) the frames are independent, so they are dealt out as
    replicas -- rank r analyses frames r, r + R, r + 2R, ... of the file every rank has mapped, no data-path
    collective ("the order path shards by frame") -- and rank 0 collects the (frame, timestep, fraction)
    triples and writes the file in frame order; the other ranks write nothing."""
    rank, world = 0, 1
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        dist = None
    with open_gsd(infile) as trajectory:
        mine = list(_ordered_fractions(trajectory, range(rank, len(trajectory), world) if world > 1 else None, order_fn))
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0, group=group)
        if rank != 0:
            return
        mine = sorted(row for part in gathered for row in part)
    _write(outfile, [(step, fraction) for _, step, fraction in mine])

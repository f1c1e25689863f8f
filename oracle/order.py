"""CPU restatement of the neighbour / orientational-order part of ``sdanalysis.order``
(reference ``src/sdanalysis/order.py:28-86,174-309``).  TEST INFRASTRUCTURE ONLY.
"""

import numpy as np

from . import rowan_shim as rowan
from .dynamics import create_freud_box
from .freud_shim import NearestNeighbors


def setup_neighbours(box, position, max_radius=3.5, max_neighbours=8, is_2D=True):
    """order.py:174-183."""
    nn = NearestNeighbors(max_radius, max_neighbours, strict_cut=True)
    nn.compute(create_freud_box(box, is_2D), position)
    return nn


def compute_neighbours(box, position, max_radius=3.5, max_neighbours=8):
    """order.py:186-207 -> (N, k) uint32, missing = 2**32 - 1."""
    return setup_neighbours(box, position, max_radius, max_neighbours).getNeighborList()


def _neighbour_relative_angle(neighbourlist, orientation):
    """order.py:28-65: missing neighbours are replaced by the molecule itself (angle 0)."""
    num_mols = len(orientation)
    default_vals = np.arange(num_mols).reshape(-1, 1)
    neighbourlist = np.where(neighbourlist < num_mols, neighbourlist, default_vals)
    return 2 * rowan.intrinsic_distance(orientation[neighbourlist], orientation.reshape(-1, 1, 4))


def _orientational_order(neighbourlist, orientation):
    """order.py:68-86."""
    angles = _neighbour_relative_angle(neighbourlist, orientation)
    return np.mean(np.square(np.cos(angles)), axis=1)


def relative_orientations(box, position, orientation, max_radius=3.5, max_neighbours=8):
    """order.py:210-231."""
    return _neighbour_relative_angle(
        compute_neighbours(box, position, max_radius, max_neighbours), orientation
    )


def orientational_order(box, position, orientation, max_radius=3.5, max_neighbours=8):
    """order.py:234-262."""
    return _orientational_order(
        compute_neighbours(box, position, max_radius, max_neighbours), orientation
    )


def num_neighbours(box, position, max_radius=3.5):
    """order.py:265-281: k = 9, so counts saturate at 9 (trap T8)."""
    return setup_neighbours(box, position, max_radius, 9).nlist.neighbor_counts


def relative_distances(box, position, max_radius=3.5, max_neighbours=8):
    """order.py:284-309: sqrt of r_sq_list with the -1 padding mapped to 0."""
    nn = setup_neighbours(box, position, max_radius, max_neighbours)
    distances = np.empty((len(position), max_neighbours))
    distances[:] = nn.r_sq_list
    distances[distances < 0] = 0
    return np.sqrt(distances)


def compute_voronoi_neighs(box, position, return_query=False, buff=5):
    """order.py:312-315 (``buff=5`` like the reference; a larger buffer gives the exact periodic diagram
    when cells are wider than that, which freud's buffer does not cover)."""
    from .freud_shim import Voronoi

    vor = Voronoi(create_freud_box(box), buff=buff)
    query = vor.computeNeighbors(position)
    return query if return_query else query.nlist.neighbor_counts

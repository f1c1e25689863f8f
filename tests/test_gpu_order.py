"""Parity of the cell-list neighbour kernel and the order parameter with the CPU oracle."""

import numpy as np
import pytest

from oracle import order as oorder

pytestmark = pytest.mark.gpu
MISSING = 2 ** 32 - 1


def _no_ties(rsq):
    """Rows whose listed distances are strictly increasing (tie order is implementation-defined)."""
    valid = rsq >= 0
    d = np.where(valid, rsq, np.inf)
    return np.all((np.diff(d, axis=1) > 0) | ~valid[:, 1:], axis=1)


def _compare(box, pos, quat, r, k):
    from sdanalysis_b200 import order

    got = order.neighbour_analysis(box, pos, quat, r, k, want=("nlist", "rsq", "counts", "order"))
    nn = oorder.setup_neighbours(box, pos, r, k)
    want_list, want_rsq = nn.getNeighborList(), nn.r_sq_list
    rows = _no_ties(want_rsq)
    assert rows.mean() > 0.99
    # bit-exact lists and squared distances (same f32 operations as the oracle's freud restatement)
    assert np.array_equal(got["nlist"][rows], want_list[rows])
    assert np.array_equal(got["rsq"][rows], want_rsq[rows])
    assert np.array_equal(got["counts"], nn.nlist.neighbor_counts)
    want_order = oorder._orientational_order(want_list, quat)
    np.testing.assert_allclose(got["order"][rows], want_order[rows], rtol=1e-5, atol=1e-7)
    return got


def test_liquid_frame(golden_dir):
    g = np.load(golden_dir / "liquid.npz")
    got = _compare(g["box"], g["position"], g["orientation"], 3.5, 8)
    assert (got["nlist"] == MISSING).any()  # the liquid has molecules with fewer than 8 neighbours


def test_liquid_num_neighbours_saturates(golden_dir):
    """SURVEY trap T8: counts saturate at k = 9."""
    from sdanalysis_b200 import order

    g = np.load(golden_dir / "liquid.npz")
    counts = order.num_neighbours(g["box"], g["position"], 3.5)
    assert np.array_equal(counts, oorder.num_neighbours(g["box"], g["position"], 3.5))
    assert counts.max() == 9


@pytest.mark.parametrize("n,r,k", [(4096, 3.5, 8), (5000, 3.5, 9), (3000, 6.0, 16), (700, 2.0, 3)])
def test_synthetic_liquid(n, r, k):
    from sdanalysis_b200.synthetic import SyntheticTrimer

    walk = SyntheticTrimer(n, seed=n)
    walk.advance(400)
    pos, quat = walk.frame()
    _compare(walk.box, pos, quat, r, k)


def test_small_and_tilted_boxes():
    """Boxes with fewer than 3 cells per side and a tilted box (brute-force oracle)."""
    from sdanalysis_b200 import order

    rng = np.random.default_rng(4)
    for box in ([7.5, 12.0, 1, 0, 0, 0], [9.0, 9.0, 1, 0.25, 0, 0], [40.0, 8.0, 1, 0, 0, 0]):
        box = np.array(box, dtype=np.float32)
        n = 300
        pos = np.zeros((n, 3), np.float32)
        pos[:, 0] = rng.uniform(-box[0] / 2, box[0] / 2, n)
        pos[:, 1] = rng.uniform(-box[1] / 2, box[1] / 2, n)
        got = order.neighbour_analysis(box, pos, None, 3.5, 8, want=("nlist", "rsq", "counts"))
        nn = oorder.setup_neighbours(box, pos, 3.5, 8)
        rows = _no_ties(nn.r_sq_list)
        assert np.array_equal(got["nlist"][rows], nn.getNeighborList()[rows]), box
        assert np.array_equal(got["counts"], nn.nlist.neighbor_counts), box


def test_isolated_molecules_have_no_neighbours():
    from sdanalysis_b200 import order

    pos = np.array([[0, 0, 0], [20, 20, 0], [-20, 20, 0], [20, -20, 0]], dtype=np.float32)
    box = [100, 100, 1, 0, 0, 0]
    quat = np.tile(np.array([1, 0, 0, 0], np.float32), (4, 1))
    got = order.neighbour_analysis(box, pos, quat, 3.5, 8, want=("nlist", "rsq", "counts", "order"))
    assert np.all(got["nlist"] == MISSING) and np.all(got["rsq"] == -1) and np.all(got["counts"] == 0)
    assert np.all(got["order"] == 1.0)  # missing neighbours are replaced by the molecule itself


def test_setup_neighbours_query_object(golden_dir):
    """order.setup_neighbours (reference order.py:174-183): list, squared distances and counts of one
    cell-list build agree with the individual functions."""
    from sdanalysis_b200 import order

    g = np.load(golden_dir / "liquid.npz")
    query = order.setup_neighbours(g["box"], g["position"], 3.5, 8)
    assert np.array_equal(query.getNeighborList(), order.compute_neighbours(g["box"], g["position"], 3.5, 8))
    rsq = query.nlist.r_sq_list
    assert np.allclose(np.sqrt(np.where(rsq < 0, 0, rsq)), order.relative_distances(g["box"], g["position"], 3.5, 8))
    assert np.array_equal(np.minimum(query.nlist.neighbor_counts, 8), (query.getNeighborList() != 2 ** 32 - 1).sum(axis=1))


def test_config2_size_the_three_public_functions_share_one_build():
    """BASELINE configs[2] size (32 768 molecules): compute_neighbours + orientational_order + num_neighbours on
    one frame -- read-only numpy arrays as a GSD mapping hands them out, and torch device tensors -- against the
    oracle, which (like the reference, order.py:207,258-261,279-281) builds its neighbour structure three times.
    The k = 8 list must be the first 8 columns of the shared k = 9 search, the count must saturate at 9."""
    import torch

    from sdanalysis_b200 import _lib, order
    from sdanalysis_b200.synthetic import SyntheticTrimer

    n = 32768
    walk = SyntheticTrimer(n, seed=5)
    walk.advance(400)
    pos, quat = walk.frame()
    nn = oorder.setup_neighbours(walk.box, pos, 3.5, 8)
    want_list, want_rsq = nn.getNeighborList(), nn.r_sq_list
    want_counts = oorder.num_neighbours(walk.box, pos, 3.5)
    want_order = oorder._orientational_order(want_list, quat)
    rows = _no_ties(want_rsq)
    assert rows.mean() > 0.99

    def check(nlist, oo, counts):
        assert np.array_equal(nlist[rows], want_list[rows])
        np.testing.assert_allclose(oo[rows], want_order[rows], rtol=1e-5, atol=1e-7)
        assert np.array_equal(counts, want_counts) and counts.max() <= 9

    # read-only numpy frames (what gsdio's mapping hands out): the three calls share one build
    ro_pos = np.frombuffer(pos.tobytes(), dtype=np.float32).reshape(n, 3)
    ro_quat = np.frombuffer(quat.tobytes(), dtype=np.float32).reshape(n, 4)
    launches = _lib.launch_count()
    got = (order.compute_neighbours(walk.box, ro_pos, 3.5, 8), order.orientational_order(walk.box, ro_pos, ro_quat, 3.5, 8),
           order.num_neighbours(walk.box, ro_pos, 3.5))
    assert _lib.launch_count() - launches <= 6
    check(*got)
    # writable numpy arrays are never recognised as "the same frame": three builds, same answers
    launches = _lib.launch_count()
    check(order.compute_neighbours(walk.box, pos, 3.5, 8), order.orientational_order(walk.box, pos, quat, 3.5, 8),
          order.num_neighbours(walk.box, pos, 3.5))
    assert _lib.launch_count() - launches >= 12
    # device tensors in, device tensors out, one build for the three calls
    d_pos, d_quat = torch.from_numpy(pos).cuda(), torch.from_numpy(quat).cuda()
    launches = _lib.launch_count()
    nlist = order.compute_neighbours(walk.box, d_pos, 3.5, 8)
    oo = order.orientational_order(walk.box, d_pos, d_quat, 3.5, 8)
    counts = order.num_neighbours(walk.box, d_pos, 3.5)
    assert _lib.launch_count() - launches <= 6  # bin, scan, scatter, search (+ order from the cached list)
    assert nlist.is_cuda and oo.is_cuda and counts.is_cuda
    check(nlist.cpu().numpy().view(np.uint32), oo.cpu().numpy(), counts.cpu().numpy().view(np.uint32))
    d_pos.add_(0.0)  # an in-place write makes it a different frame: a new build
    launches = _lib.launch_count()
    order.num_neighbours(walk.box, d_pos, 3.5)
    assert _lib.launch_count() - launches >= 4

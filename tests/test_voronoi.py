"""``order.compute_voronoi_neighs`` (reference ``order.py:312-315``, test ``test/order_test.py:48-50``):
the oracle's qhull-based restatement of freud's Voronoi on the reference's crystals, and the CUDA
cell-clipping kernel (``csrc/voronoi.cu``) against it."""

import numpy as np
import pytest

from oracle import order as oorder


@pytest.mark.parametrize("crystal", ["crystal_p2", "crystal_p2gg", "crystal_pg"])
def test_oracle_voronoi_on_crystals(golden_dir, crystal):
    """reference test/order_test.py:48-50: every molecule of the three crystals has 6 Voronoi neighbours."""
    g = np.load(golden_dir / f"{crystal}.npz")
    counts = oorder.compute_voronoi_neighs(g["box"], g["position"])
    assert np.all(counts == 6)


def test_oracle_voronoi_euler_relation(golden_dir):
    """A triangulated torus has exactly 3 N edges: the counts of a generic liquid sum to 6 N."""
    g = np.load(golden_dir / "liquid.npz")
    counts = oorder.compute_voronoi_neighs(g["box"], g["position"])
    assert counts.sum() == 6 * len(counts)
    assert counts.min() >= 3 and counts.max() <= 12


def _ambiguous(query, n, tol=1e-6):
    """Molecules next to a (nearly) degenerate vertex: a ridge shorter than ``tol`` means four almost
    cocircular generators, where the other diagonal is an equally valid neighbour pair."""
    bad = np.zeros(n, dtype=bool)
    short = [pair for pair, length in query.ridge_lengths.items() if length < tol]
    for i, j in short:
        bad[[i, j]] = True
        for a, b in query.pairs:  # the two other generators of the degenerate vertex are neighbours of both
            if a in (i, j):
                bad[b] = True
            if b in (i, j):
                bad[a] = True
    return bad


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["liquid", "crystal_p2", "crystal_p2gg", "crystal_pg"])
def test_cuda_voronoi_counts_match_oracle(golden_dir, name):
    from sdanalysis_b200 import order

    g = np.load(golden_dir / f"{name}.npz")
    query = oorder.compute_voronoi_neighs(g["box"], g["position"], return_query=True)
    want = query.nlist.neighbor_counts
    got = order.compute_voronoi_neighs(g["box"], g["position"])
    assert got.dtype == np.uint32 and got.shape == want.shape
    ok = ~_ambiguous(query, len(want))
    assert ok.mean() > 0.99
    assert np.array_equal(got[ok], want[ok])
    if name != "liquid":
        assert np.all(got == 6)  # reference test/order_test.py:48-50


@pytest.mark.gpu
def test_cuda_voronoi_sparse_and_clustered_points():
    """Voids and clusters: molecules whose 16 nearest neighbours do not close the cell take the second kernel."""
    from sdanalysis_b200 import order

    rng = np.random.default_rng(11)
    n = 3000
    box = np.array([100.0, 80.0, 1.0, 0.0, 0.0, 0.0], dtype=np.float32)
    centres = (rng.random((30, 2)) - 0.5) * box[:2]
    xy = centres[rng.integers(0, 30, n)] + rng.normal(0.0, 1.5, (n, 2))
    xy = (xy + box[:2] / 2) % box[:2] - box[:2] / 2
    pos = np.zeros((n, 3), dtype=np.float32)
    pos[:, :2] = xy
    # cells at the edge of a void are wider than the reference's buffer of 5: compare with the exact
    # periodic diagram (all images)
    query = oorder.compute_voronoi_neighs(box, pos, return_query=True, buff=50)
    want = query.nlist.neighbor_counts
    got = order.compute_voronoi_neighs(box, pos)
    ok = ~_ambiguous(query, n)
    assert ok.mean() > 0.99
    assert int(np.sum(got[ok] != want[ok])) == 0, np.nonzero(got != want)[0][:10]
    assert got.sum() == 6 * n or (~ok).any()


@pytest.mark.gpu
def test_cuda_voronoi_void_strip():
    """A lattice with its last rows missing (32 768 molecules on a 182 x 182 grid): the molecules along the
    void are closed by far-away neighbours in the second kernel; rows of molecules at similar distance must
    not blow up the intermediate cells."""
    from sdanalysis_b200 import order
    from sdanalysis_b200.synthetic import SyntheticTrimer

    walk = SyntheticTrimer(32768, seed=3)
    pos, _ = walk.frame()
    counts = order.compute_voronoi_neighs(walk.box, pos)
    assert counts.sum() == 6 * len(counts)
    query = oorder.compute_voronoi_neighs(walk.box, pos, return_query=True, buff=20)
    ok = ~_ambiguous(query, len(counts))
    assert np.array_equal(counts[ok], query.nlist.neighbor_counts[ok])


@pytest.mark.gpu
def test_cuda_voronoi_synthetic_liquid_euler_relation():
    """Full-size property: 1 M molecules, the counts sum to 6 N (every vertex of a generic diagram has degree 3)."""
    from sdanalysis_b200 import order
    from sdanalysis_b200.synthetic import SyntheticTrimer

    walk = SyntheticTrimer(1_000_000, seed=9)
    pos, _ = walk.frame()
    counts = order.compute_voronoi_neighs(walk.box, pos)
    assert counts.sum() == 6 * len(counts)
    assert counts.min() >= 3 and counts.max() <= 12


@pytest.mark.gpu
def test_ordering_factories(golden_dir, tmp_path):
    """reference order.py:89-171: the three classifier factories on a crystal frame."""
    import joblib
    from sklearn.tree import DecisionTreeClassifier

    from sdanalysis_b200 import order
    from sdanalysis_b200.frame import HoomdFrame  # noqa: F401  (the factories take any object with box/position/orientation)

    g = np.load(golden_dir / "crystal_p2.npz")

    class Snap:
        box, position, orientation = g["box"], g["position"], g["orientation"]

    assert np.all(order.create_neigh_ordering(6)(Snap))
    assert not np.any(order.create_neigh_ordering(5)(Snap))
    assert np.all(order.create_orient_ordering(0.6)(Snap))  # reference test/order_test.py:58-63
    features = order.relative_orientations(Snap.box, Snap.position, Snap.orientation, 3.5, 8)
    labels = (np.arange(len(features)) % 2).astype(int)
    path = tmp_path / "model.pkl"
    joblib.dump(DecisionTreeClassifier(max_depth=3, random_state=0).fit(features, labels), path)
    predicted = order.create_ml_ordering(path)(Snap)
    assert predicted.shape == (len(features),) and set(np.unique(predicted)) <= {0, 1}

"""``calculate_wave_number`` (reference ``dynamics/_util.py:57-86``): the oracle's restatement of
``freud.density.RDF`` against an independent float64 minimum-image histogram, and the CUDA pair
histogram (``csrc/rdf.cu``) against the oracle, bit for bit."""

import numpy as np
import pytest

from oracle import dynamics as odyn
from oracle import freud_shim

F32 = np.float32


def _f64_histogram(box6, pos, rmax, dr, nbins):
    """Independent check: float64, ``d - L rint(d / L)`` minimum image, orthorhombic 2D box."""
    L = np.asarray(box6[:2], dtype=np.float64)
    p = np.asarray(pos[:, :2], dtype=np.float64)
    d = p[None, :, :] - p[:, None, :]
    d -= L * np.rint(d / L)
    r = np.sqrt((d ** 2).sum(-1))
    np.fill_diagonal(r, np.inf)
    r = r[r < rmax]
    return np.bincount(np.floor(r / dr).astype(np.int64), minlength=nbins + 1)[:nbins]


def test_oracle_rdf_matches_float64_histogram(golden_dir):
    g = np.load(golden_dir / "liquid.npz")
    box = odyn.create_freud_box(g["box"], is_2D=True)
    rmax = min(box.Lx / 2.2, box.Ly / 2.2)
    rdf = freud_shim.RDF(rmax=rmax, dr=rmax / 200).compute(box, g["position"])
    assert rdf.nbins in (199, 200)
    want = _f64_histogram(g["box"], g["position"], float(F32(rmax)), float(F32(rmax / 200)), rdf.nbins)
    # f32 rounding moves a pair across a bin edge now and then: compare cumulative counts loosely and totals tightly
    assert abs(int(rdf.counts.sum()) - int(want.sum())) <= 4
    assert np.abs(np.cumsum(rdf.counts.astype(np.int64)) - np.cumsum(want)).max() <= 8
    # a liquid: g(r) -> 1 at large r, excluded core at small r
    assert rdf.RDF[:10].max() == 0.0
    assert abs(float(rdf.RDF[-40:].mean()) - 1.0) < 0.05
    assert np.all(np.diff(rdf.R) > 0) and rdf.R.dtype == np.float32


def test_oracle_wave_number_of_the_liquid_and_crystal(golden_dir):
    g = np.load(golden_dir / "liquid.npz")
    box = odyn.create_freud_box(g["box"], is_2D=True)
    k = odyn.calculate_wave_number(box, g["position"])
    # first peak of S(k) of the Trimer liquid: nearest-neighbour spacing ~2.2 -> k ~ 2 pi / 2.2 ~ 2.9
    # (the reference's CLI default and conftest use 2.90, test/conftest.py:41-48)
    assert 2.5 < k < 3.3
    assert k in np.linspace(0.5, 20, 200)
    relax = odyn.Relaxations(0, g["box"], g["position"], g["orientation"], "trimer", wave_number=None)
    assert relax.wave_number == k and np.isclose(relax.distance, np.pi / (2 * k))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["liquid", "crystal_p2", "crystal_p2gg"])
def test_cuda_pair_histogram_is_bit_exact(golden_dir, name):
    from sdanalysis_b200.dynamics import _util
    from sdanalysis_b200.util import create_freud_box

    g = np.load(golden_dir / f"{name}.npz")
    obox = odyn.create_freud_box(g["box"], is_2D=True)
    rmax = min(obox.Lx / 2.2, obox.Ly / 2.2)
    want = freud_shim.RDF(rmax=rmax, dr=rmax / 200).compute(obox, g["position"])
    got = _util.radial_distribution(create_freud_box(g["box"], is_2D=True), g["position"], rmax=rmax, dr=rmax / 200)
    assert np.array_equal(got.counts, want.counts)
    assert np.array_equal(got.R, want.R) and np.array_equal(got.RDF, want.RDF)
    assert _util.calculate_wave_number(create_freud_box(g["box"], is_2D=True), g["position"]) == odyn.calculate_wave_number(
        obox, g["position"]
    )


@pytest.mark.gpu
def test_cuda_pair_histogram_3d_box_and_tilt():
    from sdanalysis_b200.dynamics import _util
    from sdanalysis_b200.util import Box

    rng = np.random.default_rng(5)
    n = 700
    box6 = np.array([9.0, 11.0, 7.5, 0.2, -0.1, 0.15], dtype=F32)
    pos = ((rng.random((n, 3)) - 0.5) * box6[:3]).astype(F32)
    obox = freud_shim.Box(*box6)
    want = freud_shim.RDF(rmax=3.2, dr=0.05).compute(obox, pos)
    got = _util.radial_distribution(Box(*[float(v) for v in box6]), pos, rmax=3.2, dr=0.05)
    assert np.array_equal(got.counts, want.counts)
    assert np.array_equal(got.RDF, want.RDF)


@pytest.mark.gpu
def test_relaxations_without_wave_number(golden_dir):
    """``Relaxations(wave_number=None)`` estimates it like the reference (``relaxations.py:61-62``)."""
    from sdanalysis_b200 import dynamics
    from sdanalysis_b200.molecules import Trimer

    g = np.load(golden_dir / "liquid.npz")
    relax = dynamics.Relaxations(0, g["box"], g["position"], g["orientation"], molecule=Trimer(), wave_number=None)
    want = odyn.Relaxations(0, g["box"], g["position"], g["orientation"], "trimer", wave_number=None)
    assert relax.wave_number == want.wave_number
    assert relax.distance == want.distance

"""The reference's own hot-path tests, re-hosted against the CUDA drop-in classes.

Same bodies as tests/test_oracle_reference_tests.py (which pins the oracle); here the object under
test is ``sdanalysis_b200``.  Input quaternions are built with the oracle's rowan shim, exactly as
the reference tests build them with rowan.
"""

import numpy as np
import pytest
from hypothesis import assume, example, given, settings
from hypothesis import strategies as st
from hypothesis.extra.numpy import arrays

from oracle import rowan_shim as rowan

pytestmark = pytest.mark.gpu
MAX_BOX = 20.0
FEW = settings(max_examples=20, deadline=None)


@pytest.fixture(scope="module")
def sd():
    import sdanalysis_b200
    from sdanalysis_b200 import dynamics, order, util
    from sdanalysis_b200.molecules import Trimer

    class NS:
        pass

    ns = NS()
    ns.dynamics, ns.order, ns.util, ns.Trimer = dynamics, order, util, Trimer
    return ns


# ---------------------------------------------------------------- test/tracked_motion_test.py
def test_simple_orientation(sd):
    box = sd.util.Box(10, 10, 10)
    motion = sd.dynamics.TrackedMotion(box, np.zeros((1, 3)), rowan.from_euler([0], [0], [0]))
    for i in range(10):
        motion.add(np.zeros((1, 3)), rowan.from_euler([np.pi / 4 * (i + 1)], [0], [0]))
    assert np.isclose(np.linalg.norm(motion.delta_rotation), np.pi / 4 * 10)


@given(st.lists(st.floats(-np.pi / 2, np.pi / 2), min_size=1, max_size=12))
@FEW
def test_orientation(sd, rotations):
    box = sd.util.Box(10, 10, 10)
    motion = sd.dynamics.TrackedMotion(box, np.zeros((1, 3)), rowan.from_euler([0], [0], [0]))
    for orientation in np.cumsum(rotations):
        motion.add(np.zeros((1, 3)), rowan.from_euler([orientation], [0], [0]))
    # reference atol 1e-8 on float64 accumulators; ours are float32 (as for real, f32 trajectories)
    assert np.isclose(motion.delta_rotation[0, 0], np.sum(rotations), atol=4e-6)


def test_simple_translation(sd):
    box = sd.util.Box(10, 10, 10)
    motion = sd.dynamics.TrackedMotion(box, np.zeros((1, 3)), rowan.from_euler([0], [0], [0]))
    for i in range(100):
        motion.add(np.ones((1, 3)) * (i + 1), rowan.from_euler([0], [0], [0]))
    assert np.isclose(np.linalg.norm(motion.delta_translation), np.sqrt(3) * 100)


@given(arrays(np.float64, (10, 3), elements=st.floats(-4, 4)))
@FEW
def test_translation(sd, translations):
    box = sd.util.Box(10, 10, 10)
    motion = sd.dynamics.TrackedMotion(box, np.zeros((1, 3)), rowan.from_euler([0], [0], [0]))
    positions = np.cumsum(translations, axis=0)
    for position in positions:
        motion.add(position.reshape(1, 3), rowan.from_euler([0], [0], [0]))
    assert np.allclose(motion.delta_translation, positions[-1], atol=2e-5)


# ---------------------------------------------------------------- test/dynamics_test.py
@given(arrays(np.float32, (10, 3), elements=st.floats(0, np.float32(MAX_BOX / 2), width=32)))
@FEW
def test_alpha(sd, positions):
    dyn = sd.dynamics.Dynamics(0, [MAX_BOX] * 3, np.zeros_like(positions))
    dyn.add(positions)
    result = dyn.compute_alpha()
    assume(not np.isnan(result))
    assert result >= -1


@given(
    arrays(np.float32, 100, elements=st.floats(0, 10, width=32)),
    arrays(np.float32, 100, elements=st.floats(0, np.float32(2 * np.pi), width=32)),
)
@FEW
def test_overlap(sd, displacement, rotation):
    assert np.isclose(sd.dynamics.dynamics.mobile_overlap(rotation, rotation), 1)
    assert 0.0 <= sd.dynamics.dynamics.mobile_overlap(displacement, rotation) <= 1.0


def test_overlap_known_answer(sd):
    """Distinct keys: the result is fully determined (compare with a host evaluation)."""
    rng = np.random.default_rng(0)
    d = rng.permutation(5000).astype(np.float32) / 7
    r = rng.permutation(5000).astype(np.float32) / 3
    k = 500
    want = len(np.intersect1d(np.argsort(d)[-k:], np.argsort(r)[-k:])) / k
    assert sd.dynamics.dynamics.mobile_overlap(d, r) == want
    with pytest.raises(ZeroDivisionError):
        sd.dynamics.dynamics.mobile_overlap(d[:9], r[:9])


def test_float64_box_and_read_only_arrays(sd):
    rng = np.random.default_rng(0)
    init = rng.random((100, 3)).astype(np.float32)
    final = rng.random((100, 3)).astype(np.float32)
    init.flags.writeable = False
    final.flags.writeable = False
    dyn = sd.dynamics.Dynamics(0, [1, 1, 1], init)
    dyn.add(final)
    assert np.all(dyn.compute_displacement() < 1)


def test_empty_inputs_raise(sd):
    """dynamics.py:90-91, _util.py:127-128: RuntimeError on empty arrays."""
    with pytest.raises(RuntimeError):
        sd.dynamics.Dynamics(0, [1, 1, 1], np.zeros((0, 3), np.float32))
    with pytest.raises(RuntimeError):
        sd.dynamics.TrackedMotion(sd.util.Box(1, 1, 1), np.zeros((3, 3)), np.zeros((0, 4)))
    dyn = sd.dynamics.Dynamics(0, [1, 1, 1], np.zeros((5, 3), np.float32))
    with pytest.raises(ValueError):
        dyn.compute_struct_relax()


def test_MolecularRelaxation(sd):
    n, threshold = 10, 0.4
    tau = sd.dynamics.MolecularRelaxation(n, threshold)
    invalid = np.full(n, tau._max_value, dtype=np.uint32)
    move = lambda dist: np.ones(n) * dist  # noqa: E731
    tau.add(1, move(0))
    assert np.all(tau.get_status() == invalid)
    tau.add(2, move(threshold - 0.1))
    assert np.all(tau.get_status() == invalid)
    tau.add(3, move(threshold + 0.1))
    assert np.all(tau.get_status() == np.full(n, 3))
    tau.add(4, move(threshold - 0.1))
    assert np.all(tau.get_status() == np.full(n, 3))
    tau.add(4, move(threshold + 0.1))
    assert np.all(tau.get_status() == np.full(n, 3))
    with pytest.raises(RuntimeError):
        tau.add(5, np.ones(n + 1))
    tau.add(6, np.full(n, np.nan))  # NaN distances change nothing
    assert np.all(tau.get_status() == np.full(n, 3))


def test_LastMolecularRelaxation(sd):
    n, threshold = 10, 0.4
    tau = sd.dynamics.LastMolecularRelaxation(n, threshold, 1.0)
    invalid = np.full(n, tau._max_value, dtype=np.uint32)
    move = lambda dist: np.ones(n) * dist  # noqa: E731
    script = [
        (2, threshold + 0.1, invalid, 2, 1),
        (3, threshold - 0.1, invalid, 2, 0),
        (4, threshold + 0.1, invalid, 4, 1),
        (5, threshold + 0.2, invalid, 4, 1),
        (6, 1.1, np.full(n, 4), 4, 2),
        (7, threshold - 0.1, np.full(n, 4), 4, 2),
        (8, threshold + 0.1, np.full(n, 4), 4, 2),
    ]
    for timediff, dist, status, raw, state in script:
        tau.add(timediff, move(dist))
        assert np.all(tau.get_status() == status)
        assert np.all(tau._status == np.full(n, raw))
        assert np.all(tau._state == np.full(n, state, dtype=np.uint8))
    # state 0 -> 2 in a single add (relaxations.py:287)
    jump = sd.dynamics.LastMolecularRelaxation(n, threshold, 1.0)
    jump.add(9, move(5.0))
    assert np.all(jump.get_status() == 9) and np.all(jump._state == 2)


@given(arrays(np.float32, 100, elements=st.floats(0, 10, width=32)))
@example(np.full(10, np.nan, dtype=np.float32))
@FEW
def test_structural_relaxation(sd, array):
    assert isinstance(float(sd.dynamics.dynamics.structural_relax(array)), float)


def test_large_rotation(sd):
    dyn = sd.dynamics.Dynamics(0, np.ones(3), np.zeros((10, 3)), sd.util.zero_quaternion(10))
    for i in range(10):
        dyn.add(np.zeros((10, 3)), rowan.from_euler(np.ones(10) * np.pi / 4 * i, np.zeros(10), np.zeros(10)))
    assert np.allclose(np.linalg.norm(dyn.delta_rotation, axis=1), np.pi / 4 * 9)


def test_struct_translation(sd):
    dyn = sd.dynamics.Dynamics(
        0, np.ones(3) * 10, np.array([[0.0, 0.0, 0]]), sd.util.zero_quaternion(1), wave_number=3.14, molecule=sd.Trimer()
    )
    assert np.allclose(dyn.delta_rotation, np.zeros((1, 3)))
    for pos, expected in [([0.1, 0.1, 0], 1), ([0.4, 0.0, 0], 1), ([0.6, 0.0, 0], 0)]:
        dyn.add(np.array([pos]))
        assert dyn.compute_struct_relax() == expected


@given(
    arrays(np.float64, (1, 2), elements=st.floats(-4, 4)),
    arrays(np.float64, (1, 3), elements=st.floats(-np.pi / 3, np.pi / 3)),
)
@example(translation=np.array([[-0.00097108, -0.00097108]]), rotation=np.array([[0.0, 0.0, 0.0]]))
@FEW
def test_struct(sd, translation, rotation):
    translation = np.append(translation, np.zeros((1, 1)), axis=1)
    dyn = sd.dynamics.Dynamics(
        0, np.ones(3) * 10, np.zeros((1, 3)), sd.util.zero_quaternion(1), wave_number=2.9, molecule=sd.Trimer()
    )
    orientation = rowan.from_euler(rotation[:, 0], rotation[:, 1], rotation[:, 2])
    dyn.add(translation, orientation)
    assert np.allclose(dyn.delta_rotation, rotation, atol=2e-7)
    assert np.allclose(dyn.delta_translation, translation, atol=8e-7)


@pytest.mark.parametrize("wave_number", [None, 2.90])
@pytest.mark.parametrize("orientation", [True, False])
def test_compute_all(sd, golden_dir, wave_number, orientation):
    """dynamics_test.py:275-295 on the bundled liquid frame (the trajectory blob is missing)."""
    g = np.load(golden_dir / "liquid.npz")
    dyn = sd.dynamics.Dynamics(int(g["timestep"]), g["box"], g["position"], g["orientation"], wave_number=wave_number)
    moved = g["position"] + np.float32(0.01)
    if orientation:
        result = dyn.compute_all(int(g["timestep"]) + 1, moved, g["orientation"])
    else:
        result = dyn.compute_all(int(g["timestep"]) + 1, moved)
    assert sorted(result.keys()) == sorted(dyn._all_quantities)
    assert np.isnan(result.get("scattering_function"))
    assert result["time"] == 1 and np.isclose(result["msd"], 2e-4, rtol=1e-3)
    assert np.isnan(result["struct"])  # no molecule given
    assert np.isnan(result["com_struct"]) == (wave_number is None)


# ---------------------------------------------------------------- test/test_orientational_ordering.py
def test_orientational_order_parallel_antiparallel_perpendicular(sd):
    nl = np.array([[1], [0]])
    q = np.zeros((2, 4))
    q[:, 0] = 1
    assert np.allclose(sd.order._orientational_order(nl, q), 1)
    q[0, :] = [0, 0, 0, 1]
    assert np.allclose(sd.order._orientational_order(nl, q), 1)
    q[0, :] = rowan.from_euler(np.pi / 2, 0, 0)
    assert np.allclose(sd.order._orientational_order(nl, q), 0, atol=1e-6)


@given(st.floats(allow_nan=False, allow_infinity=False, min_value=-1e3, max_value=1e3))
@FEW
def test_orientational_order_properties(sd, angle):
    assume(np.cos(angle))
    nl = np.array([[1], [0]])
    q = np.array([rowan.from_euler(0, 0, 0), rowan.from_euler(angle, 0, 0)])
    # quaternions are float32 on the device: tolerance of f32 inputs
    assert np.allclose(sd.order._orientational_order(nl, q), np.square(np.cos(angle)), atol=1e-6)


# ---------------------------------------------------------------- test/order_test.py
@pytest.fixture(scope="module", params=["crystal_p2.npz", "crystal_p2gg.npz", "crystal_pg.npz"])
def crystal(request, golden_dir):
    return np.load(golden_dir / request.param)


def test_crystal_known_answers(sd, crystal):
    """order_test.py:31-45,58-63,71-81."""
    box, pos, quat = crystal["box"], crystal["position"], crystal["orientation"]
    n = len(pos)
    neighs = sd.order.compute_neighbours(box, pos, 10, 6)
    assert neighs.shape == (n, 6) and neighs.dtype == np.uint32 and np.all(neighs < n)
    assert np.all(sd.order.num_neighbours(box, pos, 3.5) == 6)
    assert np.all(sd.order.orientational_order(box, pos, quat, 3.5) > 0.60)
    assert np.all(sd.order.create_orient_ordering(0.60)(type("F", (), dict(box=box, position=pos, orientation=quat))))
    assert np.all(np.isfinite(sd.order.relative_distances(box, pos)))
    assert np.all(np.isfinite(sd.order.relative_orientations(box, pos, quat)))


def test_struct_relaxations_and_factory(sd):
    """relaxations.py:297-335: StructRelaxations reports the per-molecule mean over sites;
    create_mol_relaxations validates its arguments and picks the tracker class."""
    from sdanalysis_b200.dynamics import relaxations
    from sdanalysis_b200.molecules import Trimer

    tracker = relaxations.StructRelaxations(4, 0.5, Trimer())
    tracker.add(3, np.array([1.0] * 3 + [0.0] * 9, dtype=np.float32))
    status = tracker.get_status()
    assert status.shape == (4,) and status[0] == 3 and np.all(status[1:] == 2 ** 32 - 1)
    assert isinstance(relaxations.create_mol_relaxations(5, 0.4), relaxations.MolecularRelaxation)
    last = relaxations.create_mol_relaxations(5, 0.4, last_passage=True)
    assert isinstance(last, relaxations.LastMolecularRelaxation)
    with pytest.raises(ValueError):
        relaxations.create_mol_relaxations(5, -0.1)
    with pytest.raises(ValueError):
        relaxations.create_mol_relaxations(5, 0.4, last_passage=True, last_passage_cutoff=-1.0)
    assert relaxations.RelaxationType["rotation"] is relaxations.RelaxationType.rotation

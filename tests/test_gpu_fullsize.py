"""Size-independent properties at BASELINE.json's full molecule count (1 M), where the CPU oracle is
too slow: independent torch evaluations on the device, idempotence, periodicity, and the ordering
of passage times."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 1_000_000
WAVE = 2.9


@pytest.fixture(scope="module")
def run():
    import torch

    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import SyntheticTrimer

    dev = torch.device("cuda", 0)
    walk = SyntheticTrimer(N, seed=9, sigma_r=0.05, sigma_theta=0.12, device=dev)
    eng = DynamicsEngine(N, walk.box, WAVE, mol_sites=Trimer().positions, device=dev)
    rows = []
    frames = []
    for t in range(12):
        pos, quat = walk.frame()
        frames.append((pos.clone(), quat.clone()))
        active = [0, 1] if t < 6 else [0, 1, 2]  # origins 0 and 1 start together, 2 starts at t = 6
        res = eng.step(walk.timestep, pos, quat, active)
        rows.append(dict(zip(res.keys, res.rows())))
        walk.advance(25)
    return eng, rows, frames, walk.box, dev


def test_twin_origins_agree_bit_for_bit(run):
    """Two origins created at the same frame and fed the same frames are identical (state and rows)."""
    eng, rows, *_ = run
    for r in rows:
        for key in r[0]:
            a, b = r[0][key], r[1][key]
            assert a == b or (np.isnan(a) and np.isnan(b)), key
    import torch

    assert torch.equal(eng.origins[0].delta, eng.origins[1].delta)
    assert torch.equal(eng.origins[0].flags, eng.origins[1].flags)
    assert torch.equal(eng.origins[0].status, eng.origins[1].status)


def test_sums_match_independent_torch_reduction(run):
    """msd / mfd / msr / rot1 / com_struct of the last row against torch on the same device state."""
    import torch

    eng, rows, *_ = run
    d = eng.origins[0].delta[:, :N].double()
    d2 = d[0] ** 2 + d[1] ** 2
    th = d[2].abs()
    row = rows[-1][0]
    np.testing.assert_allclose(row["msd"], d2.mean().item(), rtol=1e-6)
    np.testing.assert_allclose(row["mfd"], (d2 ** 2).mean().item(), rtol=1e-6)
    np.testing.assert_allclose(row["mean_displacement"], d2.sqrt().mean().item(), rtol=1e-6)
    np.testing.assert_allclose(row["msr"], (th ** 2).mean().item(), rtol=1e-6)
    np.testing.assert_allclose(row["rot1"], th.cos().mean().item(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(row["rot2"], (2 * th.cos() ** 2 - 1).mean().item(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(row["gamma"] + 1, ((d2 * th ** 2).mean() / (d2.mean() * (th ** 2).mean())).item(), rtol=1e-6)
    dist = np.float32(np.pi / (2 * WAVE))
    f = eng.origins[0].delta[:, :N]
    disp32 = torch.sqrt(f[0] * f[0] + f[1] * f[1])
    assert abs(row["com_struct"] - (disp32 < float(dist)).double().mean().item()) <= 2 / N


def test_overlap_matches_torch_topk(run):
    """Exact top-10 % intersection at 1 M molecules against torch.topk (ties have measure ~0 here)."""
    import torch

    eng, rows, *_ = run
    k = N // 10
    for key in (0, 2):
        f = eng.origins[key].delta[:, :N]
        d2 = f[0] * f[0] + f[1] * f[1]
        th = f[2].abs()
        top_r = torch.zeros(N, dtype=torch.bool, device=f.device)
        top_t = torch.zeros(N, dtype=torch.bool, device=f.device)
        top_r[torch.topk(d2, k).indices] = True
        top_t[torch.topk(th, k).indices] = True
        want = (top_r & top_t).sum().item() / k
        # values equal to the k-th largest may be resolved differently: allow their count
        slack = ((d2 == torch.topk(d2, k).values[-1]).sum().item() + (th == torch.topk(th, k).values[-1]).sum().item()) / k
        assert abs(rows[-1][key]["overlap"] - want) <= slack + 1e-12


def test_struct_matches_independent_formula(run):
    """Reference-mode `struct` (site-major x molecule-major pairing, rotation about x) in torch f64."""
    import torch

    from sdanalysis_b200.molecules import Trimer

    eng, rows, *_ = run
    f = eng.origins[0].delta[:, :N].double()
    sites = torch.tensor(Trimer().positions.astype(np.float32), dtype=torch.float64, device=f.device)
    theta = f[2]
    count = 0
    dist = np.pi / (2 * WAVE)
    idx = torch.arange(N, device=f.device)
    for s in range(3):
        m = (s * N + idx) // 3
        tx, ty = f[0][m], f[1][m]
        vy = sites[s, 1]
        ey = ty - vy * (1 - torch.cos(theta))
        ez = vy * torch.sin(theta)
        count += ((tx ** 2 + ey ** 2 + ez ** 2).sqrt() < dist).sum().item()
    assert abs(rows[-1][0]["struct"] - count / (3 * N)) <= 20 / (3 * N)


def test_zero_step_is_idempotent(run):
    """Recomputing the sums without a new sample changes neither the state nor the result."""
    import torch

    eng, rows, frames, _, _ = run
    before = eng.origins[0].delta.clone()
    res = eng.step(eng.origins[0].timestep + rows[-1][0]["time"], None, None, [0], zero_step=True)
    again = res.rows()[0]
    assert torch.equal(before, eng.origins[0].delta)
    for key in ("msd", "mfd", "alpha", "msr", "rot1", "gamma", "overlap", "struct", "com_struct"):
        assert again[key] == rows[-1][0][key], key


def test_passage_times_are_ordered(run):
    """tau_F (first passage at d) <= tau_D (first passage at 3 d); an irreversible last passage
    implies a first passage no later; rotations: tau_T4 <= tau_T3 <= tau_T2."""
    eng, *_ = run
    s = eng.summary(0)
    unset = 2 ** 32 - 1
    fired_d = s["tau_D"] != unset
    assert np.all(s["tau_F"][fired_d] <= s["tau_D"][fired_d])
    last = s["tau_L"] != unset
    assert np.all(s["tau_F"][last] <= s["tau_L"][last])
    assert np.all(s["tau_T4"] <= s["tau_T3"]) and np.all(s["tau_T3"] <= s["tau_T2"])
    assert fired_d.sum() > 0 and (s["tau_T4"] != unset).sum() > N // 10


def test_periodic_image_leaves_everything_unchanged():
    """A frame shifted by a whole box vector is the same configuration."""
    import torch

    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.synthetic import SyntheticTrimer

    dev = torch.device("cuda", 0)
    n = 200_000
    walk = SyntheticTrimer(n, seed=2, device=dev)
    eng = DynamicsEngine(n, walk.box, WAVE, device=dev, trackers=None)
    pos, quat = walk.frame()
    eng.step(0, pos, quat, [0], want_sums=False)
    shifted = pos.clone()
    shifted[:, 0] += float(walk.box[0])
    shifted[:, 1] -= float(walk.box[1])
    row = eng.step(1, shifted, quat, [0]).rows()[0]
    # wrap maps the image back: the displacement is the f32 rounding of the shift (<= 1 ulp of L)
    assert row["msd"] < (2 * np.spacing(np.float32(walk.box[0]))) ** 2 * 2
    assert row["msr"] == 0.0 and row["rot1"] == 1.0


def test_predicted_overlap_brackets_match_sampled_ones(monkeypatch):
    """The history-predicted brackets of the `overlap` selection (csrc/overlap.cu) give exactly the
    counts of the sampled brackets, engage after two frames of history, never fall back to the exact
    kernel on a smooth trajectory and collect far fewer candidates."""
    import torch

    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.synthetic import SyntheticTrimer

    dev = torch.device("cuda", 0)
    n = 300_000
    walk = SyntheticTrimer(n, seed=12, sigma_r=0.01, sigma_theta=0.02, device=dev)
    monkeypatch.setenv("SDB_OV_PREDICT", "1")
    predicted = DynamicsEngine(n, walk.box, WAVE, device=dev, trackers=None)
    monkeypatch.setenv("SDB_OV_PREDICT", "0")
    sampled = DynamicsEngine(n, walk.box, WAVE, device=dev, trackers=None)
    cands = []
    for t in range(40):
        pos, quat = walk.frame()
        active = [0] if t < 7 else [0, 1]  # a second origin joins with no history
        a = predicted.step(walk.timestep, pos, quat, active).table()
        fails_p, cand_p = predicted.overlap_stats(len(active))
        b = sampled.step(walk.timestep, pos, quat, active).table()
        fails_s, cand_s = sampled.overlap_stats(len(active))
        assert np.array_equal(a["overlap"], b["overlap"]), t
        assert fails_p == 0 and fails_s == 0, t
        cands.append((int(cand_p[0]), int(cand_s[0])))
        walk.advance(1)
    assert cands[1][0] == cands[1][1] > 0  # no history yet: both engines sample
    assert all(p < s / 3 for p, s in cands[25:])  # old origin: predicted brackets are much tighter


def test_four_million_molecules_config5_shape():
    """BASELINE.json configs[4] molecule count (4 M, sigma_r = 0.03, LastMolecularRelaxation among the
    six default trackers): sums against torch on the device state, passage-time ordering, and the
    multi-tile / multi-segment indexing at the largest N of the configurations."""
    import torch

    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import SyntheticTrimer

    dev = torch.device("cuda", 0)
    n = 4_000_000
    walk = SyntheticTrimer(n, seed=5, sigma_r=0.03, sigma_theta=0.06, device=dev)
    eng = DynamicsEngine(n, walk.box, WAVE, mol_sites=Trimer().positions, device=dev)
    table = None
    for t in range(10):
        pos, quat = walk.frame()
        active = [0] if t < 4 else [0, 1, 2]
        table = eng.step(walk.timestep, pos, quat, active).table()
        walk.advance(40)
    d = eng.origins[0].delta[:, :n].double()
    d2 = d[0] ** 2 + d[1] ** 2
    th = d[2].abs()
    np.testing.assert_allclose(table["msd"][0], d2.mean().item(), rtol=1e-6)
    np.testing.assert_allclose(table["msr"][0], (th ** 2).mean().item(), rtol=1e-6)
    np.testing.assert_allclose(table["rot1"][0], th.cos().mean().item(), rtol=1e-5, atol=1e-7)
    dist = np.float32(np.pi / (2 * WAVE))
    disp32 = torch.sqrt((eng.origins[0].delta[0, :n] ** 2 + eng.origins[0].delta[1, :n] ** 2))
    assert abs(table["com_struct"][0] - (disp32 < float(dist)).double().mean().item()) <= 2 / n
    k = n // 10
    top_r = torch.topk(d2.float(), k).indices
    top_t = torch.topk(th.float(), k).indices
    both = np.intersect1d(top_r.cpu().numpy(), top_t.cpu().numpy()).size
    assert abs(table["overlap"][0] * k - both) <= 2  # ties at the cut are broken by index
    s = eng.summary(0)
    unset = 2 ** 32 - 1
    fired = s["tau_D"] != unset
    assert fired.sum() > 0 and np.all(s["tau_F"][fired] <= s["tau_D"][fired])
    last = s["tau_L"] != unset
    assert last.sum() > 0 and np.all(s["tau_F"][last] <= s["tau_L"][last])
    assert table["time"].tolist() == [360, 200, 200]


def test_per_frame_increments_pass_is_bit_identical(monkeypatch):
    """sdb_dyn_increments (phase 1 once per frame and previous sample, used when several small segments
    share it) leaves every accumulator, flag and sum bit-identical to the in-kernel phase 1; also with
    a partial last tile and origins created along the way."""
    import torch

    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import SyntheticTrimer

    dev = torch.device("cuda", 0)
    n = 70_001  # partial last tile, n not a multiple of 4
    walk = SyntheticTrimer(n, seed=8, sigma_r=0.05, sigma_theta=0.1, device=dev)
    monkeypatch.setenv("SDB_INCREMENTS", "1")
    monkeypatch.setenv("SDB_INCREMENTS_MAX_PER_SEGMENT", "64")
    with_pass = DynamicsEngine(n, walk.box, WAVE, mol_sites=Trimer().positions, device=dev)
    monkeypatch.setenv("SDB_INCREMENTS", "0")
    without = DynamicsEngine(n, walk.box, WAVE, mol_sites=Trimer().positions, device=dev)
    used = 0
    for t in range(8):
        pos, quat = walk.frame()
        active = list(range(20 if t < 4 else 24))  # 3 segments of 8; four more origins join at t = 4
        a = with_pass.step(walk.timestep, pos, quat, active)
        b = without.step(walk.timestep, pos, quat, active)
        used += with_pass._inc_ws is not None
        assert np.array_equal(a.sums(), b.sums()), t
        walk.advance(3)
    assert used > 0  # the pass really ran
    for key in (0, 7, 19, 23):
        assert torch.equal(with_pass.origins[key].delta, without.origins[key].delta)
        assert torch.equal(with_pass.origins[key].flags, without.origins[key].flags)
        assert torch.equal(with_pass.origins[key].status[:, :n], without.origins[key].status[:, :n])

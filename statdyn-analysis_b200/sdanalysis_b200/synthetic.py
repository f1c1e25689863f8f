"""Synthetic 2D Trimer-liquid trajectories for tests and benchmarks.

Number density 0.2, square box, jittered square lattice at t = 0, then an unbiased random walk:
per timestep Gaussian increments ``sigma_r`` per Cartesian component and ``sigma_theta`` in the
orientation; positions are wrapped into [-L/2, L/2) and stored as float32, orientations as
quaternions (cos(t/2), 0, 0, sin(t/2)) float32, z = 0 -- the layout of ``frame.py:162-184``.

``SyntheticTrimer`` keeps its walk state in float64 on the host (numpy) or on a CUDA device
(torch); the latter is used by ``bench.py`` to create frames directly in HBM.
"""

import numpy as np

DENSITY = 0.2


def box_for(num_mols, density=DENSITY):
    side = np.float32(np.sqrt(num_mols / density))
    return np.array([side, side, 1, 0, 0, 0], dtype=np.float32)


class SyntheticTrimer:
    def __init__(self, num_mols, seed=1, sigma_r=0.01, sigma_theta=0.02, device=None, density=DENSITY):
        self.n = int(num_mols)
        self.sigma_r, self.sigma_theta = float(sigma_r), float(sigma_theta)
        self.box = box_for(self.n, density)
        self.side = float(self.box[0])
        self.timestep = 0
        self.device = device
        cols = int(np.ceil(np.sqrt(self.n)))
        spacing = self.side / cols
        if device is None:
            self.rng = np.random.default_rng(seed)
            idx = np.arange(self.n)
            grid = np.stack([idx % cols, idx // cols], axis=1).astype(np.float64)
            self.xy = (grid + 0.5) * spacing - self.side / 2 + self.rng.uniform(-0.3, 0.3, (self.n, 2))
            self.theta = self.rng.uniform(-np.pi, np.pi, self.n)
        else:
            import torch

            self.torch = torch
            self.gen = torch.Generator(device=device)
            self.gen.manual_seed(int(seed))
            idx = torch.arange(self.n, device=device)
            grid = torch.stack([idx % cols, idx // cols], dim=1).to(torch.float64)
            jitter = (torch.rand((self.n, 2), generator=self.gen, device=device, dtype=torch.float64) - 0.5) * 0.6
            self.xy = (grid + 0.5) * spacing - self.side / 2 + jitter
            self.theta = (torch.rand(self.n, generator=self.gen, device=device, dtype=torch.float64) - 0.5) * (2 * np.pi)

    def advance(self, steps):
        """Move the walk forward by ``steps`` timesteps (one Gaussian draw of the summed increment)."""
        if steps <= 0:
            return
        scale = np.sqrt(steps)
        if self.device is None:
            self.xy += self.rng.normal(0.0, self.sigma_r * scale, (self.n, 2))
            self.theta += self.rng.normal(0.0, self.sigma_theta * scale, self.n)
        else:
            torch = self.torch
            self.xy += torch.randn((self.n, 2), generator=self.gen, device=self.device, dtype=torch.float64) * (
                self.sigma_r * scale
            )
            self.theta += torch.randn(self.n, generator=self.gen, device=self.device, dtype=torch.float64) * (
                self.sigma_theta * scale
            )
        self.timestep += int(steps)

    def frame(self):
        """(position (N,3) f32, orientation (N,4) f32) of the current state."""
        L = self.side
        if self.device is None:
            wrapped = self.xy - L * np.floor(self.xy / L + 0.5)
            pos = np.zeros((self.n, 3), dtype=np.float32)
            pos[:, :2] = wrapped
            pos[:, :2][pos[:, :2] >= np.float32(L / 2)] -= np.float32(L)
            quat = np.zeros((self.n, 4), dtype=np.float32)
            quat[:, 0] = np.cos(self.theta / 2)
            quat[:, 3] = np.sin(self.theta / 2)
            return pos, quat
        torch = self.torch
        wrapped = self.xy - L * torch.floor(self.xy / L + 0.5)
        pos = torch.zeros((self.n, 3), dtype=torch.float32, device=self.device)
        pos[:, :2] = wrapped.to(torch.float32)
        quat = torch.zeros((self.n, 4), dtype=torch.float32, device=self.device)
        quat[:, 0] = torch.cos(self.theta / 2).to(torch.float32)
        quat[:, 3] = torch.sin(self.theta / 2).to(torch.float32)
        return pos, quat


def linear_schedule(num_frames, keyframe_interval, keyframe_max):
    """(timestep, indexes) of the reference's linear mode (``read/_gsd.py:99-115``): frames at
    timesteps 0..num_frames-1, a new origin whenever ``timestep % keyframe_interval == 0`` while
    ``len(index_list) <= keyframe_max``."""
    indexes = []
    for t in range(num_frames):
        if t % keyframe_interval == 0 and len(indexes) <= keyframe_max:
            indexes.append(len(indexes))
        yield t, list(indexes)


def exponential_schedule(total_steps, num_linear, gen_steps, max_gen):
    """(timestep, indexes) of the reference's exponential mode (``read/_gsd.py:135-189``)."""
    from .StepSize import GenerateStepSeries

    series = GenerateStepSeries(total_steps, num_linear=num_linear, gen_steps=gen_steps, max_gen=max_gen)
    for step in series:
        yield int(step), list(series.get_index())


def synthetic_frames(schedule, num_mols, seed=1, sigma_r=0.01, sigma_theta=0.02, device=None):
    """Yield ``(indexes, ArrayFrame)`` for a schedule of ``(timestep, indexes)`` pairs."""
    from .frame import ArrayFrame

    walk = SyntheticTrimer(num_mols, seed, sigma_r, sigma_theta, device)
    for timestep, indexes in schedule:
        walk.advance(timestep - walk.timestep)
        pos, quat = walk.frame()
        yield indexes, ArrayFrame(timestep, walk.box, pos, quat)

"""world_size-2 gloo tests of the origin-sharded multi-GPU plumbing (no GPU needed).

The compute step is replaced by a stub that derives its "sums" from the broadcast frame, so the
tests check exactly what the distributed layer owns: every rank sees the reader rank's frame, each
origin is advanced by its owner rank only, and rank 0 re-assembles the rows in the reference's order.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdanalysis_b200.distributed import ShardedDynamics, frame_floats, owner_of, partition

N_MOLS = 1003  # not a multiple of 4: exercises the 16-byte alignment padding of the quaternions
NSUM = 16


def test_partition_is_a_round_robin_cover():
    keys = list(range(23))
    for world in (1, 2, 3, 8):
        parts = [partition(keys, r, world) for r in range(world)]
        assert sorted(k for p in parts for k in p) == keys
        assert all(owner_of(k, world) == r for r, p in enumerate(parts) for k in p)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert frame_floats(5) == 16 + 20 and frame_floats(8) == 24 + 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def step_fn(timestep, pos, quat, keys):
            # fake sums: column 0 = key, 1 = timestep, 2 = checksum of the frame seen by this rank
            checksum = float(pos.double().sum() + 3 * quat.double().sum())
            sums = torch.zeros((len(keys), NSUM), dtype=torch.float64)
            for i, k in enumerate(keys):
                sums[i, 0], sums[i, 1], sums[i, 2], sums[i, 3] = k, timestep, checksum, rank
            calls.append((timestep, list(keys)))
            return list(keys), [0 for _ in keys], sums

        # ranks may own ceil or floor(K / R) origins and pass max_local accordingly: they agree on the largest
        # (buffers shared between ranks are sized with it)
        sharded = ShardedDynamics(N_MOLS, step_fn, max_local=8 - rank, device="cpu")
        assert sharded.max_local == 8
        rng = np.random.default_rng(5)
        results = []
        handle = None
        for t in range(6):
            pos = rng.normal(size=(N_MOLS, 3)).astype(np.float32)
            quat = rng.normal(size=(N_MOLS, 4)).astype(np.float32)
            active = list(range(min(t + 1, 5)))  # origins appear one per frame, like the linear schedule
            handle = sharded.post_frame(pos, quat) if rank == 0 else sharded.post_frame()
            collect = sharded.step(100 + t, active, handle)
            keys, timediffs, sums = collect()
            if rank == 0:
                want = float(pos.astype(np.float64).sum() + 3 * quat.astype(np.float64).sum())
                results.append((keys, timediffs, sums.tolist(), want, active))
        # the same frames again with batched gathering: rows stay local until flush()
        batched = ShardedDynamics(N_MOLS, step_fn, max_local=8, device="cpu", batch=4)
        rng = np.random.default_rng(5)
        flushed = []
        for t in range(6):
            pos = rng.normal(size=(N_MOLS, 3)).astype(np.float32)
            quat = rng.normal(size=(N_MOLS, 4)).astype(np.float32)
            active = list(range(min(t + 1, 5)))
            handle = batched.post_frame(pos, quat) if rank == 0 else batched.post_frame()
            got = batched.step(100 + t, active, handle, gather="batch")
            if got:  # the batch filled up: step() flushed it
                flushed.extend(got)
        flushed.extend(batched.flush())
        out.put((rank, calls, results, [(k, td, s.tolist()) for k, td, s in flushed]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_step_broadcasts_frames_and_gathers_rows():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, calls, results, flushed = out.get(timeout=100)
        got[rank] = (calls, results, flushed)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # each rank only ever advances the origins it owns
    for rank, (calls, _, _) in got.items():
        assert all(owner_of(k, world) == rank for _, keys in calls for k in keys)
    # rank 0 sees every origin of every frame, in the order of `active`, with the right payload
    for t, (keys, timediffs, sums, want, active) in enumerate(got[0][1]):
        assert keys == active
        assert timediffs == [t - k for k in active]  # origin k is created at frame k
        sums = np.array(sums)
        assert np.array_equal(sums[:, 0], active) and np.all(sums[:, 1] == 100 + t)
        # every rank computed from the frame that rank 0 supplied
        assert np.allclose(sums[:, 2], want, rtol=1e-12)
        assert np.array_equal(sums[:, 3], [owner_of(k, world) for k in active])
    # batched gathering returns the same rows, frame by frame, on rank 0 only
    assert got[1][2] == []
    assert len(got[0][2]) == len(got[0][1])
    for (keys, timediffs, sums, _, _), (bkeys, btd, bsums) in zip(got[0][1], got[0][2]):
        assert keys == bkeys and timediffs == btd
        assert np.array_equal(np.array(sums), np.array(bsums))


def _order_worker(rank, world, port, infile, outfile, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sdanalysis_b200 import run_analysis

        seen = []

        def order_fn(box, position, orientation):  # stub of order.orientational_order: a value per particle
            seen.append(float(position[0, 0]))
            return np.where(np.arange(len(position)) < position[0, 0], 1.0, 0.0)

        run_analysis.order(infile, outfile, order_fn=order_fn)
        out.put((rank, seen))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_order_path_shards_frames_as_replicas(tmp_path):
    """run_analysis.order under torch.distributed: every rank analyses its own frames (no collective on the data
    path), rank 0 writes all rows in frame order."""
    from sdanalysis_b200.gsdio import GSDWriter

    infile, outfile = tmp_path / "t.gsd", tmp_path / "order.csv"
    with GSDWriter(infile) as w:
        for frame in range(7):
            pos = np.zeros((10, 3), dtype=np.float32)
            pos[0, 0] = frame  # the stub orders `frame` of the 10 particles
            w.append(100 * frame, [5, 5, 1, 0, 0, 0], pos, np.tile(np.array([1, 0, 0, 0], np.float32), (10, 1)))
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_order_worker, args=(r, world, port, str(infile), str(outfile), out)) for r in range(world)]
    for p in procs:
        p.start()
    seen = dict(out.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert seen[0] == [0.0, 2.0, 4.0, 6.0] and seen[1] == [1.0, 3.0, 5.0]  # frames r, r + R, ...
    lines = outfile.read_text().strip().splitlines()
    assert lines[0] == "Timestep, OrientOrder"
    assert [line.split() for line in lines[1:]] == [[str(100 * f), ",", str(f / 10)] for f in range(7)]

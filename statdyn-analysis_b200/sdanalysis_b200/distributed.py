"""Multi-GPU driver: time-origins sharded across ranks, frames broadcast from the reader rank.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).  The reference is a
single process, so this layer is new: per-origin state is only ever touched by its own
origin, hence

* origin ``k`` lives on rank ``k % world`` (round-robin keeps the progressively activated origins of
  the linear schedule balanced) and never moves;
* the only shared input is the current frame: the reader rank (0) uploads it and broadcasts the
  28 B/molecule over NVLink; the broadcast of frame t+1 is issued before the kernels of frame t are
  waited for, so it overlaps with compute;
* the per-(frame, origin) rows (16 doubles) are gathered to rank 0, which assembles the ``dynamics``
  table in the reference's row order; passage-time tables stay sharded.

The compute step is injected (``step_fn``) so the plumbing can be exercised on CPU tensors.
"""

import numpy as np

from . import _lib

SDB_NSUM = _lib.SDB_NSUM


def owner_of(key, world):
    return int(key) % world


def partition(keys, rank, world):
    """The subsequence of ``keys`` owned by ``rank`` (order preserved)."""
    return [k for k in keys if owner_of(k, world) == rank]


def frame_floats(n):
    """Length of the broadcast buffer: positions (3n, padded to 16 bytes) + quaternions (4n)."""
    return (3 * n + 3) // 4 * 4 + 4 * n


class ShardedDynamics:
    """Origin-sharded wrapper around one compute step per rank.

    ``step_fn(timestep, pos, quat, local_keys) -> (keys, timediffs, sums)`` advances the local
    origins and returns a ``(len(keys), SDB_NSUM)`` float64 tensor on ``device`` (the CUDA engine or,
    in tests, a stub).  ``max_local`` bounds the number of local origins active in one frame.
    """

    def __init__(self, num_mols, step_fn, max_local, *, device, group=None, depth=3):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n = int(num_mols)
        self.device = torch.device(device)
        self.step_fn = step_fn
        self.max_local = int(max_local)
        self.qoff = (3 * self.n + 3) // 4 * 4
        self.frames = [torch.zeros(frame_floats(self.n), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self._slot = 0
        # gather buffers: one block of max_local rows per rank; rank 0 keeps `depth` receive sets so a
        # result can be read while later frames are in flight
        self._send = [torch.zeros((self.max_local, SDB_NSUM), dtype=torch.float64, device=self.device) for _ in range(depth)]
        self._recv = (
            [[torch.zeros_like(self._send[0]) for _ in range(self.world)] for _ in range(depth)]
            if self.rank == 0
            else None
        )
        self._start = {}  # origin key -> timestep of its first frame (known to every rank)
        # frame uploads and broadcasts run on a side stream so that they overlap the kernels of the
        # previous frame; `_consumed[slot]` marks the end of the kernels that read a frame buffer
        import os

        self._cuda = self.device.type == "cuda"
        # Measured on 2 x B200 (round 1): issuing the NCCL broadcast from a side stream so that it
        # overlaps the kernels of the previous frame slows those kernels down 1.5-8x (independent of
        # NCCL_MAX_NCHANNELS), so the exchange is issued in-stream by default.
        self._use_side = self._cuda and os.environ.get("SDB_SHARD_SIDE", "0") == "1"
        self._async_gather = os.environ.get("SDB_SHARD_ASYNC_GATHER", "1") == "1"
        self._side = torch.cuda.Stream(device=self.device) if self._use_side else None
        self._consumed = [None] * depth
        self._gathers = [None] * depth

    # -- frame exchange ---------------------------------------------------------------------------
    def post_frame(self, position=None, orientation=None):
        """Start the broadcast of the next frame (rank 0 supplies it; others pass nothing).

        ``position`` (n,3) / ``orientation`` (n,4) may be host arrays, pinned CPU tensors or tensors
        on ``device``.  Returns a handle for :meth:`step`."""
        torch = self.torch
        self._slot = (self._slot + 1) % len(self.frames)
        buf = self.frames[self._slot]

        def exchange():
            if self.rank == 0:
                if position is None:
                    raise ValueError("rank 0 is the reader rank and must supply the frame")
                pos = torch.as_tensor(position)
                quat = torch.as_tensor(orientation)
                buf[: 3 * self.n].view(self.n, 3).copy_(pos, non_blocking=True)
                buf[self.qoff : self.qoff + 4 * self.n].view(self.n, 4).copy_(quat, non_blocking=True)
            if self.world > 1:
                return self.dist.broadcast(buf, src=0, group=self.group, async_op=True)
            return None

        if not self._use_side:
            return (self._slot, exchange(), None)
        with torch.cuda.stream(self._side):
            consumed = self._consumed[self._slot]
            if consumed is not None:
                self._side.wait_event(consumed)  # the kernels that read this buffer have finished
            work = exchange()
            if work is not None:
                work.wait()  # orders the side stream after the broadcast
            ready = torch.cuda.Event()
            ready.record(self._side)
        return (self._slot, None, ready)

    def _frame_views(self, slot):
        buf = self.frames[slot]
        pos = buf[: 3 * self.n].view(self.n, 3)
        quat = buf[self.qoff : self.qoff + 4 * self.n].view(self.n, 4)
        return pos, quat

    # -- one frame ------------------------------------------------------------------------------------
    def step(self, timestep, active, handle, gather=True):
        """Advance the local share of ``active`` with the frame behind ``handle``.

        Returns, on rank 0 with ``gather``, a callable producing ``(keys, timediffs, sums)`` for *all*
        origins in the order of ``active``; on the other ranks (or without ``gather``) the local
        triple."""
        torch, dist = self.torch, self.dist
        slot, work, ready = handle
        if work is not None:
            work.wait()
        if ready is not None:
            torch.cuda.current_stream(self.device).wait_event(ready)  # frame uploaded and broadcast
        pos, quat = self._frame_views(slot)
        local = partition(active, self.rank, self.world)
        if len(local) > self.max_local:
            raise ValueError(f"{len(local)} local origins exceed max_local={self.max_local}")
        for k in active:
            self._start.setdefault(k, int(timestep))
        keys, timediffs, sums = self.step_fn(timestep, pos, quat, local) if local else ([], [], None)
        if keys != local:
            # the engine may reorder origins (grouping by previous sample): restore `local` order
            order = [keys.index(k) for k in local]
            sums = sums[torch.tensor(order, device=sums.device)] if sums is not None else None
            keys, timediffs = [keys[i] for i in order], [timediffs[i] for i in order]
        if self._use_side:
            done = torch.cuda.Event()
            done.record()
            self._consumed[slot] = done
        if not gather or self.world == 1:
            return lambda: (keys, timediffs, None if sums is None else sums.cpu().numpy())
        send = self._send[slot % len(self._send)]
        previous = self._gathers[slot % len(self._send)]
        if previous is not None:
            previous.wait()  # the gather that last used these buffers (long finished) is ordered first
        if keys:
            send[: len(keys)].copy_(sums, non_blocking=True)
        recv = self._recv[slot % len(self._send)] if self.rank == 0 else None
        gathered = dist.gather(send, recv, dst=0, group=self.group, async_op=True)
        if not self._async_gather:
            gathered.wait()
        self._gathers[slot % len(self._send)] = gathered
        if self.rank != 0:
            return lambda: (gathered.wait(), (keys, timediffs, None))[1]
        world, start = self.world, self._start

        def collect():
            gathered.wait()
            rows = {}
            for r, block in enumerate(recv):
                mine = partition(active, r, world)
                if mine:
                    host = block[: len(mine)].cpu().numpy()
                    for k, row in zip(mine, host):
                        rows[k] = row
            out = np.stack([rows[k] for k in active]) if active else np.zeros((0, SDB_NSUM))
            return list(active), [int(timestep) - start[k] for k in active], out

        return collect

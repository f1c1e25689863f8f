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
ROW = SDB_NSUM + 2  # gathered rows carry (key, timediff) behind the sums


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

    def __init__(self, num_mols, step_fn, max_local, *, device, group=None, depth=2):
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
        # gather buffers: one row block per rank, padded to max_local (+1 header row: count)
        self._send = torch.zeros((self.max_local + 1, ROW), dtype=torch.float64, device=self.device)
        self._recv = (
            [torch.zeros_like(self._send) for _ in range(self.world)] if self.rank == 0 else None
        )
        self._pending = None

    # -- frame exchange ---------------------------------------------------------------------------
    def post_frame(self, position=None, orientation=None):
        """Start the broadcast of the next frame (rank 0 supplies it; others pass nothing).

        ``position`` (n,3) / ``orientation`` (n,4) may be host arrays, pinned CPU tensors or tensors
        on ``device``.  Returns a handle for :meth:`step`."""
        torch = self.torch
        self._slot = (self._slot + 1) % len(self.frames)
        buf = self.frames[self._slot]
        if self.rank == 0:
            if position is None:
                raise ValueError("rank 0 is the reader rank and must supply the frame")
            pos = torch.as_tensor(position)
            quat = torch.as_tensor(orientation)
            buf[: 3 * self.n].view(self.n, 3).copy_(pos, non_blocking=True)
            buf[self.qoff : self.qoff + 4 * self.n].view(self.n, 4).copy_(quat, non_blocking=True)
        work = None
        if self.world > 1:
            work = self.dist.broadcast(buf, src=0, group=self.group, async_op=True)
        return (self._slot, work)

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
        slot, work = handle
        if work is not None:
            work.wait()  # orders the current stream after the broadcast
        pos, quat = self._frame_views(slot)
        local = partition(active, self.rank, self.world)
        if len(local) > self.max_local:
            raise ValueError(f"{len(local)} local origins exceed max_local={self.max_local}")
        keys, timediffs, sums = self.step_fn(timestep, pos, quat, local) if local else ([], [], None)
        if not gather or self.world == 1:
            return lambda: (keys, timediffs, None if sums is None else sums.cpu().numpy())
        send = self._send
        send.zero_()
        send[0, 0] = len(keys)
        if keys:
            send[1 : 1 + len(keys), :SDB_NSUM] = sums
            meta = torch.tensor([[k, t] for k, t in zip(keys, timediffs)], dtype=torch.float64, device=self.device)
            send[1 : 1 + len(keys), SDB_NSUM:] = meta
        dist.gather(send, self._recv if self.rank == 0 else None, dst=0, group=self.group)
        if self.rank != 0:
            return lambda: (keys, timediffs, None)
        recv = [r.clone() for r in self._recv]

        def collect():
            rows = {}
            for block in recv:
                host = block.cpu().numpy()
                count = int(host[0, 0])
                for row in host[1 : 1 + count]:
                    rows[int(row[SDB_NSUM])] = (int(row[SDB_NSUM + 1]), row[:SDB_NSUM].copy())
            out_keys = [k for k in active if k in rows]
            out_td = [rows[k][0] for k in out_keys]
            out = np.stack([rows[k][1] for k in out_keys]) if out_keys else np.zeros((0, SDB_NSUM))
            return out_keys, out_td, out

        return collect

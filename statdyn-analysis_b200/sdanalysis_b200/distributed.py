"""Multi-GPU driver: time-origins sharded across ranks, frames broadcast from the reader rank.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).  The reference is a
single process, so this layer is new: per-origin state is only ever touched by its own
origin, hence

* origin ``k`` lives on rank ``k % world`` (round-robin keeps the progressively activated origins of
  the linear schedule balanced) and never moves;
* the only shared input is the current frame: the reader rank (0) uploads it and pushes the
  28 B/molecule into every rank's frame ring over NVLink with multicast stores (``csrc/exchange.cu``)
  on a side stream, so the exchange of frame t+1 overlaps the kernels of frame t;
* the per-(frame, origin) rows (16 doubles) are gathered to rank 0 -- per frame or once per batch of
  frames -- which assembles the ``dynamics`` table in the reference's row order; passage-time tables
  stay sharded.

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

    Frame exchange (``exchange``):

    ``"multicast"``  (CUDA, default) every rank holds a ring of frame slots in torch symmetric memory.
        The reader rank pushes frame t into slot t % depth of *every* rank with one pass of
        ``multimem.st`` stores through the NVSwitch multicast address (``csrc/exchange.cu``), on a side
        stream, followed by a device-side barrier; compute waits on an event.  The reader's NVLink
        egress is 28 B per molecule whatever the number of ranks, and no collective kernel competes
        with the compute kernels.  Without multicast support the peers pull the slot from the reader
        with copy-engine peer copies instead.
    ``"collective"`` ``torch.distributed.broadcast`` (NCCL on GPUs, gloo in the CPU tests).

    Rows: ``step(..., gather=True)`` gathers every frame's rows to rank 0 with a collective (CPU tests,
    small runs); ``gather="batch"`` keeps them in a device buffer that :meth:`flush` gathers once per
    ``batch`` frames -- nothing but the frame exchange happens per frame.
    """

    def __init__(self, num_mols, step_fn, max_local, *, device, group=None, depth=3, exchange=None, batch=64, direct_rows=False):
        import os

        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n = int(num_mols)
        self.device = torch.device(device)
        self.step_fn = step_fn
        # direct_rows: step_fn(timestep, pos, quat, keys, rows_out=ptr) writes the rows of the frame in `keys`
        # order straight to the device address `ptr` -- with the native engine that is one peer copy into
        # rank 0's row buffer, enqueued by the same native call as the kernels
        self.direct_rows = bool(direct_rows)
        self.max_local = int(max_local)
        if self.world > 1:
            # symmetric buffers and the batched gather are sized with max_local: every rank must use the same value
            # (ranks own ceil or floor(K / R) origins): agree on the largest one
            t = torch.tensor([self.max_local], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            self.max_local = int(t.item())
        self.depth = int(depth)
        self.qoff = (3 * self.n + 3) // 4 * 4
        self.slot_floats = (frame_floats(self.n) + 3) // 4 * 4  # 16-byte multiple
        self._cuda = self.device.type == "cuda"
        self._slot = 0
        self._start = {}  # origin key -> timestep of its first frame (known to every rank)
        self._consumed = [None] * self.depth
        self._view_cache = {}
        self._slice_stage = None
        self.h2d_bytes = 0
        self._last_active, self._last_local = None, None
        self.stalls = [] if os.environ.get("SDB_SHARD_PROFILE") == "1" else None
        self.exchange_times = []
        self.exchange_note = ""

        if exchange is None:
            exchange = os.environ.get("SDB_SHARD_EXCHANGE", "multicast" if self._cuda and self.world > 1 else "collective")
        self.exchange = exchange
        self._hdl = None
        if exchange == "multicast":
            try:
                self._setup_symmetric()
            except Exception as exc:  # no symmetric memory on this system: collectives still work
                # LOUD: the data plane changes (NCCL broadcast instead of multicast stores); callers that
                # report the exchange (bench.py's config.parallelism) read self.exchange / exchange_note
                import sys

                self.exchange_note = f"multicast unavailable: {type(exc).__name__}: {exc}"
                print(
                    f"[sdanalysis_b200 rank {self.rank}] WARNING: symmetric-memory multicast frame exchange unavailable "
                    f"({type(exc).__name__}: {exc}); falling back to torch.distributed.broadcast",
                    file=sys.stderr, flush=True,
                )
                if os.environ.get("SDB_SHARD_STRICT", "0") == "1":
                    raise
                self.exchange = "collective"
        if self.exchange == "collective":
            self.frames = [
                torch.zeros(self.slot_floats, dtype=torch.float32, device=self.device) for _ in range(self.depth)
            ]
        # Measured on 2 x B200 (round 1): issuing the NCCL broadcast from a side stream so that it
        # overlaps the kernels of the previous frame slows those kernels down 1.5-8x (independent of
        # NCCL_MAX_NCHANNELS), so the collective exchange is issued in-stream by default.
        self._use_side = self._cuda and (self.exchange != "collective" or os.environ.get("SDB_SHARD_SIDE", "0") == "1")
        self._side = torch.cuda.Stream(device=self.device, priority=-1) if self._use_side else None

        # per-frame gather (gather=True): one block of max_local rows per rank; rank 0 keeps `depth`
        # receive sets so a result can be read while later frames are in flight
        self._async_gather = os.environ.get("SDB_SHARD_ASYNC_GATHER", "1") == "1"
        self._send = self._recv = None
        self._gathers = [None] * self.depth
        # batched gather (gather="batch")
        self.batch = int(batch)
        self._batch_buf = None
        self._batch_recv = self._batch_host = None
        self._rows_r0 = None
        self._batch_meta = []

    # -- symmetric-memory ring ------------------------------------------------------------------------
    def _setup_symmetric(self):
        torch, dist = self.torch, self.dist
        import torch.distributed._symmetric_memory as symm_mem

        group = self.group if self.group is not None else dist.group.WORLD
        self._ring = symm_mem.empty(self.depth * self.slot_floats, dtype=torch.float32, device=self.device)
        self._hdl = symm_mem.rendezvous(self._ring, group.group_name)
        self._ring.zero_()
        self.frames = [self._ring[i * self.slot_floats : (i + 1) * self.slot_floats] for i in range(self.depth)]
        self._mc_ptr = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
        import os

        self._mc_blocks = int(os.environ.get("SDB_MC_BLOCKS", "32"))
        if self.rank == 0:
            # the reader stages a frame here (H2D or D2D) before pushing it to every rank
            self._stage = [torch.zeros(self.slot_floats, dtype=torch.float32, device=self.device) for _ in range(2)]
            self._stage_i = 0
            self._stage_free = [None, None]
        else:
            self._reader_ring = self._hdl.get_buffer(0, (self.depth * self.slot_floats,), torch.float32)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def _ensure_gather_buffers(self):
        if self._send is None:
            torch = self.torch
            self._send = [
                torch.zeros((self.max_local, SDB_NSUM), dtype=torch.float64, device=self.device) for _ in range(self.depth)
            ]
            if self.rank == 0:
                self._recv = [[torch.zeros_like(self._send[0]) for _ in range(self.world)] for _ in range(self.depth)]

    # -- frame exchange ---------------------------------------------------------------------------
    def _views(self, buf):
        """(pos (n,3), quat (n,4)) views of a frame buffer, cached (slicing tensors is slow)."""
        key = buf.data_ptr()
        views = self._view_cache.get(key)
        if views is None:
            views = (buf[: 3 * self.n].view(self.n, 3), buf[self.qoff : self.qoff + 4 * self.n].view(self.n, 4))
            self._view_cache[key] = views
        return views

    def _fill(self, buf, position, orientation):
        torch = self.torch
        if position is None:
            raise ValueError("rank 0 is the reader rank and must supply the frame")
        pos_view, quat_view = self._views(buf)
        pos_view.copy_(torch.as_tensor(position), non_blocking=True)
        quat_view.copy_(torch.as_tensor(orientation), non_blocking=True)

    def slice_bounds(self, rank=None):
        """Rows ``[lo, hi)`` of a frame that ``rank`` uploads in the sliced exchange (multiples of 4)."""
        r = self.rank if rank is None else rank
        per = (self.n // self.world) // 4 * 4
        lo = r * per
        hi = self.n if r == self.world - 1 else lo + per
        return lo, hi

    def post_frame_sliced(self, position, orientation):
        """Exchange of a frame that EVERY rank can read in host memory (e.g. the same memory-mapped GSD
        file, or a shared pinned buffer): rank r uploads only rows ``slice_bounds(r)`` over its own PCIe
        link and multicasts that slice into the frame slot of every rank, so no single host link carries
        the whole frame.  A rank may instead pass just its rows as device tensors (a trajectory resident
        in HBM, sharded by rows): they are multicast in place.  Called by all ranks with the same frame;
        returns a handle for :meth:`step`."""
        torch = self.torch
        if self.exchange != "multicast" or not self._mc_ptr:
            return self.post_frame(position if self.rank == 0 else None, orientation if self.rank == 0 else None)
        self._slot = (self._slot + 1) % self.depth
        slot = self._slot
        lo, hi = self.slice_bounds()
        pos_t, quat_t = torch.as_tensor(position), torch.as_tensor(orientation)
        # a rank that already holds ITS rows of the frame on the device (shape (hi - lo, ...)) multicasts
        # them straight from there
        direct = (
            pos_t.is_cuda and quat_t.is_cuda and pos_t.shape[0] == hi - lo and quat_t.shape[0] == hi - lo
            and pos_t.dtype == torch.float32 and quat_t.dtype == torch.float32
            and pos_t.is_contiguous() and quat_t.is_contiguous()
            and (hi - lo) % 4 == 0 and pos_t.data_ptr() % 16 == 0 and quat_t.data_ptr() % 16 == 0
        )
        own_rows = pos_t.shape[0] == hi - lo and quat_t.shape[0] == hi - lo and self.world > 1
        if not direct and not own_rows and pos_t.shape[0] != self.n:
            # every rank reaches the barrier below or none does: a rank that cannot take part says so
            # BEFORE anybody waits (all ranks evaluate the same shape test on their own argument)
            raise ValueError("post_frame_sliced needs the whole frame in host memory or this rank's rows on the device")
        # the multicast copy moves whole 16-byte words: the last rank's rows (n need not be a multiple of
        # 4) are staged in a buffer padded to 4 rows; the pad lands in the slot's own padding before qoff
        rows_pad = (hi - lo + 3) // 4 * 4
        if not direct and self._slice_stage is None:
            self._slice_stage = [
                (torch.zeros((rows_pad, 3), dtype=torch.float32, device=self.device),
                 torch.zeros((rows_pad, 4), dtype=torch.float32, device=self.device))
                for _ in range(2)
            ]
            self._slice_free = [None, None]
            self._slice_i = 0
        main = _lib.current_stream(self.device) if direct else None
        with torch.cuda.stream(self._side):
            nxt = self._consumed[(slot + 1) % self.depth]
            if nxt is not None:
                self._side.wait_event(nxt)
            if self.stalls is not None:
                x0 = torch.cuda.Event(enable_timing=True)
                x0.record(self._side)
            lib = _lib.load()
            base = self._mc_ptr + 4 * slot * self.slot_floats
            if direct:
                self._side.wait_event(main.record_event())  # the caller's tensors are complete
                spos, squat = pos_t, quat_t
            else:
                self._slice_i ^= 1
                spos, squat = self._slice_stage[self._slice_i]
                free = self._slice_free[self._slice_i]
                if free is not None:
                    self._side.wait_event(free)
                if own_rows:  # this rank's rows, but not 16-byte tileable (n % 4 != 0 on the last rank): staged
                    if pos_t.is_cuda:
                        self._side.wait_event(_lib.current_stream(self.device).record_event())
                    spos[: hi - lo].copy_(pos_t, non_blocking=True)
                    squat[: hi - lo].copy_(quat_t, non_blocking=True)
                else:
                    self._upload_rows(lib, spos, squat, pos_t, quat_t, lo, hi)
            pos_bytes = (12 * (hi - lo) + 15) // 16 * 16  # <= 4 * qoff - 12 * lo: stays inside the position part
            _lib.check(lib.sdb_multicast_copy(spos.data_ptr(), base + 12 * lo, pos_bytes, self._mc_blocks, self._side.cuda_stream))
            _lib.check(lib.sdb_multicast_copy(squat.data_ptr(), base + 4 * self.qoff + 16 * lo, 16 * (hi - lo), self._mc_blocks, self._side.cuda_stream))
            if not direct:
                self._slice_free[self._slice_i] = self._side.record_event()
            self._hdl.barrier(channel=0)
            ready = torch.cuda.Event(enable_timing=self.stalls is not None)
            ready.record(self._side)
            if self.stalls is not None:
                self.exchange_times.append((x0, ready))
        self.h2d_bytes = 0 if direct else 28 * (hi - lo)
        return (slot, None, ready)

    def _upload_rows(self, lib, spos, squat, pos_t, quat_t, lo, hi):
        """Rows ``[lo, hi)`` of a host frame into this rank's device staging buffers, on the side stream, without
        ever synchronising it: page-locked sources (pinned tensors, a registered trajectory mapping) are
        uploaded in place with one asynchronous copy each; pageable ones go through a page-locked staging pair
        first (a host copy by the library's copy threads).  torch's own ``copy_`` is not used here: for a source
        it does not recognise as pinned it synchronises the stream, which stops the host from running ahead of
        the device (measured: +0.2 ms per step at 2 GPUs)."""
        torch = self.torch
        stream = self._side.cuda_stream
        count = hi - lo
        if pos_t.dtype != torch.float32 or quat_t.dtype != torch.float32 or not pos_t.is_contiguous() or not quat_t.is_contiguous():
            pos_t, quat_t = pos_t.to(torch.float32).contiguous(), quat_t.to(torch.float32).contiguous()
        src_pos = pos_t.data_ptr() + 12 * lo
        src_quat = quat_t.data_ptr() + 16 * lo
        pinned = (pos_t.is_pinned() and quat_t.is_pinned()) or (
            _lib.is_pinned(src_pos, 12 * count) and _lib.is_pinned(src_quat, 16 * count)
        )
        if not pinned:
            if getattr(self, "_slice_pinned", None) is None:
                self._slice_pinned = [
                    (torch.empty((spos.shape[0], 3), dtype=torch.float32).pin_memory(),
                     torch.empty((squat.shape[0], 4), dtype=torch.float32).pin_memory())
                    for _ in range(2)
                ]
                self._slice_pinned_done = [None, None]
            hpos, hquat = self._slice_pinned[self._slice_i]
            done = self._slice_pinned_done[self._slice_i]
            if done is not None:
                done.synchronize()  # the upload that last read this staging pair (two frames ago)
            lib.sdb_test_host_copy(hpos.data_ptr(), src_pos, 12 * count)
            lib.sdb_test_host_copy(hquat.data_ptr(), src_quat, 16 * count)
            src_pos, src_quat = hpos.data_ptr(), hquat.data_ptr()
        _lib.check(lib.sdb_upload_async(spos.data_ptr(), src_pos, 12 * count, stream))
        _lib.check(lib.sdb_upload_async(squat.data_ptr(), src_quat, 16 * count, stream))
        if not pinned:
            self._slice_pinned_done[self._slice_i] = self._side.record_event()
        self._upload_keep = (pos_t, quat_t)  # the source stays referenced until the next upload replaces it

    def post_frame(self, position=None, orientation=None):
        """Start the exchange of the next frame (rank 0 supplies it; others pass nothing).

        ``position`` (n,3) / ``orientation`` (n,4) may be host arrays, pinned CPU tensors or tensors
        on ``device``.  Returns a handle for :meth:`step`."""
        torch = self.torch
        self._slot = (self._slot + 1) % self.depth
        slot = self._slot
        buf = self.frames[slot]

        if self.exchange == "collective":
            def exchange():
                if self.rank == 0:
                    self._fill(buf, position, orientation)
                if self.world > 1:
                    return self.dist.broadcast(buf, src=0, group=self.group, async_op=True)
                return None

            if not self._use_side:
                return (slot, exchange(), None)
            with torch.cuda.stream(self._side):
                consumed = self._consumed[slot]
                if consumed is not None:
                    self._side.wait_event(consumed)  # the kernels that read this buffer have finished
                work = exchange()
                if work is not None:
                    work.wait()  # orders the side stream after the broadcast
                ready = torch.cuda.Event()
                ready.record(self._side)
            return (slot, None, ready)

        # symmetric-memory ring: see the class docstring.  Slot reuse: before arriving at the barrier of
        # frame u every rank waits for its own kernels on the slot of frame u + 1; the reader writes that
        # slot only after the barrier, so nobody is still reading it.
        main = _lib.current_stream(self.device) if self._cuda else None
        with torch.cuda.stream(self._side):
            nxt = self._consumed[(slot + 1) % self.depth]
            if nxt is not None:
                self._side.wait_event(nxt)
            if self.stalls is not None:
                x0 = torch.cuda.Event(enable_timing=True)
                x0.record(self._side)
            if self.rank == 0:
                self._side.wait_event(main.record_event())  # the caller's frame tensors are complete
                pos_t, quat_t = torch.as_tensor(position), torch.as_tensor(orientation)
                direct = (
                    self._mc_ptr
                    and pos_t.is_cuda and quat_t.is_cuda
                    and pos_t.dtype == torch.float32 and quat_t.dtype == torch.float32
                    and pos_t.is_contiguous() and quat_t.is_contiguous()
                    and (12 * self.n) % 16 == 0 and pos_t.data_ptr() % 16 == 0 and quat_t.data_ptr() % 16 == 0
                )
                lib = _lib.load()
                base = self._mc_ptr + 4 * slot * self.slot_floats
                if direct:
                    # device-resident frame: multicast straight from the caller's tensors (the caller keeps
                    # them alive and unchanged until the step that consumes this frame has been enqueued)
                    _lib.check(lib.sdb_multicast_copy(pos_t.data_ptr(), base, 12 * self.n, self._mc_blocks, self._side.cuda_stream))
                    _lib.check(lib.sdb_multicast_copy(quat_t.data_ptr(), base + 4 * self.qoff, 16 * self.n, self._mc_blocks, self._side.cuda_stream))
                else:
                    self._stage_i ^= 1
                    stage = self._stage[self._stage_i]
                    free = self._stage_free[self._stage_i]
                    if free is not None:
                        self._side.wait_event(free)
                    self._fill(stage, pos_t, quat_t)
                    if self._mc_ptr:
                        _lib.check(lib.sdb_multicast_copy(stage.data_ptr(), base, 4 * self.slot_floats, self._mc_blocks, self._side.cuda_stream))
                    else:
                        buf.copy_(stage, non_blocking=True)
                    self._stage_free[self._stage_i] = self._side.record_event()
            self._hdl.barrier(channel=0)
            if self.rank != 0 and not self._mc_ptr:
                lo = slot * self.slot_floats
                buf.copy_(self._reader_ring[lo : lo + self.slot_floats], non_blocking=True)
            ready = torch.cuda.Event(enable_timing=self.stalls is not None)
            ready.record(self._side)
            if self.stalls is not None:
                self.exchange_times.append((x0, ready))
        return (slot, None, ready)

    def _frame_views(self, slot):
        return self._views(self.frames[slot])

    # -- one frame ------------------------------------------------------------------------------------
    def step(self, timestep, active, handle, gather=True):
        """Advance the local share of ``active`` with the frame behind ``handle``.

        ``gather=True``: returns, on rank 0, a callable producing ``(keys, timediffs, sums)`` for *all*
        origins in the order of ``active`` (on the other ranks the local triple).  ``gather="batch"``:
        the rows stay on the device until :meth:`flush`; returns None.  ``gather=False``: a callable
        producing the local triple."""
        torch, dist = self.torch, self.dist
        slot, work, ready = handle
        if work is not None:
            work.wait()
        main = None
        if ready is not None:
            main = _lib.current_stream(self.device)
            if self.stalls is not None:  # diagnostics: time the compute stream spends waiting for the frame
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(main)
                main.wait_event(ready)
                e1.record(main)
                self.stalls.append((e0, e1))
            else:
                main.wait_event(ready)  # frame uploaded and delivered
        pos, quat = self._frame_views(slot)
        if active == self._last_active:  # steady state: the same origins as in the previous frame
            local = self._last_local
        else:
            local = partition(active, self.rank, self.world)
            if len(local) > self.max_local:
                raise ValueError(f"{len(local)} local origins exceed max_local={self.max_local}")
            for k in active:
                self._start.setdefault(k, int(timestep))
            self._last_active, self._last_local = list(active), local
        rows_ptr = 0
        if self.direct_rows and gather == "batch" and self.world > 1:
            if self._batch_buf is None:
                self._setup_batch_rows()
            i = len(self._batch_meta)
            row_bytes = self.max_local * SDB_NSUM * 8
            if self._rows_r0 is not None:
                rows_ptr = self._rows_r0_ptr[self._rows_parity] + (self.rank * self.batch + i) * row_bytes
            else:
                rows_ptr = self._batch_buf.data_ptr() + i * row_bytes
        if not local:
            keys, timediffs, sums = [], [], None
        elif rows_ptr:
            keys, timediffs, sums = self.step_fn(timestep, pos, quat, local, rows_out=rows_ptr)
        else:
            keys, timediffs, sums = self.step_fn(timestep, pos, quat, local)
        if keys != local:
            # the engine may reorder origins (grouping by previous sample): restore `local` order
            order = [keys.index(k) for k in local]
            sums = sums[torch.tensor(order, device=sums.device)] if sums is not None else None
            keys, timediffs = [keys[i] for i in order], [timediffs[i] for i in order]
        if self._use_side:
            done = torch.cuda.Event()
            done.record(main if main is not None else _lib.current_stream(self.device))
            self._consumed[slot] = done
        if gather == "batch" and self.world > 1:
            if self._batch_buf is None:
                self._setup_batch_rows()
            i = len(self._batch_meta)
            if rows_ptr:
                pass  # already written by step_fn
            elif keys and sums is not None:
                if self._rows_r0 is not None:
                    # straight into rank 0's row buffer over NVLink (symmetric memory): flush() then needs no
                    # collective, only a barrier and one device-to-host copy on rank 0
                    self._rows_r0[self._rows_parity][self.rank, i, : len(keys)].copy_(sums, non_blocking=True)
                else:
                    self._batch_buf[i, : len(keys)].copy_(sums, non_blocking=True)
            self._batch_meta.append((int(timestep), list(active)))
            if len(self._batch_meta) == self.batch:
                self._flushed = self.flush()
                return self._flushed
            return None
        if not gather or self.world == 1:
            return lambda: (keys, timediffs, None if sums is None else sums.cpu().numpy())
        self._ensure_gather_buffers()
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

    def _setup_batch_rows(self):
        """Buffers of the batched row gather.  With symmetric memory (CUDA, multicast exchange): two
        alternating row buffers ``[world][batch][max_local][SDB_NSUM]`` on rank 0 that every rank writes its
        rows into directly; otherwise a local buffer that :meth:`flush` gathers with a collective."""
        torch = self.torch
        shape = (self.batch, self.max_local, SDB_NSUM)
        self._batch_buf = torch.zeros(shape, dtype=torch.float64, device=self.device)
        self._rows_r0 = None
        import os

        if self.exchange == "multicast" and self._hdl is not None and os.environ.get("SDB_SHARD_ROWS", "symmetric") == "symmetric":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                group = self.group if self.group is not None else self.dist.group.WORLD
                per = self.world * self.batch * self.max_local * SDB_NSUM
                self._rows_sym = symm_mem.empty(2 * per, dtype=torch.float64, device=self.device)
                self._rows_hdl = symm_mem.rendezvous(self._rows_sym, group.group_name)
                self._rows_sym.zero_()
                full = (2, self.world) + shape
                on_r0 = self._rows_sym.view(full) if self.rank == 0 else self._rows_hdl.get_buffer(0, full, torch.float64)
                self._rows_r0 = [on_r0[0], on_r0[1]]
                self._rows_r0_ptr = [on_r0[0].data_ptr(), on_r0[1].data_ptr()]
                self._rows_parity = 0
                torch.cuda.synchronize(self.device)
                self.dist.barrier(group=self.group)
            except Exception as exc:
                import warnings

                warnings.warn(f"symmetric-memory row buffer unavailable ({exc}); gathering rows with a collective")
                self._rows_r0 = None

    def flush(self):
        """Gather the rows of the frames stepped with ``gather="batch"`` since the last flush.

        Returns, on rank 0, a list of ``(keys, timediffs, sums)`` per frame (all origins, in the order
        of ``active``); an empty list on the other ranks.  One collective for up to ``batch`` frames."""
        torch, dist = self.torch, self.dist
        meta, self._batch_meta = self._batch_meta, []
        if not meta or self._batch_buf is None:
            return []
        count = len(meta)
        if self._rows_r0 is not None:
            # every rank has written its rows into rank 0's buffer: a device-side barrier orders those writes
            # before rank 0's copy to the host.  The buffers alternate, and a rank reaches the next flush's
            # barrier only after rank 0 has finished reading this one, so nothing is overwritten early.
            parity = self._rows_parity
            self._rows_parity ^= 1
            self._rows_hdl.barrier(channel=0)
            if self.rank != 0:
                return []
            if self._batch_host is None:
                self._batch_host = torch.empty((self.world,) + tuple(self._batch_buf.shape), dtype=torch.float64, pin_memory=True)
            # the whole (small, contiguous) buffer: a strided slice would go through a staging copy
            self._batch_host.copy_(self._rows_r0[parity])
            return self._assemble(meta, self._batch_host.numpy())
        send = self._batch_buf[:count]  # leading rows of a contiguous buffer: contiguous
        if self.rank == 0:
            if self._batch_recv is None:
                self._batch_recv = torch.empty((self.world,) + tuple(self._batch_buf.shape), dtype=torch.float64, device=self.device)
                self._batch_host = torch.empty(tuple(self._batch_recv.shape), dtype=torch.float64, pin_memory=self._cuda)
            recv = [self._batch_recv[r, :count] for r in range(self.world)]
        else:
            recv = None
        dist.gather(send, recv, dst=0, group=self.group)
        if self.rank != 0:
            return []
        # one device-to-host copy for all ranks and frames
        self._batch_host[:, :count].copy_(self._batch_recv[:, :count])
        return self._assemble(meta, self._batch_host.numpy())

    def _assemble(self, meta, host):
        """Rows of every frame in the order of its ``active`` list from ``host[rank][frame][local row]``."""
        count = len(meta)
        out = []
        i = 0
        while i < count:
            # a run of frames with the same active list shares its scatter pattern: one assignment per rank
            active = meta[i][1]
            j = i + 1
            while j < count and meta[j][1] == active:
                j += 1
            index = {k: c for c, k in enumerate(active)}
            t0 = np.array([self._start[k] for k in active], dtype=np.int64)
            sums = np.zeros((j - i, len(active), SDB_NSUM))
            for r in range(self.world):
                place = np.array([index[k] for k in partition(active, r, self.world)], dtype=np.int64)
                if len(place):
                    sums[:, place] = host[r, i:j, : len(place)]
            for f in range(i, j):
                out.append((list(active), (meta[f][0] - t0).tolist(), sums[f - i]))
            i = j
        return out

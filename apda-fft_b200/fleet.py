"""Fleet sweep plumbing: windows are independent, so a batch is sharded contiguously over one process per GPU and
the only exchange is a gather of the fixed-size peak records to rank 0 (torch.distributed: NCCL on GPUs, gloo in the
CPU tests).  No collective touches the data path."""
from __future__ import annotations

import os


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Rank r owns windows [r*ceil(total/world), min(total, (r+1)*ceil(total/world)))."""
    per = -(-total // world) if world > 0 else total
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def shard_capacity(total: int, world: int) -> int:
    return -(-total // world)


def gather_records(local_recs, total: int, dst: int = 0, group=None):
    """local_recs: uint8 tensor [shard_capacity(total, world), rec_bytes] (rows past the shard's own windows are
    padding).  Returns the [total, rec_bytes] table on rank ``dst`` (None elsewhere); row order == window order."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_recs[:total]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = shard_capacity(total, world)
    if rank == dst:
        # gather straight into one table: rank r's rows land at [r*per, (r+1)*per), which IS window order
        table = torch.empty((world * per,) + tuple(local_recs.shape[1:]), dtype=local_recs.dtype,
                            device=local_recs.device)
        parts = list(table.view(world, per, *local_recs.shape[1:]).unbind(0))
        dist.gather(local_recs, gather_list=parts, dst=dst, group=group)
        return table[:total]
    dist.gather(local_recs, gather_list=None, dst=dst, group=group)
    return None


class RecordGatherer:
    """Gather of the records in slices, overlapped with the compute of the following slices.

    A shard is analysed in a few sub-batches; as soon as the records of one sub-batch are complete its rows are sent
    to rank ``dst`` asynchronously (the collective runs on the backend's own stream / thread) while the next sub-batch
    is being analysed.  Only the last slice's transfer is exposed.  Row order of the table == window order, exactly as
    ``gather_records`` (rank r's rows at [r*per, (r+1)*per)).

        g = RecordGatherer(per, rec_bytes, device)         # once
        for lo, hi in g.slices(parts):                     # every step
            ... analyse windows [lo, hi) of the local shard into local_recs[lo:hi] ...
            g.start(local_recs, lo, hi)
        table = g.finish()                                 # rank dst: [world*per, rec_bytes]; others: None
    """

    def __init__(self, per: int, rec_bytes: int, device, dst: int = 0, group=None):
        import torch
        import torch.distributed as dist
        self.per, self.dst, self.group = per, dst, group
        self.active = dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.table = None
        if self.active and self.rank == dst:
            self.table = torch.empty((self.world * per, rec_bytes), dtype=torch.uint8, device=device)
        self.pending = []
        self.local = None

    def slices(self, parts: int):
        """Contiguous [lo, hi) row ranges of the local shard (one slice when there is nothing to overlap)."""
        parts = max(1, min(parts if self.active else 1, self.per))
        step = -(-self.per // parts)
        return [(lo, min(self.per, lo + step)) for lo in range(0, self.per, step)]

    def start(self, local_recs, lo: int, hi: int):
        import torch.distributed as dist
        self.local = local_recs
        if not self.active:
            return
        rows = local_recs[lo:hi]
        if self.rank == self.dst:
            view = self.table.view(self.world, self.per, -1)
            parts = [view[r, lo:hi] for r in range(self.world)]
            self.pending.append(dist.gather(rows, gather_list=parts, dst=self.dst, group=self.group, async_op=True))
        else:
            self.pending.append(dist.gather(rows, gather_list=None, dst=self.dst, group=self.group, async_op=True))

    def finish(self):
        """Wait for the outstanding slices (the current stream waits; the host does not block on NCCL)."""
        for work in self.pending:
            work.wait()
        self.pending = []
        if not self.active:
            return self.local
        return self.table


class PeerRecordTable:
    """The fleet's record table in the destination rank's HBM, mapped into every rank (CUDA IPC over NVLink).

    Rank ``dst`` allocates ``buffers`` tables of ``world * per`` records; the others open the allocation and hand
    ``table(step) + rank * per * rec_bytes`` to the pickers as their record pointer, so the K3 kernels' own 128-byte
    epilogue stores land in the owner's memory: the gather of SURVEY 8(e) fused into the producing kernel, with no
    collective, no staging copy and no SM time.  ``torch.distributed`` only carries the 64-byte handle and barriers.

    Steps are numbered 1, 2, ...; step s uses table ``s % buffers``.  Flow control is on the device timelines in both
    directions: a producer publishes ``s`` after its pickers (``signal``), the owner's stream holds until every rank
    published ``s`` (``wait``), and once the owner has consumed the table it acknowledges ``s`` (``release``); a
    producer about to write step ``s`` first holds its stream until step ``s - buffers`` is acknowledged (``begin``),
    so it can never overwrite rows the owner is still reading.

        t = PeerRecordTable(analyzer.ctx, per, 128, device)       # once (collective: every rank calls it)
        for s in 1, 2, ...:
            ptr = t.begin(s)                                      # every rank: back-pressure, then this step's rows
            ... an.peaks_device(..., d_rec=ptr, ...) ...          # any number of launches
            t.signal(s)
            if t.owner:
                table = t.wait(s)                                 # uint8 [world*per, rec_bytes], valid until release(s)
                ... consume on the same stream ...
                t.release(s)
        t.close()
    """

    def __init__(self, ctx, per: int, rec_bytes: int, device, dst: int = 0, group=None, buffers: int = 2):
        import ctypes

        import torch
        import torch.distributed as dist
        self.ctx, self.per, self.rec_bytes, self.dst, self.group = ctx, per, rec_bytes, dst, group
        self.device = device
        self.buffers = max(1, int(buffers))
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.owner = self.rank == dst
        self.nbytes = self.world * per * rec_bytes                      # one table
        self.stride = (self.nbytes + 255) & ~255
        # control words behind the tables: [0, 4*world) step counters, +128 owner's time-out word, +132 acknowledged step,
        # +256 + 4*rank the producers' own time-out words
        self.flags_off = self.stride * self.buffers
        self.ctl_bytes = 512
        base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        err = None
        if self.owner:
            try:
                ctx.call("apda_peer_table_create", ctypes.c_int64(self.flags_off + self.ctl_bytes), ctypes.byref(base), handle)
            except Exception as exc:  # noqa: BLE001 - reported after the collective steps below, never before
                err = exc
        if self.world > 1:
            h = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            dist.broadcast(h, src=dst, group=group)
            if not self.owner:
                try:
                    if os.environ.get("APDA_TEST_PEER_OPEN_FAILS"):  # exercises the collective fail-over (tests only)
                        raise RuntimeError("peer table open disabled by APDA_TEST_PEER_OPEN_FAILS")
                    raw = (ctypes.c_ubyte * 64)(*h.cpu().tolist())
                    ctx.call("apda_peer_table_open", raw, ctypes.byref(base))
                except Exception as exc:  # noqa: BLE001
                    err = exc
            # every rank learns whether all mappings exist, so that a failure raises everywhere instead of hanging
            good = torch.tensor([0.0 if err else 1.0], device=device)
            dist.all_reduce(good, op=dist.ReduceOp.MIN, group=group)
            if float(good[0]) < 1.0:
                if base.value:
                    ctx.call("apda_peer_table_destroy" if self.owner else "apda_peer_table_close", ctypes.c_void_p(base.value))
                raise err or RuntimeError("peer table: another rank could not map the allocation")
        elif err:
            raise err
        self.base = int(base.value)
        self.local_ptr = self.base + self.rank * per * rec_bytes        # rows of this rank in table 0
        self._closed = False
        if self.owner:  # counters start at 0
            view = self._raw(self.flags_off, self.ctl_bytes)
            view.zero_()
            torch.cuda.synchronize(device)
        if self.world > 1:
            dist.barrier(group=group)

    def table_ptr(self, step: int) -> int:
        return self.base + (int(step) % self.buffers) * self.stride

    def row_ptr(self, row: int, step: int = 0) -> int:
        """Device pointer of local row ``row`` in the table of ``step`` (for sub-batch launches)."""
        return self.table_ptr(step) + (self.rank * self.per + row) * self.rec_bytes

    def _raw(self, offset: int, nbytes: int):
        import torch

        class _Mem:  # the allocation belongs to libapda_b200, torch only views it
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.base + offset, False),
                                      "version": 3, "strides": None}
        return torch.as_tensor(m, device=self.device)

    def _tensor(self, step: int = 0):
        return self._raw((int(step) % self.buffers) * self.stride, self.nbytes).view(self.world * self.per, self.rec_bytes)

    def begin(self, step: int, timeout_s: float = 5.0) -> int:
        """Every rank, before the first launch that writes step ``step``: hold this rank's stream until the owner has
        acknowledged step ``step - buffers`` (whose table is about to be overwritten); returns this rank's row pointer."""
        import ctypes
        if step > self.buffers:
            self.ctx.call("apda_peer_wait", ctypes.c_void_p(self.base + self.flags_off + 132), 1,
                          (int(step) - self.buffers) & 0xffffffff, float(timeout_s),
                          ctypes.c_void_p(self.base + self.flags_off + 256 + 4 * self.rank))
        return self.row_ptr(0, step)

    def signal(self, step: int):
        """Enqueue (after this rank's pickers, same stream): publish `step` - my rows of this step are in the table."""
        import ctypes
        self.ctx.call("apda_peer_signal", ctypes.c_void_p(self.base + self.flags_off + 4 * self.rank), int(step) & 0xffffffff)

    def wait(self, step: int, timeout_s: float = 5.0):
        """Owner only: hold the stream until every rank has published `step`; returns that step's table view, valid
        until ``release(step)``."""
        import ctypes
        assert self.owner
        self.ctx.call("apda_peer_wait", ctypes.c_void_p(self.base + self.flags_off), self.world, int(step) & 0xffffffff,
                      float(timeout_s), ctypes.c_void_p(self.base + self.flags_off + 128))
        return self._tensor(step)

    def release(self, step: int):
        """Owner only, enqueued after the consumer of ``wait(step)``'s table on the same stream: the table of ``step`` may
        be overwritten (by step ``step + buffers``)."""
        import ctypes
        assert self.owner
        self.ctx.call("apda_peer_signal", ctypes.c_void_p(self.base + self.flags_off + 132), int(step) & 0xffffffff)

    def timed_out(self) -> bool:
        """Owner only (synchronises): did the LAST wait of the owner, or the last back-pressure wait of any producer, give up?"""
        words = self._raw(self.flags_off, self.ctl_bytes).cpu().numpy().view("int32")
        return bool(words[32] != 0 or (words[64:64 + self.world] != 0).any())

    def complete(self, step: int = 0):
        """Every rank: wait for the local stream's kernels, then a process barrier; the owner gets the table view."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        return self._tensor(step) if self.owner else None

    def close(self):
        import ctypes
        if self._closed:
            return
        self._closed = True
        if self.owner:
            self.ctx.call("apda_peer_table_destroy", ctypes.c_void_p(self.base))
        else:
            self.ctx.call("apda_peer_table_close", ctypes.c_void_p(self.base))

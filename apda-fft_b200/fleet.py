"""Fleet sweep plumbing: windows are independent, so a batch is sharded contiguously over one process per GPU and
the only exchange is a gather of the fixed-size peak records to rank 0 (torch.distributed: NCCL on GPUs, gloo in the
CPU tests).  No collective touches the data path."""
from __future__ import annotations


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

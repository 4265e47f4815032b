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

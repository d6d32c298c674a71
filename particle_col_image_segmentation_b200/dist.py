"""Slice sharding across the GPUs of one box and the gather of the per-label tables.

Every 2-D slice is an independent problem (tiff_analysis.py:727-737 only ever sees
one image; split_zstack.py:52 iterates slices; labels restart at 1 per slice), so a
stack shards by contiguous blocks of slices with no halo and no label
reconciliation.  The only exchange is the per-label table: ranks all-gather their
row counts, pad to the largest and all-gather the rows (NCCL on device tensors; the
same code runs over gloo on CPU tensors in the tests).
"""

import torch
import torch.distributed as dist


def shard_range(n_slices, rank, world_size):
    """Contiguous block ``[z0, z1)`` of slices owned by ``rank`` (remainder to the low ranks)."""
    base, rem = divmod(n_slices, world_size)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def gather_tables(local_table, group=None):
    """All-gather ``(n_r, C)`` float64 tables into one ``(sum n_r, C)`` table in rank order.

    Rows keep their global ``z`` column, so the result equals the single-GPU table."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_table
    world = dist.get_world_size(group)
    if world == 1:
        return local_table
    dev = local_table.device
    n_local = torch.tensor([local_table.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap, local_table.shape[1]), dtype=local_table.dtype, device=dev)
    padded[: local_table.shape[0]] = local_table
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def gather_tables_padded(ftable, offsets, group=None):
    """The same gather for a table still in its padded device form (``SegmentResult.table_padded()``:
    ``(cap, C)`` rows of which the first ``offsets[-1]`` are valid).  Row counts are exchanged first and
    read back in ONE host synchronisation; every rank then contributes its first ``max(counts)`` rows."""
    n_local = offsets[-1:].to(torch.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ftable[: int(n_local.item())]
    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=torch.int64, device=ftable.device)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts = counts.cpu().tolist()
    cap = min(max(max(counts), 1), ftable.shape[0])
    flat = torch.empty((world * cap, ftable.shape[1]), dtype=ftable.dtype, device=ftable.device)
    dist.all_gather_into_tensor(flat, ftable[:cap].contiguous(), group=group)
    parts = flat.view(world, cap, ftable.shape[1])
    return torch.cat([parts[r, :c] for r, c in enumerate(counts)], dim=0)


def segment_zstack_sharded(stack_local, z0, group=None, **kwargs):
    """Run the pipeline on this rank's slices (global index of the first one: ``z0``) and
    gather the region tables.  Returns ``(SegmentResult for the local slices, global table)``."""
    from . import split_zstack

    res = split_zstack.segment_zstack_device(stack_local, z0=z0, **kwargs)
    return res, gather_tables(res.table_device(), group)

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


class GatheredTable:
    """Result of ``TableGather``: every rank's first ``cap`` rows and all row counts, still on the device.
    ``compact()`` waits for the exchange, reads the counts (the one host synchronisation) and returns the
    ``(sum n_r, C)`` table."""

    def __init__(self, parts, counts, cap, redo, ready=None):
        self.parts, self.counts, self.cap, self._redo, self.ready = parts, counts, cap, redo, ready

    def compact(self):
        if self.ready is not None:
            self.ready.synchronize()
        counts = self.counts.cpu().tolist()
        if max(counts) > self.cap:  # the speculative capacity was too small: exchange again with the true sizes
            return self._redo()
        return torch.cat([self.parts[r, :c] for r, c in enumerate(counts)], dim=0)


class TableGather:
    """All-gather of the padded per-rank region tables without a host synchronisation in the steady state,
    overlapped with the next step's kernels.

    The number of rows a rank contributes is only known on the device.  Instead of reading it back before
    every exchange (``gather_tables_padded``), the exchange ships the first ``cap`` rows of every rank, where
    ``cap`` is 1.25 x the largest count seen so far; the counts travel alongside and are checked when the
    table is consumed (``GatheredTable.compact``).  A count above ``cap`` redoes that exchange with the true
    sizes and raises the capacity, so the result is always exact.  On CUDA the rows are first copied to a
    staging buffer on the caller's stream and the two collectives run on a side stream, so the pipeline can
    overwrite its table for the next stack while the previous one is still travelling."""

    def __init__(self, group=None, cap=0):
        self.group, self.cap = group, int(cap)
        self.comm = self.done = self.stage = self.stage_n = None

    def __call__(self, ftable, offsets):
        n_local = offsets[-1:].to(torch.int64)
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        if self.cap <= 0:  # first use: learn the sizes (one synchronisation)
            if world > 1:
                counts = torch.empty(world, dtype=torch.int64, device=ftable.device)
                dist.all_gather_into_tensor(counts, n_local, group=self.group)
            else:
                counts = n_local
            self.cap = min(max(int(int(counts.max().item()) * 1.25) + 1, 16), ftable.shape[0])
            self.stage = None
        cap, C = self.cap, ftable.shape[1]

        def redo():
            self.cap = 0
            return gather_tables_padded(ftable, offsets, self.group)

        def exchange(rows, n):
            counts = torch.empty(world, dtype=torch.int64, device=ftable.device)
            flat = torch.empty((world * cap, C), dtype=ftable.dtype, device=ftable.device)
            if world > 1:
                dist.all_gather_into_tensor(counts, n, group=self.group)
                dist.all_gather_into_tensor(flat, rows, group=self.group)
            else:
                counts.copy_(n)
                flat.copy_(rows)
            return flat.view(world, cap, C), counts

        if not ftable.is_cuda:
            parts, counts = exchange(ftable[:cap].contiguous(), n_local)
            return GatheredTable(parts, counts, cap, redo)
        main = torch.cuda.current_stream()
        if self.comm is None:
            self.comm = torch.cuda.Stream(device=ftable.device)
        if self.stage is None:
            self.stage = torch.empty((cap, C), dtype=ftable.dtype, device=ftable.device)
            self.stage_n = torch.empty(1, dtype=torch.int64, device=ftable.device)
        if self.done is not None:
            main.wait_event(self.done)  # the previous exchange has read the staging buffers
        self.stage.copy_(ftable[:cap])
        self.stage_n.copy_(n_local)
        staged = torch.cuda.Event()
        staged.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(staged)
            parts, counts = exchange(self.stage, self.stage_n)
            self.done = torch.cuda.Event()
            self.done.record(self.comm)
        return GatheredTable(parts, counts, cap, redo, ready=self.done)


def segment_zstack_sharded(stack_local, z0, group=None, **kwargs):
    """Run the pipeline on this rank's slices (global index of the first one: ``z0``) and
    gather the region tables.  Returns ``(SegmentResult for the local slices, global table)``."""
    from . import split_zstack

    res = split_zstack.segment_zstack_device(stack_local, z0=z0, **kwargs)
    return res, gather_tables(res.table_device(), group)

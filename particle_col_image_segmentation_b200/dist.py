"""Slice sharding across the GPUs of one box and the gather of the per-label tables.

Every 2-D slice is an independent problem (tiff_analysis.py:727-737 only ever sees
one image; split_zstack.py:52 iterates slices; labels restart at 1 per slice), so a
stack shards by contiguous blocks of slices with no halo and no label
reconciliation.  The only exchange is the per-label table.  ``gather_tables`` /
``gather_tables_padded`` are the plain forms (counts, then rows padded to the largest);
``TableGather`` is what a running pipeline uses: one gather-to-root per step, no host
synchronisation, overlapped with the next step (NCCL on device tensors; the same code
runs over gloo on CPU tensors in the tests).
"""

import torch
import torch.distributed as dist

from . import _lib


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
        n = int(n_local.item())
        if n > ftable.shape[0]:
            raise _lib.PcsError(f"region table overflow: {n} regions, capacity {ftable.shape[0]}; raise max_regions_per_slice")
        return ftable[:n]
    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=torch.int64, device=ftable.device)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts = counts.cpu().tolist()
    if max(counts) > ftable.shape[0]:
        raise _lib.PcsError(f"region table overflow: {max(counts)} regions, capacity {ftable.shape[0]}; raise max_regions_per_slice")
    cap = min(max(max(counts), 1), ftable.shape[0])
    flat = torch.empty((world * cap, ftable.shape[1]), dtype=ftable.dtype, device=ftable.device)
    dist.all_gather_into_tensor(flat, ftable[:cap].contiguous(), group=group)
    parts = flat.view(world, cap, ftable.shape[1])
    return torch.cat([parts[r, :c] for r, c in enumerate(counts)], dim=0)


class TableStaging:
    """One message buffer of the table gather: ``hdr_rows`` header rows (the row count of every chunk as float64)
    followed by ``caps[i]`` rows per chunk.  ``rows[i]`` / ``counts[i]`` are views a ``SegmentPlan(staging=...)``
    finalises its tables into, so the exchange needs no copy on the pipeline's stream."""

    def __init__(self, caps, table_caps, C, device, dtype=torch.float64):
        self.caps, self.table_caps = [int(c) for c in caps], [int(c) for c in table_caps]
        nch = len(self.caps)
        self.hdr_rows = (nch + C - 1) // C
        self.buf = torch.zeros((self.hdr_rows + sum(self.caps), C), dtype=dtype, device=device)
        hdr = self.buf[: self.hdr_rows].view(-1)
        self.counts = [hdr[i : i + 1] for i in range(nch)]
        self.rows, at = [], self.hdr_rows
        for c in self.caps:
            self.rows.append(self.buf[at : at + c])
            at += c
        self.busy = None  # event of the exchange that last read this buffer


class GatheredTable:
    """Result of ``TableGather``: on the root (or on every rank with ``all_ranks=True``) the staged rows and row
    counts of all ranks, still on the device.  ``compact()`` waits for the exchange, reads the counts (the one host
    synchronisation) and returns the ``(sum n, C)`` table in (rank, chunk) order, i.e. global slice order; on ranks
    that received nothing it returns ``None``."""

    def __init__(self, owner, recv, caps, hdr_rows, table_caps, ready=None):
        self.owner, self.recv, self.caps, self.hdr_rows, self.table_caps, self.ready = owner, recv, caps, hdr_rows, table_caps, ready

    def compact(self):
        if self.ready is not None:
            self.ready.synchronize()
        if self.recv is None:
            return None
        nch, C = len(self.caps), self.recv.shape[2]
        counts = self.recv[:, : self.hdr_rows].reshape(self.recv.shape[0], -1)[:, :nch].cpu().to(torch.int64)
        for i, (cap, tcap) in enumerate(zip(self.caps, self.table_caps)):
            worst = int(counts[:, i].max())
            if worst > tcap:
                raise _lib.PcsError(f"region table overflow: {worst} regions in a chunk, capacity {tcap}; raise max_regions_per_slice")
            if worst > cap:
                # The staged copy holds only `cap` rows and the pipeline may have overwritten its table since: the rows
                # are gone.  Later exchanges use the larger capacity; this one has to be repeated by the caller.
                self.owner.caps = None
                self.owner.min_caps = [max(m, int(counts[:, j].max())) for j, m in enumerate(self.owner.min_caps or [0] * nch)]
                raise _lib.PcsError(f"table gather: a chunk holds {worst} regions, the speculative capacity was {cap}; "
                                    "the capacity has been raised, run the step and the exchange again")
        parts, row0 = [], self.hdr_rows
        starts = []
        for cap in self.caps:
            starts.append(row0)
            row0 += cap
        for r in range(self.recv.shape[0]):
            for i in range(nch):
                n = int(counts[r, i])
                if n:
                    parts.append(self.recv[r, starts[i] : starts[i] + n])
        if not parts:
            return torch.zeros((0, C), dtype=self.recv.dtype, device=self.recv.device)
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


class TableGather:
    """Gather of the per-rank region tables of one step: ONE collective per step, to the root, without a host
    synchronisation in the steady state and without making the pipeline wait for the previous exchange.

    ``gather(pads)`` takes the padded device tables of all chunks of a step (``SegmentResult.table_padded()``:
    ``(offsets, ftable)`` pairs, rows ``[0, offsets[-1])`` valid).  The number of valid rows is only known on the
    device, so every rank ships a fixed-size message: a header row with its row counts followed by the first
    ``cap_i`` rows of every chunk, where ``cap_i`` is 1.25 x the largest count seen when the sizes were learnt
    (first call: one all-reduce of the counts and one synchronisation).  Counts are checked when the table is
    consumed (``GatheredTable.compact``); a count above the shipped capacity raises -- the staged copy cannot be
    completed afterwards because the pipeline may already be overwriting its tables -- and enlarges the capacity
    for the following steps.  On CUDA the message is assembled in one of two staging buffers on the caller's stream
    and the collective (``dist.gather``: grouped NCCL send / recv, nothing lands on the non-root ranks) runs on a
    side stream; a staging buffer is reused only after the exchange that read it two steps earlier, so the kernels
    of step n + 1 overlap the exchange of step n.  ``all_ranks=True`` turns the gather into an all-gather."""

    def __init__(self, group=None, root=0, all_ranks=False, slack=1.25):
        self.group, self.root, self.all_ranks, self.slack = group, int(root), bool(all_ranks), float(slack)
        self.caps = self.min_caps = None
        self.comm = None
        self.stage, self.recv, self.done, self.k = [None, None], [None, None], [None, None], 0

    def _world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def _learn(self, pads):
        counts = torch.stack([off[-1].to(torch.int64) for off, _ in pads])
        if self._world() > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.MAX, group=self.group)
        worst = counts.cpu().tolist()  # the one synchronisation
        floor = self.min_caps or [0] * len(pads)
        self.caps = [min(max(int(max(w, m) * self.slack) + 16, 16), int(ft.shape[0])) for w, m, (_, ft) in zip(worst, floor, pads)]
        self.stage, self.recv, self.done = [None, None], [None, None], [None, None]

    def gather(self, pads):
        if isinstance(pads, tuple) and len(pads) == 2 and torch.is_tensor(pads[0]):
            pads = [pads]
        world = self._world()
        rank = dist.get_rank(self.group) if world > 1 else 0
        if self.caps is None or len(self.caps) != len(pads):
            self._learn(pads)
        caps, nch = self.caps, len(pads)
        ft0 = pads[0][1]
        C, dev = int(ft0.shape[1]), ft0.device
        hdr_rows = (nch + C - 1) // C
        rows = hdr_rows + sum(caps)
        i = self.k % 2
        self.k += 1
        cuda = ft0.is_cuda
        if cuda:
            main = torch.cuda.current_stream()
            if self.comm is None:
                self.comm = torch.cuda.Stream(device=dev)
            if self.done[i] is not None:
                main.wait_event(self.done[i])  # the exchange two steps back has read this staging buffer
        if self.stage[i] is None or self.stage[i].shape[0] != rows:
            self.stage[i] = torch.zeros((rows, C), dtype=ft0.dtype, device=dev)
            want = self.all_ranks or rank == self.root
            self.recv[i] = torch.empty((world, rows, C), dtype=ft0.dtype, device=dev) if want else None
        st, rv = self.stage[i], self.recv[i]
        hdr = st[:hdr_rows].view(-1)
        at = hdr_rows
        for j, ((off, ft), cap) in enumerate(zip(pads, caps)):
            hdr[j : j + 1].copy_(off[-1:])  # int32 -> float64, exact
            st[at : at + cap].copy_(ft[:cap])
            at += cap

        def exchange():
            if world == 1:
                rv[0].copy_(st)
            elif self.all_ranks:
                dist.all_gather_into_tensor(rv.view(world * rows, C), st, group=self.group)
            else:
                dist.gather(st, [rv[r] for r in range(world)] if rank == self.root else None, dst=dist.get_global_rank(self.group, self.root) if self.group is not None else self.root, group=self.group)

        ready = None
        if cuda:
            staged = torch.cuda.Event()
            staged.record(main)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(staged)
                exchange()
                ready = torch.cuda.Event()
                ready.record(self.comm)
            self.done[i] = ready
        else:
            exchange()
        return GatheredTable(self, rv, list(caps), hdr_rows, [int(ft.shape[0]) for _, ft in pads], ready)

    __call__ = gather

    def make_staging(self, counts_seen, table_caps, C, device, n=2):
        """``n`` message buffers (double buffering) sized for the largest per-chunk row counts seen so far
        (``counts_seen``: one int per chunk, already reduced over the ranks) plus the usual slack."""
        caps = [min(max(int(c * self.slack) + 16, 16), int(t)) for c, t in zip(counts_seen, table_caps)]
        return [TableStaging(caps, table_caps, C, device) for _ in range(n)]

    def exchange(self, stage):
        """Gather a message buffer a plan has just finalised its tables into (``SegmentPlan(staging=stage)``).  Nothing is
        copied and nothing runs on the caller's stream except an event record; before replaying the plan bound to
        ``stage`` again, the caller's stream must wait for ``stage.busy`` (``wait_free(stage)``)."""
        world = self._world()
        rank = dist.get_rank(self.group) if world > 1 else 0
        buf = stage.buf
        rows, C = int(buf.shape[0]), int(buf.shape[1])
        want = self.all_ranks or rank == self.root
        key = id(stage)
        if not hasattr(self, "_recv"):
            self._recv = {}
        if key not in self._recv:
            self._recv[key] = torch.empty((world, rows, C), dtype=buf.dtype, device=buf.device) if want else None
        rv = self._recv[key]

        def run():
            if world == 1:
                rv[0].copy_(buf)
            elif self.all_ranks:
                dist.all_gather_into_tensor(rv.view(world * rows, C), buf, group=self.group)
            else:
                dist.gather(buf, [rv[r] for r in range(world)] if rank == self.root else None, dst=dist.get_global_rank(self.group, self.root) if self.group is not None else self.root, group=self.group)

        ready = None
        if buf.is_cuda:
            main = torch.cuda.current_stream()
            if self.comm is None:
                self.comm = torch.cuda.Stream(device=buf.device)
            filled = torch.cuda.Event()
            filled.record(main)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(filled)
                run()
                ready = torch.cuda.Event()
                ready.record(self.comm)
            stage.busy = ready
        else:
            run()
        return GatheredTable(self, rv, list(stage.caps), stage.hdr_rows, list(stage.table_caps), ready)

    @staticmethod
    def wait_free(stage):
        if stage.busy is not None:
            torch.cuda.current_stream().wait_event(stage.busy)


def segment_zstack_sharded(stack_local, z0, group=None, **kwargs):
    """Run the pipeline on this rank's slices (global index of the first one: ``z0``) and
    gather the region tables.  Returns ``(SegmentResult for the local slices, global table)``."""
    from . import split_zstack

    res = split_zstack.segment_zstack_device(stack_local, z0=z0, **kwargs)
    return res, gather_tables(res.table_device(), group)

"""Z-stack layout (``split_zstack.py``) and the full per-slice segment pipeline.

``split_zstack.py:50-65`` reads a ``(Z, C, Y, X)`` TIFF stack and writes every
``z_slice[channel]`` plane to its own file; ilastik (an external GUI classifier)
then turns the planes into class images for ``tiff_analysis.py``.  Here the same
layout contract feeds the device pipeline directly:

    threshold (Otsu)  ->  5x5 median  ->  label  ->  regionprops
                      ->  refine (small objects out, holes filled)  ->  EDT

Each stage is the library call the reference makes at the cited line (or, for
Otsu / small-object removal, the one the north_star names); the CPU statement of
the same composition is ``oracle/pipeline.py``.  Every slice is an independent
2-D problem (tiff_analysis.py:727-737), so a stack is processed as batched
kernel launches over chunks of slices and shards across GPUs by slice with no
halo (``dist.py``).
"""

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _io, _lib, ops

CHANNEL_MAP_4 = {0: "CY5", 1: "RFP", 2: "GFP", 3: "DAPI"}  # split_zstack.py:39
CHANNEL_MAP_2 = {0: "RFP", 1: "GFP"}  # split_zstack.py:54

TABLE_COLUMNS = ("z", "label", "area", "centroid_y", "centroid_x", "min_row", "min_col", "max_row", "max_col", "first_row", "first_col", "intensity_sum", "intensity_mean")


def split_channels(zstack, channel_indices=(1, 2)):
    """split_zstack.py:52-65 without the file I/O: ``{channel name: (Z, Y, X) planes}``.
    A stack whose slices do not have 4 channels is treated as ``RFP, GFP`` (:53-55)."""
    if zstack.ndim != 4:
        raise ValueError(f"expected a (Z, C, Y, X) stack, got {zstack.shape}")
    cmap = CHANNEL_MAP_4
    if zstack.shape[1] != 4:
        cmap, channel_indices = CHANNEL_MAP_2, (0, 1)
    return {cmap[c]: zstack[:, c] for c in channel_indices}


def get_clean_file_name(input_file):
    """``(channel suffix, name without it)`` as split_zstack.py:21-32 derives them: the stem up to the
    first dot, minus the channel list and the ``_zstack`` / ``_mip`` markers."""
    stem = input_file.split(".")[0]
    for marker in ("CY5_RFP_GFP_DAPI", "RFP_GFP"):
        if marker + "_" in stem:
            suffix = "_" + marker
            return suffix, stem.replace(suffix, "").replace("_zstack", "").replace("_mip", "")
    return "", stem


def channel_folder_name(destination, used_channels, channel_name):
    """Folder a channel's planes go to (split_zstack.py:34-38, without creating it)."""
    return destination.replace(".tif", "").replace("_mip", "").replace(used_channels, "") + "_" + channel_name


def process_tif(input_file, channel_indices=(1, 2), move=True):
    """split_zstack.py:40-65 with this repo's TIFF reader / writer: the stack file is moved into a folder
    named after it, read as ``(Z, C, Y, X)`` and every selected ``z_slice[channel]`` plane is written to
    ``<folder>_<channel>/<name>_z<i>_<channel>.tif``.  Returns the list of files written.
    ``move=False`` leaves the input where it is (the planes still go next to the would-be destination)."""
    import os

    from . import tiff_io

    stem_end = input_file.split("/")[-1].split(".")[0]
    used_channels, clean = get_clean_file_name(input_file)
    os.makedirs(clean, exist_ok=True)
    destination = os.path.join(clean, os.path.basename(input_file))
    if move:
        os.rename(input_file, destination)
    if not input_file.endswith(".tif"):
        return []
    zstack = tiff_io.read_stack(destination if move else input_file)
    if zstack.ndim != 4:
        raise ValueError(f"expected a (Z, C, Y, X) stack in {input_file}, got shape {zstack.shape}")
    planes = split_channels(zstack, tuple(channel_indices))
    written = []
    for name, stack in planes.items():
        folder = channel_folder_name(destination, used_channels, name)
        os.makedirs(folder, exist_ok=True)
        for z in range(stack.shape[0]):
            out = os.path.join(folder, plane_name(stem_end.replace(used_channels, ""), z, name))
            tiff_io.write_plane(out, stack[z])
            written.append(out)
    return written


def segment_tif(path, channel=1, outputs=("mask", "labels", "refined", "edt"), **kwargs):
    """Read a ``(Z, C, Y, X)`` (or ``(Z, Y, X)``) uint16 stack file into pinned memory and run the segment
    pipeline on one channel; returns the dict of pinned host tensors of ``segment_zstack_pinned``."""
    from . import tiff_io

    t = tiff_io.read_stack_pinned(path)
    if t.dim() == 4:
        t = t[:, channel].contiguous()
        if torch.cuda.is_available():
            t = t.pin_memory()
    if t.dim() != 3 or t.dtype != torch.uint16:
        raise _lib.PcsError(f"segment_tif expects a uint16 (Z, [C,] Y, X) stack, got {tuple(t.shape)} {t.dtype}")
    out = alloc_host_outputs(*t.shape, outputs=outputs)
    segment_zstack_pinned(t, out, outputs=outputs, **kwargs)
    return out


def plane_name(base, z, channel):
    """File name the reference would give the plane (split_zstack.py:63)."""
    return f"{base}_z{z}_{channel}.tif"


@dataclass
class SegmentResult:
    """Device-resident outputs of ``segment_zstack_device`` (16 B / voxel)."""

    mask: torch.Tensor  # (Z, H, W) uint8
    labels: torch.Tensor  # (Z, H, W) int32
    refined: torch.Tensor  # (Z, H, W) uint8
    edt: torch.Tensor  # (Z, H, W) float64
    threshold: torch.Tensor  # (Z,) int32
    counts: torch.Tensor  # (Z,) int32
    tables: list = field(default_factory=list)  # per chunk: (z0, offsets[B+1], int64 table, float64 (cap, 13) table)
    z0: int = 0

    def table_padded(self):
        """Per chunk ``(offsets[B + 1], float64 (cap, 13) table)`` as left on the device by the pipeline:
        rows ``[0, offsets[-1])`` are valid.  No host synchronisation."""
        return [(offsets, ftable) for _, offsets, _, ftable in self.tables]

    def table_device(self):
        """Compact ``(n, 13)`` float64 table on the device (one sync to learn the sizes)."""
        parts = []
        for _, offsets, table, ftable in self.tables:
            n = int(offsets[-1])
            if n > table.shape[1]:
                raise _lib.PcsError(f"region table overflow: {n} regions in a chunk, capacity {table.shape[1]}; raise max_regions_per_slice")
            if n > ftable.shape[0]:
                raise _lib.PcsError(f"the staged table holds {ftable.shape[0]} rows, the chunk has {n} regions: enlarge the gather's staging")
            if n:
                parts.append(ftable[:n])
        if not parts:
            return torch.zeros((0, len(TABLE_COLUMNS)), dtype=torch.float64, device=self.labels.device)
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)

    def to_numpy(self):
        return {
            "threshold": self.threshold.cpu().numpy().astype(np.int64),
            "mask": self.mask.cpu().numpy().astype(bool),
            "labels": self.labels.cpu().numpy(),
            "refined": self.refined.cpu().numpy().astype(bool),
            "edt": self.edt.cpu().numpy(),
            "table": self.table_device().cpu().numpy(),
            "counts": self.counts.cpu().numpy().astype(np.int64),
        }


class SegmentPlan:
    """The pipeline bound to fixed device buffers, optionally captured in a CUDA graph.

    One ``pcs_segment_chunk`` call per chunk of slices is enqueued on the current stream; with
    ``graph=True`` the calls are captured once and replayed, which removes the launch gaps
    between the ~34 kernels of a chunk (a z-stack is many small 2-D problems, so launch
    latency matters).  ``plan()`` is asynchronous and returns the bound ``SegmentResult``.
    """

    def __init__(self, stack, denoise_size=5, min_size=20, chunk=16, max_regions_per_slice=1 << 14, out=None, z0=0, graph=False, streams=1, staging=None):
        ops.require_cuda(stack, "stack")
        if stack.dtype != torch.uint16 or stack.dim() != 3:
            raise _lib.PcsError(f"segment pipeline expects a (Z, H, W) uint16 tensor, got {tuple(stack.shape)} {stack.dtype}")
        self.stack = stack
        Z, H, W = (int(v) for v in stack.shape)
        dev = stack.device
        if out is None:
            out = SegmentResult(
                mask=torch.empty((Z, H, W), dtype=torch.uint8, device=dev),
                labels=torch.empty((Z, H, W), dtype=torch.int32, device=dev),
                refined=torch.empty((Z, H, W), dtype=torch.uint8, device=dev),
                edt=torch.empty((Z, H, W), dtype=torch.float64, device=dev),
                threshold=torch.empty(Z, dtype=torch.int32, device=dev),
                counts=torch.empty(Z, dtype=torch.int32, device=dev),
            )
        self.out = out
        out.z0 = z0
        self.dn, self.ms = int(denoise_size or 0), int(min_size or 0)
        self.lib = _lib.load()
        chunk = max(1, min(int(chunk), Z))
        self.chunk = chunk
        out.tables = []
        self.calls = []
        # staging (dist.TableStaging, optional): the float64 rows of chunk i are finalised straight into the message buffer of
        # the table gather -- staging.rows[i], with the true count at staging.counts[i] -- instead of a private table that
        # the exchange would have to copy on the pipeline's stream
        self.staging = staging
        self.fin = []
        nws = self.lib.pcs_segment_workspace_bytes(chunk, H, W)
        n_chunks = (Z + chunk - 1) // chunk
        # chunks are independent (slices are): with several streams their kernels overlap, which fills
        # the SMs while the latency-bound union-find / search kernels of another chunk are in flight
        self.n_streams = max(1, min(int(streams), n_chunks))
        self.side = [torch.cuda.Stream(device=dev) for _ in range(self.n_streams - 1)]
        self.ws = [torch.empty(nws, dtype=torch.uint8, device=dev) for _ in range(self.n_streams)]  # owned: stable pointers for graph replays
        P_ = ops._p
        for i, a in enumerate(range(0, Z, chunk)):
            b = min(Z, a + chunk)
            B = b - a
            offsets = torch.empty(B + 1, dtype=torch.int32, device=dev)
            cap = int(max_regions_per_slice) * B
            table = torch.empty((ops.TABLE_COLS, cap), dtype=torch.int64, device=dev)
            if staging is not None:
                if len(staging.rows) != n_chunks:
                    raise _lib.PcsError(f"staging was made for {len(staging.rows)} chunks, the plan has {n_chunks}")
                ftable = staging.rows[i]
                self.fin.append((P_(table), cap, P_(offsets), B, W, float(z0 + a), P_(ftable), int(ftable.shape[0]), P_(staging.counts[i])))
            else:
                ftable = torch.empty((cap, len(TABLE_COLUMNS)), dtype=torch.float64, device=dev)
            out.tables.append((a, offsets, table, ftable))
            P = ops._p
            ws = self.ws[i % self.n_streams]
            self.calls.append((P(stack[a:b]), B, H, W, self.dn, self.ms, P(out.mask[a:b]), P(out.labels[a:b]), P(out.refined[a:b]), P(out.edt[a:b]),
                               P(out.threshold[a:b]), P(out.counts[a:b]), P(offsets), P(table), cap, 0 if staging is not None else P(ftable), int(z0 + a), P(ws), nws))
        self.graph = None
        if graph:
            self._enqueue()  # warm-up: one-time initialisations must not land in the capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g

    def _enqueue(self):
        main = torch.cuda.current_stream()
        for s in self.side:
            s.wait_stream(main)  # fork
        for i, c in enumerate(self.calls):
            j = i % self.n_streams
            st = main if j == 0 else self.side[j - 1]
            with torch.cuda.stream(st):
                _lib.check(self.lib.pcs_segment_chunk(*c, st.cuda_stream), "pcs_segment_chunk")
                if self.fin:  # rows + count of this chunk into the gather's message buffer
                    _lib.check(self.lib.pcs_table_finalize_ex(*self.fin[i], st.cuda_stream), "pcs_table_finalize_ex")
        for s in self.side:
            main.wait_stream(s)  # join

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        return self.out


def segment_zstack_device(stack, denoise_size=5, min_size=20, chunk=16, max_regions_per_slice=1 << 14, out=None, z0=0):
    """Full pipeline over a device-resident ``(Z, H, W)`` uint16 stack.

    Asynchronous on the current stream; no host synchronisation inside.  ``out`` may
    carry preallocated output tensors (a previous ``SegmentResult``) to reuse.
    """
    return SegmentPlan(stack, denoise_size, min_size, chunk, max_regions_per_slice, out=out, z0=z0)()


def segment_zstack(stack, denoise_size=5, min_size=20, chunk=16, max_regions_per_slice=1 << 14, z0=0):
    """numpy ``(Z, H, W)`` (or a single ``(H, W)`` slice) in, dict of numpy arrays out:
    ``threshold, mask, labels, refined, edt, table, counts`` -- the same keys and dtypes
    as ``oracle.pipeline.segment_zstack``."""
    np_in = _io.is_numpy(stack)
    t = _io.to_device(stack)
    single = t.dim() == 2
    if single:
        t = t.unsqueeze(0)
    res = segment_zstack_device(t, denoise_size, min_size, chunk, max_regions_per_slice, z0=z0)
    if not np_in:
        return res
    d = res.to_numpy()
    if single:
        d = {k: (v[0] if k not in ("table",) else v) for k, v in d.items()}
    return d


# ---------------------------------------------------------------- host-buffer (end-to-end) path
_E2E_CACHE = {}


IMAGE_OUTPUTS = {"mask": torch.uint8, "labels": torch.int32, "refined": torch.uint8, "edt": torch.float64}


def alloc_host_outputs(Z, H, W, outputs=("mask", "labels", "refined", "edt")):
    """Pinned host buffers for the chosen image outputs of a ``(Z, H, W)`` stack (plus the per-slice threshold and
    region count, which always come back, like the region table).  What is not listed never crosses PCIe: the
    float64 EDT alone is 8 of the 14 output bytes per voxel."""
    for k in outputs:
        if k not in IMAGE_OUTPUTS:
            raise ValueError(f"unknown output {k!r}; choose from {sorted(IMAGE_OUTPUTS)}")
    out = {k: torch.empty((Z, H, W), dtype=IMAGE_OUTPUTS[k]).pin_memory() for k in outputs}
    out["threshold"] = torch.empty(Z, dtype=torch.int32).pin_memory()
    out["counts"] = torch.empty(Z, dtype=torch.int32).pin_memory()
    return out


def segment_zstack_pinned(host_in, host_out, denoise_size=5, min_size=20, chunk=16, max_regions_per_slice=1 << 14, outputs=None):
    """Host buffers in, host buffers out.  The stack moves through the device chunk by chunk on three
    streams -- H2D copy of chunk i+1, the pipeline on chunk i and the D2H copy of the outputs of chunk
    i-1 overlap, so the call costs about as much as its largest leg (with every output: the 14 B/voxel
    coming back over PCIe).  ``outputs`` picks the image outputs that are copied back (default: the
    ones ``host_out`` has buffers for); the table, thresholds and counts always are.  Device buffers
    are cached between calls.  Returns the number of table rows (``host_out['table']`` holds them)."""
    if outputs is None:
        outputs = tuple(k for k in ("edt", "labels", "mask", "refined") if k in host_out)
    for k in outputs:
        if k not in IMAGE_OUTPUTS or k not in host_out:
            raise ValueError(f"output {k!r} is unknown or has no buffer in host_out")
    dev = _io.device()
    key = (str(dev), tuple(host_in.shape), denoise_size, min_size, chunk, max_regions_per_slice)
    st = _E2E_CACHE.get(key)
    if st is None:
        d_in = torch.empty(tuple(host_in.shape), dtype=torch.uint16, device=dev)
        plan = SegmentPlan(d_in, denoise_size, min_size, chunk, max_regions_per_slice)
        n = len(plan.calls)
        st = {"in": d_in, "plan": plan, "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev),
              "ev_in": [torch.cuda.Event() for _ in range(n)], "ev_out": [torch.cuda.Event() for _ in range(n)]}
        _E2E_CACHE.clear()
        _E2E_CACHE[key] = st
    plan, d_in = st["plan"], st["in"]
    res = plan.out
    main = torch.cuda.current_stream()
    st["h2d"].wait_stream(main)
    st["d2h"].wait_stream(main)
    Z = host_in.shape[0]
    for i, c in enumerate(plan.calls):
        a, b = i * plan.chunk, min(Z, (i + 1) * plan.chunk)
        with torch.cuda.stream(st["h2d"]):
            d_in[a:b].copy_(host_in[a:b], non_blocking=True)
            st["ev_in"][i].record()
        main.wait_event(st["ev_in"][i])
        _lib.check(plan.lib.pcs_segment_chunk(*c, main.cuda_stream), "pcs_segment_chunk")
        st["ev_out"][i].record(main)
        with torch.cuda.stream(st["d2h"]):
            st["d2h"].wait_event(st["ev_out"][i])
            for k in outputs:
                host_out[k][a:b].copy_(getattr(res, k)[a:b], non_blocking=True)
    main.wait_stream(st["d2h"])
    for k in ("threshold", "counts"):
        host_out[k].copy_(getattr(res, k), non_blocking=True)
    table = res.table_device()
    host_out["table"] = table.cpu()
    torch.cuda.synchronize()
    return int(table.shape[0])

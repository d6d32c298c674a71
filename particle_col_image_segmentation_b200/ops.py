"""Batched device-level operators over ``(B, H, W)`` torch CUDA tensors.

Thin, allocation-explicit wrappers over the C ABI (``include/pcs.h``).  torch is
used only to own device buffers and the stream; every computation is a libpcs
kernel.  Binary masks travel as *bit images*: ``uint32`` tensors of shape
``(B, H, ceil(W/32))`` viewed here as ``int32`` (torch has no arithmetic on
uint32, none is needed).
"""

import numpy as np
import torch

from . import _lib

TABLE_COLS = 10
T_AREA, T_SUMY, T_SUMX, T_MINY, T_MINX, T_MAXY, T_MAXX, T_FIRST, T_SUMI, T_OVERLAP = range(10)
CMP = {">": 0, ">=": 1, "<": 2, "<=": 3, "==": 4, "!=": 5}
_DTYPE_CODE = {torch.uint8: 0, torch.uint16: 1, torch.int32: 2, torch.float32: 3, torch.float64: 4, torch.int64: 5, torch.bool: 0}


def _stream():
    # the raw handle of torch's current stream; ~10x cheaper than building a torch.cuda.Stream object per launch, which
    # mattered once the drop-in functions were down to ~60 launches of a few microseconds each
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _p(t):
    return 0 if t is None else t.data_ptr()


def words(W):
    return (W + 31) // 32


def require_cuda(t, name="input"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PcsError(f"{name} must be a CUDA tensor: this package has no CPU path")
    if not t.is_contiguous():
        raise _lib.PcsError(f"{name} must be contiguous")
    return t


def _bhw(t):
    if t.dim() != 3:
        raise _lib.PcsError(f"expected a (B, H, W) tensor, got {tuple(t.shape)}")
    return int(t.shape[0]), int(t.shape[1]), int(t.shape[2])


def new_bits(B, H, W, device):
    return torch.empty((B, H, words(W)), dtype=torch.int32, device=device)


class Workspace:
    """Grow-only scratch buffer reused across calls (the library never allocates)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != torch.device(device):
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self.buf


_default_ws = {}


def _ws(nbytes, device, slot="main"):
    key = (str(device), slot)
    w = _default_ws.setdefault(key, Workspace())
    return w.get(nbytes, device)


# ---------------------------------------------------------------- K2
def compare(img, op, thr, thr_dev=None, want_bits=True, want_mask=False):
    """bits / uint8 mask of ``img <op> thr`` (per-slice thresholds if ``thr_dev``)."""
    require_cuda(img)
    B, H, W = _bhw(img)
    bits = new_bits(B, H, W, img.device) if want_bits else None
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=img.device) if want_mask else None
    fn = {torch.uint8: "pcs_compare_u8", torch.uint16: "pcs_compare_u16", torch.int32: "pcs_compare_i32", torch.float32: "pcs_compare_f32", torch.float64: "pcs_compare_f64"}.get(img.dtype)
    if fn is None:
        raise _lib.PcsError(f"compare: unsupported dtype {img.dtype}")
    if img.dtype in (torch.float32, torch.float64):
        thr = float(thr)
    else:
        thr = int(thr)
    _lib.call(fn, _p(img), thr, _p(thr_dev), CMP[op], _p(bits), _p(mask), B, H, W, _stream())
    return bits, mask


def member_u8(img, values, want_bits=True, want_mask=False):
    require_cuda(img)
    B, H, W = _bhw(img)
    tab = np.zeros(256, dtype=np.uint8)
    for v in values:
        if 0 <= int(v) < 256:
            tab[int(v)] = 1
    tab_d = torch.from_numpy(tab).to(img.device)
    bits = new_bits(B, H, W, img.device) if want_bits else None
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=img.device) if want_mask else None
    _lib.call("pcs_member_u8", _p(img), _p(tab_d), _p(bits), _p(mask), B, H, W, _stream())
    return bits, mask


def pack(mask):
    """uint8 / bool image -> bit image (non-zero = set)."""
    require_cuda(mask)
    m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    return compare(m, "!=", 0)[0]


def unpack(bits, W, dtype=torch.uint8):
    require_cuda(bits)
    B, H, _ = bits.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=bits.device)
    _lib.call("pcs_unpack_bits", _p(bits), _p(out), B, H, W, _stream())
    return out.view(torch.bool) if dtype == torch.bool else out


def logic(a, b, op, W):
    code = {"and": 0, "or": 1, "andnot": 2, "xor": 3, "not": 4}[op]
    out = torch.empty_like(a)
    B, H, _ = a.shape
    _lib.call("pcs_bits_logic", _p(a), _p(b), _p(out), code, B, H, W, _stream())
    return out


def count(bits, W):
    B, H, _ = bits.shape
    out = torch.empty(B, dtype=torch.int64, device=bits.device)
    _lib.call("pcs_bits_count", _p(bits), _p(out), B, H, W, _stream())
    return out


def lut_u8_(img, lut):
    """In-place 256-entry LUT."""
    require_cuda(img)
    lut_d = torch.as_tensor(np.asarray(lut, dtype=np.uint8)).to(img.device)
    _lib.call("pcs_lut_u8", _p(img), _p(lut_d), img.numel(), _stream())
    return img


def assign_where_u8_(img, bits, value):
    B, H, W = _bhw(img)
    _lib.call("pcs_assign_where_u8", _p(img), _p(bits), int(value), B, H, W, _stream())
    return img


def max_label(labels):
    """Largest label of an int32 / int64 label image (one reduction kernel, one read-back)."""
    out = torch.empty(1, dtype=torch.int64, device=labels.device)
    _lib.call("pcs_max_label", _p(labels), 4 if labels.dtype == torch.int32 else 8, int(labels.numel()), _p(out), _stream())
    return int(out.item())


def transpose2d(t):
    """(H, W) uint8 / bool / int32 CUDA tensor -> contiguous (W, H) transpose."""
    H, W = (int(v) for v in t.shape)
    src = t.view(torch.uint8) if t.dtype == torch.bool else t
    if src.element_size() not in (1, 4):
        raise _lib.PcsError(f"transpose2d: unsupported dtype {t.dtype}")
    out = torch.empty((W, H), dtype=src.dtype, device=t.device)
    _lib.call("pcs_transpose", _p(src.contiguous()), _p(out), src.element_size(), H, W, _stream())
    return out


def gather(img, slice_idx, lin_idx):
    """``img[slice_idx, lin_idx]`` over flattened slices -> int64."""
    n = int(lin_idx.numel())
    out = torch.empty(n, dtype=torch.int64, device=img.device)
    if n:
        slice_elems = img[0].numel()
        _lib.call("pcs_gather", _p(img), _DTYPE_CODE[img.dtype], _p(slice_idx), _p(lin_idx), _p(out), n, slice_elems, _stream())
    return out


# ---------------------------------------------------------------- K1
def otsu_u16(img, return_hist=False):
    """Per-slice Otsu threshold (int32 tensor of length B)."""
    require_cuda(img)
    B, H, W = _bhw(img)
    hist = torch.empty((B, 65536), dtype=torch.int32, device=img.device)
    thr = torch.empty(B, dtype=torch.int32, device=img.device)
    _lib.call("pcs_histogram_u16", _p(img), _p(hist), B, H, W, _stream())
    _lib.call("pcs_otsu_u16", _p(hist), _p(thr), 0, B, H * W, _stream())
    return (thr, hist) if return_hist else thr


# ---------------------------------------------------------------- K3
def median_u8(img, size=5):
    require_cuda(img)
    B, H, W = _bhw(img)
    out = torch.empty_like(img)
    _lib.call("pcs_median_u8", _p(img), _p(out), int(size), B, H, W, _stream())
    return out


def majority(bits, W, size=5):
    B, H, _ = bits.shape
    out = torch.empty_like(bits)
    _lib.call("pcs_majority_bits", _p(bits), _p(out), int(size), B, H, W, _stream())
    return out


# ---------------------------------------------------------------- K4
def _ccl_ws(B, H, W, aux, device):
    n = _lib.load().pcs_ccl_workspace_bytes(B, H, W, int(aux))
    return _ws(n, device, "ccl"), n


def label_bits(bits, W, connectivity=8, invert=False, dtype=torch.int32, first_out=None, cap=0):
    """Label a bit image.  Returns ``(labels, counts[B], offsets[B+1])``."""
    B, H, _ = bits.shape
    labels = torch.empty((B, H, W), dtype=dtype, device=bits.device)
    counts = torch.empty(B, dtype=torch.int32, device=bits.device)
    offsets = torch.empty(B + 1, dtype=torch.int32, device=bits.device)
    ws, n = _ccl_ws(B, H, W, 0, bits.device)
    _lib.call("pcs_label_bits", _p(bits), B, H, W, connectivity, int(invert), _p(labels), labels.element_size(), _p(counts), _p(offsets), _p(first_out), int(cap), _p(ws), n, _stream())
    return labels, counts, offsets


def conn_planes(img, all_fg=False, connectivity=8, want_higher=False):
    require_cuda(img)
    B, H, W = _bhw(img)
    planes = torch.empty((6, B, H, words(W)), dtype=torch.int32, device=img.device)
    higher = new_bits(B, H, W, img.device) if want_higher else None
    src = img.view(torch.uint8) if img.dtype == torch.bool else img
    _lib.call("pcs_conn_planes", _p(src), _DTYPE_CODE[img.dtype], _p(planes), _p(higher), int(all_fg), connectivity, B, H, W, _stream())
    return planes, higher


def label_values(img, connectivity=8, dtype=torch.int64, first_out=None, cap=0):
    """Multi-valued labelling (equal non-zero neighbours).  ``(labels, counts, offsets)``."""
    B, H, W = _bhw(img)
    planes, _ = conn_planes(img, all_fg=False, connectivity=connectivity)
    labels = torch.empty((B, H, W), dtype=dtype, device=img.device)
    counts = torch.empty(B, dtype=torch.int32, device=img.device)
    offsets = torch.empty(B + 1, dtype=torch.int32, device=img.device)
    ws, n = _ccl_ws(B, H, W, 0, img.device)
    _lib.call("pcs_label_conn", _p(planes), B, H, W, connectivity, _p(labels), labels.element_size(), _p(counts), _p(offsets), _p(first_out), int(cap), _p(ws), n, _stream())
    return labels, counts, offsets


def fill_holes(bits, W):
    B, H, _ = bits.shape
    out = torch.empty_like(bits)
    ws, n = _ccl_ws(B, H, W, 0, bits.device)
    _lib.call("pcs_fill_holes_bits", _p(bits), _p(out), B, H, W, _p(ws), n, _stream())
    return out


def remove_small(bits, W, min_size, connectivity=4):
    B, H, _ = bits.shape
    out = torch.empty_like(bits)
    ws, n = _ccl_ws(B, H, W, 1, bits.device)
    _lib.call("pcs_remove_small_bits", _p(bits), _p(out), B, H, W, connectivity, int(min_size), _p(ws), n, _stream())
    return out


def select_components(bits, seeds, W, connectivity=8):
    B, H, _ = bits.shape
    out = torch.empty_like(bits)
    ws, n = _ccl_ws(B, H, W, 0, bits.device)
    _lib.call("pcs_select_components_bits", _p(bits), _p(seeds), _p(out), B, H, W, connectivity, _p(ws), n, _stream())
    return out


def local_maxima(img, connectivity=8):
    """Plateau maxima of a (B, H, W) image as a bit image."""
    B, H, W = _bhw(img)
    planes, higher = conn_planes(img, all_fg=True, connectivity=connectivity, want_higher=True)
    out = new_bits(B, H, W, img.device)
    counts = torch.empty(B, dtype=torch.int32, device=img.device)
    ws, n = _ccl_ws(B, H, W, 0, img.device)
    _lib.call("pcs_local_maxima_conn", _p(planes), _p(higher), _p(out), _p(counts), B, H, W, connectivity, _p(ws), n, _stream())
    return out


# ---------------------------------------------------------------- K5
def footprint_runs(footprint, reflect=False):
    """Flat footprint -> int32[n][3] runs ``(dy, lo, hi)`` (pieces of at most 32)."""
    fp = np.asarray(footprint) != 0
    if fp.ndim != 2:
        raise _lib.PcsError("footprint must be 2-D")
    if reflect:
        fp = fp[::-1, ::-1]
        cy, cx = (fp.shape[0] - 1) - fp.shape[0] // 2, (fp.shape[1] - 1) - fp.shape[1] // 2
    else:
        cy, cx = fp.shape[0] // 2, fp.shape[1] // 2
    runs = []
    for i in range(fp.shape[0]):
        row = np.concatenate([[False], fp[i], [False]])
        d = np.diff(row.astype(np.int8))
        for a, b in zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1)):
            lo, hi = a - cx, b - 1 - cx
            while lo <= hi:
                runs.append((i - cy, lo, min(hi, lo + 31)))
                lo += 32
    return np.asarray(runs, dtype=np.int32).reshape(-1, 3)


def dilate(bits, W, footprint, border_value=0):
    B, H, _ = bits.shape
    runs = footprint_runs(footprint)
    out = torch.empty_like(bits)
    if len(runs) == 0:
        return out.zero_()
    runs_d = torch.from_numpy(runs).to(bits.device)
    _lib.call("pcs_dilate_bits", _p(bits), _p(out), _p(runs_d), len(runs), 0, int(bool(border_value)), 0, B, H, W, _stream())
    return out


def erode(bits, W, footprint, border_value=0):
    B, H, _ = bits.shape
    runs = footprint_runs(footprint, reflect=True)
    out = torch.empty_like(bits)
    if len(runs) == 0:
        return logic(out.zero_(), None, "not", W)
    runs_d = torch.from_numpy(runs).to(bits.device)
    _lib.call("pcs_dilate_bits", _p(bits), _p(out), _p(runs_d), len(runs), 1, int(not border_value), 1, B, H, W, _stream())
    return out


# ---------------------------------------------------------------- K7
def edt(bits, W, invert=False, want_dist=True, want_sq=False, thr_sq=None):
    """Exact EDT of ``bits ^ invert``.  Returns ``(dist f64, sq i32, thr_bits)`` (None if not requested)."""
    B, H, _ = bits.shape
    dev = bits.device
    dist = torch.empty((B, H, W), dtype=torch.float64, device=dev) if want_dist else None
    sq = torch.empty((B, H, W), dtype=torch.int32, device=dev) if want_sq else None
    tb = new_bits(B, H, W, dev) if thr_sq is not None else None
    n = _lib.load().pcs_edt_workspace_bytes(B, H, W)
    ws = _ws(n, dev, "edt")
    _lib.call("pcs_edt_bits", _p(bits), int(invert), B, H, W, _p(dist), _p(sq), _p(tb), int(thr_sq or 0), _p(ws), n, _stream())
    return dist, sq, tb


def dilate_disk(bits, W, radius):
    """``binary_dilation(mask, disk(r))`` as ``EDT(~mask)^2 <= r^2`` (bit-exact)."""
    return edt(bits, W, invert=True, want_dist=False, thr_sq=int(radius) * int(radius))[2]


# ---------------------------------------------------------------- K8
def new_table(cap, device):
    t = torch.empty((TABLE_COLS, int(cap)), dtype=torch.int64, device=device)
    _lib.call("pcs_table_init", _p(t), int(cap), _stream())
    return t


def region_table(labels, offsets, table, intensity=None, fg_bits=None, ov_bits=None):
    B, H, W = _bhw(labels)
    idt = -1
    if intensity is not None:
        idt = {torch.uint8: 0, torch.uint16: 1}.get(intensity.dtype)
        if idt is None:
            raise _lib.PcsError(f"region_table: unsupported intensity dtype {intensity.dtype}")
    _lib.call("pcs_region_table", _p(labels), labels.element_size(), _p(intensity), idt, _p(fg_bits), _p(ov_bits), _p(offsets), _p(table), int(table.shape[1]), B, H, W, _stream())
    return table


def refine_labeled(bits, labels, table, offsets, min_size, W, want_mask=False):
    """``remove_small_objects(min_size, 8-connected)`` then ``binary_fill_holes`` of a bit image whose
    components are already labelled (``pcs_refine_labeled_bits``).  Returns bits (and the uint8 mask)."""
    B, H, _ = bits.shape
    lib = _lib.load()
    n = lib.pcs_fill_holes_table_workspace_bytes(B, H, W)
    ws = torch.empty(n, dtype=torch.uint8, device=bits.device)
    out = torch.empty_like(bits)
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=bits.device) if want_mask else None
    _lib.call("pcs_refine_labeled_bits", _p(bits), _p(labels), _p(table), int(table.shape[1]) if table is not None else 0, _p(offsets), int(min_size),
              _p(out), _p(mask), B, H, W, _p(ws), n, _stream())
    return (out, mask) if want_mask else out


def select_labels(labels, keep):
    """Bit image of pixels whose label is flagged in ``keep`` (uint8 ``(B, n)``)."""
    B, H, W = _bhw(labels)
    out = new_bits(B, H, W, labels.device)
    _lib.call("pcs_select_labels", _p(labels), labels.element_size(), _p(keep), int(keep.shape[1]), _p(out), B, H, W, _stream())
    return out


def select_by_area(labels, fg_bits, table, offsets, min_size):
    B, H, W = _bhw(labels)
    out = torch.empty_like(fg_bits)
    _lib.call("pcs_select_by_area", _p(labels), _p(fg_bits), _p(table), int(table.shape[1]), _p(offsets), int(min_size), _p(out), B, H, W, _stream())
    return out


def roi_sums(labels, planes, n_rois):
    K = int(planes.shape[0])
    out = torch.zeros((n_rois, K), dtype=torch.float64, device=labels.device)
    _lib.call("pcs_roi_sums_f64", _p(labels), _p(planes), K, int(labels.numel()), int(n_rois), _p(out), _stream())
    return out


def min_dist(a_xy, b_xy):
    a_xy, b_xy = a_xy.contiguous(), b_xy.contiguous()
    out = torch.empty(int(a_xy.shape[0]), dtype=torch.float64, device=a_xy.device)
    _lib.call("pcs_min_dist_f64", _p(a_xy), int(a_xy.shape[0]), _p(b_xy), int(b_xy.shape[0]), _p(out), _stream())
    return out


def nearest(a_xy, b_xy, exclude_self=False):
    """Distance to and index of the nearest ``b`` point for every ``a`` point (``(n, 2)`` float64 CUDA
    tensors); ``exclude_self`` skips ``j == i`` (use with ``b_xy is a_xy``)."""
    a_xy, b_xy = a_xy.contiguous(), b_xy.contiguous()
    na = int(a_xy.shape[0])
    d = torch.empty(na, dtype=torch.float64, device=a_xy.device)
    j = torch.empty(na, dtype=torch.int64, device=a_xy.device)
    _lib.call("pcs_nearest_f64", _p(a_xy), na, _p(b_xy), int(b_xy.shape[0]), int(bool(exclude_self)), _p(d), _p(j), _stream())
    return d, j

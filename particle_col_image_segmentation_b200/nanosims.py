"""Device-backed per-ROI reductions of the NanoSIMS MATLAB script
(``HCN_nanosims_rois_activity_distance_5iso_YG.m``).

* ROI labelling in MATLAB's column-major order        .m:104, :173
* ``sum(sum(plane .* roimask))`` per plane and ROI      .m:126-132, :190-196
* isotope activities                                    .m:136-139, :200-203
* ROI centroids ``(x, y)``, 1-based                     .m:164-165, :228-229
* nearest neighbour between the two ROI sets            .m:260-267
* distance to the aggregate boundary pixels             .m:290-308
* activity-vs-distance binning (north_star; the script writes one row per ROI)

The decisions listed in ``oracle/nanosims.py`` (ROI order, identity ``imresize``,
boundary-pixel definition, the reference's (row, col) vs (x, y) mix at .m:301) apply
here unchanged.
"""

import numpy as np
import torch

from . import _io, _lib, ops

PLANES_7 = ("12C", "13C", "14N12C", "15N12C", "16O", "17O", "18O")
ACTIVITIES_7 = (("13C", (1, (1, 0))), ("15N", (3, (2, 3))), ("17O", (5, (6, 5, 4))), ("18O", (6, (6, 5, 4))))
ACTIVITIES_5 = (("13C", (1, (1, 0))), ("15N", (3, (2, 3))))


def matlab_label(mask):
    """8-connected components numbered in column-major order: raster labelling of the
    transposed mask.  Returns ``(device int32 labels (H, W), n)``."""
    t = _io.image_2d(mask)
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    tt = ops.transpose2d(t[0]).unsqueeze(0)
    bits = ops.compare(tt, "!=", 0)[0]
    lab, counts, _ = ops.label_bits(bits, int(tt.shape[2]), connectivity=8, dtype=torch.int32)
    return ops.transpose2d(lab[0]), int(counts[0].item())


def _cubic(x):
    ax = np.abs(x)
    ax2, ax3 = ax * ax, ax * ax * ax
    return (1.5 * ax3 - 2.5 * ax2 + 1) * (ax <= 1) + (-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2) * ((1 < ax) & (ax <= 2))


def resize_contributions(in_length, out_length, antialiasing=True):
    """Tap table of one dimension of MATLAB's ``imresize`` (bicubic, a = -0.5; the kernel is stretched by 1 / scale
    when shrinking with antialiasing): ``(idx int32 (out, P), weights float64 (out, P))``, source indices 0-based and
    mirrored at the ends, weights normalised per output index (imresize.m ``contributions``).  Host work: out x P
    numbers."""
    scale = out_length / in_length
    kernel_width = 4.0
    if scale < 1 and antialiasing:
        h = lambda x: scale * _cubic(scale * x)  # noqa: E731
        kernel_width = kernel_width / scale
    else:
        h = _cubic
    x = np.arange(1, out_length + 1, dtype=np.float64)[:, None]
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kernel_width / 2)
    P = int(np.ceil(kernel_width)) + 2
    indices = left + np.arange(P)[None, :]
    weights = h(u - indices)
    weights = weights / weights.sum(axis=1, keepdims=True)
    aux = np.concatenate([np.arange(1, in_length + 1), np.arange(in_length, 0, -1)])
    indices = aux[np.mod(indices.astype(np.int64) - 1, aux.size)]
    keep = np.any(weights != 0, axis=0)
    return (indices[:, keep] - 1).astype(np.int32), np.ascontiguousarray(weights[:, keep])


def _transpose_taps(idx, w, in_length):
    """The adjoint tap table: for every SOURCE index the outputs it feeds and with which weight, padded with -1."""
    out_i, p_i = np.nonzero(w != 0)
    src = idx[out_i, p_i]
    order = np.lexsort((out_i, src))  # by source index, outputs ascending within it
    src, out_i, wv = src[order], out_i[order], w[out_i, p_i][order]
    counts = np.bincount(src, minlength=in_length)
    Q = max(1, int(counts.max()))
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    col = np.arange(src.size) - start[src]
    tidx = np.full((in_length, Q), -1, dtype=np.int32)
    tw = np.zeros((in_length, Q), dtype=np.float64)
    tidx[src, col] = out_i
    tw[src, col] = wv
    return tidx, tw


def _apply_taps(t, idx, w, dim):
    """One dimension of a resize on the device: ``t`` (H, W) float64 -> the same with ``dim`` replaced by len(idx)."""
    H, W = (int(v) for v in t.shape)
    n_main = int(idx.shape[0])
    out = torch.empty((n_main, W) if dim == 0 else (H, n_main), dtype=torch.float64, device=t.device)
    d_idx = torch.from_numpy(np.ascontiguousarray(idx)).to(t.device)
    d_w = torch.from_numpy(np.ascontiguousarray(w)).to(t.device)
    if dim == 0:
        args = (n_main, W, W, 1, W, 1)
    else:
        args = (n_main, H, 1, W, 1, n_main)
    _lib.call("pcs_resize_taps_f64", ops._p(t), ops._p(out), ops._p(d_idx), ops._p(d_w), int(idx.shape[1]), *args, ops._stream())
    return out


def imresize(a, out_shape, antialiasing=True):
    """MATLAB ``imresize(A, [rows cols])`` for a double image with the defaults the script relies on (.m:125, :189):
    bicubic, antialiased when shrinking, one dimension at a time, the smaller scale first.  numpy in -> numpy out,
    CUDA tensor in -> CUDA tensor out."""
    t = _f64_image(a)
    scales = [out_shape[0] / t.shape[0], out_shape[1] / t.shape[1]]
    for dim in sorted((0, 1), key=lambda d: scales[d]):
        idx, w = resize_contributions(int(t.shape[dim]), int(out_shape[dim]), antialiasing)
        t = _apply_taps(t, idx, w, dim)
    return _io.back(t, _io.is_numpy(a))


def _adjoint_resize(plane, roi_shape):
    """R^T applied to an acquisition-size plane, R = imresize from ``roi_shape`` to ``plane.shape``:
    ``sum(plane .* imresize(holder))  ==  sum((R^T plane) .* holder)`` for every ROI's ``holder``, so the per-ROI sums
    under resized masks (.m:125-132) cost one adjoint resize per ion plane plus the ordinary masked sums, not one
    resize per ROI.  The dimensions are undone in the reverse of imresize's order."""
    t = plane
    scales = [plane.shape[0] / roi_shape[0], plane.shape[1] / roi_shape[1]]
    for dim in reversed(sorted((0, 1), key=lambda d: scales[d])):
        idx, w = resize_contributions(int(roi_shape[dim]), int(plane.shape[dim]))
        tidx, tw = _transpose_taps(idx, w, int(roi_shape[dim]))
        t = _apply_taps(t, tidx, tw, dim)
    return t


def _set_table(planes_d, mask, spec, set_id):
    lab, n = matlab_label(mask)
    if tuple(lab.shape) != tuple(planes_d.shape[1:]):  # ROI image and acquisition differ in size: .m:125 resizes every ROI mask
        planes_d = torch.stack([_adjoint_resize(planes_d[k], tuple(lab.shape)) for k in range(planes_d.shape[0])])
    sums = ops.roi_sums(lab, planes_d, n).cpu().numpy()
    tab = ops.new_table(max(1, n), lab.device)
    ops.region_table(lab.unsqueeze(0), None, tab)
    t = tab.cpu().numpy()
    area = t[ops.T_AREA, :n].astype(np.float64)
    xy = np.column_stack([t[ops.T_SUMX, :n] / area + 1.0, t[ops.T_SUMY, :n] / area + 1.0]) if n else np.zeros((0, 2))
    cols = []
    for _, (num, den) in spec:
        d = np.zeros(n)
        for j in den:
            d = d + sums[:, j]
        with np.errstate(divide="ignore", invalid="ignore"):
            cols.append(sums[:, num] / d)
    act = np.column_stack(cols) if cols else np.zeros((n, 0))
    rows = np.column_stack([np.full(n, float(set_id)), np.arange(1, n + 1, dtype=np.float64), sums, act, act * 100.0])
    return rows, xy


def boundary_pixels(mask):
    """Mask pixels with a 4-neighbour outside the mask, ``(row, col)`` 1-based, raster order
    (.m:290-291): the mask minus its erosion by the 4-neighbourhood cross (outside = False)."""
    bits, H, W = _io.mask_bits(mask)
    cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)
    inner = ops.erode(bits, W, cross, border_value=0)
    edge = ops.logic(bits, inner, "andnot", W)
    m = ops.unpack(edge, W, torch.uint8)[0].cpu().numpy()  # a few hundred pixels of a 256^2 mask: listed on the host
    return torch.from_numpy(np.argwhere(m).astype(np.float64) + 1.0).to(bits.device)


def analyse(planes, red_mask, green_mask, agg_mask, raster=19.0, acq=512.0):
    """Rows ``[set, i, sums..., act..., act*100..., x, y, nearest_um, boundary_um]`` for the
    red then the green ROIs (.m:154, :216, :249-252, :265-268, :306-309)."""
    planes_d = _io.to_device(planes, torch.float64)
    k = int(planes_d.shape[0])
    spec = ACTIVITIES_7 if k >= 7 else ACTIVITIES_5
    ra, axy = _set_table(planes_d, red_mask, spec, 1)
    rb, bxy = _set_table(planes_d, green_mask, spec, 2)
    dev = planes_d.device
    a_d, b_d = torch.from_numpy(axy).to(dev), torch.from_numpy(bxy).to(dev)
    bd = boundary_pixels(agg_mask)
    scale = raster / acq
    near = np.concatenate([ops.min_dist(a_d, b_d).cpu().numpy(), ops.min_dist(b_d, a_d).cpu().numpy()]) * scale
    bdist = np.concatenate([ops.min_dist(a_d, bd).cpu().numpy(), ops.min_dist(b_d, bd).cpu().numpy()]) * scale
    return np.column_stack([np.concatenate([ra, rb]), np.concatenate([axy, bxy]), near, bdist])


def activity_vs_distance(activity, distance, edges):
    """``np.digitize`` + ``np.bincount`` over the per-ROI rows (host: a few hundred ROIs)."""
    idx = np.digitize(distance, edges)
    nb = len(edges) + 1
    cnt = np.bincount(idx, minlength=nb).astype(np.float64)
    tot = np.bincount(idx, weights=activity, minlength=nb)
    with np.errstate(divide="ignore", invalid="ignore"):
        return cnt, tot, tot / cnt


# ---------------------------------------------------------------- ratio images (.m:17-69)
def _f64_image(a):
    t = _io.to_device(a, torch.float64)
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D plane, got shape {tuple(t.shape)}")
    return t


def imgaussfilt(a, sigma):
    """MATLAB ``imgaussfilt(A, sigma)`` with its defaults (.m:43): kernel ``2*ceil(2*sigma)+1``, replicate
    border, double precision.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    t = _f64_image(a)
    out, tmp = torch.empty_like(t), torch.empty_like(t)
    H, W = t.shape
    _lib.call("pcs_gauss_f64", ops._p(t), ops._p(out), ops._p(tmp), float(sigma), 1, int(H), int(W), ops._stream())
    return _io.back(out, _io.is_numpy(a))


def scaled_uint8(num, dens=()):
    """``uint8(R .* (255 / max(R(:))))`` with ``R = num ./ (dens[0] + dens[1] + ...)`` (``R = num`` without
    denominators): the conversion behind every ``...img`` variable of the script (.m:31-37, :45-69)."""
    n = _f64_image(num)
    ds = [_f64_image(d) for d in dens]
    if len(ds) > 3:
        raise ValueError("at most three denominators")
    ratio = torch.empty_like(n)
    maxv = torch.empty(1, dtype=torch.float64, device=n.device)
    p = [ops._p(d) for d in ds] + [0] * (3 - len(ds))
    _lib.call("pcs_ratio_f64", ops._p(n), p[0], p[1], p[2], ops._p(ratio), ops._p(maxv), int(n.numel()), ops._stream())
    out = torch.empty(n.shape, dtype=torch.uint8, device=n.device)
    _lib.call("pcs_scale_u8_f64", ops._p(ratio), ops._p(maxv), ops._p(out), int(n.numel()), ops._stream())
    return _io.back(out, _io.is_numpy(num))


def ratio_images(ions):
    """The image block at the top of the script (.m:17-69).  ``ions`` maps the seven plane names of
    ``PLANES_7`` (and optionally ``"Esi"``) to the raw square ``IM`` arrays; returns the uint8 images under the
    script's own variable names.  The one-pixel frame is cropped first (``IM(2:n-1, 2:n-1)``, .m:18-28)."""
    raw = {k: _f64_image(np.asarray(v, dtype=np.float64)[1:-1, 1:-1]) for k, v in ions.items()}
    g1 = {k: imgaussfilt(raw[k], 1) for k in ("15N12C", "14N12C", "16O", "17O", "18O") if k in raw}
    g15 = {k: imgaussfilt(raw[k], 1.5) for k in ("12C", "13C", "Esi") if k in raw}
    out = {}
    for k, name in (("12C", "C12img"), ("13C", "C13img"), ("14N12C", "N14C12img"), ("15N12C", "N15C12img"), ("16O", "O16img"), ("17O", "O17img"), ("18O", "O18img")):
        if k in raw:
            out[name] = scaled_uint8(raw[k])
    if "15N12C" in raw and "14N12C" in raw:
        out["N15ratioimg"] = scaled_uint8(g1["15N12C"], (g1["15N12C"], g1["14N12C"]))  # .m:45
        out["N15ratimg"] = scaled_uint8(raw["15N12C"], (raw["15N12C"], raw["14N12C"]))  # .m:65
    if "12C" in raw and "13C" in raw:
        out["C13ratioimg"] = scaled_uint8(g15["13C"], (g15["13C"], g15["12C"]))  # .m:54
        out["C13ratimg"] = scaled_uint8(raw["13C"], (raw["13C"], raw["12C"]))  # .m:66
        if "14N12C" in raw:
            out["N14C12C12ratio"] = scaled_uint8(g1["14N12C"], (g15["12C"],))  # .m:53
    if all(k in raw for k in ("16O", "17O", "18O")):
        dens_g = (g1["18O"], g1["17O"], g1["16O"])
        dens_r = (raw["18O"], raw["17O"], raw["16O"])
        out["O17ratioimg"] = scaled_uint8(g1["17O"], dens_g)  # .m:59
        out["O18ratioimg"] = scaled_uint8(g1["18O"], dens_g)  # .m:60
        out["O17ratimg"] = scaled_uint8(raw["17O"], dens_r)  # .m:67
        out["O18ratimg"] = scaled_uint8(raw["18O"], dens_r)  # .m:68
    if "Esi" in raw and "14N12C" in raw:
        out["N14C12ESIratio"] = scaled_uint8(raw["14N12C"], (raw["Esi"],))  # .m:64 overwrites .m:63
    return {k: v.cpu().numpy() for k, v in out.items()}

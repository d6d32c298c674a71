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

from . import _io, ops

PLANES_7 = ("12C", "13C", "14N12C", "15N12C", "16O", "17O", "18O")
ACTIVITIES_7 = (("13C", (1, (1, 0))), ("15N", (3, (2, 3))), ("17O", (5, (6, 5, 4))), ("18O", (6, (6, 5, 4))))
ACTIVITIES_5 = (("13C", (1, (1, 0))), ("15N", (3, (2, 3))))


def matlab_label(mask):
    """8-connected components numbered in column-major order: raster labelling of the
    transposed mask.  Returns ``(device int32 labels (H, W), n)``."""
    t = _io.image_2d(mask)
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    tt = t[0].t().contiguous().unsqueeze(0)
    bits = ops.compare(tt, "!=", 0)[0]
    lab, counts, _ = ops.label_bits(bits, int(tt.shape[2]), connectivity=8, dtype=torch.int32)
    return lab[0].t().contiguous(), int(counts[0].item())


def _set_table(planes_d, mask, spec, set_id):
    lab, n = matlab_label(mask)
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
    m = ops.unpack(edge, W, torch.bool)[0]
    return (torch.nonzero(m).to(torch.float64) + 1.0).contiguous()  # nonzero() hands back a transposed view


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

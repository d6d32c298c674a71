"""CPU restatement of the reference's image-analysis (L2) functions.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Each function cites the
``tiff_analysis.py`` lines it follows.  The restatement is table-oriented (one
pass over a per-label table instead of the reference's per-region Python loops)
but returns the same structures, so it can be compared entry by entry with the
reference module loaded through ``oracle.ref_loader`` (done in
``tests/test_oracle.py::test_l2_matches_live_reference`` and frozen in ``tests/golden``).
"""

import numpy as np
from scipy import ndimage as ndi

from .skimage_shim.measure import label, regionprops
from .skimage_shim.morphology import binary_dilation, disk

# tiff_analysis.py:47-82 (module constants; values become kernel parameters)
CELL_TYPES = ["3D05", "6B07", "C3M10"]
BASE_TYPE_MAP = {1: "3D05", 2: "6B07", 3: "C3M10", 4: "Particle", 5: "Background"}
STRAIN_MAP = {"3D05": "RFP", "6B07": "DAPI", "C3M10": "GFP"}
MIN_CELL_AREA = {"3D05": 20, "6B07": 20, "C3M10": 20}
MIN_CLUSTER_AREA = {"3D05": 200, "6B07": 200, "C3M10": 370}
DENOISE_SIZE = 5
DILATION_RADIUS = 20
DISTANCE_THRESHOLD = 2
CELL_CLUSTER_DISTANCE_THRESHOLD = 5
DAPI_RFP_OVERLAP_THRESHOLD = 0.1
PX_TO_UM_CONV = 9.95


def normalize_ds_arr(ds_arr, side=2048):
    """tiff_analysis.py:727-737 -- squeeze ``(S,S,1)`` / ``(1,S,S)`` / ``(S,S)`` to
    2-D.  The reference hard-codes ``S == 2048``; ``side=None`` lifts that gate."""
    if ds_arr.shape[-1] == 1:
        return np.squeeze(ds_arr)
    if ds_arr.shape[0] == 1:
        return ds_arr[0]
    if ds_arr.ndim == 2 and (side is None or (ds_arr.shape[0] == side and ds_arr.shape[1] == side)):
        return ds_arr
    raise ValueError(f"DS arr shape is not ({side},{side},1) or (1,{side},{side}) or ({side},{side}). Shape: {ds_arr.shape}")


def denoise(ds_arr):
    """tiff_analysis.py:122, :643 -- ``median_filter(ds_arr, size=5)`` (mode reflect)."""
    return ndi.median_filter(ds_arr, size=DENOISE_SIZE)


def get_cell_positions_and_areas(z_slice, cell_types, merged=False):
    """tiff_analysis.py:742-789.

    Multi-valued 8-connected labelling (:743), one region per label (:746), class
    of a region = class value at its first raster pixel (:755, :1041-1044),
    particle area = sum of particle-region areas (:760), cells have
    ``MIN_CELL_AREA <= area < MIN_CLUSTER_AREA`` (:769), clusters
    ``area >= MIN_CLUSTER_AREA`` (:772), ``cluster.cells = int(area // mean cell
    area)`` (:776-781).
    """
    label_im = label(z_slice)
    regions = regionprops(label_im)
    cell_pos, cell_clusters = {}, {}
    particle_area = 0
    for region in regions:
        r0, c0 = region.coords[0]
        cell_type = cell_types[z_slice[r0, c0]]
        if cell_type not in CELL_TYPES:
            if cell_type == "Particle":
                particle_area += region.area
            continue
        cell_pos.setdefault(cell_type, [])
        cell_clusters.setdefault(cell_type, [])
        a = region.area
        if MIN_CELL_AREA[cell_type] <= a < MIN_CLUSTER_AREA[cell_type]:
            cell_pos[cell_type].append(region)
        if a >= MIN_CLUSTER_AREA[cell_type]:
            cell_clusters[cell_type].append(region)
    for cell_type, clusters in cell_clusters.items():
        mean_area = np.average([c.area for c in cell_pos[cell_type]])
        for cluster in clusters:
            cluster.cells = int(cluster.area // mean_area)
    merged_clusters = {}
    if merged:
        merged_clusters, _ = get_cell_clusters_from_distances(z_slice, cell_pos, cell_clusters, cell_types)
    return cell_pos, cell_clusters, particle_area, merged_clusters


def get_cell_clusters_from_distances(z_slice, cell_pos, cell_clusters, cell_types):
    """tiff_analysis.py:791-824 -- per cell class, and for the union of all classes,
    merge regions whose dilated masks touch."""
    merged_regions, merged_images = {}, {}
    img_vals, everything = [], []
    for key in set(cell_pos) | set(cell_clusters):
        regs = cell_pos.get(key, []) + cell_clusters.get(key, [])
        val = next((v for v, name in cell_types.items() if name == key), 0)
        img_vals.append(val)
        everything.extend(regs)
        merged_regions[key], merged_images[key] = get_merged_regions(z_slice == val, regs)
    union = np.isin(z_slice, img_vals) if img_vals else np.zeros_like(z_slice, dtype=bool)
    merged_regions["combined"], merged_images["combined"] = get_merged_regions(union, everything)
    return merged_regions, merged_images


def get_merged_regions(binary_image, og_cell_regions):
    """tiff_analysis.py:826-883.

    Dilate by ``disk(5 // 2)`` (:827-828), label (:829), key every region by the
    dilated label under ``(int(cy), int(cx))`` (:843-852), and emit one merged
    record per key in order of first appearance: summed area (:855), area-weighted
    mean centroid via ``np.average`` (:856-858), union bbox (:860-864).  The merged
    image is the OR of the selected dilated components (:878), hole-filled (:880).
    """
    dilated = binary_dilation(binary_image, disk(CELL_CLUSTER_DISTANCE_THRESHOLD // 2))
    dilated_labels = label(dilated)
    keys = []
    for r in og_cell_regions:
        cy, cx = r.centroid
        keys.append(int(dilated_labels[int(cy), int(cx)]))
    merged, seen = [], []
    for k in keys:
        if k <= 0 or k in seen:
            continue
        seen.append(k)
        members = [r for r, kk in zip(og_cell_regions, keys) if kk == k]
        areas = [m.area for m in members]
        bbs = np.array([m.bbox for m in members])
        merged.append(
            {
                "area": sum(areas),
                "centroid": np.average([m.centroid for m in members], axis=0, weights=areas),
                "regions": members,
                "bbox": (int(bbs[:, 0].min()), int(bbs[:, 1].min()), int(bbs[:, 2].max()), int(bbs[:, 3].max())),
            }
        )
    merged_image = np.isin(dilated_labels, seen) if seen else np.zeros_like(binary_image, dtype=bool)
    merged_image = ndi.binary_fill_holes(merged_image)
    return merged, merged_image


def fill_particle_area(ds_arr, particle_label, cell_label, overlap_label):
    """tiff_analysis.py:982-1015 -- cell pixels inside the ``disk(20)``-dilated particle
    (:990, :1004) or closer than ``DISTANCE_THRESHOLD`` to it (:996-1000) become
    ``overlap_label``; returns the new image and the pixel count (:1015)."""
    particle = ds_arr == particle_label
    cell = ds_arr == cell_label
    dilated = binary_dilation(particle, disk(DILATION_RADIUS))
    dist = ndi.distance_transform_edt(~particle)
    overlap = (cell & (dist < DISTANCE_THRESHOLD)) | (cell & dilated)
    updated = ds_arr.copy()
    updated[overlap] = overlap_label
    return updated, np.sum(overlap)


def recreate_particle_area(ds_arr, cell_types, particle_area):
    """tiff_analysis.py:931-950 -- chain ``fill_particle_area`` over the cell classes."""
    particle_label = None
    for key, value in cell_types.items():
        if value == "Particle":
            particle_label = key
    for cell_label, cell_type in cell_types.items():
        if cell_type not in CELL_TYPES:
            continue
        ds_arr, n = fill_particle_area(ds_arr, particle_label, cell_label, overlap_label=particle_label)
        particle_area += n
    return ds_arr, particle_area


def combine_cell_positions_and_clusters(dapi_channel, other_channel):
    """tiff_analysis.py:252-287 -- DAPI cells (class 1) whose overlap with the other
    channel's class-1 mask exceeds 10 % of their area are rewritten to class 2.
    The per-cell full-image passes (:268-279) are a bincount here."""
    dapi_mask = dapi_channel == 1
    other_mask = other_channel == 1
    lab = label(dapi_mask)
    n = int(lab.max())
    area = np.bincount(lab.ravel(), minlength=n + 1).astype(np.float64)
    ov = np.bincount(lab[other_mask].ravel(), minlength=n + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        frac = ov / area
    remove = frac > DAPI_RFP_OVERLAP_THRESHOLD
    remove[0] = False
    out = dapi_channel.copy()
    out[remove[lab]] = 2
    return out


def get_rfp_base_arr(rfp_arr, cell_strains):
    """tiff_analysis.py:224-231 -- ordered in-place remap to the base class numbering."""
    if cell_strains == ["6B07"] or cell_strains == ["6B07", "C3M10"]:
        steps = [(1, 4), (2, 5)]
    else:
        steps = [(2, 4), (3, 5)]
    for a, b in steps:
        rfp_arr[rfp_arr == a] = b
    return rfp_arr


def relabel_other_channel(other_channel, other_channel_name):
    """tiff_analysis.py:177-181 -- ``3->5`` then ``2->4`` (then ``1->3`` for GFP) on a copy."""
    out = other_channel.copy()
    out[out == 3] = 5
    out[out == 2] = 4
    if other_channel_name == "GFP":
        out[out == 1] = 3
    return out


def combine_channels(rfp_base, channel_ds_arrs, cell_strains):
    """tiff_analysis.py:233-249 -- overwrite the base image with each non-3D05 strain's
    base class number where that strain's channel holds class 1."""
    for strain in cell_strains:
        if strain == "3D05":
            continue
        val = next(v for v, name in BASE_TYPE_MAP.items() if name == strain)
        rfp_base[channel_ds_arrs[STRAIN_MAP[strain]] == 1] = val
    return rfp_base


def get_cell_counts_and_densities(cell_pos, cell_clusters, particle_area):
    """tiff_analysis.py:1018-1038."""
    cell_count, cell_density, cell_area_ratio = {}, {}, {}
    particle_area = particle_area / (PX_TO_UM_CONV**2)
    for cell_type, cells in cell_pos.items():
        if cell_type not in CELL_TYPES:
            continue
        clusters = cell_clusters[cell_type]
        cell_count[cell_type] = len(cells) + sum(c.cells for c in clusters)
        cell_area = np.sum([c.area for c in cells])
        for c in clusters:
            cell_area += c["area"]
        area = cell_area / (PX_TO_UM_CONV**2)
        cell_density[cell_type] = round(cell_count[cell_type] / particle_area, 5)
        cell_area_ratio[cell_type] = round(area / particle_area, 5)
    return cell_count, cell_density, cell_area_ratio


# --------------------------------------------------------------------------
# comparison helpers (plain data out of region lists)
# --------------------------------------------------------------------------
def region_row(r):
    return (int(r.label), float(r.area), tuple(float(c) for c in r.centroid), tuple(int(b) for b in r.bbox), int(getattr(r, "cells", -1)))


def summarize_positions(result):
    """Flatten ``get_cell_positions_and_areas`` output into comparable plain data."""
    cell_pos, cell_clusters, particle_area, merged = result
    out = {
        "cell_pos": {k: [region_row(r) for r in v] for k, v in cell_pos.items()},
        "cell_clusters": {k: [region_row(r) for r in v] for k, v in cell_clusters.items()},
        "particle_area": float(particle_area),
        "merged": {},
    }
    for k, recs in merged.items():
        out["merged"][k] = [
            (float(m["area"]), tuple(float(c) for c in m["centroid"]), tuple(int(b) for b in m["bbox"]), [int(r.label) for r in m["regions"]])
            for m in recs
        ]
    return out


def get_cell_cell_distances(cell_pos):
    """Brute-force statement of the nearest-neighbour distances (refine_boundaries.py:8-12 goal 3; model
    .m:260-263 ``pdist2`` + ``min``): ``{(a, b): (distance, index)}`` per cell of strain ``a``; within a strain
    the cell itself is excluded.  Same operation order as the device kernel: (dx*dx + dy*dy), sqrt last."""
    pts = {k: np.array([r.centroid for r in v], dtype=np.float64).reshape(-1, 2) for k, v in cell_pos.items()}
    out = {}
    for a, pa in pts.items():
        for b, pb in pts.items():
            if len(pa) == 0:
                out[(a, b)] = (np.zeros(0), np.zeros(0, dtype=np.int64))
                continue
            if len(pb) == 0:
                out[(a, b)] = (np.full(len(pa), np.inf), np.full(len(pa), -1, dtype=np.int64))
                continue
            dx = pa[:, None, 0] - pb[None, :, 0]
            dy = pa[:, None, 1] - pb[None, :, 1]
            d2 = dx * dx + dy * dy
            if a == b:
                np.fill_diagonal(d2, np.inf)
            j = np.argmin(d2, axis=1)
            d = np.sqrt(d2[np.arange(len(pa)), j])
            j = np.where(np.isinf(d), -1, j)
            out[(a, b)] = (d, j.astype(np.int64))
    return out

"""Device-backed mirror of the reference's image-analysis functions.

Same names, arguments, return structures and error behaviour as the L2 functions of
``/root/reference/tiff_analysis.py`` (function level), so that code
written against the reference runs unchanged:

    get_cell_positions_and_areas     tiff_analysis.py:742-789
    get_cell_clusters_from_distances tiff_analysis.py:791-824
    get_merged_regions               tiff_analysis.py:826-883
    recreate_particle_area           tiff_analysis.py:931-950
    fill_particle_area               tiff_analysis.py:982-1015
    combine_cell_positions_and_clusters  tiff_analysis.py:252-287
    get_rfp_base_arr / combine_channels  tiff_analysis.py:224-249
    get_cell_counts_and_densities    tiff_analysis.py:1018-1038
    normalize_ds_arr / get_type      tiff_analysis.py:727-737, :1041-1044

The pixel work (labelling, reductions, dilation, EDT, hole filling, mask algebra)
runs in libpcs kernels; what stays on the host is the reference's own bookkeeping
over the per-label table (a few thousand rows): class lookup, area filters,
group-by of regions under a dilated label, rounding.  The reference's per-region
full-image passes (:268-279, :843-878) become one table pass each.
"""

import numpy as np
import torch

from . import _io, ops
from .measure import LabelHolder, RegionProperties
from .morphology import disk
from .ndimage import median_filter  # noqa: F401  (tiff_analysis.py:42)

# tiff_analysis.py:47-82 -- module constants, user-editable in the reference
BASE_TYPE_MAP = {1: "3D05", 2: "6B07", 3: "C3M10", 4: "Particle", 5: "Background"}
CELL_TYPES = ["3D05", "6B07", "C3M10"]
CHANNELS = ["RFP", "DAPI", "GFP"]
CHANNEL_MAP = {"RFP": "3D05", "DAPI": "6B07", "GFP": "C3M10"}
STRAIN_MAP = {"3D05": "RFP", "6B07": "DAPI", "C3M10": "GFP"}
MIN_CELL_AREA = {"3D05": 20, "6B07": 20, "C3M10": 20}
MIN_CLUSTER_AREA = {"3D05": 200, "6B07": 200, "C3M10": 370}
DENOISE_SIZE = 5
DILATION_RADIUS = 20
DISTANCE_THRESHOLD = 2
CELL_CLUSTER_DISTANCE_THRESHOLD = 5
DAPI_RFP_OVERLAP_THRESHOLD = 0.1
PX_TO_UM_CONV = 9.95


def normalize_ds_arr(ds_arr):
    """tiff_analysis.py:727-737.  The reference accepts only 2048-pixel sides; any
    ``(H, W, 1)``, ``(1, H, W)`` or ``(H, W)`` array is accepted here."""
    if ds_arr.shape[-1] == 1:
        return np.squeeze(ds_arr)
    if ds_arr.shape[0] == 1:
        return ds_arr[0]
    if ds_arr.ndim == 2:
        return ds_arr
    raise ValueError(f"DS arr shape is not (H,W,1) or (1,H,W) or (H,W). Shape: {ds_arr.shape}")


def get_type(region, data):
    """tiff_analysis.py:1041-1044 -- class value at the region's first raster pixel."""
    r, c = region.first_pixel if hasattr(region, "first_pixel") else region.coords[0]
    return data[r, c]


# ---------------------------------------------------------------- device helpers
def _u8_image(a):
    t = _io.image_2d(a)
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    if t.dtype != torch.uint8:
        raise TypeError(f"class images must be uint8, got {t.dtype}")
    return t


def _label_and_table(t, bits=None):
    """Label a (1, H, W) uint8 class image (multi-valued) or a bit image; return
    ``(labels, n, host table, class value per label)``."""
    if bits is None:
        labels, counts, _ = ops.label_values(t, connectivity=8, dtype=torch.int64)
    else:
        labels, counts, _ = ops.label_bits(bits, int(t.shape[2]), connectivity=8, dtype=torch.int32)
    n = int(counts[0].item())
    table = ops.new_table(max(1, n), t.device)
    ops.region_table(labels, None, table, fg_bits=bits)
    cls = ops.gather(t, None, table[ops.T_FIRST, :n].contiguous()) if n else torch.zeros(0, dtype=torch.int64, device=t.device)
    return labels, n, table.cpu().numpy(), cls.cpu().numpy()


# ---------------------------------------------------------------- L2 functions
def get_cell_positions_and_areas(z_slice, cell_types, merged=False):
    """tiff_analysis.py:742-789: ``(cell_pos, cell_clusters, particle_area, merged_clusters)``.

    The reference walks every connected component in Python -- label noise included, tens of thousands of one-pixel
    specks on an ilastik image -- and asks each for its class, area and size bracket.  Here those are array passes
    over the per-label table (one D2H copy); region objects are created only for the components that survive the
    area filters.  Dict keys keep the reference's insertion order (a cell type enters both dicts when its first
    component of ANY size is met)."""
    t = _u8_image(z_slice)
    shape = (int(t.shape[1]), int(t.shape[2]))
    labels, n, tab, cls = _label_and_table(t)
    holder = LabelHolder(labels[0])
    area = tab[ops.T_AREA, :n]
    # class value -> cell type name, once per distinct value (a value missing from cell_types raises KeyError, as :756 does)
    values, first_at, inv = np.unique(cls[:n], return_index=True, return_inverse=True)
    names = [cell_types[int(v)] for v in values]
    cell_pos, cell_clusters = {}, {}
    particle_area = 0
    is_particle = np.array([nm == "Particle" for nm in names], dtype=bool)
    if n and is_particle[inv].any():
        particle_area = np.float64(area[is_particle[inv]].sum())  # integer areas: any summation order is exact
    for vi in np.argsort(first_at, kind="stable").tolist():  # cell types in order of their first component
        cell_type = names[vi]
        if cell_type not in CELL_TYPES:
            continue
        of_type = inv == vi
        if cell_type not in cell_pos:
            cell_pos[cell_type] = []
            cell_clusters[cell_type] = []
        lo, hi = MIN_CELL_AREA[cell_type], MIN_CLUSTER_AREA[cell_type]
        cells = np.flatnonzero(of_type & (area >= lo) & (area < hi))
        clusters = np.flatnonzero(of_type & (area >= hi))
        # two class values may map to one cell type: keep label order within the type
        cell_pos[cell_type] = _merge_by_label(cell_pos[cell_type], [RegionProperties(i + 1, tab[:, i], holder, None, shape) for i in cells.tolist()])
        cell_clusters[cell_type] = _merge_by_label(cell_clusters[cell_type], [RegionProperties(i + 1, tab[:, i], holder, None, shape) for i in clusters.tolist()])
    cell_area_averages = {}
    for cell_type, cell_array in cell_pos.items():
        cell_area_averages[cell_type] = np.average([cell.area for cell in cell_array])
    for cell_type, cluster_array in cell_clusters.items():
        for cluster in cluster_array:
            cluster.cells = int(cluster.area // cell_area_averages[cell_type])
    if merged:
        merged_clusters, _ = _clusters_from_distances(t, cell_pos, cell_clusters, cell_types)
    else:
        merged_clusters = {}
    return cell_pos, cell_clusters, particle_area, merged_clusters


def _merge_by_label(a, b):
    if not a:
        return b
    return sorted(a + b, key=lambda r: r.label)


def _clusters_from_distances(t, cell_pos, cell_clusters, cell_types, want_numpy=True):
    combined = {}
    for key in set(cell_pos) | set(cell_clusters):
        combined[key] = cell_pos.get(key, []) + cell_clusters.get(key, [])
    merged_regions, merged_images = {}, {}
    img_vals, combined_regions = [], []
    for cell_type, cell_regions in combined.items():
        cell_img_val = 0
        for cell_val, cell_temp_type in cell_types.items():
            if cell_temp_type == cell_type:
                cell_img_val = cell_val
                break
        img_vals.append(cell_img_val)
        combined_regions.extend(cell_regions)
        bits = ops.compare(t, "==", cell_img_val)[0]
        merged_regions[cell_type], merged_images[cell_type] = _merged_regions(bits, t, cell_regions, want_numpy)
    bits = ops.member_u8(t, img_vals)[0]
    merged_regions["combined"], merged_images["combined"] = _merged_regions(bits, t, combined_regions, want_numpy)
    return merged_regions, merged_images


def get_cell_clusters_from_distances(z_slice, cell_pos, cell_clusters, cell_types):
    """tiff_analysis.py:791-824."""
    return _clusters_from_distances(_u8_image(z_slice), cell_pos, cell_clusters, cell_types, _io.is_numpy(z_slice))


def _merged_regions(bits, t, og_cell_regions, want_numpy=True):
    """tiff_analysis.py:826-883.  The reference groups the regions under each dilated component with a nested Python
    loop (O(R^2)) and ORs one full-image ``dilated_labels == v`` per group; here the group-by is one pass over the
    keys (``np.unique`` + unbuffered ``ufunc.at``, which adds in list order like ``np.average`` over the rows does,
    so the area-weighted centroids come out bit-identical) and the selected components are one label-LUT kernel."""
    H, W = int(t.shape[1]), int(t.shape[2])
    dilated = ops.dilate(bits, W, disk(CELL_CLUSTER_DISTANCE_THRESHOLD // 2))
    dlabels, dcounts, _ = ops.label_bits(dilated, W, connectivity=8, dtype=torch.int32)
    merged_regions = []
    processed = []
    if og_cell_regions:
        R = len(og_cell_regions)
        rows = np.stack([r._row for r in og_cell_regions]) if all(hasattr(r, "_row") for r in og_cell_regions) else None
        if rows is not None:
            area = rows[:, ops.T_AREA].astype(np.float64)
            cent = np.column_stack([rows[:, ops.T_SUMY] / area, rows[:, ops.T_SUMX] / area])
            bbox = np.column_stack([rows[:, ops.T_MINY], rows[:, ops.T_MINX], rows[:, ops.T_MAXY] + 1, rows[:, ops.T_MAXX] + 1])
        else:  # foreign region objects (anything with .area / .centroid / .bbox)
            area = np.array([r.area for r in og_cell_regions], dtype=np.float64)
            cent = np.array([r.centroid for r in og_cell_regions], dtype=np.float64).reshape(R, 2)
            bbox = np.array([r.bbox for r in og_cell_regions], dtype=np.int64).reshape(R, 4)
        iy, ix = cent[:, 0].astype(np.int64), cent[:, 1].astype(np.int64)  # int(centroid): truncation (:849)
        inside = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)               # the reference's bounds check (:850)
        keys = np.zeros(R, dtype=np.int64)
        if inside.any():
            lin = (iy[inside] * W + ix[inside]).astype(np.int64)
            keys[inside] = ops.gather(dlabels, None, torch.from_numpy(lin).to(t.device)).cpu().numpy()
        member = np.flatnonzero(keys > 0)
        if member.size:
            uniq, first_at, inv = np.unique(keys[member], return_index=True, return_inverse=True)
            order = np.argsort(first_at, kind="stable")          # groups in order of their first region (:843-878)
            rank = np.empty_like(order)
            rank[order] = np.arange(order.size)
            g = rank[inv]                                        # group index of every member region
            G = order.size
            a_m, c_m, b_m = area[member], cent[member], bbox[member]
            tot = np.zeros(G)
            np.add.at(tot, g, a_m)
            wsum = np.zeros((G, 2))
            np.add.at(wsum, g, c_m * a_m[:, None])               # row order within a group = list order
            lo = np.full((G, 2), np.iinfo(np.int64).max, dtype=np.int64)
            hi = np.full((G, 2), np.iinfo(np.int64).min, dtype=np.int64)
            np.minimum.at(lo, g, b_m[:, :2])
            np.maximum.at(hi, g, b_m[:, 2:])
            members_of = [[] for _ in range(G)]
            for idx, gi in zip(member.tolist(), g.tolist()):
                members_of[gi].append(og_cell_regions[idx])
            centroid = wsum / tot[:, None]
            for gi in range(G):
                merged_regions.append({"area": tot[gi], "centroid": centroid[gi], "regions": members_of[gi],
                                       "bbox": (int(lo[gi, 0]), int(lo[gi, 1]), int(hi[gi, 0]), int(hi[gi, 1]))})
            processed = uniq[order].tolist()
    n = int(dcounts[0].item())
    keep = np.zeros((1, n + 1), dtype=np.uint8)
    keep[0, processed] = 1
    sel = ops.select_labels(dlabels, torch.from_numpy(keep).to(t.device))
    filled = ops.fill_holes(sel, W)
    return merged_regions, _io.bits_to_bool(filled, W, want_numpy)


def get_merged_regions(binary_image, og_cell_regions):
    """tiff_analysis.py:826-883: ``(list of merged-region dicts, hole-filled bool image)``."""
    bits, H, W = _io.mask_bits(binary_image)
    t = torch.empty((1, H, W), dtype=torch.uint8, device=bits.device)  # shape carrier only
    return _merged_regions(bits, t, og_cell_regions, _io.is_numpy(binary_image))


def fill_particle_area(ds_arr, particle_label, cell_label, overlap_label):
    """tiff_analysis.py:982-1015: ``(updated image, number of relabelled pixels)``.

    One exact squared EDT of the complement of the particle mask serves both tests:
    ``dist < DISTANCE_THRESHOLD`` (:1000) and membership of the ``disk(DILATION_RADIUS)``
    dilation (:990, ``EDT^2 <= r^2``)."""
    np_in = _io.is_numpy(ds_arr)
    t = _u8_image(ds_arr)
    W = int(t.shape[2])
    particle = ops.compare(t, "==", int(particle_label))[0]
    cell = ops.compare(t, "==", int(cell_label))[0]
    thr_dist = int(np.ceil(float(DISTANCE_THRESHOLD) ** 2)) - 1  # d < T  <=>  d^2 <= ceil(T^2) - 1
    thr = max(thr_dist, int(DILATION_RADIUS) ** 2)
    has_particle = int(ops.count(particle, W)[0].item()) > 0
    if has_particle:
        near = ops.edt(particle, W, invert=True, want_dist=False, thr_sq=thr)[2]
    else:
        # no particle pixel: the dilation is empty and scipy's EDT measures to the virtual
        # point (-1, 0) -- reproduce that for the distance test only
        sq = ops.edt(particle, W, invert=True, want_dist=False, want_sq=True)[1]
        near = ops.compare(sq, "<=", thr_dist)[0]
    overlap = ops.logic(cell, near, "and", W)
    updated = t.clone()
    ops.assign_where_u8_(updated, overlap, int(overlap_label))
    count = ops.count(overlap, W)
    n = np.int64(count[0].item())
    return _io.back(updated[0], np_in), n


def recreate_particle_area(ds_arr, cell_types, particle_area):
    """tiff_analysis.py:931-950."""
    particle_label = None
    for key, value in cell_types.items():
        if value == "Particle":
            particle_label = key
    for cell_type_label, cell_type in cell_types.items():
        if cell_type not in CELL_TYPES:
            continue
        if particle_label is None:
            # no "Particle" entry: the reference's `ds_arr == None` mask is all False, its EDT then measures to the virtual
            # point (-1, 0) and cells within DISTANCE_THRESHOLD of it would be assigned None -- a numpy error there.  Nothing
            # can be relabelled without a particle class: leave the image as it is.
            continue
        updated_ds_arr, overlap_area = fill_particle_area(ds_arr, particle_label, cell_type_label, overlap_label=particle_label)
        particle_area += overlap_area
        ds_arr = updated_ds_arr
    return ds_arr, particle_area


def combine_cell_positions_and_clusters(dapi_channel, other_channel):
    """tiff_analysis.py:252-287: DAPI cells overlapping the other channel's cells by more
    than ``DAPI_RFP_OVERLAP_THRESHOLD`` of their area are rewritten to class 2."""
    np_in = _io.is_numpy(dapi_channel)
    cell_to_be_removed = 2
    d = _u8_image(dapi_channel)
    o = _u8_image(other_channel)
    W = int(d.shape[2])
    dapi_mask = ops.compare(d, "==", 1)[0]
    rfp_mask = ops.compare(o, "==", 1)[0]
    labels, counts, _ = ops.label_bits(dapi_mask, W, connectivity=8, dtype=torch.int32)
    n = int(counts[0].item())
    table = ops.new_table(max(1, n), d.device)
    ops.region_table(labels, None, table, fg_bits=dapi_mask, ov_bits=rfp_mask)
    tab = table.cpu().numpy()
    keep = np.zeros((1, n + 1), dtype=np.uint8)
    if n:
        frac = tab[ops.T_OVERLAP, :n] / tab[ops.T_AREA, :n].astype(np.float64)
        keep[0, 1:] = frac > DAPI_RFP_OVERLAP_THRESHOLD
    cells_to_remove = ops.select_labels(labels, torch.from_numpy(keep).to(d.device))
    dapi_combined = d.clone()
    ops.assign_where_u8_(dapi_combined, cells_to_remove, cell_to_be_removed)
    return _io.back(dapi_combined[0], np_in)


def _apply_lut_inplace(arr, steps):
    lut = np.arange(256, dtype=np.uint8)
    for a, b in steps:  # ordered arr[arr == a] = b chain -> one table
        lut[lut == a] = b
    if _io.is_numpy(arr):
        t = _u8_image(arr)
        ops.lut_u8_(t, lut)
        arr[...] = t[0].cpu().numpy()
    else:
        ops.lut_u8_(arr, lut)
    return arr


def get_rfp_base_arr(rfp_arr, cell_strains):
    """tiff_analysis.py:224-231 (in place)."""
    if cell_strains == ["6B07"] or cell_strains == ["6B07", "C3M10"]:
        return _apply_lut_inplace(rfp_arr, [(1, 4), (2, 5)])
    return _apply_lut_inplace(rfp_arr, [(2, 4), (3, 5)])


def relabel_other_channel(other_channel, other_channel_name):
    """tiff_analysis.py:177-181 on a copy: 3->5, 2->4 and, for GFP, 1->3."""
    out = other_channel.copy() if _io.is_numpy(other_channel) else other_channel.clone()
    steps = [(3, 5), (2, 4)] + ([(1, 3)] if other_channel_name == "GFP" else [])
    return _apply_lut_inplace(out, steps)


def combine_channels(rfp_base, channel_ds_arrs, cell_strains):
    """tiff_analysis.py:233-249 (in place on ``rfp_base``)."""
    np_in = _io.is_numpy(rfp_base)
    base = _u8_image(rfp_base) if np_in else rfp_base.unsqueeze(0)
    W = int(base.shape[2])
    for strain in cell_strains:
        if strain == "3D05":
            continue
        channel_name = STRAIN_MAP[strain]
        for val, strain_name in BASE_TYPE_MAP.items():
            if strain_name == strain:
                bits = ops.compare(_u8_image(channel_ds_arrs[channel_name]), "==", 1)[0]
                ops.assign_where_u8_(base, bits, val)
    if np_in:
        rfp_base[...] = base[0].cpu().numpy()
    return rfp_base


def get_cell_cell_distances(cell_pos):
    """Nearest-neighbour distances between cells, within and across strains -- goal 3 of the author's notes
    (refine_boundaries.py:8-12: "the distance between each cell of a given strain and its nearest neighbor of
    the same strain, and ... of a different strain"), not yet written in the reference; the MATLAB script's
    ``pdist2`` + ``min`` block (.m:260-263) is the model.  ``cell_pos`` is the dict of region lists that
    ``get_cell_positions_and_areas`` returns.  Result: ``{(strain_a, strain_b): (distances, indices)}`` with one
    entry per cell of ``strain_a`` (centroid to centroid, pixels; ``inf`` / ``-1`` when ``strain_b`` has no
    other cell), indices into ``cell_pos[strain_b]``."""
    import torch

    from . import _io, ops

    dev = _io.device()
    pts = {k: torch.as_tensor(np.array([r.centroid for r in v], dtype=np.float64).reshape(-1, 2), device=dev) for k, v in cell_pos.items()}
    out = {}
    for a, pa in pts.items():
        for b, pb in pts.items():
            if pa.shape[0] == 0:
                out[(a, b)] = (np.zeros(0), np.zeros(0, dtype=np.int64))
                continue
            d, j = ops.nearest(pa, pb, exclude_self=(a == b))
            out[(a, b)] = (d.cpu().numpy(), j.cpu().numpy())
    return out


def get_cell_counts_and_densities(cell_pos, cell_clusters, particle_area):
    """tiff_analysis.py:1018-1038 -- host arithmetic over the region lists."""
    cell_count, cell_density, cell_area_ratio = {}, {}, {}
    particle_area = particle_area / (PX_TO_UM_CONV**2)
    for cell_type, cell_array in cell_pos.items():
        if cell_type not in CELL_TYPES:
            continue
        cluster_cells = 0
        for cluster in cell_clusters[cell_type]:
            cluster_cells += cluster.cells
        cell_count[cell_type] = len(cell_array) + cluster_cells
        cell_area = np.sum([cell.area for cell in cell_array])
        for cluster in cell_clusters[cell_type]:
            cell_area += cluster["area"]
        area = cell_area / (PX_TO_UM_CONV**2)
        cell_density[cell_type] = round(cell_count[cell_type] / particle_area, 5)
        cell_area_ratio[cell_type] = round(area / particle_area, 5)
    return cell_count, cell_density, cell_area_ratio


def process_single_array(ds_arr, cell_types):
    """The pixel part of ``process_single_h5_file`` (tiff_analysis.py:642-651): normalise,
    denoise, measure, recreate the particle area.  File I/O, plots and CSVs stay with the caller."""
    ds_arr = normalize_ds_arr(ds_arr)
    ds_arr_denoised = median_filter(ds_arr, size=DENOISE_SIZE)
    cell_positions, cell_clusters, particle_area, merged_clusters = get_cell_positions_and_areas(ds_arr_denoised, cell_types, merged=True)
    cell_count, cell_density, cell_area_ratio = get_cell_counts_and_densities(cell_positions, cell_clusters, particle_area)
    ds_arr_recreated, particle_area = recreate_particle_area(ds_arr_denoised, cell_types, particle_area)
    return {
        "denoised": ds_arr_denoised,
        "cell_positions": cell_positions,
        "cell_clusters": cell_clusters,
        "merged_clusters": merged_clusters,
        "cell_count": cell_count,
        "cell_density": cell_density,
        "cell_area_ratio": cell_area_ratio,
        "recreated": ds_arr_recreated,
        "particle_area": particle_area,
    }


def load_class_image(full_file_path):
    """The read at the top of ``process_single_h5_file`` / ``process_multiple_h5_files`` (tiff_analysis.py:118-121,
    :639-642): the first dataset of the ilastik export, squeezed to 2-D.  ``h5_io`` stands in for h5py, which is not
    installable here (its compatibility with files written by libhdf5 itself is unverified, see ``h5_io``)."""
    from . import h5_io

    return normalize_ds_arr(h5_io.read_first_dataset(full_file_path))


def process_single_h5_file(full_file_path, cell_types):
    """tiff_analysis.py:627-671 without the plots and CSV files: read the export, then ``process_single_array``."""
    return process_single_array(load_class_image(full_file_path), cell_types)

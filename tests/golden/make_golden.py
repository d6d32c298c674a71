"""Generate ``tests/golden/*.npz|json`` by running the REFERENCE's own functions.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

It imports ``/root/reference/tiff_analysis.py`` unmodified through
``oracle.ref_loader`` (scipy + the scikit-image shim underneath) and freezes the
outputs of its L2 functions on small seeded class images.  The GPU box has no
``/root/reference``; tests there read these files.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import l2, ref_loader  # noqa: E402
from particle_col_image_segmentation_b200 import synth  # noqa: E402


def canon(summary):
    # "combined" order follows Python set iteration order in the reference
    # (tiff_analysis.py:794-795) -> sort by first member label
    for k in summary["merged"]:
        summary["merged"][k] = sorted(summary["merged"][k], key=lambda m: m[3][0])
    return summary


def main():
    ta = ref_loader.load_tiff_analysis()
    arrays, meta = {}, {}

    # case A: single-channel file path (tiff_analysis.py:627-671)
    raw = synth.class_image(384, 384, seed=4321, noise=0.02)
    den = ta.median_filter(raw, size=ta.DENOISE_SIZE)
    types = {1: "3D05", 2: "Particle", 3: "Background"}
    res = ta.get_cell_positions_and_areas(den, types, merged=True)
    meta["A_positions"] = canon(l2.summarize_positions(res))
    cnt, dens, ratio = ta.get_cell_counts_and_densities(res[0], res[1], res[2])
    meta["A_counts"] = {"count": {k: int(v) for k, v in cnt.items()}, "density": dens, "ratio": ratio}
    rec, area = ta.recreate_particle_area(den, types, res[2])
    meta["A_particle_area"] = float(area)
    arrays.update(A_raw=raw, A_denoised=den, A_recreated=rec)
    _, merged_images = ta.get_cell_clusters_from_distances(den, res[0], res[1], types)
    for k, im in merged_images.items():
        arrays[f"A_merged_image_{k}"] = im

    # case B: combined-channel image with the base type map (tiff_analysis.py:206)
    comb = synth.multi_class_image(320, 320, seed=99)
    resb = ta.get_cell_positions_and_areas(comb, ta.BASE_TYPE_MAP, merged=True)
    meta["B_positions"] = canon(l2.summarize_positions(resb))
    arrays["B_image"] = comb
    recb, areab = ta.recreate_particle_area(comb, ta.BASE_TYPE_MAP, resb[2])
    meta["B_particle_area"] = float(areab)
    arrays["B_recreated"] = recb

    # case C: DAPI / RFP overlap removal and channel combining (tiff_analysis.py:167-204)
    dapi = ta.median_filter(synth.class_image(256, 256, seed=7, noise=0.01), size=5)
    rfp = ta.median_filter(synth.class_image(256, 256, seed=8, noise=0.01), size=5)
    rng = np.random.default_rng(5)
    # make a third of the DAPI cells coincide with RFP cells
    take = (dapi == 1) & (rng.random(dapi.shape) < 0.4)
    rfp = rfp.copy()
    rfp[take] = 1
    upd = ta.combine_cell_positions_and_clusters(dapi, rfp)
    arrays.update(C_dapi=dapi, C_rfp=rfp, C_dapi_updated=upd)
    strains = ["3D05", "6B07"]
    base = ta.get_rfp_base_arr(rfp.copy(), strains)
    arrays["C_rfp_base"] = base.copy()
    arrays["C_combined"] = ta.combine_channels(base.copy(), {"RFP": rfp, "DAPI": dapi}, strains)
    other = rfp.copy()
    other[other == 3] = 5
    other[other == 2] = 4
    arrays["C_other_updated"] = other
    up, n = ta.fill_particle_area(dapi, 2, 1, 2)
    arrays["C_fill"] = up
    meta["C_fill_count"] = int(n)

    np.savez_compressed(os.path.join(HERE, "reference_l2.npz"), **arrays)
    with open(os.path.join(HERE, "reference_l2.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(arrays), "arrays;", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in meta.items()})


if __name__ == "__main__":
    main()

"""Generate ``tests/golden/*.npz|json`` by running the REFERENCE's own functions.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py [--skimage auto|real|shim]

``--skimage real`` insists on a real scikit-image (searched on ``sys.path`` and in ``baseline/_ref``, where the
driver's reference install would put it) and fails without one; ``auto`` (default) uses a real one if importable
and the restatement in ``oracle/skimage_shim`` otherwise; ``shim`` forces the restatement.  Which one produced the
fixtures is recorded in ``reference_l2.json`` under ``_generated_with``.

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


def pick_skimage(mode):
    """-> description of the scikit-image layer the reference will import."""
    root = os.path.dirname(os.path.dirname(HERE))
    ref_install = os.path.join(root, "baseline", "_ref")
    if mode in ("auto", "real") and os.path.isdir(ref_install) and ref_install not in sys.path:
        sys.path.insert(0, ref_install)
    real = None
    if mode != "shim":
        try:
            import skimage

            if not getattr(skimage, "__version__", "").endswith("+shim"):
                real = skimage.__version__
        except ImportError:
            pass
    if mode == "real" and real is None:
        raise SystemExit("make_golden: --skimage real, but no real scikit-image is importable (not in this image, not under baseline/_ref)")
    if real is None:
        for name in [m for m in sys.modules if m == "skimage" or m.startswith("skimage.")]:
            del sys.modules[name]  # a half-imported real package must not shadow the shim
        return "oracle/skimage_shim (restatement of scikit-image 0.25.2; the real library is not installable offline)"
    return f"scikit-image {real} (real library)"


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":
        # the reference iterates a set of cell-type names (tiff_analysis.py:794-795): with hash randomisation the member
        # order of the "combined" groups, and with it the last bits of their area-weighted centroids, changes from run
        # to run.  Fix the hash seed so that regenerating the fixtures reproduces them byte for byte.
        os.execve(sys.executable, [sys.executable] + sys.argv, dict(os.environ, PYTHONHASHSEED="0"))
    mode = "auto"
    if "--skimage" in sys.argv:
        mode = sys.argv[sys.argv.index("--skimage") + 1]
    if mode not in ("auto", "real", "shim"):
        raise SystemExit("--skimage takes auto, real or shim")
    layer = pick_skimage(mode)
    ta = ref_loader.load_tiff_analysis()  # installs the shim only if no real scikit-image is importable
    arrays, meta = {}, {"_generated_with": {"skimage": layer, "reference": "tiff_analysis.py, unmodified, via oracle/ref_loader.py"}}

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

"""Randomised differential run of the tiff_analysis mirrors (L2 functions) against oracle/l2.py, the
restatement pinned to the unmodified reference by tests/golden (not part of the test suite)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from scipy import ndimage as ndi
from oracle import l2 as ol2
from helpers import assert_summary_equal
from particle_col_image_segmentation_b200 import synth, tiff_analysis as ta

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(time.time()) % 100000)
t0 = time.time(); n = 0; bad = 0
types = {1: "C3M10", 2: "Particle", 3: "Background"}
while time.time() - t0 < budget:
    n += 1
    H = int(rng.integers(40, 400)); W = int(rng.integers(40, 500)); seed = int(rng.integers(1 << 30))
    noise = float(rng.choice([0.0, 0.01, 0.05, 0.15]))
    info = (H, W, seed, noise)
    try:
        raw = synth.class_image(H, W, seed=seed, noise=noise)
        den = ta.median_filter(raw, size=5)
        assert np.array_equal(den, ndi.median_filter(raw, size=5)), "median"
        try:
            want = ol2.get_cell_positions_and_areas(den, types, merged=True)
        except ValueError as e_ref:  # the reference divides by the mean of an empty list when an image has clusters but no cells (tiff_analysis.py:781)
            try:
                ta.get_cell_positions_and_areas(den, types, merged=True)
                raise AssertionError("reference raises, device path does not")
            except ValueError as e_dev:
                assert str(e_dev) == str(e_ref), "different error"
                continue
        got = ta.get_cell_positions_and_areas(den, types, merged=True)
        assert_summary_equal(ol2.summarize_positions(got), ol2.summarize_positions(want))
        g, ng = ta.recreate_particle_area(den, types, got[2])
        w, nw = ol2.recreate_particle_area(den, types, want[2])
        assert np.array_equal(g, w) and ng == nw, "recreate"
        other = synth.class_image(H, W, seed=seed + 50, noise=0.0)
        assert np.array_equal(ta.combine_cell_positions_and_clusters(den, other), ol2.combine_cell_positions_and_clusters(den, other)), "combine"
        up, cnt = ta.fill_particle_area(den, 2, 1, 2)
        up2, cnt2 = ol2.fill_particle_area(den, 2, 1, 2)
        assert np.array_equal(up, up2) and int(cnt) == int(cnt2), "fill_particle_area"
        c1 = ta.get_cell_counts_and_densities(got[0], got[1], got[2]); c2 = ol2.get_cell_counts_and_densities(want[0], want[1], want[2])
        assert c1 == c2, "counts"
    except AssertionError as e:
        bad += 1; print("MISMATCH", str(e)[:120], info, flush=True)
    except Exception as e:  # noqa: BLE001
        bad += 1; print("ERROR", type(e).__name__, str(e)[:160], info, flush=True)
print(f"fuzz L2: {n} cases, {bad} bad, {time.time() - t0:.0f} s")

"""Randomised differential run of the device pipeline and primitives against the oracle (not part of the
test suite; run on a GPU box: python scratch/fuzz_pipeline.py [seconds])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy import ndimage as ndi
from oracle import pipeline as opipe
from oracle.skimage_shim.morphology import remove_small_objects
from particle_col_image_segmentation_b200 import split_zstack, synth, ops, ndimage as pnd, measure as pm

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(time.time()) % 100000)
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < budget:
    n += 1
    kind = rng.integers(0, 4)
    Z = int(rng.integers(1, 4)); H = int(rng.integers(1, 300)); W = int(rng.integers(1, 700))
    if kind == 0:
        st = synth.zstack_u16(Z, max(H, 8), max(W, 8), seed=int(rng.integers(1 << 30)))
    elif kind == 1:
        st = rng.integers(0, int(rng.choice([2, 300, 65536])), (Z, H, W)).astype(np.uint16)
    elif kind == 2:
        st = (rng.random((Z, H, W)) < rng.uniform(0.05, 0.95)).astype(np.uint16) * int(rng.integers(1, 60000))
        st += rng.integers(0, 3, st.shape).astype(np.uint16)
    else:
        base = ndi.gaussian_filter(rng.random((Z, H, W)), (0, rng.uniform(0.5, 6), rng.uniform(0.5, 6)))
        st = (base * 60000).astype(np.uint16)
    dn = int(rng.choice([0, 3, 5, 7])); ms = int(rng.choice([1, 2, 20, 200])); ch = int(rng.integers(1, 4))
    try:
        got = split_zstack.segment_zstack(st, denoise_size=dn, min_size=ms, chunk=ch)
        want = opipe.segment_zstack(st, denoise_size=dn, min_size=ms)
        for k in ("threshold", "mask", "labels", "refined", "edt", "table", "counts"):
            if not np.array_equal(got[k], want[k]):
                bad += 1
                print("MISMATCH", k, st.shape, kind, dn, ms, ch, flush=True)
                np.save(f"/tmp/fuzz_bad_{bad}.npy", st)
                break
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("ERROR", type(e).__name__, str(e)[:200], st.shape, kind, dn, ms, ch, flush=True)
print(f"fuzz: {n} cases, {bad} bad, {time.time() - t0:.0f} s")

"""Seeded differential fuzz of the device path against the oracle (fixed seed lists; the open-ended versions are
scratch/fuzz_*.py, which draw their cases from the same generators in tests/fuzz_cases.py)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fuzz_cases  # noqa: E402

pytestmark = pytest.mark.gpu


def test_fuzz_pipeline_seeded():
    """The whole segment pipeline on random stacks: sizes 1x1 .. 300x700, blobs / noise / binary / smooth fields,
    every denoise size, min_size and chunking; every output bit-exact."""
    for seed in range(7000, 7160):
        fuzz_cases.check_pipeline(seed)


def test_fuzz_primitives_seeded():
    """Labelling (both connectivities, multi-valued), hole filling, EDT, disk dilation / erosion, small objects,
    local maxima, median, Otsu on random images."""
    for seed in range(8000, 8200):
        fuzz_cases.check_primitives(seed)


def test_fuzz_l2_seeded():
    """The tiff_analysis mirrors against oracle/l2.py, including the reference's ValueError on images with clusters
    but no cells (tiff_analysis.py:781)."""
    outcomes = [fuzz_cases.check_l2(seed) for seed in range(9000, 9080)]
    assert outcomes.count("ok") >= 40

"""Shared comparison helpers for the parity tests."""

import numpy as np


def canon_summary(summary):
    """Sort merged groups by first member label (the reference's "combined" order
    follows Python set iteration, tiff_analysis.py:794-795) and turn tuples into lists."""
    import json

    s = json.loads(json.dumps(summary))
    for k in s["merged"]:
        s["merged"][k] = sorted(s["merged"][k], key=lambda m: m[3][0])
    return s


def assert_summary_equal(got, want, exact_groups=("cell_pos", "cell_clusters")):
    got, want = canon_summary(got), canon_summary(want)
    assert got["particle_area"] == want["particle_area"]
    for grp in exact_groups:
        assert set(got[grp]) == set(want[grp]), grp
        for k in want[grp]:
            assert got[grp][k] == want[grp][k], (grp, k)  # bit-exact areas, centroids, bboxes
    assert set(got["merged"]) == set(want["merged"])
    for k in want["merged"]:
        assert len(got["merged"][k]) == len(want["merged"][k]), k
        for g, w in zip(got["merged"][k], want["merged"][k]):
            assert g[0] == w[0] and g[2] == w[2] and sorted(g[3]) == sorted(w[3]), k
            if k == "combined":  # member order depends on set iteration -> fp summation order
                np.testing.assert_allclose(g[1], w[1], rtol=1e-12)
            else:
                assert g[1] == w[1], k

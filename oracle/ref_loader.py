"""Import the reference's own ``tiff_analysis.py`` UNMODIFIED, for fixtures only.

TEST INFRASTRUCTURE ONLY.  Works only where ``/root/reference`` exists (the
build container).  ``tiff_analysis.py:36-45`` imports h5py, matplotlib and
scikit-image, none of which is installed; h5py/matplotlib are never reached by
the L2 functions, so empty stub modules stand in for them, and scikit-image is
replaced by ``oracle.skimage_shim``.  The GPU box has no ``/root/reference``:
nothing that runs there calls this module (``available()`` is False).
"""

import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("PCS_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "tiff_analysis.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_tiff_analysis():
    """Return the reference module ``tiff_analysis`` (tiff_analysis.py:1-1138)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    from . import skimage_shim

    skimage_shim.install()
    _stub("h5py")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.colors = _stub("matplotlib.colors")
    mpl.patches = _stub("matplotlib.patches", Rectangle=object)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    return importlib.import_module("tiff_analysis")

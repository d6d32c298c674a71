"""Minimal CPU restatement of the scikit-image 0.25.2 calls the reference makes.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

scikit-image is pinned by the reference (``uv.lock:859-860``) but is absent from
this image and cannot be installed, so each function below restates the
published algorithm on top of ``scipy.ndimage`` / numpy.  Call sites in the
reference:

* ``skimage.measure.label``       tiff_analysis.py:260, :743, :829; refine_boundaries.py:64
* ``skimage.measure.regionprops`` tiff_analysis.py:263, :746
* ``skimage.morphology.disk``     tiff_analysis.py:827, :990
* ``skimage.morphology.binary_dilation``  tiff_analysis.py:828, :990
* ``skimage.morphology.local_maxima``     refine_boundaries.py:63
* ``skimage.segmentation.watershed``      refine_boundaries.py:73

plus the functions the north_star names without a reference call site
(``filters.threshold_otsu``, ``morphology.binary_erosion/opening/closing``,
``morphology.remove_small_objects``).

``install()`` registers the shim in ``sys.modules`` under the name ``skimage``
so that the reference's own modules can be imported unmodified.
"""

import sys

from . import filters, measure, morphology, segmentation  # noqa: F401
from .measure import _regionprops  # noqa: F401

__version__ = "0.25.2+shim"


def install():
    """Register this package as ``skimage`` (only if the real one is missing)."""
    try:  # pragma: no cover - the real library is never present in this image
        import skimage  # noqa: F401

        if not getattr(skimage, "__version__", "").endswith("+shim"):
            return False
    except ImportError:
        pass
    me = sys.modules[__name__]
    sys.modules["skimage"] = me
    sys.modules["skimage.measure"] = measure
    sys.modules["skimage.measure._regionprops"] = _regionprops
    sys.modules["skimage.morphology"] = morphology
    sys.modules["skimage.filters"] = filters
    sys.modules["skimage.segmentation"] = segmentation
    io_stub = type(sys)("skimage.io")
    sys.modules["skimage.io"] = io_stub
    me.io = io_stub
    return True

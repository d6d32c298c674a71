"""ctypes binding of ``libpcs.so`` (the C ABI declared in ``include/pcs.h``).

The product has NO CPU fallback: if the shared library is missing and cannot be
built, importing any compute entry point raises.  Calls that return a negative
status raise ``PcsError`` with ``pcs_last_error_string()``.
"""

import ctypes
import os
import shutil
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint16, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCS_LIB_PATH") or os.path.join(HERE, "libpcs.so")  # override: A/B builds while tuning


class PcsError(RuntimeError):
    pass


_P = c_void_p  # every device pointer crosses the ABI as a plain address
_I, _L, _Z = c_int, c_int64, c_size_t

# name -> (restype, argtypes); mirrors include/pcs.h one to one
SIGNATURES = {
    "pcs_version": (c_int, []),
    "pcs_last_error_string": (c_char_p, []),
    "pcs_device_sm_count": (c_int, [_I]),
    "pcs_kernel_launches": (c_uint64, []),
    "pcs_profile_enable": (c_int, [_I]),
    "pcs_profile_collect": (c_int, [_P, _P, _P, _I]),
    "pcs_compare_u8": (c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P]),
    "pcs_compare_u16": (c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P]),
    "pcs_compare_i32": (c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P]),
    "pcs_compare_f32": (c_int, [_P, c_float, _P, _I, _P, _P, _I, _I, _I, _P]),
    "pcs_compare_f64": (c_int, [_P, c_double, _P, _I, _P, _P, _I, _I, _I, _P]),
    "pcs_member_u8": (c_int, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "pcs_unpack_bits": (c_int, [_P, _P, _I, _I, _I, _P]),
    "pcs_bits_logic": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "pcs_bits_count": (c_int, [_P, _P, _I, _I, _I, _P]),
    "pcs_lut_u8": (c_int, [_P, _P, _L, _P]),
    "pcs_assign_where_u8": (c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "pcs_gather": (c_int, [_P, _I, _P, _P, _P, _L, _L, _P]),
    "pcs_max_label": (c_int, [_P, _I, _L, _P, _P]),
    "pcs_transpose": (c_int, [_P, _P, _I, _I, _I, _P]),
    "pcs_fill_u32": (c_int, [_P, c_uint32, _Z, _P]),
    "pcs_zero_background": (c_int, [_P, _Z, _I, _P]),
    "pcs_histogram_bytes": (_Z, [_I]),
    "pcs_histogram_u16": (c_int, [_P, _P, _I, _I, _I, _P]),
    "pcs_otsu_u16": (c_int, [_P, _P, _P, _I, _L, _P]),
    "pcs_median_u8": (c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "pcs_majority_bits": (c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "pcs_majority_bits_mask": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "pcs_ccl_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "pcs_label_bits": (c_int, [_P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _L, _P, _Z, _P]),
    "pcs_conn_planes_bytes": (_Z, [_I, _I, _I]),
    "pcs_conn_planes": (c_int, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "pcs_label_conn": (c_int, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _L, _P, _Z, _P]),
    "pcs_fill_holes_bits": (c_int, [_P, _P, _I, _I, _I, _P, _Z, _P]),
    "pcs_fill_holes_table_workspace_bytes": (_Z, [_I, _I, _I]),
    "pcs_fill_holes_table_bits": (c_int, [_P, _P, _L, _P, _L, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "pcs_refine_labeled_bits": (c_int, [_P, _P, _P, _L, _P, _L, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "pcs_remove_small_bits": (c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "pcs_select_components_bits": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "pcs_local_maxima_conn": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "pcs_dilate_bits": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pcs_edt_workspace_bytes": (_Z, [_I, _I, _I]),
    "pcs_edt_bits": (c_int, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _Z, _P]),
    "pcs_table_init": (c_int, [_P, _L, _P]),
    "pcs_table_init_rows": (c_int, [_P, _L, _P, _I, _P]),
    "pcs_region_table": (c_int, [_P, _I, _P, _I, _P, _P, _P, _P, _L, _I, _I, _I, _P]),
    "pcs_select_labels": (c_int, [_P, _I, _P, _L, _P, _I, _I, _I, _P]),
    "pcs_select_by_area": (c_int, [_P, _P, _P, _L, _P, _L, _P, _I, _I, _I, _P]),
    "pcs_roi_sums_f64": (c_int, [_P, _P, _I, _L, _I, _P, _P]),
    "pcs_min_dist_f64": (c_int, [_P, _L, _P, _L, _P, _P]),
    "pcs_nearest_f64": (c_int, [_P, _L, _P, _L, _I, _P, _P, _P]),
    "pcs_watershed_workspace_bytes": (_Z, [_I, _I, _I]),
    "pcs_watershed_f64": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "pcs_gauss_f64": (c_int, [_P, _P, _P, c_double, _I, _I, _I, _P]),
    "pcs_ratio_f64": (c_int, [_P, _P, _P, _P, _P, _P, _L, _P]),
    "pcs_scale_u8_f64": (c_int, [_P, _P, _P, _L, _P]),
    "pcs_resize_taps_f64": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _L, _L, _L, _P]),
    "pcs_segment_workspace_bytes": (_Z, [_I, _I, _I]),
    "pcs_segment_chunk": (c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _I, _P, _Z, _P]),
    "pcs_table_finalize": (c_int, [_P, _L, _P, _I, _I, c_double, _P, _P]),
    "pcs_table_finalize_ex": (c_int, [_P, _L, _P, _I, _I, c_double, _P, _L, _P, _P]),
}

_lib = None


def _build_if_possible():
    from . import build as _build

    if shutil.which(_build.NVCC) or os.path.exists(_build.NVCC):
        _build.build()


def load():
    """Load (building first if a compiler is at hand) and bind libpcs.so."""
    global _lib
    if _lib is not None:
        return _lib
    default_lib = "PCS_LIB_PATH" not in os.environ
    stale = False
    if default_lib:
        from . import build as _build

        stale = _build.needs_build()  # missing, or built from other sources than the ones in the tree (content hash)
    if not os.path.exists(LIB_PATH) or stale:
        try:
            _build_if_possible()
        except Exception as e:  # noqa: BLE001
            raise PcsError(f"libpcs.so is {'stale' if os.path.exists(LIB_PATH) else 'missing'} and could not be built: {e}") from e
        if default_lib and _build.needs_build():
            raise PcsError("libpcs.so does not match the sources in csrc/ and no nvcc is at hand to rebuild it; run `python __graft_entry__.py` where nvcc is. A stale library is never used.")
    if not os.path.exists(LIB_PATH):
        raise PcsError(f"libpcs.so not found at {LIB_PATH}; run `python __graft_entry__.py` (build()) first. There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.pcs_version() != 200:
        raise PcsError(f"libpcs.so version {lib.pcs_version()} does not match the bindings (200)")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().pcs_last_error_string()
        raise PcsError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def call(name, *args):
    """Invoke an int-returning entry point and raise on a negative status."""
    check(getattr(load(), name)(*args), name)

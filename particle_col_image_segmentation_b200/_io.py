"""numpy <-> device plumbing for the drop-in entry points.

The reference's boundary is "numpy arrays in, numpy arrays out" (SURVEY.md 8b).
Every public function accepts a numpy array (copied to the current CUDA device)
or a CUDA torch tensor (used in place) and answers in kind.
"""

import numpy as np
import torch

from . import _lib


def device():
    if not torch.cuda.is_available():
        raise _lib.PcsError("no CUDA device: particle_col_image_segmentation_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def is_numpy(a):
    return not isinstance(a, torch.Tensor)


def to_device(a, dtype=None):
    """-> contiguous CUDA tensor (no leading batch axis added)."""
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(device())
    else:
        arr = np.ascontiguousarray(a)
        if arr.dtype == np.bool_:
            arr = arr.view(np.uint8)
        if not arr.flags.writeable:
            arr = arr.copy()
        t = torch.from_numpy(arr).to(device())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def image_2d(a, dtype=None):
    """2-D image -> (1, H, W) CUDA tensor."""
    t = to_device(a, dtype)
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D image, got shape {tuple(t.shape)}")
    return t.unsqueeze(0)


def mask_bits(a):
    """2-D mask (bool / any numeric, non-zero = True) -> ((1, H, WW) bit image, H, W)."""
    from . import ops

    t = image_2d(a)
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    if t.dtype not in (torch.uint8, torch.uint16, torch.int32, torch.float32, torch.float64):
        t = (t != 0).view(torch.uint8) if t.dtype != torch.uint8 else t
    bits = ops.compare(t, "!=", 0)[0]
    return bits, int(t.shape[1]), int(t.shape[2])


def bits_to_bool(bits, W, like_numpy=True):
    from . import ops

    m = ops.unpack(bits, W, dtype=torch.bool)[0]
    return m.cpu().numpy() if like_numpy else m


def back(t, like_numpy=True):
    return t.cpu().numpy() if like_numpy else t

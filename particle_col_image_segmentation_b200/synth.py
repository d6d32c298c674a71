"""Seeded synthetic inputs for the five BASELINE.json configs.

The reference ships no data (all ``.tif/.h5/.png`` are git-ignored,
``.gitignore:13-20``), so every test and benchmark input is generated here.
numpy generators are used for parity-sized inputs; ``zstack_u16_device`` renders
the large stacks directly in HBM (torch is used for device buffers only).
"""

import numpy as np


def _render_blobs(img, cy, cx, peak, sigma):
    """Add Gaussian-profile blobs to a float32 image in local windows."""
    h, w = img.shape
    for y0, x0, p, s in zip(cy, cx, peak, sigma):
        r = int(np.ceil(4 * s))
        ya, yb = max(0, int(y0) - r), min(h, int(y0) + r + 1)
        xa, xb = max(0, int(x0) - r), min(w, int(x0) + r + 1)
        if ya >= yb or xa >= xb:
            continue
        yy = np.arange(ya, yb, dtype=np.float32)[:, None] - np.float32(y0)
        xx = np.arange(xa, xb, dtype=np.float32)[None, :] - np.float32(x0)
        img[ya:yb, xa:xb] += np.float32(p) * np.exp(-(yy * yy + xx * xx) / np.float32(2 * s * s))


def blob_params(rng, h, w, n_blobs):
    return (
        rng.uniform(0, h, n_blobs),
        rng.uniform(0, w, n_blobs),
        rng.uniform(3000, 20000, n_blobs),
        rng.uniform(3, 8, n_blobs),
    )


def slice_u16(h=512, w=512, n_blobs=None, seed=1001, rng=None, params=None):
    """Config 1: background ``N(500, 40)`` clipped at 0 plus Gaussian-profile blobs
    (peak 3000-20000, sigma 3-8 px); ~60 blobs per 512x512."""
    rng = np.random.default_rng(seed) if rng is None else rng
    if n_blobs is None:
        n_blobs = max(1, int(round(60 * h * w / (512 * 512))))
    img = rng.normal(500.0, 40.0, (h, w)).astype(np.float32)
    if params is None:
        params = blob_params(rng, h, w, n_blobs)
    _render_blobs(img, *params)
    return np.clip(np.rint(img), 0, 65535).astype(np.uint16)


def zstack_u16(z=4, h=512, w=512, n_blobs=None, seed=1002, drift=2.0):
    """Config 2 (one channel of the ``(Z, C, Y, X)`` stack, split_zstack.py:50-58):
    the blob centres drift by up to ``drift`` px per slice."""
    rng = np.random.default_rng(seed)
    if n_blobs is None:
        n_blobs = max(1, int(round(1500 * h * w / (2048 * 2048))))
    cy, cx, peak, sigma = blob_params(rng, h, w, n_blobs)
    out = np.empty((z, h, w), dtype=np.uint16)
    for i in range(z):
        out[i] = slice_u16(h, w, rng=rng, params=(cy, cx, peak, sigma))
        cy = cy + rng.uniform(-drift, drift, n_blobs)
        cx = cx + rng.uniform(-drift, drift, n_blobs)
    return out


def zstack_u16_device(z, h, w, seed, device, n_blobs=None, drift=2.0):
    """Large stacks rendered in HBM: blob parameters from the numpy RNG, pixels
    and noise with torch on ``device`` (host RAM cannot hold config 5)."""
    import torch

    rng = np.random.default_rng(seed)
    if n_blobs is None:
        n_blobs = max(1, int(round(1500 * h * w / (2048 * 2048))))
    cy, cx, peak, sigma = blob_params(rng, h, w, n_blobs)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((z, h, w), dtype=torch.uint16, device=device)
    R = 32  # window half-size = 4 * max sigma
    for i in range(z):
        img = torch.empty((h, w), dtype=torch.float32, device=device).normal_(500.0, 40.0, generator=gen)
        tcy = torch.as_tensor(cy, device=device, dtype=torch.float32)
        tcx = torch.as_tensor(cx, device=device, dtype=torch.float32)
        tp = torch.as_tensor(peak, device=device, dtype=torch.float32)
        ts = torch.as_tensor(sigma, device=device, dtype=torch.float32)
        iy = tcy.floor().long()[:, None] + torch.arange(-R, R + 1, device=device)[None, :]
        ix = tcx.floor().long()[:, None] + torch.arange(-R, R + 1, device=device)[None, :]
        gy = torch.exp(-((iy.float() - tcy[:, None]) ** 2) / (2 * ts[:, None] ** 2))
        gx = torch.exp(-((ix.float() - tcx[:, None]) ** 2) / (2 * ts[:, None] ** 2))
        vals = tp[:, None, None] * gy[:, :, None] * gx[:, None, :]
        ok = ((iy >= 0) & (iy < h))[:, :, None] & ((ix >= 0) & (ix < w))[:, None, :]
        lin = iy.clamp(0, h - 1)[:, :, None] * w + ix.clamp(0, w - 1)[:, None, :]
        img.view(-1).index_add_(0, lin[ok], vals[ok])
        out[i] = img.round_().clamp_(0, 65535).to(torch.int32).to(torch.uint16)
        cy = cy + rng.uniform(-drift, drift, n_blobs)
        cx = cx + rng.uniform(-drift, drift, n_blobs)
    return out


def class_image(h=2048, w=2048, seed=1234, noise=0.01):
    """ilastik-style uint8 class image (exercises tiff_analysis.py:727-1015):
    3 = background, discs of class 2 (particle), small rectangles of class 1
    (cells), and ``noise`` uniform label noise.  Counts scale with the area."""
    rng = np.random.default_rng(seed)
    scale = h * w / (2048.0 * 2048.0)
    img = np.full((h, w), 3, dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(max(1, int(round(6 * scale)))):
        r = rng.uniform(120, 260) * min(1.0, np.sqrt(scale) * 1.5)
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        img[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 2
    for _ in range(max(4, int(round(1500 * scale)))):
        rh, rw = int(rng.integers(2, 7)), int(rng.integers(4, 13))
        if rng.random() < 0.5:
            rh, rw = rw, rh
        y0, x0 = int(rng.integers(0, max(1, h - rh))), int(rng.integers(0, max(1, w - rw)))
        img[y0 : y0 + rh, x0 : x0 + rw] = 1
    # a few large cell clusters so the cluster branch (tiff_analysis.py:772-781) is taken
    for _ in range(max(2, int(round(12 * scale)))):
        r = rng.uniform(9, 16)
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        img[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 1
    flip = rng.random((h, w)) < noise
    img[flip] = rng.integers(1, 4, size=int(flip.sum()), dtype=np.uint8)
    return img


def multi_class_image(h=512, w=512, seed=77, n_classes=3):
    """Combined-channel style image: classes 1..n_classes cells, n+1 particle, n+2 background."""
    rng = np.random.default_rng(seed)
    base = class_image(h, w, seed=seed, noise=0.0)
    img = np.where(base == 3, n_classes + 2, np.where(base == 2, n_classes + 1, 1)).astype(np.uint8)
    cells = img == 1
    # split the cell pixels among the classes by coarse blocks so regions stay coherent
    blocks = rng.integers(1, n_classes + 1, size=(h // 16 + 1, w // 16 + 1), dtype=np.uint8)
    cls = np.kron(blocks, np.ones((16, 16), dtype=np.uint8))[:h, :w]
    img[cells] = cls[cells]
    return img


def touching_particles(h=4096, w=4096, seed=1003, pitch=40.96):
    """Config 3: a jittered grid of touching / overlapping disks (radius 21-25 px at
    ``pitch``) as a bool mask, and an ilastik-like boundary-probability map
    (1 within 1.5 px of a disk edge, else 0, plus ``U(0, 0.2)`` noise) as float32."""
    rng = np.random.default_rng(seed)
    ny, nx = int(h / pitch), int(w / pitch)
    mask = np.zeros((h, w), dtype=bool)
    edge = np.zeros((h, w), dtype=bool)
    for gy in range(ny):
        for gx in range(nx):
            cy = (gy + 0.5) * pitch + rng.uniform(-3, 3)
            cx = (gx + 0.5) * pitch + rng.uniform(-3, 3)
            r = rng.uniform(21, 25) * min(1.0, pitch / 40.96)
            R = int(r + 3)
            ya, yb = max(0, int(cy) - R), min(h, int(cy) + R + 1)
            xa, xb = max(0, int(cx) - R), min(w, int(cx) + R + 1)
            yy = np.arange(ya, yb)[:, None] - cy
            xx = np.arange(xa, xb)[None, :] - cx
            d = np.sqrt(yy * yy + xx * xx)
            mask[ya:yb, xa:xb] |= d <= r
            edge[ya:yb, xa:xb] |= np.abs(d - r) <= 1.5
    prob = edge.astype(np.float32) + rng.uniform(0, 0.2, (h, w)).astype(np.float32)
    return mask, np.minimum(prob, 1.0).astype(np.float32)


def nanosims_stack(n=256, k=5, n_rois=120, seed=1004):
    """Config 4: ``k`` Poisson ion-count planes (float64), an int32 ROI label image
    holding two ROI sets (labels 1..Ra = "red", Ra+1..Ra+Rb = "green"; .m:91-104,
    :173), the set id per ROI, and a bool aggregate mask (.m:271-283)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n]
    roi = np.zeros((n, n), dtype=np.int32)
    lab = 0
    for _ in range(n_rois * 4):
        if lab >= n_rois:
            break
        a, b = rng.uniform(3, 9), rng.uniform(3, 9)
        cy, cx = rng.uniform(8, n - 8), rng.uniform(8, n - 8)
        th = rng.uniform(0, np.pi)
        u = (yy - cy) * np.cos(th) + (xx - cx) * np.sin(th)
        v = -(yy - cy) * np.sin(th) + (xx - cx) * np.cos(th)
        e = (u / a) ** 2 + (v / b) ** 2 <= 1.0
        grow = np.zeros_like(e)
        grow[1:, :] |= e[:-1, :]
        grow[:-1, :] |= e[1:, :]
        grow[:, 1:] |= e[:, :-1]
        grow[:, :-1] |= e[:, 1:]
        if (roi[e | grow] != 0).any() or e.sum() < 30:
            continue
        lab += 1
        roi[e] = lab
    set_id = (rng.random(lab) < 0.5).astype(np.int32) + 1  # 1 = red, 2 = green
    rates = rng.uniform(5, 500, k)
    planes = np.stack([rng.poisson(r * (1.0 + 0.5 * (roi > 0)), (n, n)).astype(np.float64) for r in rates])
    agg = (yy - n / 2) ** 2 / (0.42 * n) ** 2 + (xx - n / 2) ** 2 / (0.35 * n) ** 2 <= 1.0
    return planes, roi, set_id, agg

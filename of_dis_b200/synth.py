"""Seeded synthetic image pairs with known flow.

image : sum over sigma in {1.5,3,6,12,24} of sigma * gaussian_blur(N(0,1), sigma)
        on a (H+64)x(W+64)xC canvas, min-max scaled to [0,255]
flow  : u = A sin(4x/W+.3) cos(3y/H),  v = .6A cos(2.5x/W) sin(5y/H+.7)
        (stereo: u = -(2 + A(.5+.5 sin(3.1x/W + 2y/H))), v = 0)
I0    : centre crop of the canvas;  I1(x) = canvas(x - flow(x)) (cubic), so that
        I0(x) ~= I1(x + flow(x)) (the reference's convention, patch.cpp:217)
Both are quantised to uint8 because the reference CLI reads 8-bit images.
"""
from __future__ import annotations

import numpy as np


def synthetic_flow(h: int, w: int, amp: float = 6.0, stereo: bool = False):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    if stereo:
        u = -(2.0 + amp * (0.5 + 0.5 * np.sin(3.1 * x / w + 2.0 * y / h)))
        v = np.zeros_like(u)
    else:
        u = amp * np.sin(4.0 * x / w + 0.3) * np.cos(3.0 * y / h)
        v = 0.6 * amp * np.cos(2.5 * x / w) * np.sin(5.0 * y / h + 0.7)
    return u, v


def synthetic_pair(h: int, w: int, channels: int = 1, seed: int = 0, amp: float = 6.0, stereo: bool = False):
    """Returns (img0_u8, img1_u8, flow_gt[h,w,2]); images are (h,w) or (h,w,3)."""
    from scipy import ndimage

    rng = np.random.default_rng(seed)
    m = 32
    hc, wc = h + 2 * m, w + 2 * m
    canv = np.zeros((hc, wc, channels), dtype=np.float64)
    for c in range(channels):
        acc = np.zeros((hc, wc))
        for sigma in (1.5, 3.0, 6.0, 12.0, 24.0):
            acc += sigma * ndimage.gaussian_filter(rng.standard_normal((hc, wc)), sigma, mode="reflect")
        canv[..., c] = acc
    lo, hi = canv.min(), canv.max()
    canv = (canv - lo) / (hi - lo) * 255.0
    u, v = synthetic_flow(h, w, amp, stereo)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img0 = canv[m:m + h, m:m + w]
    # I1(x + f(x)) = I0(x); first-order inverse: I1(x) = canvas(x - f(x))
    img1 = np.empty_like(img0)
    for c in range(channels):
        img1[..., c] = ndimage.map_coordinates(canv[..., c], [yy + m - v, xx + m - u], order=3, mode="nearest")
    q = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    i0, i1 = q(img0), q(img1)
    if channels == 1:
        i0, i1 = i0[..., 0], i1[..., 0]
    return np.ascontiguousarray(i0), np.ascontiguousarray(i1), np.stack([u, v], -1).astype(np.float32)

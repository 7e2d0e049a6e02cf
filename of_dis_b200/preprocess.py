"""Host-side pre/post-processing around the hot path (numpy).

This is the arithmetic of /root/reference/run_dense.cpp:298-344 (pad to
2^lv_f, float conversion, x0.5 pyramid, Sobel/8 gradients, border padding) and
run_dense.cpp:407-414 (x2^lv_l upsampling, crop).  It is OUTSIDE the hot path
(SURVEY.md section 8f rank 1-2): the hot path's input is the padded pyramid.

For 8-bit input images every value of every pyramid level is a dyadic rational
that float32 holds exactly (8 integer bits + 2 fraction bits per level), so the
box mean and the Sobel/8 sums below are exact and therefore bit-identical to
OpenCV's evaluation order; tests/test_preprocess.py checks that against cv2
when it is importable.
"""
from __future__ import annotations

import math

import numpy as np


def pad_to_multiple(img: np.ndarray, lv_f: int):
    """run_dense.cpp:299-311: replicate-pad so width, height are divisible by 2^lv_f.
    Returns (padded, padw, padh)."""
    scfct = 2 ** lv_f
    h, w = img.shape[:2]
    padw = (scfct - w % scfct) % scfct
    padh = (scfct - h % scfct) % scfct
    if padw or padh:
        t, b = int(math.floor(padh / 2.0)), int(math.ceil(padh / 2.0))
        l, r = int(math.floor(padw / 2.0)), int(math.ceil(padw / 2.0))
        pads = ((t, b), (l, r)) + (((0, 0),) if img.ndim == 3 else ())
        img = np.pad(img, pads, mode="edge")
    return img, padw, padh


def half_size(img: np.ndarray) -> np.ndarray:
    """cv::resize(.5,.5,INTER_LINEAR) on even-sized float32 images == 2x2 box mean
    (run_dense.cpp:150)."""
    a = img[0::2, 0::2]
    b = img[0::2, 1::2]
    c = img[1::2, 0::2]
    d = img[1::2, 1::2]
    return (((a + b) + (c + d)) * np.float32(0.25)).astype(np.float32)


def sobel8(img: np.ndarray):
    """cv::Sobel(CV_32F, 3x3, scale 1/8, BORDER_DEFAULT=reflect101), run_dense.cpp:156-157.
    Returns (dx, dy)."""
    pads = ((1, 1), (1, 1)) + (((0, 0),) if img.ndim == 3 else ())
    p = np.pad(img, pads, mode="reflect")
    # dx: row filter [-1 0 1], column filter [1 2 1]/8
    t = p[:, 2:] - p[:, :-2]
    dx = (t[:-2] * np.float32(0.125) + t[1:-1] * np.float32(0.25)) + t[2:] * np.float32(0.125)
    # dy: row filter [1 2 1]/8 ... applied as column [-1 0 1] of the row-smoothed image
    s = (p[:, :-2] * np.float32(0.125) + p[:, 1:-1] * np.float32(0.25)) + p[:, 2:] * np.float32(0.125)
    dy = s[2:] - s[:-2]
    return dx.astype(np.float32), dy.astype(np.float32)


def build_pyramid(img_f32: np.ndarray, lv_f: int, imgpadding: int):
    """ConstructImgPyramide (run_dense.cpp:130-178): returns three lists indexed by
    level 0..lv_f of C-contiguous float32 arrays padded by `imgpadding` on all
    sides (image: replicate, gradients: zero)."""
    imgs, dxs, dys = [], [], []
    cur = np.ascontiguousarray(img_f32, dtype=np.float32)
    for i in range(lv_f + 1):
        if i > 0:
            cur = half_size(cur)
        dx, dy = sobel8(cur)
        pads = ((imgpadding, imgpadding), (imgpadding, imgpadding)) + (((0, 0),) if cur.ndim == 3 else ())
        imgs.append(np.ascontiguousarray(np.pad(cur, pads, mode="edge")))
        dxs.append(np.ascontiguousarray(np.pad(dx, pads, mode="constant")))
        dys.append(np.ascontiguousarray(np.pad(dy, pads, mode="constant")))
    return imgs, dxs, dys


class PairPyramids:
    """Everything OFClass's constructor takes for one image pair (oflow.h:84-111)."""

    def __init__(self, img0_u8: np.ndarray, img1_u8: np.ndarray, lv_f: int, imgpadding: int):
        assert img0_u8.shape == img1_u8.shape
        self.height_org, self.width_org = img0_u8.shape[:2]
        a, self.padw, self.padh = pad_to_multiple(img0_u8, lv_f)
        b, _, _ = pad_to_multiple(img1_u8, lv_f)
        self.height, self.width = a.shape[:2]
        self.noc = 1 if a.ndim == 2 else a.shape[2]
        self.lv_f = lv_f
        self.imgpadding = imgpadding
        self.i0, self.i0x, self.i0y = build_pyramid(a.astype(np.float32), lv_f, imgpadding)
        self.i1, self.i1x, self.i1y = build_pyramid(b.astype(np.float32), lv_f, imgpadding)

    def level_shape(self, lv: int):
        return self.height >> lv, self.width >> lv


def upsample_linear(flow: np.ndarray, s: int) -> np.ndarray:
    """cv::resize(fx=fy=s, INTER_LINEAR) for integer s: src = (dst+.5)/s-.5, edge clamped
    (run_dense.cpp:410)."""
    h, w = flow.shape[:2]

    def taps(n_src, n_dst):
        x = (np.arange(n_dst, dtype=np.float32) + np.float32(0.5)) / np.float32(s) - np.float32(0.5)
        x0 = np.floor(x).astype(np.int64)
        f = (x - x0).astype(np.float32)
        f[x0 < 0] = 0
        i0 = np.clip(x0, 0, n_src - 1)
        i1 = np.clip(x0 + 1, 0, n_src - 1)
        return i0, i1, f

    x0, x1, fx = taps(w, w * s)
    y0, y1, fy = taps(h, h * s)
    fl = flow.reshape(h, w, -1)
    fxb = fx[None, :, None]
    fyb = fy[:, None, None]
    rows = fl[:, x0] * (np.float32(1) - fxb) + fl[:, x1] * fxb
    out = rows[y0] * (np.float32(1) - fyb) + rows[y1] * fyb
    return out.astype(np.float32).reshape((h * s, w * s) + flow.shape[2:])


def postprocess(flow_level: np.ndarray, lv_l: int, padw: int, padh: int, width_org: int, height_org: int):
    """run_dense.cpp:407-414: scale by 2^lv_l, upsample, crop the divisibility padding."""
    out = flow_level
    if lv_l != 0:
        sc = 2 ** lv_l
        out = upsample_linear(out * np.float32(sc), sc)
    x0, y0 = int(math.floor(padw / 2.0)), int(math.floor(padh / 2.0))
    return np.ascontiguousarray(out[y0:y0 + height_org, x0:x0 + width_org])


def write_flo(path: str, flow: np.ndarray) -> None:
    """SaveFlowFile (run_dense.cpp:16-57): 'PIEH', int32 w, int32 h, float32 row-major."""
    h, w = flow.shape[:2]
    with open(path, "wb") as f:
        f.write(b"PIEH")
        np.array([w, h], dtype="<i4").tofile(f)
        np.ascontiguousarray(flow, dtype="<f4").tofile(f)


def read_flo(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        tag = f.read(4)
        if tag != b"PIEH":
            raise ValueError("not a .flo file")
        w, h = np.fromfile(f, dtype="<i4", count=2)
        data = np.fromfile(f, dtype="<f4")
    return data.reshape(h, w, -1)


def write_pfm(path: str, disp: np.ndarray) -> None:
    """SavePFMFile (run_dense.cpp:60-81): 'Pf', scale -1 (little endian), rows bottom-up, negated."""
    h, w = disp.shape[:2]
    with open(path, "wb") as f:
        f.write(("Pf\n%d %d\n%f\n" % (w, h, -1.0)).encode())
        np.ascontiguousarray(-disp.reshape(h, w)[::-1], dtype="<f4").tofile(f)

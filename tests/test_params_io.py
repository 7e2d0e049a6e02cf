"""Host-side logic that needs no GPU: operating-point presets (run_dense.cpp:225-268, README.md:48-64),
the 20-number form (run_dense.cpp:269-294) and the .flo/.pfm writers (run_dense.cpp:16-81)."""
import numpy as np
import pytest

from of_dis_b200 import params, preprocess


@pytest.mark.parametrize("width,lv_f", [(640, 5), (1024, 5), (1920, 6), (2880, 7), (256, 3), (100, 2)])
def test_first_scale_follows_the_width(width, lv_f):
    # AutoFirstScaleSelect (run_dense.cpp:180-183): floor(log2(2*width / (5 * patch size)))
    assert params.operating_point(2, width).sc_f == lv_f


def test_operating_point_presets():
    # (patch, overlap, levels below lv_f, iterations, refinement) per preset, run_dense.cpp:239-267
    want = {1: (8, 0.3, 2, 16, 0), 2: (8, 0.4, 2, 12, 1), 3: (12, 0.75, 4, 16, 1), 4: (12, 0.75, 5, 128, 1)}
    for op, (P, ov, dl, it, tv) in want.items():
        p = params.operating_point(op, 2880)
        assert (p.p_samp_s, p.max_iter, p.min_iter, p.usetvref) == (P, it, it, tv)
        assert abs(p.patove - ov) < 1e-7
        assert p.sc_l == max(p.sc_f - dl, 0)
        # shared defaults (run_dense.cpp:227-231)
        assert (p.usefbcon, p.patnorm, p.costfct, p.tv_innerit, p.tv_solverit) == (0, 1, 0, 1, 3)
        assert (p.dp_thresh, p.dr_thresh, p.res_thresh) == (0.05, 0.95, 0.0)
        assert (p.tv_alpha, p.tv_gamma, p.tv_delta, p.tv_sor) == (10.0, 10.0, 5.0, 1.6)
    assert params.operating_point(7, 1024).p_samp_s == 8  # unknown digit -> preset 2 (the switch's default)
    # steps = max(1, floor(P * (1 - overlap))) in float (oflow.cpp:91): 5, 4, 3, 3
    assert [params.operating_point(op, 1024).steps for op in (1, 2, 3, 4)] == [5, 4, 3, 3]


def test_twenty_number_form_uses_the_cli_order():
    # CLI order is ... usefbcon, patnorm, costfct, usetvref ... (README.md:80-81), the class API order differs
    p = params.from_cli_numbers("5 3 12 10 0.1 0.9 0.5 12 0.75 1 0 2 1 3 4 5 2 6 1.9 1".split(), noc=3, nop=1)
    assert (p.sc_f, p.sc_l, p.max_iter, p.min_iter, p.p_samp_s) == (5, 3, 12, 10, 12)
    assert (p.usefbcon, p.patnorm, p.costfct, p.usetvref) == (1, 0, 2, 1)
    assert (p.tv_alpha, p.tv_gamma, p.tv_delta, p.tv_innerit, p.tv_solverit, p.tv_sor) == (3.0, 4.0, 5.0, 2, 6, 1.9)
    assert (p.noc, p.nop, p.verbosity) == (3, 1, 1)
    with pytest.raises(ValueError):
        params.from_cli_numbers(["1"] * 7)
    c = p.to_c()
    assert (c.usefbcon, c.costfct, c.noc, c.patnorm) == (1, 2, 3, 0)


def test_flo_and_pfm_files(tmp_path):
    rng = np.random.default_rng(0)
    flow = rng.standard_normal((7, 11, 2)).astype(np.float32)
    path = str(tmp_path / "a.flo")
    preprocess.write_flo(path, flow)
    raw = open(path, "rb").read()
    assert raw[:4] == b"PIEH" and np.frombuffer(raw[4:12], "<i4").tolist() == [11, 7]   # run_dense.cpp:16-57
    assert np.array_equal(preprocess.read_flo(path), flow)
    disp = rng.standard_normal((5, 9, 1)).astype(np.float32)
    path = str(tmp_path / "a.pfm")
    preprocess.write_pfm(path, disp)
    with open(path, "rb") as f:                                                            # run_dense.cpp:60-81
        assert f.readline() == b"Pf\n"
        assert f.readline().split() == [b"9", b"5"]
        assert float(f.readline()) == -1.0
        data = np.fromfile(f, "<f4").reshape(5, 9)
    assert np.array_equal(-data[::-1], disp[..., 0])     # bottom-up rows, negated disparity


def _imgdump():
    import os

    from of_dis_b200 import build

    return os.path.join(build.build_host(), "ofdis_imgdump")


@pytest.mark.parametrize("ext", ["png", "ppm"])
def test_image_loader_equals_cv2_imread(ext, tmp_path):
    """run_* read images without OpenCV (host/run_dense.cpp load_image): a colour file read as gray must
    give cv2.imread(IMREAD_GRAYSCALE)'s bytes -- libpng's rgb_to_gray for PNG, cvtColor's fixed point for
    PPM --, read as colour its BGR bytes; gray files pass through."""
    import subprocess

    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    col = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    col[:5] = col[:5, :, :1]  # some r == g == b pixels
    gray = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    gext = "pgm" if ext == "ppm" else "png"
    fc, fg = str(tmp_path / ("c." + ext)), str(tmp_path / ("g." + gext))
    assert cv2.imwrite(fc, col) and cv2.imwrite(fg, gray)
    tool = _imgdump()
    for src, mode, flag in ((fc, "gray", cv2.IMREAD_GRAYSCALE), (fc, "color", cv2.IMREAD_COLOR),
                            (fg, "gray", cv2.IMREAD_GRAYSCALE), (fg, "color", cv2.IMREAD_COLOR)):
        out = str(tmp_path / "o.pnm")
        assert subprocess.run([tool, src, mode, out]).returncode == 0
        exp = cv2.imread(src, flag)
        with open(out, "rb") as f:
            assert f.readline() in (b"P5\n", b"P6\n")
            w, h = map(int, f.readline().split())
            f.readline()
            got = np.frombuffer(f.read(), np.uint8).reshape(exp.shape)
        assert (w, h) == (53, 37)
        assert np.array_equal(got, exp), (src, mode, int((got != exp).sum()))


def test_image_loader_rejects_malformed_files(tmp_path):
    """Truncated / hostile headers make load_image fail (exit code 1), never crash."""
    import struct
    import subprocess
    import zlib

    tool = _imgdump()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)

    sig = b"\x89PNG\r\n\x1a\n"
    bad = {
        "short_ihdr.png": sig + chunk(b"IHDR", b"\0\0\0\x04\0\0\0\x04\x08") + chunk(b"IEND", b""),
        "huge.png": sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 1 << 30, 1 << 30, 8, 2, 0, 0, 0)) + chunk(b"IEND", b""),
        "no_plte.png": sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 3, 0, 0, 0)) +
        chunk(b"IDAT", zlib.compress(b"\0\5\7\0\1\2")) + chunk(b"IEND", b""),
        "overflow.pgm": b"P5\n99999999999 99999999999\n255\n" + b"\0" * 16,
        "truncated.pgm": b"P5\n64 64\n255\n" + b"\0" * 100,
        "empty.png": b"",
    }
    for name, data in bad.items():
        p = tmp_path / name
        p.write_bytes(data)
        r = subprocess.run([tool, str(p), "gray", str(tmp_path / "o.pnm")])
        assert r.returncode == 1, (name, r.returncode)

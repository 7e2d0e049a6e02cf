"""The command-line drop-in (of_dis_b200/host/run_dense.cpp): argument grammar on CPU,
end-to-end .flo/.pfm against the Python pipeline (same preprocessing, oracle flow) on GPU."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from of_dis_b200 import build, params, preprocess, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bindir():
    return build.build_host()


def write_pnm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (w, h) if img.ndim == 2 else b"P6\n# c\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img).tobytes())


def write_png(path, img):
    """8-bit gray / RGB, filter type 'Paeth' on odd rows and 'Sub' on even rows to exercise the decoder."""
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else 3
    rows = img.reshape(h, w * ch).astype(np.int32)
    raw = bytearray()
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        cur = rows[y]
        left = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        if y % 2 == 0:
            raw.append(1)
            raw += bytes(((cur - left) & 255).astype(np.uint8))
        else:
            ul = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
            p = left + prev - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            raw.append(4)
            raw += bytes(((cur - pred) & 255).astype(np.uint8))
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if ch == 1 else 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b""))


def test_usage_and_argument_count(bindir):
    exe = os.path.join(bindir, "run_OF_INT")
    assert subprocess.run([exe], capture_output=True).returncode == 2
    # 7 numbers is neither an op-point nor the 20-parameter form (the reference reads argv blindly)
    r = subprocess.run([exe, "a", "b", "c"] + ["1"] * 7, capture_output=True)
    assert r.returncode == 2 and b"20" in r.stderr
    r = subprocess.run([exe, "/nonexistent/a.png", "/nonexistent/b.png", "/tmp/x.flo"], capture_output=True)
    assert r.returncode == 1


@pytest.mark.gpu
def test_host_classes_selftest(bindir):
    r = subprocess.run([os.path.join(bindir, "ofdis_host_selftest")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("exe,ch,nop,args,fmt", [
    ("run_OF_INT", 1, 2, ["2"], "png"),
    ("run_OF_INT", 1, 2, [], "pgm"),
    ("run_OF_RGB", 3, 2, "3 1 16 16 0.05 0.95 0 12 0.75 0 1 1 1 10 10 5 1 3 1.6 0".split(), "png"),
    ("run_DE_INT", 1, 1, "3 1 24 24 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split(), "pgm"),
    # README parameter 10 (usefbcon) = 1: forward-backward consistency
    ("run_OF_INT", 1, 2, "3 1 8 8 0.05 0.95 0 8 0.4 1 1 0 1 10 10 5 1 3 1.6 0".split(), "pgm"),
    ("run_DE_RGB", 3, 1, "3 1 8 8 0.05 0.95 0 8 0.4 1 1 0 1 10 10 5 1 3 1.6 0".split(), "png"),
])
def test_cli_output_equals_python_pipeline(tmp_path, bindir, oracle_port, exe, ch, nop, args, fmt):
    h, w = (150, 250) if len(args) > 1 else (218, 500)  # not divisible by 2^lv_f: exercises the padding/crop
    i0, i1, _ = synth.synthetic_pair(h, w, ch, seed=21, stereo=(nop == 1))
    fa, fb = str(tmp_path / ("a." + fmt)), str(tmp_path / ("b." + fmt))
    for path, img in ((fa, i0), (fb, i1)):
        rgb = img if ch == 1 else img[..., ::-1]  # files store RGB, the pipeline works in BGR like cv::imread
        (write_png if fmt == "png" else write_pnm)(path, np.ascontiguousarray(rgb))
    out = str(tmp_path / ("out.flo" if nop == 2 else "out.pfm"))
    r = subprocess.run([os.path.join(bindir, exe), fa, fb, out] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # same run with the pre/post-processing on the host (float pyramids handed to OFClass like the
    # reference's main): the two outputs must be the same file
    out_host = out + ".host"
    r2 = subprocess.run([os.path.join(bindir, exe), fa, fb, out_host] + args, capture_output=True, text=True,
                        env=dict(os.environ, OFDIS_HOST_PYRAMID="1"))
    assert r2.returncode == 0, r2.stdout + r2.stderr
    assert open(out, "rb").read() == open(out_host, "rb").read()
    if len(args) <= 1:
        assert "TIME (O.Flow Run-Time   ) (ms):" in r.stdout and "TIME (Sc:" in r.stdout  # verbosity 2 lines
        prm = params.operating_point(int(args[0]) if args else 2, w, noc=ch, nop=nop)
    else:
        prm = params.from_cli_numbers(args, noc=ch, nop=nop)
    pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
    exp = preprocess.postprocess(oracle_port.port_run(pyr, prm), prm.sc_l, pyr.padw, pyr.padh, w, h)
    if nop == 2:
        got = preprocess.read_flo(out)
    else:
        with open(out, "rb") as f:
            assert f.readline() == b"Pf\n"
            ww, hh = map(int, f.readline().split())
            assert float(f.readline()) == -1.0
            got = -np.fromfile(f, "<f4").reshape(hh, ww)[::-1].reshape(hh, ww, 1)
    assert got.shape == exp.shape
    assert np.array_equal(np.ascontiguousarray(got).view(np.uint32), np.ascontiguousarray(exp).view(np.uint32)), \
        float(np.abs(got - exp).max())


def test_batch_front_end_usage(bindir):
    exe = os.path.join(bindir, "run_OF_INT_batch")
    assert subprocess.run([exe], capture_output=True).returncode == 2
    r = subprocess.run([exe, "/nonexistent/list.txt"], capture_output=True)
    assert r.returncode == 1
    r = subprocess.run([exe, "/nonexistent/list.txt", "1", "2", "3"], capture_output=True)
    assert r.returncode == 2 and b"20" in r.stderr


@pytest.mark.gpu
def test_batch_front_end_writes_the_files_of_the_single_pair_binary(tmp_path, bindir):
    """run_OF_INT_batch (list file, pairs grouped by size, many pairs per launch) == run_OF_INT per pair."""
    sizes = [(218, 500), (218, 500), (218, 500), (150, 250), (150, 250)]
    lines, singles = [], []
    for k, (h, w) in enumerate(sizes):
        i0, i1, _ = synth.synthetic_pair(h, w, 1, seed=40 + k)
        fa, fb = str(tmp_path / ("a%d.pgm" % k)), str(tmp_path / ("b%d.pgm" % k))
        write_pnm(fa, i0)
        write_pnm(fb, i1)
        lines.append("%s %s %s" % (fa, fb, tmp_path / ("batch%d.flo" % k)))
        singles.append((fa, fb, str(tmp_path / ("single%d.flo" % k))))
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(lines) + "\n")
    r = subprocess.run([os.path.join(bindir, "run_OF_INT_batch"), str(lst), "--batch", "2", "2"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "5 pairs" in r.stdout
    for k, (fa, fb, out) in enumerate(singles):
        r = subprocess.run([os.path.join(bindir, "run_OF_INT"), fa, fb, out, "2"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert open(out, "rb").read() == open(str(tmp_path / ("batch%d.flo" % k)), "rb").read(), k

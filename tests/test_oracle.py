"""CPU tests of the checker itself: the C restatement (oracle/dis_oracle.c) must
equal (a) the committed golden fixtures produced by the reference build and
(b), where oracle/_ref exists, the reference build itself -- bit for bit."""
import glob
import os

import numpy as np
import pytest

from of_dis_b200 import params, preprocess, synth
from oracle import ref_driver

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    prm = params.from_cli_numbers(z["cli"], noc=int(z["noc"]), nop=int(z["nop"]))
    pyr = preprocess.PairPyramids(z["img0"], z["img1"], prm.sc_f, prm.p_samp_s)
    return z, prm, pyr


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_port_matches_golden_bitwise(path, oracle_port):
    z, prm, pyr = _load(path)
    flow = oracle_port.port_run(pyr, prm)
    assert np.array_equal(bits(flow), bits(z["flow"]))
    lvl = oracle_port.port_level_patches(pyr, prm, prm.sc_l, z["flow_prev"])
    assert np.array_equal(bits(lvl["p"]), bits(z["p"]))
    assert np.array_equal(lvl["conv"], z["conv"]) and np.array_equal(lvl["cnt"], z["cnt"])
    assert np.array_equal(bits(lvl["dense"]), bits(z["dense"]))


@pytest.mark.skipif(not ref_driver.ref_available("m1c1"), reason="oracle/_ref not built")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_reference_build_reproduces_golden(path):
    z, prm, pyr = _load(path)
    if not ref_driver.ref_available(prm.flavour()):
        pytest.skip("flavour not built")
    assert np.array_equal(bits(ref_driver.ref_run(pyr, prm)), bits(z["flow"]))


@pytest.mark.skipif(not ref_driver.ref_available("m1c1"), reason="oracle/_ref not built")
def test_port_vs_reference_cfg1_and_stages(oracle_port):
    """BASELINE config 1 (640x480 gray, op-point 2) whole run, plus per-stage checks."""
    i0, i1, _ = synth.synthetic_pair(480, 640, 1, seed=3)
    prm = params.operating_point(2, 640)
    pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
    ref = ref_driver.ref_run(pyr, prm)
    assert np.array_equal(bits(ref), bits(oracle_port.port_run(pyr, prm)))
    # variational refinement alone, on an arbitrary smooth-ish flow
    lv = prm.sc_l
    h, w = pyr.level_shape(lv)
    rng = np.random.default_rng(0)
    fl = (rng.standard_normal((h, w, 2)) * 0.7).astype(np.float32)
    assert np.array_equal(bits(ref_driver.ref_level_varref(pyr, prm, lv, fl)),
                          bits(oracle_port.port_level_varref(pyr, prm, lv, fl)))
    st = oracle_port.varref_stages(pyr, prm, lv, fl)
    out = np.stack([st["uu"], st["vv"]], -1)
    assert np.array_equal(bits(out), bits(oracle_port.port_level_varref(pyr, prm, lv, fl)))


def test_packet_order_sum_is_the_documented_order(oracle_port):
    import ctypes

    rng = np.random.default_rng(5)
    for n in (1, 3, 4, 7, 8, 12, 36, 64, 100, 144, 432):
        v = (rng.standard_normal(n) * 100).astype(np.float32)
        got = np.float32(oracle_port.lib().dis_sum_packet_order(v.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n))
        n4, n8 = n // 4 * 4, n // 8 * 8
        if n4:
            a = v[0:4].copy()
            if n4 > 4:
                b = v[4:8].copy()
                for i in range(8, n8, 8):
                    a = a + v[i:i + 4]
                    b = b + v[i + 4:i + 8]
                a = a + b
                if n4 > n8:
                    a = a + v[n8:n8 + 4]
            r = np.float32(np.float32(a[0] + a[2]) + np.float32(a[1] + a[3]))
            for i in range(n4, n):
                r = np.float32(r + v[i])
        else:
            r = v[0]
            for i in range(1, n):
                r = np.float32(r + v[i])
        assert got == r


def test_properties_zero_flow_and_translation(oracle_port):
    """identical images -> zero flow; integer shift recovered."""
    i0, _, _ = synth.synthetic_pair(128, 192, 1, seed=11)
    prm = params.from_cli_numbers("3 1 12 12 0.05 0.95 0 8 0.4 0 1 0 1 10 10 5 1 3 1.6 0".split())
    pyr = preprocess.PairPyramids(i0, i0, prm.sc_f, prm.p_samp_s)
    assert np.abs(oracle_port.port_run(pyr, prm)).max() == 0.0
    big, _, _ = synth.synthetic_pair(128, 192 + 8, 1, seed=12)
    a, b = np.ascontiguousarray(big[:, 4:-4]), np.ascontiguousarray(big[:, 2:-6])  # I1(x+2) = I0(x)
    pyr = preprocess.PairPyramids(a, b, prm.sc_f, prm.p_samp_s)
    fl = preprocess.postprocess(oracle_port.port_run(pyr, prm), prm.sc_l, pyr.padw, pyr.padh, 192, 128)
    inner = fl[16:-16, 16:-16]
    assert abs(np.median(inner[..., 0]) - 2.0) < 0.1 and abs(np.median(inner[..., 1])) < 0.1


# ---- the port against the reference build on everything the GPU tests use it for ----------------
@pytest.mark.skipif(not ref_driver.ref_available("m1c1"), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(40))
def test_port_vs_reference_on_the_random_configurations_of_the_gpu_suite(seed, oracle_port):
    """The 40 seeded parameter sets of tests/test_gpu_parity.py::test_random_configurations_vs_oracle:
    the port (the GPU tests' checker) must equal the reference build on each of them."""
    from test_gpu_parity import _random_config

    rng = np.random.default_rng(1000 + seed)
    numbers, ch, nop, size, amp = _random_config(rng)
    prm = params.from_cli_numbers(numbers, noc=ch, nop=nop)
    if not ref_driver.ref_available(prm.flavour()):
        pytest.skip("flavour not built")
    i0, i1, _ = synth.synthetic_pair(size[0], size[1], ch, seed=200 + seed, stereo=(nop == 1), amp=amp)
    pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
    assert np.array_equal(bits(ref_driver.ref_run(pyr, prm)), bits(oracle_port.port_run(pyr, prm)))


BASELINE_CASES = {
    "cfg2_seed1": (436, 1024, 1, lambda: params.operating_point(2, 1024), 1, False),
    "cfg2_seed7": (436, 1024, 1, lambda: params.operating_point(2, 1024), 7, False),
    "cfg3_1920x1080_rgb_l1": (1080, 1920, 3, lambda: params.from_cli_numbers(
        "6 2 16 16 0.05 0.95 0 12 0.75 0 1 1 1 10 10 5 1 3 1.6 0".split(), noc=3), 2, False),
    "cfg5_2880x1988_stereo_op4": (1988, 2880, 1, lambda: params.operating_point(4, 2880, noc=1, nop=1), 4, True),
}


@pytest.mark.skipif(not ref_driver.ref_available("m1c1"), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", list(BASELINE_CASES))
def test_port_vs_reference_at_baseline_sizes(name, oracle_port):
    """BASELINE configs[1], [2] and [4] at full size (the inputs of the GPU suite's full-size tests)."""
    h, w, ch, mk, seed, stereo = BASELINE_CASES[name]
    prm = mk()
    if not ref_driver.ref_available(prm.flavour()):
        pytest.skip("flavour not built")
    i0, i1, _ = synth.synthetic_pair(h, w, ch, seed=seed, stereo=stereo, amp=6.0)
    pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
    assert np.array_equal(bits(ref_driver.ref_run(pyr, prm)), bits(oracle_port.port_run(pyr, prm)))


@pytest.mark.skipif(not ref_driver.ref_available("m1c1"), reason="oracle/_ref not built")
def test_native_thread_pool_drivers_reproduce_single_runs():
    """ofdis_ref_run_many / ofdis_ref_run_many_u8 (the CPU baseline's drivers): same flows as one
    ofdis_ref_run per pair, and the restated pyramid / upsampling equal of_dis_b200/preprocess.py
    (which tests/test_preprocess.py pins to cv2), bit for bit."""
    prm = params.from_cli_numbers("3 1 8 8 0.05 0.95 0 8 0.4 0 1 0 1 10 10 5 1 3 1.6 0".split())
    pairs = [synth.synthetic_pair(121, 203, 1, seed=30 + s)[:2] for s in range(3)]  # odd size: padding + crop
    pyrs = [preprocess.PairPyramids(a, b, prm.sc_f, prm.p_samp_s) for a, b in pairs]
    _, flows = ref_driver.ref_run_many(pyrs, prm, nrep=2, threads=3)
    frames = np.stack([np.stack([a, b]) for a, b in pairs])[..., None]
    _, full = ref_driver.ref_run_many_u8(frames, prm, nrep=1, threads=2)
    for q, p in enumerate(pyrs):
        one = ref_driver.ref_run(p, prm)
        assert np.array_equal(bits(flows[q]), bits(one))
        exp = preprocess.postprocess(one, prm.sc_l, p.padw, p.padh, 203, 121)
        assert np.array_equal(bits(full[q]), bits(exp.reshape(full[q].shape)))

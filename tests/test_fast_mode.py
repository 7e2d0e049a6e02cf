"""The opt-in "fast" refinement (ofdis_set_option "sor_fast": red-black instead of lexicographic SOR, SURVEY 8f
rank 4).  It is NOT bit-identical to the reference and never covered by the parity claim; these tests pin what it
is: (1) exactly a red-black SOR of the same linear system (numpy restatement, bitwise, on a level of several
tiles: the temporal-blocking halo must not show), (2) close to the exact mode and as accurate against the ground
truth of the synthetic pairs, (3) deterministic and off by default."""
import numpy as np
import pytest

from of_dis_b200 import params, preprocess, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from of_dis_b200 import api as _api

    _api.lib()
    return _api


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def redblack_numpy(it, K, omega):
    """K red-black sweeps from du = dv = 0 on the records of one inner iteration (flow)."""
    f32 = np.float32
    a11, a12, a22, b1, b2, sh, sv = (np.asarray(it[k], f32) for k in ("a11_inv", "a12_inv", "a22_inv", "b1", "b2", "sh", "sv"))
    h, w = b1.shape
    shl = np.zeros_like(sh); shl[:, 1:] = sh[:, :-1]
    svt = np.zeros_like(sv); svt[1:, :] = sv[:-1, :]
    yy, xx = np.mgrid[0:h, 0:w]
    du, dv = np.zeros((h, w), f32), np.zeros((h, w), f32)
    om = f32(omega)

    def nb(a):
        l = np.zeros_like(a); l[:, 1:] = a[:, :-1]
        r = np.zeros_like(a); r[:, :-1] = a[:, 1:]
        t = np.zeros_like(a); t[1:, :] = a[:-1, :]
        b = np.zeros_like(a); b[:-1, :] = a[1:, :]
        return l, r, t, b

    for s in range(2 * K):
        m = ((xx + yy) & 1) == (s & 1)
        ul, ur, ut, ub = nb(du)
        vl, vr, vt, vb = nb(dv)
        B1 = b1 + (((shl * ul + sh * ur) + svt * ut) + sv * ub)
        B2 = b2 + (((shl * vl + sh * vr) + svt * vt) + sv * vb)
        nu = du + om * (a11 * B1 + a12 * B2 - du)
        nv = dv + om * (a12 * B1 + a22 * B2 - dv)
        du, dv = np.where(m, nu, du), np.where(m, nv, dv)
    return du, dv


@pytest.mark.parametrize("size,numbers", [((436, 1024), None), ((200, 320), "3 1 8 8 0.05 0.95 0 6 0.5 0 0 0 1 10 10 5 2 5 1.5 0")])
def test_fast_mode_is_a_red_black_sor_of_the_same_system(size, numbers, api, oracle_port):
    prm = params.operating_point(2, size[1]) if numbers is None else params.from_cli_numbers(numbers.split())
    i0, i1, _ = synth.synthetic_pair(size[0], size[1], 1, seed=1)
    pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
    lv = prm.sc_l
    hh, ww = pyr.level_shape(lv)
    rng = np.random.default_rng(3)
    dense = (rng.standard_normal((hh, ww, 2)) * 1.5).astype(np.float32)
    st = oracle_port.varref_stages(pyr, prm, lv, dense, n_iters=1)
    exp_du, exp_dv = redblack_numpy(st["iters"][0], prm.tv_solverit, prm.tv_sor)
    ctx = api.Context(prm, pyr.width, pyr.height, pyr.imgpadding, 2)
    ctx.set_option("sor_fast", 1)
    for f in range(2):
        ctx.upload_pyramids(f, pyr)
        ctx.set_flow(f, lv, dense)
    ctx.varref_refine(lv, 0, 2, n_inner=1)
    rec = ctx.debug_get("rec", 1, lv)
    for idx, key in enumerate(("a11_inv", "a12_inv", "a22_inv", "b1", "b2", "sh", "sv")):
        assert np.array_equal(bits(rec[..., idx]), bits(st["iters"][0][key])), key  # the system itself is the reference's
    dudv = ctx.debug_get("dudv", 1, lv)
    assert np.array_equal(bits(dudv[..., 0]), bits(exp_du)), float(np.abs(dudv[..., 0] - exp_du).max())
    assert np.array_equal(bits(dudv[..., 1]), bits(exp_dv)), float(np.abs(dudv[..., 1] - exp_dv).max())
    ctx.close()


@pytest.mark.parametrize("nop,ch", [(2, 1), (1, 1), (2, 3)])
def test_fast_mode_stays_close_to_the_exact_mode(nop, ch, api):
    prm = params.operating_point(2, 1024, noc=ch, nop=nop)
    errs = {}
    flows = {}
    for fast in (0, 1):
        out = []
        for seed in (5, 6):
            i0, i1, gt = synth.synthetic_pair(436, 1024, ch, seed=seed, stereo=(nop == 1))
            pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
            ctx = api.Context(prm, pyr.width, pyr.height, pyr.imgpadding, 1)
            if fast:
                ctx.set_option("sor_fast", 1)
            ctx.upload_pyramids(0, pyr)
            ctx.set_graph_mode(True)
            ctx.run(1)
            ctx.run(1)  # replay must give the same result
            a = ctx.get_flow(0, prm.sc_l)
            ctx.close()
            full = preprocess.postprocess(a, prm.sc_l, pyr.padw, pyr.padh, 1024, 436)
            gtf = gt if nop == 2 else gt[..., :1]
            out.append((full, float(np.sqrt(((full - gtf) ** 2).sum(-1)).mean())))
        flows[fast] = [o[0] for o in out]
        errs[fast] = np.mean([o[1] for o in out])
    delta = np.mean([np.abs(a - b).mean() for a, b in zip(flows[0], flows[1])])
    assert 0 < delta < 0.05, delta            # different iterate, a few hundredths of a pixel (full resolution)
    assert errs[1] < errs[0] * 1.05 + 0.01, errs  # as accurate against the ground truth as the exact mode


def test_fast_mode_is_off_by_default_and_needs_the_refinement(api):
    prm = params.operating_point(1, 1024)  # operating point 1: no refinement
    ctx = api.Context(prm, 1024, 448, 8, 1)
    with pytest.raises(api.OfdisError):
        ctx.set_option("sor_fast", 1)
    ctx.close()

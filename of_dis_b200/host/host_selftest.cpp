// Self-test of the reference-shaped C++ classes: the level loop of OFClass
// (oflow.cpp:184-295) written out with PatGridClass + VarRefClass must give the
// same bits as OFC::OFClass.  Usage: ofdis_host_selftest  (exit code 0 = ok)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ofdis_host.h"

using namespace OFC;

static std::vector<float> make_level(int w, int h, int pad, int shift, bool zero_border) {
  std::vector<float> v((size_t)(w + 2 * pad) * (h + 2 * pad));
  for (int y = 0; y < h + 2 * pad; ++y)
    for (int x = 0; x < w + 2 * pad; ++x) {
      int xx = x - pad, yy = y - pad;
      const bool out = xx < 0 || yy < 0 || xx >= w || yy >= h;
      xx = xx < 0 ? 0 : (xx >= w ? w - 1 : xx);
      yy = yy < 0 ? 0 : (yy >= h ? h - 1 : yy);
      const float val = 128.f + 60.f * sinf(0.21f * (xx + shift) + 0.13f * yy) + 40.f * cosf(0.17f * yy - 0.05f * (xx + shift));
      v[(size_t)y * (w + 2 * pad) + x] = (zero_border && out) ? 0.f : val;
    }
  return v;
}

int main() {
  const int W = 256, H = 128, P = 8, sc_f = 3, sc_l = 1;
  std::vector<std::vector<float>> a(sc_f + 1), ax(sc_f + 1), ay(sc_f + 1), b(sc_f + 1), bx(sc_f + 1), by(sc_f + 1);
  std::vector<const float*> pa(sc_f + 1), pax(sc_f + 1), pay(sc_f + 1), pb(sc_f + 1), pbx(sc_f + 1), pby(sc_f + 1);
  for (int l = 0; l <= sc_f; ++l) {
    const int w = W >> l, h = H >> l;
    a[l] = make_level(w, h, P, 0, false);
    b[l] = make_level(w, h, P, 1, false);
    // any gradient field works for an API equivalence test; use a finite difference of the image
    ax[l] = make_level(w, h, P, 0, true);
    ay[l] = make_level(w, h, P, 0, true);
    for (size_t i = 0; i + 1 < ax[l].size(); ++i) ax[l][i] = ax[l][i] == 0.f ? 0.f : 0.5f * (a[l][i + 1] - a[l][i]);
    for (size_t i = 0; i + (w + 2 * P) < ay[l].size(); ++i) ay[l][i] = ay[l][i] == 0.f ? 0.f : 0.5f * (a[l][i + w + 2 * P] - a[l][i]);
    bx[l] = ax[l];
    by[l] = ay[l];
    pa[l] = a[l].data(); pax[l] = ax[l].data(); pay[l] = ay[l].data();
    pb[l] = b[l].data(); pbx[l] = bx[l].data(); pby[l] = by[l].data();
  }
  const int wl = W >> sc_l, hl = H >> sc_l;
  std::vector<float> ref((size_t)wl * hl * 2), got;
  OFClass ofc(pa.data(), pax.data(), pay.data(), pb.data(), pbx.data(), pby.data(), P, ref.data(), nullptr, W, H, sc_f,
              sc_l, 12, 12, 0.05f, 0.95f, 0.f, P, 0.4f, false, 0, 1, 1, true, 10.f, 10.f, 5.f, 1, 3, 1.6f, 0);
  optparam op;
  FillOptParam(op, 2, sc_f, sc_l, 12, 12, 0.05f, 0.95f, 0.f, P, 0.4f, false, 0, 1, 1, true, 10.f, 10.f, 5.f, 1, 3, 1.6f, 0);
  std::vector<float> prev;
  for (int sl = sc_f; sl >= sc_l; --sl) {
    camparam cpl, cpr;
    FillCamParam(cpl, op, W, H, sl, P, 0);
    FillCamParam(cpr, op, W, H, sl, P, 1);
    PatGridClass grid(&cpl, &cpr, &op);
    grid.InitializeGrid(pa[sl], pax[sl], pay[sl]);
    grid.SetTargetImage(pb[sl], pbx[sl], pby[sl]);
    if (sl < sc_f) grid.InitializeFromCoarserOF(prev.data());
    grid.Optimize();
    std::vector<float> cur((size_t)cpl.width * cpl.height * 2);
    grid.AggregateFlowDense(cur.data());
    const Vector2f d = grid.GetQuePatchDis(0), r = grid.GetRefPatchPos(0), q = grid.GetQuePatchPos(0);
    if (d[0] != r[0] - q[0] || d[1] != r[1] - q[1] || grid.GetNoPatches() != grid.GetNopw() * grid.GetNoph()) {
      printf("FAIL: patch accessors inconsistent\n");
      return 1;
    }
    VarRefClass vr(pa[sl], pax[sl], pay[sl], pb[sl], pbx[sl], pby[sl], &cpl, &cpr, &op, cur.data());
    prev = cur;
  }
  got = prev;
  if (got.size() != ref.size() || memcmp(got.data(), ref.data(), sizeof(float) * ref.size())) {
    printf("FAIL: per-level classes differ from OFClass\n");
    return 1;
  }
  // Forward-backward couple (oflow.cpp:162-170,191-215): two stand-alone grids joined by SetComplGrid
  // must give the flow OFClass computes with usefbcon = 1 (one level, no refinement).
  {
    const int sl = 2, w2 = W >> sl, h2 = H >> sl;
    std::vector<float> fb_ref((size_t)w2 * h2 * 2), fw((size_t)w2 * h2 * 2), bw((size_t)w2 * h2 * 2);
    OFClass fbc(pa.data(), pax.data(), pay.data(), pb.data(), pbx.data(), pby.data(), P, fb_ref.data(), nullptr, W, H, sl, sl,
                12, 12, 0.05f, 0.95f, 0.f, P, 0.4f, true, 0, 1, 1, false, 10.f, 10.f, 5.f, 1, 3, 1.6f, 0);
    optparam op2;
    FillOptParam(op2, 2, sl, sl, 12, 12, 0.05f, 0.95f, 0.f, P, 0.4f, true, 0, 1, 1, false, 10.f, 10.f, 5.f, 1, 3, 1.6f, 0);
    camparam cpl, cpr;
    FillCamParam(cpl, op2, W, H, sl, P, 0);
    FillCamParam(cpr, op2, W, H, sl, P, 1);
    PatGridClass gfw(&cpl, &cpr, &op2), gbw(&cpr, &cpl, &op2);
    gfw.SetComplGrid(&gbw);
    gbw.SetComplGrid(&gfw);
    gfw.InitializeGrid(pa[sl], pax[sl], pay[sl]);
    gfw.SetTargetImage(pb[sl], pbx[sl], pby[sl]);
    gbw.InitializeGrid(pb[sl], pbx[sl], pby[sl]);
    gbw.SetTargetImage(pa[sl], pax[sl], pay[sl]);
    gfw.Optimize();
    gbw.Optimize();
    gfw.AggregateFlowDense(fw.data());
    gbw.AggregateFlowDense(bw.data());
    if (memcmp(fw.data(), fb_ref.data(), sizeof(float) * fw.size())) {
      printf("FAIL: SetComplGrid couple differs from OFClass(usefbcon=1)\n");
      return 1;
    }
    double sb = 0;
    for (float v : bw) sb += fabs(v);
    if (!(sb > 0) || gbw.GetQuePatchDis(0)[0] != gbw.GetRefPatchPos(0)[0] - gbw.GetQuePatchPos(0)[0]) {
      printf("FAIL: backward grid of the couple is empty\n");
      return 1;
    }
  }
  double s = 0;
  for (float v : ref) s += fabs(v);
  printf("ok: %zu flow values bitwise equal (mean |flow| %.4f)\n", ref.size(), s / ref.size());
  return 0;
}

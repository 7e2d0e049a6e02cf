// oracle/_ref wrapper -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" entry points around the UNMODIFIED reference classes
// (compiled in place from /root/reference by oracle/Makefile against
// oracle/eigen_shim).  Used by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg only; the product never links this.
//
//   ofdis_ref_run            -> OFC::OFClass ctor            (oflow.cpp:32-363)
//   ofdis_ref_level_patches  -> PatGridClass one level       (patchgrid.cpp:98-397)
//   ofdis_ref_level_varref   -> VarRefClass one level        (refine_variational.cpp:25-116)
//
// One shared object is built per (SELECTMODE, SELECTCHANNEL) pair because the
// reference selects flow/stereo and gray/RGB at compile time (CMakeLists.txt:25-46).
#include <iostream>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <sys/time.h>

#include <Eigen/Core>

// test-harness access to per-patch state (p_iter, pweight, flags); the
// reference sources themselves are not touched.
#define private public
#include "patchgrid.h"
#include "refine_variational.h"
#undef private

using namespace OFC;

extern "C" {

struct ofdis_ref_params {
  int sc_f, sc_l, max_iter, min_iter;
  float dp_thresh, dr_thresh, res_thresh;
  int p_samp_s;
  float patove;
  int usefbcon, costfct, noc, patnorm, usetvref;
  float tv_alpha, tv_gamma, tv_delta;
  int tv_innerit, tv_solverit;
  float tv_sor;
  int verbosity;
};

int ofdis_ref_mode(void) { return SELECTMODE; }
int ofdis_ref_channels(void) { return SELECTCHANNEL; }

// Whole coarse-to-fine run == reference "O.Flow Run-Time" region.
void ofdis_ref_run(const float** i0, const float** i0x, const float** i0y,
                   const float** i1, const float** i1x, const float** i1y,
                   int imgpadding, float* outflow, const float* initflow,
                   int width, int height, const ofdis_ref_params* p) {
  OFClass ofc(i0, i0x, i0y, i1, i1x, i1y, imgpadding, outflow, initflow, width, height,
              p->sc_f, p->sc_l, p->max_iter, p->min_iter, p->dp_thresh, p->dr_thresh,
              p->res_thresh, p->p_samp_s, p->patove, p->usefbcon != 0, p->costfct, p->noc,
              p->patnorm, p->usetvref != 0, p->tv_alpha, p->tv_gamma, p->tv_delta,
              p->tv_innerit, p->tv_solverit, p->tv_sor, p->verbosity);
}

// optparam / camparam exactly as OFClass derives them (oflow.cpp:76-108, 138-157).
static void fill_params(optparam& op, camparam& cpl, camparam& cpr, int width_in, int height_in,
                        int sl, int imgpadding_in, const ofdis_ref_params* p) {
#if (SELECTMODE == 1)
  op.nop = 2;
#else
  op.nop = 1;
#endif
  op.p_samp_s = p->p_samp_s;
  op.outlierthresh = (float)op.p_samp_s / 2;
  op.patove = p->patove;
  op.sc_f = p->sc_f;
  op.sc_l = p->sc_l;
  op.max_iter = p->max_iter;
  op.min_iter = p->min_iter;
  op.dp_thresh = p->dp_thresh * p->dp_thresh;
  op.dr_thresh = p->dr_thresh;
  op.res_thresh = p->res_thresh;
  op.steps = std::max(1, (int)floor(op.p_samp_s * (1 - op.patove)));
  op.novals = p->noc * (p->p_samp_s) * (p->p_samp_s);
  op.usefbcon = p->usefbcon != 0;
  op.costfct = p->costfct;
  op.noc = p->noc;
  op.patnorm = p->patnorm;
  op.verbosity = p->verbosity;
  op.noscales = op.sc_f - op.sc_l + 1;
  op.usetvref = p->usetvref != 0;
  op.tv_alpha = p->tv_alpha;
  op.tv_gamma = p->tv_gamma;
  op.tv_delta = p->tv_delta;
  op.tv_innerit = p->tv_innerit;
  op.tv_solverit = p->tv_solverit;
  op.tv_sor = p->tv_sor;
  op.normoutlier_tmpbsq = (v4sf){op.normoutlier * op.normoutlier, op.normoutlier * op.normoutlier,
                                 op.normoutlier * op.normoutlier, op.normoutlier * op.normoutlier};
  op.normoutlier_tmp2bsq = __builtin_ia32_mulps(op.normoutlier_tmpbsq, op.twos);
  op.normoutlier_tmp4bsq = __builtin_ia32_mulps(op.normoutlier_tmpbsq, op.fours);

  float sc_fct = pow(2, -sl);
  cpl.sc_fct = sc_fct;
  cpl.height = height_in * sc_fct;
  cpl.width = width_in * sc_fct;
  cpl.imgpadding = imgpadding_in;
  cpl.tmp_lb = -(float)op.p_samp_s / 2;
  cpl.tmp_ubw = (float)(cpl.width + op.p_samp_s / 2 - 2);
  cpl.tmp_ubh = (float)(cpl.height + op.p_samp_s / 2 - 2);
  cpl.tmp_w = cpl.width + 2 * imgpadding_in;
  cpl.tmp_h = cpl.height + 2 * imgpadding_in;
  cpl.curr_lv = sl;
  cpl.camlr = 0;
  cpr = cpl;
  cpr.camlr = 1;
}

// One pyramid level of the patch stage: InitializeGrid, SetTargetImage,
// [InitializeFromCoarserOF], Optimize, AggregateFlowDense (oflow.cpp:191-267).
// Returns the number of patches; any output pointer may be null.
int ofdis_ref_level_patches(const float* i0, const float* i0x, const float* i0y, const float* i1,
                            const float* i1x, const float* i1y, int width_full, int height_full,
                            int level, int imgpadding, const ofdis_ref_params* p,
                            const float* flow_prev, float* p_out, float* pweight_out,
                            int* conv_out, int* cnt_out, float* dense_out) {
  optparam op;
  camparam cpl, cpr;
  fill_params(op, cpl, cpr, width_full, height_full, level, imgpadding, p);
  PatGridClass grid(&cpl, &cpr, &op);
  grid.InitializeGrid(i0, i0x, i0y);
  grid.SetTargetImage(i1, i1x, i1y);
  if (flow_prev) grid.InitializeFromCoarserOF(flow_prev);
  grid.Optimize();
  const int np = grid.GetNoPatches();
  for (int i = 0; i < np; ++i) {
    const patchstate* pc = grid.pat[i]->pc;
    if (p_out)
      for (int k = 0; k < op.nop; ++k) p_out[i * op.nop + k] = pc->p_iter[k];
    if (pweight_out) std::memcpy(pweight_out + (size_t)i * op.novals, pc->pweight.data(), sizeof(float) * op.novals);
    if (conv_out) conv_out[i] = pc->hasconverged ? 1 : 0;
    if (cnt_out) cnt_out[i] = pc->cnt;
  }
  if (dense_out) grid.AggregateFlowDense(dense_out);
  return np;
}

// One pyramid level of the variational refinement, in place on flow.
void ofdis_ref_level_varref(const float* i0, const float* i0x, const float* i0y, const float* i1,
                            const float* i1x, const float* i1y, int width_full, int height_full,
                            int level, int imgpadding, const ofdis_ref_params* p, float* flow) {
  optparam op;
  camparam cpl, cpr;
  fill_params(op, cpl, cpr, width_full, height_full, level, imgpadding, p);
  VarRefClass varref(i0, i0x, i0y, i1, i1x, i1y, &cpl, &cpr, &op, flow);
}

// Wall-clock helper for the CPU baseline: runs ofdis_ref_run `reps` times and
// returns the per-run times in ms (timer placement == oflow.cpp:113-114,355-360).
void ofdis_ref_time_run(const float** i0, const float** i0x, const float** i0y, const float** i1,
                        const float** i1x, const float** i1y, int imgpadding, float* outflow,
                        int width, int height, const ofdis_ref_params* p, int reps, double* ms_out) {
  for (int r = 0; r < reps; ++r) {
    struct timeval a, b;
    gettimeofday(&a, nullptr);
    ofdis_ref_run(i0, i0x, i0y, i1, i1x, i1y, imgpadding, outflow, nullptr, width, height, p);
    gettimeofday(&b, nullptr);
    ms_out[r] = (b.tv_sec - a.tv_sec) * 1000.0 + (b.tv_usec - a.tv_usec) / 1000.0;
  }
}

}  // extern "C"

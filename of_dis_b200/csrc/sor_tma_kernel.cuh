// sor_tma_kernel -- the lexicographic-SOR wavefront (see varref_kernels.cu) with a TMA
// producer warp.  Included inside namespace ofdis::{anonymous} by varref_kernels.cu.
//
// Measured on the LDG version (tools/sor_timing.py): the sweep-0 warps, which issue the most
// variable-latency global loads, took ~3x longer per super-step than the others (twelve loads in
// flight share the six scoreboard slots with the shared-memory traffic, so the recurrence keeps
// waiting on slots that also guard global loads), and every warp waits for them at the barrier.
// Here no compute warp reads global memory:
//   * the LAST warp is a producer: each super-step one elected lane arms an mbarrier and issues
//     two bulk copies (cp.async.bulk -> UBLKCP) -- the whole record diagonal n (NQ*hpad float4,
//     contiguous thanks to the skewed layout) and the (du,dv) diagonal n+1 -- into a ring of
//     shared-memory stages, PF super-steps ahead of sweep 0;
//   * a diagonal stays resident while sweeps 0..K-1 consume it (super-steps n .. n+2(K-1)), so
//     the records are fetched from L2 once instead of K times;
//   * the producer also observes completion (mbarrier wait one super-step ahead of use) so the
//     compute warps just read their own 16-byte columns with conflict-free LDS.128 after the
//     CTA barrier.
// Stage reuse needs no "empty" barriers: the per-super-step __syncthreads orders the consumers'
// last read of a stage before the producer's next copy into it (plus a proxy fence).
#pragma once

__device__ __forceinline__ void mbar_init(unsigned a, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ float lds32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}


// ---------------------------------------------------------------------------
// One 4-pixel block of the lexicographic SOR, shared by both kernel variants.
// F: record fields of the block (flow: a11^-1 a12^-1 a22^-1 b1 b2 sh sv sv_top; stereo: A11 b1 sh
// sv sv_top), one float4 per field.  own_*: previous-sweep values of the block, rf_*: previous-sweep
// value of the first column of the next block, top_*: this sweep's values of the row above,
// bot_*: previous-sweep values of the row below.  du_l/dv_l/hl carry the left neighbour and its sh.
// The expressions are the reference's (solver.c:204-210 middle, :122-123 first, :259-260 last line;
// stereo :438-462); row-class and border cases select between both candidate values.
__device__ __forceinline__ float f4c(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

template <int NOP>
__device__ __forceinline__ void sor_block_update(const float4* F, const float4& own_u, const float4& own_v,
                                                 float rf_u, float rf_v, const float4& top_u, const float4& top_v,
                                                 const float4& bot_u, const float4& bot_v, bool first_row,
                                                 bool last_row, int col0, int w, float omega, float& du_l,
                                                 float& dv_l, float& hl, float* nu, float* nv) {
  const float ou[5] = {own_u.x, own_u.y, own_u.z, own_u.w, rf_u};
  const float ov[5] = {own_v.x, own_v.y, own_v.z, own_v.w, rf_v};
  if (NOP == 2) {
    // everything that does not depend on the left neighbour first (ILP) ...
    float s1[4], s2[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool has_r = (col0 + c + 1 < w);
      const float du_r = has_r ? ou[c + 1] : 0.0f, dv_r = has_r ? ov[c + 1] : 0.0f;
      const float b1 = f4c(F[3], c), b2 = f4c(F[4], c), hh = f4c(F[5], c), vv = f4c(F[6], c), vt = f4c(F[7], c);
      const float t1u = hh * du_r, t1v = hh * dv_r;
      const float t2u = t1u + vt * f4c(top_u, c), t2v = t1v + vt * f4c(top_v, c);
      const float bsu = first_row ? t1u : t2u, bsv = first_row ? t1v : t2v;
      const float t3u = bsu + vv * f4c(bot_u, c), t3v = bsv + vv * f4c(bot_v, c);
      s1[c] = (last_row ? bsu : t3u) + b1;
      s2[c] = (last_row ? bsv : t3v) + b2;
    }
    // ... then the sequential recurrence along the row
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float a11 = f4c(F[0], c), a12 = f4c(F[1], c), a22 = f4c(F[2], c);
      const float B1w = hl * du_l + s1[c], B2w = hl * dv_l + s2[c];
      const bool has_l = (col0 + c > 0);
      const float B1 = has_l ? B1w : s1[c], B2 = has_l ? B2w : s2[c];
      du_l = ou[c] + omega * (a11 * B1 + a12 * B2 - ou[c]);
      dv_l = ov[c] + omega * (a12 * B1 + a22 * B2 - ov[c]);
      hl = f4c(F[5], c);
      nu[c] = du_l;
      nv[c] = dv_l;
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = col0 + c;
      const float du_r = ou[c + 1];
      const float A11 = f4c(F[0], c), b1 = f4c(F[1], c), hh = f4c(F[2], c), vv = f4c(F[3], c), vt = f4c(F[4], c);
      float sg = 0.0f;  // sigma accumulates top, left, bottom, right
      const float s_t = sg - vt * f4c(top_u, c);
      sg = first_row ? sg : s_t;
      const float s_l = sg - hl * du_l;
      sg = (col > 0) ? s_l : sg;
      const float s_b = sg - vv * f4c(bot_u, c);
      sg = last_row ? sg : s_b;
      const float s_r = sg - hh * du_r;
      sg = (col < w - 1) ? s_r : sg;
      const float B1 = b1 - sg;
      du_l = (1.0f - omega) * ou[c] + omega * (B1 / A11);
      hl = hh;
      nu[c] = du_l;
      nv[c] = 0.f;
    }
  }
}

constexpr int SOR_TMA_PF = 3;  // producer lead (super-steps)
// ring depth: diagonal n is read by sweep k at super-step n+2k, and its (du,dv) part by sweep 0 at
// super-step n+1; it may be overwritten PF super-steps before its successor is first needed
__host__ __device__ inline int sor_tma_stages(int K) { return 2 * K + SOR_TMA_PF; }

// HPAD (rows padded to 32/64/128/256) is a template parameter so that every shared-memory
// address is `base + immediate`; stage indices advance incrementally (no modulo in the loop).
template <int NOP, int HPAD>
__global__ void __launch_bounds__(HPAD <= 64 ? 288 : 448, 1)  // HPAD 128: up to 3 sweeps x 128 rows + producer warp
    sor_tma_kernel(LevelGeom g, VarRefPlanes pl, VarRefParams vp, int K) {
  constexpr int hpad = HPAD;
  extern __shared__ __align__(128) float4 s_dyn[];
  constexpr int NF = (NOP == 2) ? 2 : 1;  // board entry: du x4, (dv x4)
  constexpr int NQ = (NOP == 2) ? 8 : 5;  // record fields (float4) per block
  constexpr int PF = SOR_TMA_PF;
  const int NR = sor_tma_stages(K);
  const int fr = blockIdx.x;
  const int w = g.w, h = g.h;
  const int tid = threadIdx.x;
  const int hb = h + 2;
  const int W4 = (w + 3) >> 2;
  const int S = W4 + h + 2 * K - 2;
  const int dmax = W4 + h - 1;
  // shared memory: [NR stages x (NQ+2)*hpad float4][board 2*K*hb*NF float4][NR mbarriers]
  const unsigned stage_bytes = (unsigned)(NQ + 2) * hpad * 16u;
  const unsigned rec_bytes = (unsigned)NQ * hpad * 16u, dud_bytes = 2u * hpad * 16u;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s_dyn);
  const unsigned board = sbase + (unsigned)NR * stage_bytes;
  const unsigned bufbytes = (unsigned)(K * hb * NF) * 16u;
  const unsigned mbar0 = board + 2u * bufbytes;
  const float4* const rec_g = pl.rec + (size_t)fr * pl.rec_stride;
  float4* const dud_g = pl.dudv + (size_t)fr * pl.dudv_stride;

  if (tid == 0) {
    for (int i = 0; i < NR; ++i) mbar_init(mbar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- producer warp ------------------------------------------------------------------------
  if (tid >= K * hpad) {
    const bool lead = (tid == K * hpad);
    unsigned ist = 0;  // stage of the next load to issue
    auto issue = [&](int n) {  // load n -> stage n % NR: records of diagonal n, (du,dv) of n+1
      const unsigned dst = sbase + ist * stage_bytes, mb = mbar0 + 8u * ist;
      const int d = n > dmax ? dmax : n, d1 = n + 1 > dmax ? dmax : n + 1;
      mbar_expect_tx(mb, rec_bytes + dud_bytes);
      bulk_g2s(dst, rec_g + (size_t)d * NQ * hpad, rec_bytes, mb);
      bulk_g2s(dst + rec_bytes, dud_g + (size_t)d1 * 2 * hpad, dud_bytes, mb);
      ist = (ist + 1 == (unsigned)NR) ? 0u : ist + 1;
    };
    if (lead)
      for (int n = 0; n < PF && n < S; ++n) issue(n);
    // Completion is observed by the producer, not by the consumers: before the barrier that ends
    // super-step T-1 the producer waits until load T has landed (it was issued PF-1 super-steps
    // earlier), so after that barrier every compute warp may read loads <= T without touching an
    // mbarrier (a try_wait on a completed phase still cost ~260 cycles per warp and super-step).
    mbar_wait(mbar0, 0);  // load 0, needed by sweep 0 in super-step 0
    __syncthreads();
    unsigned wst = 1, wpar = 0;  // stage / phase parity of load T+1
    if (NR == 1) { wst = 0; wpar = 1; }
    for (int T = 0; T < S; ++T) {
      if (lead && T + PF < S) {
        // the consumers' reads of this stage (generic proxy) were ordered by the barrier that
        // ended super-step T-1; order them before the async-proxy write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(T + PF);
      }
      if (T + 1 < S) mbar_wait(mbar0 + 8u * wst, wpar);
      if (++wst == (unsigned)NR) { wst = 0; wpar ^= 1u; }
      __syncthreads();
    }
    return;
  }

  // ---- compute warps ---------------------------------------------------------------------------
  const int k = tid / hpad, jraw = tid - k * hpad;
  const bool valid = jraw < h;
  const int j = valid ? jraw : h - 1;  // idle lanes shadow the last row, never store
  const unsigned a_me = board + (unsigned)((k * hb + j + 1) * NF) * 16u;
  const unsigned a_top = board + (unsigned)((k * hb + j) * NF) * 16u;
  const int km = k > 0 ? k - 1 : 0;
  const unsigned a_right = board + (unsigned)((km * hb + j + 1) * NF) * 16u;
  const unsigned a_bot = board + (unsigned)((km * hb + j + 2) * NF) * 16u;
  const bool first_row = (j == 0), last_row = (j == h - 1);
  const bool k0 = (k == 0), klast = (k == K - 1);
  const float omega = vp.omega;
  const unsigned lane_off = (unsigned)j * 16u, rowb = (unsigned)hpad * 16u;
  const unsigned jb_off = (unsigned)((j + 1 < hpad) ? j + 1 : j) * 16u;  // row below, same diagonal row
  const int jw_lo = jraw & ~31, jw_hi = (jw_lo + 31 < h - 1) ? jw_lo + 31 : h - 1;  // rows of this warp
  const int tstart = j + 2 * k;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float du_l = 0.f, dv_l = 0.f, hl = 0.f;
  float4 own_u = z4, own_v = z4;  // sweeps > 0: previous-sweep values of the current block
  unsigned prevb = bufbytes, curb = 0;
  unsigned st = 0, stp = 0;  // stages of load max(n,0) and of load n-1
  int I = -tstart;
  __syncthreads();  // load 0 has landed (producer waited for it)
#pragma unroll 1
  for (int T = 0; T < S; ++T, ++I) {
    const bool in_range = valid & (I >= 0) & (I < W4);
    SOR_STAMP(0, omega, omega);
    const int n = T - 2 * k;  // load number == diagonal of this warp's blocks
    SOR_STAMP(1, omega, omega);
    // Warp-uniform: does any row of this warp hold a block this super-step, or start one in the
    // next (that lane must fetch its previous-sweep block now)?  Rows jw_lo..jw_hi, block I = n - j,
    // wanted -1 <= I < W4.  Idle warps (the ramp-up and ramp-down of the wavefront, 28 % of the
    // warp-steps on a 128x54 level) only keep the ring and board indices moving.  Warps made of
    // shadow lanes only (rows >= h; possible when hpad = 128) never run the body: joining late they
    // would carry a wrong left-neighbour state into the board slot shared with the real last row.
    if (jw_lo < h && jw_lo <= n + 1 && jw_hi > n - W4) {
    const unsigned sa = sbase + st * stage_bytes;
    float4 F[NQ];
#pragma unroll
    for (int f = 0; f < NQ; ++f) F[f] = lds128(sa + f * rowb + lane_off);
    float4 bot_u, bot_v = z4, nxt_u = z4, nxt_v = z4;
    float rf_u, rf_v = 0.f;
    if (k0) {
      // previous values: own = (du,dv) diagonal n, staged with load n-1; the row below and the
      // first column of the next block are on diagonal n+1, staged with load n
      const unsigned sn = sa + rec_bytes;
      if (n >= 1) {
        const unsigned sp = sbase + stp * stage_bytes + rec_bytes;
        own_u = lds128(sp + lane_off);
        if (NOP == 2) own_v = lds128(sp + rowb + lane_off);
      } else {  // diagonal 0 has no predecessor stage; its only block is (I=0, j=0)
        own_u = dud_g[j];
        if (NOP == 2) own_v = dud_g[hpad + j];
      }
      bot_u = lds128(sn + jb_off);
      rf_u = lds32(sn + lane_off);
      if (NOP == 2) {
        bot_v = lds128(sn + rowb + jb_off);
        rf_v = lds32(sn + rowb + lane_off);
      }
    } else {  // previous-sweep values come from the board (written one super-step ago)
      nxt_u = lds128(a_right + prevb);
      bot_u = lds128(a_bot + prevb);
      if (NOP == 2) {
        nxt_v = lds128(a_right + prevb + 16);
        bot_v = lds128(a_bot + prevb + 16);
      }
      rf_u = nxt_u.x;
      rf_v = nxt_v.x;
    }
    const float4 top_u = lds128(a_top + prevb);
    const float4 top_v = (NOP == 2) ? lds128(a_top + prevb + 16) : z4;
    SOR_STAMP(2, top_u.w, F[NQ - 1].x);
    float nu[4], nv[4];
    const int col0 = 4 * I;
    sor_block_update<NOP>(F, own_u, own_v, rf_u, rf_v, top_u, top_v, bot_u, bot_v, first_row, last_row, col0, w,
                          omega, du_l, dv_l, hl, nu, nv);
    SOR_STAMP(4, nu[3], nv[3]);
    sts128(a_me + curb, make_float4(nu[0], nu[1], nu[2], nu[3]));
    if (NOP == 2) sts128(a_me + curb + 16, make_float4(nv[0], nv[1], nv[2], nv[3]));
    if (klast && in_range) {  // coalesced: lanes of a warp share the diagonal
      float4* dst = dud_g + (size_t)(I + j) * 2 * hpad + j;
      dst[0] = make_float4(nu[0], nu[1], nu[2], nu[3]);
      if (NOP == 2) dst[hpad] = make_float4(nv[0], nv[1], nv[2], nv[3]);
    }
    if (!k0) {  // the next block of the previous sweep is this thread's block one super-step on
      own_u = nxt_u;
      own_v = nxt_v;
    }
    }
    SOR_STAMP(5, omega, omega);
    __syncthreads();
    SOR_STAMP(6, omega, omega);
    const unsigned tmp = prevb;
    prevb = curb;
    curb = tmp;
    if (n >= 0) {  // advance to the stage of the next diagonal
      stp = st;
      st = (st + 1 == (unsigned)NR) ? 0u : st + 1;
    }
  }
}

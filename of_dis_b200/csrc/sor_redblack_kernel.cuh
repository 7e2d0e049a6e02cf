// sor_redblack_kernel -- the OPT-IN "fast" solver of the variational refinement (ofdis_set_option
// "sor_fast").  Included inside namespace ofdis::{anonymous} by varref_kernels.cu.
//
// NOT the reference's arithmetic: sor_coupled (solver.c:77-421) sweeps the pixels in raster order, this
// kernel in red-black (checkerboard) order -- same linear system, same omega, same number of sweeps, but a
// different iterate after K sweeps.  The flow differs from the reference build's by a few hundredths of a
// pixel on average (measured and reported by bench.py / tests/test_fast_mode.py; SURVEY finding 1: up to
// 0.17 px), so this mode is never covered by the 1e-3 / bitwise claim.  It exists because a lexicographic
// sweep is a W/4 + h deep dependency chain while a red-black half-sweep is embarrassingly
// parallel: the only route to a bandwidth-bound solver.
//
// One CTA per 32x32 tile and frame, all K sweeps in one launch by temporal blocking: the tile is staged in
// shared memory with a halo of 2K pixels (records: 7 planes, du, dv), 2K half-sweeps run on the staged
// region -- a pixel at depth d from the region's rim is exact for the first d half-sweeps, so the interior
// (depth >= 2K) is exact after all of them -- and only the interior is written.  Neighbouring tiles
// recompute each other's halo from the same inputs in the same order: results do not depend on the tiling.
// (du,dv) are ping-pong buffers (other tiles still read the old values of this tile's interior).
#pragma once

constexpr int RB_TILE = 32;

__host__ __device__ inline size_t rb_smem_bytes(int nop, int K) {
  const int s = RB_TILE + 4 * K;
  return (size_t)s * s * ((nop == 2 ? 7 : 4) + nop) * sizeof(float);
}

template <int NOP>
__global__ void __launch_bounds__(256) sor_redblack_kernel(LevelGeom g, VarRefPlanes pl, VarRefParams vp) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ float rb_smem[];
  constexpr int NR = (NOP == 2) ? 7 : 4;  // staged record planes
  const int K = vp.n_solver, H = 2 * K, S = RB_TILE + 2 * H, SS = S * S;
  const int w = g.w, h = g.h, pitch = g.pitch;
  const int fr = blockIdx.z;
  const int x0 = blockIdx.x * RB_TILE - H, y0 = blockIdx.y * RB_TILE - H;  // image position of the staged region
  const int tid = threadIdx.x;
  float* const s_rec = rb_smem;             // [NR][S][S]
  float* const s_u = rb_smem + NR * SS;     // [S][S]
  float* const s_v = s_u + SS;              // flow only
  const float* const rec = pl.frec + (size_t)fr * pl.frec_stride;
  const float* const du_in = pl.fdu + (size_t)fr * pl.fdu_stride + (size_t)pl.fcur * 2 * pl.plane;
  float* const du_out = pl.fdu + (size_t)fr * pl.fdu_stride + (size_t)(pl.fcur ^ 1) * 2 * pl.plane;

  // stage: pixels outside the image get zero weights and values, so they never contribute
  for (int idx = tid; idx < SS; idx += 256) {
    const int ly = idx / S, lx = idx - ly * S;
    const int gx = x0 + lx, gy = y0 + ly;
    float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float u = 0.f, v = 0.f;
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      const int o = gy * pitch + gx;
      const float4 a = *reinterpret_cast<const float4*>(rec + (size_t)o * 8);
      const float4 b = *reinterpret_cast<const float4*>(rec + (size_t)o * 8 + 4);
      r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
      u = du_in[o];
      if (NOP == 2) v = du_in[pl.plane + o];
    } else if (NOP == 1) {
      r[0] = 1.0f;  // stereo divides by A11
    }
#pragma unroll
    for (int f = 0; f < NR; ++f) s_rec[f * SS + idx] = r[f];
    s_u[idx] = u;
    if (NOP == 2) s_v[idx] = v;
  }
  __syncthreads();

  // 2K half-sweeps: colour 0 = pixels with even x+y (image coordinates), then colour 1, K times
  const float omega = vp.omega;
  const int half = (S - 2 + 1) / 2;  // pixels of one colour in a row of the region without its rim
  for (int s = 0; s < 2 * K; ++s) {
    const int colour = s & 1;
    for (int idx = tid; idx < (S - 2) * half; idx += 256) {
      const int ry = idx / half, rx = idx - ry * half;
      const int ly = 1 + ry;
      const int lx = 1 + 2 * rx + ((colour + y0 + ly + x0 + 1) & 1);
      const int gx = x0 + lx, gy = y0 + ly;
      if (lx > S - 2 || gx < 0 || gx >= w || gy < 0 || gy >= h) continue;
      const int p = ly * S + lx;
      if (NOP == 2) {
        // records: 0 a11^-1, 1 a12^-1, 2 a22^-1, 3 b1, 4 b2, 5 sh (to the right), 6 sv (downwards)
        const float shl = s_rec[5 * SS + p - 1], shr = s_rec[5 * SS + p], svt = s_rec[6 * SS + p - S], svb = s_rec[6 * SS + p];
        const float B1 = s_rec[3 * SS + p] + (((shl * s_u[p - 1] + shr * s_u[p + 1]) + svt * s_u[p - S]) + svb * s_u[p + S]);
        const float B2 = s_rec[4 * SS + p] + (((shl * s_v[p - 1] + shr * s_v[p + 1]) + svt * s_v[p - S]) + svb * s_v[p + S]);
        const float a11 = s_rec[p], a12 = s_rec[SS + p], a22 = s_rec[2 * SS + p];
        const float ou = s_u[p], ov = s_v[p];
        s_u[p] = ou + omega * (a11 * B1 + a12 * B2 - ou);
        s_v[p] = ov + omega * (a12 * B1 + a22 * B2 - ov);
      } else {
        // records: 0 A11 (incl. the smoothness weights), 1 b1, 2 sh, 3 sv
        const float shl = s_rec[2 * SS + p - 1], shr = s_rec[2 * SS + p], svt = s_rec[3 * SS + p - S], svb = s_rec[3 * SS + p];
        const float B1 = s_rec[SS + p] + (((shl * s_u[p - 1] + shr * s_u[p + 1]) + svt * s_u[p - S]) + svb * s_u[p + S]);
        s_u[p] = (1.0f - omega) * s_u[p] + omega * (B1 / s_rec[p]);
      }
    }
    __syncthreads();
  }

  // the tile's interior
  for (int idx = tid; idx < RB_TILE * RB_TILE; idx += 256) {
    const int ty = idx / RB_TILE, tx = idx - ty * RB_TILE;
    const int gx = x0 + H + tx, gy = y0 + H + ty;
    if (gx >= w || gy >= h) continue;
    const int p = (H + ty) * S + H + tx, o = gy * pitch + gx;
    du_out[o] = s_u[p];
    if (NOP == 2) du_out[pl.plane + o] = s_v[p];
  }
}

// Variational refinement of the DIS hot path on sm_100a (K5..K12 of SURVEY.md):
// VarRefClass (refine_variational.cpp:25-336) and the FDF1.0.1 routines it calls
// (opticalflow_aux.c:17-548, image.c:376-502, solver.c:77-466).
//
//   warp_kernel      image_warp + copyimage + the 0.5*(I0+I1w), I1w-I0 pass of get_derivatives
//   deriv1_kernel    Ix, Iy, Ixz, Iyz      (5-tap, horizontal replicate / vertical folded coeffs)
//   deriv2_kernel    Ixx, Ixy, Iyy
//   assemble_kernel  compute_smoothness + compute_data[_DE] + sub_laplacian (x2) + the 2x2 block
//                    inversion of sor_coupled's first sweep, fused; smoothness staged through
//                    shared-memory tiles; writes one 32-byte SOR record per pixel
//   sor_kernel       all sweeps of the lexicographic SOR as a systolic wavefront
//                    (one thread per (sweep,row), step t handles column t-row-2*sweep)
//
// Every expression keeps the reference's operand order; the TU is compiled with
// -fmad=false so nothing is contracted (bit-exactness).
#include "ofdis_internal.cuh"

namespace ofdis {

namespace {

#define DATANORM (0.1f * 0.1f)       /* opticalflow_aux.c:10 */
#define EPS_COLOR (0.001f * 0.001f)  /* :11 */
#define EPS_GRAD (0.001f * 0.001f)   /* :12 */
#define EPS_SMOOTH (0.001f * 0.001f) /* :14 */

// convolve_extract_coeffs(even=0) of {0,-8/12,1/12} and {0,-0.5} (image.c:338-342,
// refine_variational.cpp:45-48)
struct Coef5 { float c0, c1, c2, c3, c4; };
__device__ __forceinline__ Coef5 coef5() {
  Coef5 c;
  c.c0 = 1.0f / 12.0f;
  c.c1 = -8.0f / 12.0f;
  c.c2 = -0.0f;
  c.c3 = -(-8.0f / 12.0f);
  c.c4 = -(1.0f / 12.0f);
  return c;
}

// convolve_horiz_fast_5 (image.c:466-502): replicate borders, five products
__device__ __forceinline__ float conv_h5(const float* row, int w, int i, const Coef5& c) {
  return c.c0 * row[clampi(i - 2, w)] + c.c1 * row[clampi(i - 1, w)] + c.c2 * row[i] +
         c.c3 * row[clampi(i + 1, w)] + c.c4 * row[clampi(i + 2, w)];
}
// convolve_vert_fast_5 (image.c:401-434): border rows fold the coefficients
__device__ __forceinline__ float conv_v5(const float* q, int pitch, int h, int j, const Coef5& c) {
  if (j == 0) return (c.c0 + c.c1 + c.c2) * q[0] + c.c3 * q[pitch] + c.c4 * q[2 * pitch];
  if (j == 1) return (c.c0 + c.c1) * q[-pitch] + c.c2 * q[0] + c.c3 * q[pitch] + c.c4 * q[2 * pitch];
  if (j == h - 2) return c.c0 * q[-2 * pitch] + c.c1 * q[-pitch] + c.c2 * q[0] + (c.c3 + c.c4) * q[pitch];
  if (j == h - 1) return c.c0 * q[-2 * pitch] + c.c1 * q[-pitch] + (c.c2 + c.c3 + c.c4) * q[0];
  return c.c0 * q[-2 * pitch] + c.c1 * q[-pitch] + c.c2 * q[0] + c.c3 * q[pitch] + c.c4 * q[2 * pitch];
}

// ---------------------------------------------------------------------------
// image_warp (opticalflow_aux.c:17-60) on the padded interleaved I1, fused with
// the first loop of get_derivatives (opticalflow_aux.c:80-84).
template <int C, int NOP>
__global__ void __launch_bounds__(256) warp_kernel(LevelGeom g, VarRefPlanes pl, int f0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  const int fr = blockIdx.z, frame = f0 + fr;
  if (i >= g.w || j >= g.h) return;
  const float* fl = g.flow + (size_t)frame * g.flow_frame_stride + ((size_t)j * g.w + i) * NOP;
  const float wx = fl[0], wy = (NOP == 2) ? fl[1] : 0.0f;
  const float xx = i + wx, yy = j + wy;
  const int x = (int)floorf(xx), y = (int)floorf(yy);
  const float dx = xx - x, dy = yy - y;
  const int x1 = clampi(x, g.w), x2 = clampi(x + 1, g.w), y1 = clampi(y, g.h), y2 = clampi(y + 1, g.h);
  const int o = j * g.pitch + i;
  pl.mask[(size_t)fr * pl.plane + o] =
      (xx >= 0 && xx <= g.w - 1 && yy >= 0 && yy <= g.h - 1) ? 1.0f : 0.0f;
  const float* i1 = g.img[3] + (size_t)frame * g.img_frame_stride;
  const float* i0 = g.img[0] + (size_t)frame * g.img_frame_stride;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float s11 = i1[((y1 + g.pad) * g.tmp_w + x1 + g.pad) * C + c];
    const float s12 = i1[((y1 + g.pad) * g.tmp_w + x2 + g.pad) * C + c];
    const float s21 = i1[((y2 + g.pad) * g.tmp_w + x1 + g.pad) * C + c];
    const float s22 = i1[((y2 + g.pad) * g.tmp_w + x2 + g.pad) * C + c];
    const float wv = s11 * (1.0f - dx) * (1.0f - dy) + s12 * dx * (1.0f - dy) + s21 * (1.0f - dx) * dy +
                     s22 * dx * dy;
    const float im1 = i0[((j + g.pad) * g.tmp_w + i + g.pad) * C + c];
    const size_t po = ((size_t)fr * C + c) * pl.plane + o;
    pl.avg[po] = 0.5f * (wv + im1);
    pl.deriv[2][po] = wv - im1;  // Iz
  }
}

// get_derivatives, first-order planes (opticalflow_aux.c:86-87,91-92)
template <int C>
__global__ void __launch_bounds__(256) deriv1_kernel(LevelGeom g, VarRefPlanes pl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.w || j >= g.h) return;
  const Coef5 c5 = coef5();
  const size_t base = (size_t)blockIdx.z * pl.plane;  // blockIdx.z = frame*C + c
  const int o = j * g.pitch + i;
  const float* avg = pl.avg + base;
  const float* iz = pl.deriv[2] + base;
  pl.deriv[0][base + o] = conv_h5(avg + j * g.pitch, g.w, i, c5);
  pl.deriv[1][base + o] = conv_v5(avg + o, g.pitch, g.h, j, c5);
  pl.deriv[6][base + o] = conv_h5(iz + j * g.pitch, g.w, i, c5);
  pl.deriv[7][base + o] = conv_v5(iz + o, g.pitch, g.h, j, c5);
}

// get_derivatives, second-order planes (opticalflow_aux.c:88-90)
template <int C>
__global__ void __launch_bounds__(256) deriv2_kernel(LevelGeom g, VarRefPlanes pl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.w || j >= g.h) return;
  const Coef5 c5 = coef5();
  const size_t base = (size_t)blockIdx.z * pl.plane;
  const int o = j * g.pitch + i;
  const float* ix = pl.deriv[0] + base;
  const float* iy = pl.deriv[1] + base;
  pl.deriv[3][base + o] = conv_h5(ix + j * g.pitch, g.w, i, c5);
  pl.deriv[4][base + o] = conv_v5(ix + o, g.pitch, g.h, j, c5);
  pl.deriv[5][base + o] = conv_v5(iy + o, g.pitch, g.h, j, c5);
}

// ---------------------------------------------------------------------------
// One inner fixed-point iteration, everything except the solver:
// compute_smoothness (opticalflow_aux.c:123-165), compute_data / compute_data_DE
// (:309-548), sub_laplacian on b1 (and b2) (:172-199), and for flow the in-place
// 2x2 inversion of sor_coupled's first sweep (solver.c:115-120).
constexpr int TX = 32, TY = 8;

template <int C, int NOP>
__global__ void __launch_bounds__(TX * TY) assemble_kernel(LevelGeom g, VarRefPlanes pl, VarRefParams vp,
                                                             int f0, int first) {
  __shared__ float2 s_uv[TY + 4][TX + 4];
  __shared__ float s_s[TY + 2][TX + 2];
  const int fr = blockIdx.z, frame = f0 + fr;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int w = g.w, h = g.h, pitch = g.pitch;
  const float* flow = g.flow + (size_t)frame * g.flow_frame_stride;
  const float2* dudv = pl.dudv + (size_t)fr * pl.plane;

  // uu = wx + du (vv likewise); first iteration: uu = wx (refine_variational.cpp:189-190).
  // Coordinates are clamped, which also realises the replicate border of the 3-tap
  // horizontal derivative (image.c:448-454).
  for (int idx = tid; idx < (TY + 4) * (TX + 4); idx += TX * TY) {
    const int cy = idx / (TX + 4), cx = idx - cy * (TX + 4);
    const int gx = clampi(x0 - 2 + cx, w), gy = clampi(y0 - 2 + cy, h);
    const float* f = flow + ((size_t)gy * w + gx) * NOP;
    float2 uv;
    uv.x = f[0];
    uv.y = (NOP == 2) ? f[1] : 0.0f;
    if (!first) {
      const float2 d = dudv[gy * pitch + gx];
      if (NOP == 2) {
        uv.x = uv.x + d.x;
        uv.y = uv.y + d.y;
      } else {  // minps / maxps with zero (refine_variational.cpp:299-314)
        const float t = uv.x + d.x;
        uv.x = (g.camlr == 0) ? (t < 0.0f ? t : 0.0f) : (t > 0.0f ? t : 0.0f);
      }
    }
    s_uv[cy][cx] = uv;
  }
  __syncthreads();

  // smoothness weight s = quarter_alpha / sqrt(ux^2+uy^2+vx^2+vy^2+eps) on the tile + 1 halo
  {
    const float c0 = -0.5f, c1 = -0.0f, c2 = 0.5f;  // {0,-0.5} -> [-0.5,-0,0.5]
    for (int idx = tid; idx < (TY + 2) * (TX + 2); idx += TX * TY) {
      const int cy = idx / (TX + 2), cx = idx - cy * (TX + 2);
      const int gx = x0 - 1 + cx, gy = y0 - 1 + cy;
      if (gx < 0 || gx >= w || gy < 0 || gy >= h) continue;
      const int sx = cx + 1, sy = cy + 1;  // position in s_uv
      const float2 l = s_uv[sy][sx - 1], m = s_uv[sy][sx], r = s_uv[sy][sx + 1];
      const float2 t = s_uv[sy - 1][sx], b = s_uv[sy + 1][sx];
      const float ux = c0 * l.x + c1 * m.x + c2 * r.x;
      const float vx = c0 * l.y + c1 * m.y + c2 * r.y;
      float uy, vy;
      if (gy == 0) {  // convolve_vert_fast_3 (image.c:383-398)
        uy = (c0 + c1) * m.x + c2 * b.x;
        vy = (c0 + c1) * m.y + c2 * b.y;
      } else if (gy == h - 1) {
        uy = c0 * t.x + (c1 + c2) * m.x;
        vy = c0 * t.y + (c1 + c2) * m.y;
      } else {
        uy = c0 * t.x + c1 * m.x + c2 * b.x;
        vy = c0 * t.y + c1 * m.y + c2 * b.y;
      }
      s_s[cy][cx] = vp.quarter_alpha / sqrtf(ux * ux + uy * uy + vx * vx + vy * vy + EPS_SMOOTH);
    }
  }
  __syncthreads();

  const int i = x0 + threadIdx.x, j = y0 + threadIdx.y;
  if (i >= w || j >= h) return;
  const int cx = threadIdx.x + 1, cy = threadIdx.y + 1;
  const float sc = s_s[cy][cx];
  const float hh = (i < w - 1) ? sc + s_s[cy][cx + 1] : 0.0f;    // sh(i,j)   (opticalflow_aux.c:150-154)
  const float hl = (i > 0) ? s_s[cy][cx - 1] + sc : 0.0f;        // sh(i-1,j)
  const float vv = (j < h - 1) ? sc + s_s[cy + 1][cx] : 0.0f;    // sv(i,j)   (:159-163)
  const float vt = (j > 0) ? s_s[cy - 1][cx] + sc : 0.0f;        // sv(i,j-1)

  const int o = j * pitch + i;
  const float m = pl.mask[(size_t)fr * pl.plane + o];
  const float2 d = dudv[o];
  const float u = d.x, v = (NOP == 2) ? d.y : 0.0f;
  const float hdo3 = vp.half_delta_over3, hgo3 = vp.half_gamma_over3;
  float A11 = 0.f, A12 = 0.f, A22 = 0.f, B1 = 0.f, B2 = 0.f;
#define DRV(k, c) pl.deriv[k][((size_t)fr * C + (c)) * pl.plane + o]
  if (C == 1) {
    const float ix = DRV(0, 0), iy = DRV(1, 0), iz = DRV(2, 0), ixx = DRV(3, 0), ixy = DRV(4, 0),
                iyy = DRV(5, 0), ixz = DRV(6, 0), iyz = DRV(7, 0);
    float t, t2, nn, n2;
    if (hdo3 != 0.0f) {
      t = (NOP == 2) ? iz + ix * u + iy * v : iz + ix * u;
      nn = ix * ix + iy * iy + DATANORM;
      t = m * hdo3 / sqrtf(3 * t * t / nn + EPS_COLOR);
      t /= nn;
      A11 += t * ix * ix;
      B1 -= t * iz * ix;
      if (NOP == 2) {
        A12 += t * ix * iy;
        A22 += t * iy * iy;
        B2 -= t * iz * iy;
      }
    }
    nn = ixx * ixx + ixy * ixy + DATANORM;
    n2 = iyy * iyy + ixy * ixy + DATANORM;
    t = (NOP == 2) ? ixz + ixx * u + ixy * v : ixz + ixx * u;
    t2 = (NOP == 2) ? iyz + ixy * u + iyy * v : iyz + ixy * u;
    t = m * hgo3 / sqrtf(3 * t * t / nn + 3 * t2 * t2 / n2 + EPS_GRAD);
    t2 = t / n2;
    t /= nn;
    A11 += t * ixx * ixx + t2 * ixy * ixy;
    B1 -= t * ixx * ixz + t2 * ixy * iyz;
    if (NOP == 2) {
      A12 += t * ixx * ixy + t2 * ixy * iyy;
      A22 += t2 * iyy * iyy + t * ixy * ixy;
      B2 -= t2 * iyy * iyz + t * ixy * ixz;
    }
    A11 *= 3;
    B1 *= 3;
    if (NOP == 2) {
      A12 *= 3;
      A22 *= 3;
      B2 *= 3;
    }
  } else {
    float tc[3], nc[3], tg[6], ng[6], acc, t;
    if (hdo3 != 0.0f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float ix = DRV(0, c), iy = DRV(1, c), iz = DRV(2, c);
        tc[c] = (NOP == 2) ? iz + ix * u + iy * v : iz + ix * u;
        nc[c] = ix * ix + iy * iy + DATANORM;
      }
      acc = tc[0] * tc[0] / nc[0] + tc[1] * tc[1] / nc[1] + tc[2] * tc[2] / nc[2] + EPS_COLOR;
      t = m * hdo3 / sqrtf(acc);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float ix = DRV(0, c), iy = DRV(1, c), iz = DRV(2, c);
        const float tt = t / nc[c];
        A11 += tt * ix * ix;
        B1 -= tt * iz * ix;
        if (NOP == 2) {
          A12 += tt * ix * iy;
          A22 += tt * iy * iy;
          B2 -= tt * iz * iy;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ixx = DRV(3, c), ixy = DRV(4, c), iyy = DRV(5, c), ixz = DRV(6, c), iyz = DRV(7, c);
      ng[2 * c] = ixx * ixx + ixy * ixy + DATANORM;
      ng[2 * c + 1] = iyy * iyy + ixy * ixy + DATANORM;
      tg[2 * c] = (NOP == 2) ? ixz + ixx * u + ixy * v : ixz + ixx * u;
      tg[2 * c + 1] = (NOP == 2) ? iyz + ixy * u + iyy * v : iyz + ixy * u;
    }
    acc = tg[0] * tg[0] / ng[0] + tg[1] * tg[1] / ng[1] + tg[2] * tg[2] / ng[2] + tg[3] * tg[3] / ng[3] +
          tg[4] * tg[4] / ng[4] + tg[5] * tg[5] / ng[5] + EPS_GRAD;
    t = m * hgo3 / sqrtf(acc);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ixx = DRV(3, c), ixy = DRV(4, c), iyy = DRV(5, c), ixz = DRV(6, c), iyz = DRV(7, c);
      const float ta = t / ng[2 * c], tb = t / ng[2 * c + 1];
      A11 += ta * ixx * ixx + tb * ixy * ixy;
      B1 -= ta * ixx * ixz + tb * ixy * iyz;
      if (NOP == 2) {
        A12 += ta * ixx * ixy + tb * ixy * iyy;
        A22 += tb * iyy * iyy + ta * ixy * ixy;
        B2 -= tb * iyy * iyz + ta * ixy * ixz;
      }
    }
  }
#undef DRV

  // sub_laplacian (opticalflow_aux.c:172-199), per-pixel order -h(i-1) +h(i) -v(j-1) +v(j)
  {
    const float* fc = flow + ((size_t)j * w + i) * NOP;
    const float wxc = fc[0];
    if (i > 0) B1 -= hl * (wxc - fc[-NOP]);
    if (i < w - 1) B1 += hh * (fc[NOP] - wxc);
    if (j > 0) B1 -= vt * (wxc - fc[-w * NOP]);
    if (j < h - 1) B1 += vv * (fc[w * NOP] - wxc);
    if (NOP == 2) {
      const float wyc = fc[1];
      if (i > 0) B2 -= hl * (wyc - fc[-NOP + 1]);
      if (i < w - 1) B2 += hh * (fc[NOP + 1] - wyc);
      if (j > 0) B2 -= vt * (wyc - fc[-w * NOP + 1]);
      if (j < h - 1) B2 += vv * (fc[w * NOP + 1] - wyc);
    }
  }

  if (NOP == 2) {
    // solver.c:115-120 (+ twins for first/last line): invert the 2x2 block
    float dps;
    if (j == 0) dps = hl + hh + vv;
    else if (j == h - 1) dps = hl + hh + vt;
    else dps = hl + hh + vt + vv;
    const float iA11 = A22 + dps, iA22 = A11 + dps;
    const float det = iA11 * iA22 - A12 * A12;
    float4* rec = pl.rec + ((size_t)fr * pl.plane + o) * 2;
    rec[0] = make_float4(iA11 / det, A12 / -det, iA22 / det, B1);
    rec[1] = make_float4(B2, hh, vv, 0.0f);
  } else {
    // sor_coupled_slow_but_readable_DE (solver.c:438-460): A11 = a11 + sum_dpsis (top,left,bottom,right)
    float sum = 0.0f;
    if (j > 0) sum += vt;
    if (i > 0) sum += hl;
    if (j < h - 1) sum += vv;
    if (i < w - 1) sum += hh;
    pl.rec[(size_t)fr * pl.plane + o] = make_float4(A11 + sum, B1, hh, vv);
  }
}

// ---------------------------------------------------------------------------
// Lexicographic SOR as a systolic wavefront.
//
// sor_coupled (solver.c:77-421) visits pixels in raster order; pixel (i,j) of
// sweep k reads left/top of sweep k and right/bottom (and itself) of sweep k-1.
// With the schedule  t = i + j + 2k  every value is produced exactly one step
// before its consumers need it, so thread (k,j) walks row j one pixel per step
// and exchanges (du,dv,sv) with its neighbours through a double-buffered
// shared-memory board; one __syncthreads per step.  The arithmetic per pixel is
// the reference's expression, hence bit-identical.  Sweep 0 takes the previous
// values from global memory (prefetched PF steps ahead), the last sweep writes
// the result (and, on the last inner iteration, flow = w + dw; K12).
template <int NOP, int PF>
__global__ void __launch_bounds__(PF > 2 ? 512 : 1024)
    sor_kernel(LevelGeom g, VarRefPlanes pl, VarRefParams vp, int f0, int K, int write_flow) {
  extern __shared__ float4 s_pub[];  // [2][K][h+2]
  const int fr = blockIdx.x, frame = f0 + fr;
  const int w = g.w, h = g.h, pitch = g.pitch;
  const int tid = threadIdx.x;
  const bool valid = tid < K * h;
  const int k = valid ? tid / h : 0, j = valid ? tid - k * h : 0;
  const int hb = h + 2;
  float4* pub0 = s_pub;
  float4* pub1 = s_pub + K * hb;
  constexpr int RF = (NOP == 2) ? 2 : 1;
  const float4* rec = pl.rec + ((size_t)fr * pl.plane + (size_t)j * pitch) * RF;
  float2* drow = pl.dudv + (size_t)fr * pl.plane + (size_t)j * pitch;
  const float omega = vp.omega;
  float* flow = g.flow + (size_t)frame * g.flow_frame_stride + (size_t)j * w * NOP;

  // prefetch ring: slot q holds column (t%PF==q) data
  float4 ra[PF], rb[PF];
  float2 rr[PF], rbt[PF];
#pragma unroll
  for (int q = 0; q < PF; ++q) {
    ra[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    rb[q] = ra[q];
    rr[q] = make_float2(0.f, 0.f);
    rbt[q] = rr[q];
  }
  // warm-up: columns 0..PF-1 are needed at steps t0..t0+PF-1, t0 = j+2k
  const int tstart = j + 2 * k;
  if (valid) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      // column c is consumed at step tstart + c, slot (tstart + c) % PF
      const int c = q;
      const int slot = (tstart + c) % PF;
      if (c < w) {
#pragma unroll
        for (int s = 0; s < PF; ++s)
          if (s == slot) {
            ra[s] = __ldg(rec + (size_t)c * RF);
            if (NOP == 2) rb[s] = __ldg(rec + (size_t)c * RF + 1);
            if (k == 0) {
              if (c + 1 < w) rr[s] = drow[c + 1];
              if (j < h - 1) rbt[s] = drow[pitch + c];
            }
          }
      }
    }
  }
  float2 own = make_float2(0.f, 0.f);
  if (valid && k == 0) own = drow[0];
  float du_l = 0.f, dv_l = 0.f, hl = 0.f;

  const int S = w + h + 2 * K - 3;  // last step: column w-1 of row h-1 in sweep K-1
  for (int tb = 0; tb < S; tb += PF) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      const int t = tb + q;
      if (t < S) {
        const int i = t - tstart;
        float4* pprev = (t & 1) ? pub0 : pub1;
        float4* pcur = (t & 1) ? pub1 : pub0;
        if (valid && i >= -1 && i < w) {
          // "right" neighbour = column i+1 of the previous sweep
          float2 right = make_float2(0.f, 0.f);
          if (k == 0) {
            if (i < 0) right = own;            // column 0 was loaded before the loop
            else if (i + 1 < w) right = rr[q];
          } else if (i + 1 < w) {
            const float4 v = pprev[(k - 1) * hb + j + 1];
            right = make_float2(v.x, v.y);
          }
          if (i >= 0) {
            const float4 A = ra[q];
            float4 B = rb[q];
            float2 bot = make_float2(0.f, 0.f);
            if (j < h - 1) {
              if (k == 0) bot = rbt[q];
              else {
                const float4 v = pprev[(k - 1) * hb + j + 2];
                bot = make_float2(v.x, v.y);
              }
            }
            float4 top = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j > 0) top = pprev[k * hb + j];
            float du, dv = 0.f, svc;
            if (NOP == 2) {
              const float a11 = A.x, a12 = A.y, a22 = A.z, b1 = A.w, b2 = B.x, hh = B.y, vv = B.z;
              const float vt = top.z;
              float s1, s2;
              if (j == 0) {
                s1 = hh * right.x + vv * bot.x + b1;
                s2 = hh * right.y + vv * bot.y + b2;
              } else if (j == h - 1) {
                s1 = hh * right.x + vt * top.x + b1;
                s2 = hh * right.y + vt * top.y + b2;
              } else {
                s1 = hh * right.x + vt * top.x + vv * bot.x + b1;
                s2 = hh * right.y + vt * top.y + vv * bot.y + b2;
              }
              float B1 = s1, B2 = s2;
              if (i > 0) {
                B1 = hl * du_l + s1;
                B2 = hl * dv_l + s2;
              }
              du = own.x + omega * (a11 * B1 + a12 * B2 - own.x);
              dv = own.y + omega * (a12 * B1 + a22 * B2 - own.y);
              hl = hh;
              svc = vv;
            } else {
              const float A11 = A.x, b1 = A.y, hh = A.z, vv = A.w;
              float sigma = 0.0f;
              if (j > 0) sigma -= top.z * top.x;
              if (i > 0) sigma -= hl * du_l;
              if (j < h - 1) sigma -= vv * bot.x;
              if (i < w - 1) sigma -= hh * right.x;
              const float B1 = b1 - sigma;
              du = (1.0f - omega) * own.x + omega * (B1 / A11);
              hl = hh;
              svc = vv;
            }
            pcur[k * hb + j + 1] = make_float4(du, dv, svc, 0.f);
            du_l = du;
            dv_l = dv;
            if (k == K - 1) {
              drow[i] = make_float2(du, dv);
              if (write_flow) {
                if (NOP == 2) {
                  float2* f2 = reinterpret_cast<float2*>(flow) + i;
                  const float2 wv = *f2;
                  *f2 = make_float2(wv.x + du, wv.y + dv);
                } else {
                  const float tsum = flow[i] + du;
                  flow[i] = (g.camlr == 0) ? (tsum < 0.0f ? tsum : 0.0f) : (tsum > 0.0f ? tsum : 0.0f);
                }
              }
            }
            // refill this slot with column i+PF
            const int c = i + PF;
            if (c < w) {
              ra[q] = __ldg(rec + (size_t)c * RF);
              if (NOP == 2) rb[q] = __ldg(rec + (size_t)c * RF + 1);
              if (k == 0) {
                if (c + 1 < w) rr[q] = drow[c + 1];
                if (j < h - 1) rbt[q] = drow[pitch + c];
              }
            }
          }
          own = right;
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace

template <int C, int NOP>
static int launch_varref_t(const LevelGeom& g, const VarRefPlanes& pl, const VarRefParams& vp, int f0, int f1,
                           cudaStream_t st, Profiler* prof) {
  int launches = 0;
  const int nf = f1 - f0;
  const dim3 block(TX, TY), grid((g.w + TX - 1) / TX, (g.h + TY - 1) / TY, nf);
  const dim3 gridc(grid.x, grid.y, nf * C);
  {
    ProfScope scope(prof, KC_VR_SETUP);
    warp_kernel<C, NOP><<<grid, block, 0, st>>>(g, pl, f0);
    deriv1_kernel<C><<<gridc, block, 0, st>>>(g, pl);
    deriv2_kernel<C><<<gridc, block, 0, st>>>(g, pl);
    cudaMemsetAsync(pl.dudv, 0, sizeof(float2) * pl.plane * nf, st);
  }
  launches += 3;
  // sweeps per SOR launch: all of them when (sweeps x rows) fits one CTA
  const int K = vp.n_solver;
  const bool fused = (K >= 1) && (K * g.h <= 1024);
  const int kl = fused ? K : 1;
  const int nthreads = ((kl * g.h + 31) / 32) * 32;
  const size_t smem = sizeof(float4) * 2 * kl * (g.h + 2);
  const bool big = nthreads > 512;
  if (K >= 1) {
    if (big) cudaFuncSetAttribute(sor_kernel<NOP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(sor_kernel<NOP, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  for (int it = 0; it < vp.n_inner; ++it) {
    {
      ProfScope scope(prof, KC_VR_ASSEMBLE);
      assemble_kernel<C, NOP><<<grid, block, 0, st>>>(g, pl, vp, f0, it == 0 ? 1 : 0);
    }
    ++launches;
    const int last = (it == vp.n_inner - 1) ? 1 : 0;
    const int nl = fused ? 1 : K;
    for (int s = 0; s < nl; ++s) {
      const int wf = (last && s == nl - 1) ? 1 : 0;
      ProfScope scope(prof, KC_VR_SOR);
      if (big) sor_kernel<NOP, 2><<<nf, nthreads, smem, st>>>(g, pl, vp, f0, kl, wf);
      else sor_kernel<NOP, 8><<<nf, nthreads, smem, st>>>(g, pl, vp, f0, kl, wf);
      ++launches;
    }
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

int launch_varref(const LevelGeom& g, const VarRefPlanes& pl, const VarRefParams& vp, int f0, int f1,
                  cudaStream_t st, Profiler* prof) {
  if (g.noc == 1 && g.nop == 2) return launch_varref_t<1, 2>(g, pl, vp, f0, f1, st, prof);
  if (g.noc == 3 && g.nop == 2) return launch_varref_t<3, 2>(g, pl, vp, f0, f1, st, prof);
  if (g.noc == 1 && g.nop == 1) return launch_varref_t<1, 1>(g, pl, vp, f0, f1, st, prof);
  if (g.noc == 3 && g.nop == 1) return launch_varref_t<3, 1>(g, pl, vp, f0, f1, st, prof);
  return -1;
}

}  // namespace ofdis

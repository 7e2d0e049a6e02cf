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
//   sor_wave_kernel  all sweeps of the lexicographic SOR as a systolic wavefront, one CTA or one
//                    thread-block cluster per frame (sor_wave_kernel.cuh)
//
// Every expression keeps the reference's operand order; the TU is compiled with
// -fmad=false so nothing is contracted (bit-exactness).
#include "ofdis_internal.cuh"

#include <type_traits>

namespace ofdis {

namespace {

#ifdef OFDIS_SOR_TIMING
// debug build only: per-warp cycle stamps of a few super-steps of frame 0 (tools/sor_timing.py)
__device__ long long g_sor_times[64 * 8 * 16];  // sor_lane_kernel: [warp 16][chunk 32][4 stamps]
#define SOR_STAMP(slot, dep1, dep2)                                                           \
  do {                                                                                        \
    if (fr == 0 && (tid & 31) == 0 && T >= 40 && T < 48) {                                    \
      long long t__;                                                                          \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t__) : "f"(dep1), "f"(dep2) : "memory");   \
      g_sor_times[((tid >> 5) * 8 + (T - 40)) * 16 + (slot)] = t__;                           \
    }                                                                                         \
  } while (0)
#else
#define SOR_STAMP(slot, dep1, dep2) do { } while (0)
#endif

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
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  const int fr = blockIdx.z, frame = frame_of(g, f0, fr);
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
  const float* i1 = g.img[3] + (size_t)frame * g.img_fs[3];
  const float* i0 = g.img[0] + (size_t)frame * g.img_fs[0];
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
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
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
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
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

// R rows per thread (tile 32 x 8R): the halo work of the two staging phases -- (TH+4)x36 flow
// values and (TH+2)x34 smoothness weights per 32 x TH pixels -- shrinks from 1.69x / 1.33x (R=1)
// to 1.27x / 1.13x (R=4).  All indices inside a frame are 32-bit.
// MODE: 0 = records for sor_wave_kernel (band_f4 lane rows), 1 = fast mode (natural layout), 2 = sor_lane_kernel
// (lane-skewed layout); a template parameter so that the layout arithmetic of the other modes costs nothing.
template <int C, int NOP, int R, int MODE>
__global__ void __launch_bounds__(TX * TY) assemble_kernel(LevelGeom g, VarRefPlanes pl, VarRefParams vp,
                                                             int f0, int first) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int TH = TY * R;
  __shared__ float2 s_uv[TH + 4][TX + 4];
  __shared__ float s_s[TH + 2][TX + 2];
  const int fr = blockIdx.z, frame = frame_of(g, f0, fr);
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TH;
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int w = g.w, h = g.h, pitch = g.pitch;
  const float* const flow = g.flow + (size_t)frame * g.flow_frame_stride;
  // Records and (du,dv) of pixel (x,y).  Exact mode: the band-skewed lane rows (band_f4): chunk f of a
  // block holds field f of its 4 pixels, du is chunk nq, dv chunk nq+1.  Fast mode (red-black SOR):
  // natural layout, 8 floats per pixel, (du,dv) in the current ping-pong planes.
  // Lane mode (sor_lane_kernel): lane-skewed layout, the record is two float4 (lane_rec_f4), (du,dv) one float2.
  constexpr bool fast = (MODE == 1), lane = (MODE == 2);
  float* const rec = fast ? pl.frec + (size_t)fr * pl.frec_stride : reinterpret_cast<float*>(pl.rec + (size_t)fr * pl.rec_stride);
  float* const dudv = fast ? pl.fdu + (size_t)fr * pl.fdu_stride + (size_t)pl.fcur * 2 * pl.plane : rec;
  const int fs = fast ? 1 : 4;                       // floats between consecutive record fields of a pixel
  const int dv_off = lane ? 1 : (fast ? (int)pl.plane : 4);  // from du to dv
  auto rec_idx = [&pl, pitch](int x, int y) {
    return lane ? (int)lane_rec_f4(pl, x, y, 0) * 4 : (fast ? (y * pitch + x) * 8 : (int)band_f4(pl, x >> 2, y, 0) * 4 + (x & 3));
  };
  auto du_idx = [&pl, pitch](int x, int y) {
    return lane ? (int)lane_dudv_f(pl, x, y) : (fast ? y * pitch + x : (int)band_f4(pl, x >> 2, y, pl.nq) * 4 + (x & 3));
  };

  // uu = wx + du (vv likewise); first iteration: uu = wx (refine_variational.cpp:189-190).
  // Coordinates are clamped, which also realises the replicate border of the 3-tap
  // horizontal derivative (image.c:448-454).
  for (int idx = tid; idx < (TH + 4) * (TX + 4); idx += TX * TY) {
    const int cy = idx / (TX + 4), cx = idx - cy * (TX + 4);
    const int gx = clampi(x0 - 2 + cx, w), gy = clampi(y0 - 2 + cy, h);
    const float* f = flow + (gy * w + gx) * NOP;
    float2 uv;
    uv.x = f[0];
    uv.y = (NOP == 2) ? f[1] : 0.0f;
    if (!first) {
      const int b = du_idx(gx, gy);
      const float dx = dudv[b];
      if (NOP == 2) {
        uv.x = uv.x + dx;
        uv.y = uv.y + dudv[b + dv_off];
      } else {  // minps / maxps with zero (refine_variational.cpp:299-314)
        const float t = uv.x + dx;
        uv.x = (camlr_of(g, frame) == 0) ? (t < 0.0f ? t : 0.0f) : (t > 0.0f ? t : 0.0f);
      }
    }
    s_uv[cy][cx] = uv;
  }
  __syncthreads();

  // smoothness weight s = quarter_alpha / sqrt(ux^2+uy^2+vx^2+vy^2+eps) on the tile + 1 halo
  {
    const float c0 = -0.5f, c1 = -0.0f, c2 = 0.5f;  // {0,-0.5} -> [-0.5,-0,0.5]
    for (int idx = tid; idx < (TH + 2) * (TX + 2); idx += TX * TY) {
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

  const int i = x0 + threadIdx.x;
  if (i >= w) return;
  const float hdo3 = vp.half_delta_over3, hgo3 = vp.half_gamma_over3;
  const float* const maskp = pl.mask + (size_t)fr * pl.plane;

#pragma unroll 1
  for (int rr = 0; rr < R; ++rr) {
  const int ly = threadIdx.y + TY * rr, j = y0 + ly;
  if (j >= h) break;
  const int cx = threadIdx.x + 1, cy = ly + 1;
  const float sc = s_s[cy][cx];
  const float hh = (i < w - 1) ? sc + s_s[cy][cx + 1] : 0.0f;    // sh(i,j)   (opticalflow_aux.c:150-154)
  const float hl = (i > 0) ? s_s[cy][cx - 1] + sc : 0.0f;        // sh(i-1,j)
  const float vv = (j < h - 1) ? sc + s_s[cy + 1][cx] : 0.0f;    // sv(i,j)   (:159-163)
  const float vt = (j > 0) ? s_s[cy - 1][cx] + sc : 0.0f;        // sv(i,j-1)

  const int o = j * pitch + i;
  const float m = maskp[o];
  const int b0 = rec_idx(i, j), bd = du_idx(i, j);
  // du, dv of this pixel; the first inner iteration starts from zero and resets the stored values (which the
  // SOR's first sweep reads): refine_variational.cpp:181-182.  The last thread of a row also clears the
  // columns >= w of the row's last block, which the exact SOR updates but nobody reads.
  float u = 0.0f, v = 0.0f;
  if (first) {
    dudv[bd] = 0.0f;
    dudv[bd + dv_off] = 0.0f;
    if (!fast && !lane && i == w - 1)
      for (int t = (i & 3) + 1; t < 4; ++t) dudv[bd + t - (i & 3)] = dudv[bd + 4 + t - (i & 3)] = 0.0f;
  } else {
    u = dudv[bd];
    v = (NOP == 2) ? dudv[bd + dv_off] : 0.0f;
  }
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
    const float* fc = flow + (j * w + i) * NOP;
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
    // record of the 4-pixel block, SoA: float4 f of the block holds field f of its 4 pixels;
    // fields: a11^-1, a12^-1, a22^-1, b1, b2, sh, sv, sv(row above)
    if (lane) {
      float4* const r4 = reinterpret_cast<float4*>(rec + b0);
      r4[0] = make_float4(iA11 / det, A12 / -det, iA22 / det, B1);
      r4[32] = make_float4(B2, hh, vv, vt);
    } else {
    rec[b0] = iA11 / det;
    rec[b0 + fs] = A12 / -det;
    rec[b0 + 2 * fs] = iA22 / det;
    rec[b0 + 3 * fs] = B1;
    rec[b0 + 4 * fs] = B2;
    rec[b0 + 5 * fs] = hh;
    rec[b0 + 6 * fs] = vv;
    rec[b0 + 7 * fs] = vt;
    }
  } else {
    // sor_coupled_slow_but_readable_DE (solver.c:438-460): A11 = a11 + sum_dpsis (top,left,bottom,right)
    float sum = 0.0f;
    if (j > 0) sum += vt;
    if (i > 0) sum += hl;
    if (j < h - 1) sum += vv;
    if (i < w - 1) sum += hh;
    // stereo record fields: A11 = a11 + sum, b1, sh, sv, sv(row above)
    if (lane) {
      float4* const r4 = reinterpret_cast<float4*>(rec + b0);
      r4[0] = make_float4(A11 + sum, B1, hh, vv);
      r4[32] = make_float4(vt, 0.f, 0.f, 0.f);
    } else {
    rec[b0] = A11 + sum;
    rec[b0 + fs] = B1;
    rec[b0 + 2 * fs] = hh;
    rec[b0 + 3 * fs] = vv;
    rec[b0 + 4 * fs] = vt;
    }
  }
  }  // rows of this thread
}

// ---------------------------------------------------------------------------
// Shared-memory access helpers of the SOR kernel (explicit 128-bit forms).
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// predicated forms: lanes without a block issue no shared-memory wavefronts.  The result registers of such a
// lane keep whatever they held (no zero-fill: 44 CS2R per super-step); its arithmetic runs on that garbage and
// is never stored, and no branch of the kernel depends on data of a lane without a block.
__device__ __forceinline__ float4 lds128_if(bool p, unsigned addr) {
  float4 v;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr), "r"((unsigned)p)
               : "memory");
  return v;
}
__device__ __forceinline__ float lds32_if(bool p, unsigned addr) {
  float v;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.f32 %0, [%1];\n\t}" : "=f"(v) : "r"(addr), "r"((unsigned)p) : "memory");
  return v;
}
__device__ __forceinline__ void sts128_if(bool p, unsigned addr, const float4& v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n\t}" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "r"((unsigned)p)
               : "memory");
}
__device__ __forceinline__ void sts128(unsigned addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#include "sor_wave_kernel.cuh"
#include "sor_lane_kernel.cuh"
#include "sor_redblack_kernel.cuh"

// K12 at the end of the level: flow = w + dw (refine_variational.cpp:210-221; stereo clamp
// :299-314).  Kept out of the SOR kernel: the flow array is row-major per frame, so reading it
// from one-thread-per-row SOR lanes costs 32 cache lines per load instruction.
template <int NOP>
__global__ void __launch_bounds__(256) flow_update_kernel(LevelGeom g, VarRefPlanes pl, int f0) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  const int fr = blockIdx.z, frame = frame_of(g, f0, fr);
  if (i >= g.w || j >= g.h) return;
  const bool fast = pl.fast != 0, lane = pl.lane != 0;
  const float* dudv = fast ? pl.fdu + (size_t)fr * pl.fdu_stride + (size_t)pl.fcur * 2 * pl.plane
                           : reinterpret_cast<const float*>(pl.rec + (size_t)fr * pl.rec_stride);
  const size_t b = lane ? lane_dudv_f(pl, i, j) : (fast ? (size_t)j * g.pitch + i : band_f4(pl, i >> 2, j, pl.nq) * 4 + (i & 3));
  const size_t dv_off = lane ? 1 : (fast ? pl.plane : 4);
  float* f = g.flow + (size_t)frame * g.flow_frame_stride + ((size_t)j * g.w + i) * NOP;
  if (NOP == 2) {
    const float2 wv = *reinterpret_cast<const float2*>(f);
    *reinterpret_cast<float2*>(f) = make_float2(wv.x + dudv[b], wv.y + dudv[b + dv_off]);
  } else {
    const float t = f[0] + dudv[b];
    f[0] = (camlr_of(g, frame) == 0) ? (t < 0.0f ? t : 0.0f) : (t > 0.0f ? t : 0.0f);
  }
}

}  // namespace

// Largest number of sweeps one launch can keep in flight: K*hpad compute threads + the producer
// warp within the kernel's launch bound, stage ring + board within the 227 KB of an SM.
static int sor_sweeps_per_launch(int nop, int hpad, int rt, int K) {
  int kl = K < 1 ? 1 : K;
  while (kl > 1 && (kl * hpad + 32 > sor_max_threads(hpad) || sor_smem_bytes(nop, hpad, rt, kl) > 227 * 1024)) --kl;
  return kl;
}

template <int NOP, int HPAD, int RT, bool CL>
static cudaError_t launch_sor_t(const LevelGeom& g, const VarRefPlanes& pl, const VarRefParams& vp, int nf, int kl,
                                cudaStream_t st) {
  auto kern = sor_wave_kernel<NOP, HPAD, RT, CL>;
  const size_t smem = sor_smem_bytes(NOP, HPAD, RT, kl);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  // opt-in shared memory (and cluster size) once per device and instantiation
  static size_t smem_set[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && smem_set[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && CL) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    smem_set[dev] = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nf * (CL ? pl.nb : 1)));
  cfg.blockDim = dim3((unsigned)(kl * HPAD + 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)pl.nb;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g.pdl) {  // see pdl_wait (ofdis_internal.cuh)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, g, pl, vp, kl);
}

template <int NOP, int RT>
static cudaError_t launch_sor_rt(const LevelGeom& g, const VarRefPlanes& pl, const VarRefParams& vp, int nf, int kl,
                                 cudaStream_t st) {
  const bool cl = pl.nb > 1;
  switch (pl.hpad) {
    case 32: return cl ? launch_sor_t<NOP, 32, RT, true>(g, pl, vp, nf, kl, st) : launch_sor_t<NOP, 32, RT, false>(g, pl, vp, nf, kl, st);
    case 64: return cl ? launch_sor_t<NOP, 64, RT, true>(g, pl, vp, nf, kl, st) : launch_sor_t<NOP, 64, RT, false>(g, pl, vp, nf, kl, st);
    case 128: return cl ? launch_sor_t<NOP, 128, RT, true>(g, pl, vp, nf, kl, st) : launch_sor_t<NOP, 128, RT, false>(g, pl, vp, nf, kl, st);
    case 256: return cl ? launch_sor_t<NOP, 256, RT, true>(g, pl, vp, nf, kl, st) : launch_sor_t<NOP, 256, RT, false>(g, pl, vp, nf, kl, st);
  }
  return cudaErrorInvalidValue;
}

template <int NOP>
static cudaError_t launch_sor(const LevelGeom& g, const VarRefPlanes& pl, const VarRefParams& vp, int nf, int kl,
                              cudaStream_t st) {
  if (pl.rt == 1) return launch_sor_rt<NOP, 1>(g, pl, vp, nf, kl, st);
  if (pl.rt == 2) return launch_sor_rt<NOP, 2>(g, pl, vp, nf, kl, st);
  if (pl.rt == 4) return launch_sor_rt<NOP, 4>(g, pl, vp, nf, kl, st);
  return cudaErrorInvalidValue;
}

template <int C, int NOP>
static int launch_varref_t(const LevelGeom& g, const VarRefPlanes& pl_in, const VarRefParams& vp, int f0, int f1,
                           cudaStream_t st, Profiler* prof) {
  VarRefPlanes pl = pl_in;  // fast mode toggles the (du,dv) ping-pong buffer
  const bool pdl = g.pdl != 0 && prof == nullptr;  // the profiler's events between launches would serialise them anyway
  pl.fcur = 0;
  int launches = 0;
  const int nf = f1 - f0;
  const dim3 block(TX, TY), grid((g.w + TX - 1) / TX, (g.h + TY - 1) / TY, nf);
  const dim3 gridc(grid.x, grid.y, nf * C);
  // assemble_kernel: as many rows per thread (less halo work) as still leave >= 256 CTAs
  int rows_per_thread = 1;
  for (int r = 4; r > 1; r >>= 1)
    if ((long)grid.x * ((g.h + TY * r - 1) / (TY * r)) * nf >= 256) { rows_per_thread = r; break; }
  const dim3 grid_a(grid.x, (g.h + TY * rows_per_thread - 1) / (TY * rows_per_thread), nf);
  {
    ProfScope scope(prof, KC_VR_SETUP);
    launch_k(pdl, warp_kernel<C, NOP>, grid, block, 0, st, g, pl, f0);
    launch_k(pdl, deriv1_kernel<C>, gridc, block, 0, st, g, pl);
    launch_k(pdl, deriv2_kernel<C>, gridc, block, 0, st, g, pl);
  }
  launches += 3;
  // SOR: band plan of the level (pl.hpad rows per band, pl.nb bands == CTAs of a cluster) and as
  // many sweeps per launch as the CTA's thread and shared-memory budgets hold (sweeps are sequential,
  // so K sweeps in ceil(K / kl) launches give the same result)
  const int K = vp.n_solver;
  const int kl = sor_sweeps_per_launch(NOP, pl.hpad, pl.rt, K);
  for (int it = 0; it < vp.n_inner; ++it) {
    {
      ProfScope scope(prof, KC_VR_ASSEMBLE);
      auto launch_asm = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        if (rows_per_thread == 4) launch_k(pdl, assemble_kernel<C, NOP, 4, MODE>, grid_a, block, 0, st, g, pl, vp, f0, it == 0 ? 1 : 0);
        else if (rows_per_thread == 2) launch_k(pdl, assemble_kernel<C, NOP, 2, MODE>, grid_a, block, 0, st, g, pl, vp, f0, it == 0 ? 1 : 0);
        else launch_k(pdl, assemble_kernel<C, NOP, 1, MODE>, grid_a, block, 0, st, g, pl, vp, f0, it == 0 ? 1 : 0);
      };
      if (pl.fast) launch_asm(std::integral_constant<int, 1>{});
      else if (pl.lane) launch_asm(std::integral_constant<int, 2>{});
      else launch_asm(std::integral_constant<int, 0>{});
    }
    ++launches;
    if (pl.fast) {  // opt-in red-black solver: all K sweeps in one launch, (du,dv) ping-pong
      ProfScope scope(prof, KC_VR_SOR);
      const size_t smem = rb_smem_bytes(NOP, K);
      if (K < 1 || smem > 227 * 1024) return -1;
      static size_t smem_set[64] = {0};
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev >= 0 && dev < 64 && smem_set[dev] < smem) {
        if (cudaFuncSetAttribute(sor_redblack_kernel<NOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        smem_set[dev] = smem;
      }
      const dim3 grid_rb((g.w + RB_TILE - 1) / RB_TILE, (g.h + RB_TILE - 1) / RB_TILE, nf);
      launch_k(pdl, sor_redblack_kernel<NOP>, grid_rb, dim3(256), smem, st, g, pl, vp);
      pl.fcur ^= 1;
      ++launches;
      continue;
    }
    if (pl.lane) {  // pixel wavefront, warps synchronised through shared-memory flags (levels of few 32-row bands)
      const int kll = sl_sweeps_per_launch(pl.nb, K);
      if (kll < 1) return -1;
      for (int s = 0; s < K; s += kll) {
        ProfScope scope(prof, KC_VR_SOR);
        const int kk = (K - s < kll) ? K - s : kll;
        const size_t smem = sl_smem_bytes(pl.nb, kk);
        static size_t smem_set[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && smem_set[dev] < smem) {
          if (cudaFuncSetAttribute(sor_lane_kernel<NOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
          smem_set[dev] = smem;
        }
        launch_k(pdl, sor_lane_kernel<NOP>, dim3(nf), dim3(pl.nb * kk * 32), smem, st, g, pl, vp, kk);
        ++launches;
      }
      continue;
    }
    for (int s = 0; s < K; s += kl) {
      ProfScope scope(prof, KC_VR_SOR);
      if (launch_sor<NOP>(g, pl, vp, nf, (K - s < kl) ? K - s : kl, st) != cudaSuccess) return -1;
      ++launches;
    }
  }
  if (vp.n_inner > 0) {
    ProfScope scope(prof, KC_VR_SETUP);
    launch_k(pdl, flow_update_kernel<NOP>, grid, block, 0, st, g, pl, f0);
    ++launches;
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

#ifdef OFDIS_SOR_TIMING
extern "C" int ofdis_debug_sor_times(long long* dst) {
  return cudaMemcpyFromSymbol(dst, g_sor_times, sizeof(g_sor_times)) == cudaSuccess ? 0 : -1;
}
#endif

bool sor_lane_fits(int h, int K) { return sl_sweeps_per_launch((h + 31) / 32, K) >= 1; }
// Where the lane kernel beats the block wavefront (tools/lane_ab.py, profiles/): one or two bands with all sweeps
// in one launch.  Every further band adds 33 steps of start-up skew, and 1920x1080's 136- and 68-row levels
// (5 and 3 bands) run 20-50 % slower with it.
bool sor_lane_preferred(int h, int K) { return (h + 31) / 32 <= 2 && sl_sweeps_per_launch((h + 31) / 32, K) >= (K < 1 ? 1 : K); }

bool rb_smem_limit_exceeded(int nop, int K) { return K < 1 || rb_smem_bytes(nop, K) > 227 * 1024; }

bool sor_fits(int nop, int hpad, int rt, int K) {
  return K * hpad + 32 <= sor_max_threads(hpad) && sor_smem_bytes(nop, hpad, rt, K) <= 227 * 1024;
}

int sor_max_cluster_size() {
  // 16 CTAs is a non-portable cluster size: ask the occupancy calculator whether one such cluster
  // of the largest SOR configuration (256-row bands, one sweep) fits this device
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 8;
  if (cached[dev]) return cached[dev];
  int best = 8;
  auto kern = sor_wave_kernel<2, 128, 1, true>;
  const size_t smem = sor_smem_bytes(2, 128, 1, 3);  // 128-row bands, 3 sweeps in flight: the largest common configuration
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(3 * 128 + 32);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1) best = 16;
  }
  cudaGetLastError();  // a refused query must not poison later launches
  cached[dev] = best;
  return best;
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

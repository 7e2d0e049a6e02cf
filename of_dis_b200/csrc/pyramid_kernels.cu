// Image pyramid, gradients and output upsampling on the device -- the callers either side of the
// hot path: ConstructImgPyramide (run_dense.cpp:130-178), the
// divisibility padding (run_dense.cpp:298-311) and the final resize/crop (run_dense.cpp:407-414).
// Expression order follows of_dis_b200/preprocess.py, which for 8-bit input is bit-identical to
// OpenCV (every intermediate is a dyadic rational that float32 holds exactly).
#include <cuda_runtime.h>

#include "ofdis_internal.cuh"

namespace ofdis {

namespace {

// Level `g.level` of both images of a pair, straight from the 8-bit frames: the 2^l x 2^l box sum
// is exact in int32 and (for l <= 8) in float32, so it equals l successive cv::resize(0.5) steps.
// The divisibility padding (replicate, floor(pad/2) left/top) and the per-level border padding
// (replicate, g.pad) are folded into the index clamps.  One thread per padded destination pixel.
__global__ void __launch_bounds__(256) pyr_from_u8_kernel(LevelGeom g, int f0, PyrSourceU8 s) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y * blockDim.y + threadIdx.y;
  if (xp >= g.tmp_w || yp >= g.tmp_h) return;
  const int fr = blockIdx.z >> 1, k = blockIdx.z & 1;  // k: 0 = I0, 1 = I1
  const int C = g.noc, sh = g.level, n = 1 << sh;
  const unsigned char* src = s.frames + ((size_t)fr * 2 + k) * s.image_bytes;
  const int arr = k ? 3 : 0;
  float* dst = const_cast<float*>(g.img[arr]) + (size_t)frame_of(g, f0, fr) * g.img_fs[arr] + ((size_t)yp * g.tmp_w + xp) * C;
  const int x = clampi(xp - g.pad, g.w), y = clampi(yp - g.pad, g.h);
  const float scale = __int_as_float((127 - 2 * sh) << 23);  // 4^-l
  int sum[3] = {0, 0, 0};
  for (int dy = 0; dy < n; ++dy) {
    const int Y = clampi((y << sh) + dy - s.pad_top, s.h_org);
    const unsigned char* row = src + (size_t)Y * s.w_org * C;
    for (int dx = 0; dx < n; ++dx) {
      const int X = clampi((x << sh) + dx - s.pad_left, s.w_org);
      for (int c = 0; c < C; ++c) sum[c] += (int)__ldg(row + X * C + c);
    }
  }
  for (int c = 0; c < C; ++c) dst[c] = (float)sum[c] * scale;
}

// Border padding of un-padded float images of level g.level ([frame][2][h][w][C], I0 then I1).
__global__ void __launch_bounds__(256) pyr_from_level_kernel(LevelGeom g, int f0, const float* stage) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y * blockDim.y + threadIdx.y;
  if (xp >= g.tmp_w || yp >= g.tmp_h) return;
  const int fr = blockIdx.z >> 1, k = blockIdx.z & 1;
  const int C = g.noc, arr = k ? 3 : 0;
  const float* src = stage + ((size_t)fr * 2 + k) * ((size_t)g.w * g.h * C);
  float* dst = const_cast<float*>(g.img[arr]) + (size_t)frame_of(g, f0, fr) * g.img_fs[arr] + ((size_t)yp * g.tmp_w + xp) * C;
  const int x = clampi(xp - g.pad, g.w), y = clampi(yp - g.pad, g.h);
  for (int c = 0; c < C; ++c) dst[c] = __ldg(src + ((size_t)y * g.w + x) * C + c);
}

// cv::resize(0.5, 0.5, INTER_LINEAR) of an even-sized image == 2x2 box mean (run_dense.cpp:150),
// ((a+b)+(c+d))*0.25 with a,b the even row; reads the interior of the padded level gs, writes
// level gd = gs+1 including its replicate border.
__global__ void __launch_bounds__(256) pyr_down_kernel(LevelGeom gs, LevelGeom gd, int f0) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y * blockDim.y + threadIdx.y;
  if (xp >= gd.tmp_w || yp >= gd.tmp_h) return;
  const int fr = blockIdx.z >> 1, k = blockIdx.z & 1;
  const int C = gd.noc, arr = k ? 3 : 0;
  const float* src = gs.img[arr] + (size_t)frame_of(gd, f0, fr) * gs.img_fs[arr];
  float* dst = const_cast<float*>(gd.img[arr]) + (size_t)frame_of(gd, f0, fr) * gd.img_fs[arr] + ((size_t)yp * gd.tmp_w + xp) * C;
  const int x = clampi(xp - gd.pad, gd.w), y = clampi(yp - gd.pad, gd.h);
  const float* r0 = src + ((size_t)(2 * y + gs.pad) * gs.tmp_w + (2 * x + gs.pad)) * C;
  const float* r1 = r0 + (size_t)gs.tmp_w * C;
  for (int c = 0; c < C; ++c) dst[c] = ((r0[c] + r0[C + c]) + (r1[c] + r1[C + c])) * 0.25f;
}

// Gradients of I0 on the device (the first "next" row of SURVEY 8f): cv::Sobel(CV_32F, 3x3,
// scale 1/8, BORDER_DEFAULT = reflect101) on the un-padded level image, zero border of width
// `pad` (run_dense.cpp:156-157,171-172).  Same expression order as of_dis_b200/preprocess.py
// (row difference first, then the [1 2 1]/8 column sum, and vice versa for dy); for images that
// come from 8-bit input every intermediate is exact, so this equals OpenCV bit for bit.
__global__ void __launch_bounds__(256) sobel_kernel(LevelGeom g, int f0) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y * blockDim.y + threadIdx.y;
  const int frame = frame_of(g, f0, blockIdx.z);
  if (xp >= g.tmp_w || yp >= g.tmp_h) return;
  const int C = g.noc, w = g.w, h = g.h, P = g.pad;
  const float* im = g.img[0] + (size_t)frame * g.img_fs[0];
  float* gx = const_cast<float*>(g.img[1]) + (size_t)frame * g.img_fs[1];
  float* gy = const_cast<float*>(g.img[2]) + (size_t)frame * g.img_fs[2];
  const int x = xp - P, y = yp - P;
  const size_t o = ((size_t)yp * g.tmp_w + xp) * C;
  if (x < 0 || y < 0 || x >= w || y >= h) {
    for (int c = 0; c < C; ++c) {
      gx[o + c] = 0.f;
      gy[o + c] = 0.f;
    }
    return;
  }
  auto r101 = [](int v, int n) { return n == 1 ? 0 : (v < 0 ? -v : (v >= n ? 2 * (n - 1) - v : v)); };
  const int xm = r101(x - 1, w) + P, x0 = x + P, xq = r101(x + 1, w) + P;
  const int ym = r101(y - 1, h) + P, y0 = y + P, yq = r101(y + 1, h) + P;
  for (int c = 0; c < C; ++c) {
#define IM(X, Y) im[((size_t)(Y) * g.tmp_w + (X)) * C + c]
    const float t0 = IM(xq, ym) - IM(xm, ym), t1 = IM(xq, y0) - IM(xm, y0), t2 = IM(xq, yq) - IM(xm, yq);
    gx[o + c] = (t0 * 0.125f + t1 * 0.25f) + t2 * 0.125f;
    const float s0 = (IM(xm, ym) * 0.125f + IM(x0, ym) * 0.25f) + IM(xq, ym) * 0.125f;
    const float s2 = (IM(xm, yq) * 0.125f + IM(x0, yq) * 0.25f) + IM(xq, yq) * 0.125f;
    gy[o + c] = s2 - s0;
#undef IM
  }
}

// Output stage of run_dense.cpp:407-414: flow * 2^lv_l, cv::resize(x 2^lv_l, INTER_LINEAR)
// (src = (dst + .5)/s - .5, edge clamped, horizontal pass first), crop of the divisibility
// padding.  One thread per full-resolution pixel; expression order of preprocess.upsample_linear.
template <int NOP>
__global__ void __launch_bounds__(256) flow_upsample_kernel(LevelGeom g, int f0, float* out, int w_org, int h_org,
                                                            int crop_x, int crop_y) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
  if (X >= w_org || Y >= h_org) return;
  const int fr = blockIdx.z;
  const float* fl = g.flow + (size_t)frame_of(g, f0, fr) * g.flow_frame_stride;
  float* o = out + ((size_t)fr * h_org * w_org + (size_t)Y * w_org + X) * NOP;
  const int s = 1 << g.level;
  if (s == 1) {
    const float* q = fl + ((size_t)(Y + crop_y) * g.w + (X + crop_x)) * NOP;
    for (int c = 0; c < NOP; ++c) o[c] = q[c];
    return;
  }
  const float fs = (float)s;
  auto tap = [fs](int d, int n, int& i0, int& i1, float& f) {
    const float x = ((float)d + 0.5f) / fs - 0.5f;
    const float xf = floorf(x);
    const int x0 = (int)xf;
    f = x0 < 0 ? 0.f : x - xf;
    i0 = clampi(x0, n);
    i1 = clampi(x0 + 1, n);
  };
  int xa, xb, ya, yb;
  float fx, fy;
  tap(X + crop_x, g.w, xa, xb, fx);
  tap(Y + crop_y, g.h, ya, yb, fy);
  const float gx = 1.0f - fx, gy = 1.0f - fy;
  for (int c = 0; c < NOP; ++c) {
    const float a00 = fl[((size_t)ya * g.w + xa) * NOP + c] * fs, a01 = fl[((size_t)ya * g.w + xb) * NOP + c] * fs;
    const float a10 = fl[((size_t)yb * g.w + xa) * NOP + c] * fs, a11 = fl[((size_t)yb * g.w + xb) * NOP + c] * fs;
    const float r0 = a00 * gx + a01 * fx, r1 = a10 * gx + a11 * fx;
    o[c] = r0 * gy + r1 * fy;
  }
}

}  // namespace

static dim3 padded_grid(const LevelGeom& g, int nz) { return dim3((g.tmp_w + 31) / 32, (g.tmp_h + 7) / 8, nz); }

int launch_sobel(const LevelGeom& g, int f0, int f1, cudaStream_t st) {
  sobel_kernel<<<padded_grid(g, f1 - f0), dim3(32, 8), 0, st>>>(g, f0);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_pyr_from_u8(const LevelGeom& g, int f0, int f1, const PyrSourceU8& s, cudaStream_t st) {
  pyr_from_u8_kernel<<<padded_grid(g, 2 * (f1 - f0)), dim3(32, 8), 0, st>>>(g, f0, s);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_pyr_from_level(const LevelGeom& g, int f0, int f1, const float* stage, cudaStream_t st) {
  pyr_from_level_kernel<<<padded_grid(g, 2 * (f1 - f0)), dim3(32, 8), 0, st>>>(g, f0, stage);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_pyr_down(const LevelGeom& gs, const LevelGeom& gd, int f0, int f1, cudaStream_t st) {
  pyr_down_kernel<<<padded_grid(gd, 2 * (f1 - f0)), dim3(32, 8), 0, st>>>(gs, gd, f0);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_flow_upsample(const LevelGeom& g, int f0, int f1, float* out, int w_org, int h_org, int crop_x, int crop_y,
                         cudaStream_t st) {
  const dim3 block(32, 8), grid((w_org + 31) / 32, (h_org + 7) / 8, f1 - f0);
  if (g.nop == 2) flow_upsample_kernel<2><<<grid, block, 0, st>>>(g, f0, out, w_org, h_org, crop_x, crop_y);
  else flow_upsample_kernel<1><<<grid, block, 0, st>>>(g, f0, out, w_org, h_org, crop_x, crop_y);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ofdis

// run_OF_INT / run_OF_RGB / run_DE_INT / run_DE_RGB -- command-line drop-in for the
// reference binaries (argument grammar and TIME lines of /root/reference/run_dense.cpp:185-431,
// README.md:48-88), built on the B200 hot path through OFC::OFClass (ofdis_host.h).
//
//   run_*_* image1 image2 outputfile [oppoint | p1 .. p20]
//
// No OpenCV: images are read by a small built-in decoder (binary PGM/PPM, and
// 8-bit non-interlaced PNG through zlib); pyramid, Sobel/8 gradients, padding,
// x2^lv_l upsampling, crop and .flo/.pfm writing restate run_dense.cpp:130-178,
// 298-344,384-421 (exact for 8-bit input, see of_dis_b200/preprocess.py).
// SELECTMODE 1 = optical flow, 2 = stereo; SELECTCHANNEL 1 = gray, 3 = RGB
// (CMakeLists.txt:25-46).
#include <sys/time.h>
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "ofdis_host.h"

#ifndef SELECTMODE
#define SELECTMODE 1
#endif
#ifndef SELECTCHANNEL
#define SELECTCHANNEL 1
#endif

using namespace std;

namespace {

struct Image8 {
  int w = 0, h = 0, c = 0;  // c channels, interleaved; colour order B,G,R like cv::imread
  vector<uint8_t> px;
};

bool read_file(const char* path, vector<uint8_t>& buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? n : 0);
  size_t got = n > 0 ? fread(buf.data(), 1, n, f) : 0;
  fclose(f);
  return (long)got == n && n > 0;
}

// ---- binary PNM -------------------------------------------------------------
bool decode_pnm(const vector<uint8_t>& b, Image8& im, vector<uint8_t>& rgb, int& ch) {
  if (b.size() < 3 || b[0] != 'P' || (b[1] != '5' && b[1] != '6')) return false;
  ch = b[1] == '5' ? 1 : 3;
  size_t pos = 2;
  int vals[3], nv = 0;
  while (nv < 3 && pos < b.size()) {
    while (pos < b.size() && isspace(b[pos])) ++pos;
    if (pos < b.size() && b[pos] == '#') {
      while (pos < b.size() && b[pos] != '\n') ++pos;
      continue;
    }
    int v = 0, d = 0;
    while (pos < b.size() && isdigit(b[pos])) {
      if (++d > 6) return false;  // no header number has more than 6 digits: no int overflow
      v = v * 10 + (b[pos++] - '0');
    }
    if (!d) return false;
    vals[nv++] = v;
  }
  ++pos;  // single whitespace after maxval
  if (nv < 3 || vals[2] != 255 || vals[0] <= 0 || vals[1] <= 0 || vals[0] > (1 << 15) || vals[1] > (1 << 15)) return false;
  im.w = vals[0];
  im.h = vals[1];
  const size_t n = (size_t)im.w * im.h * ch;
  if (pos + n > b.size()) return false;
  rgb.assign(b.begin() + pos, b.begin() + pos + n);
  return true;
}

// ---- PNG (8-bit, colour types 0,2,3,4,6, non-interlaced) ----------------------
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

bool decode_png(const vector<uint8_t>& b, Image8& im, vector<uint8_t>& rgb, int& ch) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (b.size() < 8 || memcmp(b.data(), sig, 8)) return false;
  size_t pos = 8;
  int depth = 0, ctype = 0, interlace = 0;
  vector<uint8_t> idat, plte;
  while (pos + 8 <= b.size()) {
    const uint32_t len = be32(&b[pos]);
    const char* type = (const char*)&b[pos + 4];
    const uint8_t* data = &b[pos + 8];
    if (pos + 12 + len > b.size()) return false;
    if (!memcmp(type, "IHDR", 4)) {
      if (len < 13) return false;
      im.w = be32(data);
      im.h = be32(data + 4);
      depth = data[8];
      ctype = data[9];
      interlace = data[12];
    } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
    else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
    else if (!memcmp(type, "IEND", 4)) break;
    pos += 12 + len;
  }
  if (depth != 8 || interlace != 0 || im.w <= 0 || im.h <= 0 || im.w > (1 << 15) || im.h > (1 << 15)) return false;
  int spp;  // samples per pixel in the file
  switch (ctype) {
    case 0: spp = 1; break;
    case 2: spp = 3; break;
    case 3: spp = 1; break;
    case 4: spp = 2; break;
    case 6: spp = 4; break;
    default: return false;
  }
  const size_t stride = (size_t)im.w * spp;
  vector<uint8_t> raw((stride + 1) * im.h);
  uLongf rawlen = raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return false;
  vector<uint8_t> img(stride * im.h), zero(stride, 0);
  for (int y = 0; y < im.h; ++y) {
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t* in = &raw[(stride + 1) * y + 1];
    uint8_t* out = &img[stride * y];
    const uint8_t* up = y ? &img[stride * (y - 1)] : zero.data();
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)spp ? out[x - spp] : 0, bb = up[x], c = x >= (size_t)spp ? up[x - spp] : 0;
      int pr = 0;
      switch (ft) {
        case 0: pr = 0; break;
        case 1: pr = a; break;
        case 2: pr = bb; break;
        case 3: pr = (a + bb) >> 1; break;
        case 4: {
          const int p = a + bb - c, pa = abs(p - a), pb = abs(p - bb), pc = abs(p - c);
          pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c);
          break;
        }
        default: return false;
      }
      out[x] = (uint8_t)(in[x] + pr);
    }
  }
  ch = (ctype == 0 || ctype == 4) ? 1 : 3;
  rgb.resize((size_t)im.w * im.h * ch);
  for (size_t i = 0; i < (size_t)im.w * im.h; ++i) {
    const uint8_t* s = &img[i * spp];
    if (ctype == 0 || ctype == 4) rgb[i] = s[0];
    else if (ctype == 3) {
      if ((size_t)s[0] * 3 + 3 > plte.size()) return false;  // palette index beyond PLTE
      const uint8_t* e = &plte[(size_t)s[0] * 3];
      rgb[i * 3] = e[0]; rgb[i * 3 + 1] = e[1]; rgb[i * 3 + 2] = e[2];
    } else {
      rgb[i * 3] = s[0]; rgb[i * 3 + 1] = s[1]; rgb[i * 3 + 2] = s[2];
    }
  }
  return true;
}

// cv::imread semantics: COLOR -> BGR; GRAYSCALE -> 1 channel.  A colour PNM goes through cvtColor's
// 14-bit BT.601 fixed point, a colour PNG through libpng's png_set_rgb_to_gray(0.299, 0.587): 15-bit
// coefficients 9797 / 19234 / 3737, truncating (checked against cv2 4.13: both formulas reproduce
// cv2.imread(..., IMREAD_GRAYSCALE) exactly on random colour images, tests/test_params_io.py).
bool load_image(const char* path, int want_channels, Image8& im) {
  vector<uint8_t> file, rgb;
  int ch = 0;
  if (!read_file(path, file)) return false;
  bool from_png = false;
  try {
    if (!decode_pnm(file, im, rgb, ch)) {
      if (!decode_png(file, im, rgb, ch)) return false;
      from_png = true;
    }
  } catch (const std::exception&) {  // bad_alloc / length errors on malformed headers
    return false;
  }
  im.c = want_channels;
  const size_t n = (size_t)im.w * im.h;
  im.px.resize(n * want_channels);
  for (size_t i = 0; i < n; ++i) {
    if (want_channels == 1) {
      if (ch == 1) im.px[i] = rgb[i];
      else {
        const int r = rgb[i * 3], g = rgb[i * 3 + 1], b = rgb[i * 3 + 2];
        im.px[i] = from_png ? (uint8_t)((r * 9797 + g * 19234 + b * 3737) >> 15)
                            : (uint8_t)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14);
      }
    } else {
      if (ch == 1) im.px[i * 3] = im.px[i * 3 + 1] = im.px[i * 3 + 2] = rgb[i];
      else { im.px[i * 3] = rgb[i * 3 + 2]; im.px[i * 3 + 1] = rgb[i * 3 + 1]; im.px[i * 3 + 2] = rgb[i * 3]; }
    }
  }
  return true;
}

// ---- float images ---------------------------------------------------------------
struct ImageF {
  int w = 0, h = 0, c = 1;
  vector<float> px;
  float& at(int x, int y, int k) { return px[((size_t)y * w + x) * c + k]; }
  float at(int x, int y, int k) const { return px[((size_t)y * w + x) * c + k]; }
};

int clampi(int v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }
int reflect101(int v, int n) {
  if (n == 1) return 0;
  while (v < 0 || v >= n) v = v < 0 ? -v : 2 * (n - 1) - v;
  return v;
}

// copyMakeBorder: replicate (image) or constant zero (gradients)
ImageF pad(const ImageF& s, int t, int b, int l, int r, bool replicate) {
  ImageF d;
  d.w = s.w + l + r;
  d.h = s.h + t + b;
  d.c = s.c;
  d.px.assign((size_t)d.w * d.h * d.c, 0.f);
  for (int y = 0; y < d.h; ++y)
    for (int x = 0; x < d.w; ++x) {
      const int sx = x - l, sy = y - t;
      if (!replicate && (sx < 0 || sy < 0 || sx >= s.w || sy >= s.h)) continue;
      for (int k = 0; k < s.c; ++k) d.at(x, y, k) = s.at(clampi(sx, s.w), clampi(sy, s.h), k);
    }
  return d;
}

// cv::resize(.5,.5,INTER_LINEAR) on even sizes (run_dense.cpp:150)
ImageF half_size(const ImageF& s) {
  ImageF d;
  d.w = s.w / 2;
  d.h = s.h / 2;
  d.c = s.c;
  d.px.resize((size_t)d.w * d.h * d.c);
  for (int y = 0; y < d.h; ++y)
    for (int x = 0; x < d.w; ++x)
      for (int k = 0; k < s.c; ++k)
        d.at(x, y, k) = ((s.at(2 * x, 2 * y, k) + s.at(2 * x + 1, 2 * y, k)) +
                         (s.at(2 * x, 2 * y + 1, k) + s.at(2 * x + 1, 2 * y + 1, k))) * 0.25f;
  return d;
}

// cv::Sobel(CV_32F, 3x3, scale 1/8, BORDER_DEFAULT) (run_dense.cpp:156-157)
void sobel8(const ImageF& s, ImageF& dx, ImageF& dy) {
  dx = s;
  dy = s;
  for (int y = 0; y < s.h; ++y)
    for (int x = 0; x < s.w; ++x)
      for (int k = 0; k < s.c; ++k) {
        const int xm = reflect101(x - 1, s.w), xp = reflect101(x + 1, s.w), ym = reflect101(y - 1, s.h),
                  yp = reflect101(y + 1, s.h);
        const float t0 = s.at(xp, ym, k) - s.at(xm, ym, k), t1 = s.at(xp, y, k) - s.at(xm, y, k),
                    t2 = s.at(xp, yp, k) - s.at(xm, yp, k);
        dx.at(x, y, k) = (t0 * 0.125f + t1 * 0.25f) + t2 * 0.125f;
        const float s0 = (s.at(xm, ym, k) * 0.125f + s.at(x, ym, k) * 0.25f) + s.at(xp, ym, k) * 0.125f;
        const float s2 = (s.at(xm, yp, k) * 0.125f + s.at(x, yp, k) * 0.25f) + s.at(xp, yp, k) * 0.125f;
        dy.at(x, y, k) = s2 - s0;
      }
}

// ConstructImgPyramide (run_dense.cpp:130-178)
void ConstructImgPyramide(const ImageF& img, vector<ImageF>& pyr, vector<ImageF>& pyr_dx, vector<ImageF>& pyr_dy,
                          const float** img_pyr, const float** dx_pyr, const float** dy_pyr, int lv_f,
                          int imgpadding) {
  pyr.resize(lv_f + 1);
  pyr_dx.resize(lv_f + 1);
  pyr_dy.resize(lv_f + 1);
  for (int i = 0; i <= lv_f; ++i) {
    pyr[i] = i == 0 ? img : half_size(pyr[i - 1]);
    sobel8(pyr[i], pyr_dx[i], pyr_dy[i]);
  }
  for (int i = 0; i <= lv_f; ++i) {
    pyr[i] = pad(pyr[i], imgpadding, imgpadding, imgpadding, imgpadding, true);
    pyr_dx[i] = pad(pyr_dx[i], imgpadding, imgpadding, imgpadding, imgpadding, false);
    pyr_dy[i] = pad(pyr_dy[i], imgpadding, imgpadding, imgpadding, imgpadding, false);
    img_pyr[i] = pyr[i].px.data();
    dx_pyr[i] = pyr_dx[i].px.data();
    dy_pyr[i] = pyr_dy[i].px.data();
  }
}

int AutoFirstScaleSelect(int imgwidth, int fratio, int patchsize) {  // run_dense.cpp:180-183
  return std::max(0, (int)std::floor(log2((2.0f * (float)imgwidth) / ((float)fratio * (float)patchsize))));
}

// cv::resize(fx=fy=s, INTER_LINEAR): src = (dst+.5)/s-.5, clamped (run_dense.cpp:410)
ImageF upsample_linear(const ImageF& s, int sc) {
  ImageF d;
  d.w = s.w * sc;
  d.h = s.h * sc;
  d.c = s.c;
  d.px.resize((size_t)d.w * d.h * d.c);
  auto taps = [&](int n_src, int n_dst, vector<int>& i0, vector<int>& i1, vector<float>& f) {
    i0.resize(n_dst); i1.resize(n_dst); f.resize(n_dst);
    for (int x = 0; x < n_dst; ++x) {
      const float fx = ((float)x + 0.5f) / (float)sc - 0.5f;
      const int x0 = (int)floorf(fx);
      f[x] = x0 < 0 ? 0.f : fx - (float)x0;
      i0[x] = clampi(x0, n_src);
      i1[x] = clampi(x0 + 1, n_src);
    }
  };
  vector<int> x0, x1, y0, y1;
  vector<float> fx, fy;
  taps(s.w, d.w, x0, x1, fx);
  taps(s.h, d.h, y0, y1, fy);
  ImageF rows;
  rows.w = d.w; rows.h = s.h; rows.c = s.c;
  rows.px.resize((size_t)rows.w * rows.h * rows.c);
  for (int y = 0; y < s.h; ++y)
    for (int x = 0; x < d.w; ++x)
      for (int k = 0; k < s.c; ++k)
        rows.at(x, y, k) = s.at(x0[x], y, k) * (1.0f - fx[x]) + s.at(x1[x], y, k) * fx[x];
  for (int y = 0; y < d.h; ++y)
    for (int x = 0; x < d.w; ++x)
      for (int k = 0; k < s.c; ++k)
        d.at(x, y, k) = rows.at(x, y0[y], k) * (1.0f - fy[y]) + rows.at(x, y1[y], k) * fy[y];
  return d;
}

// SaveFlowFile (run_dense.cpp:16-57)
void SaveFlowFile(const ImageF& img, const char* filename) {
  FILE* stream = fopen(filename, "wb");
  if (stream == 0) {
    cout << "WriteFile: could not open file" << endl;
    return;
  }
  fprintf(stream, "PIEH");
  if ((int)fwrite(&img.w, sizeof(int), 1, stream) != 1 || (int)fwrite(&img.h, sizeof(int), 1, stream) != 1)
    cout << "WriteFile: problem writing header" << endl;
  if (fwrite(img.px.data(), sizeof(float), img.px.size(), stream) != img.px.size())
    cout << "WriteFile: problem writing data" << endl;
  fclose(stream);
}

// SavePFMFile (run_dense.cpp:60-81)
void SavePFMFile(const ImageF& img, const char* filename) {
  FILE* stream = fopen(filename, "wb");
  if (stream == 0) {
    cout << "WriteFile: could not open file" << endl;
    return;
  }
  fprintf(stream, "Pf\n%d %d\n%f\n", img.w, img.h, (float)-1.0f);
  for (int y = img.h - 1; y >= 0; --y)
    for (int x = 0; x < img.w; ++x) {
      float tmp = -img.at(x, y, 0);
      if ((int)fwrite(&tmp, sizeof(float), 1, stream) != 1) cout << "WriteFile: problem writing data" << endl;
    }
  fclose(stream);
}

// Parameter block of the command line: nothing, one operating-point digit, or the 20 explicit
// numbers (run_dense.cpp:219-294, README.md:66-88).
struct CliParams {
  int lv_f, lv_l, maxiter, miniter, patchsz, patnorm, costfct, tv_innerit, tv_solverit, verbosity;
  float mindprate, mindrrate, minimgerr, poverl, tv_alpha, tv_gamma, tv_delta, tv_sor;
  bool usefbcon, usetvref;
};

// The four operating points (README.md:44-64 of the reference; run_dense.cpp:236-262): patch size, overlap, how many
// levels below the automatically selected coarsest one the pyramid descends, Gauss-Newton iterations (max = min),
// variational refinement on/off.  Index 0 is unused.
struct OperatingPoint { int patchsz; float poverl; int levels_down, iters; bool usetvref; };
const OperatingPoint kOperatingPoints[5] = {
    {0, 0.f, 0, 0, false}, {8, 0.3f, 2, 16, false}, {8, 0.4f, 2, 12, true}, {12, 0.75f, 4, 16, true}, {12, 0.75f, 5, 128, true}};

void parse_cli_params(int nnum, char** num, int width_org, CliParams& P) {
  // defaults shared by all operating points
  P.mindprate = 0.05f; P.mindrrate = 0.95f; P.minimgerr = 0.0f;
  P.usefbcon = false; P.patnorm = 1; P.costfct = 0;
  P.tv_alpha = 10.0f; P.tv_gamma = 10.0f; P.tv_delta = 5.0f;
  P.tv_innerit = 1; P.tv_solverit = 3; P.tv_sor = 1.6f;
  P.verbosity = 2;
  if (nnum <= 1) {
    int op = nnum == 1 ? atoi(num[0]) : 2;
    if (op < 1 || op > 4) op = 2;  // anything else selects operating point 2, like the reference's default branch
    const OperatingPoint& o = kOperatingPoints[op];
    P.patchsz = o.patchsz;
    P.poverl = o.poverl;
    P.lv_f = AutoFirstScaleSelect(width_org, 5, o.patchsz);
    P.lv_l = std::max(P.lv_f - o.levels_down, 0);
    P.maxiter = P.miniter = o.iters;
    P.usetvref = o.usetvref;
    return;
  }
  // the 20 explicit numbers, in the order of the reference's README (README.md:66-88)
  const struct { char kind; void* dst; } fields[20] = {
      {'i', &P.lv_f}, {'i', &P.lv_l}, {'i', &P.maxiter}, {'i', &P.miniter}, {'f', &P.mindprate}, {'f', &P.mindrrate},
      {'f', &P.minimgerr}, {'i', &P.patchsz}, {'f', &P.poverl}, {'b', &P.usefbcon}, {'i', &P.patnorm}, {'i', &P.costfct},
      {'b', &P.usetvref}, {'f', &P.tv_alpha}, {'f', &P.tv_gamma}, {'f', &P.tv_delta}, {'i', &P.tv_innerit},
      {'i', &P.tv_solverit}, {'f', &P.tv_sor}, {'i', &P.verbosity}};
  for (int k = 0; k < 20; ++k) {
    if (fields[k].kind == 'i') *static_cast<int*>(fields[k].dst) = atoi(num[k]);
    else if (fields[k].kind == 'f') *static_cast<float*>(fields[k].dst) = (float)atof(num[k]);
    else *static_cast<bool*>(fields[k].dst) = atoi(num[k]) != 0;
  }
}

double elapsed_ms(timeval& a) {
  timeval b;
  gettimeofday(&b, NULL);
  double tt = (b.tv_sec - a.tv_sec) * 1000.0f + (b.tv_usec - a.tv_usec) / 1000.0f;
  a = b;
  return tt;
}

}  // namespace

#ifdef OFDIS_IMGDUMP
// Test tool (no GPU): ofdis_imgdump <in.png|pgm|ppm> <gray|color> <out.pnm> -- what load_image() hands to
// the pipeline, so that tests can compare the decoders with cv2.imread (tests/test_params_io.py).
int main(int argc, char** argv) {
  if (argc != 4) return 2;
  Image8 im;
  const int want = !strcmp(argv[2], "gray") ? 1 : 3;
  if (!load_image(argv[1], want, im)) return 1;
  FILE* f = fopen(argv[3], "wb");
  if (!f) return 3;
  fprintf(f, "P%d\n%d %d\n255\n", want == 1 ? 5 : 6, im.w, im.h);
  fwrite(im.px.data(), 1, im.px.size(), f);  // colour: BGR order, as cv::imread returns it
  fclose(f);
  return 0;
}
#elif defined(OFDIS_BATCH)
// Batch front-end: many pairs per launch through the C-ABI's frame dimension.
//
//   run_*_*_batch listfile [--batch N] [oppoint | p1 .. p20]
//
// listfile: one "image1 image2 outputfile" triple per line.  Consecutive pairs of the same size are
// grouped into batches of up to N (default 64): 8-bit frames up, pyramid / hot path / upsampling
// on the device, full-resolution flows back.  Every output is byte-identical to what the
// single-pair binary writes for that pair.
int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s listfile [--batch N] [oppoint | 20 parameters (README.md:66-88)]\n", argv[0]);
    return 2;
  }
  int maxb = 64, first_num = 2;
  if (argc >= 4 && !strcmp(argv[2], "--batch")) {
    maxb = atoi(argv[3]);
    first_num = 4;
  }
  const int nnum = argc - first_num;
  if (maxb < 1 || (nnum > 1 && nnum != 20)) {
    fprintf(stderr, "error: expected 0, 1 or exactly 20 numbers, got %d\n", nnum);
    return 2;
  }
  struct Job { string a, b, out; };
  vector<Job> jobs;
  {
    FILE* f = fopen(argv[1], "r");
    if (!f) {
      fprintf(stderr, "error: cannot read %s\n", argv[1]);
      return 1;
    }
    char a[4096], b[4096], o[4096];
    while (fscanf(f, "%4095s %4095s %4095s", a, b, o) == 3) jobs.push_back({a, b, o});
    fclose(f);
  }
  const int nochannels = (SELECTCHANNEL == 3) ? 3 : 1;
  const int nop = (SELECTMODE == 1) ? 2 : 1;
  timeval tv;
  gettimeofday(&tv, NULL);
  size_t done = 0;
  ofdis_ctx* ctx = nullptr;
  int ctx_w = -1, ctx_h = -1, verbosity = 0;
  vector<uint8_t> frames;
  vector<float> flows;
  size_t j0 = 0;
  while (j0 < jobs.size()) {
    // load up to maxb pairs of one size
    Image8 a8, b8;
    int w = 0, h = 0, n = 0;
    frames.clear();
    while (j0 + n < jobs.size() && n < maxb) {
      const Job& jb = jobs[j0 + n];
      if (!load_image(jb.a.c_str(), nochannels, a8) || !load_image(jb.b.c_str(), nochannels, b8) || a8.w != b8.w ||
          a8.h != b8.h) {
        fprintf(stderr, "error: cannot read the pair %s %s (binary PGM/PPM or 8-bit PNG of equal size)\n",
                jb.a.c_str(), jb.b.c_str());
        if (ctx) ofdis_destroy(ctx);
        return 1;
      }
      if (n == 0) { w = a8.w; h = a8.h; }
      else if (a8.w != w || a8.h != h) break;  // next group
      frames.insert(frames.end(), a8.px.begin(), a8.px.end());
      frames.insert(frames.end(), b8.px.begin(), b8.px.end());
      ++n;
    }
    CliParams P;
    parse_cli_params(nnum, argv + first_num, w, P);
    verbosity = P.verbosity;
    if (w != ctx_w || h != ctx_h) {
      if (ctx) ofdis_destroy(ctx);
      ctx = nullptr;
      ofdis_params p;
      memset(&p, 0, sizeof(p));
      p.sc_f = P.lv_f; p.sc_l = P.lv_l; p.max_iter = P.maxiter; p.min_iter = P.miniter;
      p.dp_thresh = P.mindprate; p.dr_thresh = P.mindrrate; p.res_thresh = P.minimgerr;
      p.p_samp_s = P.patchsz; p.patove = P.poverl; p.usefbcon = P.usefbcon ? 1 : 0; p.costfct = P.costfct;
      p.noc = nochannels; p.patnorm = P.patnorm; p.usetvref = P.usetvref ? 1 : 0;
      p.tv_alpha = P.tv_alpha; p.tv_gamma = P.tv_gamma; p.tv_delta = P.tv_delta;
      p.tv_innerit = P.tv_innerit; p.tv_solverit = P.tv_solverit; p.tv_sor = P.tv_sor; p.verbosity = P.verbosity;
      const int scf = 1 << P.lv_f;
      const int rc = ofdis_create(&ctx, 0, nullptr, &p, nop, (w + scf - 1) / scf * scf, (h + scf - 1) / scf * scf,
                                  P.patchsz, maxb);
      if (rc != OFDIS_OK) {
        fprintf(stderr, "error: ofdis_create failed with status %d for %dx%d frames\n", rc, w, h);
        return 1;
      }
      ofdis_set_graph_mode(ctx, 1);
      ctx_w = w;
      ctx_h = h;
    }
    flows.resize((size_t)n * w * h * nop);
    int rc = ofdis_upload_frames_u8(ctx, 0, n, frames.data(), w, h, OFDIS_MEM_HOST);
    if (rc == OFDIS_OK) rc = ofdis_run(ctx, n, 0);
    if (rc == OFDIS_OK) rc = ofdis_get_flow_fullres(ctx, 0, n, flows.data(), w, h, OFDIS_MEM_HOST);
    if (rc == OFDIS_OK) rc = ofdis_sync(ctx);
    if (rc != OFDIS_OK) {
      fprintf(stderr, "error: %s\n", ofdis_last_error(ctx));
      ofdis_destroy(ctx);
      return 1;
    }
    ImageF out;
    out.w = w; out.h = h; out.c = nop;
    for (int k = 0; k < n; ++k) {
      out.px.assign(flows.begin() + (size_t)k * w * h * nop, flows.begin() + (size_t)(k + 1) * w * h * nop);
      if (SELECTMODE == 1) SaveFlowFile(out, jobs[j0 + k].out.c_str());
      else SavePFMFile(out, jobs[j0 + k].out.c_str());
    }
    j0 += n;
    done += n;
  }
  if (ctx) ofdis_destroy(ctx);
  if (verbosity > 0) printf("TIME (%zu pairs, load + flow + save) (ms): %3g\n", done, elapsed_ms(tv));
  return 0;
}
#else
int main(int argc, char** argv) {
  timeval tv;
  gettimeofday(&tv, NULL);
  if (argc < 4) {
    fprintf(stderr, "usage: %s image1 image2 outputfile [oppoint | 20 parameters (README.md:66-88)]\n", argv[0]);
    return 2;
  }
  if (argc > 5 && argc != 24) {
    fprintf(stderr, "error: expected 0, 1 or exactly 20 numbers after the three paths, got %d\n", argc - 4);
    return 2;
  }
  const char *imgfile_ao = argv[1], *imgfile_bo = argv[2], *outfile = argv[3];
  const int nochannels = (SELECTCHANNEL == 3) ? 3 : 1;
  const int nop = (SELECTMODE == 1) ? 2 : 1;
  Image8 a8, b8;
  if (!load_image(imgfile_ao, nochannels, a8) || !load_image(imgfile_bo, nochannels, b8)) {
    fprintf(stderr, "error: cannot read input images (supported: binary PGM/PPM, 8-bit non-interlaced PNG)\n");
    return 1;
  }
  if (a8.w != b8.w || a8.h != b8.h) {
    fprintf(stderr, "error: image sizes differ\n");
    return 1;
  }
  const int width_org = a8.w, height_org = a8.h;

  // *** parameters (run_dense.cpp:219-294)
  CliParams P;
  parse_cli_params(argc - 4, argv + 4, width_org, P);
  const int lv_f = P.lv_f, lv_l = P.lv_l, maxiter = P.maxiter, miniter = P.miniter, patchsz = P.patchsz,
            patnorm = P.patnorm, costfct = P.costfct, tv_innerit = P.tv_innerit, tv_solverit = P.tv_solverit,
            verbosity = P.verbosity;
  const float mindprate = P.mindprate, mindrrate = P.mindrrate, minimgerr = P.minimgerr, poverl = P.poverl,
              tv_alpha = P.tv_alpha, tv_gamma = P.tv_gamma, tv_delta = P.tv_delta, tv_sor = P.tv_sor;
  const bool usefbcon = P.usefbcon, usetvref = P.usetvref;

  // *** pad so that width/height are divisible by 2^lv_f (run_dense.cpp:298-311)
  int padw = 0, padh = 0;
  const int scfct = (int)pow(2, lv_f);
  int div = width_org % scfct;
  if (div > 0) padw = scfct - div;
  div = height_org % scfct;
  if (div > 0) padh = scfct - div;
  auto to_float = [&](const Image8& s) {
    ImageF f;
    f.w = s.w; f.h = s.h; f.c = s.c;
    f.px.resize(s.px.size());
    for (size_t i = 0; i < s.px.size(); ++i) f.px[i] = (float)s.px[i];
    return pad(f, (int)floor((float)padh / 2.0f), (int)ceil((float)padh / 2.0f), (int)floor((float)padw / 2.0f),
               (int)ceil((float)padw / 2.0f), true);
  };
  // Default: pyramid, gradients, paddings, upsampling and crop run on the device, bit-identical to
  // the host restatement below (tests/test_gpu_parity.py); OFDIS_HOST_PYRAMID=1 keeps them on the
  // host and hands OFClass the float pyramids exactly like run_dense.cpp:391-400.
  const char* hp = getenv("OFDIS_HOST_PYRAMID");
  if (!(hp && atoi(hp))) {
    if (verbosity > 1) printf("TIME (Image loading     ) (ms): %3g\n", elapsed_ms(tv));
    ImageF out;
    out.w = width_org;
    out.h = height_org;
    out.c = nop;
    out.px.assign((size_t)out.w * out.h * nop, 0.f);
    if (verbosity > 1) printf("TIME (Pyramide+Gradients) (ms): %3g\n", elapsed_ms(tv));  // inside the run below
    try {
      OFC::OFClass ofc(a8.px.data(), b8.px.data(), width_org, height_org, out.px.data(), nullptr, lv_f, lv_l, maxiter,
                       miniter, mindprate, mindrrate, minimgerr, patchsz, poverl, usefbcon, costfct, nochannels,
                       patnorm, usetvref, tv_alpha, tv_gamma, tv_delta, tv_innerit, tv_solverit, tv_sor, verbosity, nop);
    } catch (const std::exception& e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
    if (verbosity > 1) gettimeofday(&tv, NULL);
    if (SELECTMODE == 1) SaveFlowFile(out, outfile);
    else SavePFMFile(out, outfile);
    if (verbosity > 1) printf("TIME (Saving flow file  ) (ms): %3g\n", elapsed_ms(tv));
    return 0;
  }
  ImageF img_ao_fmat = to_float(a8), img_bo_fmat = to_float(b8);
  const int szw = img_ao_fmat.w, szh = img_ao_fmat.h;
  if (verbosity > 1) printf("TIME (Image loading     ) (ms): %3g\n", elapsed_ms(tv));

  // *** pyramids (run_dense.cpp:325-344)
  vector<const float*> img_ao_pyr(lv_f + 1), img_bo_pyr(lv_f + 1), img_ao_dx_pyr(lv_f + 1), img_ao_dy_pyr(lv_f + 1),
      img_bo_dx_pyr(lv_f + 1), img_bo_dy_pyr(lv_f + 1);
  vector<ImageF> pa, pax, pay, pb, pbx, pby;
  ConstructImgPyramide(img_ao_fmat, pa, pax, pay, img_ao_pyr.data(), img_ao_dx_pyr.data(), img_ao_dy_pyr.data(), lv_f, patchsz);
  ConstructImgPyramide(img_bo_fmat, pb, pbx, pby, img_bo_pyr.data(), img_bo_dx_pyr.data(), img_bo_dy_pyr.data(), lv_f, patchsz);
  if (verbosity > 1) printf("TIME (Pyramide+Gradients) (ms): %3g\n", elapsed_ms(tv));

  // *** main algorithm (run_dense.cpp:383-400)
  const int sc_fct = (int)pow(2, lv_l);
  ImageF flowout;
  flowout.w = szw / sc_fct;
  flowout.h = szh / sc_fct;
  flowout.c = nop;
  flowout.px.assign((size_t)flowout.w * flowout.h * nop, 0.f);
  try {
    OFC::OFClass ofc(img_ao_pyr.data(), img_ao_dx_pyr.data(), img_ao_dy_pyr.data(), img_bo_pyr.data(),
                     img_bo_dx_pyr.data(), img_bo_dy_pyr.data(), patchsz, flowout.px.data(), nullptr, szw, szh, lv_f,
                     lv_l, maxiter, miniter, mindprate, mindrrate, minimgerr, patchsz, poverl, usefbcon, costfct,
                     nochannels, patnorm, usetvref, tv_alpha, tv_gamma, tv_delta, tv_innerit, tv_solverit, tv_sor,
                     verbosity, nop);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  if (verbosity > 1) gettimeofday(&tv, NULL);

  // *** resize to original scale, crop, save (run_dense.cpp:406-421)
  if (lv_l != 0) {
    for (float& v : flowout.px) v = v * (float)sc_fct;
    flowout = upsample_linear(flowout, sc_fct);
  }
  ImageF out;
  out.w = width_org;
  out.h = height_org;
  out.c = nop;
  out.px.resize((size_t)out.w * out.h * nop);
  const int ox = (int)floor((float)padw / 2.0f), oy = (int)floor((float)padh / 2.0f);
  for (int y = 0; y < out.h; ++y)
    for (int x = 0; x < out.w; ++x)
      for (int k = 0; k < nop; ++k) out.at(x, y, k) = flowout.at(x + ox, y + oy, k);
  if (SELECTMODE == 1) SaveFlowFile(out, outfile);
  else SavePFMFile(out, outfile);
  if (verbosity > 1) printf("TIME (Saving flow file  ) (ms): %3g\n", elapsed_ms(tv));
  return 0;
}
#endif  // OFDIS_BATCH

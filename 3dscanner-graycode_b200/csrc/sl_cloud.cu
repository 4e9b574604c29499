// Stage 2: pair extraction, undistortion, law-of-sines triangulation and ordered compaction.
//
// The reference emits points x-major (outer loop over camera x, inner over y:
// scanner/triangulation/triangulate.py:52-53).  To reproduce that order deterministically and
// without atomics the image is cut into bands of 32 rows:
//   K3a count_kernel       valid pixels per (scan, band, column)
//   K3b bandprefix_kernel  exclusive prefix down each column + column totals
//   K3c colscan_kernel     exclusive scan of the column totals in (scan, x) order -> packed bases
//   K2  emit_kernel        one warp per column of a 32-row band (lane = row): warp ballot gives the
//                          rank inside the band, so a warp writes a contiguous run of points.
// K2 stages its tile through shared memory so the global reads stay row-major / coalesced.
#include <cmath>
#include "sl_common.cuh"
#include "sl_pixel.cuh"
#include "sl_kernels.h"

namespace sl {

constexpr int kBandRows = 32;
constexpr int kTileCols = 64;
constexpr int kEmitThreads = 256;

// ------------------------------------------------------------------ workspace
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t cloud_workspace_bytes(int B, int H, int W) {
  const size_t bands = (size_t)(H + kBandRows - 1) / kBandRows;
  return align_up((size_t)B * bands * W * 4, 256) + align_up((size_t)B * W * 4, 256) +
         align_up((size_t)B * W * 8, 256);
}

CloudWorkspace carve_workspace(void* ws, int B, int H, int W) {
  CloudWorkspace c;
  c.bands = (H + kBandRows - 1) / kBandRows;
  uint8_t* p = static_cast<uint8_t*>(ws);
  c.band_cnt = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)B * c.bands * W * 4, 256);
  c.col_tot = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)B * W * 4, 256);
  c.col_base = reinterpret_cast<long long*>(p);
  return c;
}

// ------------------------------------------------------------------ K3a..c offsets
__global__ void __launch_bounds__(128) count_kernel(const int16_t* __restrict__ h,
                                                     const int16_t* __restrict__ v, int H, int W,
                                                     uint32_t* __restrict__ band_cnt) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int band = blockIdx.y, b = blockIdx.z, bands = gridDim.y;
  if (x >= W) return;
  const size_t img = (size_t)b * H * W;
  const int y0 = band * kBandRows;
  const int y1 = min(H, y0 + kBandRows);
  uint32_t c = 0;
#pragma unroll 8
  for (int y = y0; y < y1; ++y) {
    const size_t i = img + (size_t)y * W + x;
    c += (h[i] != -1 && v[i] != -1) ? 1u : 0u;     // triangulate.py:56
  }
  band_cnt[((size_t)b * bands + band) * W + x] = c;
}

__global__ void __launch_bounds__(256) bandprefix_kernel(uint32_t* __restrict__ band_cnt,
                                                          uint32_t* __restrict__ col_tot, int B,
                                                          int bands, int W) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * W) return;
  const int b = (int)(i / W), x = (int)(i - (long long)b * W);
  uint32_t run = 0;
  for (int k = 0; k < bands; ++k) {
    uint32_t* p = band_cnt + ((size_t)b * bands + k) * W + x;
    const uint32_t c = *p;
    *p = run;
    run += c;
  }
  col_tot[i] = run;
}

// single block: exclusive scan of m uint32 values into int64, plus per-scan totals
__global__ void __launch_bounds__(1024) colscan_kernel(const uint32_t* __restrict__ vals,
                                                        long long* __restrict__ out, long long m,
                                                        int B, int W, long long* __restrict__ counts) {
  __shared__ long long warp_sum[32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < m; base += 1024) {
    const long long i = base + tid;
    const long long x = i < m ? (long long)vals[i] : 0;
    long long inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sum[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long before = carry + (warp ? warp_sum[warp - 1] : 0) + inc - x;
    if (i < m) out[i] = before;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (counts) {
    const long long total = carry_s;
    for (int b = tid; b < B; b += 1024) {
      const long long lo = out[(long long)b * W];
      const long long hi = (b + 1 < B) ? out[(long long)(b + 1) * W] : total;
      counts[b] = hi - lo;
    }
    if (tid == 0) counts[B] = total;
  }
}

// ------------------------------------------------------------------ camera models
template <class T>
struct Lens {
  T fx, fy, cx, cy, ifx, ify;
  T k[8];
  T R[9];
};

struct CalibDev {
  Lens<double> cam_d, proj_d;
  Lens<float> cam_f, proj_f;
  double T[3], t_len;
  float Tf[3], t_len2f;
  int proj_w, proj_h;
};

template <class T>
static Lens<T> make_lens(const double* K, const double* dist, const double* R) {
  Lens<T> l;
  l.fx = (T)K[0]; l.fy = (T)K[4]; l.cx = (T)K[2]; l.cy = (T)K[5];
  l.ifx = (T)(1.0 / K[0]); l.ify = (T)(1.0 / K[4]);
  for (int i = 0; i < 8; ++i) l.k[i] = (T)dist[i];
  static const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) l.R[i] = (T)(R ? R[i] : eye[i]);
  return l;
}

static CalibDev make_calib(const SlCalib& c) {
  CalibDev d;
  d.cam_d = make_lens<double>(c.cam_K, c.cam_dist, c.R);
  d.proj_d = make_lens<double>(c.proj_K, c.proj_dist, nullptr);
  d.cam_f = make_lens<float>(c.cam_K, c.cam_dist, c.R);
  d.proj_f = make_lens<float>(c.proj_K, c.proj_dist, nullptr);
  double s = 0;
  for (int i = 0; i < 3; ++i) { d.T[i] = c.T[i]; d.Tf[i] = (float)c.T[i]; s += c.T[i] * c.T[i]; }
  d.t_len = sqrt(s);
  d.t_len2f = (float)s;
  d.proj_w = c.proj_w; d.proj_h = c.proj_h;
  return d;
}

// cv2.undistortPoints (default criteria: exactly 5 fixed-point iterations), fp64, one rounding
// per operation like the compiled OpenCV loop the reference calls at triangulate.py:84-85.
__device__ __forceinline__ void undistort_fp64(float uf, float vf, const Lens<double>& c, float& ox, float& oy) {
  const double x0 = dmul(dsub((double)uf, c.cx), c.ifx);
  const double y0 = dmul(dsub((double)vf, c.cy), c.ify);
  double x = x0, y = y0;
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    const double r2 = dadd(dmul(x, x), dmul(y, y));
    const double num = dadd(1.0, dmul(dadd(dmul(dadd(dmul(c.k[7], r2), c.k[6]), r2), c.k[5]), r2));
    const double den = dadd(1.0, dmul(dadd(dmul(dadd(dmul(c.k[4], r2), c.k[1]), r2), c.k[0]), r2));
    const double icdist = ddiv(num, den);
    if (icdist < 0) { x = x0; y = y0; break; }
    const double dx = dadd(dmul(dmul(dmul(2.0, c.k[2]), x), y), dmul(c.k[3], dadd(r2, dmul(dmul(2.0, x), x))));
    const double dy = dadd(dmul(c.k[2], dadd(r2, dmul(dmul(2.0, y), y))), dmul(dmul(dmul(2.0, c.k[3]), x), y));
    x = dmul(dsub(x0, dx), icdist);
    y = dmul(dsub(y0, dy), icdist);
  }
  const double xx = dadd(dadd(dmul(c.R[0], x), dmul(c.R[1], y)), c.R[2]);
  const double yy = dadd(dadd(dmul(c.R[3], x), dmul(c.R[4], y)), c.R[5]);
  const double ww = ddiv(1.0, dadd(dadd(dmul(c.R[6], x), dmul(c.R[7], y)), c.R[8]));
  ox = (float)dmul(xx, ww);
  oy = (float)dmul(yy, ww);
}

// Same iteration in fp32 with FMA contraction allowed (throughput mode).
__device__ __forceinline__ void undistort_fp32(float u, float v, const Lens<float>& c, float& ox, float& oy) {
  const float x0 = (u - c.cx) * c.ifx, y0 = (v - c.cy) * c.ify;
  float x = x0, y = y0;
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    const float r2 = x * x + y * y;
    const float num = 1.0f + ((c.k[7] * r2 + c.k[6]) * r2 + c.k[5]) * r2;
    const float den = 1.0f + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2;
    const float icdist = __fdividef(num, den);
    if (icdist < 0.0f) { x = x0; y = y0; break; }
    const float dx = 2.0f * c.k[2] * x * y + c.k[3] * (r2 + 2.0f * x * x);
    const float dy = c.k[2] * (r2 + 2.0f * y * y) + 2.0f * c.k[3] * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  const float xx = c.R[0] * x + c.R[1] * y + c.R[2];
  const float yy = c.R[3] * x + c.R[4] * y + c.R[5];
  const float ww = __fdividef(1.0f, c.R[6] * x + c.R[7] * y + c.R[8]);
  ox = xx * ww;
  oy = yy * ww;
}

__device__ __forceinline__ void rays_to_point_fp64(float cx, float cy, float px, float py, const CalibDev& c,
                                                   double& X, double& Y, double& Z);
__device__ __forceinline__ void rays_to_point_fp32(float cx, float cy, float px, float py, const CalibDev& c,
                                                   float& X, float& Y, float& Z);

// triangulate.py:86-95 with the reference's dtype flow: rays f32, camera ray normalised in f32,
// angles / lengths f64.
__device__ __forceinline__ void triangulate_fp64(float u, float v, float pu, float pv, const CalibDev& c,
                                                 double& X, double& Y, double& Z) {
  float cx, cy, px, py;
  undistort_fp64(u, v, c.cam_d, cx, cy);
  undistort_fp64(pu, pv, c.proj_d, px, py);
  rays_to_point_fp64(cx, cy, px, py, c, X, Y, Z);
}

// fp32: same triangle, but sin/cos of the angles taken from cross/dot products
// (sin(gamma) = sin(alpha+beta)); no arccos.  With a = T x c, b = T x p:
//   P = c * |T|^2 |b| / (|a| (T.p) - (T.c) |b|)
__device__ __forceinline__ void triangulate_fp32(float u, float v, float pu, float pv, const CalibDev& c,
                                                 float& X, float& Y, float& Z) {
  float cx, cy, px, py;
  undistort_fp32(u, v, c.cam_f, cx, cy);
  undistort_fp32(pu, pv, c.proj_f, px, py);
  rays_to_point_fp32(cx, cy, px, py, c, X, Y, Z);
}

// Same two formulas starting from already-undistorted rays (lens tables).
__device__ __forceinline__ void rays_to_point_fp64(float cx, float cy, float px, float py, const CalibDev& c,
                                                   double& X, double& Y, double& Z) {
  const float cn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), 1.0f));
  const float lx = __fdiv_rn(cx, cn), ly = __fdiv_rn(cy, cn), lz = __fdiv_rn(1.0f, cn);
  const double d_tl = dadd(dadd(dmul(c.T[0], (double)lx), dmul(c.T[1], (double)ly)), dmul(c.T[2], (double)lz));
  const double alpha = acos(ddiv(-d_tl, c.t_len));
  const float pn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), 1.0f));
  const double d_tp = dadd(dadd(dmul(c.T[0], (double)px), dmul(c.T[1], (double)py)), c.T[2]);
  const double beta = acos(ddiv(d_tp, dmul(c.t_len, (double)pn)));
  const double gamma = dsub(dsub(3.141592653589793, alpha), beta);
  const double p_len = ddiv(dmul(c.t_len, sin(beta)), sin(gamma));
  X = dmul((double)lx, p_len);
  Y = dmul((double)ly, p_len);
  Z = dmul((double)lz, p_len);
}

__device__ __forceinline__ void rays_to_point_fp32(float cx, float cy, float px, float py, const CalibDev& c,
                                                   float& X, float& Y, float& Z) {
  const float t0 = c.Tf[0], t1 = c.Tf[1], t2 = c.Tf[2];
  const float ax = t1 - t2 * cy, ay = t2 * cx - t0, az = t0 * cy - t1 * cx;     // T x (cx,cy,1)
  const float bx = t1 - t2 * py, by = t2 * px - t0, bz = t0 * py - t1 * px;     // T x (px,py,1)
  const float an = sqrtf(ax * ax + ay * ay + az * az);
  const float bn = sqrtf(bx * bx + by * by + bz * bz);
  const float tc = t0 * cx + t1 * cy + t2;
  const float tp = t0 * px + t1 * py + t2;
  const float s = c.t_len2f * bn / (an * tp - tc * bn);
  X = cx * s; Y = cy * s; Z = s;
}

// Lens table: the undistorted, normalised ray of every lattice point (u, v) in [0,w) x [0,h),
// computed ONCE per calibration with the exact fp64 iteration (bit-identical to
// cv2.undistortPoints, triangulate.py:84-85).  The camera pixel grid and the projector pixel grid
// never change between scans, so the per-scan kernels look the rays up instead of iterating.
__global__ void __launch_bounds__(256) lens_table_kernel(const Lens<double> lens, int w, int h,
                                                         float2* __restrict__ table) {
  const long long n = (long long)w * h;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (long long)y * w);
    float ox, oy;
    undistort_fp64((float)x, (float)y, lens, ox, oy);
    table[i] = make_float2(ox, oy);
  }
}

// ------------------------------------------------------------------ K2 emit
enum { kModePairs = 0, kModeCloud32 = 1, kModeCloud64 = 2, kModeTable32 = 3, kModeTable64 = 4 };

struct EmitArgs {
  const int16_t* h;
  const int16_t* v;
  const uint8_t* white;     // nullable
  const uint32_t* band_pre;
  const long long* col_base;
  float* out_a;             // pairs: cam_pts (N,2)   cloud: xyz (N,3)
  float* out_b;             // pairs: proj_pts (N,2)
  uint8_t* rgb;             // nullable
  const float2* cam_rays;   // table modes: [H][W] undistorted camera rays (rotated by R)
  const float2* proj_rays;  // table modes: [proj_h][proj_w] undistorted projector rays
  int H, W, bands;
  CalibDev cal;
};

constexpr int kIdxPitch = kTileCols + 2;          // int16 elements; 33 words -> conflict-free column reads
constexpr int kRgbPitch = kTileCols * 3 + 4;      // bytes; 49 words

template <int MODE>
__global__ void __launch_bounds__(kEmitThreads) emit_kernel(const EmitArgs a) {
  __shared__ int16_t sh[kBandRows * kIdxPitch];
  __shared__ int16_t sv[kBandRows * kIdxPitch];
  __shared__ __align__(4) uint8_t sc[kBandRows * kRgbPitch];
  const int x0 = blockIdx.x * kTileCols, band = blockIdx.y, b = blockIdx.z;
  const int y0 = band * kBandRows;
  const int rows = min(kBandRows, a.H - y0), cols = min(kTileCols, a.W - x0);
  const size_t img = (size_t)b * a.H * a.W;
  const bool has_rgb = a.white != nullptr && a.rgb != nullptr;

  // row-major, coalesced staging of the tile
  for (int i = threadIdx.x; i < kBandRows * kTileCols; i += kEmitThreads) {
    const int r = i / kTileCols, cidx = i - r * kTileCols;
    int16_t hv = -1, vv = -1;
    if (r < rows && cidx < cols) {
      const size_t g = img + (size_t)(y0 + r) * a.W + x0 + cidx;
      hv = a.h[g]; vv = a.v[g];
    }
    sh[r * kIdxPitch + cidx] = hv;
    sv[r * kIdxPitch + cidx] = vv;
  }
  if (has_rgb) {
    for (int i = threadIdx.x; i < kBandRows * kTileCols * 3; i += kEmitThreads) {
      const int r = i / (kTileCols * 3), cb = i - r * (kTileCols * 3);
      uint8_t val = 0;
      if (r < rows && cb < cols * 3) val = a.white[(img + (size_t)(y0 + r) * a.W + x0) * 3 + cb];
      sc[r * kRgbPitch + cb] = val;
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cidx = warp; cidx < cols; cidx += kEmitThreads / 32) {
    const int hv = sh[lane * kIdxPitch + cidx], vv = sv[lane * kIdxPitch + cidx];
    const bool valid = (hv != -1) && (vv != -1);
    const uint32_t m = __ballot_sync(0xffffffffu, valid);
    if (m == 0) continue;
    const int x = x0 + cidx;
    const long long base = a.col_base[(size_t)b * a.W + x] +
                           (long long)a.band_pre[((size_t)b * a.bands + band) * a.W + x];
    if (!valid) continue;
    const long long o = base + __popc(m & ((1u << lane) - 1u));
    const float u = (float)x, w = (float)(y0 + lane);
    const float pu = (float)min(a.cal.proj_w - 1, hv), pv = (float)min(a.cal.proj_h - 1, vv);   // triangulate.py:60-61
    if (MODE == kModePairs) {
      reinterpret_cast<float2*>(a.out_a)[o] = make_float2(u, w);
      reinterpret_cast<float2*>(a.out_b)[o] = make_float2(pu, pv);
    } else if (MODE == kModeCloud32) {
      float X, Y, Z;
      triangulate_fp32(u, w, pu, pv, a.cal, X, Y, Z);
      a.out_a[o * 3 + 0] = X; a.out_a[o * 3 + 1] = Y; a.out_a[o * 3 + 2] = Z;
    } else if (MODE == kModeTable32 || MODE == kModeTable64) {
      const float2 c = __ldg(a.cam_rays + (size_t)(y0 + lane) * a.W + x);
      const float2 p = __ldg(a.proj_rays + (size_t)min(a.cal.proj_h - 1, vv) * a.cal.proj_w + min(a.cal.proj_w - 1, hv));
      if (MODE == kModeTable32) {
        float X, Y, Z;
        rays_to_point_fp32(c.x, c.y, p.x, p.y, a.cal, X, Y, Z);
        a.out_a[o * 3 + 0] = X; a.out_a[o * 3 + 1] = Y; a.out_a[o * 3 + 2] = Z;
      } else {
        double X, Y, Z;
        rays_to_point_fp64(c.x, c.y, p.x, p.y, a.cal, X, Y, Z);
        a.out_a[o * 3 + 0] = (float)X; a.out_a[o * 3 + 1] = (float)Y; a.out_a[o * 3 + 2] = (float)Z;
      }
    } else {
      double X, Y, Z;
      triangulate_fp64(u, w, pu, pv, a.cal, X, Y, Z);
      a.out_a[o * 3 + 0] = (float)X; a.out_a[o * 3 + 1] = (float)Y; a.out_a[o * 3 + 2] = (float)Z;
    }
    if (has_rgb) {
      const uint8_t* s = sc + lane * kRgbPitch + cidx * 3;
      a.rgb[o * 3 + 0] = s[0]; a.rgb[o * 3 + 1] = s[1]; a.rgb[o * 3 + 2] = s[2];     // triangulate.py:64
    }
  }
}

// ------------------------------------------------------------------ points API
template <bool FP64>
__global__ void __launch_bounds__(256) points_kernel(const float2* __restrict__ cam,
                                                     const float2* __restrict__ proj, long long N,
                                                     const CalibDev cal, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    const float2 c = cam[i], p = proj[i];
    double X, Y, Z;
    if (FP64) {
      triangulate_fp64(c.x, c.y, p.x, p.y, cal, X, Y, Z);
    } else {
      float x, y, z;
      triangulate_fp32(c.x, c.y, p.x, p.y, cal, x, y, z);
      X = x; Y = y; Z = z;
    }
    out[i] = X; out[N + i] = Y; out[2 * N + i] = Z;
  }
}

// ------------------------------------------------------------------ box filter (triangulate.py:99-122)
constexpr int kFilterBlock = 1024;   // points per block (256 threads x 4)

__device__ __forceinline__ bool keep_pt(const double* pts, long long N, long long i, double thr) {
  const double x = pts[i], y = pts[N + i], z = pts[2 * N + i];
  return (z < thr) && (z > -thr) && (y < thr) && (y > -thr) && (x < thr) && (x > -thr);
}

__global__ void __launch_bounds__(256) filter_count_kernel(const double* __restrict__ pts, long long N,
                                                            double thr, uint32_t* __restrict__ blk_cnt) {
  const long long base = (long long)blockIdx.x * kFilterBlock;
  uint32_t c = 0;
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < N && keep_pt(pts, N, i, thr)) ++c;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ uint32_t ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += ws[i];
    blk_cnt[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) filter_emit_kernel(const double* __restrict__ pts,
                                                           const double* __restrict__ colors, long long N,
                                                           double thr, const long long* __restrict__ blk_base,
                                                           double* __restrict__ out_pts,
                                                           double* __restrict__ out_colors) {
  __shared__ uint32_t ws[8];
  const long long base = (long long)blockIdx.x * kFilterBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long run = blk_base[blockIdx.x];
  for (int k = 0; k < 4; ++k) {          // chunks of 256 consecutive points keep the order
    const long long i = base + k * 256 + threadIdx.x;
    const bool keep = i < N && keep_pt(pts, N, i, thr);
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) ws[warp] = __popc(m);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w = 0; w < 8; ++w) { const uint32_t c = ws[w]; if (w < warp) before += c; total += c; }
    if (keep) {
      const long long o = run + before + __popc(m & ((1u << lane) - 1u));
      out_pts[o] = pts[i]; out_pts[N + o] = pts[N + i]; out_pts[2 * N + o] = pts[2 * N + i];
      if (colors) {
        out_colors[o * 3 + 0] = colors[i * 3 + 0];
        out_colors[o * 3 + 1] = colors[i * 3 + 1];
        out_colors[o * 3 + 2] = colors[i * 3 + 2];
      }
    }
    run += total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ launchers
int launch_compact_offsets(const int16_t* h, const int16_t* v, int B, int H, int W, void* ws,
                           long long* counts, cudaStream_t st) {
  const CloudWorkspace c = carve_workspace(ws, B, H, W);
  dim3 g1((W + 127) / 128, c.bands, B);
  count_kernel<<<g1, 128, 0, st>>>(h, v, H, W, c.band_cnt);
  count_launch();
  const long long m = (long long)B * W;
  bandprefix_kernel<<<(int)((m + 255) / 256), 256, 0, st>>>(c.band_cnt, c.col_tot, B, c.bands, W);
  count_launch();
  colscan_kernel<<<1, 1024, 0, st>>>(c.col_tot, c.col_base, m, B, W, counts);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

static EmitArgs make_emit(const int16_t* h, const int16_t* v, const uint8_t* white, int B, int H, int W,
                          const void* ws, float* out_a, float* out_b, uint8_t* rgb) {
  const CloudWorkspace c = carve_workspace(const_cast<void*>(ws), B, H, W);
  EmitArgs a;
  a.h = h; a.v = v; a.white = white; a.band_pre = c.band_cnt; a.col_base = c.col_base;
  a.out_a = out_a; a.out_b = out_b; a.rgb = rgb; a.H = H; a.W = W; a.bands = c.bands;
  a.cam_rays = nullptr; a.proj_rays = nullptr;
  return a;
}

int launch_emit_pairs(const int16_t* h, const int16_t* v, const uint8_t* white, int B, int H, int W,
                      int proj_w, int proj_h, const void* ws, float* cam_pts, float* proj_pts,
                      uint8_t* rgb, cudaStream_t st) {
  EmitArgs a = make_emit(h, v, white, B, H, W, ws, cam_pts, proj_pts, rgb);
  a.cal = CalibDev();
  a.cal.proj_w = proj_w; a.cal.proj_h = proj_h;
  dim3 g((W + kTileCols - 1) / kTileCols, a.bands, B);
  emit_kernel<kModePairs><<<g, kEmitThreads, 0, st>>>(a);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_lens_tables(const SlCalib& calib, int cam_w, int cam_h, float* cam_rays, float* proj_rays,
                       cudaStream_t st) {
  const CalibDev cal = make_calib(calib);
  auto blocks = [](long long n) { long long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); };
  lens_table_kernel<<<blocks((long long)cam_w * cam_h), 256, 0, st>>>(cal.cam_d, cam_w, cam_h, (float2*)cam_rays);
  count_launch();
  lens_table_kernel<<<blocks((long long)cal.proj_w * cal.proj_h), 256, 0, st>>>(cal.proj_d, cal.proj_w, cal.proj_h,
                                                                               (float2*)proj_rays);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_emit_cloud(const int16_t* h, const int16_t* v, const uint8_t* white, const SlCalib& calib,
                      int precision, const float* cam_rays, const float* proj_rays, int B, int H, int W,
                      const void* ws, float* xyz, uint8_t* rgb, cudaStream_t st) {
  EmitArgs a = make_emit(h, v, white, B, H, W, ws, xyz, nullptr, rgb);
  a.cal = make_calib(calib);
  a.cam_rays = (const float2*)cam_rays; a.proj_rays = (const float2*)proj_rays;
  dim3 g((W + kTileCols - 1) / kTileCols, a.bands, B);
  if (cam_rays) {
    if (precision == SL_PREC_FP64) emit_kernel<kModeTable64><<<g, kEmitThreads, 0, st>>>(a);
    else emit_kernel<kModeTable32><<<g, kEmitThreads, 0, st>>>(a);
  } else if (precision == SL_PREC_FP64) emit_kernel<kModeCloud64><<<g, kEmitThreads, 0, st>>>(a);
  else emit_kernel<kModeCloud32><<<g, kEmitThreads, 0, st>>>(a);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_triangulate_points(const float* cam_pts, const float* proj_pts, long long N,
                              const SlCalib& calib, int precision, double* out, cudaStream_t st) {
  if (N == 0) return kOk;
  const CalibDev cal = make_calib(calib);
  long long blocks = (N + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (precision == SL_PREC_FP64)
    points_kernel<true><<<(int)blocks, 256, 0, st>>>((const float2*)cam_pts, (const float2*)proj_pts, N, cal, out);
  else
    points_kernel<false><<<(int)blocks, 256, 0, st>>>((const float2*)cam_pts, (const float2*)proj_pts, N, cal, out);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

size_t filter_workspace_bytes(long long N) {
  const size_t blocks = (size_t)((N + kFilterBlock - 1) / kFilterBlock) + 1;
  return align_up(blocks * 4, 256) + align_up(blocks * 8, 256);
}

int launch_filter_box(const double* pts, const double* colors, long long N, double threshold,
                      double* out_pts, double* out_colors, long long* out_count, void* ws,
                      cudaStream_t st) {
  const long long blocks = (N + kFilterBlock - 1) / kFilterBlock;
  uint32_t* blk_cnt = static_cast<uint32_t*>(ws);
  long long* blk_base = reinterpret_cast<long long*>(static_cast<uint8_t*>(ws) + align_up((size_t)(blocks + 1) * 4, 256));
  if (blocks > 0) {
    filter_count_kernel<<<(int)blocks, 256, 0, st>>>(pts, N, threshold, blk_cnt);
    count_launch();
  }
  // B = 0: the scan writes only counts[0] = total kept
  colscan_kernel<<<1, 1024, 0, st>>>(blk_cnt, blk_base, blocks, 0, 1, out_count);
  count_launch();
  if (blocks > 0) {
    filter_emit_kernel<<<(int)blocks, 256, 0, st>>>(pts, colors, N, threshold, blk_base, out_pts, out_colors);
    count_launch();
  }
  SL_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace sl

// Stage 2: pair extraction, undistortion, law-of-sines triangulation and ordered compaction.
//
// The reference emits points x-major (outer loop over camera x, inner over y:
// scanner/triangulation/triangulate.py:52-53).  To reproduce that order deterministically and
// without atomics the image is cut into bands of 32 rows:
//   (K1 writes a flat validity bitmap, 1 bit per pixel; K3a builds it from caller-supplied indices)
//   K3b strip_prefix_kernel  one warp per 32-column strip: 32x32 ballot transposes of the bitmap give
//                            per-(band, column) counts, prefixed down each column + column totals
//   K3c (inside K3b)         the last strip CTA to finish scans the column totals in (scan, x) order
//                            -> packed bases + per-scan counts (colscan_kernel is the standalone form)
//   K2  emit_kernel          one warp per 32-row x 16-column tile; compute phase has lane = row, so a
//                            warp ballot gives the rank inside the band and the warp writes a
//                            contiguous run of points per column.
// K2 stages its tile through shared memory so the global reads stay row-major / coalesced.
#include <cmath>
#include "sl_common.cuh"
#include "sl_pixel.cuh"
#include "sl_kernels.h"

namespace sl {

#ifndef SL_EMIT_UNROLL
#define SL_EMIT_UNROLL 1
#endif
constexpr int kBandRows = 32;       // rows per band = lanes per warp
constexpr int kStripCols = 32;      // columns per strip in the offsets kernel
constexpr int kEmitCols = 16;       // columns per warp tile in the emit kernel
#ifndef SL_EMIT_GROUP
#define SL_EMIT_GROUP 4
#endif
constexpr int kEmitGroup = SL_EMIT_GROUP;     // row groups (of 8 rows) loaded per pass: 1, 2 or 4
#ifndef SL_EMIT_WARPS
#define SL_EMIT_WARPS 4
#endif
constexpr int kEmitWarps = SL_EMIT_WARPS;

// ------------------------------------------------------------------ workspace
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t strips_of(int W) { return (size_t)(W + 31) / 32; }      // kStripCols = 32

static size_t bitmap_words(long long npix) { return (size_t)((npix + 31) / 32) + 1; }   // +1: funnel reads

size_t cloud_workspace_bytes(int B, int H, int W) {
  const size_t bands = (size_t)(H + kBandRows - 1) / kBandRows;
  return align_up((size_t)B * bitmap_words((long long)H * W) * 4, 256) +      // validity bitmap
         align_up((size_t)B * bands * W * 4, 256) +                            // band_pre
         align_up((size_t)B * bands * W * 4, 256) +                            // col_bits
         align_up((size_t)B * W * 4, 256) +                                    // col_tot
         align_up((size_t)B * W * 8, 256) +                                    // col_base
         align_up((size_t)B * (strips_of(W) + 1) * 8, 256);                    // published strip / scan totals
}

CloudWorkspace carve_workspace(void* ws, int B, int H, int W) {
  CloudWorkspace c;
  c.bands = (H + kBandRows - 1) / kBandRows;
  c.bitmap_words = (long long)bitmap_words((long long)H * W);
  uint8_t* p = static_cast<uint8_t*>(ws);
  c.bitmap = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)B * c.bitmap_words * 4, 256);
  c.band_cnt = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)B * c.bands * W * 4, 256);
  c.col_bits = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)B * c.bands * W * 4, 256);
  c.col_tot = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)B * W * 4, 256);
  c.col_base = reinterpret_cast<long long*>(p);
  p += align_up((size_t)B * W * 8, 256);
  c.flags = reinterpret_cast<unsigned long long*>(p);
  c.n_flags = (int)((size_t)B * (strips_of(W) + 1));
  return c;
}

// ------------------------------------------------------------------ K3 offsets
// K3a (only when the indices come from the caller rather than from the decode kernel):
// validity bitmap, bit i of scan b = (h[i] != -1 && v[i] != -1)   (triangulate.py:56)
__global__ void __launch_bounds__(256) bitmap_from_idx_kernel(const int16_t* __restrict__ h,
                                                               const int16_t* __restrict__ v, long long npix,
                                                               long long words_per_scan, int B,
                                                               uint32_t* __restrict__ bitmap,
                                                               unsigned long long* __restrict__ flags, int n_flags) {
  if (blockIdx.x == 0)                                        // the offsets kernel's published totals start out empty
    for (int i = threadIdx.x; i < n_flags; i += blockDim.x) flags[i] = 0;
  const long long padded = (npix + 31) & ~31LL;
  const long long total = padded * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / padded, px = i - b * padded;
    bool ok = false;
    if (px < npix) ok = (h[b * npix + px] != -1) && (v[b * npix + px] != -1);
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) bitmap[b * words_per_scan + (px >> 5)] = m;
  }
}

// Exclusive scan of m uint32 values into int64 by ONE thread block of NT threads, plus per-scan totals
// (counts[b] = points of scan b, counts[B] = total).
template <int NT>
__device__ __forceinline__ void block_scan_columns(const uint32_t* __restrict__ vals, long long* __restrict__ out,
                                                   long long m, int B, int W, long long* __restrict__ counts) {
  __shared__ long long warp_sum[32];
  __shared__ long long carry_s;
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < m; base += NT * 4) {       // 4 consecutive values per thread and round
    const long long i0 = base + (long long)tid * 4;
    uint32_t x[4] = {0, 0, 0, 0};
    if (i0 + 3 < m) {
      const uint4 q = *reinterpret_cast<const uint4*>(vals + i0);        // vals is 256-byte aligned, i0 % 4 == 0
      x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (i0 + k < m) x[k] = vals[i0 + k];
    }
    const long long p1 = (long long)x[0], p2 = p1 + x[1], p3 = p2 + x[2], sum = p3 + x[3];
    long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      long long w = lane < NW ? warp_sum[lane] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sum[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long before = carry + (warp ? warp_sum[warp - 1] : 0) + inc - sum;
    const long long o[4] = {before, before + p1, before + p2, before + p3};
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i0 + k < m) out[i0 + k] = o[k];
    __syncthreads();
    if (tid == NT - 1) carry_s = carry + warp_sum[NW - 1];
    __syncthreads();
  }
  if (counts) {
    const long long total = carry_s;
    for (int b = tid; b < B; b += NT) {
      const long long lo = out[(long long)b * W];
      const long long hi = (b + 1 < B) ? out[(long long)(b + 1) * W] : total;
      counts[b] = hi - lo;
    }
    if (tid == 0) counts[B] = total;
  }
}

__global__ void __launch_bounds__(1024) colscan_kernel(const uint32_t* __restrict__ vals,
                                                        long long* __restrict__ out, long long m,
                                                        int B, int W, long long* __restrict__ counts) {
  block_scan_columns<1024>(vals, out, m, B, W, counts);
}

// K3b: one CTA per strip of 32 columns, one warp per band (round-robin).  Lane r fetches the 32
// validity bits of row (band*32 + r) in the strip (funnel shift out of the flat bitmap); 32 ballots
// transpose the 32x32 bit block so lane c holds the column word of column x0+c, whose popcount is
// the number of valid pixels of that column in the band.  The counts meet in shared memory, are
// prefixed down each column (one warp scan per column, lanes = bands), and leave as band_pre[b][band][x]
// (exclusive) and col_tot[b][x]; the column words themselves are kept (col_bits) -- the emit kernel ranks pixels
// with them.
// K3c, the packed base of every column (scan-major, then x-major), needs the totals of everything before the strip.
// No second kernel, no last-CTA pass: every CTA publishes its strip total (a 64-bit word: valid bit + value, release
// store), then sums the published totals of the lower strips of its scan and the published totals of the lower scans
// (each scan's last strip publishes the scan total), spinning on words that are not there yet.  A CTA only ever
// waits for CTAs with a lower block index, which the hardware dispatches first, so the waits always end; the words are
// zeroed by the kernel that produced the bitmap (decode / bitmap_from_idx), which precedes this one in stream order.
constexpr int kStripWarps = 16;
constexpr int kMaxBands = 1024;      // H <= 32768
constexpr int kCntPitch = 33;        // shared counts [band][33]: the column-wise scans read it conflict-free
constexpr unsigned long long kFlagValid = 1ull << 63;

__device__ __forceinline__ void publish_total(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v | kFlagValid) : "memory");
}
__device__ __forceinline__ unsigned long long await_total(const unsigned long long* p) {
  unsigned long long v;
  do {
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  } while (!(v & kFlagValid));
  return v & ~kFlagValid;
}

__global__ void __launch_bounds__(kStripWarps * 32) strip_prefix_kernel(const uint32_t* __restrict__ bitmap,
                                                                         long long words_per_scan, int H, int W,
                                                                         int bands, int strips,
                                                                         uint32_t* __restrict__ band_pre,
                                                                         uint32_t* __restrict__ col_bits,
                                                                         uint32_t* __restrict__ col_tot,
                                                                         long long* __restrict__ col_base, int B,
                                                                         long long* __restrict__ counts,
                                                                         unsigned long long* __restrict__ flags) {
  extern __shared__ uint32_t cnt[];            // [bands][kCntPitch] counts, then exclusive prefixes
  __shared__ uint32_t tot_s[kStripCols];
  const int b = blockIdx.x / strips, strip = blockIdx.x - b * strips;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = strip * kStripCols, x = x0 + lane;
  const uint32_t* bm = bitmap + (size_t)b * words_per_scan;
  const uint32_t colmask = (W - x0 >= 32) ? 0xffffffffu : ((1u << (W - x0)) - 1u);
  constexpr int kPrefetch = 4;                  // bands whose row words are fetched before any is transposed
  for (int band0 = warp; band0 < bands; band0 += kStripWarps * kPrefetch) {
    uint32_t rw[kPrefetch];
#pragma unroll
    for (int j = 0; j < kPrefetch; ++j) {
      const int band = band0 + j * kStripWarps;
      const int y = band * kBandRows + lane;
      rw[j] = 0;
      if (band < bands && y < H) {
        const long long i = (long long)y * W + x0;
        rw[j] = __funnelshift_r(__ldg(bm + (i >> 5)), __ldg(bm + (i >> 5) + 1), (uint32_t)(i & 31)) & colmask;
      }
    }
#pragma unroll
    for (int j = 0; j < kPrefetch; ++j) {
      const int band = band0 + j * kStripWarps;
      if (band >= bands) break;                 // warp-uniform
      const uint32_t rowword = rw[j];
      uint32_t colword = 0;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const uint32_t bits = __ballot_sync(0xffffffffu, (rowword >> c) & 1u);
        if (lane == c) colword = bits;
      }
      cnt[band * kCntPitch + lane] = __popc(colword);
      if (x < W) col_bits[((size_t)b * bands + band) * W + x] = colword;    // bit r = pixel (band*32+r, x) valid
    }
  }
  __syncthreads();
  // prefix down each column: warp w takes columns 2w and 2w + 1, lane = band (32 bands per pass)
#pragma unroll
  for (int cc = 0; cc < kStripCols / kStripWarps; ++cc) {
    const int c = warp * (kStripCols / kStripWarps) + cc;
    uint32_t carry = 0;
    for (int p0 = 0; p0 < bands; p0 += 32) {
      const int band = p0 + lane;
      const uint32_t v = band < bands ? cnt[band * kCntPitch + c] : 0u;
      uint32_t inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
      }
      if (band < bands) cnt[band * kCntPitch + c] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) tot_s[c] = carry;
  }
  __syncthreads();
  if (warp == 0) {
    // ---- K3c: packed bases of this strip's columns
    const uint32_t tot = tot_s[lane];
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    const unsigned long long strip_total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned long long* my_flags = flags + (size_t)b * (strips + 1);
    if (lane == 0) publish_total(my_flags + strip, strip_total);
    if (x < W) col_tot[(size_t)b * W + x] = tot;
    unsigned long long before = 0;                           // lower strips of this scan + all lower scans
    for (int i = lane; i < strip; i += 32) before += await_total(my_flags + i);
    for (int i = lane; i < b; i += 32) before += await_total(flags + (size_t)i * (strips + 1) + strips);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
    if (x < W) col_base[(size_t)b * W + x] = (long long)(before + inc - tot);
    if (strip == strips - 1) {                               // the scan's last strip knows the scan's total
      unsigned long long lower_scans = 0;
      for (int i = lane; i < b; i += 32) lower_scans += await_total(flags + (size_t)i * (strips + 1) + strips);
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) lower_scans += __shfl_xor_sync(0xffffffffu, lower_scans, d);
      const unsigned long long scan_total = before + strip_total - lower_scans;
      if (lane == 0) {
        publish_total(my_flags + strips, scan_total);
        counts[b] = (long long)scan_total;
        if (b == B - 1) counts[B] = (long long)(lower_scans + scan_total);
      }
    }
  } else {
    for (int band = warp - 1; band < bands; band += kStripWarps - 1)
      if (x < W) band_pre[((size_t)b * bands + band) * W + x] = cnt[band * kCntPitch + lane];
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
  int row0;     // first image row of the band being processed (row-band sharding), 0 for whole images
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
  d.row0 = c.cam_row0;
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

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void rays_to_point_fp32(float cx, float cy, float px, float py, const CalibDev& c,
                                                   float& X, float& Y, float& Z) {
  const float t0 = c.Tf[0], t1 = c.Tf[1], t2 = c.Tf[2];
  const float ax = t1 - t2 * cy, ay = t2 * cx - t0, az = t0 * cy - t1 * cx;     // T x (cx,cy,1)
  const float bx = t1 - t2 * py, by = t2 * px - t0, bz = t0 * py - t1 * px;     // T x (px,py,1)
  const float an = sqrt_approx(ax * ax + ay * ay + az * az);     // ~1 ulp; the fp32 mode's tolerance is 1e-4
  const float bn = sqrt_approx(bx * bx + by * by + bz * bz);
  const float tc = t0 * cx + t1 * cy + t2;
  const float tp = t0 * px + t1 * py + t2;
  const float s = c.t_len2f * bn * rcp_approx(an * tp - tc * bn);     // 2 instructions; __fdividef adds range scaling
  X = cx * s; Y = cy * s; Z = s;
}

// The same point from the camera ray c = (cx, cy, 1) and the projector entry (bn, tp) = (|T x p|, T.p):
//   |T x c|^2 = |T|^2 |c|^2 - (T.c)^2   (no cross product),   P = c * |T|^2 bn / (|T x c| tp - (T.c) bn)
// 10 FFMA/FMUL + 2 MUFU per point.  The subtraction cancels only when the camera ray runs along the baseline, where the
// triangle is degenerate anyway; measured against the fp64 reference on the synthetic scans: 5.0e-7 relative, the same
// as the cross-product form.
__device__ __forceinline__ void rays_to_point_fp32_bt(float cx, float cy, float bn, float tp, float t0, float t1,
                                                      float t2, float t_len2, float& X, float& Y, float& Z) {
  const float tc = fmaf(t0, cx, fmaf(t1, cy, t2));
  const float c2 = fmaf(cx, cx, fmaf(cy, cy, 1.0f));
  const float an = sqrt_approx(fmaf(t_len2, c2, -tc * tc));
  const float s = t_len2 * bn * rcp_approx(fmaf(an, tp, -tc * bn));
  X = cx * s; Y = cy * s; Z = s;
}

// Lens table: the undistorted, normalised ray of every lattice point (u, v) in [0,w) x [0,h),
// computed ONCE per calibration with the exact fp64 iteration (bit-identical to
// cv2.undistortPoints, triangulate.py:84-85).  The camera pixel grid and the projector pixel grid
// never change between scans, so the per-scan kernels look the rays up instead of iterating.
// BT: store (|T x p|, T.p) of the exact ray instead of the ray (projector table of the fp32 mode).
template <bool BT>
__global__ void __launch_bounds__(256) lens_table_kernel(const Lens<double> lens, int w, int h,
                                                         float2* __restrict__ table, float t0, float t1, float t2) {
  const long long n = (long long)w * h;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (long long)y * w);
    float ox, oy;
    undistort_fp64((float)x, (float)y, lens, ox, oy);
    if (BT) {
      const float bx = t1 - t2 * oy, by = t2 * ox - t0, bz = t0 * oy - t1 * ox;     // T x (px, py, 1)
      table[i] = make_float2(sqrt_approx(bx * bx + by * by + bz * bz), t0 * ox + t1 * oy + t2);
    } else {
      table[i] = make_float2(ox, oy);
    }
  }
}

// ------------------------------------------------------------------ K2 emit
enum { kModePairs = 0, kModeCloud32 = 1, kModeCloud64 = 2, kModeTable64 = 4, kModeTableBT32 = 5 };
// kModeTableBT32 is the throughput path: the projector table holds, per panel pixel, (|T x p|, T.p) of its exact ray p --
// everything the fp32 formula needs from the projector side (sl_lens_tables with SL_PREC_FP32); kModeTable64 looks the
// exact rays themselves up and mirrors the reference's dtype flow (tables built with SL_PREC_FP64).

struct EmitArgs {
  const int16_t* h;
  const int16_t* v;
  const uint8_t* white;     // nullable
  const uint32_t* band_pre;
  const uint32_t* col_bits;
  const long long* col_base;
  float* out;               // (N,4) records, see emit_kernel
  float* out_extra[kMaxWorld - 1];   // further destinations of the same records (peer-mapped clouds of other ranks)
  int n_extra;
  int remote;               // `out` may be a peer mapping: plain stores (no L2 policy hint)
  int aligned;              // h, v, white, out, cam_rays meet the vector path's alignment (else scalar loads)
  uint32_t* out_rgb32;      // pairs mode only, nullable: (N,) packed colour
  const float2* cam_rays;   // table modes: [H][W] undistorted camera rays (rotated by R)
  const float2* proj_rays;  // table modes: [proj_h][proj_w] undistorted projector rays
  int H, W, bands, tiles_x;
  int off32;                // every packed offset fits 32 bits (B*H*W < 2^32): cheaper warp shuffles
  long long total_tiles;
  CalibDev cal;
};

// Per-warp shared memory: for each of the tile's columns a run of up to 32 compacted 16-byte records
// (+1 record of padding per column against bank conflicts).
constexpr int kRunPitch = kBandRows + 1;               // uint4 records per column
template <bool AUX>
struct EmitTile {
  uint4 rec[kEmitCols * kRunPitch];
  uint32_t aux[AUX ? kEmitCols * kRunPitch : 1];       // pairs mode: the colour word of each point
};

// One warp per tile of 32 rows x kEmitCols columns.
//  Phase A (row-major; a lane owns 4 neighbouring pixels of a row): vector loads of the indices, the
//    colour and the camera rays, gathers of the projector rays (neighbouring pixels hit neighbouring
//    projector pixels), the triangulation.  The rank of a valid pixel inside its column is
//    popc(col_bits & rows-above), so every 16-byte record goes straight to its slot of the column's run.
//  Phase B: each column's run leaves shared memory as one contiguous, coalesced block of records.
//    Runs of successive bands / columns abut in the output: the reference's x-major order
//    (triangulate.py:52-53).
// Record layouts (out, 4 words per point):
//    cloud modes  {x, y, z, rgb}     rgb = R | G<<8 | B<<16 (0 without a colour image)
//    kModePairs   {cam_x, cam_y, proj_x, proj_y}, colour word (same packing) to out_rgb32 when given
#ifndef SL_EMIT_MIN_CTAS
#define SL_EMIT_MIN_CTAS 5      // <= 102 registers: an emit CTA fits beside a resident decode CTA (see sl_decode.cu)
#endif
template <int MODE>
__global__ void __launch_bounds__(kEmitWarps * 32, SL_EMIT_MIN_CTAS) emit_kernel(const EmitArgs a) {
  extern __shared__ __align__(16) uint8_t emit_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * kEmitWarps + warp;
  if (t >= a.total_tiles) return;
  using Tile = EmitTile<MODE == kModePairs>;
  Tile& tile = reinterpret_cast<Tile*>(emit_smem)[warp];
  // tile -> (scan, band, column block); 32-bit arithmetic (total_tiles < 2^31 is checked by the launcher)
  const uint32_t ti = (uint32_t)t;
  const uint32_t per_scan = (uint32_t)a.tiles_x * (uint32_t)a.bands;
  const int b = (int)(ti / per_scan);
  const uint32_t in_scan = ti - (uint32_t)b * per_scan;
  const int band = (int)(in_scan / (uint32_t)a.tiles_x);
  const int tx = (int)(in_scan - (uint32_t)band * (uint32_t)a.tiles_x);
  const int x0 = tx * kEmitCols, y0 = band * kBandRows;
  const int cols = min(kEmitCols, a.W - x0);
  const size_t img = (size_t)b * a.H * a.W;
  const size_t cell = ((size_t)b * a.bands + band) * a.W + x0;       // (scan, band, column x0)
  const bool has_rgb = a.white != nullptr;
  const bool vec = ((a.W & 3) == 0) && cols == kEmitCols && a.aligned;
  constexpr bool kTable = (MODE == kModeTable64 || MODE == kModeTableBT32);

  // lane c <-> column x0+c: packed output offset of the column's first point in this band, and its bits
  long long my_base = 0;
  uint32_t my_bits = 0;
  if (lane < cols) {
    my_base = a.col_base[(size_t)b * a.W + x0 + lane] + (long long)a.band_pre[cell + lane];
    my_bits = a.col_bits[cell + lane];
  }

  // ---- phase A
  const int rr = lane >> 2, c4 = (lane & 3) * 4;
  const int rows = min(kBandRows, a.H - y0);
  // result of one pixel -> its slot in the column's run
  auto store_point = [&](int k, int r, uint32_t bits, int hc, int vc, float2 cray, float2 pray, uint32_t colour) {
    float X, Y, Z = 0.f;
    uint32_t wword = colour;
    const float u = (float)(x0 + c4 + k), w = (float)(a.cal.row0 + y0 + r);
    if (MODE == kModePairs) {
      X = u; Y = w; Z = (float)hc; wword = __float_as_uint((float)vc);
    } else if (MODE == kModeCloud32) {
      triangulate_fp32(u, w, (float)hc, (float)vc, a.cal, X, Y, Z);
    } else if (MODE == kModeCloud64) {
      double dx, dy, dz;
      triangulate_fp64(u, w, (float)hc, (float)vc, a.cal, dx, dy, dz);
      X = (float)dx; Y = (float)dy; Z = (float)dz;
    } else if (MODE == kModeTableBT32) {
      rays_to_point_fp32_bt(cray.x, cray.y, pray.x, pray.y, a.cal.Tf[0], a.cal.Tf[1], a.cal.Tf[2], a.cal.t_len2f, X, Y, Z);
    } else {
      double dx, dy, dz;
      rays_to_point_fp64(cray.x, cray.y, pray.x, pray.y, a.cal, dx, dy, dz);
      X = (float)dx; Y = (float)dy; Z = (float)dz;
    }
    const int rank = __popc(bits & ((1u << r) - 1u));
    tile.rec[(c4 + k) * kRunPitch + rank] = make_uint4(__float_as_uint(X), __float_as_uint(Y), __float_as_uint(Z), wword);
    if (MODE == kModePairs) tile.aux[(c4 + k) * kRunPitch + rank] = colour;
  };

  // L2 policy: the lens tables are re-read by every scan (keep them: evict_last); indices, colours and
  // the cloud stream through once (evict_first)
  const uint64_t keep = policy_evict_last(), stream = policy_evict_first();
  if (MODE == kModeTableBT32 && vec) {
    // Throughput path.  Everything per pixel is straight-line code: packed 16-bit clamps, the gathers predicated on the
    // validity bit, the 12-instruction point formula evaluated for all 16 pixels of the lane (no branch per pixel; an
    // invalid pixel computes garbage that is never stored), one predicated 16-byte shared-memory store per pixel.
    const float t0 = a.cal.Tf[0], t1 = a.cal.Tf[1], t2 = a.cal.Tf[2], tl2 = a.cal.t_len2f;
    const uint32_t pw1 = (uint32_t)(a.cal.proj_w - 1), ph1 = (uint32_t)(a.cal.proj_h - 1);
    const uint32_t pwx2 = pw1 | (pw1 << 16), phx2 = ph1 | (ph1 << 16);
    const uint32_t pw = (uint32_t)a.cal.proj_w;
    uint32_t cb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cb[k] = __shfl_sync(0xffffffffu, my_bits, c4 + k);
    const uint32_t rec_base = smem_u32(&tile.rec[c4 * kRunPitch]);
#pragma unroll 1
    for (int it0 = 0; it0 < 4; it0 += kEmitGroup) {
      uint2 h2[kEmitGroup], v2[kEmitGroup];
      uint32_t w0[kEmitGroup], w1[kEmitGroup], w2[kEmitGroup];
      float4 r0[kEmitGroup], r1[kEmitGroup];
#pragma unroll
      for (int j = 0; j < kEmitGroup; ++j) {
        const int r = min(rr + 8 * (it0 + j), rows - 1);        // rows past the image repeat the last one (never valid)
        const size_t g = img + (size_t)(y0 + r) * a.W + x0 + c4;
        h2[j] = ldg_u2_hint(a.h + g, stream);
        v2[j] = ldg_u2_hint(a.v + g, stream);
        if (has_rgb) {
          const uint32_t* wp = reinterpret_cast<const uint32_t*>(a.white + g * 3);      // 12 bytes = 4 RGB pixels
          w0[j] = ldg_u32_hint(wp, stream); w1[j] = ldg_u32_hint(wp + 1, stream); w2[j] = ldg_u32_hint(wp + 2, stream);
        } else {
          w0[j] = w1[j] = w2[j] = 0;
        }
        const float4* rp = reinterpret_cast<const float4*>(a.cam_rays + (size_t)(a.cal.row0 + y0 + r) * a.W + x0 + c4);
        r0[j] = ldg_f4_hint(rp, keep); r1[j] = ldg_f4_hint(rp + 1, keep);
      }
      float2 bt[kEmitGroup][4];
#pragma unroll
      for (int j = 0; j < kEmitGroup; ++j) {
        const int r = rr + 8 * (it0 + j);
        // triangulate.py:60-61: min(proj_w - 1, h), min(proj_h - 1, v) on both 16-bit halves at once (-1 = 0xffff clamps too)
        const uint32_t hx = __vminu2(h2[j].x, pwx2), hy = __vminu2(h2[j].y, pwx2);
        const uint32_t vx = __vminu2(v2[j].x, phx2), vy = __vminu2(v2[j].y, phx2);
        const uint32_t idx[4] = {(vx & 0xffffu) * pw + (hx & 0xffffu), (vx >> 16) * pw + (hx >> 16),
                                 (vy & 0xffffu) * pw + (hy & 0xffffu), (vy >> 16) * pw + (hy >> 16)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          bt[j][k] = make_float2(0.f, 0.f);
          if ((cb[k] >> r) & 1u) bt[j][k] = ldg_f2_hint(a.proj_rays + idx[k], keep);        // triangulate.py:56
        }
      }
#pragma unroll
      for (int j = 0; j < kEmitGroup; ++j) {
        const int r = rr + 8 * (it0 + j);
        const uint32_t below = (1u << r) - 1u;
        const uint32_t col[4] = {w0[j] & 0xffffffu, (w0[j] >> 24) | ((w1[j] & 0xffffu) << 8),
                                 (w1[j] >> 16) | ((w2[j] & 0xffu) << 16), w2[j] >> 8};   // triangulate.py:64
        const float cx[4] = {r0[j].x, r0[j].z, r1[j].x, r1[j].z}, cy[4] = {r0[j].y, r0[j].w, r1[j].y, r1[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float X, Y, Z;
          rays_to_point_fp32_bt(cx[k], cy[k], bt[j][k].x, bt[j][k].y, t0, t1, t2, tl2, X, Y, Z);
          if ((cb[k] >> r) & 1u) {
            const uint32_t slot = rec_base + (uint32_t)(k * kRunPitch + __popc(cb[k] & below)) * 16u;
            sts_u4(slot, __float_as_uint(X), __float_as_uint(Y), __float_as_uint(Z), col[k]);
          }
        }
      }
    }
  } else if (vec) {
    uint32_t cb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cb[k] = __shfl_sync(0xffffffffu, my_bits, c4 + k);
    // The tile's 4 row groups are processed kEmitGroup at a time: more groups per pass = more loads
    // in flight per warp, fewer = fewer registers (more resident warps).
#pragma unroll 1
    for (int it0 = 0; it0 < 4; it0 += kEmitGroup) {
      // (1) every load that depends on nothing, back to back
      uint2 h2[kEmitGroup], v2[kEmitGroup];
      uint32_t w0[kEmitGroup], w1[kEmitGroup], w2[kEmitGroup];
      float4 r0[kEmitGroup], r1[kEmitGroup];
#pragma unroll
      for (int j = 0; j < kEmitGroup; ++j) {
        const int r = min(rr + 8 * (it0 + j), rows - 1);        // rows past the image repeat the last one (never valid)
        const size_t g = img + (size_t)(y0 + r) * a.W + x0 + c4;
        h2[j] = ldg_u2_hint(a.h + g, stream);
        v2[j] = ldg_u2_hint(a.v + g, stream);
        if (has_rgb) {
          const uint32_t* wp = reinterpret_cast<const uint32_t*>(a.white + g * 3);      // 12 bytes = 4 RGB pixels
          w0[j] = ldg_u32_hint(wp, stream); w1[j] = ldg_u32_hint(wp + 1, stream); w2[j] = ldg_u32_hint(wp + 2, stream);
        } else {
          w0[j] = w1[j] = w2[j] = 0;
        }
        if (kTable) {
          const float4* rp = reinterpret_cast<const float4*>(a.cam_rays + (size_t)(a.cal.row0 + y0 + r) * a.W + x0 + c4);
          r0[j] = ldg_f4_hint(rp, keep); r1[j] = ldg_f4_hint(rp + 1, keep);
        }
      }
      // (2) the projector-ray gathers (neighbouring pixels hit neighbouring table entries)
      float2 pray[kEmitGroup][4];
      int hc[kEmitGroup][4], vc[kEmitGroup][4];
#pragma unroll
      for (int j = 0; j < kEmitGroup; ++j) {
        const int r = rr + 8 * (it0 + j);
        const uint32_t hw[2] = {h2[j].x, h2[j].y}, vw[2] = {v2[j].x, v2[j].y};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int hq = (hw[k >> 1] >> (16 * (k & 1))) & 0xffff, vq = (vw[k >> 1] >> (16 * (k & 1))) & 0xffff;
          hc[j][k] = min(a.cal.proj_w - 1, hq); vc[j][k] = min(a.cal.proj_h - 1, vq);     // triangulate.py:60-61
          pray[j][k] = make_float2(0.f, 0.f);
          if (kTable && ((cb[k] >> r) & 1u))                                              // triangulate.py:56
            pray[j][k] = ldg_f2_hint(a.proj_rays + (size_t)vc[j][k] * a.cal.proj_w + hc[j][k], keep);
        }
      }
      // (3) arithmetic + store
#pragma unroll
      for (int j = 0; j < kEmitGroup; ++j) {
        const int r = rr + 8 * (it0 + j);
        const uint32_t col[4] = {w0[j] & 0xffffffu, (w0[j] >> 24) | ((w1[j] & 0xffffu) << 8),
                                 (w1[j] >> 16) | ((w2[j] & 0xffu) << 16), w2[j] >> 8};   // triangulate.py:64
        const float2 cray[4] = {make_float2(r0[j].x, r0[j].y), make_float2(r0[j].z, r0[j].w),
                                make_float2(r1[j].x, r1[j].y), make_float2(r1[j].z, r1[j].w)};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((cb[k] >> r) & 1u) store_point(k, r, cb[k], hc[j][k], vc[j][k], cray[k], pray[j][k], col[k]);
      }
    }
  } else {
    // ragged tiles / row pitches that are not a multiple of 4 pixels: scalar loads, one row group at a time
    uint32_t cb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cb[k] = __shfl_sync(0xffffffffu, my_bits, c4 + k);
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
      const int r = rr + 8 * it;
      const size_t g = img + (size_t)(y0 + r) * a.W + x0 + c4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!((cb[k] >> r) & 1u)) continue;             // a set bit implies the pixel is inside the image
        const int hq = (uint16_t)a.h[g + k], vq = (uint16_t)a.v[g + k];
        uint32_t colour = 0;
        if (has_rgb) {
          const uint8_t* wp = a.white + (g + k) * 3;
          colour = (uint32_t)wp[0] | ((uint32_t)wp[1] << 8) | ((uint32_t)wp[2] << 16);
        }
        const int hcl = min(a.cal.proj_w - 1, hq), vcl = min(a.cal.proj_h - 1, vq);
        float2 cray = make_float2(0.f, 0.f), pray = cray;
        if (kTable) {
          cray = __ldg(a.cam_rays + (size_t)(a.cal.row0 + y0 + r) * a.W + x0 + c4 + k);
          pray = __ldg(a.proj_rays + (size_t)vcl * a.cal.proj_w + hcl);
        }
        store_point(k, r, cb[k], hcl, vcl, cray, pray, colour);
      }
    }
  }
  __syncwarp();

  // ---- phase B
  uint4* out = reinterpret_cast<uint4*>(a.out);
  for (int cidx = 0; cidx < cols; ++cidx) {
    const int cnt = __popc(__shfl_sync(0xffffffffu, my_bits, cidx));
    const long long base = a.off32 ? (long long)__shfl_sync(0xffffffffu, (unsigned int)my_base, cidx)
                                   : __shfl_sync(0xffffffffu, my_base, cidx);
    if (lane < cnt) {
      const uint4 rec = tile.rec[cidx * kRunPitch + lane];
      if (a.remote) out[base + lane] = rec;
      else stg_u4_hint(out + base + lane, rec, stream);
      for (int k = 0; k < a.n_extra; ++k) reinterpret_cast<uint4*>(a.out_extra[k])[base + lane] = rec;
      if (MODE == kModePairs && a.out_rgb32) a.out_rgb32[base + lane] = tile.aux[cidx * kRunPitch + lane];
    }
  }
}

// ------------------------------------------------------------------ K2f emit, throughput path
// The same emission for the common case (fp32, lens tables with the (|T x p|, T.p) projector table, W a multiple of 16,
// aligned buffers): no mode switches, and a tile of ROWS = 16 rows x 16 columns per warp (half a band) instead of 32.
// The kernel is bound by memory latency, not by instructions or bytes (ncu: 26 % of the warp slots, long-scoreboard
// stalls 5 per issued instruction, every phase of a tile -- first-level loads, table gathers, arithmetic, stores --
// waits for the previous one), so the lever is more, smaller, lighter warps in flight: half the registers' worth of
// loads per warp, twice the tiles, and the column runs of a tile leave two columns at a time (half a warp each).
// A half-band tile ranks its pixels inside the band's column word as before; its run starts popc(bits below its first
// row) records into the band's run, so the output is byte-identical to the one-band-per-warp kernel.
// ROWS = 16 for local clouds; ROWS = 32 (whole bands, runs of up to 512 bytes) when the records go to peer-mapped
// clouds: NVLink write requests are 128 bytes, a 16-byte-aligned run of 256 bytes wastes 1 request in 3, one of 512
// bytes 1 in 5 (measured on 2 GPUs: 398 GB/s into the gathering rank with 16-row tiles, 434 GB/s with 32-row tiles).
#ifndef SL_EMITF_MIN_CTAS
#define SL_EMITF_MIN_CTAS 7      // <= 73 registers: 7 CTAs (28 warps) per SM; measured 15.0 us against 16.1 us with 6
#endif
constexpr int fast_smem(int rows) { return kEmitWarps * kEmitCols * (rows + 3) * 16; }

template <int ROWS, bool MULTI>
__global__ void __launch_bounds__(kEmitWarps * 32, ROWS == 16 ? SL_EMITF_MIN_CTAS : 5) emit_fast_kernel(const EmitArgs a) {
  constexpr int kFastRows = ROWS;                       // 16 or 32
  constexpr int kFastGroups = kFastRows / 8;            // row groups of 8 rows (a lane owns 4 neighbouring pixels of a row)
  constexpr int kFastParts = kBandRows / kFastRows;     // tiles per band and column block
  // Run buffer of column c: kFastPitch 16-byte slots starting at slot c * kFastPitch + 2 * (c >> 3).  A quarter warp (the
  // unit of a 128-bit shared store) holds 2 rows x the 4 column quads c4 = 0, 4, 8, 12; with a pitch = 3 (mod 8) and the
  // extra 2 slots for the upper 8 columns their 8 records fall into 8 different 16-byte bank groups when the ranks of
  // neighbouring rows differ by one (the common case); a plain pitch of rows + 1 left every such store 2-way conflicted.
  constexpr int kFastPitch = kFastRows + 3;             // 19 (16-row tiles) / 35 (32-row tiles): both = 3 (mod 8)
  extern __shared__ __align__(16) uint8_t emit_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * kEmitWarps + warp;
  if (t >= a.total_tiles) return;
  // tile -> (scan, band, part of the band, column block): consecutive warps take consecutive column blocks
  const uint32_t ti = (uint32_t)t;
  const uint32_t per_band = (uint32_t)a.tiles_x * kFastParts;
  const uint32_t per_scan = per_band * (uint32_t)a.bands;
  const int b = (int)(ti / per_scan);
  const uint32_t in_scan = ti - (uint32_t)b * per_scan;
  const int band = (int)(in_scan / per_band);
  const uint32_t in_band = in_scan - (uint32_t)band * per_band;
  const int part = (int)(in_band / (uint32_t)a.tiles_x);
  const int tx = (int)(in_band - (uint32_t)part * (uint32_t)a.tiles_x);
  const int x0 = tx * kEmitCols, y0 = band * kBandRows + part * kFastRows;
  if (y0 >= a.H) return;                                   // lower half of a ragged last band
  const size_t img = (size_t)b * a.H * a.W;
  const size_t cell = ((size_t)b * a.bands + band) * a.W + x0;
  const uint32_t rec_warp = smem_u32(emit_smem) + (uint32_t)warp * (kEmitCols * kFastPitch * 16);

  // lane c < 16 <-> column x0 + c: where this tile's run of the column starts in the packed output, and its rows' bits
  long long my_base = 0;
  uint32_t my_bits = 0;
  if (lane < kEmitCols) {
    const uint32_t bits32 = a.col_bits[cell + lane];
    my_base = a.col_base[(size_t)b * a.W + x0 + lane] + (long long)a.band_pre[cell + lane];
    if (kFastParts > 1) {
      my_base += __popc(bits32 & ((1u << (part * kFastRows)) - 1u));
      my_bits = (bits32 >> (part * kFastRows)) & (kFastRows >= 32 ? 0xffffffffu : ((1u << (kFastRows & 31)) - 1u));
    } else {
      my_bits = bits32;
    }
  }

  // ---- phase A: row-major; lane owns pixels (row rr + 8 g, columns c4 .. c4 + 3)
  const int rr = lane >> 2, c4 = (lane & 3) * 4;
  const int rows = min(kFastRows, a.H - y0);
  const bool has_rgb = a.white != nullptr;
  const uint64_t keep = policy_evict_last(), stream = policy_evict_first();
  const float t0 = a.cal.Tf[0], t1 = a.cal.Tf[1], t2 = a.cal.Tf[2], tl2 = a.cal.t_len2f;
  const uint32_t pw1 = (uint32_t)(a.cal.proj_w - 1), ph1 = (uint32_t)(a.cal.proj_h - 1);
  const uint32_t pwx2 = pw1 | (pw1 << 16), phx2 = ph1 | (ph1 << 16);
  const uint32_t pw = (uint32_t)a.cal.proj_w;
  uint32_t cb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) cb[k] = __shfl_sync(0xffffffffu, my_bits, c4 + k);
  const uint32_t rec_base = rec_warp + (uint32_t)(c4 * kFastPitch + 2 * (c4 >> 3)) * 16u;

  uint2 h2[kFastGroups], v2[kFastGroups];
  uint32_t w0[kFastGroups], w1[kFastGroups], w2[kFastGroups];
  float4 r0[kFastGroups], r1[kFastGroups];
#pragma unroll
  for (int j = 0; j < kFastGroups; ++j) {
    const int r = min(rr + 8 * j, rows - 1);                // rows past the image repeat the last one (never valid)
    const size_t g = img + (size_t)(y0 + r) * a.W + x0 + c4;
    h2[j] = ldg_u2_hint(a.h + g, stream);
    v2[j] = ldg_u2_hint(a.v + g, stream);
    if (has_rgb) {
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(a.white + g * 3);      // 12 bytes = 4 RGB pixels
      w0[j] = ldg_u32_hint(wp, stream); w1[j] = ldg_u32_hint(wp + 1, stream); w2[j] = ldg_u32_hint(wp + 2, stream);
    } else {
      w0[j] = w1[j] = w2[j] = 0;
    }
    const float4* rp = reinterpret_cast<const float4*>(a.cam_rays + (size_t)(a.cal.row0 + y0 + r) * a.W + x0 + c4);
    r0[j] = ldg_f4_hint(rp, keep); r1[j] = ldg_f4_hint(rp + 1, keep);
  }
  float2 bt[kFastGroups][4];
#pragma unroll
  for (int j = 0; j < kFastGroups; ++j) {
    const int r = rr + 8 * j;
    // triangulate.py:60-61: min(proj_w - 1, h), min(proj_h - 1, v), both 16-bit halves at once (-1 = 0xffff clamps too)
    const uint32_t hx = __vminu2(h2[j].x, pwx2), hy = __vminu2(h2[j].y, pwx2);
    const uint32_t vx = __vminu2(v2[j].x, phx2), vy = __vminu2(v2[j].y, phx2);
    const uint32_t idx[4] = {(vx & 0xffffu) * pw + (hx & 0xffffu), (vx >> 16) * pw + (hx >> 16),
                             (vy & 0xffffu) * pw + (hy & 0xffffu), (vy >> 16) * pw + (hy >> 16)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bt[j][k] = make_float2(0.f, 0.f);
      if ((cb[k] >> r) & 1u) bt[j][k] = ldg_f2_hint(a.proj_rays + idx[k], keep);            // triangulate.py:56
    }
  }
#pragma unroll
  for (int j = 0; j < kFastGroups; ++j) {
    const int r = rr + 8 * j;
    const uint32_t below = (1u << r) - 1u;
    const uint32_t col[4] = {w0[j] & 0xffffffu, (w0[j] >> 24) | ((w1[j] & 0xffffu) << 8),
                             (w1[j] >> 16) | ((w2[j] & 0xffu) << 16), w2[j] >> 8};       // triangulate.py:64
    const float cx[4] = {r0[j].x, r0[j].z, r1[j].x, r1[j].z}, cy[4] = {r0[j].y, r0[j].w, r1[j].y, r1[j].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float X, Y, Z;
      rays_to_point_fp32_bt(cx[k], cy[k], bt[j][k].x, bt[j][k].y, t0, t1, t2, tl2, X, Y, Z);
      if ((cb[k] >> r) & 1u)
        sts_u4(rec_base + (uint32_t)(k * kFastPitch + __popc(cb[k] & below)) * 16u, __float_as_uint(X),
               __float_as_uint(Y), __float_as_uint(Z), col[k]);
    }
  }
  __syncwarp();

  // ---- phase B: a column's run is one contiguous block of records.  Half-band tiles: two columns per step, half a warp
  // each (runs of <= 16 records).  Whole-band tiles (peer-mapped clouds): one column per step, the whole warp on one run
  // of up to 512 contiguous bytes -- the larger the contiguous store, the fewer partial 128-byte NVLink write requests.
  uint4* out = reinterpret_cast<uint4*>(a.out);
  if (kFastRows == 16) {
    const int l16 = lane & 15, hsel = lane >> 4;
#pragma unroll
    for (int it = 0; it < kEmitCols / 2; ++it) {
      const int c = 2 * it + hsel;
      const int cnt = __popc(__shfl_sync(0xffffffffu, my_bits, c));
      const long long base = a.off32 ? (long long)__shfl_sync(0xffffffffu, (unsigned int)my_base, c)
                                     : __shfl_sync(0xffffffffu, my_base, c);
      if (l16 < cnt) {
        const uint4 rec = lds_u4(rec_warp + (uint32_t)(c * kFastPitch + 2 * (c >> 3) + l16) * 16u);
        if (MULTI) {
          out[base + l16] = rec;
          for (int k = 0; k < a.n_extra; ++k) reinterpret_cast<uint4*>(a.out_extra[k])[base + l16] = rec;
        } else {
          stg_u4_hint(out + base + l16, rec, stream);
        }
      }
    }
  } else {
#pragma unroll 4
    for (int c = 0; c < kEmitCols; ++c) {
      const int cnt = __popc(__shfl_sync(0xffffffffu, my_bits, c));
      const long long base = a.off32 ? (long long)__shfl_sync(0xffffffffu, (unsigned int)my_base, c)
                                     : __shfl_sync(0xffffffffu, my_base, c);
      if (lane < cnt) {
        const uint4 rec = lds_u4(rec_warp + (uint32_t)(c * kFastPitch + 2 * (c >> 3) + lane) * 16u);
        if (MULTI) {
          out[base + lane] = rec;
          for (int k = 0; k < a.n_extra; ++k) reinterpret_cast<uint4*>(a.out_extra[k])[base + lane] = rec;
        } else {
          stg_u4_hint(out + base + lane, rec, stream);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ K2r emit, lane = row
// emit_fast_kernel above is bound by the LSU: a lane owns 4 neighbouring pixels of a row, so every load instruction of the
// warp spans 8 image rows = 8 cache lines, whatever its width (ncu on a 16-scan batch: 36 L1 wavefronts per 32 pixels, the
// l1tex data pipe at 75 %), and the x-major order is restored by staging the 16-byte records in shared memory.
// This kernel transposes on the INPUT side instead, where a pixel is 7 bytes and not 16:
//   * a warp takes a tile of 32 rows x 32 columns; the row-major inputs (indices, colours) are fetched with full-width
//     16-byte loads (4 / 6 lanes per row) and parked in shared memory (indices: 16-byte chunks XOR-swizzled so that both the
//     row-wise stores and the per-lane 16-byte reads are conflict-free; colours: 104-byte row pitch);
//   * then lane = row.  For a column c the 32 lanes hold the 32 rows of that column, i.e. the column's run in order:
//     rank = popc(ballot(valid) below the lane), and the valid lanes write their records straight from registers to
//     consecutive slots -- one coalesced store of up to 512 bytes per column, no record ever touches shared memory;
//   * the lens tables are read along a column, so they are stored TRANSPOSED (cam_t [W][rows], proj_t [proj_w][proj_h],
//     built once per calibration): the camera rays of a column are contiguous (2 lines per load), and the projector
//     entries of a column are too, because along a camera column the decoded projector row advances with the camera row
//     while the projector column stays put.
// Per 32 pixels that is ~17 L1 wavefronts: staging 3 + 1.5, shared reads 1.75, bases 1, camera rays 2, gather ~3, store
// ~4.5.  The records are bit-identical to emit_fast_kernel's (same table values, same formula).
struct EmitRowsArgs {
  const int16_t* h;
  const int16_t* v;
  const uint8_t* white;     // nullable
  const uint32_t* band_pre;
  const uint32_t* col_bits;
  const long long* col_base;
  uint4* out;
  uint4* out_extra[kMaxWorld - 1];   // MULTI: further destinations of the same records (peer-mapped clouds)
  int n_extra;
  const float2* cam_t;      // [W][cam_rows]
  const float2* proj_t;     // [proj_w][proj_h]: (|T x p|, T.p)
  int H, W, bands, tiles_x, cam_rows, row0;
  uint32_t pw1, ph1, ph;
  float t0, t1, t2, tl2;
  long long total_tiles;
};
#ifndef SL_ROWS_WARPS
#define SL_ROWS_WARPS 4
#endif
constexpr int kRowsWarps = SL_ROWS_WARPS;
constexpr int kRowsWhitePitch = 104;                     // bytes; 26 words: 16 consecutive rows start in 16 different even banks
constexpr int kRowsIdxBytes = 32 * 64;                   // one index tile: 32 rows x 32 int16
constexpr int kRowsSmemWarp = 2 * kRowsIdxBytes + 32 * kRowsWhitePitch + 32 * 8;     // h, v, colours, column bases
constexpr int kRowsSmem = kRowsWarps * kRowsSmemWarp;

__device__ __forceinline__ uint4 ldg_u4_hint(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

#ifndef SL_ROWS_MIN_CTAS
#define SL_ROWS_MIN_CTAS 4
#endif
#ifndef SL_ROWS_UNROLL
#define SL_ROWS_UNROLL 1
#endif
#define SL_PRAGMA_(x) _Pragma(#x)
#define SL_UNROLL_N(n) SL_PRAGMA_(unroll n)
// MULTI: `out` and out_extra[] may be peer mappings (multi-GPU gather, sl_peer.cu): plain stores, every record to all of them
template <bool MULTI>
__global__ void __launch_bounds__(kRowsWarps * 32, SL_ROWS_MIN_CTAS) emit_rows_kernel(const EmitRowsArgs a) {
  extern __shared__ __align__(16) uint8_t rows_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * kRowsWarps + warp;
  if (t >= a.total_tiles) return;
  const uint32_t ti = (uint32_t)t;
  const uint32_t per_scan = (uint32_t)a.tiles_x * (uint32_t)a.bands;
  const int b = (int)(ti / per_scan);
  const uint32_t in_scan = ti - (uint32_t)b * per_scan;
  const int band = (int)(in_scan / (uint32_t)a.tiles_x);
  const int tx = (int)(in_scan - (uint32_t)band * (uint32_t)a.tiles_x);
  const int x0 = tx * 32, y0 = band * kBandRows;
  const int rows = min(kBandRows, a.H - y0);
  const size_t img = (size_t)b * a.H * a.W;
  const size_t cell = ((size_t)b * a.bands + band) * a.W + x0;
  const uint32_t h_s = smem_u32(rows_smem) + (uint32_t)warp * kRowsSmemWarp;
  const uint32_t v_s = h_s + kRowsIdxBytes, w_s = v_s + kRowsIdxBytes, b_s = w_s + 32 * kRowsWhitePitch;
  const bool has_rgb = a.white != nullptr;
  const uint64_t keep = policy_evict_last(), stream = policy_evict_first();

  // ---- the row-major inputs of the tile, full-width loads (chunk q of 16 bytes: row q / 4, columns 8 (q % 4) ..)
  uint4 hq[4], vq[4], wq[6];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = j * 32 + lane;
    const int r = min(q >> 2, rows - 1);                    // rows past the image repeat the last one (never valid)
    const size_t g = img + (size_t)(y0 + r) * a.W + x0 + (q & 3) * 8;
    hq[j] = ldg_u4_hint(a.h + g, stream);
    vq[j] = ldg_u4_hint(a.v + g, stream);
  }
  if (has_rgb) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int q = j * 32 + lane;
      const int r6 = q / 6, k = q - 6 * r6;
      const int r = min(r6, rows - 1);
      wq[j] = ldg_u4_hint(a.white + (img + (size_t)(y0 + r) * a.W + x0) * 3 + k * 16, stream);
    }
  }
  // column metadata, lane <-> column x0 + lane
  const uint32_t my_bits = a.col_bits[cell + lane];
  const long long my_base = a.col_base[(size_t)b * a.W + x0 + lane] + (long long)a.band_pre[cell + lane];
  sts_u2(b_s + (uint32_t)lane * 8u, (uint32_t)(unsigned long long)my_base, (uint32_t)((unsigned long long)my_base >> 32));
  // validity of MY ROW's 32 columns: the transpose of the column words, one vote per row
  uint32_t rowbits = 0;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const uint32_t m = __ballot_sync(0xffffffffu, (my_bits >> r) & 1u);
    if (lane == r) rowbits = m;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = j * 32 + lane;
    const int r = q >> 2, k = q & 3;
    const uint32_t off = (uint32_t)(r * 64 + ((k ^ ((r >> 1) & 3)) << 4));
    sts_u4(h_s + off, hq[j].x, hq[j].y, hq[j].z, hq[j].w);
    sts_u4(v_s + off, vq[j].x, vq[j].y, vq[j].z, vq[j].w);
  }
  if (has_rgb) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int q = j * 32 + lane;
      const int r6 = q / 6, k = q - 6 * r6;
      const uint32_t off = w_s + (uint32_t)(r6 * kRowsWhitePitch + k * 16);
      sts_u2(off, wq[j].x, wq[j].y);
      sts_u2(off + 8, wq[j].z, wq[j].w);
    }
  }
  __syncwarp();

  // ---- lane = row
  const size_t cam_row = (size_t)(a.row0 + y0 + min(lane, rows - 1));
  const uint32_t below = (1u << lane) - 1u;
  const uint32_t swz = (uint32_t)((lane >> 1) & 3);
#if defined(SL_ROWS_PREFETCH)
  float2 cam_next[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cam_next[k] = ldg_f2_hint(a.cam_t + (size_t)(x0 + k) * a.cam_rows + cam_row, keep);
#endif
  SL_UNROLL_N(SL_ROWS_UNROLL)
  for (int g = 0; g < 4; ++g) {
    const uint32_t ioff = (uint32_t)(lane * 64) + ((((uint32_t)g) ^ swz) << 4);
    const uint4 h8 = lds_u4(h_s + ioff), v8 = lds_u4(v_s + ioff);
    uint32_t cw[7];                                        // 24 colour bytes of columns 8g .. 8g+7 (+ a zero word)
    if (has_rgb) {
      const uint32_t woff = w_s + (uint32_t)(lane * kRowsWhitePitch + g * 24);
      const uint2 c0 = lds_u2(woff), c1 = lds_u2(woff + 8), c2 = lds_u2(woff + 16);
      cw[0] = c0.x; cw[1] = c0.y; cw[2] = c1.x; cw[3] = c1.y; cw[4] = c2.x; cw[5] = c2.y;
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) cw[i] = 0;
    }
    cw[6] = 0;
    const uint32_t hw[4] = {h8.x, h8.y, h8.z, h8.w}, vw[4] = {v8.x, v8.y, v8.z, v8.w};
    float2 cam[8], pe[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
#if defined(SL_ROWS_PREFETCH)
      cam[k] = cam_next[k];
#else
      cam[k] = ldg_f2_hint(a.cam_t + (size_t)(x0 + c) * a.cam_rows + cam_row, keep);
#endif
      // triangulate.py:60-61: min(proj_w - 1, h), min(proj_h - 1, v) (-1 = 0xffff clamps too; such a pixel is not valid)
      const uint32_t hc = min((hw[k >> 1] >> (16 * (k & 1))) & 0xffffu, a.pw1);
      const uint32_t vc = min((vw[k >> 1] >> (16 * (k & 1))) & 0xffffu, a.ph1);
      pe[k] = make_float2(0.f, 0.f);
      if ((rowbits >> c) & 1u) pe[k] = ldg_f2_hint(a.proj_t + (size_t)hc * a.ph + vc, keep);        // triangulate.py:56
    }
#if defined(SL_ROWS_PREFETCH)
    if (g < 3) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        cam_next[k] = ldg_f2_hint(a.cam_t + (size_t)(x0 + g * 8 + 8 + k) * a.cam_rows + cam_row, keep);
    }
#endif
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      const bool valid = (rowbits >> c) & 1u;
      float X, Y, Z;
      rays_to_point_fp32_bt(cam[k].x, cam[k].y, pe[k].x, pe[k].y, a.t0, a.t1, a.t2, a.tl2, X, Y, Z);
      // the colour of column 8g + k: bytes 3k .. 3k+2 of the 24 (triangulate.py:64)
      const uint32_t colour = __funnelshift_r(cw[(3 * k) >> 2], cw[((3 * k) >> 2) + 1], 8 * ((3 * k) & 3)) & 0xffffffu;
      const uint32_t colword = __ballot_sync(0xffffffffu, valid);       // = the column's bits: lanes are its rows
      const uint2 bb = lds_u2(b_s + (uint32_t)c * 8u);
      const long long base = (long long)(((unsigned long long)bb.y << 32) | bb.x);
      if (valid) {
        const uint4 rec = make_uint4(__float_as_uint(X), __float_as_uint(Y), __float_as_uint(Z), colour);
        const long long at = base + __popc(colword & below);
        if (MULTI) {
          a.out[at] = rec;
          for (int e = 0; e < a.n_extra; ++e) a.out_extra[e][at] = rec;
        } else {
          stg_u4_hint(a.out + at, rec, stream);
        }
      }
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

__device__ __forceinline__ bool inside_box(double x, double y, double z, double thr) {
  return (z < thr) && (z > -thr) && (y < thr) && (y > -thr) && (x < thr) && (x > -thr);     // triangulate.py:119
}

// REC = false: pts is the reference's (3,N) float64 array (+ optional (N,3) colours);
// REC = true : pts is the packed (N,4) float32 XYZRGB cloud of sl_emit_cloud.
template <bool REC>
__device__ __forceinline__ bool keep_pt(const void* pts, long long N, long long i, double thr) {
  if (REC) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(pts) + i);
    return inside_box((double)q.x, (double)q.y, (double)q.z, thr);
  }
  const double* p = reinterpret_cast<const double*>(pts);
  return inside_box(p[i], p[N + i], p[2 * N + i], thr);
}

template <bool REC>
__global__ void __launch_bounds__(256) filter_count_kernel(const void* __restrict__ pts, long long N,
                                                            double thr, uint32_t* __restrict__ blk_cnt) {
  const long long base = (long long)blockIdx.x * kFilterBlock;
  uint32_t c = 0;
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < N && keep_pt<REC>(pts, N, i, thr)) ++c;
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

template <bool REC>
__global__ void __launch_bounds__(256) filter_emit_kernel(const void* __restrict__ pts,
                                                           const double* __restrict__ colors, long long N,
                                                           double thr, const long long* __restrict__ blk_base,
                                                           void* __restrict__ out_pts,
                                                           double* __restrict__ out_colors) {
  __shared__ uint32_t ws[8];
  const long long base = (long long)blockIdx.x * kFilterBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long run = blk_base[blockIdx.x];
  for (int k = 0; k < 4; ++k) {          // chunks of 256 consecutive points keep the order
    const long long i = base + k * 256 + threadIdx.x;
    const bool keep = i < N && keep_pt<REC>(pts, N, i, thr);
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) ws[warp] = __popc(m);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w = 0; w < 8; ++w) { const uint32_t c = ws[w]; if (w < warp) before += c; total += c; }
    if (keep) {
      const long long o = run + before + __popc(m & ((1u << lane) - 1u));
      if (REC) {
        reinterpret_cast<float4*>(out_pts)[o] = __ldg(reinterpret_cast<const float4*>(pts) + i);
      } else {
        const double* p = reinterpret_cast<const double*>(pts);
        double* q = reinterpret_cast<double*>(out_pts);
        q[o] = p[i]; q[N + o] = p[N + i]; q[2 * N + o] = p[2 * N + i];
        if (colors) {
          out_colors[o * 3 + 0] = colors[i * 3 + 0];
          out_colors[o * 3 + 1] = colors[i * 3 + 1];
          out_colors[o * 3 + 2] = colors[i * 3 + 2];
        }
      }
    }
    run += total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ box filter as a predicate of the compaction
// Triangulate.filter_3d_pts (triangulate.py:99-122: keep |x|, |y|, |z| < threshold) folded into the validity bitmap: bit i
// = pixel i decoded (triangulate.py:56) AND the point the emission will produce for it lies inside the box.  The point
// is computed by the very device functions the emit kernels use (same inputs, same roundings), so offsets + emission
// on this bitmap give exactly sl_filter_cloud_box(emission of everything) -- without the unfiltered cloud ever existing.
template <int MODE>
__global__ void __launch_bounds__(256) box_bitmap_kernel(const int16_t* __restrict__ h, const int16_t* __restrict__ v,
                                                          int H, int W, long long words_per_scan, int B,
                                                          const float2* __restrict__ cam_rays,
                                                          const float2* __restrict__ proj_rays, const CalibDev cal,
                                                          double thr, uint32_t* __restrict__ bitmap,
                                                          unsigned long long* __restrict__ flags, int n_flags) {
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_flags; i += blockDim.x) flags[i] = 0;
  const long long npix = (long long)H * W;
  const long long padded = (npix + 31) & ~31LL;
  const long long total = padded * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / padded, px = i - b * padded;
    bool ok = false;
    if (px < npix) {
      const int hq = h[b * npix + px], vq = v[b * npix + px];
      if (hq != -1 && vq != -1) {
        const int y = (int)(px / W), x = (int)(px - (long long)y * W);
        const int hc = min(cal.proj_w - 1, (int)(uint16_t)hq), vc = min(cal.proj_h - 1, (int)(uint16_t)vq);   // triangulate.py:60-61
        float X, Y, Z;
        if (MODE == kModeTableBT32) {
          const float2 c = __ldg(cam_rays + (size_t)(cal.row0 + y) * W + x);
          const float2 e = __ldg(proj_rays + (size_t)vc * cal.proj_w + hc);
          rays_to_point_fp32_bt(c.x, c.y, e.x, e.y, cal.Tf[0], cal.Tf[1], cal.Tf[2], cal.t_len2f, X, Y, Z);
        } else if (MODE == kModeTable64) {
          const float2 c = __ldg(cam_rays + (size_t)(cal.row0 + y) * W + x);
          const float2 p = __ldg(proj_rays + (size_t)vc * cal.proj_w + hc);
          double dx, dy, dz;
          rays_to_point_fp64(c.x, c.y, p.x, p.y, cal, dx, dy, dz);
          X = (float)dx; Y = (float)dy; Z = (float)dz;
        } else if (MODE == kModeCloud32) {
          triangulate_fp32((float)x, (float)(cal.row0 + y), (float)hc, (float)vc, cal, X, Y, Z);
        } else {
          double dx, dy, dz;
          triangulate_fp64((float)x, (float)(cal.row0 + y), (float)hc, (float)vc, cal, dx, dy, dz);
          X = (float)dx; Y = (float)dy; Z = (float)dz;
        }
        ok = inside_box((double)X, (double)Y, (double)Z, thr);       // on the f32 record, like sl_filter_cloud_box
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) bitmap[b * words_per_scan + (px >> 5)] = m;
  }
}

// ------------------------------------------------------------------ launchers
// Tuning knob (SL_STAGE2_PRIORITY): launch the stage-2 kernels with the highest stream priority, so that the short
// offsets / emit kernels of one scan are placed ahead of the pending decode CTAs of the other streams.
template <class Kernel, class... Args>
static cudaError_t launch_stage2(Kernel k, int grid, int block, size_t smem, cudaStream_t st, Args... args) {
#if defined(SL_STAGE2_PRIORITY)
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributePriority;
  attr[0].val.priority = hi;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, args...);
#else
  k<<<grid, block, smem, st>>>(args...);
  return cudaSuccess;
#endif
}

static int grid_for(long long items, int per_block, int cap = 1 << 20) {
  long long g = (items + per_block - 1) / per_block;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

// offsets from a validity bitmap that is already in the workspace (written by the decode kernel)
int launch_offsets_from_bitmap(int B, int H, int W, void* ws, long long* counts, cudaStream_t st) {
  const CloudWorkspace c = carve_workspace(ws, B, H, W);
  const int strips = (W + kStripCols - 1) / kStripCols;
  const long long warps = (long long)B * strips;
  (void)warps;
  if (c.bands > kMaxBands) { set_error("image too tall (%d rows)", H); return kErrBadArg; }
  const size_t strip_smem = (size_t)c.bands * kCntPitch * sizeof(uint32_t);
  if (strip_smem > 48 * 1024)     // rare (more than 12288 rows): opt in to the larger carve-out
    SL_CUDA(cudaFuncSetAttribute(strip_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)strip_smem));
  SL_CUDA(launch_stage2(strip_prefix_kernel, B * strips, kStripWarps * 32, strip_smem, st,
                        (const uint32_t*)c.bitmap, c.bitmap_words, H, W, c.bands, strips, c.band_cnt, c.col_bits, c.col_tot,
                        c.col_base, B, counts, c.flags));
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_compact_offsets(const int16_t* h, const int16_t* v, int B, int H, int W, void* ws,
                           long long* counts, cudaStream_t st) {
  const CloudWorkspace c = carve_workspace(ws, B, H, W);
  const long long npix = (long long)H * W;
  const long long padded = ((npix + 31) & ~31LL) * B;
  bitmap_from_idx_kernel<<<grid_for(padded, 256, 148 * 16), 256, 0, st>>>(h, v, npix, c.bitmap_words, B, c.bitmap, c.flags, c.n_flags);
  count_launch();
  return launch_offsets_from_bitmap(B, H, W, ws, counts, st);
}

int launch_compact_offsets_boxed(const int16_t* h, const int16_t* v, const SlCalib& calib, int precision,
                                 const float* cam_rays, const float* proj_rays, double threshold, int B, int H, int W,
                                 void* ws, long long* counts, cudaStream_t st) {
  const CloudWorkspace c = carve_workspace(ws, B, H, W);
  const CalibDev cal = make_calib(calib);
  const long long npix = (long long)H * W;
  const long long padded = ((npix + 31) & ~31LL) * B;
  const int g = grid_for(padded, 256, 148 * 16);
  const float2* cr = (const float2*)cam_rays;
  const float2* pr = (const float2*)proj_rays;
  if (cam_rays && precision == SL_PREC_FP32)
    box_bitmap_kernel<kModeTableBT32><<<g, 256, 0, st>>>(h, v, H, W, c.bitmap_words, B, cr, pr, cal, threshold, c.bitmap, c.flags, c.n_flags);
  else if (cam_rays)
    box_bitmap_kernel<kModeTable64><<<g, 256, 0, st>>>(h, v, H, W, c.bitmap_words, B, cr, pr, cal, threshold, c.bitmap, c.flags, c.n_flags);
  else if (precision == SL_PREC_FP32)
    box_bitmap_kernel<kModeCloud32><<<g, 256, 0, st>>>(h, v, H, W, c.bitmap_words, B, cr, pr, cal, threshold, c.bitmap, c.flags, c.n_flags);
  else
    box_bitmap_kernel<kModeCloud64><<<g, 256, 0, st>>>(h, v, H, W, c.bitmap_words, B, cr, pr, cal, threshold, c.bitmap, c.flags, c.n_flags);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return launch_offsets_from_bitmap(B, H, W, ws, counts, st);
}

uint32_t* workspace_bitmap(void* ws, int B, int H, int W, long long* words_per_scan, unsigned long long** flags,
                           int* n_flags) {
  const CloudWorkspace c = carve_workspace(ws, B, H, W);
  *words_per_scan = c.bitmap_words;
  *flags = c.flags;
  *n_flags = c.n_flags;
  return c.bitmap;
}

constexpr int kEmitSmem = (int)(sizeof(EmitTile<true>) * kEmitWarps);       // pairs mode
constexpr int kEmitSmemCloud = (int)(sizeof(EmitTile<false>) * kEmitWarps);

// > 48 KB of dynamic shared memory needs an opt-in per kernel: once per device, from sl_capi.cu's device gate
int cloud_device_init() {
  SL_CUDA(cudaFuncSetAttribute(emit_kernel<kModePairs>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmitSmem));
  SL_CUDA(cudaFuncSetAttribute(emit_kernel<kModeCloud32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmitSmem));
  SL_CUDA(cudaFuncSetAttribute(emit_kernel<kModeCloud64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmitSmem));
  SL_CUDA(cudaFuncSetAttribute(emit_kernel<kModeTableBT32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmitSmem));
  SL_CUDA(cudaFuncSetAttribute(emit_kernel<kModeTable64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmitSmem));
  SL_CUDA(cudaFuncSetAttribute(emit_fast_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast_smem(16)));
  SL_CUDA(cudaFuncSetAttribute(emit_fast_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast_smem(32)));
  return kOk;
}

static int check_emit_size(const EmitArgs& a) {
  if (a.total_tiles >= (1LL << 31)) { set_error("batch too large for one emit launch"); return kErrBadArg; }
  return kOk;
}

static EmitArgs make_emit(const int16_t* h, const int16_t* v, const uint8_t* white, int B, int H, int W,
                          const void* ws, float* out, uint32_t* out_rgb32) {
  const CloudWorkspace c = carve_workspace(const_cast<void*>(ws), B, H, W);
  EmitArgs a;
  a.h = h; a.v = v; a.white = white; a.band_pre = c.band_cnt; a.col_bits = c.col_bits; a.col_base = c.col_base;
  a.out = out; a.out_rgb32 = out_rgb32; a.H = H; a.W = W; a.bands = c.bands;
  a.cam_rays = nullptr; a.proj_rays = nullptr;
  a.n_extra = 0; a.remote = 0;
  for (int k = 0; k < kMaxWorld - 1; ++k) a.out_extra[k] = nullptr;
  // vector path: 8-byte index loads, 4-byte colour loads, 16-byte record stores (cam_rays is checked by the caller)
  a.aligned = ((uintptr_t)h % 8 == 0) && ((uintptr_t)v % 8 == 0) && ((uintptr_t)white % 4 == 0);
  a.off32 = ((long long)B * H * W < (1LL << 32)) ? 1 : 0;
  a.tiles_x = (W + kEmitCols - 1) / kEmitCols;
  a.total_tiles = (long long)a.tiles_x * a.bands * B;      // callers reject >= 2^31 (check_emit_size)
  return a;
}

int launch_emit_pairs(const int16_t* h, const int16_t* v, const uint8_t* white, int B, int H, int W,
                      int proj_w, int proj_h, const void* ws, float* pairs, uint32_t* rgb32, cudaStream_t st) {
  EmitArgs a = make_emit(h, v, white, B, H, W, ws, pairs, rgb32);
  a.cal = CalibDev();
  a.cal.proj_w = proj_w; a.cal.proj_h = proj_h; a.cal.row0 = 0;
  int rc = check_emit_size(a);
  if (rc) return rc;
  emit_kernel<kModePairs><<<grid_for(a.total_tiles, kEmitWarps), kEmitWarps * 32, kEmitSmem, st>>>(a);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_lens_tables(const SlCalib& calib, int cam_w, int cam_h, int precision, float* cam_rays, float* proj_table,
                       cudaStream_t st) {
  const CalibDev cal = make_calib(calib);
  auto blocks = [](long long n) { long long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); };
  lens_table_kernel<false><<<blocks((long long)cam_w * cam_h), 256, 0, st>>>(cal.cam_d, cam_w, cam_h, (float2*)cam_rays,
                                                                             0.f, 0.f, 0.f);
  count_launch();
  const int pb = blocks((long long)cal.proj_w * cal.proj_h);
  if (precision == SL_PREC_FP64)
    lens_table_kernel<false><<<pb, 256, 0, st>>>(cal.proj_d, cal.proj_w, cal.proj_h, (float2*)proj_table, 0.f, 0.f, 0.f);
  else
    lens_table_kernel<true><<<pb, 256, 0, st>>>(cal.proj_d, cal.proj_w, cal.proj_h, (float2*)proj_table, cal.Tf[0],
                                                cal.Tf[1], cal.Tf[2]);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_emit_cloud(const int16_t* h, const int16_t* v, const uint8_t* white, const SlCalib& calib,
                      int precision, const float* cam_rays, const float* proj_rays, int B, int H, int W,
                      const void* ws, float* cloud, cudaStream_t st, float* const* extra, int n_extra, int remote,
                      long long max_points) {
  EmitArgs a = make_emit(h, v, white, B, H, W, ws, cloud, nullptr);
  a.cal = make_calib(calib);
  a.cam_rays = (const float2*)cam_rays; a.proj_rays = (const float2*)proj_rays;
  if (n_extra < 0 || n_extra > kMaxWorld - 1) { set_error("at most %d clouds per emission", kMaxWorld); return kErrBadArg; }
  for (int k = 0; k < n_extra; ++k) a.out_extra[k] = extra[k];
  a.n_extra = n_extra; a.remote = remote;
  if (max_points > 0) a.off32 = (max_points < (1LL << 32)) ? 1 : 0;     // offsets rebased by the multi-GPU exchange
  const int g = grid_for(a.total_tiles, kEmitWarps);
  const int thr = kEmitWarps * 32;
  const int rc = check_emit_size(a);
  if (rc) return rc;
  if (cam_rays && precision == SL_PREC_FP32 && (W % kEmitCols) == 0 && a.aligned) {      // the throughput path
    const bool peer = remote || n_extra > 0;
    a.total_tiles *= peer ? 1 : 2;                       // whole bands for peer-mapped clouds, half bands otherwise
    if (a.total_tiles >= (1LL << 31)) { set_error("batch too large for one emit launch"); return kErrBadArg; }
    const int gf = grid_for(a.total_tiles, kEmitWarps);
    if (peer) emit_fast_kernel<32, true><<<gf, thr, fast_smem(32), st>>>(a);
    else SL_CUDA(launch_stage2(emit_fast_kernel<16, false>, gf, thr, (size_t)fast_smem(16), st, a));
    count_launch();
    SL_CUDA(cudaGetLastError());
    return kOk;
  }
  if (cam_rays) {
    if (precision == SL_PREC_FP64) emit_kernel<kModeTable64><<<g, thr, kEmitSmemCloud, st>>>(a);
    else SL_CUDA(launch_stage2(emit_kernel<kModeTableBT32>, g, thr, (size_t)kEmitSmemCloud, st, a));
  } else if (precision == SL_PREC_FP64) emit_kernel<kModeCloud64><<<g, thr, kEmitSmemCloud, st>>>(a);
  else emit_kernel<kModeCloud32><<<g, thr, kEmitSmemCloud, st>>>(a);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

// dst[c][r] = src[r][c] for a table of float2 (32 x 32 tiles through shared memory); once per calibration
__global__ void __launch_bounds__(256) transpose_f2_kernel(const float2* __restrict__ src, int rows, int cols,
                                                            float2* __restrict__ dst) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8 threads
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = src[(size_t)(r0 + i) * cols + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) dst[(size_t)(c0 + i) * rows + r0 + tx] = tile[tx][i];
}

int launch_transpose_f2(const float* src, int rows, int cols, float* dst, cudaStream_t st) {
  const dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  if (grid.y > 65535u) { set_error("table too tall to transpose"); return kErrBadArg; }
  transpose_f2_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(src), rows, cols, reinterpret_cast<float2*>(dst));
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_emit_rows(const int16_t* h, const int16_t* v, const uint8_t* white, const SlCalib& calib, const float* cam_t,
                     int cam_rows, const float* proj_t, int B, int H, int W, const void* ws, float* cloud, cudaStream_t st,
                     float* const* extra, int n_extra, int remote) {
  if (W % 32) { set_error("the row-mapped emission needs W to be a multiple of 32 (got %d)", W); return kErrBadArg; }
  if (calib.cam_row0 < 0 || calib.cam_row0 + H > cam_rows) { set_error("camera table has %d rows, the band needs %d", cam_rows, calib.cam_row0 + H); return kErrBadArg; }
  const CloudWorkspace c = carve_workspace(const_cast<void*>(ws), B, H, W);
  const CalibDev cal = make_calib(calib);
  EmitRowsArgs a;
  a.h = h; a.v = v; a.white = white; a.band_pre = c.band_cnt; a.col_bits = c.col_bits; a.col_base = c.col_base;
  a.out = reinterpret_cast<uint4*>(cloud);
  if (n_extra < 0 || n_extra > kMaxWorld - 1) { set_error("at most %d clouds per emission", kMaxWorld); return kErrBadArg; }
  for (int k = 0; k < kMaxWorld - 1; ++k) a.out_extra[k] = k < n_extra ? reinterpret_cast<uint4*>(extra[k]) : nullptr;
  a.n_extra = n_extra;
  a.cam_t = reinterpret_cast<const float2*>(cam_t); a.proj_t = reinterpret_cast<const float2*>(proj_t);
  a.H = H; a.W = W; a.bands = c.bands; a.tiles_x = W / 32; a.cam_rows = cam_rows; a.row0 = cal.row0;
  a.pw1 = (uint32_t)(cal.proj_w - 1); a.ph1 = (uint32_t)(cal.proj_h - 1); a.ph = (uint32_t)cal.proj_h;
  a.t0 = cal.Tf[0]; a.t1 = cal.Tf[1]; a.t2 = cal.Tf[2]; a.tl2 = cal.t_len2f;
  a.total_tiles = (long long)B * a.bands * a.tiles_x;
  if (a.total_tiles >= (1LL << 31)) { set_error("batch too large for one emit launch"); return kErrBadArg; }
  const int grid = (int)((a.total_tiles + kRowsWarps - 1) / kRowsWarps);
  if (remote || n_extra > 0) emit_rows_kernel<true><<<grid, kRowsWarps * 32, kRowsSmem, st>>>(a);
  else SL_CUDA(launch_stage2(emit_rows_kernel<false>, grid, kRowsWarps * 32, (size_t)kRowsSmem, st, a));
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

template <bool REC>
static int filter_impl(const void* pts, const double* colors, long long N, double threshold, void* out_pts,
                       double* out_colors, long long* out_count, void* ws, cudaStream_t st) {
  const long long blocks = (N + kFilterBlock - 1) / kFilterBlock;
  uint32_t* blk_cnt = static_cast<uint32_t*>(ws);
  long long* blk_base = reinterpret_cast<long long*>(static_cast<uint8_t*>(ws) + align_up((size_t)(blocks + 1) * 4, 256));
  if (blocks > 0) {
    filter_count_kernel<REC><<<(int)blocks, 256, 0, st>>>(pts, N, threshold, blk_cnt);
    count_launch();
  }
  // B = 0: the scan writes only counts[0] = total kept
  colscan_kernel<<<1, 1024, 0, st>>>(blk_cnt, blk_base, blocks, 0, 1, out_count);
  count_launch();
  if (blocks > 0) {
    filter_emit_kernel<REC><<<(int)blocks, 256, 0, st>>>(pts, colors, N, threshold, blk_base, out_pts, out_colors);
    count_launch();
  }
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_filter_box(const double* pts, const double* colors, long long N, double threshold,
                      double* out_pts, double* out_colors, long long* out_count, void* ws,
                      cudaStream_t st) {
  return filter_impl<false>(pts, colors, N, threshold, out_pts, out_colors, out_count, ws, st);
}

int launch_filter_cloud(const float* cloud, long long N, double threshold, float* out_cloud,
                        long long* out_count, void* ws, cudaStream_t st) {
  return filter_impl<true>(cloud, nullptr, N, threshold, out_cloud, nullptr, out_count, ws, st);
}

}  // namespace sl

// Statistical outlier removal on the packed XYZRGB cloud (sm_100a) -- the clean-up the reference applies
// right after the hot path: pcd.remove_statistical_outlier(nb_neighbors=20, std_ratio=0.5),
// scanner/utils/visualize.py:104-105 (Open3D 0.17 PointCloud::RemoveStatisticalOutliers).
//
// Exact k-nearest-neighbour mean distances on a hashed uniform grid, no KD-tree:
//   * cells of side s are hashed (64-bit cell id -> open-addressing table of >= 2N slots), points are
//     counting-sorted by cell (atomics only; the order inside a cell does not influence any result
//     because each query keeps its k smallest squared distances SORTED and sums them nearest first);
//   * a query scans the 3x3x3 block of cells around its own cell (nearest cells first, cells beyond the current
//     k-th best skipped) and is final when its k-th distance is not larger than the distance to the faces of
//     that block (nothing outside can be closer); a near miss scans the 5x5x5 shell at the same level;
//   * the few queries that are still not final (sparse noise points) are retried on a grid with 4x larger
//     cells, up to kLevels times (cut into parts over several CTAs there: the block of a sparse query can hold a
//     large part of the cloud); the last level covers the whole bounding box with <= 2 cells per axis.
//   s of level 0 is chosen from a pilot build so that occupied cells hold ~kTargetOcc points.
// Distances: float32 coordinates (the reference narrows to float32, visualize.py:100), differences,
// squares and the (dx^2 + dy^2) + dz^2 sum in float64 with one rounding per operation, sqrt and the mean in
// float64 -- the per-point mean distances are bit-identical to the oracle's.  The cloud mean / standard
// deviation are fixed-order tree reductions (deterministic; last-ulp differences from a sequential sum).
// Everything is stream-ordered: no host synchronisation, work sizes live in device memory.
#include <cuda_runtime.h>
#include <math.h>
#include "sl_common.cuh"
#include "sl_kernels.h"

namespace sl {

constexpr int kLevels = 12;             // cell side s0 * 4^level; s0 >= extent / 2^20, so level 11 has <= 2 cells per axis
#ifndef SL_SOR_TARGET_OCC
#define SL_SOR_TARGET_OCC 8.0           // measured on the 1.8 M-point cloud of a 1080p scan: 5.4 / 4.8 / 4.4 / 4.3 / 3.8 / 4.3 / 4.0 / 4.4 ms
#endif                                  // for 3 / 4 / 5 / 6 / 8 / 10 / 12 / 16 (larger cells: fewer probes, more coherent warps;
constexpr double kTargetOcc = SL_SOR_TARGET_OCC;   // the box test and the float32 filter discard the extra candidates)
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kSorBlocks = 148 * 4;     // grid of the grid-stride kernels
constexpr int kRedBlocks = 256;         // partial sums of the two reductions
constexpr int kCompactBlock = 1024;

struct SorParams {
  uint32_t bb[6];                  // order-preserving uint images of min x,y,z / max x,y,z
  uint32_t n_finite;
  uint32_t occupied;               // distinct cells of the pilot build
  uint32_t cursor;                 // range allocator of the counting sort
  uint32_t counts[kLevels + 1];    // queries pending at each level
  double org[3], ext[3];
  double cell[kLevels];            // cell side per level
  int dim[kLevels][3];
  double mean, stddev, thr;
};

struct SorWs {
  SorParams* prm;
  unsigned long long* keys;   // [M]
  uint32_t* cnt;              // [M]
  uint32_t* start;            // [M]  after the scatter: END of the cell's range in `sorted`
  uint32_t* slot;             // [N]
  float4* sorted;             // [N]  xyz + original index
  double* avg;                // [N]
  uint32_t* pend[2];          // [N]
  double* part;               // [2 * kRedBlocks]
  uint32_t* blk_cnt;          // [blocks + 1]
  long long* blk_base;        // [blocks + 1]
  double* part_best;          // [kPartSlots * 128] sorted k-best lists of the query parts (sor_query_part_kernel)
  uint32_t* q_total;          // [kPartSlots] candidates of each sparse query
  uint32_t M;
};
constexpr size_t kPartSlots = 8192 + 148 * 8;      // >= queries x parts on any level handled by the part kernel

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
static uint32_t table_size(long long N) {
  uint64_t m = 1024;
  while (m < (uint64_t)N * 2) m <<= 1;
  return (uint32_t)m;
}

static size_t sor_carve(void* base, long long N, SorWs* w) {
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  const uint32_t M = table_size(N);
  const size_t n = (size_t)(N < 1 ? 1 : N);
  const size_t blocks = (n + kCompactBlock - 1) / kCompactBlock + 1;
  auto take = [&](size_t bytes) { uint8_t* q = p ? p + off : nullptr; off += up256(bytes); return q; };
  SorWs s;
  s.prm = (SorParams*)take(sizeof(SorParams));
  s.keys = (unsigned long long*)take((size_t)M * 8);
  s.cnt = (uint32_t*)take((size_t)M * 4);
  s.start = (uint32_t*)take((size_t)M * 4);
  s.slot = (uint32_t*)take(n * 4);
  s.sorted = (float4*)take(n * 16);
  s.avg = (double*)take(n * 8);
  s.pend[0] = (uint32_t*)take(n * 4);
  s.pend[1] = (uint32_t*)take(n * 4);
  s.part = (double*)take(2 * kRedBlocks * 8);
  s.blk_cnt = (uint32_t*)take(blocks * 4);
  s.blk_base = (long long*)take(blocks * 8);
  s.part_best = (double*)take(kPartSlots * 128 * 8);
  s.q_total = (uint32_t*)take(kPartSlots * 4);
  s.M = M;
  if (w) *w = s;
  return off;
}

size_t outlier_workspace_bytes(long long N) { return sor_carve(nullptr, N, nullptr); }

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t ord_u32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_f32(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ bool finite3(const float4& q) { return isfinite(q.x) && isfinite(q.y) && isfinite(q.z); }

__device__ __forceinline__ int cell_of(double u, double inv, int dim) {
  int c = (int)floor(u * inv);
  return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}
__device__ __forceinline__ unsigned long long cell_key(int cx, int cy, int cz) {
  return (unsigned long long)cx | ((unsigned long long)cy << 21) | ((unsigned long long)cz << 42);
}
__device__ __forceinline__ uint32_t hash_key(unsigned long long k, uint32_t mask) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k & mask;
}

// Exact squared distance (float32 coordinates widened to float64, one rounding per operation) and its cheap
// float32 estimate.  The estimate is within 1e-6 relative of the exact value (the inputs are exact, every float32
// operation adds 2^-24), so `est > bound_up(limit)` proves `exact > limit` and the float64 path -- three
// conversions and eight fp64 operations per candidate, all on slow pipes -- runs only for candidates that may count.
__device__ __forceinline__ double dist2_exact(const float4& p, double qx, double qy, double qz) {
  const double ex = (double)p.x - qx, ey = (double)p.y - qy, ez = (double)p.z - qz;
  return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
}
__device__ __forceinline__ float dist2_est(const float4& p, const float4& q) {
  const float ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
  return ex * ex + ey * ey + ez * ez;
}
__device__ __forceinline__ float bound_up(double limit) { return __double2float_ru(limit * 1.00001); }

// ------------------------------------------------------------------ planning
__global__ void sor_init_kernel(SorParams* prm) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { prm->bb[i] = 0xffffffffu; prm->bb[3 + i] = 0u; }
    prm->n_finite = 0; prm->occupied = 0; prm->cursor = 0;
    for (int i = 0; i <= kLevels; ++i) prm->counts[i] = 0;
  }
}

__global__ void __launch_bounds__(256) sor_bbox_kernel(const float4* __restrict__ cloud, long long N, SorParams* prm,
                                                        double* __restrict__ avg) {
  uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u}, nf = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const float4 q = __ldg(cloud + i);
    if (!finite3(q)) { avg[i] = -1.0; continue; }     // dropped: never a neighbour, never an inlier
    const uint32_t u[3] = {ord_u32(q.x), ord_u32(q.y), ord_u32(q.z)};
    for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], u[a]); hi[a] = max(hi[a], u[a]); }
    ++nf;
  }
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  nf = __reduce_add_sync(0xffffffffu, nf);
  if ((threadIdx.x & 31) == 0 && nf) {
    for (int a = 0; a < 3; ++a) { atomicMin(&prm->bb[a], lo[a]); atomicMax(&prm->bb[3 + a], hi[a]); }
    atomicAdd(&prm->n_finite, nf);
  }
}

__device__ void sor_set_levels(SorParams* prm, double s0) {
  double lmax = fmax(prm->ext[0], fmax(prm->ext[1], prm->ext[2]));
  const double floor_s = lmax / 1048576.0;           // <= 2^20 cells per axis (21-bit key fields)
  if (!(s0 > floor_s)) s0 = floor_s;
  if (!(s0 > 0.0)) s0 = 1.0;                          // all points coincide
  for (int l = 0; l < kLevels; ++l) {
    prm->cell[l] = s0;
    for (int a = 0; a < 3; ++a) prm->dim[l][a] = (int)floor(prm->ext[a] / s0) + 1;
    s0 *= 4.0;
  }
}

// pilot cell side from the bounding box: a surface of area (two largest extents) holding n_finite points
__global__ void sor_plan_pilot_kernel(SorParams* prm) {
  if (threadIdx.x) return;
  const uint32_t nf = prm->n_finite;
  if (nf == 0) { for (int a = 0; a < 3; ++a) { prm->org[a] = 0; prm->ext[a] = 0; } sor_set_levels(prm, 1.0); return; }
  double e[3];
  for (int a = 0; a < 3; ++a) {
    const double lo = (double)ord_f32(prm->bb[a]), hi = (double)ord_f32(prm->bb[3 + a]);
    prm->org[a] = lo; prm->ext[a] = e[a] = hi - lo;
  }
  const double l1 = fmax(e[0], fmax(e[1], e[2])), l3 = fmin(e[0], fmin(e[1], e[2]));
  const double l2 = e[0] + e[1] + e[2] - l1 - l3;
  double s = sqrt(kTargetOcc * l1 * l2 / (double)nf);
  if (!(s > 0.0)) s = kTargetOcc * l1 / (double)nf;
  sor_set_levels(prm, s);
}

// level-0 cell side from the pilot's occupancy (points per occupied cell scale ~ s^2 on a surface)
__global__ void sor_plan_final_kernel(SorParams* prm) {
  if (threadIdx.x) return;
  const uint32_t nf = prm->n_finite;
  if (nf == 0) return;
  const double occ = (double)nf / (double)(prm->occupied ? prm->occupied : 1u);
  sor_set_levels(prm, prm->cell[0] * sqrt(kTargetOcc / occ));
  prm->counts[0] = nf;
}

// ------------------------------------------------------------------ grid build (one level)
__global__ void __launch_bounds__(256) sor_clear_kernel(SorParams* prm, int level, unsigned long long* __restrict__ keys,
                                                         uint32_t* __restrict__ cnt, uint32_t M) {
  if (level >= 0 && prm->counts[level] == 0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) prm->cursor = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    keys[i] = kEmptyKey;
    cnt[i] = 0;
  }
}

__global__ void __launch_bounds__(256) sor_insert_kernel(const float4* __restrict__ cloud, long long N, SorParams* prm,
                                                          int level, unsigned long long* keys, uint32_t* cnt,
                                                          uint32_t mask, uint32_t* __restrict__ slot) {
  const int lv = level < 0 ? 0 : level;
  if (level >= 0 && prm->counts[level] == 0) return;
  const double inv = 1.0 / prm->cell[lv];
  const double ox = prm->org[0], oy = prm->org[1], oz = prm->org[2];
  const int dx = prm->dim[lv][0], dy = prm->dim[lv][1], dz = prm->dim[lv][2];
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count (i - lane is the same for all lanes): the lanes of a warp that fall into the same cell
  // -- nearly all of them on the coarse levels -- share one atomic
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i - lane < N; i += (long long)gridDim.x * blockDim.x) {
    const float4 q = i < N ? __ldg(cloud + i) : make_float4(NAN, 0.f, 0.f, 0.f);
    const bool valid = finite3(q);
    uint32_t h = 0x80000000u | (uint32_t)lane;                       // never equal to a slot or to another lane's
    if (valid) {
      const unsigned long long key = cell_key(cell_of((double)q.x - ox, inv, dx), cell_of((double)q.y - oy, inv, dy),
                                              cell_of((double)q.z - oz, inv, dz));
      h = hash_key(key, mask);
      for (;;) {
        if (*reinterpret_cast<volatile unsigned long long*>(&keys[h]) == key) break;   // common case: no atomic at all
        const unsigned long long old = atomicCAS(&keys[h], kEmptyKey, key);
        if (old == kEmptyKey) { if (level < 0) atomicAdd(&prm->occupied, 1u); break; }
        if (old == key) break;
        h = (h + 1) & mask;
      }
    }
    if (level >= 0) {
      const uint32_t peers = __match_any_sync(0xffffffffu, h);
      if (valid) {
        if (lane == __ffs(peers) - 1) atomicAdd(&cnt[h], (uint32_t)__popc(peers));
        slot[i] = h;
      }
    }
  }
}

__global__ void __launch_bounds__(256) sor_alloc_kernel(SorParams* prm, int level, const uint32_t* __restrict__ cnt,
                                                         uint32_t* __restrict__ start, uint32_t M) {
  if (prm->counts[level] == 0) return;
  const int lane = threadIdx.x & 31;
  // one range allocation per warp (a warp-wide prefix of the cell sizes), not one atomic per occupied cell
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i - lane < M; i += gridDim.x * blockDim.x) {
    const uint32_t c = i < M ? cnt[i] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    uint32_t base = 0;
    if (lane == 31 && incl) base = atomicAdd(&prm->cursor, incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (c) start[i] = base + incl - c;
  }
}

__global__ void __launch_bounds__(256) sor_scatter_kernel(const float4* __restrict__ cloud, long long N, SorParams* prm,
                                                           int level, const uint32_t* __restrict__ slot,
                                                           uint32_t* start, float4* __restrict__ sorted) {
  if (prm->counts[level] == 0) return;
  const int lane = threadIdx.x & 31;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i - lane < N; i += (long long)gridDim.x * blockDim.x) {
    float4 q = i < N ? __ldg(cloud + i) : make_float4(NAN, 0.f, 0.f, 0.f);
    const bool valid = finite3(q);
    const uint32_t h = valid ? slot[i] : (0x80000000u | (uint32_t)lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, h);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(&start[h], (uint32_t)__popc(peers));   // start[] ends up as the END
    base = __shfl_sync(0xffffffffu, base, leader);                                       // of each cell's range
    if (valid) {
      q.w = __uint_as_float((uint32_t)i);
      sorted[base + __popc(peers & ((1u << lane) - 1u))] = q;
    }
  }
}

// ------------------------------------------------------------------ k-NN query, level 0: one thread per point
// The k best squared distances of a thread live in a shared-memory column (best[j] at sm[j * blockDim + tid]:
// 64-bit accesses, bank pair fixed by tid, so lanes inserting at different depths never conflict); the current
// k-th best stays in a register, so a rejected candidate costs no memory access.  (In local memory the divergent
// insertions touched a different cache line per lane and the kernel ran at L2 speed: 6.1 ms vs 0.x ms per 1.8 M points.)
// The 27 cells are visited nearest first (own cell, 6 face, 12 edge, 8 corner neighbours) and a cell is skipped once
// k candidates are held and the cell's box is farther than the k-th best: most of the final neighbours are found in
// the first cells, so the later ones are mostly rejected by one compare or skipped whole.
__constant__ int8_t kCellOrder[27][3] = {
    {0, 0, 0},
    {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1},
    {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0}, {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1}, {1, 0, 1},
    {0, -1, -1}, {0, 1, -1}, {0, -1, 1}, {0, 1, 1},
    {-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, 1}, {1, 1, 1}};

constexpr int kQ0Threads = 128;
constexpr uint32_t kThreadModeMin = 8192;   // levels >= 1 with at least this many queries keep one thread per query
__global__ void __launch_bounds__(kQ0Threads) sor_query_kernel(const float4* __restrict__ cloud, SorParams* prm, int level,
                                                                const float4* __restrict__ sorted,
                                                                const unsigned long long* __restrict__ keys,
                                                                const uint32_t* __restrict__ cnt,
                                                                const uint32_t* __restrict__ start, uint32_t mask, int k,
                                                                const uint32_t* __restrict__ pend_in,
                                                                uint32_t* __restrict__ pend_out, double* __restrict__ avg) {
  extern __shared__ double q0_best[];
  const uint32_t nq = prm->counts[level];
  if (nq == 0 || (level > 0 && nq < kThreadModeMin)) return;
  double* best = q0_best + threadIdx.x;
  constexpr int kS = kQ0Threads;                 // stride between a thread's entries
  const double s = prm->cell[level], inv = 1.0 / s;
  const double ox = prm->org[0], oy = prm->org[1], oz = prm->org[2];
  const int dimx = prm->dim[level][0], dimy = prm->dim[level][1], dimz = prm->dim[level][2];
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nq; t += gridDim.x * blockDim.x) {
    float4 q;
    uint32_t orig;
    if (level == 0) { q = sorted[t]; orig = __float_as_uint(q.w); }   // cell order: neighbouring threads scan the same cells
    else { orig = pend_in[t]; q = __ldg(cloud + orig); }
    const double qx = (double)q.x, qy = (double)q.y, qz = (double)q.z;
    const double ux = qx - ox, uy = qy - oy, uz = qz - oz;
    const int cx = cell_of(ux, inv, dimx), cy = cell_of(uy, inv, dimy), cz = cell_of(uz, inv, dimz);

    int found = 0;
    double worst = INFINITY;                      // best[k-1] once k candidates are held
    float worst_f = INFINITY;                     // its float32 image, rounded up
    // distance from the query to the neighbour cells along each axis (0 for the own cell), shortened by a margin
    // that covers the rounding of the cell assignment
    const double marg = 1e-6 * s;
    const double lox = fmax(ux - (double)cx * s - marg, 0.0), hix = fmax((double)(cx + 1) * s - ux - marg, 0.0);
    const double loy = fmax(uy - (double)cy * s - marg, 0.0), hiy = fmax((double)(cy + 1) * s - uy - marg, 0.0);
    const double loz = fmax(uz - (double)cz * s - marg, 0.0), hiz = fmax((double)(cz + 1) * s - uz - marg, 0.0);
    // one cell at offset (ax, ay, az), |a| <= 2, from the query's cell
    auto gap = [&](int a, double lo, double hi) {
      return a < 0 ? lo + (double)(-a - 1) * s : (a > 0 ? hi + (double)(a - 1) * s : 0.0);
    };
    auto scan_cell = [&](int ax, int ay, int az) {
      const int x = cx + ax, y = cy + ay, z = cz + az;
      if (x < 0 || x >= dimx || y < 0 || y >= dimy || z < 0 || z >= dimz) return;
      const double gx = gap(ax, lox, hix), gy = gap(ay, loy, hiy), gz = gap(az, loz, hiz);
      if (gx * gx + gy * gy + gz * gz > worst) return;        // nothing in that cell can enter the k best
      const unsigned long long key = cell_key(x, y, z);
      uint32_t h = hash_key(key, mask);
      unsigned long long kk;
      while ((kk = keys[h]) != key && kk != kEmptyKey) h = (h + 1) & mask;
      if (kk != key) return;
      const uint32_t end = start[h];
      for (uint32_t j = end - cnt[h]; j < end; ++j) {
        const float4 p = sorted[j];
        if (dist2_est(p, q) > worst_f) continue;
        const double d2 = dist2_exact(p, qx, qy, qz);
        if (!(d2 < worst)) continue;
        int pos = found < k ? found++ : k - 1;
        while (pos > 0) {
          const double above = best[(pos - 1) * kS];
          if (!(above > d2)) break;
          best[pos * kS] = above;
          --pos;
        }
        best[pos * kS] = d2;
        if (found == k) { worst = best[(k - 1) * kS]; worst_f = bound_up(worst); }
      }
    };
    // distance from the query to the faces of the (2r+1)^3 block that have unscanned cells behind them
    auto reach = [&](int r) {
      double g = INFINITY;
      if (cx - r > 0) g = fmin(g, ux - (double)(cx - r) * s);
      if (cx + r < dimx - 1) g = fmin(g, (double)(cx + r + 1) * s - ux);
      if (cy - r > 0) g = fmin(g, uy - (double)(cy - r) * s);
      if (cy + r < dimy - 1) g = fmin(g, (double)(cy + r + 1) * s - uy);
      if (cz - r > 0) g = fmin(g, uz - (double)(cz - r) * s);
      if (cz + r < dimz - 1) g = fmin(g, (double)(cz + r + 1) * s - uz);
      return g - marg;                              // rounding of the cell assignment, with a wide margin
    };

    for (int c = 0; c < 27; ++c) scan_cell(kCellOrder[c][0], kCellOrder[c][1], kCellOrder[c][2]);
    double g = reach(1);
    bool final_ = isinf(g) || (found >= k && g > 0.0 && worst <= g * g);
    if (!final_ && found >= k) {
      // The k-th neighbour lies just outside the block (a query close to a face of its cell).  If it is within reach
      // of the 5x5x5 block, scanning that shell settles the query at this level -- the k-th distance can only shrink
      // -- instead of sending it to a grid with 4x larger cells; nearly all of its 98 cells fail the box test.
      const double g2 = reach(2);
      if (isinf(g2) || (g2 > 0.0 && worst <= g2 * g2)) {
        for (int az = -2; az <= 2; ++az)
          for (int ay = -2; ay <= 2; ++ay)
            for (int ax = -2; ax <= 2; ++ax)
              if (max(max(abs(ax), abs(ay)), abs(az)) == 2) scan_cell(ax, ay, az);
        final_ = true;
      }
    }
    final_ = final_ || level == kLevels - 1;
    if (final_) {
      double sum = 0.0;
      for (int j = 0; j < found; ++j) sum = __dadd_rn(sum, sqrt(best[j * kS]));      // nearest first
      avg[orig] = found ? sum / (double)found : -1.0;
    } else {
      pend_out[atomicAdd(&prm->counts[level + 1], 1u)] = orig;
    }
  }
}

// Levels >= 1 with few queries left (< kThreadModeMin).  The queries that get this far are the sparse ones (noise
// far from the surface); their 3x3x3 block can hold a large part of the cloud, so the candidates of ONE query are
// spread over the threads of a CTA and, when there are fewer queries than CTAs, over several CTAs ("parts",
// contiguous slices of the candidate list).  Keeping a sorted k-best list per thread would make nearly every
// candidate step a divergent insertion, so a part is scanned twice without any list instead:
//   pass 1  every thread takes the minimum (float32 estimate) of its strided share; the k-th smallest of those
//           minima bounds the part's k-th distance from above (they are k distinct candidates);
//   pass 2  the candidates under the bound -- a few dozen -- get their exact distance and go to a shared list.
// The k smallest of the list are drawn one at a time (block-wide minimum, the owner advances) and written, sorted,
// to the part's slot of a scratch array; sor_query_merge_kernel merges the parts of a query (nearest first) and
// applies the same finality test and summation as the thread-per-query kernel.  If the list overflows (thousands
// of exact ties at the bound) the CTA falls back to per-thread sorted lists, which is exact for any input.
constexpr int kQbThreads = 256;
constexpr int kQbBatch = 4;          // candidate loads in flight per thread
constexpr int kQbList = 4096;        // capacity of the shared candidate list
constexpr int kQbItems = 148 * 8;    // CTAs of the part kernel; a query is cut into up to kQbMaxParts parts
constexpr int kQbMaxParts = 64;
constexpr int kMaxNeighbors = 128;

__device__ __forceinline__ int parts_per_query(uint32_t nq) {
  const int g = kQbItems / (int)(nq ? nq : 1u);
  return g < 1 ? 1 : (g > kQbMaxParts ? kQbMaxParts : g);
}

// minimum over the CTA of (v, owner = tid); ties go to the lowest tid.  All threads get the result.
__device__ __forceinline__ void block_min(double& v, int& o, double* w_val, int* w_own) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double v2 = __shfl_down_sync(0xffffffffu, v, off);
    const int o2 = __shfl_down_sync(0xffffffffu, o, off);
    if (v2 < v || (v2 == v && o2 < o)) { v = v2; o = o2; }
  }
  if (lane == 0) { w_val[warp] = v; w_own[warp] = o; }
  __syncthreads();
  v = w_val[0]; o = w_own[0];
#pragma unroll
  for (int w = 1; w < kQbThreads / 32; ++w)
    if (w_val[w] < v || (w_val[w] == v && w_own[w] < o)) { v = w_val[w]; o = w_own[w]; }
  __syncthreads();
}

template <int KMAX>
__global__ void __launch_bounds__(kQbThreads) sor_query_part_kernel(
    const float4* __restrict__ cloud, const SorParams* prm, int level, const float4* __restrict__ sorted,
    const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ start,
    uint32_t mask, int k, const uint32_t* __restrict__ pend_in, double* __restrict__ part_best,
    uint32_t* __restrict__ q_total) {
  const uint32_t nq = prm->counts[level];
  if (nq == 0 || nq >= kThreadModeMin) return;
  __shared__ uint32_t c_begin[27], c_pre[28];      // candidate ranges of the 27 cells, prefix of their sizes
  __shared__ double w_val[kQbThreads / 32];
  __shared__ int w_own[kQbThreads / 32];
  __shared__ double list[kQbList];
  __shared__ unsigned int list_n;
  const double inv = 1.0 / prm->cell[level];
  const double ox = prm->org[0], oy = prm->org[1], oz = prm->org[2];
  const int dimx = prm->dim[level][0], dimy = prm->dim[level][1], dimz = prm->dim[level][2];
  const int tid = threadIdx.x;
  const int G = parts_per_query(nq);
  for (uint32_t item = blockIdx.x; item < nq * (uint32_t)G; item += gridDim.x) {
    const uint32_t t = item / (uint32_t)G;
    const int part = (int)(item - t * (uint32_t)G);
    const float4 q = __ldg(cloud + pend_in[t]);
    const double qx = (double)q.x, qy = (double)q.y, qz = (double)q.z;
    const int cx = cell_of(qx - ox, inv, dimx), cy = cell_of(qy - oy, inv, dimy), cz = cell_of(qz - oz, inv, dimz);
    if (tid < 27) {
      const int x = cx + tid % 3 - 1, y = cy + (tid / 3) % 3 - 1, z = cz + tid / 9 - 1;
      uint32_t n = 0, b0 = 0;
      if (x >= 0 && x < dimx && y >= 0 && y < dimy && z >= 0 && z < dimz) {
        const unsigned long long key = cell_key(x, y, z);
        uint32_t h = hash_key(key, mask);
        unsigned long long kk;
        while ((kk = keys[h]) != key && kk != kEmptyKey) h = (h + 1) & mask;
        if (kk == key) { n = cnt[h]; b0 = start[h] - n; }
      }
      c_begin[tid] = b0;
      c_pre[tid + 1] = n;
    }
    if (tid == 32) list_n = 0;
    __syncthreads();
    if (tid == 0) {
      c_pre[0] = 0;
      for (int i = 0; i < 27; ++i) c_pre[i + 1] += c_pre[i];
    }
    __syncthreads();
    const uint32_t total = c_pre[27];
    if (part == 0 && tid == 0) q_total[t] = total;
    const uint32_t slice = (total + (uint32_t)G - 1) / (uint32_t)G;
    const uint32_t lo = min(total, (uint32_t)part * slice), hi = min(total, lo + slice);

    // f(point) for every candidate of this thread's share of [lo, hi); the loads of a batch are issued together
    auto scan = [&](auto f) {
      int cell = 0;
      for (uint32_t j0 = lo + tid; j0 < hi; j0 += kQbThreads * kQbBatch) {
        float4 p[kQbBatch];
#pragma unroll
        for (int u = 0; u < kQbBatch; ++u) {
          const uint32_t j = j0 + u * kQbThreads;
          if (j < hi) {
            while (j >= c_pre[cell + 1]) ++cell;
            p[u] = sorted[c_begin[cell] + (j - c_pre[cell])];
          }
        }
#pragma unroll
        for (int u = 0; u < kQbBatch; ++u) {
          if (j0 + u * kQbThreads >= hi) break;
          f(p[u]);
        }
      }
    };

    float thr_f = INFINITY;                // small part: every candidate goes to the list
    if (hi - lo > kQbList) {
      float tmin = INFINITY;               // float32 estimates are enough for a bound
      scan([&](const float4& p) { tmin = fminf(tmin, dist2_est(p, q)); });
      double kth = INFINITY;
      for (int r = 0; r < k; ++r) {        // k-th smallest of the per-thread minima (k <= 128 < kQbThreads)
        double v = (double)tmin;
        int o = tid;
        block_min(v, o, w_val, w_own);
        if (o == tid) tmin = INFINITY;
        kth = v;
      }
      thr_f = bound_up(kth * 1.00001);     // estimate -> exact -> estimate: two margins
    }
    scan([&](const float4& p) {
      if (dist2_est(p, q) <= thr_f) {
        const unsigned int i = atomicAdd(&list_n, 1u);
        if (i < kQbList) list[i] = dist2_exact(p, qx, qy, qz);
      }
    });
    __syncthreads();
    const bool overflow = list_n > kQbList;

    double best[KMAX];                     // this thread's candidates for the final selection, sorted
    int found = 0;
    auto insert = [&](double d2, int cap) {
      int pos;
      if (found < cap) pos = found++;
      else if (d2 < best[cap - 1]) pos = cap - 1;
      else return;
      while (pos > 0 && best[pos - 1] > d2) { best[pos] = best[pos - 1]; --pos; }
      best[pos] = d2;
    };
    if (!overflow) {
      for (uint32_t i = tid; i < list_n; i += kQbThreads) insert(list[i], kQbList / kQbThreads);
    } else {
      scan([&](const float4& p) { insert(dist2_exact(p, qx, qy, qz), k); });
    }

    // k-way selection over the threads' sorted lists -> the part's k best, sorted (INFINITY-padded)
    double* out = part_best + (size_t)item * k;
    int head = 0;
    for (int r = 0; r < k; ++r) {
      double v = head < found ? best[head] : INFINITY;
      int o = tid;
      block_min(v, o, w_val, w_own);
      if (o == tid) { if (head < found) ++head; out[r] = v; }
    }
    __syncthreads();
  }
}

// one warp per query: merge the sorted k-best lists of its parts, then the finality test of the level
__global__ void __launch_bounds__(128) sor_query_merge_kernel(const float4* __restrict__ cloud, SorParams* prm, int level,
                                                               int k, const uint32_t* __restrict__ pend_in,
                                                               const double* __restrict__ part_best,
                                                               const uint32_t* __restrict__ q_total,
                                                               uint32_t* __restrict__ pend_out, double* __restrict__ avg) {
  const uint32_t nq = prm->counts[level];
  if (nq == 0 || nq >= kThreadModeMin) return;
  __shared__ double sel_all[4][kMaxNeighbors];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sel = sel_all[warp];
  const int G = parts_per_query(nq);
  const double s = prm->cell[level], inv = 1.0 / s;
  const int dimx = prm->dim[level][0], dimy = prm->dim[level][1], dimz = prm->dim[level][2];
  for (uint32_t t = blockIdx.x * 4 + warp; t < nq; t += gridDim.x * 4) {
    // lane l owns the lists of parts l and l + 32 (G <= 64)
    const double* la = part_best + ((size_t)t * G + lane) * k;
    const double* lb = part_best + ((size_t)t * G + lane + 32) * k;
    int ha = lane < G ? 0 : k, hb = lane + 32 < G ? 0 : k;
    const uint32_t total = q_total[t];
    const int want = total < (uint32_t)k ? (int)total : k;
    for (int r = 0; r < want; ++r) {
      const double va = ha < k ? la[ha] : INFINITY, vb = hb < k ? lb[hb] : INFINITY;
      double v = fmin(va, vb);
      int o = lane;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, off);
        const int o2 = __shfl_xor_sync(0xffffffffu, o, off);
        if (v2 < v || (v2 == v && o2 < o)) { v = v2; o = o2; }
      }
      if (o == lane) { if (va <= vb) ++ha; else ++hb; }
      if (lane == 0) sel[r] = v;
    }
    __syncwarp();
    if (lane == 0) {
      const uint32_t orig = pend_in[t];
      const float4 q = __ldg(cloud + orig);
      const double ux = (double)q.x - prm->org[0], uy = (double)q.y - prm->org[1], uz = (double)q.z - prm->org[2];
      const int cx = cell_of(ux, inv, dimx), cy = cell_of(uy, inv, dimy), cz = cell_of(uz, inv, dimz);
      double g = INFINITY;
      if (cx - 1 > 0) g = fmin(g, ux - (double)(cx - 1) * s);
      if (cx + 1 < dimx - 1) g = fmin(g, (double)(cx + 2) * s - ux);
      if (cy - 1 > 0) g = fmin(g, uy - (double)(cy - 1) * s);
      if (cy + 1 < dimy - 1) g = fmin(g, (double)(cy + 2) * s - uy);
      if (cz - 1 > 0) g = fmin(g, uz - (double)(cz - 1) * s);
      if (cz + 1 < dimz - 1) g = fmin(g, (double)(cz + 2) * s - uz);
      g -= 1e-6 * s;
      const bool final_ = isinf(g) || (want >= k && g > 0.0 && sel[k - 1] <= g * g) || level == kLevels - 1;
      if (final_) {
        double sum = 0.0;
        for (int j = 0; j < want; ++j) sum = __dadd_rn(sum, sqrt(sel[j]));
        avg[orig] = want ? sum / (double)want : -1.0;
      } else {
        pend_out[atomicAdd(&prm->counts[level + 1], 1u)] = orig;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ statistics (fixed-order reductions)
// pass 0: sum of the positive mean distances; pass 1: sum of (avg - mean)^2 over the same points
__global__ void __launch_bounds__(256) sor_partial_kernel(const double* __restrict__ avg, long long N,
                                                           const SorParams* prm, int pass, double* __restrict__ part) {
  __shared__ double sh[256];
  const double mean = pass ? prm->mean : 0.0;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const double a = avg[i];
    if (a > 0.0) acc += pass ? (a - mean) * (a - mean) : a;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[pass * kRedBlocks + blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(kRedBlocks) sor_final_kernel(SorParams* prm, int pass, const double* __restrict__ part,
                                                                double std_ratio, double* __restrict__ stats) {
  __shared__ double sh[kRedBlocks];
  sh[threadIdx.x] = part[pass * kRedBlocks + threadIdx.x];
  __syncthreads();
  for (int w = kRedBlocks / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x) return;
  const double nf = (double)prm->n_finite;
  if (pass == 0) { prm->mean = sh[0] / nf; return; }
  prm->stddev = sqrt(sh[0] / (nf - 1.0));                       // Bessel's correction, as Open3D
  prm->thr = prm->mean + std_ratio * prm->stddev;
  if (stats) {
    stats[0] = prm->mean; stats[1] = prm->stddev; stats[2] = prm->thr; stats[3] = nf;
    for (int l = 0; l < kLevels; ++l) stats[4 + l] = (double)prm->counts[l];
  }
}

// ------------------------------------------------------------------ ordered compaction of the inliers
__device__ __forceinline__ bool sor_keep(double a, double thr) { return a > 0.0 && a < thr; }

__global__ void __launch_bounds__(256) sor_count_kernel(const double* __restrict__ avg, long long N, const SorParams* prm,
                                                         uint32_t* __restrict__ blk_cnt) {
  const double thr = prm->thr;
  const long long base = (long long)blockIdx.x * kCompactBlock;
  uint32_t c = 0;
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < N && sor_keep(avg[i], thr)) ++c;
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

// exclusive scan of the block counts by one CTA (chunks of 1024 with a running carry)
__global__ void __launch_bounds__(1024) sor_scan_kernel(const uint32_t* __restrict__ blk_cnt, long long blocks,
                                                         long long* __restrict__ blk_base, long long* __restrict__ total) {
  __shared__ long long sh[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long c0 = 0; c0 < blocks; c0 += 1024) {
    const long long i = c0 + threadIdx.x;
    const long long v = i < blocks ? (long long)blk_cnt[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const long long add = (int)threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < blocks) blk_base[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) sor_emit_kernel(const float4* __restrict__ cloud, const double* __restrict__ avg,
                                                        long long N, const SorParams* prm,
                                                        const long long* __restrict__ blk_base,
                                                        float4* __restrict__ out_cloud, long long* __restrict__ out_index) {
  __shared__ uint32_t ws[8];
  const double thr = prm->thr;
  const long long base = (long long)blockIdx.x * kCompactBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long run = blk_base[blockIdx.x];
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    const bool keep = i < N && sor_keep(avg[i], thr);
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) ws[warp] = __popc(m);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w = 0; w < 8; ++w) { const uint32_t c = ws[w]; if (w < warp) before += c; total += c; }
    if (keep) {
      const long long o = run + before + __popc(m & ((1u << lane) - 1u));
      if (out_cloud) out_cloud[o] = __ldg(cloud + i);
      if (out_index) out_index[o] = i;
    }
    run += total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ launcher
int launch_outlier_removal(const float* cloud_f, long long N, int nb_neighbors, double std_ratio, float* out_cloud,
                           long long* out_index, double* mean_dist, double* stats, long long* out_count, void* ws,
                           cudaStream_t st) {
  SorWs w;
  sor_carve(ws, N, &w);
  const float4* cloud = reinterpret_cast<const float4*>(cloud_f);
  int launches = 0;
  if (N == 0) {
    SL_CUDA(cudaMemsetAsync(out_count, 0, sizeof(long long), st));
    if (stats) SL_CUDA(cudaMemsetAsync(stats, 0xff, 16 * sizeof(double), st));      // NaN
    return kOk;
  }
  const uint32_t mask = w.M - 1;
  sor_init_kernel<<<1, 32, 0, st>>>(w.prm); ++launches;
  sor_bbox_kernel<<<kSorBlocks, 256, 0, st>>>(cloud, N, w.prm, w.avg); ++launches;
  sor_plan_pilot_kernel<<<1, 32, 0, st>>>(w.prm); ++launches;
  sor_clear_kernel<<<kSorBlocks, 256, 0, st>>>(w.prm, -1, w.keys, w.cnt, w.M); ++launches;
  sor_insert_kernel<<<kSorBlocks, 256, 0, st>>>(cloud, N, w.prm, -1, w.keys, w.cnt, mask, w.slot); ++launches;
  sor_plan_final_kernel<<<1, 32, 0, st>>>(w.prm); ++launches;
  for (int l = 0; l < kLevels; ++l) {
    sor_clear_kernel<<<kSorBlocks, 256, 0, st>>>(w.prm, l, w.keys, w.cnt, w.M);
    sor_insert_kernel<<<kSorBlocks, 256, 0, st>>>(cloud, N, w.prm, l, w.keys, w.cnt, mask, w.slot);
    sor_alloc_kernel<<<kSorBlocks, 256, 0, st>>>(w.prm, l, w.cnt, w.start, w.M);
    sor_scatter_kernel<<<kSorBlocks, 256, 0, st>>>(cloud, N, w.prm, l, w.slot, w.start, w.sorted);
    const uint32_t* pin = w.pend[l & 1];
    uint32_t* pout = w.pend[(l + 1) & 1];
    // the two query kernels decide on the device which of them handles this level (kThreadModeMin)
    const size_t q0_smem = (size_t)nb_neighbors * kQ0Threads * sizeof(double);      // <= 128 KB at 128 neighbours
    if (l == 0 && q0_smem > 48 * 1024)
      SL_CUDA(cudaFuncSetAttribute(sor_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q0_smem));
    sor_query_kernel<<<kSorBlocks * 4, kQ0Threads, q0_smem, st>>>(cloud, w.prm, l, w.sorted, w.keys, w.cnt, w.start, mask,
                                                                 nb_neighbors, pin, pout, w.avg);
    if (l > 0) {
      if (nb_neighbors <= 32)
        sor_query_part_kernel<32><<<kQbItems, kQbThreads, 0, st>>>(cloud, w.prm, l, w.sorted, w.keys, w.cnt, w.start, mask,
                                                                   nb_neighbors, pin, w.part_best, w.q_total);
      else
        sor_query_part_kernel<128><<<kQbItems, kQbThreads, 0, st>>>(cloud, w.prm, l, w.sorted, w.keys, w.cnt, w.start, mask,
                                                                    nb_neighbors, pin, w.part_best, w.q_total);
      sor_query_merge_kernel<<<kSorBlocks, 128, 0, st>>>(cloud, w.prm, l, nb_neighbors, pin, w.part_best, w.q_total, pout,
                                                         w.avg);
    }
    launches += l > 0 ? 7 : 5;
  }
  sor_partial_kernel<<<kRedBlocks, 256, 0, st>>>(w.avg, N, w.prm, 0, w.part);
  sor_final_kernel<<<1, kRedBlocks, 0, st>>>(w.prm, 0, w.part, std_ratio, stats);
  sor_partial_kernel<<<kRedBlocks, 256, 0, st>>>(w.avg, N, w.prm, 1, w.part);
  sor_final_kernel<<<1, kRedBlocks, 0, st>>>(w.prm, 1, w.part, std_ratio, stats);
  const long long blocks = (N + kCompactBlock - 1) / kCompactBlock;
  sor_count_kernel<<<(int)blocks, 256, 0, st>>>(w.avg, N, w.prm, w.blk_cnt);
  sor_scan_kernel<<<1, 1024, 0, st>>>(w.blk_cnt, blocks, w.blk_base, out_count);
  sor_emit_kernel<<<(int)blocks, 256, 0, st>>>(cloud, w.avg, N, w.prm, w.blk_base, reinterpret_cast<float4*>(out_cloud),
                                               out_index);
  launches += 7;
  if (mean_dist) SL_CUDA(cudaMemcpyAsync(mean_dist, w.avg, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < launches; ++i) count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace sl

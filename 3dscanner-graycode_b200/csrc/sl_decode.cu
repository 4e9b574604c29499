// Stage 1: fused Gray-code decode kernels (sm_100a).
//
// K1  decode_tma_kernel     persistent, one CTA per SM (1 producer + 16 consumer warps that claim
//                           128-px sub-tiles dynamically).  The producer thread streams [F x 512 px]
//                           tiles into a ring of shared-memory stages, ONE TMA tensor load per
//                           tile (cp.async.bulk.tensor.3d + mbarrier complete_tx; the stack is
//                           described as a (pixels/4 x F x B) uint32 tensor); consumer warps read them as
//                           conflict-free 32-bit words (4 pixels / thread), keep every per-pixel
//                           decision in registers as byte-SWAR flags and write two int16 indices.
//                           Bound: HBM.  Algorithmic bytes / pixel = F + 4.
// K1g decode_generic_kernel any alignment / any eps: direct loads, fp64 compares exactly as
//                           scanner/grayCode/decode_codes.py:172-182 writes them.
// K1m codes_to_pixels_kernel run merge (src/3-capture_decode.py:95-96) + Gray->decimal (:99-100).
#include <atomic>
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through the runtime)
#include <cuda_runtime.h>
#if !defined(SL_NO_BINV_TABLE) && !defined(SL_BINV_TABLE)
#define SL_BINV_TABLE 1      // measured: 20.1 -> 18.6 us per pipelined 1080p scan, 85 % -> 90 % of HBM peak batched
#endif
#if defined(SL_BINV_TABLE)
// b_inv = white / (white + black) for every pair of 8-bit samples, filled once per device by
// binv_table_kernel with the same correctly rounded fp64 division (NaN for 0/0); 512 KB, L2-resident.
namespace sl { __device__ double g_binv_table[65536]; }
#define SL_BINV_LOOKUP(w, b) __ldg(&::sl::g_binv_table[((w) << 8) | (b)])
#endif
#include "sl_common.cuh"
#include "sl_decode_quad.cuh"
#include "sl_kernels.h"

namespace sl {

#if defined(SL_BINV_TABLE)
__global__ void binv_table_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // i = white << 8 | black
  if (i < 65536) g_binv_table[i] = __ddiv_rn((double)(i >> 8), (double)((i >> 8) + (i & 255)));
}
#endif

// Called once per device from the first C-ABI entry on that device (sl_capi.cu: device_gate, under its mutex).
// The table is filled on a private blocking-free stream and the call returns only when the fill has completed, so
// no launch on any stream or thread can ever read a partly filled table.  If the calling thread is capturing a
// CUDA graph, the thread's capture mode is relaxed for the duration: the fill runs immediately, outside the capture.
static int fill_binv_table() {
#if defined(SL_BINV_TABLE)
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  SL_CUDA(cudaThreadExchangeStreamCaptureMode(&mode));
  cudaStream_t st = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    binv_table_kernel<<<256, 256, 0, st>>>();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
  cudaThreadExchangeStreamCaptureMode(&mode);
  if (e != cudaSuccess) return cuda_fail(e, "b_inv table initialisation");
  count_launch();
#endif
  return kOk;
}

// ------------------------------------------------------------------ K1 configuration
#ifndef SL_TILE_PX
#define SL_TILE_PX 512
#endif
#ifndef SL_CONSUMER_WARPS
#define SL_CONSUMER_WARPS 16
#endif
#ifndef SL_MAX_STAGES
#define SL_MAX_STAGES 8
#endif
constexpr int kTilePx = SL_TILE_PX;           // pixels per stage row (bytes per frame per stage)
constexpr int kSubTiles = kTilePx / 128;      // a warp handles 128 px (32 lanes * 4 px) at a time
constexpr int kConsumerWarps = SL_CONSUMER_WARPS;
constexpr int kDecodeThreads = (kConsumerWarps + 1) * 32;   // + 1 producer warp
constexpr int kMaxStages = SL_MAX_STAGES;
constexpr int kSmemBudget = 227 * 1024;

// four int16 indices of a quad: one 8-byte store, or four 2-byte stores when scan b's plane is only 2-byte aligned
// (a batch whose scans hold a number of pixels that is not a multiple of 4)
__device__ __forceinline__ void store_idx4(int16_t* p, uint32_t lo, uint32_t hi, int vec) {
  if (vec) {
    st_global_v2(p, lo, hi);
  } else {
    p[0] = (int16_t)(lo & 0xffffu); p[1] = (int16_t)(lo >> 16);
    p[2] = (int16_t)(hi & 0xffffu); p[3] = (int16_t)(hi >> 16);
  }
}

struct DecodeArgs {
  const uint8_t* frames;
  size_t frame_pitch, scan_pitch;
  int16_t* h_idx;
  int16_t* v_idx;
  int8_t* h_codes;
  int8_t* v_codes;
  uint32_t* bitmap;      // nullable: flat validity bits, scan b at bitmap + b*bitmap_words
  long long bitmap_words;
  unsigned long long* flags;   // nullable: the offsets kernel's published totals, zeroed here (see strip_prefix_kernel)
  int n_flags;
  long long npix;        // pixels per scan
  long long npix4;       // npix rounded down to a multiple of 4 (part handled by this kernel)
  long long total_tiles; // B * tiles_per_scan
  int tiles_per_scan;
  int n, F, e, stages;
  unsigned long long* stats;   // nullable: [0] += quads decoded, [1] += quads that took the full rule evaluation
  int idx_vec;           // every scan's index planes are 8-byte aligned (B == 1 or npix % 4 == 0): vector stores
};

// Register cap: the decode CTA (544 threads, 188 KB of shared memory) leaves room on the SM for one emit
// CTA of another stream only if it stays at 72 registers (measured: 38.4 -> 36.0 us per pipelined scan,
// decode itself unchanged).
#ifndef SL_DECODE_MAXREG
#define SL_DECODE_MAXREG 72
#endif
#define SL_DECODE_REG_ATTR __maxnreg__(SL_DECODE_MAXREG)
template <int NB, bool CODES>
__global__ void SL_DECODE_REG_ATTR decode_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                        const DecodeArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int n = NB ? NB : a.n;
  const int F = 4 * n + 2;
  const int stages = a.stages;
  const uint32_t stage_bytes = (uint32_t)F * kTilePx;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  // bars[0..stages) = full, bars[stages..2*stages) = empty; then the sub-tile claim counter
  unsigned int* next_sub = reinterpret_cast<unsigned int*>(bars + 2 * stages);
  unsigned int* warp_stats = next_sub + 2;          // [kConsumerWarps][2], only touched when a.stats is given
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[stages + s]), kSubTiles);
    }
    *next_sub = 0;
    mbar_init_fence();
  }
  __syncthreads();
  pdl_wait();          // everything above overlapped the previous kernel of the stream; global memory from here on
  if (blockIdx.x == 0 && a.flags)
    for (int i = threadIdx.x; i < a.n_flags; i += blockDim.x) a.flags[i] = 0;

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------ producer (one lane)
    if (lane == 0) {
      tma_prefetch_desc(&tmap);
      const uint64_t pol = policy_evict_first();
      int s = 0;
      uint32_t ph = 0;
      // 32-bit tile bookkeeping (total_tiles < 2^31 is checked by the launcher); (b, tile) advance
      // incrementally, a division only when a scan boundary is crossed
      const uint32_t total = (uint32_t)a.total_tiles, tps = (uint32_t)a.tiles_per_scan, grid = gridDim.x;
      uint32_t tile = blockIdx.x, b = 0;
      for (uint32_t t = blockIdx.x; t < total; t += grid, tile += grid) {
        if (tile >= tps) { const uint32_t q = tile / tps; b += q; tile -= q * tps; }
        mbar_wait(smem_u32(&bars[stages + s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bars[s]);
        mbar_arrive_expect_tx(full, stage_bytes);          // the box is always written in full
        tma_load_3d(smem_u32(smem + (size_t)s * stage_bytes), &tmap, (int)(tile * (kTilePx / 4)), 0, (int)b, full, pol);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  // All consumer warps draw 128-pixel sub-tiles from one counter, in ring order.  A warp that hits the
  // slow path (full rule evaluation) therefore delays nobody: the others keep claiming work up to
  // `stages - 1` tiles ahead of it.
  const int off_lane = lane * 4;
  // 32-bit bookkeeping, advanced incrementally from claim to claim (claims of a warp are monotonic):
  // ring slot s / phase ph of iteration `it`, and (scan b, tile in scan) of tile t = blockIdx.x + it * grid
  const uint32_t total = (uint32_t)a.total_tiles, tps = (uint32_t)a.tiles_per_scan, grid = gridDim.x;
  uint32_t it_prev = 0, s = 0, ph = 0, tile = blockIdx.x, b = 0;
  if (a.stats && lane == 0) { warp_stats[2 * warp] = 0; warp_stats[2 * warp + 1] = 0; }
  for (;;) {
    unsigned int sub = 0;
    if (lane == 0) sub = atomicAdd(next_sub, 1u);
    sub = __shfl_sync(0xffffffffu, sub, 0);
    const uint32_t it = sub / kSubTiles;                 // kSubTiles is a power of two
    const int part = (int)(sub % kSubTiles);
    const uint32_t t = blockIdx.x + it * grid;
    if (t >= total) {
      if (a.stats && lane == 0) {
        atomicAdd(a.stats, (unsigned long long)warp_stats[2 * warp]);
        atomicAdd(a.stats + 1, (unsigned long long)warp_stats[2 * warp + 1]);
      }
      break;
    }
    const uint32_t delta = it - it_prev;
    it_prev = it;
    s += delta;
    while (s >= (uint32_t)stages) { s -= (uint32_t)stages; ph ^= 1u; }
    tile += delta * grid;
    if (tile >= tps) { const uint32_t q = tile / tps; b += q; tile -= q * tps; }
    mbar_wait(smem_u32(&bars[s]), ph);

    const long long px0 = (long long)tile * kTilePx;
    const long long rem = a.npix4 - px0;
    const int cnt = (int)(rem < kTilePx ? rem : kTilePx);
    const int off = part * 128 + off_lane;                                 // pixel offset in tile
    uint32_t nib = 0, slow = 0;
    if (off < cnt) {
#if !defined(SL_DECODE_GENERIC_LD)
      // explicit ld.shared with a 32-bit address.  Through the lambda the compiler only sees a generic pointer and
      // emits LD.E (generic loads with 64-bit addresses); measured 24.6 -> 22.9 us per isolated 1080p launch and
      // 18.0 -> 17.3 us pipelined (88 % -> 92 % of the measured HBM peak).
      const uint32_t base_s = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)off;
      auto ld = [base_s](int f) { return lds_u32(base_s + (uint32_t)f * kTilePx); };
#else
      const uint32_t* base = reinterpret_cast<const uint32_t*>(smem + (size_t)s * stage_bytes) + (off >> 2);
      auto ld = [&](int f) { return base[f * (kTilePx / 4)]; };
#endif
      const long long px = px0 + off;
      int8_t* hc = CODES ? a.h_codes + (size_t)b * n * a.npix + px : nullptr;
      int8_t* vc = CODES ? a.v_codes + (size_t)b * n * a.npix + px : nullptr;
      auto st = [&](int axis, int bit, uint32_t tri) {
        int8_t* dst = (axis ? vc : hc) + (size_t)bit * a.npix;
        *reinterpret_cast<uint32_t*>(dst) = tri;
      };
      const QuadOut o = decode_quad<NB, CODES>(ld, st, n, a.e);
      store_idx4(a.h_idx + (size_t)b * a.npix + px, o.h01, o.h23, a.idx_vec);
      store_idx4(a.v_idx + (size_t)b * a.npix + px, o.v01, o.v23, a.idx_vec);
      nib = ((o.valid & kMsb4) * 0x00204081u) >> 28;      // the 4 msb flags gathered into bits 0..3
      slow = o.slow;
    }
    if (a.stats) {
      const unsigned int m_all = __ballot_sync(0xffffffffu, off < cnt), m_slow = __ballot_sync(0xffffffffu, slow != 0);
      if (lane == 0) { warp_stats[2 * warp] += __popc(m_all); warp_stats[2 * warp + 1] += __popc(m_slow); }
    }
    if (a.bitmap) {
      // 8 neighbouring lanes hold the 32 validity bits of 32 consecutive pixels
      const uint32_t word = __reduce_or_sync(0xffu << (lane & 24), nib << (4 * (lane & 7)));
      if ((lane & 7) == 0 && px0 + off < a.npix)
        a.bitmap[(size_t)b * a.bitmap_words + ((px0 + off) >> 5)] = word;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[stages + s]));
  }
}

// ------------------------------------------------------------------ K1c decode straight from BGR camera frames
// The capture loop converts every frame with cv2.cvtColor(BGR2GRAY) before it enters the stack
// (src/3-capture_decode.py:66).  This variant streams the 3-byte pixels themselves (3F + 4 algorithmic
// bytes per pixel instead of 3F + F for the conversion pass plus F + 4 for the decode): same persistent
// producer / consumer structure, 128-pixel tiles (one warp each, F x 384 bytes per stage).  A consumer
// lane converts its own 4 pixels of every frame (gray4_from_bgr12: two IDP4A per pixel) into a private
// column of a per-warp gray tile in shared memory, hands the BGR stage back to the producer, and then runs
// the unchanged per-quad decode on the gray tile.
constexpr int kBgrTilePx = 128;
constexpr int kBgrRowBytes = kBgrTilePx * 3;
constexpr int kBgrFullBars = 32;     // per-tile "landed" barriers (> stages + consumer warps, see the kernel)
static_assert(kConsumerWarps + 12 <= kBgrFullBars, "claims may run stages + warps - 1 tiles ahead");

template <int NB>
__global__ void SL_DECODE_REG_ATTR decode_bgr_tma_kernel(const __grid_constant__ CUtensorMap tmap, const DecodeArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int n = NB ? NB : a.n;
  const int F = 4 * n + 2;
  const int stages = a.stages;
  const uint32_t stage_bytes = (uint32_t)F * kBgrRowBytes;
  uint8_t* gray_base = smem + (size_t)stages * stage_bytes;                 // [kConsumerWarps][F][32 words]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gray_base + (size_t)kConsumerWarps * F * kBgrTilePx);
  // bars[0..kBgrFullBars) = "tile landed", one per tile modulo kBgrFullBars; then `stages` "stage free" barriers.
  // There are more consumer warps (one tile each) than stages, so a warp can hold a claim that is a whole revolution
  // of the stage ring ahead; a parity wait on a per-STAGE barrier could not tell that apart from "already complete".
  // Claims never run more than stages + kConsumerWarps - 1 (<= 27) tiles ahead of the oldest unreleased tile (the
  // producer stalls on that tile's stage), so with 32 per-TILE barriers the previous use of a barrier has always
  // landed when a warp waits for the next one, and the parity (tile / 32) & 1 is unambiguous.
  unsigned int* next_tile = reinterpret_cast<unsigned int*>(bars + kBgrFullBars + stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (blockIdx.x == 0 && a.flags)
    for (int i = threadIdx.x; i < a.n_flags; i += blockDim.x) a.flags[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBgrFullBars + stages; ++i) mbar_init(smem_u32(&bars[i]), 1);
    *next_tile = 0;
    mbar_init_fence();
  }
  __syncthreads();

  const uint32_t total = (uint32_t)a.total_tiles, tps = (uint32_t)a.tiles_per_scan, grid = gridDim.x;
  if (warp == kConsumerWarps) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap);
      const uint64_t pol = policy_evict_first();
      int s = 0;
      uint32_t ph = 0, tile = blockIdx.x, b = 0, it = 0;
      for (uint32_t t = blockIdx.x; t < total; t += grid, tile += grid, ++it) {
        if (tile >= tps) { const uint32_t q = tile / tps; b += q; tile -= q * tps; }
        mbar_wait(smem_u32(&bars[kBgrFullBars + s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bars[it % kBgrFullBars]);
        mbar_arrive_expect_tx(full, stage_bytes);
        tma_load_3d(smem_u32(smem + (size_t)s * stage_bytes), &tmap, (int)(tile * (kBgrRowBytes / 4)), 0, (int)b, full, pol);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }

  const uint32_t gray_s = smem_u32(gray_base + (size_t)warp * F * kBgrTilePx) + lane * 4;   // this lane's gray column
  uint32_t it_prev = 0, s = 0, tile = blockIdx.x, b = 0;
  for (;;) {
    unsigned int it = 0;
    if (lane == 0) it = atomicAdd(next_tile, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    const uint32_t t = blockIdx.x + it * grid;
    if (t >= total) break;
    const uint32_t delta = it - it_prev;
    it_prev = it;
    s += delta;
    while (s >= (uint32_t)stages) s -= (uint32_t)stages;
    tile += delta * grid;
    if (tile >= tps) { const uint32_t q = tile / tps; b += q; tile -= q * tps; }
    mbar_wait(smem_u32(&bars[it % kBgrFullBars]), (it / kBgrFullBars) & 1u);

    // BGR stage -> this lane's gray column (out-of-range pixels of the last tile are zero-filled by the TMA)
    // plain loads from the stage (free to be scheduled several frames ahead; the mbarrier wait above is a compiler
    // barrier), explicit st.shared into the gray column
    const uint32_t* src = reinterpret_cast<const uint32_t*>(smem + (size_t)s * stage_bytes) + lane * 3;
    constexpr int kGroup = 4;                              // frames whose loads are issued before any is converted
    const int nf = NB ? 4 * NB + 2 : F;                    // compile-time trip count: immediate smem offsets
#pragma unroll
    for (int f0 = 0; f0 < nf; f0 += kGroup) {
      uint32_t w[kGroup][3];
#pragma unroll
      for (int u = 0; u < kGroup; ++u)
        if (f0 + u < nf) {
          const uint32_t* r = src + (f0 + u) * (kBgrRowBytes / 4);
          w[u][0] = r[0]; w[u][1] = r[1]; w[u][2] = r[2];
        }
#pragma unroll
      for (int u = 0; u < kGroup; ++u)
        if (f0 + u < nf) sts_u32(gray_s + (f0 + u) * kBgrTilePx, gray4_from_bgr12(w[u][0], w[u][1], w[u][2]));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[kBgrFullBars + s]));  // the producer may refill the stage now

    const long long px0 = (long long)tile * kBgrTilePx;
    const long long rem = a.npix4 - px0;
    const int cnt = (int)(rem < kBgrTilePx ? rem : kBgrTilePx);
    const int off = lane * 4;
    uint32_t nib = 0;
    if (off < cnt) {
      auto ld = [gray_s](int f) { return lds_u32(gray_s + f * kBgrTilePx); };
      auto st = [](int, int, uint32_t) {};
      const long long px = px0 + off;
      const QuadOut o = decode_quad<NB, false>(ld, st, n, a.e);
      store_idx4(a.h_idx + (size_t)b * a.npix + px, o.h01, o.h23, a.idx_vec);
      store_idx4(a.v_idx + (size_t)b * a.npix + px, o.v01, o.v23, a.idx_vec);
      nib = ((o.valid & kMsb4) * 0x00204081u) >> 28;
    }
    if (a.bitmap) {
      const uint32_t word = __reduce_or_sync(0xffu << (lane & 24), nib << (4 * (lane & 7)));
      if ((lane & 7) == 0 && px0 + off < a.npix)
        a.bitmap[(size_t)b * a.bitmap_words + ((px0 + off) >> 5)] = word;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ K1g generic
struct GenericArgs {
  const uint8_t* frames;
  size_t frame_pitch, scan_pitch;
  int16_t* h_idx;
  int16_t* v_idx;
  int8_t* h_codes;
  int8_t* v_codes;
  long long npix, px_begin, px_end;   // per-scan pixel range [px_begin, px_end)
  int n, B;
  double eps;
  uint32_t* bitmap;       // nullable: validity bits are OR-ed in (the launcher zeroes the words first)
  long long bitmap_words;
  unsigned long long* flags;   // nullable, zeroed (see DecodeArgs)
  int n_flags;
  int bgr;                // frames hold 3-byte BGR pixels, converted like cv2.cvtColor(BGR2GRAY) on load
  const double* ld_in;    // nullable: caller-supplied L_d / L_g (get_is_lit, decode_codes.py:125)
  const double* lg_in;
  double* ld_out;         // nullable: get_direct_indirect outputs (decode_codes.py:90-122)
  double* lg_out;
};

__device__ __forceinline__ int classify_fp64(double p, double q, double ld, double lge, bool a, double eps) {
  // decode_codes.py:172-182, later rule wins.  Every sum is a single rounded fp64 add.
  const double pe = dadd(p, eps), qe = dadd(q, eps);
  int r = -1;
  if (a && p > qe) r = 1;
  if (a && pe < q) r = 0;
  if (pe < ld && q > lge) r = 0;
  if (p > lge && qe < ld) r = 1;
  return r;
}

__global__ void __launch_bounds__(256) decode_generic_kernel(const GenericArgs a) {
  if (a.flags && blockIdx.x == 0)
    for (int i = threadIdx.x; i < a.n_flags; i += blockDim.x) a.flags[i] = 0;
  const long long span = a.px_end - a.px_begin;
  const long long total = span * a.B;
  const int n = a.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / span;
    const long long px = a.px_begin + (i - b * span);
    const uint8_t* f0 = a.frames + (size_t)b * a.scan_pitch + (size_t)px * (a.bgr ? 3 : 1);
    auto ld = [&](int f) {
      const uint8_t* p = f0 + (size_t)f * a.frame_pitch;
      if (!a.bgr) return (double)__ldg(p);
      return (double)((3735u * __ldg(p) + 19235u * __ldg(p + 1) + 9798u * __ldg(p + 2) + 16384u) >> 15);
    };
    const double black = ld(0), white = ld(1);
    double lmax = ld(2 + 2 * (n - 1)), lmin = ld(3);
    lmax = fmax(lmax, ld(2 + 2 * (n - 2)));  lmax = fmax(lmax, ld(2 + 2 * (n - 3)));
    lmax = fmax(lmax, ld(2 + 2 * n + 2 * (n - 1)));
    lmax = fmax(lmax, ld(2 + 2 * n + 2 * (n - 2)));
    lmax = fmax(lmax, ld(2 + 2 * n + 2 * (n - 3)));
    lmin = fmin(lmin, ld(5));  lmin = fmin(lmin, ld(7));
    lmin = fmin(lmin, ld(3 + 2 * n));  lmin = fmin(lmin, ld(5 + 2 * n));  lmin = fmin(lmin, ld(7 + 2 * n));
    const double b_inv = ddiv(white, dadd(white, black));               // decode_codes.py:113
    double l_d = dmul(dsub(lmax, lmin), b_inv);                          // :119
    double l_g = dmul(dmul(2.0, dsub(lmax, l_d)), b_inv);                // :120
    if (a.ld_out) {
      a.ld_out[(size_t)b * a.npix + px] = l_d;
      a.lg_out[(size_t)b * a.npix + px] = l_g;
      if (!a.h_idx) continue;
    }
    if (a.ld_in) {
      l_d = a.ld_in[(size_t)b * a.npix + px];
      l_g = a.lg_in[(size_t)b * a.npix + px];
    }
    const double lge = dadd(l_g, a.eps);
    const bool big = l_d > lge;

    uint32_t gh = 0, gv = 0;
    bool okh = true, okv = true;
    for (int j = 0; j < n; ++j) {        // column bits, MSB first
      const int r = classify_fp64(ld(2 + 2 * j), ld(2 + 2 * n + 2 * j), l_d, lge, big, a.eps);
      okh &= (r >= 0);
      gh = (gh << 1) | (uint32_t)(r > 0);
      if (a.h_codes) a.h_codes[((size_t)b * n + j) * a.npix + px] = (int8_t)r;
    }
    for (int k = n - 1; k >= 0; --k) {   // row bits are stored LSB first
      const int r = classify_fp64(ld(3 + 2 * k), ld(3 + 2 * n + 2 * k), l_d, lge, big, a.eps);
      okv &= (r >= 0);
      gv = (gv << 1) | (uint32_t)(r > 0);
      if (a.v_codes) a.v_codes[((size_t)b * n + k) * a.npix + px] = (int8_t)r;
    }
    if (a.h_idx) {
      a.h_idx[(size_t)b * a.npix + px] = okh ? (int16_t)gray_to_binary16(gh) : (int16_t)-1;
      a.v_idx[(size_t)b * a.npix + px] = okv ? (int16_t)gray_to_binary16(gv) : (int16_t)-1;
    }
    if (a.bitmap && okh && okv) atomicOr(a.bitmap + (size_t)b * a.bitmap_words + (px >> 5), 1u << (px & 31));
  }
}

// ------------------------------------------------------------------ K1m merge + convert
__global__ void __launch_bounds__(256) codes_to_pixels_kernel(const int8_t* __restrict__ h_codes,
                                                               const int8_t* __restrict__ v_codes,
                                                               int R, int n, long long npix,
                                                               size_t run_pitch,
                                                               int16_t* __restrict__ h_idx,
                                                               int16_t* __restrict__ v_idx) {
  for (long long px = blockIdx.x * (long long)blockDim.x + threadIdx.x; px < npix;
       px += (long long)gridDim.x * blockDim.x) {
    uint32_t gh = 0, gv = 0;
    bool okh = true, okv = true;
    for (int j = 0; j < n; ++j) {
      int mh = -1, mv = -1;
      for (int r = 0; r < R; ++r) {     // np.max over runs: 1 > 0 > -1
        mh = max(mh, (int)h_codes[(size_t)r * run_pitch + (size_t)j * npix + px]);
        mv = max(mv, (int)v_codes[(size_t)r * run_pitch + (size_t)(n - 1 - j) * npix + px]);
      }
      okh &= (mh >= 0);  okv &= (mv >= 0);
      gh = (gh << 1) | (uint32_t)(mh > 0);
      gv = (gv << 1) | (uint32_t)(mv > 0);
    }
    h_idx[px] = okh ? (int16_t)gray_to_binary16(gh) : (int16_t)-1;
    v_idx[px] = okv ? (int16_t)gray_to_binary16(gv) : (int16_t)-1;
  }
}

// ------------------------------------------------------------------ launchers
// `sm_reserve` (a launch argument of the C ABI): SMs the persistent decode grid leaves to other work.  A caller that
// pipelines independent scans over several streams gets more out of the GPU when the decode grid does not cover
// every SM: the offsets / emit kernels of the other scans then always find SMs of their own instead of the CTA slot
// beside a resident decode CTA.  Measured on the bench step (4 streams): 34.2 us per scan with 148 CTAs, 33.4 with
// 145, 32.4 with 140, 32.6 with 138 / 135; a launch on its own gets slower instead (21.9 -> 23.0 us at 140).
static int decode_grid_sms(int num_sms, int sm_reserve) {
  const int r = sm_reserve < 0 ? 0 : sm_reserve;
  return num_sms > 2 * r ? num_sms - r : num_sms;
}

// cuTensorMapEncodeTiled, fetched through the runtime so libcuda is not a link dependency
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static std::atomic<EncodeTiledFn> g_encode_tiled{nullptr};

static int make_frames_tensor_map(CUtensorMap* out, const uint8_t* frames, long long npix4, int F, int B,
                                  size_t frame_pitch, size_t scan_pitch, int bytes_per_px = 1,
                                  int box_words = kTilePx / 4) {
  EncodeTiledFn encode = g_encode_tiled.load(std::memory_order_acquire);
  if (!encode) { set_error("cuTensorMapEncodeTiled not available (device not initialised)"); return kErrCuda; }
  // tensor of uint32 (4 pixels): dim0 = npix4/4 along a frame, dim1 = frame, dim2 = scan
  const cuuint64_t dims[3] = {(cuuint64_t)(npix4 / 4) * bytes_per_px, (cuuint64_t)F, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)frame_pitch, (cuuint64_t)(B > 1 ? scan_pitch : frame_pitch * F)};
  const cuuint32_t box[3] = {(cuuint32_t)box_words, (cuuint32_t)F, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(frames), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (npix4=%lld F=%d B=%d pitch=%zu/%zu)", (int)r, npix4, F, B,
              frame_pitch, scan_pitch);
    return kErrCuda;
  }
  return kOk;
}

template <int NB, bool CODES>
static int launch_tma(const CUtensorMap& tmap, const DecodeArgs& a, int grid, size_t smem, cudaStream_t st) {
  auto k = decode_tma_kernel<NB, CODES>;     // > 48 KB of dynamic shared memory: opted in by decode_device_init
  SL_CUDA(launch_pdl(k, dim3(grid), dim3(kDecodeThreads), smem, st, tmap, a));
  count_launch();
  return kOk;
}

template <bool CODES>
static int dispatch_tma(const CUtensorMap& tmap, const DecodeArgs& a, int grid, size_t smem, cudaStream_t st) {
  switch (a.n) {
    case 8:  return launch_tma<8, CODES>(tmap, a, grid, smem, st);
    case 9:  return launch_tma<9, CODES>(tmap, a, grid, smem, st);
    case 10: return launch_tma<10, CODES>(tmap, a, grid, smem, st);
    case 11: return launch_tma<11, CODES>(tmap, a, grid, smem, st);
    case 12: return launch_tma<12, CODES>(tmap, a, grid, smem, st);
    default: return launch_tma<0, CODES>(tmap, a, grid, smem, st);
  }
}

template <int NB>
static int launch_bgr_tma(const CUtensorMap& tmap, const DecodeArgs& a, int grid, size_t smem, cudaStream_t st) {
  auto k = decode_bgr_tma_kernel<NB>;
  k<<<grid, kDecodeThreads, smem, st>>>(tmap, a);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

static int dispatch_bgr_tma(const CUtensorMap& tmap, const DecodeArgs& a, int grid, size_t smem, cudaStream_t st) {
  switch (a.n) {
    case 10: return launch_bgr_tma<10>(tmap, a, grid, smem, st);
    case 11: return launch_bgr_tma<11>(tmap, a, grid, smem, st);
    case 12: return launch_bgr_tma<12>(tmap, a, grid, smem, st);
    default: return launch_bgr_tma<0>(tmap, a, grid, smem, st);
  }
}

// Once per device, called by sl_capi.cu's device gate under its mutex: everything the launch path would otherwise
// initialise lazily (and racily): the b_inv table, the tensor-map encoder, the shared-memory opt-ins.
template <int NB>
static int opt_in_smem() {
  SL_CUDA(cudaFuncSetAttribute(decode_tma_kernel<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  SL_CUDA(cudaFuncSetAttribute(decode_tma_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  return kOk;
}
template <int NB>
static int opt_in_smem_bgr() {
  SL_CUDA(cudaFuncSetAttribute(decode_bgr_tma_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  return kOk;
}

int decode_device_init() {
  int rc = fill_binv_table();
  if (rc) return rc;
  if (!g_encode_tiled.load(std::memory_order_acquire)) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    SL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled not available from this driver");
      return kErrCuda;
    }
    g_encode_tiled.store(reinterpret_cast<EncodeTiledFn>(fn), std::memory_order_release);
  }
  if ((rc = opt_in_smem<0>()) || (rc = opt_in_smem<8>()) || (rc = opt_in_smem<9>()) || (rc = opt_in_smem<10>()) ||
      (rc = opt_in_smem<11>()) || (rc = opt_in_smem<12>())) return rc;
  if ((rc = opt_in_smem_bgr<0>()) || (rc = opt_in_smem_bgr<10>()) || (rc = opt_in_smem_bgr<11>()) ||
      (rc = opt_in_smem_bgr<12>())) return rc;
  return kOk;
}

int launch_decode(const uint8_t* frames, int B, int F, long long npix, size_t frame_pitch,
                  size_t scan_pitch, double eps, int16_t* h_idx, int16_t* v_idx, int8_t* h_codes,
                  int8_t* v_codes, uint32_t* valid_bitmap, long long bitmap_words, unsigned long long* flags, int n_flags,
                  int num_sms, cudaStream_t st, bool bgr, int sm_reserve, unsigned long long* stats) {
  const int n = (F - 2) / 4;
  const bool int_eps = eps >= 0.0 && eps <= 127.0 && eps == (double)(int)eps;
  const bool aligned = ((uintptr_t)frames % 16 == 0) && (frame_pitch % 16 == 0) &&
                       (scan_pitch % 16 == 0) && ((uintptr_t)h_idx % 8 == 0) &&
                       ((uintptr_t)v_idx % 8 == 0) &&
                       (!h_codes || (((uintptr_t)h_codes % 4 == 0) && ((uintptr_t)v_codes % 4 == 0) &&
                                     npix % 4 == 0));
  long long tma_px = (int_eps && aligned) ? (npix & ~3LL) : 0;
  if (frame_pitch >= (1ull << 40) || scan_pitch >= (1ull << 40)) tma_px = 0;   // tensor-map stride limit
  if (valid_bitmap && tma_px < npix)   // some pixels go through the generic kernel, which ORs its bits in
    SL_CUDA(cudaMemsetAsync(valid_bitmap, 0, (size_t)B * bitmap_words * sizeof(uint32_t), st));

  if (tma_px > 0 && bgr) {
    DecodeArgs a;
    a.frames = frames; a.frame_pitch = frame_pitch; a.scan_pitch = scan_pitch;
    a.h_idx = h_idx; a.v_idx = v_idx; a.h_codes = nullptr; a.v_codes = nullptr;
    a.bitmap = valid_bitmap; a.bitmap_words = bitmap_words; a.flags = flags; a.n_flags = n_flags; a.stats = nullptr;
    a.npix = npix; a.npix4 = tma_px;
    a.tiles_per_scan = (int)((tma_px + kBgrTilePx - 1) / kBgrTilePx);
    a.total_tiles = (long long)a.tiles_per_scan * B;
    a.n = n; a.F = F; a.e = (int)eps; a.idx_vec = (npix % 4 == 0 || B == 1) ? 1 : 0;
    const size_t stage_bytes = (size_t)F * kBgrRowBytes;
    const size_t gray_bytes = (size_t)kConsumerWarps * F * kBgrTilePx;
    int stages = (int)((kSmemBudget - 512 - gray_bytes) / stage_bytes);
    if (stages > 12) stages = 12;
    if (stages < 2) { set_error("frame count too large for the BGR decode"); return kErrBadArg; }
    if (a.total_tiles >= (1LL << 31)) { set_error("batch too large for one decode launch"); return kErrBadArg; }
    a.stages = stages;
    const size_t smem = stages * stage_bytes + gray_bytes + (kBgrFullBars + stages) * sizeof(uint64_t) + 16;
    const int sms = decode_grid_sms(num_sms, sm_reserve);
    const int grid = (int)(a.total_tiles < sms ? a.total_tiles : sms);
    CUtensorMap tmap;
    int rc = make_frames_tensor_map(&tmap, frames, tma_px, F, B, frame_pitch, scan_pitch, 3, kBgrRowBytes / 4);
    if (rc != kOk) return rc;
    if ((rc = dispatch_bgr_tma(tmap, a, grid, smem, st)) != kOk) return rc;
  } else if (tma_px > 0) {
    DecodeArgs a;
    a.frames = frames; a.frame_pitch = frame_pitch; a.scan_pitch = scan_pitch;
    a.h_idx = h_idx; a.v_idx = v_idx; a.h_codes = h_codes; a.v_codes = v_codes;
    a.bitmap = valid_bitmap; a.bitmap_words = bitmap_words; a.flags = flags; a.n_flags = n_flags; a.stats = stats;
    a.npix = npix; a.npix4 = tma_px;
    a.tiles_per_scan = (int)((tma_px + kTilePx - 1) / kTilePx);
    a.total_tiles = (long long)a.tiles_per_scan * B;
    a.n = n; a.F = F; a.e = (int)eps; a.idx_vec = (npix % 4 == 0 || B == 1) ? 1 : 0;
    const size_t stage_bytes = (size_t)F * kTilePx;
    int stages = (int)((kSmemBudget - 512) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) { tma_px = 0; }
    else {
      if (a.total_tiles >= (1LL << 31)) { set_error("batch too large for one decode launch"); return kErrBadArg; }
      a.stages = stages;
      const size_t smem = stages * stage_bytes + 2 * stages * sizeof(uint64_t) + 16 + 8 * kConsumerWarps;
      const int sms = decode_grid_sms(num_sms, sm_reserve);
      const int grid = (int)(a.total_tiles < sms ? a.total_tiles : sms);
      CUtensorMap tmap;
      int rc = make_frames_tensor_map(&tmap, frames, tma_px, F, B, frame_pitch, scan_pitch);
      if (rc != kOk) return rc;
      rc = h_codes ? dispatch_tma<true>(tmap, a, grid, smem, st) : dispatch_tma<false>(tmap, a, grid, smem, st);
      if (rc != kOk) return rc;
    }
  }
  if (tma_px < npix) {
    GenericArgs g;
    g.frames = frames; g.frame_pitch = frame_pitch; g.scan_pitch = scan_pitch;
    g.h_idx = h_idx; g.v_idx = v_idx; g.h_codes = h_codes; g.v_codes = v_codes;
    g.npix = npix; g.px_begin = tma_px; g.px_end = npix; g.n = n; g.B = B; g.eps = eps;
    g.bitmap = valid_bitmap; g.bitmap_words = bitmap_words; g.flags = flags; g.n_flags = n_flags;
    g.ld_in = g.lg_in = nullptr; g.ld_out = g.lg_out = nullptr; g.bgr = bgr ? 1 : 0;
    const long long total = (npix - tma_px) * B;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
    decode_generic_kernel<<<(int)blocks, 256, 0, st>>>(g);
    count_launch();
    SL_CUDA(cudaGetLastError());
  }
  return kOk;
}

// get_direct_indirect (ld_out/lg_out) and get_is_lit with caller-supplied L_d/L_g (ld_in/lg_in):
// the reference API lets the two halves of get_codes be called separately.
int launch_decode_parts(const uint8_t* frames, int F, long long npix, size_t frame_pitch, double eps,
                        const double* ld_in, const double* lg_in, double* ld_out, double* lg_out,
                        int8_t* h_codes, int8_t* v_codes, int num_sms, cudaStream_t st) {
  GenericArgs g;
  g.frames = frames; g.frame_pitch = frame_pitch; g.scan_pitch = 0;
  g.h_idx = nullptr; g.v_idx = nullptr; g.h_codes = h_codes; g.v_codes = v_codes;
  g.npix = npix; g.px_begin = 0; g.px_end = npix; g.n = (F - 2) / 4; g.B = 1; g.eps = eps;
  g.bitmap = nullptr; g.bitmap_words = 0; g.flags = nullptr; g.n_flags = 0;
  g.ld_in = ld_in; g.lg_in = lg_in; g.ld_out = ld_out; g.lg_out = lg_out; g.bgr = 0;
  long long blocks = (npix + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  if (blocks < 1) return kOk;
  decode_generic_kernel<<<(int)blocks, 256, 0, st>>>(g);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_codes_to_pixels(const int8_t* h_codes, const int8_t* v_codes, int R, int n,
                           long long npix, size_t run_pitch, int16_t* h_idx, int16_t* v_idx,
                           int num_sms, cudaStream_t st) {
  long long blocks = (npix + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  if (blocks < 1) blocks = 1;
  codes_to_pixels_kernel<<<(int)blocks, 256, 0, st>>>(h_codes, v_codes, R, n, npix, run_pitch, h_idx, v_idx);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace sl

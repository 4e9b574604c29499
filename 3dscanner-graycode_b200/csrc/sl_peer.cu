// Multi-GPU: gather of the compacted clouds over NVLink peer memory (one process per GPU).
//
// The reference is a single process; the partitioning is BASELINE.json's (north_star): independent scans are
// block-sharded over ranks, one huge capture is cut into row bands, and the only exchange is the gather of the
// compacted clouds.  Ordered compaction makes that gather a placement problem: every rank must know where its
// records start in the gathered, packed cloud (scan-major, each scan in the reference's x-major order,
// scanner/triangulation/triangulate.py:52-53) BEFORE it emits.  So the ranks exchange their counts first
// (a few hundred bytes, or k*W column totals for row bands), rebase the compaction offsets of their own workspace,
// and the emit kernel then writes its records straight into the gathering rank's buffer through a peer mapping:
// the transfer rides on the emission, run by run, and no staging copy, padding or concatenation exists.
//
//   exchange_offsets_kernel   publish my counts into every rank's control block (P2P stores + release flag),
//                             wait for everybody's, rebase col_base of the local workspace
//   (emit_kernel, sl_cloud.cu, writes to the peer-mapped cloud(s))
//   exchange_arrive_kernel    tell the gathering rank(s) that my records have landed / wait for all ranks
//   exchange_push_kernel      the other transport for scan-sharded batches: a rank's records of one step are one
//                             contiguous block of the gathered cloud, so they can also be emitted locally (at the
//                             local emission's speed) and pushed as one stream of full 128-byte lines by a few CTAs
//                             while the SMs go on with the next batch
//
// Control blocks live in cudaMalloc memory exported with CUDA IPC (sl_peer_*); flags are 64-bit step numbers
// written with st.release.sys and polled with ld.acquire.sys.  Every wait has a time-out (a dead peer must not
// hang the GPU): on expiry the error word of the local control block is set and the kernels run to completion.
#include <cstring>
#include <new>
#include "sl_common.cuh"
#include "sl_kernels.h"

namespace sl {

constexpr int kExchangeThreads = 512;

struct ExchangeLocal {            // first 256 bytes of a control block; only the owner touches it
  unsigned long long pub_step;    // number of exchange_offsets steps issued on this rank
  unsigned long long sig_step;    // number of arrive steps
  unsigned int error;             // 0 ok, 1 = time-out in exchange_offsets, 2 = time-out in arrive
  unsigned int pad;
  long long rank_base;            // diagnostics: packed offset of this rank's first record in the last step
  long long grand_total;          // diagnostics: points of all ranks in the last step
};

struct SlotHeader {               // 32 bytes, followed by counts[max_scans] (int64) and col_tot[max_width] (u32)
  unsigned long long seq;         // step number, written last with release.sys
  long long total;                // points of the source rank in this step
  long long n_counts;             // scans of the source rank (by-scan mode)
  long long reserved;
};

struct ExchangeLayout {
  size_t slot_bytes, slots_off, done_off, total_bytes;
  int max_scans, max_width;
};

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static_assert(kMaxWorld == SL_MAX_RANKS, "SL_MAX_RANKS");

static ExchangeLayout exchange_layout(int world, int max_scans, int max_width) {
  ExchangeLayout l;
  l.max_scans = max_scans; l.max_width = max_width;
  l.slot_bytes = round_up(sizeof(SlotHeader) + (size_t)max_scans * 8 + (size_t)max_width * 4, 256);
  l.slots_off = 256;
  l.done_off = l.slots_off + 2 * (size_t)world * l.slot_bytes;       // two parities x world source ranks
  l.total_bytes = round_up(l.done_off + (size_t)world * 8, 256);
  return l;
}

struct ExchangeArgs {
  uint8_t* ctrl[kMaxWorld];       // every rank's control block as mapped here (ctrl[rank] is the local one)
  int rank, world;
  unsigned long long slot_bytes, slots_off, done_off;
  int max_scans, max_width;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spin until *p >= want; false on time-out
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long want,
                                          unsigned long long timeout_ns) {
  const unsigned long long t0 = global_timer_ns();
  unsigned int spins = 0;
  while (ld_acquire_sys_u64(p) < want) {
    if ((++spins & 255u) == 0 && global_timer_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
  return true;
}

__device__ __forceinline__ uint8_t* slot_ptr(const ExchangeArgs& x, int dst, unsigned long long seq, int src) {
  return x.ctrl[dst] + x.slots_off + ((seq & 1ull) * (unsigned long long)x.world + (unsigned long long)src) * x.slot_bytes;
}

// One CTA.  ROWBAND = false (scans sharded over ranks): publishes counts[0..B], then adds the number of points of
// all lower ranks to col_base[B][W]; counts_all receives every scan's count in rank order followed by the grand total.
// ROWBAND = true (one capture, rank r holds the r-th band of rows): publishes col_tot[W]; col_base[x] becomes
// sum_{x' < x} sum_r tot_r[x'] + sum_{r' < rank} tot_r'[x], the position of this band's run inside column x of the
// whole image; counts_all = per-rank totals followed by the grand total.
template <bool ROWBAND>
__global__ void __launch_bounds__(kExchangeThreads) exchange_offsets_kernel(const ExchangeArgs x, int B, int W,
                                                                            const long long* __restrict__ counts,
                                                                            const uint32_t* __restrict__ col_tot,
                                                                            long long* __restrict__ col_base,
                                                                            long long* __restrict__ counts_all) {
  __shared__ unsigned long long seq_s;
  __shared__ int failed_s;
  __shared__ long long base_s, grand_s;
  __shared__ long long cnt_off_s[kMaxWorld + 1];
  __shared__ long long scan_s[kExchangeThreads / 32];
  const int tid = threadIdx.x;
  ExchangeLocal* local = reinterpret_cast<ExchangeLocal*>(x.ctrl[x.rank]);
  if (tid == 0) {
    seq_s = ++local->pub_step;
    failed_s = 0;
  }
  __syncthreads();
  const unsigned long long seq = seq_s;

  // ---- publish into every rank's control block (my own included: one code path)
  for (int d = 0; d < x.world; ++d) {
    uint8_t* slot = slot_ptr(x, d, seq, x.rank);
    SlotHeader* hdr = reinterpret_cast<SlotHeader*>(slot);
    if (ROWBAND) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(slot + sizeof(SlotHeader) + (size_t)x.max_scans * 8);
      for (int i = tid; i < W; i += kExchangeThreads) dst[i] = col_tot[i];
    } else {
      long long* dst = reinterpret_cast<long long*>(slot + sizeof(SlotHeader));
      for (int i = tid; i < B; i += kExchangeThreads) dst[i] = counts[i];
    }
    if (tid == 0) { hdr->total = counts[B]; hdr->n_counts = B; }
  }
  __syncthreads();
  if (tid < x.world) {
    __threadfence_system();
    st_release_sys_u64(&reinterpret_cast<SlotHeader*>(slot_ptr(x, tid, seq, x.rank))->seq, seq);
  }

  // ---- wait for every rank's slot of this step in MY control block
  if (tid < x.world) {
    const SlotHeader* hdr = reinterpret_cast<const SlotHeader*>(slot_ptr(x, x.rank, seq, tid));
    if (!wait_flag(&hdr->seq, seq, x.timeout_ns)) failed_s = 1;
  }
  __syncthreads();
  if (failed_s) {
    if (tid == 0) local->error = 1;
    return;
  }

  if (tid == 0) {
    long long base = 0, grand = 0, off = 0;
    for (int r = 0; r < x.world; ++r) {
      const SlotHeader* hdr = reinterpret_cast<const SlotHeader*>(slot_ptr(x, x.rank, seq, r));
      const long long t = __ldcg(&hdr->total);      // written by rank r: read at L2, the point of coherence
      if (r < x.rank) base += t;
      grand += t;
      cnt_off_s[r] = off;
      off += ROWBAND ? 1 : __ldcg(&hdr->n_counts);
    }
    cnt_off_s[x.world] = off;
    base_s = base; grand_s = grand;
    local->rank_base = base; local->grand_total = grand;
  }
  __syncthreads();

  if (!ROWBAND) {
    if (counts_all) {
      for (int r = 0; r < x.world; ++r) {
        const uint8_t* slot = slot_ptr(x, x.rank, seq, r);
        const long long* src = reinterpret_cast<const long long*>(slot + sizeof(SlotHeader));
        const long long n = cnt_off_s[r + 1] - cnt_off_s[r];
        for (long long i = tid; i < n; i += kExchangeThreads) counts_all[cnt_off_s[r] + i] = __ldcg(src + i);
      }
      if (tid == 0) counts_all[cnt_off_s[x.world]] = grand_s;
    }
    const long long base = base_s;
    const long long m = (long long)B * W;
    if (base != 0 && col_base)           // col_base == nullptr: counts only (the records are pushed as a block)
      for (long long i = tid; i < m; i += kExchangeThreads) col_base[i] += base;
    return;
  }

  // ---- row bands: global column bases.  Thread t owns the columns [t*chunk, (t+1)*chunk).
  if (counts_all) {
    if (tid < x.world) counts_all[tid] = __ldcg(&reinterpret_cast<const SlotHeader*>(slot_ptr(x, x.rank, seq, tid))->total);
    if (tid == 0) counts_all[x.world] = grand_s;
  }
  const int chunk = (W + kExchangeThreads - 1) / kExchangeThreads;
  const int c0 = tid * chunk, c1 = min(W, c0 + chunk);
  auto col_of = [&](int r, int c) {
    const uint8_t* slot = slot_ptr(x, x.rank, seq, r);
    return (long long)__ldcg(reinterpret_cast<const uint32_t*>(slot + sizeof(SlotHeader) + (size_t)x.max_scans * 8) + c);
  };
  long long mine = 0;
  for (int c = c0; c < c1; ++c)
    for (int r = 0; r < x.world; ++r) mine += col_of(r, c);
  // exclusive block scan of the per-thread sums
  const int lane = tid & 31, warp = tid >> 5;
  long long inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  if (lane == 31) scan_s[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    long long w = lane < kExchangeThreads / 32 ? scan_s[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    if (lane < kExchangeThreads / 32) scan_s[lane] = w;     // inclusive over warps
  }
  __syncthreads();
  long long run = (warp ? scan_s[warp - 1] : 0) + inc - mine;
  for (int c = c0; c < c1; ++c) {
    long long below = 0, all = 0;
    for (int r = 0; r < x.world; ++r) {
      const long long t = col_of(r, c);
      if (r < x.rank) below += t;
      all += t;
    }
    col_base[c] = run + below;
    run += all;
  }
}

// One warp.  Lanes < world whose bit is set in dst_mask tell that rank "my records of this step have landed"
// (the emit kernel before this one in stream order has completed; the fence orders its peer writes before the
// flag).  With `wait` the rank then waits until every rank has told it the same.
__global__ void __launch_bounds__(32) exchange_arrive_kernel(const ExchangeArgs x, unsigned int dst_mask, int wait) {
  const int lane = threadIdx.x;
  ExchangeLocal* local = reinterpret_cast<ExchangeLocal*>(x.ctrl[x.rank]);
  unsigned long long seq = 0;
  if (lane == 0) seq = ++local->sig_step;
  seq = __shfl_sync(0xffffffffu, seq, 0);
  if (lane < x.world && ((dst_mask >> lane) & 1u)) {
    __threadfence_system();
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(x.ctrl[lane] + x.done_off) + x.rank, seq);
  }
  if (wait && lane < x.world) {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(x.ctrl[x.rank] + x.done_off) + lane;
    if (!wait_flag(flag, seq, x.timeout_ns)) local->error = 2;
  }
}

struct PushDsts {
  uint4* cloud[kMaxWorld];
  int n;
};

// Copy this rank's records of the step, src[0 .. n) (src_rebased: src is laid out like the gathered cloud and the
// records start at src + rank_base), to [rank_base, rank_base + n) of every destination cloud (peer-mapped or local).
// n = *total_ptr and rank_base come from device memory: no host synchronisation.  The first few records are copied
// apart so that the warps of the main loop write whole, aligned 128-byte lines of the destination.
constexpr int kPushThreads = 512;
__global__ void __launch_bounds__(kPushThreads) exchange_push_kernel(const ExchangeArgs x, const uint4* __restrict__ src,
                                                                      int src_rebased, const long long* __restrict__ total_ptr,
                                                                      const PushDsts d, long long max_points) {
  ExchangeLocal* local = reinterpret_cast<ExchangeLocal*>(x.ctrl[x.rank]);
  if (local->error == 1) return;                          // the count exchange timed out: rank_base is not valid
  const long long base = local->rank_base;
  const long long n = *total_ptr;
  if (n <= 0) return;
  if (base < 0 || base + n > max_points) {                // the gathered cloud is too small: nothing is written
    if (blockIdx.x == 0 && threadIdx.x == 0) local->error = 3;
    return;
  }
  const uint4* s = src + (src_rebased ? base : 0);
  const long long head = min(n, (long long)((8 - (base & 7)) & 7));
  const long long tid = (long long)blockIdx.x * kPushThreads + threadIdx.x;
  const long long nthr = (long long)gridDim.x * kPushThreads;
  if (tid < head) {
    const uint4 v = s[tid];
    for (int k = 0; k < d.n; ++k) d.cloud[k][base + tid] = v;
  }
  for (long long i = head + tid; i < n; i += 4 * nthr) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i + j * nthr < n) v[j] = __ldcs(s + i + j * nthr);
    for (int k = 0; k < d.n; ++k) {
      uint4* o = d.cloud[k] + base;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i + j * nthr < n) o[i + j * nthr] = v[j];
    }
  }
}

}  // namespace sl

using namespace sl;

struct SlExchange {
  int device = 0;
  ExchangeArgs args;
  ExchangeLayout layout;
};

extern "C" {

// ---- peer memory (CUDA IPC)
int sl_peer_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) { set_error("bad argument"); return kErrBadArg; }
  *ptr = nullptr;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return kErrAlloc; }
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaMemset"); }
  *ptr = p;
  return kOk;
}

int sl_peer_free(void* ptr) {
  if (ptr) SL_CUDA(cudaFree(ptr));
  return kOk;
}

int sl_peer_export(const void* ptr, void* handle64) {
  if (!ptr || !handle64) { set_error("null argument"); return kErrBadArg; }
  static_assert(sizeof(cudaIpcMemHandle_t) == SL_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  SL_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, sizeof(h));
  return kOk;
}

int sl_peer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) { set_error("null argument"); return kErrBadArg; }
  *ptr = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  SL_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return kOk;
}

int sl_peer_close(void* ptr) {
  if (ptr) SL_CUDA(cudaIpcCloseMemHandle(ptr));
  return kOk;
}

// ---- exchange object
size_t sl_exchange_ctrl_bytes(int world, int max_scans, int max_width) {
  if (world < 1 || world > kMaxWorld || max_scans < 1 || max_width < 1) return 0;
  return exchange_layout(world, max_scans, max_width).total_bytes;
}

int sl_exchange_create(SlExchange** out, int rank, int world, void* const* ctrl_ptrs, int max_scans, int max_width,
                       double timeout_s) {
  if (!out) { set_error("out is null"); return kErrBadArg; }
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) { set_error("bad rank/world (%d/%d, at most %d ranks)", rank, world, kMaxWorld); return kErrBadArg; }
  if (max_scans < 1 || max_width < 1 || !ctrl_ptrs) { set_error("bad argument"); return kErrBadArg; }
  for (int r = 0; r < world; ++r)
    if (!ctrl_ptrs[r]) { set_error("control block of rank %d is null", r); return kErrBadArg; }
  int device = 0;
  SL_CUDA(cudaGetDevice(&device));
  SlExchange* x = new (std::nothrow) SlExchange();
  if (!x) { set_error("out of host memory"); return kErrAlloc; }
  x->device = device;
  x->layout = exchange_layout(world, max_scans, max_width);
  memset(&x->args, 0, sizeof(x->args));
  for (int r = 0; r < world; ++r) x->args.ctrl[r] = static_cast<uint8_t*>(ctrl_ptrs[r]);
  x->args.rank = rank; x->args.world = world;
  x->args.slot_bytes = x->layout.slot_bytes; x->args.slots_off = x->layout.slots_off; x->args.done_off = x->layout.done_off;
  x->args.max_scans = max_scans; x->args.max_width = max_width;
  if (!(timeout_s > 0)) timeout_s = 20.0;
  x->args.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  *out = x;
  return kOk;
}

void sl_exchange_destroy(SlExchange* x) { delete x; }

int sl_exchange_offsets_by_scan(SlExchange* x, int B, int H, int W, void* workspace, const int64_t* counts,
                                int64_t* counts_all, void* stream) {
  if (!x || !counts) { set_error("null argument"); return kErrBadArg; }
  if (B < 1 || B > x->layout.max_scans || H < 1 || W < 1) { set_error("B=%d outside 1..max_scans=%d", B, x->layout.max_scans); return kErrBadArg; }
  long long* col_base = workspace ? carve_workspace(workspace, B, H, W).col_base : nullptr;
  exchange_offsets_kernel<false><<<1, kExchangeThreads, 0, (cudaStream_t)stream>>>(
      x->args, B, W, (const long long*)counts, nullptr, col_base, (long long*)counts_all);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int sl_exchange_offsets_row_band(SlExchange* x, int H, int W, void* workspace, const int64_t* counts,
                                 int64_t* counts_all, void* stream) {
  if (!x || !workspace || !counts) { set_error("null argument"); return kErrBadArg; }
  if (H < 1 || W < 1 || W > x->layout.max_width) { set_error("W=%d outside 1..max_width=%d", W, x->layout.max_width); return kErrBadArg; }
  const CloudWorkspace c = carve_workspace(workspace, 1, H, W);
  exchange_offsets_kernel<true><<<1, kExchangeThreads, 0, (cudaStream_t)stream>>>(
      x->args, 1, W, (const long long*)counts, c.col_tot, c.col_base, (long long*)counts_all);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int sl_exchange_push(SlExchange* x, const float* src, int src_rebased, const int64_t* total, float* const* clouds,
                     int n_clouds, int64_t max_points, int n_ctas, void* stream) {
  if (!x || !src || !total || !clouds) { set_error("null argument"); return kErrBadArg; }
  if (n_clouds < 1 || n_clouds > kMaxWorld) { set_error("n_clouds must be 1..%d", kMaxWorld); return kErrBadArg; }
  if (max_points < 1) { set_error("max_points must be positive"); return kErrBadArg; }
  if ((uintptr_t)src % 16 || (uintptr_t)total % 8) { set_error("src must be 16-byte aligned, total 8-byte"); return kErrBadArg; }
  PushDsts d;
  memset(&d, 0, sizeof(d));
  for (int k = 0; k < n_clouds; ++k) {
    if (!clouds[k] || (uintptr_t)clouds[k] % 16) { set_error("cloud buffers must be non-null and 16-byte aligned"); return kErrBadArg; }
    d.cloud[k] = reinterpret_cast<uint4*>(clouds[k]);
  }
  d.n = n_clouds;
  if (n_ctas <= 0) n_ctas = 32;
  if (n_ctas > 1024) n_ctas = 1024;
  exchange_push_kernel<<<n_ctas, kPushThreads, 0, (cudaStream_t)stream>>>(x->args, reinterpret_cast<const uint4*>(src),
                                                                          src_rebased ? 1 : 0, (const long long*)total, d,
                                                                          (long long)max_points);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int sl_exchange_arrive(SlExchange* x, uint32_t dst_mask, int wait, void* stream) {
  if (!x) { set_error("null argument"); return kErrBadArg; }
  exchange_arrive_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(x->args, dst_mask, wait ? 1 : 0);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int sl_exchange_status(SlExchange* x, int* error, int64_t* rank_base, int64_t* grand_total) {
  if (!x) { set_error("null argument"); return kErrBadArg; }
  ExchangeLocal l;
  SL_CUDA(cudaMemcpy(&l, x->args.ctrl[x->args.rank], sizeof(l), cudaMemcpyDeviceToHost));
  if (error) *error = (int)l.error;
  if (rank_base) *rank_base = l.rank_base;
  if (grand_total) *grand_total = l.grand_total;
  return kOk;
}

}  // extern "C"

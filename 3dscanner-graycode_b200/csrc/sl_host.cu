// Host-side format conversions at the NumPy boundary of the reference's API (no device work in this file).
//
// The reference hands the decode a float64 frame stack (src/3-capture_decode.py:66-70: 763 MB for one 1080p scan that
// holds 95 MB of uint8 samples) and expects int64 index images, a (3,N) float64 point array and (N,3) float64 colours
// back (decode_codes.py:189-229, triangulate.py:64-69,95).  The GPU path computes on uint8 samples, int16 indices and
// packed 16-byte records; these helpers do the dtype changes on the host in one multi-threaded pass each, straight
// from / into pinned staging buffers, so that only the narrow form crosses PCIe.  Every conversion is exact or fails.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <vector>
#include "sl_common.cuh"

namespace sl {
namespace {

int pick_threads(int threads, size_t n, size_t min_per_thread) {
  if (threads <= 0) {
    const unsigned hw = std::thread::hardware_concurrency();
    threads = (int)std::min<unsigned>(hw ? hw : 1u, 32u);
  }
  const size_t useful = std::max<size_t>(1, n / std::max<size_t>(1, min_per_thread));
  return (int)std::min<size_t>((size_t)threads, useful);
}

// fn(begin, end) over [0, n) in contiguous chunks (multiples of `grain` elements), one chunk per thread
template <class Fn>
void parallel_chunks(size_t n, int threads, size_t grain, Fn fn) {
  if (threads <= 1 || n == 0) { fn((size_t)0, n); return; }
  size_t per = (n + threads - 1) / threads;
  per = (per + grain - 1) / grain * grain;
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 1; t < threads; ++t) {
    const size_t b = std::min(n, per * t), e = std::min(n, per * (t + 1));
    if (b < e) pool.emplace_back([=] { fn(b, e); });
  }
  fn((size_t)0, std::min(n, per));
  for (auto& th : pool) th.join();
}

// exact narrowing of grey levels: a value passes iff it is an integer in 0..255 (NaN, 300, -1, 0.5 do not)
template <class T>
size_t narrow_u8_range(const T* src, uint8_t* dst, size_t b, size_t e) {
  size_t bad = 0;
  for (size_t i = b; i < e; ++i) {
    const T v = src[i];
    const bool in = (v >= (T)0) && (v <= (T)255);
    const uint8_t u = in ? (uint8_t)(int)v : (uint8_t)0;
    bad += (size_t)!(in && (T)u == v);
    dst[i] = u;
  }
  return bad;
}

#if defined(__x86_64__) && defined(__GNUC__)
#define SL_HOST_AVX2 1
// 8 doubles per step: truncate to int32 (out-of-range and NaN give INT_MIN), require the round trip to be exact and the
// integer to be in 0..255, pack to bytes.  The plain loop above costs ~2.5 ns per sample and thread; this one is
// bound by reading the stack.
__attribute__((target("avx2"))) size_t narrow_f64_avx2(const double* src, uint8_t* dst, size_t b, size_t e) {
  size_t i = b;
  __m256i bad = _mm256_setzero_si256();
  const __m128i hi = _mm_set1_epi32(~255);
  for (; i + 8 <= e; i += 8) {
    const __m256d a = _mm256_loadu_pd(src + i), c = _mm256_loadu_pd(src + i + 4);
    const __m128i ia = _mm256_cvttpd_epi32(a), ic = _mm256_cvttpd_epi32(c);
    const __m256d ne_a = _mm256_cmp_pd(_mm256_cvtepi32_pd(ia), a, _CMP_NEQ_UQ);     // unordered (NaN) counts as different
    const __m256d ne_c = _mm256_cmp_pd(_mm256_cvtepi32_pd(ic), c, _CMP_NEQ_UQ);
    bad = _mm256_or_si256(bad, _mm256_castpd_si256(_mm256_or_pd(ne_a, ne_c)));
    const __m128i out_of_range = _mm_and_si128(_mm_or_si128(ia, ic), hi);           // any bit above 255, or negative
    bad = _mm256_or_si256(bad, _mm256_castsi128_si256(out_of_range));
    const __m128i w = _mm_packs_epi32(ia, ic);                                      // 8 x int16
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i), _mm_packus_epi16(w, w));  // 8 x uint8
  }
  size_t nbad = _mm256_testz_si256(bad, bad) ? 0 : 1;
  if (nbad) {                        // recount exactly (the error message reports how many)
    nbad = 0;
    for (size_t k = b; k < i; ++k) {
      const double v = src[k];
      const bool in = (v >= 0.0) && (v <= 255.0);
      nbad += (size_t)!(in && (double)(uint8_t)(int)(in ? v : 0.0) == v);
    }
  }
  return nbad + narrow_u8_range(src, dst, i, e);
}
#endif

template <class T>
size_t narrow_u8_dispatch(const T* src, uint8_t* dst, size_t b, size_t e) { return narrow_u8_range(src, dst, b, e); }
#if defined(SL_HOST_AVX2)
template <>
size_t narrow_u8_dispatch<double>(const double* src, uint8_t* dst, size_t b, size_t e) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  return avx2 ? narrow_f64_avx2(src, dst, b, e) : narrow_u8_range(src, dst, b, e);
}
#endif

template <class T>
int narrow_u8(const T* src, uint8_t* dst, size_t n, int threads) {
  std::atomic<size_t> bad{0};
  parallel_chunks(n, pick_threads(threads, n, 1u << 16), 64, [&](size_t b, size_t e) {
    const size_t k = narrow_u8_dispatch(src, dst, b, e);
    if (k) bad.fetch_add(k, std::memory_order_relaxed);
  });
  if (bad.load()) {
    set_error("frame stack must hold integer grey levels in 0..255 (%zu samples do not)", bad.load());
    return kErrBadArg;
  }
  return kOk;
}

}  // namespace
}  // namespace sl

using namespace sl;

extern "C" {

int sl_host_narrow_u8(const void* src, int src_type, size_t n, uint8_t* dst, int threads) {
  if ((!src || !dst) && n) { set_error("null buffer"); return kErrBadArg; }
  if (src_type == 0) {
    if ((uintptr_t)src % 8) { set_error("float64 source must be 8-byte aligned"); return kErrBadArg; }
    return narrow_u8((const double*)src, dst, n, threads);
  }
  if (src_type == 1) {
    if ((uintptr_t)src % 4) { set_error("float32 source must be 4-byte aligned"); return kErrBadArg; }
    return narrow_u8((const float*)src, dst, n, threads);
  }
  set_error("src_type must be 0 (float64) or 1 (float32)");
  return kErrBadArg;
}

int sl_host_widen_idx(const int16_t* src, size_t n, int64_t* dst, int threads) {
  if ((!src || !dst) && n) { set_error("null buffer"); return kErrBadArg; }
  if ((uintptr_t)src % 2 || (uintptr_t)dst % 8) { set_error("misaligned index buffer"); return kErrBadArg; }
  parallel_chunks(n, pick_threads(threads, n, 1u << 16), 64, [=](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) dst[i] = (int64_t)src[i];
  });
  return kOk;
}

int sl_host_narrow_idx(const int64_t* src, size_t n, int16_t* dst, int threads) {
  if ((!src || !dst) && n) { set_error("null buffer"); return kErrBadArg; }
  if ((uintptr_t)src % 8 || (uintptr_t)dst % 2) { set_error("misaligned index buffer"); return kErrBadArg; }
  std::atomic<size_t> bad{0};
  parallel_chunks(n, pick_threads(threads, n, 1u << 16), 64, [&, src, dst](size_t b, size_t e) {
    size_t k = 0;
    for (size_t i = b; i < e; ++i) {
      const int64_t v = src[i];
      k += (size_t)(v < -1 || v > 32767);
      dst[i] = (int16_t)v;
    }
    if (k) bad.fetch_add(k, std::memory_order_relaxed);
  });
  if (bad.load()) { set_error("decoded indices must lie in -1..32767 (%zu do not)", bad.load()); return kErrBadArg; }
  return kOk;
}

int sl_host_unpack_cloud(const float* cloud, int64_t n, double* xyz_3xn, double* colors_nx3, int threads) {
  if (n < 0) { set_error("negative point count"); return kErrBadArg; }
  if ((!cloud || !xyz_3xn) && n) { set_error("null buffer"); return kErrBadArg; }
  if ((uintptr_t)cloud % 4 || (uintptr_t)xyz_3xn % 8 || (uintptr_t)colors_nx3 % 8) { set_error("misaligned buffer"); return kErrBadArg; }
  const size_t N = (size_t)n;
  parallel_chunks(N, pick_threads(threads, N, 1u << 14), 64, [=](size_t b, size_t e) {
    double* x = xyz_3xn;
    double* y = xyz_3xn + N;
    double* z = xyz_3xn + 2 * N;
    for (size_t i = b; i < e; ++i) {
      const float* r = cloud + 4 * i;
      x[i] = (double)r[0];
      y[i] = (double)r[1];
      z[i] = (double)r[2];
    }
    if (colors_nx3) {
      for (size_t i = b; i < e; ++i) {
        uint32_t c;
        std::memcpy(&c, cloud + 4 * i + 3, 4);
        // uint8 / 255 as an IEEE double division, the reference's img_white[j, i, :] / 255 (triangulate.py:69)
        colors_nx3[3 * i + 0] = (double)(c & 255u) / 255.0;
        colors_nx3[3 * i + 1] = (double)((c >> 8) & 255u) / 255.0;
        colors_nx3[3 * i + 2] = (double)((c >> 16) & 255u) / 255.0;
      }
    }
  });
  return kOk;
}

}  // extern "C"

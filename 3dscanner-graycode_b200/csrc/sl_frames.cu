// Capture front-end helpers next to the hot path (SURVEY 8(f) "next" rows):
//   bgr_to_gray_kernel   cv2.cvtColor(BGR2GRAY) as used by src/3-capture_decode.py:66 and
//                        scanner/grayCode/decode_codes.py:70-87 (to_gray): 15-bit fixed point
//                        (3735 B + 19235 G + 9798 R + 16384) >> 15, bit-exact with OpenCV's 8-bit path.
//   pair_diff_kernel     per consecutive frame pair, the number of elements with |a - b| > thresh
//                        (scanner/grayCode/decode_codes.py:49-54, remove_bad_images).
// Both are single-pass streaming kernels: bound by HBM, 4 bytes (3 in + 1 out) resp. 2 bytes per element.
#include "sl_common.cuh"
#include "sl_kernels.h"

namespace sl {

__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
  return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// 4 pixels per thread: 12 bytes in (three aligned words), 4 bytes out.
__global__ void __launch_bounds__(256) bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, long long npix,
                                                          uint8_t* __restrict__ gray) {
  const long long quads = npix >> 2;
  const bool aligned = (((uintptr_t)bgr | (uintptr_t)gray) & 3) == 0;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads;
       q += (long long)gridDim.x * blockDim.x) {
    if (aligned) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(bgr) + q * 3;
      const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
      const uint32_t g0 = gray_of(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff);
      const uint32_t g1 = gray_of(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff);
      const uint32_t g2 = gray_of((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff);
      const uint32_t g3 = gray_of((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24);
      reinterpret_cast<uint32_t*>(gray)[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
      for (int k = 0; k < 4; ++k) {
        const uint8_t* p = bgr + (q * 4 + k) * 3;
        gray[q * 4 + k] = (uint8_t)gray_of(p[0], p[1], p[2]);
      }
    }
  }
  // tail (npix mod 4)
  const long long t = (quads << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < npix) gray[t] = (uint8_t)gray_of(bgr[t * 3], bgr[t * 3 + 1], bgr[t * 3 + 2]);
}

// counts[k] = #{ i : |frames[k+1][i] - frames[k][i]| > thresh },  k = 0 .. n_frames-2
__global__ void __launch_bounds__(256) pair_diff_kernel(const uint8_t* __restrict__ frames, long long elems,
                                                        size_t frame_pitch, int thresh,
                                                        unsigned long long* __restrict__ counts) {
  const int k = blockIdx.y;
  const uint8_t* a = frames + (size_t)k * frame_pitch;
  const uint8_t* b = a + frame_pitch;
  unsigned int c = 0;
  const bool aligned = ((((uintptr_t)a | (uintptr_t)b) & 3) == 0);
  const long long words = aligned ? (elems >> 2) : 0;
  const uint32_t t4 = 0x01010101u * (uint32_t)min(max(thresh, 0), 255);
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < words;
       w += (long long)gridDim.x * blockDim.x) {
    const uint32_t d = __vabsdiffu4(__ldg(reinterpret_cast<const uint32_t*>(a) + w),
                                    __ldg(reinterpret_cast<const uint32_t*>(b) + w));
    c += __popc(__vcmpgtu4(d, t4)) >> 3;         // 0xff per byte above the threshold
  }
  for (long long i = (words << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < elems;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = abs((int)a[i] - (int)b[i]);
    c += d > thresh ? 1u : 0u;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counts + k, (unsigned long long)c);
}

// Projector pattern sequence (scanner/grayCode/generate_codes.py:34-81) written straight into device memory:
// frame 0 black, 1 white, 2+2j column bit j (MSB first), 3+2k row bit k (LSB first), +2n the inverses;
// portrait panels (height > width) get all-black patterns like the reference (stripe width 0).
__global__ void __launch_bounds__(256) pattern_kernel(int pw, int ph, int n, uint8_t* __restrict__ out) {
  const long long npix = (long long)pw * ph;
  const long long total = npix * (4 * n + 2);
  const bool draw = pw >= ph;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / npix);
    const long long p = i - (long long)f * npix;
    const int y = (int)(p / pw), x = (int)(p - (long long)y * pw);
    uint8_t val = 0;
    if (f == 1) val = 255;
    else if (f >= 2) {
      const int g = f - 2;
      const bool inverse = g >= 2 * n;
      const int k = inverse ? g - 2 * n : g;
      if (draw) {
        uint32_t bit;
        if ((k & 1) == 0) { const uint32_t gray = (uint32_t)x ^ ((uint32_t)x >> 1); bit = (gray >> (n - 1 - (k >> 1))) & 1u; }
        else              { const uint32_t gray = (uint32_t)y ^ ((uint32_t)y >> 1); bit = (gray >> (k >> 1)) & 1u; }
        val = (bit ^ (inverse ? 1u : 0u)) ? 255 : 0;
      }
    }
    out[i] = val;
  }
}

int launch_patterns(int pw, int ph, int n, uint8_t* out, int num_sms, cudaStream_t st) {
  const long long total = (long long)pw * ph * (4 * n + 2);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms * 32) blocks = (long long)num_sms * 32;
  pattern_kernel<<<(int)blocks, 256, 0, st>>>(pw, ph, n, out);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_bgr_to_gray(const uint8_t* bgr, long long npix, uint8_t* gray, int num_sms, cudaStream_t st) {
  if (npix == 0) return kOk;
  long long blocks = ((npix >> 2) + 255) / 256;
  if (blocks > (long long)num_sms * 16) blocks = (long long)num_sms * 16;
  if (blocks < 1) blocks = 1;
  bgr_to_gray_kernel<<<(int)blocks, 256, 0, st>>>(bgr, npix, gray);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

int launch_pair_diff(const uint8_t* frames, int n_frames, long long elems, size_t frame_pitch, int thresh,
                     unsigned long long* counts, int num_sms, cudaStream_t st) {
  if (n_frames < 2) return kOk;
  SL_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n_frames - 1) * sizeof(unsigned long long), st));
  long long bx = ((elems >> 2) + 255) / 256;
  const long long cap = (long long)num_sms * 16 / (n_frames - 1) + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)(n_frames - 1));
  pair_diff_kernel<<<grid, 256, 0, st>>>(frames, elems, frame_pitch, thresh, counts);
  count_launch();
  SL_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace sl

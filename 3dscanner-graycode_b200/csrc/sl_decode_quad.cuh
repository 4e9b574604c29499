// Decode of 4 neighbouring pixels held one-per-byte-lane in 32-bit registers.
// Shared by the TMA-staged kernel (loads from shared memory) and the host logic
// checker in tests/hostcheck (loads from a NumPy buffer) -- see sl_pixel.cuh.
//
// Frame layout of the stack (scanner/grayCode/generate_codes.py:56-79 as consumed by
// scanner/grayCode/decode_codes.py:106-107,150-160), n = (F-2)/4:
//   frame 0 black, frame 1 white,
//   frame 2+2j      column ("h") bit j, MSB first        frame 2+2n+2j   its inverse
//   frame 3+2k      row ("v") bit k, LSB first           frame 3+2n+2k   its inverse
#pragma once
#include "sl_pixel.cuh"

namespace sl {

struct QuadOut {
  uint32_t h01, h23;   // int16 lanes: projector column of pixels {0,1} / {2,3}; 0xffff = -1
  uint32_t v01, v23;   // projector row
  uint32_t valid;      // bit 7 of lane k set iff pixel k decoded on both axes
  uint32_t slow;       // 1 if the quad needed the full rule evaluation (diagnostics: sl_decode_u8_stats)
};

// Tri-state bytes (1, 0, 0xff=-1) of one bit for 4 pixels (decode_codes.py:162-182 output).
SL_HD uint32_t tristate4(uint32_t one, uint32_t valid) {
  const uint32_t o = (one & kMsb4) >> 7;
  const uint32_t bad = ((~valid & kMsb4) >> 7) * 0xffu;
  return (o & ~bad) | bad;
}

// msb-flag accumulator -> n-bit code per pixel, Gray -> binary, -1 where invalid.
SL_HD void finish_axis(uint32_t accA, uint32_t accB, uint32_t allv, int n,
                       uint32_t& out01, uint32_t& out23) {
  const int cntA = n < 8 ? n : 8;
  const int cntB = n - cntA;
  // flag of code bit i sits at bit (8-cnt+i) of its lane: move to bit i
  const uint32_t a = (accA >> (8 - cntA)) & (0x01010101u * ((1u << cntA) - 1u));
  const uint32_t b = cntB ? ((accB >> (8 - cntB)) & (0x01010101u * ((1u << cntB) - 1u))) : 0u;
  uint32_t g01 = byte_perm(a, b, 0x5140u);     // lanes {A0|B0<<8, A1|B1<<8}
  uint32_t g23 = byte_perm(a, b, 0x7362u);
  // Gray -> binary in both 16-bit lanes (decode_codes.py:203-207)
  g01 ^= (g01 >> 1) & 0x7fff7fffu;  g23 ^= (g23 >> 1) & 0x7fff7fffu;
  g01 ^= (g01 >> 2) & 0x3fff3fffu;  g23 ^= (g23 >> 2) & 0x3fff3fffu;
  g01 ^= (g01 >> 4) & 0x0fff0fffu;  g23 ^= (g23 >> 4) & 0x0fff0fffu;
  g01 ^= (g01 >> 8) & 0x00ff00ffu;  g23 ^= (g23 >> 8) & 0x00ff00ffu;
  const uint32_t bad = ~allv & kMsb4;          // any bit -1 => index -1 (decode_codes.py:225-226)
  out01 = g01 | byte_perm(bad, 0u, 0x9988u);
  out23 = g23 | byte_perm(bad, 0u, 0xbbaau);
}

// Full evaluation of the four rules for both axes (decode_codes.py:172-182).  Kept out of line on the
// device: it is the slow path of decode_quad and must not add to the fast path's register budget.
template <int NB, bool CODES, class Load, class StoreCode>
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void full_rules(Load ld, StoreCode st, int n, const QuadThr& qt, uint32_t ek, QuadOut& out) {
  // column bits: code bit i (LSB = 0) is reference index j = n-1-i
  {
    uint32_t accA = 0, accB = 0, allv = kMsb4;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < n; ++i) {
      const int j = n - 1 - i;
      uint32_t one, valid;
      classify4(ld(2 + 2 * j), ld(2 + 2 * n + 2 * j), qt, ek, one, valid);
      allv &= valid;
      if (i < 8) accA = ((accA >> 1) & kLow4) | (one & kMsb4);
      else       accB = ((accB >> 1) & kLow4) | (one & kMsb4);
      if (CODES) st(0, j, tristate4(one, valid));
    }
    finish_axis(accA, accB, allv, n, out.h01, out.h23);
    out.valid = allv;
  }
  // row bits: stored LSB first (flipped by the caller at src/3-capture_decode.py:100)
  {
    uint32_t accA = 0, accB = 0, allv = kMsb4;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < n; ++i) {
      uint32_t one, valid;
      classify4(ld(3 + 2 * i), ld(3 + 2 * n + 2 * i), qt, ek, one, valid);
      allv &= valid;
      if (i < 8) accA = ((accA >> 1) & kLow4) | (one & kMsb4);
      else       accB = ((accB >> 1) & kLow4) | (one & kMsb4);
      if (CODES) st(1, i, tristate4(one, valid));
    }
    finish_axis(accA, accB, allv, n, out.v01, out.v23);
    out.valid &= allv;
  }
}

// NB > 0: compile-time bit count (fully unrolled); NB == 0: run-time n.
// ld(f) returns the 4 bytes of frame f for this quad.  st(axis, bit, tri) stores the
// tri-state bytes of one bit when CODES (axis 0 = h with bit = j MSB-first index,
// axis 1 = v with bit = k LSB-first index, i.e. the reference's array positions).
template <int NB, bool CODES, class Load, class StoreCode>
SL_HD QuadOut decode_quad(Load ld, StoreCode st, int n_rt, int e) {
  const int n = NB ? NB : n_rt;
  const uint32_t black = ld(0), white = ld(1);

  // L_max over the 3 finest column patterns + inverses, L_min over the 3 finest row
  // patterns + inverses (decode_codes.py:110-117), in two 16-bit lanes per register.
  uint32_t mxA, mxB, mnA, mnB;
  {
    const uint32_t h0 = ld(2 + 2 * (n - 1)), h1 = ld(2 + 2 * (n - 2)), h2 = ld(2 + 2 * (n - 3));
    const uint32_t i0 = ld(2 + 2 * n + 2 * (n - 1)), i1 = ld(2 + 2 * n + 2 * (n - 2)), i2 = ld(2 + 2 * n + 2 * (n - 3));
    mxA = max3_u16x2(widen_lo(h0), widen_lo(h1), widen_lo(h2));
    mxA = max3_u16x2(mxA, widen_lo(i0), widen_lo(i1));
    mxA = max3_u16x2(mxA, widen_lo(i2), 0u);
    mxB = max3_u16x2(widen_hi(h0), widen_hi(h1), widen_hi(h2));
    mxB = max3_u16x2(mxB, widen_hi(i0), widen_hi(i1));
    mxB = max3_u16x2(mxB, widen_hi(i2), 0u);
    const uint32_t v0 = ld(3), v1 = ld(5), v2 = ld(7);
    const uint32_t w0 = ld(3 + 2 * n), w1 = ld(5 + 2 * n), w2 = ld(7 + 2 * n);
    mnA = min3_u16x2(widen_lo(v0), widen_lo(v1), widen_lo(v2));
    mnA = min3_u16x2(mnA, widen_lo(w0), widen_lo(w1));
    mnA = min3_u16x2(mnA, widen_lo(w2), 0xffffffffu);
    mnB = min3_u16x2(widen_hi(v0), widen_hi(v1), widen_hi(v2));
    mnB = min3_u16x2(mnB, widen_hi(w0), widen_hi(w1));
    mnB = min3_u16x2(mnB, widen_hi(w2), 0xffffffffu);
  }

  QuadThr qt;
  qt.lo = 0; qt.hi1 = 0; qt.a = 0; qt.hard = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    const int w = (white >> (8 * k)) & 0xff, b = (black >> (8 * k)) & 0xff;
    const uint32_t mx = k < 2 ? mxA : mxB, mn = k < 2 ? mnA : mnB;
    const int lmax = (mx >> (16 * (k & 1))) & 0xffff, lmin = (mn >> (16 * (k & 1))) & 0xffff;
    const PixelThr t = pixel_thresholds(w, b, lmax, lmin, e);
    qt.lo |= (uint32_t)t.lo << (8 * k);
    qt.hi1 |= (uint32_t)t.hi1 << (8 * k);
    qt.a |= (uint32_t)(t.a << 7) << (8 * k);
    qt.hard |= (uint32_t)((t.easy ^ 1) << 7) << (8 * k);
  }
  qt.loL = qt.lo & kLow4;
  qt.hi1L = qt.hi1 & kLow4;
  const uint32_t ek = 0x01010101u * (uint32_t)(0x80 - (e + 1));

  QuadOut out;
  // ---- fast pass (both axes): single-threshold shortcut, 2 byte compares per bit.  If every bit of
  // the 4 pixels is decided by it the result is final (all bits known => all indices valid).
  bool need_full = CODES || qt.hard != 0;     // a pixel with hi1 > lo cannot use the shortcut at all
#if defined(__CUDA_ARCH__)
  // SIMT: if any quad of the warp needs the full evaluation, the warp will execute it anyway -- then every lane
  // takes it right away (it is exact for every pixel) instead of first running a fast pass beside idle lanes
  need_full = __any_sync(__activemask(), need_full);
#endif
  if (!need_full) {
    uint32_t hA = 0, hB = 0, vA = 0, vB = 0, undecided = qt.hard;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < n; ++i) {
      const int j = n - 1 - i;
      uint32_t gp, same;
      classify4_easy(ld(2 + 2 * j), ld(2 + 2 * n + 2 * j), qt, gp, same);
      undecided |= same;
      if (i < 8) hA = ((hA >> 1) & kLow4) | (gp & kMsb4);
      else       hB = ((hB >> 1) & kLow4) | (gp & kMsb4);
      classify4_easy(ld(3 + 2 * i), ld(3 + 2 * n + 2 * i), qt, gp, same);
      undecided |= same;
      if (i < 8) vA = ((vA >> 1) & kLow4) | (gp & kMsb4);
      else       vB = ((vB >> 1) & kLow4) | (gp & kMsb4);
    }
    need_full = (undecided & kMsb4) != 0;
    if (!need_full) {
      finish_axis(hA, hB, kMsb4, n, out.h01, out.h23);
      finish_axis(vA, vB, kMsb4, n, out.v01, out.v23);
      out.valid = kMsb4;
      out.slow = 0;
      return out;
    }
  }
  full_rules<NB, CODES>(ld, st, n, qt, ek, out);
  out.slow = 1;
  return out;
}

}  // namespace sl

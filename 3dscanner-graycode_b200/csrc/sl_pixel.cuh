// Per-pixel arithmetic shared by every decode kernel (and, compiled for the host,
// by the logic checker under tests/hostcheck -- test infrastructure, never a
// product fallback: the product entry points in sl_capi.cu only launch kernels).
//
// Reference semantics restated here (paths relative to the reference repo):
//   scanner/grayCode/decode_codes.py:90-122   direct/global estimate (fp64)
//   scanner/grayCode/decode_codes.py:162-182  per-bit tri-state rules, last rule wins
//   scanner/grayCode/decode_codes.py:189-229  Gray -> binary, any -1 bit => -1
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SL_HD __host__ __device__ __forceinline__
#else
#define SL_HD inline
#endif

namespace sl {

constexpr uint32_t kMsb4 = 0x80808080u;   // bit 7 of every byte lane
constexpr uint32_t kLow4 = 0x7f7f7f7fu;

// ---- IEEE fp64 ops that must never be contracted into FMAs ----------------
// The reference evaluates b_inv = w/(w+b); Ld = (Lmax-Lmin)*b_inv;
// Lg = 2.0*(Lmax-Ld)*b_inv; Lg+eps with one rounding per operation (NumPy).
#if defined(__CUDA_ARCH__)
SL_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
SL_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
SL_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
SL_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
SL_HD int d2i_ceil(double a) { return __double2int_ru(a); }
SL_HD int d2i_floor(double a) { return __double2int_rd(a); }
SL_HD uint32_t absdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
SL_HD uint32_t max3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
SL_HD uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
#else
// Host build (tests only) is compiled with -ffp-contract=off.
SL_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
SL_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
SL_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
SL_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
SL_HD int d2i_ceil(double a) {
  if (!(a == a)) return 0;
  if (a >= 2147483647.0) return 2147483647;
  if (a <= -2147483648.0) return -2147483647 - 1;
  int i = (int)a; return (double)i < a ? i + 1 : i;
}
SL_HD int d2i_floor(double a) {
  if (!(a == a)) return 0;
  if (a >= 2147483647.0) return 2147483647;
  if (a <= -2147483648.0) return -2147483647 - 1;
  int i = (int)a; return (double)i > a ? i - 1 : i;
}
SL_HD uint32_t absdiff4(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) {
    int x = (a >> (8 * k)) & 0xff, y = (b >> (8 * k)) & 0xff;
    r |= (uint32_t)(x > y ? x - y : y - x) << (8 * k);
  }
  return r;
}
SL_HD uint32_t max3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r = 0;
  for (int k = 0; k < 2; ++k) {
    uint32_t x = (a >> (16 * k)) & 0xffff, y = (b >> (16 * k)) & 0xffff, z = (c >> (16 * k)) & 0xffff;
    uint32_t m = x > y ? x : y; m = m > z ? m : z;
    r |= m << (16 * k);
  }
  return r;
}
SL_HD uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r = 0;
  for (int k = 0; k < 2; ++k) {
    uint32_t x = (a >> (16 * k)) & 0xffff, y = (b >> (16 * k)) & 0xffff, z = (c >> (16 * k)) & 0xffff;
    uint32_t m = x < y ? x : y; m = m < z ? m : z;
    r |= m << (16 * k);
  }
  return r;
}
#endif


// PRMT (byte permute).  Selector nibble k picks byte (nibble & 7) of {lo: bytes 0-3,
// hi: bytes 4-7}; nibble bit 3 replicates that byte's msb over the whole byte.
#if defined(__CUDA_ARCH__)
// (inline PTX: the __byte_perm intrinsic masks the selector with 0x7777 and drops the
// sign-replicate bit that prmt.b32 itself implements)
SL_HD uint32_t byte_perm(uint32_t lo, uint32_t hi, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
}
#else
SL_HD uint32_t byte_perm(uint32_t lo, uint32_t hi, uint32_t sel) {
  const uint64_t src = ((uint64_t)hi << 32) | lo;
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) {
    const uint32_t nib = (sel >> (4 * k)) & 0xf;
    uint32_t b = (uint32_t)(src >> (8 * (nib & 7))) & 0xff;
    if (nib & 8) b = (b & 0x80) ? 0xff : 0x00;
    r |= b << (8 * k);
  }
  return r;
}
#endif
// bytes {0,1} / {2,3} of x widened to two 16-bit lanes
SL_HD uint32_t widen_lo(uint32_t x) { return byte_perm(x, 0u, 0x4140u); }
SL_HD uint32_t widen_hi(uint32_t x) { return byte_perm(x, 0u, 0x4342u); }

// ---- BGR -> gray of 4 pixels -----------------------------------------------------
// cv2.cvtColor(frame, COLOR_BGR2GRAY) on 8-bit data (src/3-capture_decode.py:66, decode_codes.py:86):
// gray = (3735 B + 19235 G + 9798 R + 16384) >> 15.  The 12 bytes B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3 arrive
// as three little-endian words.  Evaluated as T = 2 * (3735 B + 19235 G + 9798 R) + 32768 < 2^24, whose byte 2 is
// the gray value.  The doubled coefficients are split into bytes (7470 = 29*256 + 46, 38470 = 150*256 + 70,
// 19596 = 76*256 + 140) so that T is two 4-way byte dot products (IDP4A) per source word touched; a pixel that
// straddles two words chains its dot products through the accumulator, bytes of other pixels meet zero
// coefficients.  All multiplies run on the FMA pipe -- the decode that follows is bound by the ALU pipe --
// and the only byte shuffling left is the final gather of the four result bytes (3 PRMT).
SL_HD uint32_t dot4_u8(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __dp4a(a, b, c);
#else
  for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 0xff) * ((b >> (8 * k)) & 0xff);
  return c;
#endif
}
SL_HD uint32_t gray4_from_bgr12(uint32_t w0, uint32_t w1, uint32_t w2) {
  constexpr uint32_t hB = 29u, hG = 150u, hR = 76u, lB = 46u, lG = 70u, lR = 140u, kRound = 32768u;
  const uint32_t t0 = dot4_u8(w0, hB | (hG << 8) | (hR << 16), 0u) * 256u + dot4_u8(w0, lB | (lG << 8) | (lR << 16), kRound);
  const uint32_t t1 = dot4_u8(w1, hG | (hR << 8), dot4_u8(w0, hB << 24, 0u)) * 256u +
                      dot4_u8(w1, lG | (lR << 8), dot4_u8(w0, lB << 24, kRound));
  const uint32_t t2 = dot4_u8(w2, hR, dot4_u8(w1, (hB << 16) | (hG << 24), 0u)) * 256u +
                      dot4_u8(w2, lR, dot4_u8(w1, (lB << 16) | (lG << 24), kRound));
  const uint32_t t3 = dot4_u8(w2, (hB << 8) | (hG << 16) | (hR << 24), 0u) * 256u +
                      dot4_u8(w2, (lB << 8) | (lG << 16) | (lR << 24), kRound);
  return byte_perm(byte_perm(t0, t1, 0x0062u), byte_perm(t2, t3, 0x0062u), 0x5410u);
}

// ---- per-pixel thresholds ---------------------------------------------------
// With integer-valued 8-bit samples and an integer eps=e, every per-bit test of
// decode_codes.py:172-182 reduces to byte compares against two per-pixel bytes:
//   (x + e <  Ld)   <=>  x <  lo,   lo  = ceil(Ld) - e        (Ld <= 255 always)
//   (x > Lg + e)    <=>  x >= hi1,  hi1 = floor(fl(Lg+e)) + 1
// and `A = Ld > Lg+e`.  NaN (white+black == 0) makes every comparison false:
// lo = 0, A = 0.  If hi1 would exceed 255 no byte can pass `x >= hi1`; since both
// remaining rules pair an `x >= hi1` test with an `x < lo` test, lo is forced to 0
// and hi1 is clamped.
struct PixelThr {
  int lo;    // 0..255
  int hi1;   // 0..255
  int a;     // 0/1
  int easy;  // hi1 <= lo: the single-threshold shortcut of classify4_easy applies
};

SL_HD PixelThr pixel_thresholds(int white, int black, int lmax, int lmin, int e) {
#if defined(__CUDA_ARCH__) && defined(SL_BINV_LOOKUP)
  // tuning option of sl_decode.cu: b_inv from a 256x256 table filled with the same rounded division
  const double b_inv = SL_BINV_LOOKUP(white, black);   // decode_codes.py:113
#else
  const double w = (double)white;
  const double s = (double)(white + black);           // exact
  const double b_inv = ddiv(w, s);                     // decode_codes.py:113
#endif
  const double ld = dmul((double)(lmax - lmin), b_inv);          // :119 (difference is exact)
  const double lg = dmul(dmul(2.0, dsub((double)lmax, ld)), b_inv);  // :120
  const double lge = dadd(lg, (double)e);
  PixelThr t;
  t.a = (ld > lge) ? 1 : 0;
  int lo = d2i_ceil(ld) - e;
  int hi1 = d2i_floor(lge) + 1;
  if (!(lge == lge)) { lo = 0; hi1 = 255; }            // NaN: nothing passes
  if (hi1 > 255) { lo = 0; hi1 = 255; }
  if (hi1 < 0) hi1 = 0;
  if (lo < 0) lo = 0;
  if (lo > 255) lo = 255;                               // unreachable (Ld <= 255); defensive
  t.lo = lo; t.hi1 = hi1;
  t.easy = (hi1 <= lo) ? 1 : 0;
  return t;
}

// ---- 4-lane byte SWAR --------------------------------------------------------
// a >= b per byte lane; result valid in bit 7 of each lane only.
// aH = a | kMsb4, bL = b & kLow4 are passed in so callers can hoist them.
SL_HD uint32_t ge4(uint32_t a, uint32_t aH, uint32_t b, uint32_t bL) {
  const uint32_t x = aH - bL;                 // bit7 = (a&0x7f) >= (b&0x7f), no inter-lane borrow
  return (a & ~b) | (~(a ^ b) & x);
}

struct QuadThr {           // thresholds of 4 neighbouring pixels, one per byte lane
  uint32_t lo, loL;        // lo and lo & 0x7f
  uint32_t hi1, hi1L;
  uint32_t a;              // 0x80 in lanes where A holds
  uint32_t hard;           // 0x80 in lanes where hi1 > lo (no shortcut)
};

// Shortcut for a pixel with hi1 <= lo.  If the normal and inverse samples lie on opposite sides of
// hi1 the bit is decided by rules 3/4 alone, which outrank rules 1/2 (decode_codes.py:178-182):
//   p >= hi1 and q < hi1 (<= lo)  =>  rule 4 fires                      => 1
//   p <  hi1 (<= lo) and q >= hi1 =>  rule 3 fires and rule 4 cannot    => 0
// gp = (p >= hi1) is then the bit; ~(gp ^ gq) marks bits this test cannot decide (msb flags).
SL_HD void classify4_easy(uint32_t p, uint32_t q, const QuadThr& t, uint32_t& gp, uint32_t& same) {
  gp = ge4(p, p | kMsb4, t.hi1, t.hi1L);
  const uint32_t gq = ge4(q, q | kMsb4, t.hi1, t.hi1L);
  same = ~(gp ^ gq);
}

// Classify one bit of 4 pixels.  p = normal pattern bytes, q = inverse pattern bytes.
// ek = per-lane constant (0x80 - (e+1)) replicated, requires 0 <= e <= 127.
// Outputs msb flags: one (bit value 1) and valid (bit value != -1).
SL_HD void classify4(uint32_t p, uint32_t q, const QuadThr& t, uint32_t ek,
                     uint32_t& one, uint32_t& valid) {
  const uint32_t pH = p | kMsb4, qH = q | kMsb4, qL = q & kLow4;
  const uint32_t p_ge_lo = ge4(p, pH, t.lo, t.loL);     // !(p + e < Ld)
  const uint32_t p_ge_hi = ge4(p, pH, t.hi1, t.hi1L);   //  p > Lg + e
  const uint32_t q_ge_lo = ge4(q, qH, t.lo, t.loL);
  const uint32_t q_ge_hi = ge4(q, qH, t.hi1, t.hi1L);
  const uint32_t p_ge_q = ge4(p, pH, q, qL);
  const uint32_t ad = absdiff4(p, q);
  const uint32_t far = ((ad & kLow4) + ek) | ad;        // |p-q| >= e+1
  const uint32_t r12 = t.a & far;                        // rule 1 or rule 2 fires (:172-176)
  const uint32_t r3 = ~p_ge_lo & q_ge_hi;                // :178-179 -> 0
  const uint32_t r4 = p_ge_hi & ~q_ge_lo;                // :181-182 -> 1
  valid = r12 | r3 | r4;
  one = r4 | (r12 & p_ge_q & ~r3);
}

// Gray -> binary for codes of up to 16 bits (decode_codes.py:203-207).
SL_HD uint32_t gray_to_binary16(uint32_t g) {
  g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8;
  return g;
}

}  // namespace sl

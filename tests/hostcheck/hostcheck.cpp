// TEST INFRASTRUCTURE: compiles the kernels' shared per-quad decode logic
// (csrc/sl_decode_quad.cuh) for the host so the byte-SWAR arithmetic can be checked
// against the oracle on a machine without a GPU.  Never loaded by the product.
#include <cstdint>
#include <cstring>
#include "sl_decode_quad.cuh"

extern "C" int hostcheck_decode(const uint8_t* frames, int F, long npix, int e,
                                int16_t* h_idx, int16_t* v_idx,
                                int8_t* h_codes, int8_t* v_codes) {
  if (F < 14 || (F - 2) % 4 || npix % 4) return -1;
  const int n = (F - 2) / 4;
  for (long px = 0; px < npix; px += 4) {
    auto ld = [&](int f) { uint32_t w; std::memcpy(&w, frames + (long)f * npix + px, 4); return w; };
    auto st = [&](int axis, int bit, uint32_t tri) {
      int8_t* dst = (axis ? v_codes : h_codes) + (long)bit * npix + px;
      std::memcpy(dst, &tri, 4);
    };
    sl::QuadOut o = h_codes ? sl::decode_quad<0, true>(ld, st, n, e)
                            : sl::decode_quad<0, false>(ld, st, n, e);
    std::memcpy(h_idx + px, &o.h01, 4); std::memcpy(h_idx + px + 2, &o.h23, 4);
    std::memcpy(v_idx + px, &o.v01, 4); std::memcpy(v_idx + px + 2, &o.v23, 4);
  }
  return 0;
}

// BGR -> gray of 4 pixels (sl_pixel.cuh gray4_from_bgr12): bgr holds npix*3 bytes, npix % 4 == 0
extern "C" int hostcheck_bgr_to_gray(const uint8_t* bgr, long npix, uint8_t* gray) {
  if (npix % 4) return -1;
  for (long px = 0; px < npix; px += 4) {
    uint32_t w[3];
    std::memcpy(w, bgr + px * 3, 12);
    const uint32_t g = sl::gray4_from_bgr12(w[0], w[1], w[2]);
    std::memcpy(gray + px, &g, 4);
  }
  return 0;
}

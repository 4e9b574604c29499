// Internal launcher interface between the kernel translation units and the C ABI (sl_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/slb200.h"

namespace sl {

constexpr int kMaxWorld = 16;   // ranks of one multi-GPU exchange (sl_peer.cu); also the destinations of one emission

void count_launch();   // bumps the process-wide kernel launch counter

// ---- stage 1 (sl_decode.cu)
int decode_device_init();   // once per device (under sl_capi.cu's device mutex): b_inv table, tensor-map encoder, smem opt-ins
int cloud_device_init();    // once per device: shared-memory opt-ins of the emit kernels (sl_cloud.cu)
// valid_bitmap (nullable): flat validity bits, scan b at valid_bitmap + b*bitmap_words
int launch_decode(const uint8_t* frames, int B, int F, long long npix, size_t frame_pitch,
                  size_t scan_pitch, double eps, int16_t* h_idx, int16_t* v_idx, int8_t* h_codes,
                  int8_t* v_codes, uint32_t* valid_bitmap, long long bitmap_words, unsigned long long* flags, int n_flags,
                  int num_sms, cudaStream_t st, bool bgr = false, int sm_reserve = 0,
                  unsigned long long* stats = nullptr);
int launch_decode_parts(const uint8_t* frames, int F, long long npix, size_t frame_pitch, double eps,
                        const double* ld_in, const double* lg_in, double* ld_out, double* lg_out,
                        int8_t* h_codes, int8_t* v_codes, int num_sms, cudaStream_t st);
int launch_codes_to_pixels(const int8_t* h_codes, const int8_t* v_codes, int R, int n,
                           long long npix, size_t run_pitch, int16_t* h_idx, int16_t* v_idx,
                           int num_sms, cudaStream_t st);

// ---- capture front end (sl_frames.cu)
int launch_patterns(int pw, int ph, int n, uint8_t* out, int num_sms, cudaStream_t st);
int launch_bgr_to_gray(const uint8_t* bgr, long long npix, uint8_t* gray, int num_sms, cudaStream_t st);
int launch_pair_diff(const uint8_t* frames, int n_frames, long long elems, size_t frame_pitch, int thresh,
                     unsigned long long* counts, int num_sms, cudaStream_t st);

// ---- stage 2 (sl_cloud.cu)
struct CloudWorkspace {      // views into the caller's workspace buffer
  uint32_t* bitmap;          // [B][bitmap_words] flat validity bits (bit i of word i/32 = pixel i)
  long long bitmap_words;
  uint32_t* band_cnt;        // [B][bands][W]  exclusive prefix of valid pixels down the column, per band
  uint32_t* col_bits;        // [B][bands][W]  bit r = pixel (band*32 + r, x) valid
  uint32_t* col_tot;         // [B][W]
  long long* col_base;       // [B][W] packed offset of the first point of column x of scan b
  unsigned long long* flags; // [B][strips + 1] published strip / scan totals of the offsets kernel (zeroed by the bitmap producer)
  int n_flags;
  int bands;
};
size_t cloud_workspace_bytes(int B, int H, int W);
CloudWorkspace carve_workspace(void* ws, int B, int H, int W);

int launch_compact_offsets(const int16_t* h, const int16_t* v, int B, int H, int W, void* ws,
                           long long* counts, cudaStream_t st);
int launch_offsets_from_bitmap(int B, int H, int W, void* ws, long long* counts, cudaStream_t st);
int launch_compact_offsets_boxed(const int16_t* h, const int16_t* v, const SlCalib& calib, int precision,
                                 const float* cam_rays, const float* proj_rays, double threshold, int B, int H, int W,
                                 void* ws, long long* counts, cudaStream_t st);
uint32_t* workspace_bitmap(void* ws, int B, int H, int W, long long* words_per_scan, unsigned long long** flags,
                           int* n_flags);
int launch_emit_pairs(const int16_t* h, const int16_t* v, const uint8_t* white, int B, int H, int W,
                      int proj_w, int proj_h, const void* ws, float* pairs, uint32_t* rgb32, cudaStream_t st);
int launch_lens_tables(const SlCalib& calib, int cam_w, int cam_h, int precision, float* cam_rays, float* proj_table,
                       cudaStream_t st);
int launch_emit_cloud(const int16_t* h, const int16_t* v, const uint8_t* white, const SlCalib& calib,
                      int precision, const float* cam_rays, const float* proj_rays, int B, int H, int W,
                      const void* ws, float* cloud, cudaStream_t st, float* const* extra = nullptr, int n_extra = 0,
                      int remote = 0, long long max_points = 0);
int launch_transpose_f2(const float* src, int rows, int cols, float* dst, cudaStream_t st);   // dst[c][r] = src[r][c], float2 entries
// K2r: the emission with lane = row (transposed lens tables, see sl_cloud.cu); fp32 table mode, local cloud, W % 32 == 0
int launch_emit_rows(const int16_t* h, const int16_t* v, const uint8_t* white, const SlCalib& calib, const float* cam_t,
                     int cam_rows, const float* proj_t, int B, int H, int W, const void* ws, float* cloud, cudaStream_t st,
                     float* const* extra = nullptr, int n_extra = 0, int remote = 0);
int launch_triangulate_points(const float* cam_pts, const float* proj_pts, long long N,
                              const SlCalib& calib, int precision, double* out, cudaStream_t st);
size_t filter_workspace_bytes(long long N);
int launch_filter_box(const double* pts, const double* colors, long long N, double threshold,
                      double* out_pts, double* out_colors, long long* out_count, void* ws,
                      cudaStream_t st);

int launch_filter_cloud(const float* cloud, long long N, double threshold, float* out_cloud,
                        long long* out_count, void* ws, cudaStream_t st);

// ---- cloud clean-up (sl_outlier.cu)
size_t outlier_workspace_bytes(long long N);
int launch_outlier_removal(const float* cloud, long long N, int nb_neighbors, double std_ratio, float* out_cloud,
                           long long* out_index, double* mean_dist, double* stats, long long* out_count, void* ws,
                           cudaStream_t st);

}  // namespace sl

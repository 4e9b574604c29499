// C ABI (include/slb200.h): argument checks, device gate, error strings, and the host-buffer
// pipeline (pinned-memory friendly H2D -> kernels -> D2H on a small ring of CUDA streams).
// Nothing here computes on the CPU: every entry point either launches the sm_100a kernels
// or fails.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "sl_common.cuh"
#include "sl_kernels.h"

namespace sl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return kErrCuda;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct DevInfo { int checked = 0; int ok = 0; int sms = 0; };
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

// returns the SM count of the current device, or a negative error
static int device_gate(int* sms) {
  int dev = 0;
  SL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return kErrDevice; }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevInfo& d = g_dev[dev];
  if (!d.checked) {
    int major = 0, minor = 0, n = 0;
    SL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SL_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    SL_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    const bool ok = (major == 10 && minor == 0);
    if (ok) {       // everything the launch paths need once per device, before any kernel of ours can be queued
      int trc = decode_device_init();
      if (trc == kOk) trc = cloud_device_init();
      if (trc) return trc;
    }
    d.ok = ok; d.sms = ok ? n : -(major * 10 + minor);
    d.checked = 1;
  }
  if (!d.ok) {
    set_error("slb200 is built for sm_100a only; device %d is sm_%d (no fallback path exists)", dev, -d.sms);
    return kErrDevice;
  }
  *sms = d.sms;
  return kOk;
}

// Entry points that select a device (pipelines, sl_device_check) restore the caller's current device on return.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int check_decode_shape(int B, int F, int H, int W) {
  if (B < 1 || H < 1 || W < 1) { set_error("B, H, W must be positive (got %d, %d, %d)", B, H, W); return kErrBadArg; }
  if (F < 14 || (F - 2) % 4) { set_error("F must be 4n+2 with n >= 3 (got %d)", F); return kErrBadArg; }
  if ((F - 2) / 4 > 15) { set_error("n = %d bits exceeds the int16 index range (max 15)", (F - 2) / 4); return kErrBadArg; }
  return kOk;
}

}  // namespace sl

using namespace sl;

extern "C" {

const char* sl_last_error(void) { return g_err; }
int sl_version(void) { return 100; }
int64_t sl_launch_count(void) { return (int64_t)g_launches.load(); }

int sl_device_check(int device) {
  DeviceGuard guard;
  if (device >= 0) SL_CUDA(cudaSetDevice(device));
  int sms = 0;
  return device_gate(&sms);
}

int sl_decode_u8(const uint8_t* frames, int B, int F, int H, int W, size_t frame_pitch,
                 size_t scan_pitch, double eps, int16_t* h_idx, int16_t* v_idx, int8_t* h_codes,
                 int8_t* v_codes, void* cloud_workspace, int sm_reserve, void* stream) {
  int rc = check_decode_shape(B, F, H, W);
  if (rc) return rc;
  if (!frames || !h_idx || !v_idx) { set_error("frames, h_idx, v_idx must be non-null"); return kErrBadArg; }
  if ((h_codes == nullptr) != (v_codes == nullptr)) { set_error("h_codes and v_codes must be given together"); return kErrBadArg; }
  const long long npix = (long long)H * W;
  if (frame_pitch < (size_t)npix) { set_error("frame_pitch smaller than H*W"); return kErrBadArg; }
  if (B > 1 && scan_pitch < frame_pitch * (size_t)F) { set_error("scan_pitch smaller than F*frame_pitch"); return kErrBadArg; }
  if (!(eps == eps)) { set_error("eps is NaN"); return kErrBadArg; }
  int sms = 0;
  if ((rc = device_gate(&sms))) return rc;
  uint32_t* bitmap = nullptr;
  unsigned long long* flags = nullptr;
  int n_flags = 0;
  long long words = 0;
  if (cloud_workspace) bitmap = workspace_bitmap(cloud_workspace, B, H, W, &words, &flags, &n_flags);
  if (sm_reserve < 0) { set_error("sm_reserve must be >= 0"); return kErrBadArg; }
  return launch_decode(frames, B, F, npix, frame_pitch, scan_pitch, eps, h_idx, v_idx, h_codes, v_codes,
                       bitmap, words, flags, n_flags, sms, (cudaStream_t)stream, false, sm_reserve);
}

int sl_decode_u8_stats(const uint8_t* frames, int B, int F, int H, int W, size_t frame_pitch, size_t scan_pitch,
                       double eps, int16_t* h_idx, int16_t* v_idx, uint64_t* stats, void* stream) {
  int rc = check_decode_shape(B, F, H, W);
  if (rc) return rc;
  if (!frames || !h_idx || !v_idx || !stats) { set_error("null buffer"); return kErrBadArg; }
  const long long npix = (long long)H * W;
  if (frame_pitch < (size_t)npix) { set_error("frame_pitch smaller than H*W"); return kErrBadArg; }
  if (B > 1 && scan_pitch < frame_pitch * (size_t)F) { set_error("scan_pitch smaller than F*frame_pitch"); return kErrBadArg; }
  if (!(eps == eps)) { set_error("eps is NaN"); return kErrBadArg; }
  int sms = 0;
  if ((rc = device_gate(&sms))) return rc;
  SL_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(uint64_t), (cudaStream_t)stream));
  return launch_decode(frames, B, F, npix, frame_pitch, scan_pitch, eps, h_idx, v_idx, nullptr, nullptr, nullptr, 0,
                       nullptr, 0, sms, (cudaStream_t)stream, false, 0, (unsigned long long*)stats);
}

int sl_decode_bgr_u8(const uint8_t* frames_bgr, int B, int F, int H, int W, size_t frame_pitch,
                     size_t scan_pitch, double eps, int16_t* h_idx, int16_t* v_idx,
                     void* cloud_workspace, int sm_reserve, void* stream) {
  int rc = check_decode_shape(B, F, H, W);
  if (rc) return rc;
  if (sm_reserve < 0) { set_error("sm_reserve must be >= 0"); return kErrBadArg; }
  if (!frames_bgr || !h_idx || !v_idx) { set_error("frames_bgr, h_idx, v_idx must be non-null"); return kErrBadArg; }
  const long long npix = (long long)H * W;
  if (frame_pitch < (size_t)npix * 3) { set_error("frame_pitch smaller than 3*H*W"); return kErrBadArg; }
  if (B > 1 && scan_pitch < frame_pitch * (size_t)F) { set_error("scan_pitch smaller than F*frame_pitch"); return kErrBadArg; }
  if (!(eps == eps)) { set_error("eps is NaN"); return kErrBadArg; }
  int sms = 0;
  if ((rc = device_gate(&sms))) return rc;
  uint32_t* bitmap = nullptr;
  unsigned long long* flags = nullptr;
  int n_flags = 0;
  long long words = 0;
  if (cloud_workspace) bitmap = workspace_bitmap(cloud_workspace, B, H, W, &words, &flags, &n_flags);
  return launch_decode(frames_bgr, B, F, npix, frame_pitch, scan_pitch, eps, h_idx, v_idx, nullptr, nullptr,
                       bitmap, words, flags, n_flags, sms, (cudaStream_t)stream, true, sm_reserve);
}

int sl_direct_indirect(const uint8_t* frames, int F, int H, int W, size_t frame_pitch, double* L_d,
                       double* L_g, void* stream) {
  int rc = check_decode_shape(1, F, H, W);
  if (rc) return rc;
  if (!frames || !L_d || !L_g) { set_error("null buffer"); return kErrBadArg; }
  if (frame_pitch < (size_t)H * W) { set_error("frame_pitch smaller than H*W"); return kErrBadArg; }
  int sms = 0;
  if ((rc = device_gate(&sms))) return rc;
  return launch_decode_parts(frames, F, (long long)H * W, frame_pitch, 0.0, nullptr, nullptr, L_d, L_g, nullptr,
                             nullptr, sms, (cudaStream_t)stream);
}

int sl_is_lit(const uint8_t* frames, int F, int H, int W, size_t frame_pitch, const double* L_d,
              const double* L_g, double eps, int8_t* h_codes, int8_t* v_codes, void* stream) {
  int rc = check_decode_shape(1, F, H, W);
  if (rc) return rc;
  if (!frames || !L_d || !L_g || !h_codes || !v_codes) { set_error("null buffer"); return kErrBadArg; }
  if (frame_pitch < (size_t)H * W) { set_error("frame_pitch smaller than H*W"); return kErrBadArg; }
  if (!(eps == eps)) { set_error("eps is NaN"); return kErrBadArg; }
  int sms = 0;
  if ((rc = device_gate(&sms))) return rc;
  return launch_decode_parts(frames, F, (long long)H * W, frame_pitch, eps, L_d, L_g, nullptr, nullptr, h_codes,
                             v_codes, sms, (cudaStream_t)stream);
}

int sl_codes_to_pixels(const int8_t* h_codes, const int8_t* v_codes, int R, int n, int64_t npix,
                       size_t run_pitch, int16_t* h_idx, int16_t* v_idx, void* stream) {
  if (!h_codes || !v_codes || !h_idx || !v_idx) { set_error("null buffer"); return kErrBadArg; }
  if (R < 1 || n < 1 || n > 15 || npix < 0) { set_error("bad R/n/npix (%d, %d, %lld)", R, n, (long long)npix); return kErrBadArg; }
  if (R > 1 && run_pitch < (size_t)n * (size_t)npix) { set_error("run_pitch smaller than n*npix"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  if (npix == 0) return kOk;
  return launch_codes_to_pixels(h_codes, v_codes, R, n, npix, run_pitch, h_idx, v_idx, sms, (cudaStream_t)stream);
}

int sl_pattern_sequence_u8(int proj_w, int proj_h, uint8_t* out, void* stream) {
  if (proj_w < 2 || proj_h < 1 || !out) { set_error("bad argument"); return kErrBadArg; }
  int n = 0;
  while ((1 << n) < (proj_w > proj_h ? proj_w : proj_h)) ++n;       // ceil(log2(max(w, h))), generate_codes.py:22-25
  if (n < 1 || n > 16) { set_error("panel size out of range"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_patterns(proj_w, proj_h, n, out, sms, (cudaStream_t)stream);
}

int sl_bgr_to_gray_u8(const uint8_t* bgr, int64_t npix, uint8_t* gray, void* stream) {
  if (npix < 0 || (npix > 0 && (!bgr || !gray))) { set_error("bad argument"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_bgr_to_gray(bgr, npix, gray, sms, (cudaStream_t)stream);
}

int sl_pair_diff_counts(const uint8_t* frames, int n_frames, int64_t elems_per_frame, size_t frame_pitch,
                        int thresh, uint64_t* counts, void* stream) {
  if (n_frames < 1 || elems_per_frame < 0 || (n_frames > 1 && (!frames || !counts))) { set_error("bad argument"); return kErrBadArg; }
  if (frame_pitch < (size_t)elems_per_frame) { set_error("frame_pitch smaller than the frame"); return kErrBadArg; }
  if (n_frames > 65536) { set_error("too many frames"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_pair_diff(frames, n_frames, elems_per_frame, frame_pitch, thresh, (unsigned long long*)counts, sms,
                          (cudaStream_t)stream);
}

size_t sl_cloud_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  return cloud_workspace_bytes(B, H, W);
}

int sl_compact_offsets(const int16_t* h_idx, const int16_t* v_idx, int B, int H, int W,
                       void* workspace, int64_t* counts, void* stream) {
  if (!h_idx || !v_idx || !workspace || !counts) { set_error("null buffer"); return kErrBadArg; }
  if (B < 1 || H < 1 || W < 1) { set_error("B, H, W must be positive"); return kErrBadArg; }
  if ((uintptr_t)h_idx % 2 || (uintptr_t)v_idx % 2 || (uintptr_t)workspace % 256 || (uintptr_t)counts % 8) {
    set_error("h_idx / v_idx must be 2-byte aligned, workspace 256-byte, counts 8-byte"); return kErrBadArg;
  }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_compact_offsets(h_idx, v_idx, B, H, W, workspace, (long long*)counts, (cudaStream_t)stream);
}

int sl_compact_offsets_boxed(const int16_t* h_idx, const int16_t* v_idx, const SlCalib* calib, int precision,
                             const float* cam_rays, const float* proj_rays, double threshold, int B, int H, int W,
                             void* workspace, int64_t* counts, void* stream) {
  if (!h_idx || !v_idx || !calib || !workspace || !counts) { set_error("null buffer"); return kErrBadArg; }
  if (B < 1 || H < 1 || W < 1 || calib->proj_w < 1 || calib->proj_h < 1) { set_error("bad shape"); return kErrBadArg; }
  if ((cam_rays == nullptr) != (proj_rays == nullptr)) { set_error("cam_rays and proj_rays go together"); return kErrBadArg; }
  if (precision != SL_PREC_FP32 && precision != SL_PREC_FP64) { set_error("bad precision %d", precision); return kErrBadArg; }
  if (!(threshold == threshold)) { set_error("threshold is NaN"); return kErrBadArg; }
  if ((uintptr_t)h_idx % 2 || (uintptr_t)v_idx % 2 || (uintptr_t)workspace % 256 || (uintptr_t)counts % 8 ||
      (uintptr_t)cam_rays % 8 || (uintptr_t)proj_rays % 8) {
    set_error("h_idx / v_idx must be 2-byte aligned, workspace 256-byte, counts and the lens tables 8-byte"); return kErrBadArg;
  }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_compact_offsets_boxed(h_idx, v_idx, *calib, precision, cam_rays, proj_rays, threshold, B, H, W, workspace,
                                      (long long*)counts, (cudaStream_t)stream);
}

int sl_compact_offsets_decoded(int B, int H, int W, void* workspace, int64_t* counts, void* stream) {
  if (!workspace || !counts) { set_error("null buffer"); return kErrBadArg; }
  if (B < 1 || H < 1 || W < 1) { set_error("B, H, W must be positive"); return kErrBadArg; }
  if ((uintptr_t)workspace % 256 || (uintptr_t)counts % 8) { set_error("workspace must be 256-byte aligned, counts 8-byte"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_offsets_from_bitmap(B, H, W, workspace, (long long*)counts, (cudaStream_t)stream);
}

const uint32_t* sl_workspace_col_totals(const void* workspace, int B, int H, int W) {
  if (!workspace || B < 1 || H < 1 || W < 1) return nullptr;
  return carve_workspace(const_cast<void*>(workspace), B, H, W).col_tot;
}

int sl_emit_pairs(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb, int B,
                  int H, int W, int proj_w, int proj_h, const void* workspace, float* pairs,
                  uint32_t* rgb32, void* stream) {
  if (!h_idx || !v_idx || !workspace || !pairs) { set_error("null buffer"); return kErrBadArg; }
  if (rgb32 && !white_rgb) { set_error("rgb32 needs white_rgb"); return kErrBadArg; }
  if (B < 1 || H < 1 || W < 1 || proj_w < 1 || proj_h < 1) { set_error("bad shape"); return kErrBadArg; }
  if ((uintptr_t)pairs % 16 || (uintptr_t)rgb32 % 4 || (uintptr_t)h_idx % 2 || (uintptr_t)v_idx % 2) {
    set_error("pairs must be 16-byte aligned, rgb32 4-byte, h_idx / v_idx 2-byte"); return kErrBadArg;
  }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_emit_pairs(h_idx, v_idx, white_rgb, B, H, W, proj_w, proj_h, workspace, pairs, rgb32,
                           (cudaStream_t)stream);
}

int sl_lens_tables(const SlCalib* calib, int cam_w, int cam_h, int precision, float* cam_rays, float* proj_rays,
                   void* stream) {
  if (!calib || !cam_rays || !proj_rays) { set_error("null buffer"); return kErrBadArg; }
  if (precision != SL_PREC_FP32 && precision != SL_PREC_FP64) { set_error("bad precision %d", precision); return kErrBadArg; }
  if ((uintptr_t)cam_rays % 16 || (uintptr_t)proj_rays % 8) { set_error("cam_rays must be 16-byte aligned, proj_rays 8-byte"); return kErrBadArg; }
  if (cam_w < 1 || cam_h < 1 || calib->proj_w < 1 || calib->proj_h < 1) { set_error("bad shape"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_lens_tables(*calib, cam_w, cam_h, precision, cam_rays, proj_rays, (cudaStream_t)stream);
}

static int emit_cloud_checked(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb,
                             const SlCalib* calib, int precision, const float* cam_rays, const float* proj_rays,
                             int B, int H, int W, const void* workspace, float* const* clouds, int n_clouds,
                             int remote, int64_t max_points, void* stream) {
  if (!h_idx || !v_idx || !workspace || !clouds || !calib) { set_error("null buffer"); return kErrBadArg; }
  if (n_clouds < 1 || n_clouds > kMaxWorld) { set_error("n_clouds must be 1..%d", kMaxWorld); return kErrBadArg; }
  if ((cam_rays == nullptr) != (proj_rays == nullptr)) { set_error("cam_rays and proj_rays go together"); return kErrBadArg; }
  if (B < 1 || H < 1 || W < 1 || calib->proj_w < 1 || calib->proj_h < 1) { set_error("bad shape"); return kErrBadArg; }
  if (precision != SL_PREC_FP32 && precision != SL_PREC_FP64) { set_error("bad precision %d", precision); return kErrBadArg; }
  for (int k = 0; k < n_clouds; ++k)
    if (!clouds[k] || (uintptr_t)clouds[k] % 16) { set_error("cloud buffers must be non-null and 16-byte aligned"); return kErrBadArg; }
  if ((uintptr_t)cam_rays % 16 || (uintptr_t)proj_rays % 8 || (uintptr_t)h_idx % 2 || (uintptr_t)v_idx % 2) {
    set_error("cam_rays must be 16-byte aligned, proj_rays 8-byte, h_idx / v_idx 2-byte"); return kErrBadArg;
  }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_emit_cloud(h_idx, v_idx, white_rgb, *calib, precision, cam_rays, proj_rays, B, H, W, workspace,
                           clouds[0], (cudaStream_t)stream, clouds + 1, n_clouds - 1, remote, max_points);
}

int sl_emit_cloud(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb,
                  const SlCalib* calib, int precision, const float* cam_rays, const float* proj_rays,
                  int B, int H, int W, const void* workspace, float* cloud, void* stream) {
  float* one[1] = {cloud};
  return emit_cloud_checked(h_idx, v_idx, white_rgb, calib, precision, cam_rays, proj_rays, B, H, W, workspace,
                            one, 1, 0, 0, stream);
}

int sl_emit_cloud_multi(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb,
                        const SlCalib* calib, int precision, const float* cam_rays, const float* proj_rays,
                        int B, int H, int W, const void* workspace, float* const* clouds, int n_clouds,
                        int64_t max_points, void* stream) {
  if (max_points < 1) { set_error("max_points must be positive"); return kErrBadArg; }
  return emit_cloud_checked(h_idx, v_idx, white_rgb, calib, precision, cam_rays, proj_rays, B, H, W, workspace,
                            clouds, n_clouds, 1, max_points, stream);
}

int sl_transpose_table(const float* src, int rows, int cols, float* dst, void* stream) {
  if (!src || !dst || rows < 1 || cols < 1) { set_error("bad argument"); return kErrBadArg; }
  if ((uintptr_t)src % 8 || (uintptr_t)dst % 8) { set_error("tables must be 8-byte aligned"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_transpose_f2(src, rows, cols, dst, (cudaStream_t)stream);
}

int sl_emit_cloud_rows(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb, const SlCalib* calib,
                       const float* cam_rays_t, int cam_rows, const float* proj_bt_t, int B, int H, int W,
                       const void* workspace, float* cloud, void* stream) {
  if (!h_idx || !v_idx || !workspace || !cloud || !calib || !cam_rays_t || !proj_bt_t) { set_error("null buffer"); return kErrBadArg; }
  if (B < 1 || H < 1 || W < 1 || calib->proj_w < 1 || calib->proj_h < 1 || cam_rows < 1) { set_error("bad shape"); return kErrBadArg; }
  if ((uintptr_t)h_idx % 16 || (uintptr_t)v_idx % 16 || (uintptr_t)white_rgb % 16 || (uintptr_t)cloud % 16 ||
      (uintptr_t)cam_rays_t % 8 || (uintptr_t)proj_bt_t % 8) {
    set_error("h_idx, v_idx, white_rgb and cloud must be 16-byte aligned, the tables 8-byte"); return kErrBadArg;
  }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_emit_rows(h_idx, v_idx, white_rgb, *calib, cam_rays_t, cam_rows, proj_bt_t, B, H, W, workspace, cloud,
                          (cudaStream_t)stream);
}

int sl_emit_cloud_rows_multi(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb, const SlCalib* calib,
                             const float* cam_rays_t, int cam_rows, const float* proj_bt_t, int B, int H, int W,
                             const void* workspace, float* const* clouds, int n_clouds, void* stream) {
  if (!h_idx || !v_idx || !workspace || !clouds || !calib || !cam_rays_t || !proj_bt_t) { set_error("null buffer"); return kErrBadArg; }
  if (n_clouds < 1 || n_clouds > kMaxWorld) { set_error("n_clouds must be 1..%d", kMaxWorld); return kErrBadArg; }
  if (B < 1 || H < 1 || W < 1 || calib->proj_w < 1 || calib->proj_h < 1 || cam_rows < 1) { set_error("bad shape"); return kErrBadArg; }
  for (int k = 0; k < n_clouds; ++k)
    if (!clouds[k] || (uintptr_t)clouds[k] % 16) { set_error("cloud buffers must be non-null and 16-byte aligned"); return kErrBadArg; }
  if ((uintptr_t)h_idx % 16 || (uintptr_t)v_idx % 16 || (uintptr_t)white_rgb % 16 || (uintptr_t)cam_rays_t % 8 ||
      (uintptr_t)proj_bt_t % 8) {
    set_error("h_idx, v_idx and white_rgb must be 16-byte aligned, the tables 8-byte"); return kErrBadArg;
  }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_emit_rows(h_idx, v_idx, white_rgb, *calib, cam_rays_t, cam_rows, proj_bt_t, B, H, W, workspace, clouds[0],
                          (cudaStream_t)stream, clouds + 1, n_clouds - 1, 1);
}

int sl_emit_cloud_at(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb,
                     const SlCalib* calib, int precision, const float* cam_rays, const float* proj_rays,
                     int B, int H, int W, const void* workspace, float* cloud, int64_t max_points, void* stream) {
  if (max_points < 1) { set_error("max_points must be positive"); return kErrBadArg; }
  float* one[1] = {cloud};
  return emit_cloud_checked(h_idx, v_idx, white_rgb, calib, precision, cam_rays, proj_rays, B, H, W, workspace,
                            one, 1, 0, max_points, stream);
}

int sl_decode_triangulate_fused(const uint8_t* frames, int B, int F, int H, int W, size_t frame_pitch,
                                size_t scan_pitch, double eps, const uint8_t* white_rgb, const SlCalib* calib,
                                int precision, const float* cam_rays, const float* proj_rays, int16_t* h_idx,
                                int16_t* v_idx, void* workspace, float* cloud, int64_t* counts, int sm_reserve,
                                void* stream) {
  if (!workspace || !counts) { set_error("null buffer"); return kErrBadArg; }
  int rc = sl_decode_u8(frames, B, F, H, W, frame_pitch, scan_pitch, eps, h_idx, v_idx, nullptr, nullptr, workspace,
                        sm_reserve, stream);
  if (rc) return rc;
  if ((rc = sl_compact_offsets_decoded(B, H, W, workspace, counts, stream))) return rc;
  return sl_emit_cloud(h_idx, v_idx, white_rgb, calib, precision, cam_rays, proj_rays, B, H, W, workspace, cloud, stream);
}

int sl_triangulate_points(const float* cam_pts, const float* proj_pts, int64_t N,
                          const SlCalib* calib, int precision, double* out, void* stream) {
  if (N < 0 || !calib) { set_error("bad argument"); return kErrBadArg; }
  if (N > 0 && (!cam_pts || !proj_pts || !out)) { set_error("null buffer"); return kErrBadArg; }
  if (precision != SL_PREC_FP32 && precision != SL_PREC_FP64) { set_error("bad precision %d", precision); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_triangulate_points(cam_pts, proj_pts, N, *calib, precision, out, (cudaStream_t)stream);
}

size_t sl_filter_workspace_bytes(int64_t N) { return N < 0 ? 0 : filter_workspace_bytes(N); }

int sl_filter_box(const double* pts, const double* colors, int64_t N, double threshold,
                  double* out_pts, double* out_colors, int64_t* out_count, void* workspace,
                  size_t workspace_bytes, void* stream) {
  if (N < 0 || !out_count || !workspace) { set_error("bad argument"); return kErrBadArg; }
  if (N > 0 && (!pts || !out_pts)) { set_error("null buffer"); return kErrBadArg; }
  if ((colors == nullptr) != (out_colors == nullptr)) { set_error("colors and out_colors must be given together"); return kErrBadArg; }
  if (workspace_bytes < filter_workspace_bytes(N)) { set_error("workspace too small"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_filter_box(pts, colors, N, threshold, out_pts, out_colors, (long long*)out_count, workspace,
                           (cudaStream_t)stream);
}

int sl_filter_cloud_box(const float* cloud, int64_t N, double threshold, float* out_cloud,
                        int64_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  if (N < 0 || !out_count || !workspace) { set_error("bad argument"); return kErrBadArg; }
  if (N > 0 && (!cloud || !out_cloud)) { set_error("null buffer"); return kErrBadArg; }
  if (workspace_bytes < filter_workspace_bytes(N)) { set_error("workspace too small"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_filter_cloud(cloud, N, threshold, out_cloud, (long long*)out_count, workspace, (cudaStream_t)stream);
}

size_t sl_outlier_workspace_bytes(int64_t N) { return (N < 0 || N > (1LL << 30)) ? 0 : outlier_workspace_bytes(N); }

int sl_remove_statistical_outliers(const float* cloud, int64_t N, int nb_neighbors, double std_ratio,
                                   float* out_cloud, int64_t* out_index, double* mean_dist, double* stats,
                                   int64_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  if (N < 0 || N > (1LL << 30)) { set_error("N out of range (0..2^30)"); return kErrBadArg; }
  // Open3D: "the number of neighbors and standard deviation ratio must be positive"
  if (nb_neighbors < 1 || !(std_ratio > 0.0)) { set_error("nb_neighbors and std_ratio must be positive"); return kErrBadArg; }
  if (nb_neighbors > 128) { set_error("nb_neighbors up to 128 supported (got %d)", nb_neighbors); return kErrBadArg; }
  if (!out_count || !workspace) { set_error("null buffer"); return kErrBadArg; }
  if (N > 0 && !cloud) { set_error("null buffer"); return kErrBadArg; }
  if (workspace_bytes < outlier_workspace_bytes(N)) { set_error("workspace too small"); return kErrBadArg; }
  int sms = 0, rc;
  if ((rc = device_gate(&sms))) return rc;
  return launch_outlier_removal(cloud, N, nb_neighbors, std_ratio, out_cloud, (long long*)out_index, mean_dist, stats,
                                (long long*)out_count, workspace, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ host-buffer pipeline
struct SlPipeline {
  int device = 0, max_batch = 0, F = 0, H = 0, W = 0, sms = 0;
  int sm_reserve = 0;   // sl_pipeline_set_sm_reserve
  static constexpr int kSlots = 3;
  struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t counted = nullptr;
    uint8_t* frames = nullptr;   // [F][H][W]
    uint8_t* frames_bgr = nullptr;   // [F][H][W][3], allocated by the first sl_pipeline_submit_bgr on the slot
    uint8_t* white = nullptr;    // [H][W][3]
    int16_t* h = nullptr;
    int16_t* v = nullptr;
    void* ws = nullptr;
    long long* counts = nullptr;       // device [2]
    long long* counts_pinned = nullptr; // host pinned [2]
    float* cloud = nullptr;      // [H*W][4] packed records
    // streaming API: the scan in flight on this slot (sl_pipeline_submit) and where its results go
    bool pending = false;
    float* user_cloud = nullptr;
    long long user_cap = 0;
    int64_t* user_count = nullptr;
  } slot[kSlots];
  long long submitted = 0;       // scans handed to sl_pipeline_submit so far
  bool stream_overflow = false;  // a submitted scan did not fit its cloud buffer (reported by sl_pipeline_flush)
  int8_t* h_codes = nullptr;   // lazily allocated for sl_decode_host with codes
  int8_t* v_codes = nullptr;
  // lens tables of the last calibration seen (rebuilt when the calibration changes)
  SlCalib table_calib;
  int table_precision = -1;
  bool have_tables = false;
  float* cam_rays = nullptr;
  float* proj_rays = nullptr;
  size_t proj_rays_elems = 0;
  // the same tables stored column-major for the row-mapped emission (fp32, W % 32 == 0)
  float* cam_rays_t = nullptr;
  float* proj_rays_t = nullptr;
  size_t proj_rays_t_elems = 0;
  bool have_transposed = false;
};

// the emission of one pipeline step: row-mapped where it applies, else the general one
static int pipeline_emit(SlPipeline* p, const int16_t* h, const int16_t* v, const uint8_t* white, const SlCalib& calib,
                         int precision, int B, const void* ws, float* cloud, cudaStream_t st) {
  if (p->have_transposed && precision == SL_PREC_FP32)
    return launch_emit_rows(h, v, white, calib, p->cam_rays_t, p->H, p->proj_rays_t, B, p->H, p->W, ws, cloud, st);
  return launch_emit_cloud(h, v, white, calib, precision, p->cam_rays, p->proj_rays, B, p->H, p->W, ws, cloud, st);
}

static void pipeline_free(SlPipeline* p) {
  if (!p) return;
  DeviceGuard guard;
  cudaSetDevice(p->device);
  for (auto& s : p->slot) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.frames); cudaFree(s.frames_bgr); cudaFree(s.white); cudaFree(s.h); cudaFree(s.v); cudaFree(s.ws);
    cudaFree(s.counts); cudaFreeHost(s.counts_pinned); cudaFree(s.cloud);
    if (s.counted) cudaEventDestroy(s.counted);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  cudaFree(p->h_codes); cudaFree(p->v_codes);
  cudaFree(p->cam_rays); cudaFree(p->proj_rays); cudaFree(p->cam_rays_t); cudaFree(p->proj_rays_t);
  delete p;
}

int sl_pipeline_create(SlPipeline** out, int device, int max_batch, int F, int H, int W) {
  if (!out) { set_error("out is null"); return kErrBadArg; }
  *out = nullptr;
  int rc = check_decode_shape(max_batch < 1 ? 0 : 1, F, H, W);
  if (rc) return rc;
  if (max_batch < 1) { set_error("max_batch must be >= 1"); return kErrBadArg; }
  DeviceGuard guard;
  if (device >= 0) SL_CUDA(cudaSetDevice(device));
  int sms = 0;
  if ((rc = device_gate(&sms))) return rc;
  SlPipeline* p = new (std::nothrow) SlPipeline();
  if (!p) { set_error("out of host memory"); return kErrAlloc; }
  cudaGetDevice(&p->device);
  p->max_batch = max_batch; p->F = F; p->H = H; p->W = W; p->sms = sms;
  const size_t npix = (size_t)H * W;
  cudaError_t e = cudaSuccess;
  for (auto& s : p->slot) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.counted, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&s.frames, (size_t)F * npix);
    if (e == cudaSuccess) e = cudaMalloc(&s.white, npix * 3);
    if (e == cudaSuccess) e = cudaMalloc(&s.h, npix * 2);
    if (e == cudaSuccess) e = cudaMalloc(&s.v, npix * 2);
    if (e == cudaSuccess) e = cudaMalloc(&s.ws, cloud_workspace_bytes(1, H, W));
    if (e == cudaSuccess) e = cudaMalloc(&s.counts, 2 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMallocHost(&s.counts_pinned, 2 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&s.cloud, npix * 4 * sizeof(float));
  }
  if (e != cudaSuccess) {
    pipeline_free(p);
    set_error("pipeline allocation failed: %s", cudaGetErrorString(e));
    return kErrAlloc;
  }
  *out = p;
  return kOk;
}

void sl_pipeline_destroy(SlPipeline* p) { pipeline_free(p); }

int sl_pipeline_set_sm_reserve(SlPipeline* p, int sms) {
  if (!p || sms < 0) { set_error("bad argument"); return kErrBadArg; }
  p->sm_reserve = sms;
  return kOk;
}

// lens tables of the pipeline: (re)built when the calibration or the precision changes; all slots must be idle then
static int pipeline_tables(SlPipeline* p, const SlCalib* calib, int precision) {
  const size_t npix = (size_t)p->H * p->W;
  if (calib->proj_w < 1 || calib->proj_h < 1) { set_error("bad projector size"); return kErrBadArg; }
  if (p->have_tables && p->table_precision == precision && memcmp(&p->table_calib, calib, sizeof(SlCalib)) == 0) return kOk;
  for (auto& s : p->slot) SL_CUDA(cudaStreamSynchronize(s.stream));
  const size_t pe = (size_t)calib->proj_w * calib->proj_h * 2;
  if (!p->cam_rays) SL_CUDA(cudaMalloc(&p->cam_rays, npix * 2 * sizeof(float)));
  if (pe > p->proj_rays_elems) {
    cudaFree(p->proj_rays); p->proj_rays = nullptr; p->proj_rays_elems = 0;
    SL_CUDA(cudaMalloc(&p->proj_rays, pe * sizeof(float)));
    p->proj_rays_elems = pe;
  }
  const int trc = launch_lens_tables(*calib, p->W, p->H, precision, p->cam_rays, p->proj_rays, p->slot[0].stream);
  if (trc) return trc;
  p->have_transposed = false;
  if (precision == SL_PREC_FP32 && p->W % 32 == 0) {
    if (!p->cam_rays_t) SL_CUDA(cudaMalloc(&p->cam_rays_t, npix * 2 * sizeof(float)));
    if (pe > p->proj_rays_t_elems) {
      cudaFree(p->proj_rays_t); p->proj_rays_t = nullptr; p->proj_rays_t_elems = 0;
      SL_CUDA(cudaMalloc(&p->proj_rays_t, pe * sizeof(float)));
      p->proj_rays_t_elems = pe;
    }
    int xrc = launch_transpose_f2(p->cam_rays, p->H, p->W, p->cam_rays_t, p->slot[0].stream);
    if (!xrc) xrc = launch_transpose_f2(p->proj_rays, calib->proj_h, calib->proj_w, p->proj_rays_t, p->slot[0].stream);
    if (xrc) return xrc;
    p->have_transposed = true;
  }
  SL_CUDA(cudaStreamSynchronize(p->slot[0].stream));
  p->table_calib = *calib;
  p->table_precision = precision;
  p->have_tables = true;
  return kOk;
}

// Scan i goes through slot i % kSlots: H2D -> decode -> offsets -> emit -> D2H(count).  The cloud
// D2H needs the count (packed host output), so it is issued one step later, after the slot's
// `counted` event, while the next scans' copies and kernels are already queued on other streams.
int sl_scan_to_cloud_host(SlPipeline* p, const uint8_t* frames_host, const uint8_t* white_rgb_host,
                          int B, const SlCalib* calib, int precision, double eps,
                          int16_t* h_idx_host, int16_t* v_idx_host, float* cloud_host,
                          int64_t cap_points, int64_t* counts_host) {
  if (!p || !frames_host || !calib || !cloud_host || !counts_host) { set_error("null argument"); return kErrBadArg; }
  if (B < 1 || B > p->max_batch) { set_error("B=%d outside 1..max_batch=%d", B, p->max_batch); return kErrBadArg; }
  if ((h_idx_host == nullptr) != (v_idx_host == nullptr)) { set_error("h_idx_host and v_idx_host go together"); return kErrBadArg; }
  if (precision != SL_PREC_FP32 && precision != SL_PREC_FP64) { set_error("bad precision %d", precision); return kErrBadArg; }
  for (auto& s : p->slot)
    if (s.pending) { set_error("scans submitted with sl_pipeline_submit are still in flight: call sl_pipeline_flush first"); return kErrState; }
  DeviceGuard guard;
  SL_CUDA(cudaSetDevice(p->device));
  const size_t npix = (size_t)p->H * p->W;
  const size_t stack = (size_t)p->F * npix;
  const int S = SlPipeline::kSlots;
  {
    const int trc = pipeline_tables(p, calib, precision);
    if (trc) return trc;
  }
  long long total = 0;
  bool overflow = false;
  int rc = kOk;

  auto drain = [&](int i) -> int {   // finish scan i: wait for its count, copy its cloud out
    auto& s = p->slot[i % S];
    SL_CUDA(cudaEventSynchronize(s.counted));
    const long long n = s.counts_pinned[0];
    counts_host[i] = n;
    if (total + n > cap_points) overflow = true;
    if (!overflow && n > 0)
      SL_CUDA(cudaMemcpyAsync(cloud_host + (size_t)total * 4, s.cloud, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    total += n;
    return kOk;
  };

  for (int i = 0; i < B && rc == kOk; ++i) {
    if (i >= S) {                 // the slot is reused: scan i-S must be on its way out first.
      // drain() queues that scan's D2H on the slot's stream, so stream order protects the
      // slot's buffers from the copies queued next.
      if ((rc = drain(i - S)) != kOk) break;
    }
    auto& s = p->slot[i % S];
    SL_CUDA(cudaMemcpyAsync(s.frames, frames_host + (size_t)i * stack, stack, cudaMemcpyHostToDevice, s.stream));
    if (white_rgb_host)
      SL_CUDA(cudaMemcpyAsync(s.white, white_rgb_host + (size_t)i * npix * 3, npix * 3, cudaMemcpyHostToDevice, s.stream));
    long long bm_words = 0;
    unsigned long long* flags = nullptr;
    int n_flags = 0;
    uint32_t* bm = workspace_bitmap(s.ws, 1, p->H, p->W, &bm_words, &flags, &n_flags);
    rc = launch_decode(s.frames, 1, p->F, (long long)npix, npix, stack, eps, s.h, s.v, nullptr, nullptr, bm, bm_words,
                       flags, n_flags, p->sms, s.stream, false, p->sm_reserve);
    if (rc) break;
    if (h_idx_host) {
      SL_CUDA(cudaMemcpyAsync(h_idx_host + (size_t)i * npix, s.h, npix * 2, cudaMemcpyDeviceToHost, s.stream));
      SL_CUDA(cudaMemcpyAsync(v_idx_host + (size_t)i * npix, s.v, npix * 2, cudaMemcpyDeviceToHost, s.stream));
    }
    rc = launch_offsets_from_bitmap(1, p->H, p->W, s.ws, s.counts, s.stream);
    if (rc) break;
    SL_CUDA(cudaMemcpyAsync(s.counts_pinned, s.counts, 2 * sizeof(long long), cudaMemcpyDeviceToHost, s.stream));
    rc = pipeline_emit(p, s.h, s.v, white_rgb_host ? s.white : nullptr, *calib, precision, 1, s.ws, s.cloud, s.stream);
    if (rc) break;
    SL_CUDA(cudaEventRecord(s.counted, s.stream));
  }
  if (rc == kOk) {
    for (int i = (B > S ? B - S : 0); i < B; ++i)
      if ((rc = drain(i)) != kOk) break;
  }
  for (auto& s : p->slot) {
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess && rc == kOk) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  if (rc != kOk) return rc;
  counts_host[B] = total;
  if (overflow) { set_error("cloud of %lld points exceeds cap_points=%lld", total, (long long)cap_points); return kErrBadArg; }
  return kOk;
}

// ---- streaming: scans are submitted one at a time and complete in the background ----------------------------
// the scan occupying `s` is on its way out: its count is read, its cloud copy queued on the slot's stream
static int pipeline_retire(SlPipeline* p, SlPipeline::Slot& s) {
  if (!s.pending) return kOk;
  SL_CUDA(cudaEventSynchronize(s.counted));
  const long long n = s.counts_pinned[0];
  *s.user_count = n;
  if (n > s.user_cap) p->stream_overflow = true;
  else if (n > 0)
    SL_CUDA(cudaMemcpyAsync(s.user_cloud, s.cloud, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
  s.pending = false;
  return kOk;
}

static int pipeline_submit(SlPipeline* p, const uint8_t* frames_host, bool bgr, const uint8_t* white_rgb_host,
                           const SlCalib* calib, int precision, double eps, float* cloud_host, int64_t cap_points,
                           int64_t* count_host) {
  if (!p || !frames_host || !calib || !cloud_host || !count_host) { set_error("null argument"); return kErrBadArg; }
  if (precision != SL_PREC_FP32 && precision != SL_PREC_FP64) { set_error("bad precision %d", precision); return kErrBadArg; }
  DeviceGuard guard;
  SL_CUDA(cudaSetDevice(p->device));
  int rc = pipeline_tables(p, calib, precision);
  if (rc) return rc;
  const size_t npix = (size_t)p->H * p->W;
  const size_t frame = bgr ? npix * 3 : npix;
  const size_t stack = (size_t)p->F * frame;
  auto& s = p->slot[p->submitted % SlPipeline::kSlots];
  if ((rc = pipeline_retire(p, s))) return rc;        // the scan submitted kSlots calls ago leaves the slot first
  if (bgr && !s.frames_bgr) SL_CUDA(cudaMalloc(&s.frames_bgr, stack));     // first colour scan on this slot
  uint8_t* dev_frames = bgr ? s.frames_bgr : s.frames;
  SL_CUDA(cudaMemcpyAsync(dev_frames, frames_host, stack, cudaMemcpyHostToDevice, s.stream));
  if (white_rgb_host) SL_CUDA(cudaMemcpyAsync(s.white, white_rgb_host, npix * 3, cudaMemcpyHostToDevice, s.stream));
  long long bm_words = 0;
  unsigned long long* flags = nullptr;
  int n_flags = 0;
  uint32_t* bm = workspace_bitmap(s.ws, 1, p->H, p->W, &bm_words, &flags, &n_flags);
  if ((rc = launch_decode(dev_frames, 1, p->F, (long long)npix, frame, stack, eps, s.h, s.v, nullptr, nullptr, bm, bm_words,
                          flags, n_flags, p->sms, s.stream, bgr, p->sm_reserve))) return rc;
  if ((rc = launch_offsets_from_bitmap(1, p->H, p->W, s.ws, s.counts, s.stream))) return rc;
  SL_CUDA(cudaMemcpyAsync(s.counts_pinned, s.counts, 2 * sizeof(long long), cudaMemcpyDeviceToHost, s.stream));
  if ((rc = pipeline_emit(p, s.h, s.v, white_rgb_host ? s.white : nullptr, *calib, precision, 1, s.ws, s.cloud, s.stream))) return rc;
  SL_CUDA(cudaEventRecord(s.counted, s.stream));
  s.pending = true;
  s.user_cloud = cloud_host; s.user_cap = cap_points; s.user_count = count_host;
  ++p->submitted;
  return kOk;
}

int sl_pipeline_submit(SlPipeline* p, const uint8_t* frames_host, const uint8_t* white_rgb_host, const SlCalib* calib,
                       int precision, double eps, float* cloud_host, int64_t cap_points, int64_t* count_host) {
  return pipeline_submit(p, frames_host, false, white_rgb_host, calib, precision, eps, cloud_host, cap_points, count_host);
}

int sl_pipeline_submit_bgr(SlPipeline* p, const uint8_t* frames_bgr_host, const uint8_t* white_rgb_host,
                           const SlCalib* calib, int precision, double eps, float* cloud_host, int64_t cap_points,
                           int64_t* count_host) {
  return pipeline_submit(p, frames_bgr_host, true, white_rgb_host, calib, precision, eps, cloud_host, cap_points, count_host);
}

int sl_pipeline_flush(SlPipeline* p) {
  if (!p) { set_error("null argument"); return kErrBadArg; }
  DeviceGuard guard;
  SL_CUDA(cudaSetDevice(p->device));
  int rc = kOk;
  for (int k = 0; k < SlPipeline::kSlots && rc == kOk; ++k)      // oldest first
    rc = pipeline_retire(p, p->slot[(p->submitted + k) % SlPipeline::kSlots]);
  for (auto& s : p->slot) {
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess && rc == kOk) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  if (rc != kOk) return rc;
  if (p->stream_overflow) {
    p->stream_overflow = false;
    set_error("a submitted scan produced more points than its cap_points (its count is written, its cloud is not)");
    return kErrBadArg;
  }
  return kOk;
}

int sl_decode_host(SlPipeline* p, const uint8_t* frames_host, int B, double eps,
                   int16_t* h_idx_host, int16_t* v_idx_host, int8_t* h_codes_host,
                   int8_t* v_codes_host) {
  if (!p || !frames_host || !h_idx_host || !v_idx_host) { set_error("null argument"); return kErrBadArg; }
  if (B < 1 || B > p->max_batch) { set_error("B=%d outside 1..max_batch=%d", B, p->max_batch); return kErrBadArg; }
  if ((h_codes_host == nullptr) != (v_codes_host == nullptr)) { set_error("codes buffers go together"); return kErrBadArg; }
  DeviceGuard guard;
  SL_CUDA(cudaSetDevice(p->device));
  const size_t npix = (size_t)p->H * p->W;
  const size_t stack = (size_t)p->F * npix;
  const int n = (p->F - 2) / 4;
  if (h_codes_host && !p->h_codes) {
    SL_CUDA(cudaMalloc(&p->h_codes, (size_t)n * npix));
    SL_CUDA(cudaMalloc(&p->v_codes, (size_t)n * npix));
  }
  const int S = SlPipeline::kSlots;
  int rc = kOk;
  for (int i = 0; i < B; ++i) {
    auto& s = p->slot[i % S];
    cudaStream_t st = s.stream;
    if (h_codes_host) {   // the code planes are single-buffered: serialise on them
      st = p->slot[0].stream;
    }
    auto& u = h_codes_host ? p->slot[0] : s;
    SL_CUDA(cudaMemcpyAsync(u.frames, frames_host + (size_t)i * stack, stack, cudaMemcpyHostToDevice, st));
    rc = launch_decode(u.frames, 1, p->F, (long long)npix, npix, stack, eps, u.h, u.v,
                       h_codes_host ? p->h_codes : nullptr, h_codes_host ? p->v_codes : nullptr, nullptr, 0, nullptr, 0,
                       p->sms, st, false, p->sm_reserve);
    if (rc) break;
    SL_CUDA(cudaMemcpyAsync(h_idx_host + (size_t)i * npix, u.h, npix * 2, cudaMemcpyDeviceToHost, st));
    SL_CUDA(cudaMemcpyAsync(v_idx_host + (size_t)i * npix, u.v, npix * 2, cudaMemcpyDeviceToHost, st));
    if (h_codes_host) {
      SL_CUDA(cudaMemcpyAsync(h_codes_host + (size_t)i * n * npix, p->h_codes, (size_t)n * npix, cudaMemcpyDeviceToHost, st));
      SL_CUDA(cudaMemcpyAsync(v_codes_host + (size_t)i * n * npix, p->v_codes, (size_t)n * npix, cudaMemcpyDeviceToHost, st));
    }
  }
  for (auto& s : p->slot) {
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess && rc == kOk) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  return rc;
}

}  // extern "C"

/* slb200 -- C ABI of the B200 (sm_100a) Gray-code decode + triangulation hot path.
 *
 * The reference (guillaume-charron/3DScanner-GrayCode) is pure Python and has no FFI of
 * its own; its boundary for this path is the module surface that src/3-capture_decode.py:5-8
 * and src/4-triangulate.py:5-6 import.  Each entry point below names the reference
 * function(s) (file:line, relative to the reference repo) it replaces; INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / CUDA types in the signatures
 *    (`stream` is a cudaStream_t passed as void*, NULL = default stream);
 *  - "device-pointer" entry points are asynchronous on `stream`, the caller owns every
 *    buffer; "host" entry points (suffix _host) take HOST buffers, do their own staged
 *    H2D/D2H copies and return when the results are in the caller's memory;
 *  - return 0 on success or a negative SL_ERR_* code; sl_last_error() gives the message
 *    (thread-local);
 *  - sm_100 only: every entry point fails with SL_ERR_DEVICE on another GPU and there is
 *    no CPU fallback.
 */
#ifndef SLB200_H_
#define SLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SL_OK 0
#define SL_ERR_BAD_ARG (-1)
#define SL_ERR_CUDA (-2)
#define SL_ERR_DEVICE (-3)
#define SL_ERR_ALLOC (-4)
#define SL_ERR_STATE (-5)

/* precision of the triangulation arithmetic */
#define SL_PREC_FP32 0 /* fp32 undistort + cross-product law of sines (<=1e-4 rel. of the reference) */
#define SL_PREC_FP64 1 /* mirrors the reference dtype flow: fp64 undistort -> f32 rays -> fp64 angles */

/* Calibration, converted once by the caller.  Matrices row-major.
 * proj_K must already be scaled to the panel size (scanner/triangulation/triangulate.py:28-33). */
typedef struct SlCalib {
  double cam_K[9];
  double cam_dist[8];  /* k1 k2 p1 p2 k3 k4 k5 k6, unused = 0 */
  double proj_K[9];
  double proj_dist[8];
  double R[9];         /* proj_R: camera -> projector rotation (triangulate.py:84, R=) */
  double T[3];         /* proj_T[:,0] (triangulate.py:86) */
  int32_t proj_w;      /* clamp of decoded indices (triangulate.py:60-61) */
  int32_t proj_h;
  int32_t cam_row0;    /* row-band sharding: image row of the first row of the h/v/white buffers passed to
                          sl_emit_cloud (camera coordinates and cam_rays look-ups use cam_row0 + y); else 0 */
  int32_t reserved;
} SlCalib;

const char* sl_last_error(void);
int sl_version(void);
/* 0 if `device` (or the current device when < 0) is an sm_100 GPU */
int sl_device_check(int device);

/* ---------------------------------------------------------------- decode (stage 1) */

/* Fused decode of B frame stacks.
 * Replaces get_codes -> get_direct_indirect + get_is_lit (scanner/grayCode/decode_codes.py:231-248,
 * 90-122, 125-186) and the two per-pixel gray_to_decimal loops (src/3-capture_decode.py:99-100,
 * decode_codes.py:189-229).
 *   frames      device u8; frame f of scan b starts at frames + b*scan_pitch + f*frame_pitch,
 *               H*W contiguous pixels per frame; F = 4n+2, 3 <= n <= 15, frame 0 black, 1 white
 *   eps         the reference's `eps` (decode_codes.py:125); integer values 0..127 take the
 *               byte-compare kernel, anything else the fp64-compare kernel (same results)
 *   h_idx,v_idx device int16 [B][H*W]: projector column / row, -1 = invalid
 *   h_codes,v_codes  nullable device int8 [B][n][H*W]: the tri-state arrays get_codes returns
 *   cloud_workspace  nullable device scratch of sl_cloud_workspace_bytes(B,H,W): when given, the
 *               kernel also leaves the per-pixel validity bitmap there, so the compaction can start
 *               from it (sl_compact_offsets_decoded) instead of re-reading the indices
 *   sm_reserve  launch option, normally 0: the persistent decode kernel runs one CTA per SM; `sm_reserve` SMs are
 *               left out of that grid.  Worth setting (8 on a B200) by callers that pipeline independent scans over
 *               several CUDA streams: the short offsets / emit kernels of the other scans then find SMs of their own
 *               (bench step 34.2 -> 32.4 us per 1080p scan); a decode launch running alone gets ~5 % slower.
 * Alignment: any; the TMA kernel needs frames, frame_pitch, scan_pitch multiples of 16 bytes and h_idx / v_idx
 * multiples of 8, anything else is decoded by the (slower) generic kernel with the same results.
 * The library keeps no mutable process state: every option is an argument, per-device constants are initialised once
 * under a lock by the first call on that device, and calls from several host threads / streams / devices may overlap.
 */
int sl_decode_u8(const uint8_t* frames, int B, int F, int H, int W, size_t frame_pitch,
                 size_t scan_pitch, double eps, int16_t* h_idx, int16_t* v_idx, int8_t* h_codes,
                 int8_t* v_codes, void* cloud_workspace, int sm_reserve, void* stream);

/* sl_decode_u8 plus a diagnostic: stats (device u64 [2], zeroed by the call) receives the number of 4-pixel groups
 * the byte-compare kernel decoded and how many of them needed the full evaluation of the four rules
 * (decode_codes.py:172-182) instead of the single-threshold fast pass -- the kernel's input-dependent cost. */
int sl_decode_u8_stats(const uint8_t* frames, int B, int F, int H, int W, size_t frame_pitch, size_t scan_pitch,
                       double eps, int16_t* h_idx, int16_t* v_idx, uint64_t* stats, void* stream);

/* The same decode straight from the camera's 8-bit BGR frames: the capture loop's per-frame
 * cv2.cvtColor(frame, COLOR_BGR2GRAY) (src/3-capture_decode.py:66; gray = (3735 B + 19235 G + 9798 R + 16384) >> 15)
 * is applied while the frames stream through shared memory, so no gray stack is ever written or re-read.
 *   frames_bgr  device u8; pixel p of frame f of scan b at frames_bgr + b*scan_pitch + f*frame_pitch + 3*p
 *               (B, G, R bytes); pitches in bytes, frame_pitch >= 3*H*W
 * Other arguments as sl_decode_u8 (the tri-state code planes are not offered here). */
int sl_decode_bgr_u8(const uint8_t* frames_bgr, int B, int F, int H, int W, size_t frame_pitch,
                     size_t scan_pitch, double eps, int16_t* h_idx, int16_t* v_idx,
                     void* cloud_workspace, int sm_reserve, void* stream);

/* get_direct_indirect (decode_codes.py:90-122): L_d, L_g device f64 [H*W] (NaN where
 * white+black == 0, as in the reference). */
int sl_direct_indirect(const uint8_t* frames, int F, int H, int W, size_t frame_pitch, double* L_d,
                       double* L_g, void* stream);

/* get_is_lit (decode_codes.py:125-186) with caller-supplied L_d / L_g (device f64 [H*W]):
 * fp64 compares exactly as written in the reference; `m` has no effect there and is not taken. */
int sl_is_lit(const uint8_t* frames, int F, int H, int W, size_t frame_pitch, const double* L_d,
              const double* L_g, double eps, int8_t* h_codes, int8_t* v_codes, void* stream);

/* Per-bit max over R capture runs (src/3-capture_decode.py:95-96) followed by the per-pixel
 * Gray -> decimal conversion (src/3-capture_decode.py:99-100).
 *   h_codes,v_codes  device int8, run r at + r*run_pitch, [n][npix] each (values -1/0/1)
 */
int sl_codes_to_pixels(const int8_t* h_codes, const int8_t* v_codes, int R, int n, int64_t npix,
                       size_t run_pitch, int16_t* h_idx, int16_t* v_idx, void* stream);

/* ------------------------------------------------------------- capture front end */

/* get_image_sequence(get_gray_codes(w, h), w, h) (scanner/grayCode/generate_codes.py:5-81) generated in device
 * memory: out device u8 [4n+2][proj_h][proj_w], n = ceil(log2(max(proj_w, proj_h))). */
int sl_pattern_sequence_u8(int proj_w, int proj_h, uint8_t* out, void* stream);

/* cv2.cvtColor(frame, COLOR_BGR2GRAY) on 8-bit frames, as src/3-capture_decode.py:66 and to_gray
 * (scanner/grayCode/decode_codes.py:70-87) apply it before the decode:
 * gray = (3735*B + 19235*G + 9798*R + 16384) >> 15.  bgr device u8 [npix][3], gray device u8 [npix]. */
int sl_bgr_to_gray_u8(const uint8_t* bgr, int64_t npix, uint8_t* gray, void* stream);

/* remove_bad_images' measurements (decode_codes.py:49-54): counts[k] = number of elements with
 * |frames[k+1] - frames[k]| > thresh, k = 0..n_frames-2 (the reference counts per channel element).
 * frames device u8, frame k at frames + k*frame_pitch; counts device u64 [n_frames-1]. */
int sl_pair_diff_counts(const uint8_t* frames, int n_frames, int64_t elems_per_frame, size_t frame_pitch,
                        int thresh, uint64_t* counts, void* stream);

/* ---------------------------------------------------------- triangulation (stage 2) */

/* Bytes of device scratch the compaction needs for B scans of H x W (the workspace must be 256-byte aligned). */
size_t sl_cloud_workspace_bytes(int B, int H, int W);

/* Ordered-compaction offsets: counts the valid pixels (h != -1 and v != -1) and prepares
 * `workspace` so the emit kernels write every scan in the reference's x-major order
 * (triangulate.py:52-56).  counts: device int64 [B+1]: counts[b] = points of scan b,
 * counts[B] = total.  Points of scan b start at sum(counts[0..b-1]) in the packed outputs. */
int sl_compact_offsets(const int16_t* h_idx, const int16_t* v_idx, int B, int H, int W,
                       void* workspace, int64_t* counts, void* stream);

/* The same offsets with Triangulate.filter_3d_pts (triangulate.py:99-122, as src/4-triangulate.py:71 applies it with
 * threshold 0.5) folded in as one more predicate: a pixel counts only if it decoded (triangulate.py:56) AND the point the
 * emission will produce for it has |x|, |y|, |z| < threshold (strict, compared as doubles on the float record, exactly
 * like sl_filter_cloud_box).  sl_emit_cloud / sl_emit_pairs on this workspace then emit the filtered cloud directly, in
 * the reference's order; the unfiltered cloud never exists.  calib / precision / cam_rays / proj_rays as in sl_emit_cloud
 * (the emission must be called with the same ones: the predicate is evaluated on the very point it will write). */
int sl_compact_offsets_boxed(const int16_t* h_idx, const int16_t* v_idx, const SlCalib* calib, int precision,
                             const float* cam_rays, const float* proj_rays, double threshold, int B, int H, int W,
                             void* workspace, int64_t* counts, void* stream);

/* Same, for a workspace whose validity bitmap was already written by sl_decode_u8 (same B,H,W). */
int sl_compact_offsets_decoded(int B, int H, int W, void* workspace, int64_t* counts, void* stream);

/* Per-column point counts left in the workspace by sl_compact_offsets[_decoded]: device u32 [B][W]
 * (needed to merge the clouds of row bands back into x-major order).  Returns NULL on bad arguments. */
const uint32_t* sl_workspace_col_totals(const void* workspace, int B, int H, int W);

/* Triangulate.get_cam_proj_pts (triangulate.py:39-71): emits, x-major, one 16-byte record per
 * valid pixel: pairs (N,4) f32 = [x, y, min(proj_w-1,h), min(proj_h-1,v)] (cam_pts | proj_pts) and,
 * when rgb32 and white_rgb are given, rgb32 (N,) u32 = R | G<<8 | B<<16 of white_rgb[y,x,:].
 * Requires sl_compact_offsets on the same workspace first. */
int sl_emit_pairs(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb, int B,
                  int H, int W, int proj_w, int proj_h, const void* workspace, float* pairs,
                  uint32_t* rgb32, void* stream);

/* Lens tables: the undistorted normalised ray of every camera pixel (rotated by R) and of every
 * projector pixel -- cv2.undistortPoints (triangulate.py:84-85: 5 Brown-Conrady fixed-point
 * iterations, fp64 arithmetic, f32 result) evaluated once per calibration on the two pixel
 * lattices, which never change between scans.  cam_rays device f32 [cam_h][cam_w][2] (16-byte aligned);
 * proj_table device f32 [proj_h][proj_w][2] (8-byte aligned), its content depends on `precision`:
 *   SL_PREC_FP64  the exact ray (x, y) of the panel pixel -- the fp64 mode mirrors the reference's dtype flow from it;
 *   SL_PREC_FP32  (|T x p|, T.p) of that exact ray p = (x, y, 1): all the fp32 point formula needs from the projector
 *                 side, so the per-point arithmetic is 12 instructions.
 * sl_emit_cloud must be given tables built with the precision it is called with. */
int sl_lens_tables(const SlCalib* calib, int cam_w, int cam_h, int precision, float* cam_rays, float* proj_table,
                   void* stream);

/* Fused get_cam_proj_pts + triangulate (triangulate.py:39-97): undistortion of both sides,
 * law-of-sines depth, ordered compaction.  With cam_rays/proj_rays (from sl_lens_tables, same
 * calibration, sizes and precision) the rays are looked up; with both NULL the 5 Brown-Conrady iterations
 * run in the kernel for every valid pixel (fp32 or fp64 per `precision`).
 * cloud (N,4) f32: one 16-byte XYZRGB record per point = [x, y, z, bits(R | G<<8 | B<<16)] (the
 * layout PCL calls PointXYZRGB; colour word 0 when white_rgb is NULL), packed over the batch, each
 * scan in the reference's x-major order.
 * Requires sl_compact_offsets[_decoded] on the same workspace first.
 * Alignment: cloud 16 bytes, cam_rays 16, proj_rays 8 (SL_ERR_BAD_ARG otherwise); h_idx / v_idx 8 bytes and white_rgb 4
 * for the vector loads -- buffers with less (any 2-byte aligned sub-view) take the scalar path, same results. */
int sl_emit_cloud(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb,
                  const SlCalib* calib, int precision, const float* cam_rays, const float* proj_rays,
                  int B, int H, int W, const void* workspace, float* cloud, void* stream);

/* The whole device-resident path in one call: sl_decode_u8 (validity bitmap into `workspace`) ->
 * sl_compact_offsets_decoded -> sl_emit_cloud, queued back to back on `stream`
 * (src/3-capture_decode.py:75-100 + src/4-triangulate.py:50-64).  Arguments as in those three entry points;
 * h_idx / v_idx are outputs the caller may also use.  "Fused" is about the interface: ordered compaction needs
 * every column's point count before the first record can be placed, so the decode and the emission stay
 * separate kernels joined by the offsets kernel. */
int sl_decode_triangulate_fused(const uint8_t* frames, int B, int F, int H, int W, size_t frame_pitch,
                                size_t scan_pitch, double eps, const uint8_t* white_rgb, const SlCalib* calib,
                                int precision, const float* cam_rays, const float* proj_rays, int16_t* h_idx,
                                int16_t* v_idx, void* workspace, float* cloud, int64_t* counts, int sm_reserve,
                                void* stream);

/* Triangulate.triangulate (triangulate.py:73-97) on caller-supplied point pairs.
 * cam_pts, proj_pts device f32 (N,2); out device f64 (3,N) like the reference's return. */
int sl_triangulate_points(const float* cam_pts, const float* proj_pts, int64_t N,
                          const SlCalib* calib, int precision, double* out, void* stream);

/* Triangulate.filter_3d_pts (triangulate.py:99-122): keeps points with |x|,|y|,|z| < threshold
 * (strict), order preserved.  pts device f64 (3,N); colors device f64 (N,3) (nullable).
 * out_pts device f64 with room for (3,N): row r of the result starts at out_pts + r*N and holds M
 * values; out_colors device f64 (N,3) capacity, first M rows written; out_count device int64[1] = M.
 * workspace: device scratch of sl_filter_workspace_bytes(N). */
int sl_filter_box(const double* pts, const double* colors, int64_t N, double threshold,
                  double* out_pts, double* out_colors, int64_t* out_count, void* workspace,
                  size_t workspace_bytes, void* stream);
size_t sl_filter_workspace_bytes(int64_t N);

/* The same filter on the packed XYZRGB cloud of sl_emit_cloud (src/4-triangulate.py:71 applied to the
 * fast path's output): cloud device f32 (N,4) -> out_cloud (first M records), out_count device int64[1] = M.
 * Coordinates are compared as doubles against `threshold`, like the reference compares its float64 points. */
int sl_filter_cloud_box(const float* cloud, int64_t N, double threshold, float* out_cloud,
                        int64_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* pcd.remove_statistical_outlier(nb_neighbors, std_ratio) as scanner/utils/visualize.py:104-105 applies it to
 * the filtered cloud (Open3D 0.17 PointCloud::RemoveStatisticalOutliers; the reference narrows the points to
 * float32 first, visualize.py:100): mean distance of every point to its nb_neighbors nearest points (itself
 * included), threshold = mean + std_ratio * standard deviation (Bessel) of those means, inliers
 * 0 < mean_dist < threshold in ascending index order.  Exact k-NN on a hashed uniform grid; 1 <= nb_neighbors <= 128.
 *   cloud       device f32 (N,4) XYZRGB records (only x, y, z are read; records with a non-finite coordinate
 *               are dropped and are nobody's neighbour)
 *   out_cloud   nullable device f32, capacity (N,4): the M inlier records
 *   out_index   nullable device int64 [N]: the M inlier indices (Open3D's `ind`)
 *   mean_dist   nullable device f64 [N]: the per-point mean neighbour distance (-1 for dropped records)
 *   stats       nullable device f64 [16]: cloud mean, standard deviation, threshold, finite records, then the
 *               number of queries handled at each of the 12 grid levels (diagnostics)
 *   out_count   device int64 [1] = M
 *   workspace   device scratch of sl_outlier_workspace_bytes(N) */
int sl_remove_statistical_outliers(const float* cloud, int64_t N, int nb_neighbors, double std_ratio,
                                   float* out_cloud, int64_t* out_index, double* mean_dist, double* stats,
                                   int64_t* out_count, void* workspace, size_t workspace_bytes, void* stream);
size_t sl_outlier_workspace_bytes(int64_t N);

/* -------------------------------------------------------------- host-buffer pipeline */

typedef struct SlPipeline SlPipeline;

/* A pipeline owns device buffers, pinned staging and streams for scans of one geometry.
 * max_batch scans of F frames of H x W are staged per call. */
int sl_pipeline_create(SlPipeline** out, int device, int max_batch, int F, int H, int W);
void sl_pipeline_destroy(SlPipeline* p);
/* sm_reserve (see sl_decode_u8) of this pipeline's decode launches; default 0.  The pipeline entry points select the
 * pipeline's device for the duration of the call and restore the caller's current device before returning. */
int sl_pipeline_set_sm_reserve(SlPipeline* p, int sms);

/* Whole path for B (<= max_batch) scans held in HOST memory:
 * H2D of the frame stacks and white images, decode, compaction, triangulation, D2H.
 *   frames_host     u8 [B][F][H][W]         white_rgb_host  u8 [B][H][W][3] (nullable)
 *   h_idx_host, v_idx_host   nullable int16 [B][H][W]
 *   cloud_host      f32, capacity cap_points*4: XYZRGB records as sl_emit_cloud writes them
 *   counts_host     int64 [B+1] (last = total)
 * Fails with SL_ERR_BAD_ARG if the total exceeds cap_points (counts_host is still written). */
int sl_scan_to_cloud_host(SlPipeline* p, const uint8_t* frames_host, const uint8_t* white_rgb_host,
                          int B, const SlCalib* calib, int precision, double eps,
                          int16_t* h_idx_host, int16_t* v_idx_host, float* cloud_host,
                          int64_t cap_points, int64_t* counts_host);

/* Streaming form of the same path (one scan per call, as a capture loop produces them): sl_pipeline_submit queues the
 * H2D copies, the kernels and the count read-back of one scan on the next of the pipeline's streams and returns; the
 * scan submitted 3 calls earlier is retired first (its count is written to its count_host, its cloud copy queued).
 * sl_pipeline_flush retires everything still in flight and returns when every submitted scan's cloud and count are
 * in host memory (results may only be read after it).  Host buffers must stay valid until then; pinned memory keeps
 * the copies asynchronous.
 *   frames_host  u8 [F][H][W]     white_rgb_host  u8 [H][W][3] (nullable)
 *   cloud_host   f32, capacity cap_points*4      count_host  int64 [1]
 * A scan with more than cap_points points gets its count but not its cloud; sl_pipeline_flush then fails with
 * SL_ERR_BAD_ARG.  The blocking calls fail with SL_ERR_STATE while submitted scans are in flight. */
int sl_pipeline_submit(SlPipeline* p, const uint8_t* frames_host, const uint8_t* white_rgb_host,
                       const SlCalib* calib, int precision, double eps, float* cloud_host, int64_t cap_points,
                       int64_t* count_host);
/* sl_pipeline_submit for a capture loop's own frames: frames_bgr_host u8 [F][H][W][3], 8-bit BGR as cv2.VideoCapture
 * delivers them.  cv2.cvtColor(frame, COLOR_BGR2GRAY) (src/3-capture_decode.py:66, decode_codes.py:86) happens inside
 * the decode load (as in sl_decode_bgr_u8), bit-identical; 3x the PCIe bytes, no host pass over the frames. */
int sl_pipeline_submit_bgr(SlPipeline* p, const uint8_t* frames_bgr_host, const uint8_t* white_rgb_host,
                           const SlCalib* calib, int precision, double eps, float* cloud_host, int64_t cap_points,
                           int64_t* count_host);
int sl_pipeline_flush(SlPipeline* p);

/* Decode only, HOST buffers (get_codes + gray_to_decimal loops for one or more scans). */
int sl_decode_host(SlPipeline* p, const uint8_t* frames_host, int B, double eps,
                   int16_t* h_idx_host, int16_t* v_idx_host, int8_t* h_codes_host,
                   int8_t* v_codes_host);

/* ------------------------------------------------- multi-GPU: clouds gathered over NVLink peer memory */

/* BASELINE.json's partitioning (the reference itself is a single process): one process per GPU, independent scans
 * block-sharded over the ranks (360-degree / turntable batches) or one huge capture cut into bands of rows; the only
 * exchange is the gather of the compacted clouds.  Because the cloud is an ordered compaction (scan-major, every scan
 * in the reference's x-major order, scanner/triangulation/triangulate.py:52-53), a rank must know where its records
 * start in the gathered cloud before it emits.  So the ranks first exchange their counts through small control blocks
 * in peer-mapped memory (sl_exchange_offsets_*: publishes this rank's counts to every rank, waits for everybody's,
 * rebases the offsets in this rank's compaction workspace), then sl_emit_cloud_multi writes the records straight into
 * the gathering rank's cloud through its peer mapping (the transfer rides on the emission, no staging, padding or
 * concatenation), and sl_exchange_arrive tells the gathering rank(s) that they have landed.
 * The caller provides the plumbing: one process per GPU, the 64-byte handles carried between the processes by any
 * means (torch.distributed in slb200/dist.py). */
#define SL_PEER_HANDLE_BYTES 64
#define SL_MAX_RANKS 16

/* Device memory other processes can map (cudaMalloc + CUDA IPC), zero-filled.  sl_peer_export writes the 64-byte
 * handle of an allocation made by sl_peer_alloc in THIS process; sl_peer_open maps another process's allocation into
 * this one (peer access over NVLink is enabled on first use) and returns its address here; sl_peer_close unmaps. */
int sl_peer_alloc(size_t bytes, void** ptr);
int sl_peer_free(void* ptr);
int sl_peer_export(const void* ptr, void* handle64);
int sl_peer_open(const void* handle64, void** ptr);
int sl_peer_close(void* ptr);

typedef struct SlExchange SlExchange;

/* Size of one rank's control block (allocate it with sl_peer_alloc, zero-filled, and open it on every other rank). */
size_t sl_exchange_ctrl_bytes(int world, int max_scans, int max_width);
/* ctrl_ptrs[r] = control block of rank r as mapped in this process (ctrl_ptrs[rank] is the local allocation);
 * max_scans = most scans one rank processes per step, max_width = widest image (row-band mode);
 * timeout_s: a wait for a peer gives up after this long (<= 0: 20 s) and reports through sl_exchange_status. */
int sl_exchange_create(SlExchange** out, int rank, int world, void* const* ctrl_ptrs, int max_scans, int max_width,
                       double timeout_s);
void sl_exchange_destroy(SlExchange* x);

/* Scans sharded over ranks.  After sl_compact_offsets[_decoded] of this rank's B scans (counts = its int64 [B+1]
 * output, still on the device): every offset in `workspace` is advanced by the number of points of all lower ranks,
 * so that the emission writes at the packed position of the whole batch.  counts_all (nullable, device int64
 * [total scans + 1]): every scan's point count in rank order, then the grand total.  Asynchronous on `stream`;
 * every rank must make the same sequence of sl_exchange_* calls.  workspace may be null: then only the counts are
 * exchanged (this rank's first record and the total stay in the exchange for sl_exchange_push) and the offsets keep
 * their local values. */
int sl_exchange_offsets_by_scan(SlExchange* x, int B, int H, int W, void* workspace, const int64_t* counts,
                                int64_t* counts_all, void* stream);
/* One capture cut into row bands, rank r holding the r-th band (H rows here, boundaries multiples of 32 rows, calib
 * cam_row0 = first row): after sl_compact_offsets[_decoded](1, H, W) the per-column totals of all ranks are exchanged
 * and the offsets in `workspace` become the positions of this band's runs inside the columns of the whole image.
 * counts_all (nullable, device int64 [world + 1]): points per rank, then the total. */
int sl_exchange_offsets_row_band(SlExchange* x, int H, int W, void* workspace, const int64_t* counts,
                                 int64_t* counts_all, void* stream);

/* sl_emit_cloud writing every record to n_clouds buffers (1..SL_MAX_RANKS; local or peer-mapped, each 16-byte aligned):
 * clouds[0] = the gathering rank's cloud for a gather, all ranks' clouds for an all-gather.  max_points = capacity of
 * the gathered cloud in records (the rebased offsets can exceed B*H*W). */
int sl_emit_cloud_multi(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb,
                        const SlCalib* calib, int precision, const float* cam_rays, const float* proj_rays,
                        int B, int H, int W, const void* workspace, float* const* clouds, int n_clouds,
                        int64_t max_points, void* stream);

/* sl_emit_cloud (fp32, lens tables, local cloud) with the lane = row mapping: the same records, bit for bit, from
 * TRANSPOSED lens tables -- cam_rays_t float2 [W][cam_rows] (cam_rays_t[x * cam_rows + y] = cam_rays[y * W + x]) and
 * proj_bt_t float2 [proj_w][proj_h] (the SL_PREC_FP32 projector table transposed).  Needs W % 32 == 0 and 16-byte aligned
 * h_idx / v_idx / white_rgb / cloud (SL_ERR_BAD_ARG otherwise: use sl_emit_cloud).  calib->cam_row0 + H <= cam_rows. */
/* dst[c * rows + r] = src[r * cols + c] for a table of float2 entries (rows x cols): the transposed lens tables of
 * sl_emit_cloud_rows from the ones sl_lens_tables builds (camera table: rows = h, cols = w; SL_PREC_FP32 projector table:
 * rows = proj_h, cols = proj_w).  Once per calibration. */
int sl_transpose_table(const float* src, int rows, int cols, float* dst, void* stream);
int sl_emit_cloud_rows(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb, const SlCalib* calib,
                       const float* cam_rays_t, int cam_rows, const float* proj_bt_t, int B, int H, int W,
                       const void* workspace, float* cloud, void* stream);

/* sl_emit_cloud_rows writing every record to n_clouds buffers (local or peer-mapped), as sl_emit_cloud_multi does for the
 * tiled emission: clouds[0] = the gathering rank's cloud for a gather, all ranks' clouds for an all-gather. */
int sl_emit_cloud_rows_multi(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb, const SlCalib* calib,
                             const float* cam_rays_t, int cam_rows, const float* proj_bt_t, int B, int H, int W,
                             const void* workspace, float* const* clouds, int n_clouds, void* stream);

/* sl_emit_cloud into LOCAL memory for a workspace whose offsets sl_exchange_offsets_* has rebased: cloud is laid out
 * like the gathered cloud (capacity max_points records). */
int sl_emit_cloud_at(const int16_t* h_idx, const int16_t* v_idx, const uint8_t* white_rgb,
                     const SlCalib* calib, int precision, const float* cam_rays, const float* proj_rays,
                     int B, int H, int W, const void* workspace, float* cloud, int64_t max_points, void* stream);

/* The other transport for scan-sharded batches.  A rank's records of one step are ONE contiguous block of the gathered
 * cloud, so instead of emitting through the peer mapping (NVLink store requests of 256-512 bytes at the pace of the
 * emission) the rank emits locally at full speed and n_ctas CTAs (<= 0: 32) then copy the block as whole 128-byte lines
 * while the other SMs go on with the next batch.  src = the local records: packed from src[0] (src_rebased = 0, after
 * sl_exchange_offsets_by_scan with a null workspace + sl_emit_cloud) or laid out like the gathered cloud
 * (src_rebased = 1, after the rebasing exchange + sl_emit_cloud_at); total = device pointer to this rank's point count
 * (counts + B); clouds = the n_clouds destinations (peer-mapped or local, 16-byte aligned), each of capacity max_points.
 * No host synchronisation.  If the block does not fit, nothing is written and sl_exchange_status reports error 3. */
int sl_exchange_push(SlExchange* x, const float* src, int src_rebased, const int64_t* total, float* const* clouds,
                     int n_clouds, int64_t max_points, int n_ctas, void* stream);

/* After the emission: tell the ranks in dst_mask (bit r = rank r) that this rank's records of the current step have
 * landed; with wait != 0 also wait (on the stream, not the host) until every rank has told this one the same --
 * the gathered cloud is then complete for whatever is queued next on `stream`. */
int sl_exchange_arrive(SlExchange* x, uint32_t dst_mask, int wait, void* stream);

/* Synchronous: error = 0, 1 / 2 if a wait in sl_exchange_offsets_* / sl_exchange_arrive timed out, 3 if sl_exchange_push
 * found the gathered cloud too small (results of that step are then invalid); rank_base / grand_total (nullable) = this rank's first record and the total of the last step. */
int sl_exchange_status(SlExchange* x, int* error, int64_t* rank_base, int64_t* grand_total);

/* ---- Host-side conversions at the NumPy boundary of the reference API (no device work) ----------------------------------
 * The reference's scripts hand over a float64 frame stack (src/3-capture_decode.py:66-70) and expect int64 index images
 * (decode_codes.py:189-229), a (3,N) float64 point array (triangulate.py:95) and (N,3) float64 colours (:64-69).  These
 * do the dtype changes in one multi-threaded host pass each (threads <= 0: one per hardware thread, at most 32), from /
 * into pinned staging, so that only uint8 samples, int16 indices and 16-byte records cross PCIe.  Exact or an error. */

/* src (src_type 0 = float64, 1 = float32; n values) -> uint8.  SL_ERR_BAD_ARG if any value is not an integer in 0..255. */
int sl_host_narrow_u8(const void* src, int src_type, size_t n, uint8_t* dst, int threads);
/* int16 indices -> the reference's int64 h_pixels / v_pixels. */
int sl_host_widen_idx(const int16_t* src, size_t n, int64_t* dst, int threads);
/* int64 h_pixels / v_pixels -> int16.  SL_ERR_BAD_ARG if any value lies outside -1..32767. */
int sl_host_narrow_idx(const int64_t* src, size_t n, int16_t* dst, int threads);
/* n packed records (host) -> xyz_3xn: (3,n) float64 row-major (= Triangulate.triangulate's return) and, if non-null,
 * colors_nx3: (n,3) float64 = channel / 255 as an IEEE double division (triangulate.py:69). */
int sl_host_unpack_cloud(const float* cloud, int64_t n, double* xyz_3xn, double* colors_nx3, int threads);

/* Number of kernels this library has launched in the calling process (all streams). */
int64_t sl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SLB200_H_ */

/*
 * x360.h — C-ABI of libx360 (B200 / sm_100a equirect -> rectilinear extraction).
 *
 * The reference (nicolasdiolez/360Extractor) has no FFI of its own: its boundary
 * for this path is four Python call signatures plus the loop that calls them.
 * Each entry point below names the reference interface it replaces (file:line
 * relative to the reference root).  Plain C types only; no C++ types, no
 * exceptions and no callbacks cross this ABI.  All functions return 0 on
 * success or a negative X360_E_* code; the message is in x360_last_error()
 * (thread-local).  There is no CPU fallback: every compute entry fails with
 * X360_E_CUDA when no usable device is present.
 *
 * Threading: a plan is used from one thread at a time (the reference calls this
 * path from exactly one worker thread, src/core/processor.py:52-78).
 * Ownership: the caller owns every input/output buffer and keeps it alive until
 * the stream is synchronised; a plan owns only device scratch, its constant
 * tables and its streams/events.
 */
#ifndef X360_H_
#define X360_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X360_ABI_VERSION 1

/* error codes */
#define X360_OK          0
#define X360_E_INVALID  -1   /* bad argument */
#define X360_E_CUDA     -2   /* CUDA runtime error / no device */
#define X360_E_NOMEM    -3
#define X360_E_UNSUPPORTED -4

/* src/core/processor.py:199-200 — 'lanczos' => INTER_LANCZOS4, else INTER_LINEAR */
#define X360_INTERP_LINEAR   0
#define X360_INTERP_LANCZOS4 1

/* kernel-path selector (x360_params.path): 0 lets the plan choose */
#define X360_PATH_AUTO    0
#define X360_PATH_DIRECT  1   /* taps gathered straight from global/L2 through L1   (comparison path: only in libraries */
#define X360_PATH_STAGED  2   /* the first staged bilinear kernel                    built with -DX360_LAB; X360_E_UNSUPPORTED otherwise) */
#define X360_PATH_TILED   3   /* TMA in / pair records / dp2a gather / TMA out, maps shared across yaw-shifted views
                                 (bilinear and Lanczos4; what AUTO resolves to) */

typedef struct x360_plan x360_plan;

/*
 * One plan = one job's geometry: replaces the per-job map build
 *   src/core/processor.py:245-265  (generate_views -> create_rectilinear_map per active view)
 * The maps themselves are never materialised: the plan keeps the per-view fp64
 * rotation matrices and the kernels evaluate src/core/geometry.py:141-190 on the
 * fly per output tile.
 */
typedef struct x360_params {
    int32_t struct_size;      /* = sizeof(x360_params) */
    int32_t src_w, src_h;     /* equirect frame, uint8 BGR [src_h][src_w][3]   (processor.py:250-254) */
    int32_t out_w, out_h;     /* view size; the reference always passes out_res twice (processor.py:263-265) */
    double  fov_deg;          /* horizontal FOV (geometry.py:141) */
    int32_t interp;           /* X360_INTERP_* */
    int32_t n_views;          /* emitted views (after the active_cameras filter, processor.py:259-261) */
    const double *R;          /* [n_views][9] row-major fp64 from get_rotation_matrix (geometry.py:80-119),
                                 built by the caller with the reference's own numpy ops */
    int32_t device;           /* CUDA ordinal */
    int32_t max_batch;        /* frames per pipelined chunk in x360_extract_host (0 = default 2: small chunks keep both
                                 PCIe directions busy; the kernel is far from being the bottleneck there) */
    int32_t path;             /* X360_PATH_* */
    int32_t flags;            /* X360_FLAG_* overrides of the plan's own choices (0 = let the plan choose) */
} x360_params;

/* x360_params.flags: explicit overrides of what the plan would choose by itself (every variant computes the same bytes;
 * the parity suite runs each one).  The library reads no environment variable unless it is built with -DX360_LAB. */
#define X360_FLAG_PLANES       1   /* bilinear: pair-record planes (transform pass) instead of the raw gather */
#define X360_FLAG_FUSED_SCORE  2   /* blur score fused into the extraction kernel (gray tile + 1-px halo per output tile) */
#define X360_FLAG_SPLIT_SCORE  4   /* blur score by one streaming pass over the finished views (the default: measured faster) */
#define X360_FLAG_TILE_H16     8   /* 32 x 16 output tiles */
#define X360_FLAG_TILE_H32    16   /* 32 x 32 output tiles */
#define X360_FLAG_NO_WIDE     32   /* no wide / seam boxes: tiles beyond the regular staging capacity take the direct gathers */
#define X360_FLAG_FULL_UPLOAD 64   /* host entries: upload whole frames even where the views read only a band of source rows */

int x360_plan_create(const x360_params *params, x360_plan **out_plan);
int x360_plan_destroy(x360_plan *plan);

/*
 * The batched hot path: replaces the body of the per-view loop
 *   src/core/processor.py:327-342   rect = cv2.remap(frame, map_x, map_y, interp, BORDER_WRAP)
 *                                   score = ImageUtils.calculate_blur_score(rect)
 * for n_frames frames x n_views views in one call.
 *   frames     uint8 [n_frames][src_h][src_w][3]
 *   out_views  uint8 [n_frames][n_views][out_h][out_w][3]
 *   sum_lap / sum_lap2  int64 [n_frames][n_views], both NULL to skip the score
 *       (exact integer sum of L and L^2 of the 4-neighbour Laplacian of the
 *        Q15 gray view, src/utils/image_utils.py:22-27; score = s2/N - (s/N)^2, N = out_w*out_h)
 *
 * x360_extract_device: all pointers are DEVICE pointers on the plan's device;
 *   work is enqueued on cuda_stream (a cudaStream_t, may be NULL) and the call
 *   returns without synchronising.
 * x360_extract_host: all pointers are HOST pointers (pinned for full speed);
 *   frames are uploaded, processed and downloaded in chunks of max_batch frames
 *   on the plan's own side streams (H2D / compute / D2H overlap); synchronous.
 * x360_extract: dispatches on the kind of `frames` pointer (device vs host);
 *   out/sum pointers must be of the same kind.
 */
int x360_extract_device(x360_plan *plan, const uint8_t *frames, int32_t n_frames,
                        uint8_t *out_views, int64_t *sum_lap, int64_t *sum_lap2, void *cuda_stream);
int x360_extract_host(x360_plan *plan, const uint8_t *frames, int32_t n_frames,
                      uint8_t *out_views, int64_t *sum_lap, int64_t *sum_lap2);
int x360_extract(x360_plan *plan, const uint8_t *frames, int32_t n_frames,
                 uint8_t *out_views, int64_t *sum_lap, int64_t *sum_lap2, void *cuda_stream);

/*
 * Materialise one view's float32 maps (HOST pointers, [out_h][out_w] each):
 * replaces GeometryProcessor.create_rectilinear_map (src/core/geometry.py:122-192)
 * for callers that still want the arrays (tests/test_core.py:67-78, the GUI preview).
 */
int x360_make_maps(x360_plan *plan, int32_t view, float *map_x, float *map_y);

/*
 * Explicit-map remap, HOST pointers: replaces
 *   cv2.remap(frame, map_x, map_y, interp, borderMode=cv2.BORDER_WRAP)   (processor.py:334, analyzer.py:63)
 * for a caller that holds its own float32 maps.  src uint8 [src_h][src_w][3],
 * maps float32 [out_h][out_w], dst uint8 [out_h][out_w][3].
 */
int x360_remap_maps(int32_t device, const uint8_t *src, int32_t src_h, int32_t src_w,
                    const float *map_x, const float *map_y, int32_t out_h, int32_t out_w,
                    int32_t interp, uint8_t *dst);

/*
 * Laplacian-variance sums of one HOST image (channels = 3 BGR or 1 gray):
 * replaces ImageUtils.calculate_blur_score (src/utils/image_utils.py:6-27).
 */
int x360_blur_sums(int32_t device, const uint8_t *image, int32_t h, int32_t w, int32_t channels,
                   int64_t *sum_lap, int64_t *sum_lap2);

/*
 * Host-only helpers (no device needed), exposed so the constant tables and the focal
 * length the plan derives can be checked against the oracle without a GPU:
 *   x360_lanczos4_table  int16 [32*32][8][8], entry fy*32+fx — cv2.remap's INTER_LANCZOS4 weights
 *   x360_focal           f of src/core/geometry.py:141
 */
int x360_lanczos4_table(int16_t *out_tab);
/* the same table as the factors the tiled kernel computes its weights from: k1d float32 [32][8] (cv2's 1-D Lanczos-4
 * coefficients), fix uint32 [1024] (per (fy, fx): bits 0-1 = p, the tap (4 + (p >> 1), 4 + (p & 1)) that the table's
 * sum-to-2^15 fix-up changed; bits 16-31 = the signed residue it added).  it[r][c] = sat_s16(rint(f32(k[fy][r] k[fx][c]) 2^15)) + fix */
int x360_lanczos4_factors(float *k1d, uint32_t *fix);
double x360_focal(int32_t out_w, double fov_deg);
/*
 * How the tiled bilinear path would organise a plan (host-only, no device needed):
 *   n_units       number of view units; the views of a unit differ only by a yaw (ring layouts, the four
 *                 horizontal cube faces: src/core/geometry.py:24-37,71-75), so they share ONE map evaluation
 *   unit_of_view  [n_views] unit index of each view (units are ordered largest first)
 *   shift_px      [n_views] map_x shift of the view relative to its unit's first view, in source pixels [0, W)
 *   rows_max, bw_max  [2] staging capacity chosen from a survey of the tiles' source footprints
 *                 (footprint rows; raw bytes per footprint row, a multiple of 48)
 *   tile_h        [2] output tile height the plan works in: 32, or 16 where 32-row footprints would leave
 *                 too few CTAs per SM (tiles are 32 pixels wide)
 *                 index 0: extraction without scores, index 1: with the fused blur score (1-px halo)
 * Any output pointer may be NULL.
 */
int x360_plan_preview(const x360_params *params, int32_t *n_units, int32_t *unit_of_view, double *shift_px,
                      int32_t *rows_max, int32_t *bw_max, int32_t *tile_h);
/*
 * A plan serves its views in one or two groups, each with its own kernel variant, tile height and launch: the yaw-only
 * views (map_x independent of the output row: rings at pitch 0, the four horizontal cube faces) and the others (pole
 * faces, inclined and fibonacci views).  group_of_view [n_views]; tile_h_of_view [n_views][2] (without / with a fused
 * score).  x360_plan_preview's rows_max / bw_max / tile_h describe the group of view 0.  Host-only.
 */
int x360_plan_preview_views(const x360_params *params, int32_t *group_of_view, int32_t *tile_h_of_view);

/*
 * Device-resident hand-off of a view batch to the masker: replaces the host round trip of
 * src/core/processor.py:392-420 (views -> AIService.process_batch, src/core/ai_model.py:125-167, whose model
 * normalises each BGR image to a CHW float tensor: BGR -> RGB, HWC -> CHW, / 255).
 * views: DEVICE uint8 [n_views][h][w][3] (the output of x360_extract_device, flat over frames and views);
 * select: HOST array of n_select view slots to convert — the faces 'ai_mask_cameras' names, processor.py:405-408 —
 * or NULL for all n_views in order; out: DEVICE float32 [n_select or n_views][3][h][w], value = float(v) / 255
 * (correctly rounded); to_rgb != 0 swaps B and R.  Asynchronous on cuda_stream.
 */
int x360_views_to_nchw(int32_t device, const uint8_t *views, int64_t n_views, int32_t h, int32_t w,
                       const int32_t *select, int32_t n_select, int32_t to_rgb, float *out, void *cuda_stream);

/*
 * Unsharp mask on a batch of views, the step after the blur filter:
 *   gaussian = cv2.GaussianBlur(rect_img, (0, 0), 2.0)
 *   rect_img = cv2.addWeighted(rect_img, 1.0 + strength, gaussian, -strength, 0)
 * (src/core/processor.py:379-381, src/ui/preview_widget.py:118-120; strength = settings
 * 'sharpening_strength', processor.py:229).  images / out: uint8 [n_images][h][w][3], out != images.
 * Bit-exact with OpenCV's fixed-point 13-tap Gaussian (BORDER_REFLECT_101) and its float32 addWeighted.
 * _device: DEVICE pointers, asynchronous on cuda_stream.  _host: HOST pointers, synchronous.
 */
/*
 * Make x360_extract_host return sharpened views: every chunk is unsharp-masked on the device between the
 * extraction kernel and the download (no extra trip over PCIe).  The Laplacian sums stay those of the PLAIN
 * views, as in the reference, which scores and filters before it sharpens (processor.py:341-381).
 */
int x360_plan_set_sharpen(x360_plan *plan, int32_t enabled, double strength);

int x360_unsharp_device(int32_t device, const uint8_t *images, int64_t n_images, int32_t h, int32_t w,
                        double strength, uint8_t *out, void *cuda_stream);
int x360_unsharp_host(int32_t device, const uint8_t *images, int64_t n_images, int32_t h, int32_t w,
                      double strength, uint8_t *out);

/*
 * Host-only: 1 if map_x of every view (row-major fp64 R[n_views][9], geometry.py:80-119) is independent of the
 * output row — yaw-only rotations: ring layouts at pitch 0 (geometry.py:71-75), the horizontal cube faces
 * (geometry.py:30-33).  Such plans keep one fp64 map_x row per warp instead of a tile and run 7 CTAs per SM.
 */
int x360_column_only_map_x(const double *R, int32_t n_views);

/*
 * Baseline JPEG encode of a view batch on the device (the codec step after the path): replaces the
 * reference's save step
 *     cv2.imwrite(path, rect_img, [cv2.IMWRITE_JPEG_QUALITY, quality])     src/utils/file_manager.py:21-32,
 *     save_params at src/core/processor.py:121-123, quality default 95 (src/core/settings_manager.py:24)
 * so that only compressed bytes cross PCIe.  Parity definition: the bytes of every file are IDENTICAL to what cv2.imwrite /
 * cv2.imencode of opencv-python 4.13.0 (libjpeg-turbo 3.1.2) produce for the same pixels and quality: YCbCr 4:2:0, baseline
 * sequential, integer "islow" DCT, IJG-scaled Annex-K quantisation tables, Annex-K Huffman tables, JFIF 1.01 header.
 * images: uint8 [n][h][w][3] BGR.  The files are written back to back into `out`; offsets [n + 1] (int64) receives the
 * byte offset of each file and, last, the total.
 *   x360_jpeg_create          encoder for images of h x w, at most max_images per call, quality 1..100
 *   x360_jpeg_encode_device   DEVICE pointers (images, out, offsets), asynchronous on cuda_stream; x360_jpeg_status
 *                             synchronises the stream and reports X360_E_NOMEM if out_capacity (or the encoder's scratch,
 *                             3 bytes per pixel of the batch) was too small — nothing is written past the capacity
 *   x360_jpeg_encode_host     HOST pointers, any n (chunks of max_images), synchronous
 *   x360_jpeg_max_bytes       a capacity that always suffices for n_images
 *   x360_jpeg_coefficients_host  the quantised DCT coefficients: int16 [n][mcus][6][64], per 16 x 16 MCU the blocks
 *                             Y00 Y01 Y10 Y11 Cb Cr in zigzag order (inspection / tests)
 *   x360_jpeg_header          host-only: the SOI..SOS header of an h x w file; returns its length
 */
typedef struct x360_jpeg x360_jpeg;
int x360_jpeg_create(int32_t device, int32_t h, int32_t w, int32_t max_images, int32_t quality, x360_jpeg **out_encoder);
int x360_jpeg_destroy(x360_jpeg *encoder);
int64_t x360_jpeg_max_bytes(const x360_jpeg *encoder, int32_t n_images);
int x360_jpeg_encode_device(x360_jpeg *encoder, const uint8_t *images, int32_t n_images, uint8_t *out, int64_t out_capacity,
                            int64_t *offsets, void *cuda_stream);
int x360_jpeg_status(x360_jpeg *encoder, void *cuda_stream);
int x360_jpeg_encode_host(x360_jpeg *encoder, const uint8_t *images, int64_t n_images, uint8_t *out, int64_t out_capacity,
                          int64_t *offsets);
int x360_jpeg_coefficients_host(x360_jpeg *encoder, const uint8_t *images, int32_t n_images, int16_t *coef);
int x360_jpeg_header(int32_t h, int32_t w, int32_t quality, uint8_t *out, int32_t capacity);

/*
 * x360_extract_host with the save step's encode on the device: frames in (HOST), one JPEG file per (frame, view) out (HOST).
 * Replaces src/core/processor.py:327-390 + the cv2.imwrite of :428-431 for a frame batch: extraction (+ blur sums, + the
 * unsharp mask if x360_plan_set_sharpen is on) -> x360_jpeg encode of the chunk's views -> download of the files only.
 * out: the files back to back in (frame, view) order; offsets: int64 [n_frames * n_views + 1].  Files of views the caller's
 * blur filter rejects are simply not written to disk by the caller (the sums come back as in x360_extract_host).
 */
int x360_extract_host_jpeg(x360_plan *plan, const uint8_t *frames, int32_t n_frames, int32_t quality, uint8_t *out,
                           int64_t out_capacity, int64_t *offsets, int64_t *sum_lap, int64_t *sum_lap2);

/*
 * The blur accept / reject machine of src/core/processor.py:341-376 (state: blur_history = deque(maxlen=10) and
 * consecutive_blur_skips, processor.py:224-225; settings blur_filter_enabled / smart_blur_enabled / blur_threshold,
 * processor.py:219-221), as plain C so that the pipelined host path can decide per view which pixels to download.
 * x360_blur_accept returns 1 when the view is kept; the state carries across views AND frames in iteration order.
 */
typedef struct x360_blur_state {
    int32_t enabled, smart;
    double threshold;
    double history[10];          /* accepted scores, oldest first from history[head] (ring buffer) */
    int32_t hist_len, hist_head;
    int32_t consecutive_skips;
    int32_t reserved;
    int64_t skipped, forced;
} x360_blur_state;
void x360_blur_state_init(x360_blur_state *state, int32_t enabled, int32_t smart, double threshold);
int x360_blur_accept(x360_blur_state *state, double score);
/* x360_blur_accept over a table of n scores in iteration order ((frame, view) row-major): keep[i] = 1 where the view is kept */
int x360_blur_replay(x360_blur_state *state, const double *scores, int64_t n, uint8_t *keep);
/*
 * x360_extract_host that downloads only the views the blur machine keeps: per chunk the Laplacian sums come back first
 * (a few bytes), the machine is replayed on the host in (frame, view) order, and one copy per KEPT view is queued — rejected
 * views never cross PCIe.  out_views: the kept views back to back, at most capacity_views of them; keep: uint8
 * [n_frames * n_views] (1 = kept); scores: float64 [n_frames * n_views] (may be NULL); n_kept: their count.  With the
 * machine disabled (state->enabled == 0) every view is kept and no score is computed, as in processor.py:341.
 * Sharpening (x360_plan_set_sharpen) applies to what is downloaded, after the decision, as in processor.py:379-381.
 */
int x360_extract_host_filtered(x360_plan *plan, const uint8_t *frames, int32_t n_frames, x360_blur_state *state,
                               uint8_t *out_views, int64_t capacity_views, uint8_t *keep, double *scores, int64_t *n_kept);

/* pinned host memory for the pipelined path (cudaHostAlloc / cudaFreeHost) */
int x360_host_alloc(size_t bytes, void **out_ptr);
int x360_host_free(void *ptr);

/* diagnostics */
const char *x360_last_error(void);
const char *x360_version(void);
/* number of kernels this library has launched in this process (all plans) */
int64_t x360_launch_count(void);
/* which path the plan's extract kernel uses (X360_PATH_*), and its launch geometry */
int x360_plan_info(const x360_plan *plan, int32_t *path, int32_t *grid, int32_t *block, int32_t *smem_bytes);

/* (tile, view) pairs of the LAST extract launch per gather mode: [0] staged in shared memory by one regular TMA box,
 * [1] direct aligned-word gather, [2] direct byte-wise gather (pole rows / odd geometry), [3] staged by a wide box
 * (footprints near the poles), [4] staged across the +-pi seam.  Synchronises the device. */
int x360_plan_tile_modes(x360_plan *plan, uint32_t *out5);
/* 32 x 32 output blocks (summed over the views) whose two 16-row tiles the plan processes TRANSPOSED — lanes down the block's
 * rows, because there the source row changes faster along output x than along y (the left / right quadrants of a pole face):
 * a layout choice of the gather, the pixels are the same.  0 for plans without such blocks. */
int x360_plan_transposed_blocks(const x360_plan *plan, int64_t *n_blocks);
/* The band of source rows the plan's views can read (every tap of every output pixel, plus a margin): the host entries
 * (x360_extract_host, _host_jpeg, _host_filtered) upload only these rows of each frame — a ring of 90-degree views at pitch 0
 * reads half of an equirectangular frame.  [0, src_h) for plans that see a pole or whose band covers most of the frame, and
 * with X360_FLAG_FULL_UPLOAD.  The device entries read the frames they are given. */
int x360_plan_upload_rows(const x360_plan *plan, int32_t *first_row, int32_t *n_rows);

#ifdef __cplusplus
}
#endif
#endif /* X360_H_ */

/*
 * karios_b200.h -- C ABI of the B200-native KARIOS KLT matching hot path.
 *
 * One shared library (karios_b200/_lib/libkarios_b200.so, built for sm_100a by
 * __graft_entry__.build()).  Every entry point replaces one call site of the
 * reference (telespazio-tim/karios, Python + OpenCV); the citation after each
 * declaration is the reference interface it stands in for.
 *
 * Conventions
 *   - all image / point / result pointers are DEVICE pointers owned by the
 *     caller (e.g. torch CUDA tensors); pitches are in BYTES;
 *   - every call is asynchronous on the cudaStream_t passed as `stream`
 *     (a `void*` here so the header needs no CUDA include); nothing
 *     synchronises except kr_read_stats;
 *   - return value 0 = OK, negative = error (kr_status); the message of the
 *     last failure on the calling thread is kr_last_error();
 *   - a kr_ctx owns all scratch memory; it is NOT thread-safe: use one context
 *     per host thread (the reference calls klt_tracker from a thread pool in
 *     auto-ksize mode, klt.py:526-527);
 *   - no exceptions cross this boundary and there is no CPU fallback.
 */
#ifndef KARIOS_B200_H
#define KARIOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KR_API __attribute__((visibility("default")))

typedef struct kr_ctx kr_ctx;

typedef enum {
    KR_OK = 0,
    KR_ERR_INVALID = -1,      /* bad argument */
    KR_ERR_CUDA = -2,         /* CUDA runtime error (message has the detail) */
    KR_ERR_NOMEM = -3,        /* device allocation failed */
    KR_ERR_CAPACITY = -4,     /* image / corner count exceeds the context */
    KR_ERR_UNSUPPORTED = -5   /* dtype / kernel size not implemented */
} kr_status;

typedef enum { KR_U8 = 0, KR_U16 = 1, KR_I16 = 2, KR_F32 = 3 } kr_dtype;

/* Which rounding the Sobel row filter of cornerMinEigenVal uses in the last
 * columns of a row: OpenCV's vector loop ends at a multiple of
 * the SIMD batch and the scalar tail rounds differently.  KR_TAIL_AVX512 = the
 * opencv-python 4.13 AVX-512 build (tail from 32*floor(W/32)); KR_TAIL_NONE =
 * the FMA form everywhere; any value > 0 is taken as an explicit vector batch
 * width (16 would be an AVX2 build). */
#define KR_TAIL_NONE 0
#define KR_TAIL_AVX512 32

/* Parameters of one klt_tracker call: KLTConfiguration (karios/core/
 * configuration.py:36-50) plus the constants fixed in klt.py:128-132,143. */
typedef struct {
    int32_t max_corners;        /* maxCorners (<= 0: unlimited, bounded by the context) */
    int32_t block_size;         /* blocksize */
    int32_t win_size;           /* matching_winsize */
    int32_t max_level;          /* 1   (klt.py:130) */
    int32_t max_count;          /* 30  (klt.py:131) */
    int32_t ksize_mon;          /* laplacian_kernel_size (int or dict "mon") */
    int32_t ksize_ref;          /* laplacian_kernel_size (int or dict "ref") */
    int32_t invert_mon;         /* laplacian_invert_polarity (klt.py:419) */
    int32_t tail_mode;          /* KR_TAIL_* */
    int32_t compute_zncc;       /* kr_match_tile only: also fill zncc[] */
    int32_t compute_mi;         /* kr_match_tile only: also fill mutual_info[] / mi[] */
    double quality_level;       /* qualityLevel */
    double min_distance;        /* minDistance */
    double eps;                 /* 0.03 (klt.py:131) */
    double min_eig_threshold;   /* 1e-4 (cv2 default) */
    double back_threshold;      /* 0.1  (klt.py:143) */
    double zncc_min_score;      /* confidence_threshold, api/core.py:884 (rows below get NaN) */
} kr_klt_conf;

/* Device-side scalars of the last calls on a context, read back (one small
 * D2H copy + stream synchronise) with kr_read_stats. */
typedef struct {
    double min_a, max_a;        /* kr_minmax_mask, image a */
    double min_b, max_b;        /* kr_minmax_mask, image b */
    uint64_t valid;             /* valid pixels of the (auto) mask, klt.py:276 */
    float eig_max;              /* masked maximum of the min-eigenvalue map */
    uint32_t n_candidates;      /* local maxima found by the response kernel */
    uint32_t n_above_threshold; /* ... of which above qualityLevel * max */
    uint32_t n_sorted;          /* ... of which entered the sort / NMS */
    uint32_t n_corners;         /* corners returned (Ninit, klt.py:149) */
    uint32_t n_kept;            /* rows after the back-check (klt.py:150-154) */
    uint32_t nms_rounds;        /* grid-wide rounds the parallel NMS took */
    uint32_t overflow;          /* 1: candidate list exceeded the context capacity */
    uint32_t select_incomplete; /* 1: pre-selection too small; host re-ran with all candidates */
    uint32_t two_tier;          /* 1: the two-tier corner response produced the corners */
    uint32_t two_tier_fallback; /* 1: it could not decide (flat image / list overflow) */
    uint32_t n_border_maxima;   /* two-tier: border / mask-edge pixels checked for the maximum */
    uint32_t n_exact;           /* two-tier: candidates confirmed in OpenCV's arithmetic */
    uint32_t est_cut_bits;      /* two-tier: float bits of the running cut-off estimate tier 1 used (0: none) */
    uint32_t rows_skipped;      /* two-tier: 104-pixel row pieces ruled out as a whole by that estimate */
} kr_stats;

/* SoA result rows of klt_tracker (klt.py:166-168), capacity >= max corners. */
typedef struct {
    float *x0, *y0, *dx, *dy, *score;   /* float32 columns of the DataFrame */
    double *zncc;                       /* optional (may be NULL) */
    double *mutual_info;                /* optional: Studholme NMI, api/core.py:894-897 */
    double *mi;                         /* optional: 2 MI / (Hx + Hy), api/core.py:902-907 */
    int32_t capacity;
} kr_rows;

KR_API int kr_version(void);
KR_API const char *kr_last_error(void);
/* Kernel launches issued by this library since it was loaded (process-wide,
 * monotonic).  Measurement aid: bench.py reports the difference over its timed
 * region as "gpu_launches".  Replaces nothing in the reference. */
KR_API uint64_t kr_launch_count(void);

/* Workspace sized for tiles up to max_w x max_h and max_corners corners
 * (<= 0: up to one corner per candidate).  Replaces nothing: cv2/NumPy allocate
 * their temporaries per call (klt.py:46-48,433-434). */
KR_API int kr_ctx_create(int device, int max_w, int max_h, int max_corners, kr_ctx **out);
KR_API void kr_ctx_destroy(kr_ctx *ctx);
KR_API int kr_read_stats(kr_ctx *ctx, void *stream, kr_stats *host_out);

/* Corner pre-selection: by default only the strongest ~2*maxCorners candidates
 * enter the NMS (OpenCV sorts all of them, then stops at maxCorners).  When that
 * subset yields fewer than maxCorners corners, kr_stats.select_incomplete is 1
 * and the caller repeats the call after kr_set_select_all(ctx, 1). */
KR_API int kr_set_select_all(kr_ctx *ctx, int on);

/* Corner response implementation.  0 (default): two tiers -- integer bounds of
 * the response at every pixel, OpenCV's float arithmetic only for the masked
 * maximum and the candidates that survive the value cut-off (falls back to 1 on
 * its own when it cannot decide); 1: OpenCV's arithmetic at every pixel.  Both
 * return the same corners. */
KR_API int kr_set_corner_mode(kr_ctx *ctx, int mode);

/* np.nanmin / np.nanmax of both tiles (klt.py:46) and, when mask_out != NULL,
 * the auto mask (mon != 0) & (ref != 0) & finite & != nodata with its count
 * (klt.py:268-276).  a = monitored, b = reference.  Results -> kr_stats. */
KR_API int kr_minmax_mask(kr_ctx *ctx, const void *img_a, int64_t pitch_a, const void *img_b,
                          int64_t pitch_b, int dtype, int w, int h, int has_nodata_a,
                          double nodata_a, int has_nodata_b, double nodata_b, uint8_t *mask_out,
                          int64_t mask_pitch, void *stream);

/* cv2.Laplacian(_to_uint8(img) [255 - .], cv2.CV_8U, ksize) (klt.py:42-49,
 * 419, 433-434).  slot 0 / 1: normalise with the min/max kr_minmax_mask left
 * for image a / b; slot -1: compute this image's min/max first. */
KR_API int kr_u8_laplacian(kr_ctx *ctx, const void *img, int64_t pitch, int dtype, int w, int h,
                           int slot, int ksize, int invert, uint8_t *out, int64_t out_pitch,
                           void *stream);

/* cv2.cornerMinEigenVal(img, blockSize, ksize=3): first stage of
 * goodFeaturesToTrack, exposed for parity tests. */
KR_API int kr_corner_min_eigen_val(kr_ctx *ctx, const uint8_t *img, int64_t pitch, int w, int h,
                                   int block_size, int tail_mode, float *eig, int64_t eig_pitch,
                                   void *stream);

/* cv2.goodFeaturesToTrack(img, maxCorners, qualityLevel, minDistance, mask,
 * blockSize) (klt.py:120,494).  out_xy: [capacity][2] float32 in OpenCV's
 * order; *d_count (device int32) receives the corner count. */
KR_API int kr_good_features(kr_ctx *ctx, const uint8_t *img, int64_t pitch, const uint8_t *mask,
                            int64_t mask_pitch, int w, int h, int max_corners,
                            double quality_level, double min_distance, int block_size,
                            int tail_mode, float *out_xy, int capacity, int32_t *d_count,
                            void *stream);

/* cv2.pyrDown(src) -> ((w+1)/2, (h+1)/2) (pyramid of calcOpticalFlowPyrLK). */
KR_API int kr_pyr_down(kr_ctx *ctx, const uint8_t *src, int64_t pitch, int w, int h, uint8_t *dst,
                       int64_t dst_pitch, void *stream);

/* cv2.calcOpticalFlowPyrLK(prev, next, p0, None, winSize=(win,win), maxLevel,
 * criteria=(EPS|COUNT, max_count, eps), minEigThreshold) (klt.py:134-140).
 * p0/p1: [n][2] float32; d_count (device, may be NULL) overrides n. */
KR_API int kr_pyr_lk(kr_ctx *ctx, const uint8_t *prev, int64_t prev_pitch, const uint8_t *next,
                     int64_t next_pitch, int w, int h, const float *p0, int n,
                     const int32_t *d_count, int win, int max_level, int max_count, double eps,
                     double min_eig_threshold, float *p1, uint8_t *status, float *err,
                     void *stream);

/* klt_tracker(ref_data, image_data, mask, conf, p0) (klt.py:83-172) on uint8
 * (Laplacian) planes: corners (unless p0 given), LK forward + backward,
 * back-check, score.  Rows keep OpenCV's corner order; counts -> kr_stats
 * (n_corners = Ninit, n_kept = len(DataFrame)). */
KR_API int kr_klt_track(kr_ctx *ctx, const uint8_t *ref, int64_t ref_pitch, const uint8_t *mon,
                        int64_t mon_pitch, const uint8_t *mask, int64_t mask_pitch, int w, int h,
                        const kr_klt_conf *conf, const float *p0, int n_p0, kr_rows rows,
                        void *stream);

/* ZNCCService.compute_zncc(df, monitored, reference) (zncc_service.py:162-238):
 * one float64 per row, NaN where the reference returns NaN.  d_count (device,
 * may be NULL) overrides n. */
KR_API int kr_zncc(kr_ctx *ctx, const void *ref, int64_t ref_pitch, int ref_w, int ref_h,
                   const void *mon, int64_t mon_pitch, int mon_w, int mon_h, int dtype,
                   const float *x0, const float *y0, const float *dx, const float *dy, int n,
                   const int32_t *d_count, double *out, void *stream);

/* MutualInfoService.compute_mutual_info(df, monitored, reference)
 * (mutual_info_service.py:73-138, _mutual_info :32-63) -> out_studholme, and
 * ZNCCService.compute_mi(df, monitored, reference) (zncc_service.py:240-287,
 * _mutual_information :129-151) -> out_nmi, from ONE 32 x 32 joint histogram
 * (np.histogram2d semantics) of the two 57 x 57 chips per row.  Either output
 * may be NULL.  NaN where the reference returns NaN (chip outside the raster,
 * non-finite pixels, zero joint entropy). */
KR_API int kr_mutual_info(kr_ctx *ctx, const void *ref, int64_t ref_pitch, int ref_w, int ref_h,
                          const void *mon, int64_t mon_pitch, int mon_w, int mon_h, int dtype,
                          const float *x0, const float *y0, const float *dx, const float *dy, int n,
                          const int32_t *d_count, double *out_studholme, double *out_nmi,
                          void *stream);

/* KLT._match_tile (klt.py:236-349) for one tile window of full-size rasters,
 * without any host synchronisation: auto mask (mask == NULL) or user mask,
 * min/max, uint8 + Laplacian of both windows, klt_tracker, tile offsets,
 * sort by (x0, y0), and optionally ZNCC (compute_zncc) and the two mutual-
 * information scores (compute_mi) of rows with score >= zncc_min_score
 * against the FULL rasters (api/core.py:884-907).  mon/ref/mask point at the
 * full rasters (img_w x img_h); the tile is [x_off, x_off+tile_w) x
 * [y_off, y_off+tile_h). */
KR_API int kr_match_tile(kr_ctx *ctx, const void *mon, int64_t mon_pitch, const void *ref,
                         int64_t ref_pitch, int dtype, int img_w, int img_h, const uint8_t *mask,
                         int64_t mask_pitch, int x_off, int y_off, int tile_w, int tile_h,
                         int has_nodata_mon, double nodata_mon, int has_nodata_ref,
                         double nodata_ref, const kr_klt_conf *conf, kr_rows rows, void *stream);

/* KLT._match_tile_auto_ksize (klt.py:465-545; laplacian_kernel_size == "auto") for one
 * tile as ONE stream-ordered launch sequence without host synchronisation: min/max (+ auto
 * mask when mask == NULL), the n_k Laplacians of each raster, their pyramids (once per
 * plane), goodFeaturesToTrack once per reference kernel size, the n_k * n_k LK round trips
 * (product order: mon outer, ref inner), the inlier ratios n_kept / n_init, and the winner
 * -- highest ratio, first maximum wins (strict >, klt.py:536) -- chosen on the device and
 * its rows (OpenCV corner order, no offsets) copied to `rows`.  invert_mon of `conf` is the
 * polarity (klt.py:419); ksize_* are ignored.  Not covered here (the caller falls back to
 * one kr_klt_track per pair): max_corners <= 0, outlier filtering, and a corner pass that
 * reports select_incomplete (kr_auto_result.redo == 1).
 * scratch: device memory of kr_auto_ksize_scratch_bytes(); d_result: device record. */
#define KR_AUTO_MAX_K 8
typedef struct {
    int32_t best_mon, best_ref;     /* winning kernel sizes, 0 when no pair produced a result */
    int32_t n_init, n_kept;         /* Ninit and rows of the winner */
    int32_t redo;                   /* 1: a corner pass needs the exact re-run -- use the host loop */
    int32_t n_k;
    uint64_t valid;                 /* valid pixels of the (auto) mask; 0 => tile skipped */
    int32_t counts[KR_AUTO_MAX_K * KR_AUTO_MAX_K][2];   /* (n_init, n_kept) per pair, product order */
} kr_auto_result;
KR_API int64_t kr_auto_ksize_scratch_bytes(int w, int h, int n_k, int max_corners, int win_size,
                                           int max_level);
KR_API int kr_auto_ksize(kr_ctx *ctx, const void *mon, int64_t mon_pitch, const void *ref,
                         int64_t ref_pitch, int dtype, int w, int h, const uint8_t *mask,
                         int64_t mask_pitch, int has_nodata_mon, double nodata_mon,
                         int has_nodata_ref, double nodata_ref, const kr_klt_conf *conf,
                         const int32_t *ksizes, int n_k, void *scratch, int64_t scratch_bytes,
                         kr_rows rows, kr_auto_result *d_result, void *stream);

/* ---- full-frame passes either side of the matching path; these
 * take no context. ---- */

/* Element-wise steps of skimage.registration.phase_cross_correlation(mon, ref)
 * (upsample_factor 1, normalization "phase") as LargeOffsetMatcher.match calls it
 * (karios/matcher/large_offset.py:32-41).  a, b: interleaved complex spectra of n
 * elements (float32 or float64 pairs).  a <- a conj(b) / max(|a conj(b)|, 100 eps). */
KR_API int kr_cross_power(void *a, const void *b, int64_t n, int is_double, void *stream);

/* np.argmax(np.abs(x)) over n real values: first index of the maximum -> *out_index
 * (device int64).  scratch16: 16 bytes of device scratch. */
KR_API int kr_argmax_abs(const void *x, int64_t n, int is_double, void *scratch16, int64_t *out_index,
                         void *stream);

/* shift_image(img, y_off, x_off) (karios/core/image.py:70-101): dst[y][x] =
 * src[y + y_off][x + x_off], zero where that leaves the raster.  dst != src. */
KR_API int kr_shift_image(const void *src, int64_t src_pitch, void *dst, int64_t dst_pitch, int dtype,
                          int w, int h, int x_off, int y_off, void *stream);

/* Histogram of an integer raster: bin = (value - lo) >> shift for value >= lo and
 * bin < nbins (<= 8192 per call), ADDED to hist[nbins] (device uint64).
 * np.nanpercentile(image, [2, 98]) of KariosAPI._check_quality
 * (karios/api/core.py:500-506) is read off a coarse pass and a refinement. */
KR_API int kr_histogram(const void *img, int64_t pitch, int dtype, int w, int h, int lo, int shift,
                        int nbins, uint64_t *hist, void *stream);

/* np.count_nonzero(image) with image[mask == 0] = 0 when a mask is given
 * (karios/api/core.py:285-290) -> *d_count (device uint64). */
KR_API int kr_count_valid(const void *img, int64_t pitch, int dtype, int w, int h, const uint8_t *mask,
                          int64_t mask_pitch, uint64_t *d_count, void *stream);

/* Both of the above in ONE pass over the raster (the percentile histogram of _check_quality and
 * the valid-pixel count of analyze_accuracy read the same monitored raster). */
KR_API int kr_histogram_count(const void *img, int64_t pitch, int dtype, int w, int h, int lo, int shift,
                              int nbins, uint64_t *hist, const uint8_t *mask, int64_t mask_pitch,
                              uint64_t *d_count, void *stream);

/* image.array[y0.astype(int), x0.astype(int)] as float64 (NaN outside the raster):
 * _filter_by_dn_values (karios/api/core.py:687-728), DEM altitudes (:1050-1053). */
KR_API int kr_gather_points(const void *img, int64_t pitch, int dtype, int w, int h, const float *x0,
                            const float *y0, int n, double *out, void *stream);

/* Multi-GPU exchange record of one unit (a tile of a scene pair): written on the
 * device, stream-ordered after kr_match_tile, so that the match tables of a batch
 * can be collected with ONE fixed-size all_gather and no host synchronisation
 * (the statistics karios/accuracy_analysis/accuracy_statistics.py
 * derives from dx / dy start from these moments).  128 bytes. */
typedef struct {
    int32_t n_rows;                 /* rows of the unit (kr_stats.n_kept) */
    int32_t flags;                  /* bit 0: select_incomplete, bit 1: overflow */
    double n;                       /* n_rows as float64 */
    double sum_dx, sum_dy, sum_dx2, sum_dy2;
    double min_dx, min_dy, max_dx, max_dy;      /* +inf / -inf when n_rows == 0 */
    double reserved[6];
} kr_unit_header;
KR_API int kr_unit_header_write(kr_ctx *ctx, kr_rows rows, kr_unit_header *d_out, void *stream);

/* A host raster in ordinary (pageable) memory -- what GdalRasterImage.array / .read hand out
 * (karios/core/image.py:300-388) -- to device memory: several host threads stage chunks through
 * pinned buffers in parallel and overlap the DMA (one cudaMemcpy does this on one thread).
 * Returns when the host memory has been read; `stream` is ordered after the last chunk. */
KR_API int kr_upload_pageable(void *dst_device, const void *src_host, int64_t bytes, int device, void *stream);

/* Measurement hooks (no reference counterpart).  With profiling on, kr_match_tile
 * brackets its stages with CUDA events on the caller's stream; after the stream
 * has been synchronised kr_read_stage_ms returns the KR_NUM_STAGES durations
 * (ms) of the last call, in this order: min/max+mask, Laplacian mon, Laplacian
 * ref, corner response, candidate selection, NMS, corner sort, pyramids, LK
 * round trip, row compaction+sort, ZNCC, mutual information. */
#define KR_NUM_STAGES 12
KR_API int kr_set_profiling(kr_ctx *ctx, int on);
KR_API int kr_read_stage_ms(kr_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* KARIOS_B200_H */

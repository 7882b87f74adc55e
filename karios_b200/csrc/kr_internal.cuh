// kr_internal.cuh -- shared declarations of the sm_100a KLT matching library.
// Layout of the library:
//   kr_prep.cu     K1 min/max + auto mask, LUT build, K2 normalise + Laplacian
//   kr_corners.cu  K3 min-eigenvalue response + candidates, K4 selection (NMS)
//   kr_corner_fast.cu  K3 in two tiers: integer bounds everywhere, exact values where needed
//   kr_sort.cu     chunk bitonic sort + rank merge (u64 keys)
//   kr_lk.cu       K5 pyrDown, K6 pyramidal LK (forward + backward + back-check)
//   kr_zncc.cu     K7 ZNCC
//   kr_mi.cu       K8 mutual-information scores (32 x 32 joint histogram per match)
//   kr_api.cu      context, error handling, C ABI (include/karios_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/karios_b200.h"

#define KR_MAX_LEVELS 6
#define KR_SORT_CHUNK 8192          // keys per bitonic chunk (64 KB of shared memory)

// Device-resident scalars of one context (zeroed / initialised by kr_reset_stats).
struct KrDevStats {
    int32_t min_i[3], max_i[3];         // integer rasters: slot 0 = a (mon), 1 = b (ref), 2 = scratch
    uint32_t minf_enc[3], maxf_enc[3];  // float rasters, order-preserving encoding
    unsigned long long valid;
    uint32_t eig_max_enc;               // order-preserving encoding of the masked max
    uint32_t n_cand, n_thr, n_sel, n_acc, n_corners, n_kept;
    uint32_t nms_rounds, overflow, select_incomplete;
    uint32_t thr_bits, cut_bits, hist_shift;
    uint32_t undecided[3];
    uint32_t barrier[2];
    uint32_t n_rowkeys;
    // two-tier corner response (kr_corner_fast.cu)
    uint32_t lmax_enc, umax_enc;        // bounds of the masked maximum, integer units
    uint32_t n_maxlist, n_exact;
    uint32_t cut_applied;               // the value cut-off dropped candidates above the threshold
    uint32_t fast_mode, fast_fallback;  // two-tier path in use / it cannot decide: re-run exactly
    uint32_t cut_est_bits;              // tier 1: running estimate of the value cut-off (0: none yet)
    uint32_t fa_rows;                   // tier 1: warp-rows whose candidates are in the estimate histogram
    uint32_t nms_pending[16];           // multi-launch NMS: candidates left undecided by round r
    uint32_t fa_skipped;                // tier 1: warp-rows ruled out as a whole by the running cut
    uint32_t fa_done;                   // tier 1: worker blocks that have finished
    uint32_t fa_rows_in;                // tier 1: warp-rows whose candidates have LANDED in the histogram
    uint32_t pad2[1];
};

struct kr_ctx {
    int device, num_sms;
    int max_w, max_h, max_corners;
    int64_t cand_cap;       // capacity of the candidate / key lists
    int64_t corner_cap;     // capacity of corner-sized arrays
    KrDevStats *d_stats;
    uint8_t *d_lut[3];      // 65536-entry uint8 tables (slot a, b, scratch)
    uint64_t *d_cand;       // K3 output: (float bits << 32) | (y*W + x)
    uint64_t *d_keys_a;     // selected keys / sort ping
    uint64_t *d_keys_b;     // accepted keys / sort pong
    uint64_t *d_maxlist;    // two-tier K3: pixels that may hold the masked maximum
    int64_t maxlist_cap;
    uint32_t *d_hist;       // 4096-bin value histogram of the candidates
    uint32_t *d_ghist;      // tier 1: 8192-bin histogram (float bits >> 18) of the first candidates seen
    uint32_t *d_xy;         // per selected candidate: x | y << 16
    uint8_t *d_state;       // NMS state per selected candidate
    int32_t *d_next;        // NMS cell lists: next pointer per candidate
    int32_t *d_cell_head;   // NMS cell lists: head per cell
    int64_t cell_cap;
    // planes owned by the context (pitch = plane_pitch bytes)
    int64_t plane_pitch;
    uint8_t *d_mask;        // auto mask
    uint8_t *d_lap[2];      // Laplacian planes: 0 = mon, 1 = ref
    uint8_t *d_pyr[2][KR_MAX_LEVELS];   // pyramid levels >= 1 of (prev, next)
    int64_t pyr_pitch[KR_MAX_LEVELS];
    // per-corner scratch
    float *d_p0, *d_p1;     // corners and forward-tracked points, [corner_cap][2]
    float *d_d;             // back-check distance
    uint8_t *d_keep;
    int nms_grid;           // co-resident grid of the persistent NMS kernel
    int force_select_all;   // sort every candidate above the threshold (no pre-selection)
    int no_fast_corners;    // always use the one-tier exact response kernel (kr_set_corner_mode)
    int last_dtype;         // dtype of the last min/max pass (for kr_read_stats)
    int prof_on;            // stage events enabled (kr_set_profiling)
    cudaEvent_t ev[KR_NUM_STAGES + 1];
};

// ---- error plumbing (kr_api.cu) ------------------------------------------
int kr_set_error(int code, const char *fmt, ...);
#define KR_CUDA(expr)                                                                  \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess)                                                         \
            return kr_set_error(KR_ERR_CUDA, "%s: %s (%s:%d)", #expr,                  \
                                cudaGetErrorString(_e), __FILE__, __LINE__);           \
    } while (0)
// every kernel launch of the library is followed by KR_LAUNCH_CHECK(): it also counts the launch
// (kr_launch_count, process-wide; bench.py reports the difference over its timed region)
void kr_note_launch(void);
#define KR_LAUNCH_CHECK()            \
    do {                             \
        kr_note_launch();            \
        KR_CUDA(cudaGetLastError()); \
    } while (0)
#define KR_TRY(expr)                 \
    do {                             \
        int _r = (expr);             \
        if (_r != KR_OK) return _r;  \
    } while (0)

// stage boundary i (see KR_NUM_STAGES in the header)
#define KR_MARK(ctx, i, s)                                              \
    do {                                                                \
        if ((ctx)->prof_on) KR_CUDA(cudaEventRecord((ctx)->ev[(i)], (s))); \
    } while (0)

// ---- small device helpers ------------------------------------------------
__host__ __device__ __forceinline__ int kr_reflect101(int p, int len)
{
    // cv::borderInterpolate(p, len, BORDER_REFLECT_101)
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

__device__ __forceinline__ uint32_t kr_f32_enc(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float kr_f32_dec_bits(uint32_t e, int)
{
    uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
#define KR_ENC_NEG_INF 0x007fffffu   /* kr_f32_enc(-inf) */

// Sobel derivatives (scaled) at one position from its 3x3 neighbourhood, with
// the exact rounding sequence of the cv2 build.
__device__ __forceinline__ void sobel_products(float p00, float p01, float p02, float p10, float p12,
                                               float p20, float p21, float p22, float s, bool tail,
                                               float &xx, float &xy, float &yy)
{
    const float s2 = 2.0f * s;
    float r0 = p02 - p00, r1 = p12 - p10, r2 = p22 - p20;           // exact (small integers)
    float dx = __fmaf_rn(s, r0 + r2, __fmul_rn(s2, r1));
    float t0, t2;
    if (!tail) {
        t0 = __fmaf_rn(s, p02, __fmaf_rn(s2, p01, __fmul_rn(s, p00)));
        t2 = __fmaf_rn(s, p22, __fmaf_rn(s2, p21, __fmul_rn(s, p20)));
    } else {
        t0 = __fadd_rn(__fadd_rn(__fmul_rn(s, p00), __fmul_rn(s2, p01)), __fmul_rn(s, p02));
        t2 = __fadd_rn(__fadd_rn(__fmul_rn(s, p20), __fmul_rn(s2, p21)), __fmul_rn(s, p22));
    }
    float dy = __fsub_rn(t2, t0);
    xx = __fmul_rn(dx, dx);
    xy = __fmul_rn(dx, dy);
    yy = __fmul_rn(dy, dy);
}

// calcMinEigenVal on the float64 box sums: plain float32, no contraction.
__device__ __forceinline__ float eig_from_sums(double sxx, double sxy, double syy)
{
    float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
    float t = __fsub_rn(a, c);
    return __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
}

// ---- stage launchers (each enqueues on `s`, returns a kr_status) -----------
int krl_reset_stats(kr_ctx *ctx, cudaStream_t s);
int krl_minmax_mask(kr_ctx *ctx, const void *a, int64_t pa, const void *b, int64_t pb, int dtype,
                    int w, int h, int has_nd_a, double nd_a, int has_nd_b, double nd_b,
                    uint8_t *mask, int64_t pm, cudaStream_t s);
int krl_minmax_single(kr_ctx *ctx, const void *img, int64_t pitch, int dtype, int w, int h, int slot,
                      cudaStream_t s);
int krl_laplacian(kr_ctx *ctx, const void *img, int64_t pitch, int dtype, int w, int h, int slot,
                  int ksize, int invert, uint8_t *out, int64_t out_pitch, cudaStream_t s);
int krl_good_features(kr_ctx *ctx, const uint8_t *img, int64_t pitch, const uint8_t *mask,
                      int64_t mask_pitch, int w, int h, int max_corners, double quality,
                      double min_distance, int block, int tail_mode, int select_all, float *eig_out,
                      int64_t eig_pitch, float *out_xy, int capacity, int32_t *d_count,
                      cudaStream_t s);
int kr_nms_occupancy(int *blocks_per_sm);
int krl_sort_u64(kr_ctx *ctx, uint64_t *keys, uint64_t *out, const uint32_t *d_n, int64_t cap,
                 int descending, cudaStream_t s);
int krl_pyr_down(const uint8_t *src, int64_t pitch, int w, int h, uint8_t *dst, int64_t dst_pitch,
                 cudaStream_t s);
struct KrLkArgs {
    const uint8_t *img[2][KR_MAX_LEVELS];   // [0] = prev pyramid, [1] = next pyramid
    int64_t pitch[2][KR_MAX_LEVELS];
    int w[KR_MAX_LEVELS], h[KR_MAX_LEVELS];
    int levels;                             // top level index (0 = no pyramid)
    int win, max_count;
    int use_cache;                          // J window cache in shared memory (set by lk_prepare)
    double eps2;
    float min_eig_thr;
};
int krl_build_pyramids(kr_ctx *ctx, const uint8_t *prev, int64_t pp, const uint8_t *next, int64_t np_,
                       int w, int h, int win, int max_level, KrLkArgs *args, cudaStream_t s);
int krl_pyramid_geometry(int w, int h, int64_t pitch, int win, int max_level, int *levels, int *wl,
                         int *hl, int64_t *pl);
int krl_pyramid_plane(const uint8_t *src, int levels, const int *wl, const int *hl, const int64_t *pl,
                      uint8_t *const *dst, cudaStream_t s);
int krl_lk_single(const KrLkArgs &a, const float *p0, int n, const int32_t *d_count, float *p1,
                  uint8_t *status, float *err, cudaStream_t s);
int krl_lk_roundtrip(const KrLkArgs &a, const float *p0, int n_cap, const uint32_t *d_count,
                     float back_thr, float *p1, float *dist, uint8_t *keep, cudaStream_t s);
int krl_emit_rows(kr_ctx *ctx, const float *p0, const float *p1, const float *dist,
                  const uint8_t *keep, int n_cap, const uint32_t *d_count, int sort_xy,
                  float back_thr, float x_off, float y_off, kr_rows rows, cudaStream_t s);
int krl_zncc(const void *ref, int64_t rp, int rw, int rh, const void *mon, int64_t mp, int mw, int mh,
             int dtype, const float *x0, const float *y0, const float *dx, const float *dy,
             const float *score, float min_score, int n, const uint32_t *d_count, double *out,
             cudaStream_t s);
int krl_mutual_info(const void *ref, int64_t rp, int rw, int rh, const void *mon, int64_t mp, int mw,
                    int mh, int dtype, const float *x0, const float *y0, const float *dx,
                    const float *dy, const float *score, float min_score, int n,
                    const uint32_t *d_count, double *out_studholme, double *out_nmi, cudaStream_t s);
int krl_eig_fast(kr_ctx *ctx, const uint8_t *img, int64_t pitch, const uint8_t *mask, int64_t mask_pitch,
                 int w, int h, float scale, int tail_start, uint32_t target, cudaStream_t s);
int krl_cand_hist_fast(kr_ctx *ctx, double quality, float scale, cudaStream_t s);
int krl_exact_cands(kr_ctx *ctx, const uint8_t *img, int64_t pitch, int w, int h, float scale,
                    int tail_start, double quality, int expected, const uint64_t *sel, uint64_t *keys,
                    cudaStream_t s);

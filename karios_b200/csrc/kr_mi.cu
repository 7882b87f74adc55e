// kr_mi.cu -- K8: per-match mutual-information scores over 57x57 chips.
//
// Replaces, with one histogram per match for both,
//   MutualInfoService.compute_mutual_info / _compute_mutual_info / _mutual_info
//     (karios/matcher/mutual_info_service.py:73-138, 32-63): Studholme's
//     (H(X) + H(Y)) / H(X,Y), NaN when H(X,Y) == 0;
//   ZNCCService.compute_mi / _compute_mi / _mutual_information
//     (karios/matcher/zncc_service.py:240-287, 129-151): 2 * MI / (H(X) + H(Y)),
//     NaN when H(X) + H(Y) == 0.
// Both build np.histogram2d(chip_ref, chip_mon, bins=32): per chip and per axis
// 33 edges np.linspace(min, max, 33) (min - 0.5 / max + 0.5 when min == max), bin
// = number of edges <= v, minus one, the last edge closed.  The chips are the
// full 57x57 windows around (int(x0), int(y0)) and (round(x0+dx), round(y0+dy)),
// with the border rule of the ZNCC service (NaN when a chip leaves the raster).
//
// Integer rasters: edges are the exact rationals min + k (max - min) / 32, so the
// bin is the integer quotient ((v - min) * 32) / (max - min), capped at 31.
// Float rasters: edges evaluated exactly like linspace (k * step, then + start,
// two roundings; last edge = max) and the quotient guess is corrected against
// them -- in float32 for MutualInfoService, which hands float32 chips to
// np.histogram2d (float32 min/max -> float32 linspace), and in float64 for
// compute_mi, which casts the chips to float64 first (zncc_service.py:134-135);
// so float rasters build two histograms.  Entropies come from the integer
// counts: H = ln n - sum(c ln c) / n.
//
// One warp per match: pass 1 min/max of both chips, pass 2 (the chips now sit in
// L1/L2) shared-memory joint histogram, pass 3 one histogram row per lane.
#include <math.h>
#include "kr_internal.cuh"

namespace {

constexpr int M_MARGIN = 28, M_SIDE = 57, M_N = M_SIDE * M_SIDE, M_BINS = 32;
constexpr int M_WARPS = 8;

template <typename T> struct MTraits { static constexpr bool is_float = false; };
template <> struct MTraits<float> { static constexpr bool is_float = true; };

template <typename T> __device__ __forceinline__ const T *row_ptr(const T *base, int64_t pitch, int r)
{
    return (const T *)((const char *)base + (int64_t)r * pitch);
}

__device__ __forceinline__ int wmin(int v)
{
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int wmax(int v)
{
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float wminf(float v)
{
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float wmaxf(float v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double wsumd(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// np.linspace(start, stop, 33)[k] in float64
__device__ __forceinline__ double edge_at(int k, double start, double stop, double step)
{
    return k >= M_BINS ? stop : __dadd_rn(__dmul_rn((double)k, step), start);
}

// searchsorted(edges, v, 'right') - 1 with the last edge closed
__device__ __forceinline__ int float_bin(double v, double start, double stop, double step)
{
    if (!(step > 0.0)) return M_BINS / 2;
    int k = (int)((v - start) / step);
    k = max(0, min(M_BINS - 1, k));
    while (k > 0 && v < edge_at(k, start, stop, step)) k--;
    while (k < M_BINS - 1 && v >= edge_at(k + 1, start, stop, step)) k++;
    return k;
}

// the same with float32 edges (np.linspace on float32 scalars stays float32)
__device__ __forceinline__ float edge_at32(int k, float start, float stop, float step)
{
    return k >= M_BINS ? stop : __fadd_rn(__fmul_rn((float)k, step), start);
}
__device__ __forceinline__ int float_bin32(float v, float start, float stop, float step)
{
    if (!(step > 0.f)) return M_BINS / 2;
    int k = (int)__fdiv_rn(__fsub_rn(v, start), step);
    k = max(0, min(M_BINS - 1, k));
    while (k > 0 && v < edge_at32(k, start, stop, step)) k--;
    while (k < M_BINS - 1 && v >= edge_at32(k + 1, start, stop, step)) k++;
    return k;
}

__device__ __forceinline__ double clogc(uint32_t c)
{
    return c > 1u ? (double)c * log((double)c) : 0.0;
}

// One joint histogram (a warp's shared-memory copy) -> both scores; lane = one row
// of the histogram (rotated columns: no bank conflicts).  Leaves NaN when all
// pixels share one joint bin: H(X,Y) == 0, and then also H(X) + H(Y) == 0.
__device__ __forceinline__ void scores_from_hist(const uint32_t *hist, int lane, double &r_st, double &r_mi)
{
    uint32_t rowsum = 0, colsum = 0, cmax = 0;
    double sxy = 0.0;
#pragma unroll 4
    for (int j = 0; j < M_BINS; j++) {
        const uint32_t cj = hist[lane * M_BINS + ((j + lane) & (M_BINS - 1))];
        rowsum += cj;
        cmax = max(cmax, cj);
        if (cj > 1u) sxy += (double)cj * log((double)cj);
        colsum += hist[j * M_BINS + lane];
    }
    const double sx = wsumd(clogc(rowsum)), sy = wsumd(clogc(colsum));
    sxy = wsumd(sxy);
    cmax = (uint32_t)wmax((int)cmax);
    if (cmax < (uint32_t)M_N) {
        const double ln_n = log((double)M_N), inv_n = 1.0 / (double)M_N;
        const double hx = ln_n - sx * inv_n, hy = ln_n - sy * inv_n, hxy = ln_n - sxy * inv_n;
        r_st = (hx + hy) / hxy;
        r_mi = 2.0 * (hx + hy - hxy) / (hx + hy);
    }
}

template <typename T>
__global__ void __launch_bounds__(32 * M_WARPS)
k_mutual_info(const T *__restrict__ ref, int64_t rp, int rw, int rh, const T *__restrict__ mon,
              int64_t mp, int mw, int mh, const float *__restrict__ x0, const float *__restrict__ y0,
              const float *__restrict__ dx, const float *__restrict__ dy,
              const float *__restrict__ score, float min_score, int n, const uint32_t *d_count,
              double *__restrict__ out_studholme, double *__restrict__ out_nmi)
{
    __shared__ uint32_t s_hist[M_WARPS][M_BINS * M_BINS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t *hist = s_hist[wib];
    int cnt = n;
    if (d_count) cnt = (int)min(*d_count, (uint32_t)n);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < cnt; i += warps) {
        const float fx = x0[i], fy = y0[i];
        bool ok = !(score && !(score[i] >= min_score));        // api/core.py:884
        const int ax = (int)fx, ay = (int)fy;
        const int bx = __float2int_rn(__fadd_rn(fx, dx[i]));
        const int by = __float2int_rn(__fadd_rn(fy, dy[i]));
        if (ax - M_MARGIN < 0 || ay - M_MARGIN < 0 || bx - M_MARGIN < 0 || by - M_MARGIN < 0) ok = false;
        if (ax >= rw - M_MARGIN || ay >= rh - M_MARGIN || bx >= mw - M_MARGIN || by >= mh - M_MARGIN)
            ok = false;
        double r_st = qnan, r_mi = qnan;
        if (ok) {
            const T *pa = row_ptr(ref, rp, ay - M_MARGIN) + (ax - M_MARGIN);
            const T *pb = row_ptr(mon, mp, by - M_MARGIN) + (bx - M_MARGIN);
#pragma unroll
            for (int j = 0; j < M_BINS; j++) hist[j * M_BINS + lane] = 0u;
            __syncwarp();
            if (!MTraits<T>::is_float) {
                int mna = INT_MAX, mxa = INT_MIN, mnb = INT_MAX, mxb = INT_MIN;
                int r = 0, c = lane;
                for (int p = lane; p < M_N; p += 32) {
                    const int a = (int)row_ptr(pa, rp, r)[c], b = (int)row_ptr(pb, mp, r)[c];
                    mna = min(mna, a); mxa = max(mxa, a);
                    mnb = min(mnb, b); mxb = max(mxb, b);
                    c += 32;
                    if (c >= M_SIDE) { c -= M_SIDE; r++; }
                }
                mna = wmin(mna); mxa = wmax(mxa); mnb = wmin(mnb); mxb = wmax(mxb);
                const int da = mxa - mna, db = mxb - mnb;
                r = 0; c = lane;
                for (int p = lane; p < M_N; p += 32) {
                    const int a = (int)row_ptr(pa, rp, r)[c], b = (int)row_ptr(pb, mp, r)[c];
                    const int ka = da ? min(M_BINS - 1, ((a - mna) * M_BINS) / da) : M_BINS / 2;
                    const int kb = db ? min(M_BINS - 1, ((b - mnb) * M_BINS) / db) : M_BINS / 2;
                    atomicAdd(&hist[ka * M_BINS + kb], 1u);
                    c += 32;
                    if (c >= M_SIDE) { c -= M_SIDE; r++; }
                }
            } else {
                float mna = INFINITY, mxa = -INFINITY, mnb = INFINITY, mxb = -INFINITY;
                bool fin = true;
                int r = 0, c = lane;
                for (int p = lane; p < M_N; p += 32) {
                    const float a = (float)row_ptr(pa, rp, r)[c], b = (float)row_ptr(pb, mp, r)[c];
                    fin = fin && isfinite(a) && isfinite(b);
                    mna = fminf(mna, a); mxa = fmaxf(mxa, a);
                    mnb = fminf(mnb, b); mxb = fmaxf(mxb, b);
                    c += 32;
                    if (c >= M_SIDE) { c -= M_SIDE; r++; }
                }
                ok = __all_sync(0xffffffffu, fin);              // np.histogram2d raises -> NaN
                mna = wminf(mna); mxa = wmaxf(mxa); mnb = wminf(mnb); mxb = wmaxf(mxb);
                if (ok) {
                    // float32 edges -> Studholme score
                    const float st32a = __fdiv_rn(__fsub_rn(mxa, mna), (float)M_BINS);
                    const float st32b = __fdiv_rn(__fsub_rn(mxb, mnb), (float)M_BINS);
                    r = 0; c = lane;
                    for (int p = lane; p < M_N; p += 32) {
                        const float a = (float)row_ptr(pa, rp, r)[c], b = (float)row_ptr(pb, mp, r)[c];
                        const int ka = float_bin32(a, mna, mxa, st32a), kb = float_bin32(b, mnb, mxb, st32b);
                        atomicAdd(&hist[ka * M_BINS + kb], 1u);
                        c += 32;
                        if (c >= M_SIDE) { c -= M_SIDE; r++; }
                    }
                    __syncwarp();
                    double unused = qnan;
                    scores_from_hist(hist, lane, r_st, unused);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < M_BINS; j++) hist[j * M_BINS + lane] = 0u;
                    __syncwarp();
                    // float64 edges -> compute_mi score (histogram consumed below)
                    const double sa = (double)mna, ea = (double)mxa, sb = (double)mnb, eb = (double)mxb;
                    const double stepa = (ea - sa) / M_BINS, stepb = (eb - sb) / M_BINS;
                    r = 0; c = lane;
                    for (int p = lane; p < M_N; p += 32) {
                        const double a = (double)row_ptr(pa, rp, r)[c], b = (double)row_ptr(pb, mp, r)[c];
                        const int ka = float_bin(a, sa, ea, stepa), kb = float_bin(b, sb, eb, stepb);
                        atomicAdd(&hist[ka * M_BINS + kb], 1u);
                        c += 32;
                        if (c >= M_SIDE) { c -= M_SIDE; r++; }
                    }
                }
            }
            __syncwarp();
            if (ok) {
                if (MTraits<T>::is_float) {
                    double unused = qnan;
                    scores_from_hist(hist, lane, unused, r_mi);
                } else {
                    scores_from_hist(hist, lane, r_st, r_mi);
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (out_studholme) out_studholme[i] = r_st;
            if (out_nmi) out_nmi[i] = r_mi;
        }
    }
}

template <typename T>
int launch(const void *ref, int64_t rp, int rw, int rh, const void *mon, int64_t mp, int mw, int mh,
           const float *x0, const float *y0, const float *dx, const float *dy, const float *score,
           float min_score, int n, const uint32_t *d_count, double *o1, double *o2, cudaStream_t s)
{
    int grid = (n + M_WARPS - 1) / M_WARPS;
    if (grid > 148 * 6) grid = 148 * 6;
    k_mutual_info<T><<<grid, 32 * M_WARPS, 0, s>>>((const T *)ref, rp, rw, rh, (const T *)mon, mp, mw, mh,
                                                   x0, y0, dx, dy, score, min_score, n, d_count, o1, o2);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

}  // namespace

int krl_mutual_info(const void *ref, int64_t rp, int rw, int rh, const void *mon, int64_t mp, int mw,
                    int mh, int dtype, const float *x0, const float *y0, const float *dx,
                    const float *dy, const float *score, float min_score, int n,
                    const uint32_t *d_count, double *out_studholme, double *out_nmi, cudaStream_t s)
{
    if (n <= 0) return KR_OK;
    switch (dtype) {
    case KR_U8: return launch<uint8_t>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out_studholme, out_nmi, s);
    case KR_U16: return launch<uint16_t>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out_studholme, out_nmi, s);
    case KR_I16: return launch<int16_t>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out_studholme, out_nmi, s);
    case KR_F32: return launch<float>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out_studholme, out_nmi, s);
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    }
}

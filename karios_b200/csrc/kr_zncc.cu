// kr_zncc.cu -- K7: per-match zero-mean normalised cross-correlation.
//
// Replaces ZNCCService.compute_zncc / _compute_zncc / _extract_chip / _zncc2
// (karios/matcher/zncc_service.py:162-238, 289-297, 45-126): reference chip
// centred at (int(x0), int(y0)), monitored chip at (round(x0+dx), round(y0+dy))
// -- float32 sum, round half to even --, 57x57 chips of which the central 43x43
// window is correlated with population statistics in float64; NaN when a chip
// leaves the raster or a window has zero variance.
//
// Integer rasters: exact int64 moments, converted once.
// One warp per row; each lane strides the 1849 window pixels.
#include <math.h>
#include "kr_internal.cuh"

namespace {

constexpr int Z_MARGIN = 28, Z_HALF = 21, Z_SIDE = 2 * Z_HALF + 1, Z_N = Z_SIDE * Z_SIDE;

template <typename T> struct ZTraits { static constexpr bool is_float = false; };
template <> struct ZTraits<float> { static constexpr bool is_float = true; };

__device__ __forceinline__ long long wsum(long long v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double wsum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_zncc(const T *__restrict__ ref, int64_t rp, int rw, int rh, const T *__restrict__ mon, int64_t mp,
       int mw, int mh, const float *__restrict__ x0, const float *__restrict__ y0,
       const float *__restrict__ dx, const float *__restrict__ dy, const float *__restrict__ score,
       float min_score, int n, const uint32_t *d_count, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    int cnt = n;
    if (d_count) cnt = (int)min(*d_count, (uint32_t)n);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < cnt; i += warps) {
        const float fx = x0[i], fy = y0[i];
        double res = qnan;
        bool ok = !(score && !(score[i] >= min_score));        // api/core.py:884
        const int ax = (int)fx, ay = (int)fy;                   // int(series["x0"])
        const int bx = __float2int_rn(__fadd_rn(fx, dx[i]));    // round(x0 + dx), half to even
        const int by = __float2int_rn(__fadd_rn(fy, dy[i]));
        if (ax - Z_MARGIN < 0 || ay - Z_MARGIN < 0 || bx - Z_MARGIN < 0 || by - Z_MARGIN < 0) ok = false;
        if (ax >= rw - Z_MARGIN || ay >= rh - Z_MARGIN || bx >= mw - Z_MARGIN || by >= mh - Z_MARGIN)
            ok = false;
        if (ok) {
            const T *pa = (const T *)((const char *)ref + (int64_t)(ay - Z_HALF) * rp) + (ax - Z_HALF);
            const T *pb = (const T *)((const char *)mon + (int64_t)(by - Z_HALF) * mp) + (bx - Z_HALF);
            if (!ZTraits<T>::is_float) {
                long long sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
                int r = 0, c = lane;
                if (c >= Z_SIDE) { c -= Z_SIDE; r++; }          // never for 43 > 32
                for (int p = lane; p < Z_N; p += 32) {
                    long long a = (long long)*((const T *)((const char *)pa + (int64_t)r * rp) + c);
                    long long b = (long long)*((const T *)((const char *)pb + (int64_t)r * mp) + c);
                    sa += a; sb += b; saa += a * a; sbb += b * b; sab += a * b;
                    c += 32;
                    if (c >= Z_SIDE) { c -= Z_SIDE; r++; }
                }
                sa = wsum(sa); sb = wsum(sb); saa = wsum(saa); sbb = wsum(sbb); sab = wsum(sab);
                const long long N = Z_N;
                const long long va = N * saa - sa * sa, vb = N * sbb - sb * sb;
                if (va != 0 && vb != 0)
                    res = __ddiv_rn((double)(N * sab - sa * sb),
                                    __dmul_rn(__dsqrt_rn((double)va), __dsqrt_rn((double)vb)));
            } else {
                // float rasters: two-pass float64 (mean, then centred moments)
                double sa = 0, sb = 0;
                int r = 0, c = lane;
                for (int p = lane; p < Z_N; p += 32) {
                    sa += (double)*((const T *)((const char *)pa + (int64_t)r * rp) + c);
                    sb += (double)*((const T *)((const char *)pb + (int64_t)r * mp) + c);
                    c += 32;
                    if (c >= Z_SIDE) { c -= Z_SIDE; r++; }
                }
                const double ma = wsum(sa) / Z_N, mb = wsum(sb) / Z_N;
                double caa = 0, cbb = 0, cab = 0;
                r = 0; c = lane;
                for (int p = lane; p < Z_N; p += 32) {
                    double a = (double)*((const T *)((const char *)pa + (int64_t)r * rp) + c) - ma;
                    double b = (double)*((const T *)((const char *)pb + (int64_t)r * mp) + c) - mb;
                    caa += a * a; cbb += b * b; cab += a * b;
                    c += 32;
                    if (c >= Z_SIDE) { c -= Z_SIDE; r++; }
                }
                caa = wsum(caa); cbb = wsum(cbb); cab = wsum(cab);
                const double sda = sqrt(caa / Z_N), sdb = sqrt(cbb / Z_N);
                if (sda != 0.0 && sdb != 0.0) res = (cab / Z_N) / (sda * sdb);
            }
        }
        if (lane == 0) out[i] = res;
    }
}

template <typename T>
int launch(const void *ref, int64_t rp, int rw, int rh, const void *mon, int64_t mp, int mw, int mh,
           const float *x0, const float *y0, const float *dx, const float *dy, const float *score,
           float min_score, int n, const uint32_t *d_count, double *out, cudaStream_t s)
{
    int grid = (n + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    k_zncc<T><<<grid, 256, 0, s>>>((const T *)ref, rp, rw, rh, (const T *)mon, mp, mw, mh, x0, y0, dx,
                                   dy, score, min_score, n, d_count, out);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

}  // namespace

int krl_zncc(const void *ref, int64_t rp, int rw, int rh, const void *mon, int64_t mp, int mw, int mh,
             int dtype, const float *x0, const float *y0, const float *dx, const float *dy,
             const float *score, float min_score, int n, const uint32_t *d_count, double *out,
             cudaStream_t s)
{
    if (n <= 0) return KR_OK;
    switch (dtype) {
    case KR_U8: return launch<uint8_t>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out, s);
    case KR_U16: return launch<uint16_t>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out, s);
    case KR_I16: return launch<int16_t>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out, s);
    case KR_F32: return launch<float>(ref, rp, rw, rh, mon, mp, mw, mh, x0, y0, dx, dy, score, min_score, n, d_count, out, s);
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    }
}

// kr_scene.cu -- full-frame passes either side of the matching path.
//
//   kr_cross_power / kr_argmax_abs  the element-wise steps of the whole-pixel
//       skimage.registration.phase_cross_correlation(mon, ref) that
//       LargeOffsetMatcher.match calls (karios/matcher/large_offset.py:32-41): cross-power
//       spectrum with "phase" normalisation, then the first maximum of |correlation|
//       (the FFTs themselves are cuFFT through torch.fft);
//   kr_shift_image   shift_image (karios/core/image.py:70-101);
//   kr_histogram     value histogram of an integer raster: the 2 / 98 percentiles of
//       _check_quality (karios/api/core.py:500-506) are read off it on the host;
//   kr_count_valid   np.count_nonzero of the (masked) monitored image (api/core.py:285-290);
//   kr_gather_points raster values at int(y0), int(x0): _filter_by_dn_values
//       (api/core.py:687-728) and the DEM altitude lookup (api/core.py:1050-1053).
#include <float.h>
#include <math.h>
#include "kr_internal.cuh"

namespace {

// a <- a * conj(b) / max(|a * conj(b)|, 100 eps)    (interleaved complex)
template <typename F>
__global__ void __launch_bounds__(256)
k_cross_power(F *__restrict__ a, const F *__restrict__ b, int64_t n, F floor_)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const F ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
        const F pr = ar * br + ai * bi, pi = ai * br - ar * bi;
        F m = (F)hypot((double)pr, (double)pi);
        if (m < floor_) m = floor_;
        a[2 * i] = pr / m;
        a[2 * i + 1] = pi / m;
    }
}

// first index of the maximum of |x| (np.argmax(np.abs(x))): key = (|x| bits, ~index)
template <typename F> __device__ __forceinline__ unsigned long long abs_bits(F v);
template <> __device__ __forceinline__ unsigned long long abs_bits<float>(float v)
{
    return (unsigned long long)(__float_as_uint(v) & 0x7fffffffu);
}
template <> __device__ __forceinline__ unsigned long long abs_bits<double>(double v)
{
    return (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
}

template <typename F>
__global__ void __launch_bounds__(256)
k_argmax_abs(const F *__restrict__ x, int64_t n, unsigned long long *best_val, unsigned long long *best_idx_inv)
{
    unsigned long long bv = 0, bi = 0;             // bi = ~index: larger = earlier
    bool any = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const F v = x[i];
        if (v != v) continue;                      // np.argmax would stop at a NaN; inputs are finite
        const unsigned long long a = abs_bits<F>(v), inv = ~(unsigned long long)i;
        if (!any || a > bv || (a == bv && inv > bi)) { bv = a; bi = inv; any = true; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, bv, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool oa = __shfl_xor_sync(0xffffffffu, (int)any, o);
        if (oa && (!any || ov > bv || (ov == bv && oi > bi))) { bv = ov; bi = oi; any = true; }
    }
    if ((threadIdx.x & 31) == 0 && any) {
        // two-word maximum: raise the value first; the index is settled in k_argmax_fix
        atomicMax(best_val, bv);
    }
    (void)best_idx_inv;
}

template <typename F>
__global__ void __launch_bounds__(256)
k_argmax_fix(const F *__restrict__ x, int64_t n, const unsigned long long *best_val,
             unsigned long long *best_idx_inv)
{
    const unsigned long long bv = *best_val;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const F v = x[i];
        if (v == v && abs_bits<F>(v) == bv) atomicMax(best_idx_inv, ~(unsigned long long)i);
    }
}

__global__ void k_argmax_finish(const unsigned long long *best_idx_inv, long long *out)
{
    *out = (long long)(~(*best_idx_inv));
}

template <typename T>
__global__ void __launch_bounds__(256)
k_shift_image(const T *__restrict__ src, int64_t sp, T *__restrict__ dst, int64_t dp, int w, int h,
              int x_off, int y_off)
{
    const int y = blockIdx.y;
    const int sy = y + y_off;
    T *drow = (T *)((char *)dst + (int64_t)y * dp);
    const bool row_in = sy >= 0 && sy < h;
    const T *srow = (const T *)((const char *)src + (int64_t)(row_in ? sy : 0) * sp);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) {
        const int sx = x + x_off;
        drow[x] = (row_in && sx >= 0 && sx < w) ? srow[sx] : (T)0;
    }
}

constexpr int HIST_SPAN = 8192;                    // bins per pass (32 KB of shared memory)

template <typename T>
__global__ void __launch_bounds__(256)
k_histogram(const T *__restrict__ img, int64_t pitch, int w, int h, int lo, int shift, int nbins,
            unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t sh[HIST_SPAN];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const T *row = (const T *)((const char *)img + (int64_t)y * pitch);
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            const int d = (int)row[x] - lo;
            const int b = d >> shift;
            if (d >= 0 && b < nbins) atomicAdd(&sh[b], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// One pass over an integer raster with 8-byte loads (rows 8-byte aligned): histogram of
// (value - lo) >> shift into shared memory and / or the count of non-zero pixels under a mask.
// Runs of equal bins inside a load are merged before the shared-memory atomic (neighbouring pixels
// of a smooth raster share their coarse bin), four loads are in flight per thread.
template <typename T, bool HIST, bool COUNT>
__global__ void __launch_bounds__(256)
k_scan8(const T *__restrict__ img, int64_t pitch, int w, int h, int lo, int shift, int nbins,
        unsigned long long *__restrict__ hist, const uint8_t *__restrict__ mask, int64_t mpitch, int mask_vec,
        unsigned long long *__restrict__ count)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    __shared__ uint32_t sh[HIST ? HIST_SPAN : 1];
    if (HIST) {
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0;
        __syncthreads();
    }
    const int nv = w / VEC;
    unsigned long long nz = 0;
    auto one = [&](uint2 raw, int x0, const uint8_t *mrow) {
        T px[VEC];
        memcpy(px, &raw, 8);
        if (COUNT) {
            uint8_t mk[VEC];
            if (mrow) {
                if (mask_vec) {
                    if (VEC == 4) { const uint32_t m = *reinterpret_cast<const uint32_t *>(mrow + x0); memcpy(mk, &m, 4); }
                    else { const uint2 m = *reinterpret_cast<const uint2 *>(mrow + x0); memcpy(mk, &m, 8); }
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; i++) mk[i] = mrow[x0 + i];
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; i++) nz += (px[i] != (T)0 && (!mrow || mk[i] != 0)) ? 1u : 0u;
        }
        if (HIST) {
            int cur = -1, cnt = 0;
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                const int d = (int)px[i] - lo;
                int b = d >> shift;
                if (d < 0 || b >= nbins) b = -1;
                if (b != cur) {
                    if (cur >= 0) atomicAdd(&sh[cur], (uint32_t)cnt);
                    cur = b; cnt = 0;
                }
                cnt++;
            }
            if (cur >= 0) atomicAdd(&sh[cur], (uint32_t)cnt);
        }
    };
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const char *rowb = (const char *)img + (int64_t)y * pitch;
        const uint2 *row = reinterpret_cast<const uint2 *>(rowb);
        const uint8_t *mrow = (COUNT && mask) ? mask + (int64_t)y * mpitch : nullptr;
        int v = threadIdx.x;
        for (; v + 3 * 256 < nv; v += 4 * 256) {
            const uint2 a = __ldg(row + v), b = __ldg(row + v + 256), c = __ldg(row + v + 512), d = __ldg(row + v + 768);
            one(a, v * VEC, mrow); one(b, (v + 256) * VEC, mrow); one(c, (v + 512) * VEC, mrow); one(d, (v + 768) * VEC, mrow);
        }
        for (; v < nv; v += 256) one(__ldg(row + v), v * VEC, mrow);
        // the last w % VEC pixels of the row
        const T *rowt = reinterpret_cast<const T *>(rowb);
        for (int x = nv * VEC + threadIdx.x; x < w; x += 256) {
            const T p = rowt[x];
            if (COUNT) nz += (p != (T)0 && (!mrow || mrow[x] != 0)) ? 1u : 0u;
            if (HIST) {
                const int d = (int)p - lo, b = d >> shift;
                if (d >= 0 && b < nbins) atomicAdd(&sh[b], 1u);
            }
        }
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
        if ((threadIdx.x & 31) == 0 && nz) atomicAdd(count, nz);
    }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += blockDim.x)
            if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
    }
}

template <typename T>
int launch_scan8(const void *img, int64_t pitch, int w, int h, int lo, int shift, int nbins, unsigned long long *hist,
                 const uint8_t *mask, int64_t mpitch, unsigned long long *count, cudaStream_t s)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    const int grid = h < 148 * 8 ? h : 148 * 8;
    const int mask_vec = mask && ((uintptr_t)mask % VEC == 0) && (mpitch % VEC == 0);
    const T *p = (const T *)img;
    if (hist && count) k_scan8<T, true, true><<<grid, 256, 0, s>>>(p, pitch, w, h, lo, shift, nbins, hist, mask, mpitch, mask_vec, count);
    else if (hist) k_scan8<T, true, false><<<grid, 256, 0, s>>>(p, pitch, w, h, lo, shift, nbins, hist, mask, mpitch, mask_vec, count);
    else k_scan8<T, false, true><<<grid, 256, 0, s>>>(p, pitch, w, h, lo, shift, nbins, hist, mask, mpitch, mask_vec, count);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_count_valid(const T *__restrict__ img, int64_t pitch, const uint8_t *__restrict__ mask, int64_t mpitch,
              int w, int h, unsigned long long *out)
{
    unsigned long long c = 0;
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const T *row = (const T *)((const char *)img + (int64_t)y * pitch);
        const uint8_t *mrow = mask ? mask + (int64_t)y * mpitch : nullptr;
        for (int x = threadIdx.x; x < w; x += blockDim.x)
            c += (row[x] != (T)0 && (!mrow || mrow[x] != 0)) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_gather_points(const T *__restrict__ img, int64_t pitch, int w, int h, const float *__restrict__ x0,
                const float *__restrict__ y0, int n, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)x0[i], y = (int)y0[i];                    // .astype(int): truncation
    double v = __longlong_as_double(0x7ff8000000000000ll);
    if (x >= 0 && x < w && y >= 0 && y < h) v = (double)((const T *)((const char *)img + (int64_t)y * pitch))[x];
    out[i] = v;
}

}  // namespace

extern "C" {

KR_API int kr_cross_power(void *a, const void *b, int64_t n, int is_double, void *stream)
{
    if (!a || !b || n < 0) return kr_set_error(KR_ERR_INVALID, "bad cross-power arguments");
    if (n == 0) return KR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (is_double)
        k_cross_power<double><<<grid, 256, 0, s>>>((double *)a, (const double *)b, n, 100.0 * DBL_EPSILON);
    else
        k_cross_power<float><<<grid, 256, 0, s>>>((float *)a, (const float *)b, n, 100.0f * FLT_EPSILON);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

KR_API int kr_argmax_abs(const void *x, int64_t n, int is_double, void *scratch16, int64_t *out_index,
                         void *stream)
{
    if (!x || n < 1 || !scratch16 || !out_index) return kr_set_error(KR_ERR_INVALID, "bad argmax arguments");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *bv = (unsigned long long *)scratch16, *bi = bv + 1;
    KR_CUDA(cudaMemsetAsync(scratch16, 0, 16, s));
    const int grid = 148 * 8;
    if (is_double) {
        k_argmax_abs<double><<<grid, 256, 0, s>>>((const double *)x, n, bv, bi);
        k_argmax_fix<double><<<grid, 256, 0, s>>>((const double *)x, n, bv, bi);
    } else {
        k_argmax_abs<float><<<grid, 256, 0, s>>>((const float *)x, n, bv, bi);
        k_argmax_fix<float><<<grid, 256, 0, s>>>((const float *)x, n, bv, bi);
    }
    k_argmax_finish<<<1, 1, 0, s>>>(bi, (long long *)out_index);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

KR_API int kr_shift_image(const void *src, int64_t src_pitch, void *dst, int64_t dst_pitch, int dtype,
                          int w, int h, int x_off, int y_off, void *stream)
{
    if (!src || !dst || w < 1 || h < 1 || src == dst) return kr_set_error(KR_ERR_INVALID, "bad shift arguments");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((w + 1023) / 1024 < 8 ? (w + 1023) / 1024 : 8, h);
#define KR_SHIFT(T) k_shift_image<T><<<grid, 256, 0, s>>>((const T *)src, src_pitch, (T *)dst, dst_pitch, w, h, x_off, y_off)
    switch (dtype) {
    case KR_U8: KR_SHIFT(uint8_t); break;
    case KR_U16: case KR_I16: KR_SHIFT(uint16_t); break;
    case KR_F32: KR_SHIFT(uint32_t); break;          // bit copy; zero fill = 0.0f
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    }
#undef KR_SHIFT
    KR_LAUNCH_CHECK();
    return KR_OK;
}

KR_API int kr_histogram(const void *img, int64_t pitch, int dtype, int w, int h, int lo, int shift,
                        int nbins, uint64_t *hist, void *stream)
{
    if (!img || !hist || w < 1 || h < 1 || nbins < 1 || nbins > HIST_SPAN || shift < 0 || shift > 16)
        return kr_set_error(KR_ERR_INVALID, "bad histogram arguments (1..%d bins per pass)", HIST_SPAN);
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = h < 148 * 4 ? h : 148 * 4;
    unsigned long long *hh = (unsigned long long *)hist;
    if ((uintptr_t)img % 8 == 0 && pitch % 8 == 0) {
        switch (dtype) {
        case KR_U8: return launch_scan8<uint8_t>(img, pitch, w, h, lo, shift, nbins, hh, nullptr, 0, nullptr, s);
        case KR_U16: return launch_scan8<uint16_t>(img, pitch, w, h, lo, shift, nbins, hh, nullptr, 0, nullptr, s);
        case KR_I16: return launch_scan8<int16_t>(img, pitch, w, h, lo, shift, nbins, hh, nullptr, 0, nullptr, s);
        default: return kr_set_error(KR_ERR_UNSUPPORTED, "histogram needs an integer raster (dtype %d)", dtype);
        }
    }
    switch (dtype) {
    case KR_U8: k_histogram<uint8_t><<<grid, 256, 0, s>>>((const uint8_t *)img, pitch, w, h, lo, shift, nbins, hh); break;
    case KR_U16: k_histogram<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)img, pitch, w, h, lo, shift, nbins, hh); break;
    case KR_I16: k_histogram<int16_t><<<grid, 256, 0, s>>>((const int16_t *)img, pitch, w, h, lo, shift, nbins, hh); break;
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "histogram needs an integer raster (dtype %d)", dtype);
    }
    KR_LAUNCH_CHECK();
    return KR_OK;
}

KR_API int kr_count_valid(const void *img, int64_t pitch, int dtype, int w, int h, const uint8_t *mask,
                          int64_t mask_pitch, uint64_t *d_count, void *stream)
{
    if (!img || !d_count || w < 1 || h < 1) return kr_set_error(KR_ERR_INVALID, "bad count arguments");
    cudaStream_t s = (cudaStream_t)stream;
    KR_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
    const int grid = h < 148 * 8 ? h : 148 * 8;
    unsigned long long *o = (unsigned long long *)d_count;
    if ((uintptr_t)img % 8 == 0 && pitch % 8 == 0 && dtype != KR_F32) {
        if (dtype == KR_U8) return launch_scan8<uint8_t>(img, pitch, w, h, 0, 0, 1, nullptr, mask, mask_pitch, o, s);
        return launch_scan8<uint16_t>(img, pitch, w, h, 0, 0, 1, nullptr, mask, mask_pitch, o, s);
    }
    switch (dtype) {
    case KR_U8: k_count_valid<uint8_t><<<grid, 256, 0, s>>>((const uint8_t *)img, pitch, mask, mask_pitch, w, h, o); break;
    case KR_U16: case KR_I16: k_count_valid<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)img, pitch, mask, mask_pitch, w, h, o); break;
    case KR_F32: k_count_valid<float><<<grid, 256, 0, s>>>((const float *)img, pitch, mask, mask_pitch, w, h, o); break;
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    }
    KR_LAUNCH_CHECK();
    return KR_OK;
}

KR_API int kr_histogram_count(const void *img, int64_t pitch, int dtype, int w, int h, int lo, int shift,
                              int nbins, uint64_t *hist, const uint8_t *mask, int64_t mask_pitch,
                              uint64_t *d_count, void *stream)
{
    if (!img || !hist || !d_count || w < 1 || h < 1 || nbins < 1 || nbins > HIST_SPAN || shift < 0 || shift > 16)
        return kr_set_error(KR_ERR_INVALID, "bad histogram arguments (1..%d bins per pass)", HIST_SPAN);
    cudaStream_t s = (cudaStream_t)stream;
    if ((uintptr_t)img % 8 == 0 && pitch % 8 == 0) {
        KR_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
        unsigned long long *hh = (unsigned long long *)hist, *o = (unsigned long long *)d_count;
        switch (dtype) {
        case KR_U8: return launch_scan8<uint8_t>(img, pitch, w, h, lo, shift, nbins, hh, mask, mask_pitch, o, s);
        case KR_U16: return launch_scan8<uint16_t>(img, pitch, w, h, lo, shift, nbins, hh, mask, mask_pitch, o, s);
        case KR_I16: return launch_scan8<int16_t>(img, pitch, w, h, lo, shift, nbins, hh, mask, mask_pitch, o, s);
        default: return kr_set_error(KR_ERR_UNSUPPORTED, "histogram needs an integer raster (dtype %d)", dtype);
        }
    }
    KR_TRY(kr_histogram(img, pitch, dtype, w, h, lo, shift, nbins, hist, stream));
    return kr_count_valid(img, pitch, dtype, w, h, mask, mask_pitch, d_count, stream);
}

KR_API int kr_gather_points(const void *img, int64_t pitch, int dtype, int w, int h, const float *x0,
                            const float *y0, int n, double *out, void *stream)
{
    if (!img || !x0 || !y0 || !out || n < 0) return kr_set_error(KR_ERR_INVALID, "bad gather arguments");
    if (n == 0) return KR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = (n + 255) / 256;
    switch (dtype) {
    case KR_U8: k_gather_points<uint8_t><<<grid, 256, 0, s>>>((const uint8_t *)img, pitch, w, h, x0, y0, n, out); break;
    case KR_U16: k_gather_points<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)img, pitch, w, h, x0, y0, n, out); break;
    case KR_I16: k_gather_points<int16_t><<<grid, 256, 0, s>>>((const int16_t *)img, pitch, w, h, x0, y0, n, out); break;
    case KR_F32: k_gather_points<float><<<grid, 256, 0, s>>>((const float *)img, pitch, w, h, x0, y0, n, out); break;
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    }
    KR_LAUNCH_CHECK();
    return KR_OK;
}

}  // extern "C"

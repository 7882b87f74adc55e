// kr_prep.cu -- K1 (min/max + auto mask + valid count), LUT build and K2
// (uint8 normalisation fused with the integer Laplacian).
//
// Reference call sites replaced:
//   karios/matcher/klt.py:42-49    _to_uint8 (np.nanmin/np.nanmax, float64 scale, truncation)
//   karios/matcher/klt.py:268-276  auto mask + valid pixel count
//   karios/matcher/klt.py:419      255 - uint8(mon)  (polarity)
//   karios/matcher/klt.py:433-434  cv2.Laplacian(u8, cv2.CV_8U, ksize)
#include <float.h>
#include <limits.h>
#include <stdlib.h>
#include <type_traits>
#include "kr_internal.cuh"

namespace {

template <typename T> struct PixTraits;
template <> struct PixTraits<uint8_t> { typedef int acc_t; static constexpr bool is_float = false; };
template <> struct PixTraits<uint16_t> { typedef int acc_t; static constexpr bool is_float = false; };
template <> struct PixTraits<int16_t> { typedef int acc_t; static constexpr bool is_float = false; };
template <> struct PixTraits<float> { typedef float acc_t; static constexpr bool is_float = true; };

template <typename T> struct alignas(sizeof(T) * 4) Vec4 { T v[4]; };

__device__ __forceinline__ int acc_min(int a, int b) { return min(a, b); }
__device__ __forceinline__ int acc_max(int a, int b) { return max(a, b); }
__device__ __forceinline__ float acc_min(float a, float b) { return fminf(a, b); }   // NaN-ignoring
__device__ __forceinline__ float acc_max(float a, float b) { return fmaxf(a, b); }

template <typename A> __device__ __forceinline__ A warp_min(A v)
{
    for (int o = 16; o > 0; o >>= 1) v = acc_min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename A> __device__ __forceinline__ A warp_max(A v)
{
    for (int o = 16; o > 0; o >>= 1) v = acc_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ void publish_minmax(KrDevStats *st, int slot, int mn, int mx)
{
    atomicMin(&st->min_i[slot], mn);
    atomicMax(&st->max_i[slot], mx);
}
__device__ __forceinline__ void publish_minmax(KrDevStats *st, int slot, float mn, float mx)
{
    // +inf / -inf are the identities when no finite value was seen by this block
    atomicMin(&st->minf_enc[slot], kr_f32_enc(mn));
    atomicMax(&st->maxf_enc[slot], kr_f32_enc(mx));
}

// K1.  One block walks rows blockIdx.x, blockIdx.x + gridDim.x, ...; threads
// stride the row with 4-pixel vector loads when every row is suitably aligned.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
k_minmax_mask(const T *__restrict__ a, int64_t pa, const T *__restrict__ b, int64_t pb, int w, int h,
              int slot_a, int slot_b, int has_nd_a, double nd_a, int has_nd_b, double nd_b,
              uint8_t *__restrict__ mask, int64_t pm, KrDevStats *st)
{
    typedef typename PixTraits<T>::acc_t A;
    const bool isf = PixTraits<T>::is_float;
    A mn_a = isf ? (A)INFINITY : (A)INT_MAX, mx_a = isf ? (A)-INFINITY : (A)INT_MIN;
    A mn_b = mn_a, mx_b = mx_a;
    unsigned cnt = 0;

    auto one = [&](T va, T vb, bool have_b) -> uint8_t {
        A xa = (A)va;
        mn_a = acc_min(mn_a, xa);
        mx_a = acc_max(mx_a, xa);
        bool ok = va != (T)0;
        if (isf) ok = ok && isfinite((float)va);
        if (has_nd_a) ok = ok && ((double)va != nd_a);
        if (have_b) {
            A xb = (A)vb;
            mn_b = acc_min(mn_b, xb);
            mx_b = acc_max(mx_b, xb);
            ok = ok && vb != (T)0;
            if (isf) ok = ok && isfinite((float)vb);
            if (has_nd_b) ok = ok && ((double)vb != nd_b);
        }
        return ok ? 1 : 0;
    };

    const bool have_b = b != nullptr;
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const T *ra = (const T *)((const char *)a + (int64_t)y * pa);
        const T *rb = have_b ? (const T *)((const char *)b + (int64_t)y * pb) : nullptr;
        uint8_t *rm = mask ? mask + (int64_t)y * pm : nullptr;
        int x_scalar = 0;
        if (VEC) {
            int nv = w >> 2;
            for (int i = threadIdx.x; i < nv; i += blockDim.x) {
                Vec4<T> qa = *reinterpret_cast<const Vec4<T> *>(ra + 4 * i);
                Vec4<T> qb = qa;
                if (have_b) qb = *reinterpret_cast<const Vec4<T> *>(rb + 4 * i);
                uchar4 m;
                m.x = one(qa.v[0], qb.v[0], have_b);
                m.y = one(qa.v[1], qb.v[1], have_b);
                m.z = one(qa.v[2], qb.v[2], have_b);
                m.w = one(qa.v[3], qb.v[3], have_b);
                cnt += m.x + m.y + m.z + m.w;
                if (rm) *reinterpret_cast<uchar4 *>(rm + 4 * i) = m;
            }
            x_scalar = nv << 2;
        }
        for (int x = x_scalar + threadIdx.x; x < w; x += blockDim.x) {
            uint8_t m = one(ra[x], have_b ? rb[x] : (T)0, have_b);
            cnt += m;
            if (rm) rm[x] = m;
        }
    }

    // block reduction: warp shuffles, then one atomic per block
    __shared__ A s_mn_a[8], s_mx_a[8], s_mn_b[8], s_mx_b[8];
    __shared__ unsigned s_cnt[8];
    mn_a = warp_min(mn_a); mx_a = warp_max(mx_a);
    mn_b = warp_min(mn_b); mx_b = warp_max(mx_b);
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        s_mn_a[wid] = mn_a; s_mx_a[wid] = mx_a; s_mn_b[wid] = mn_b; s_mx_b[wid] = mx_b;
        s_cnt[wid] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long c = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) {
            mn_a = acc_min(mn_a, s_mn_a[i]); mx_a = acc_max(mx_a, s_mx_a[i]);
            mn_b = acc_min(mn_b, s_mn_b[i]); mx_b = acc_max(mx_b, s_mx_b[i]);
            c += s_cnt[i];
        }
        publish_minmax(st, slot_a, mn_a, mx_a);
        if (have_b) publish_minmax(st, slot_b, mn_b, mx_b);
        if (mask) atomicAdd(&st->valid, c);
    }
}

// K1, lean form for the common case: two uint16 rasters, 8-byte aligned rows, no
// no-data value, auto mask wanted.  Packed 16-bit min / max (two pixels per
// instruction), zero test by product, 4 mask bytes per store, several
// independent 8-byte loads in flight per thread.
__global__ void __launch_bounds__(256)
k_minmax_mask_u16(const uint16_t *__restrict__ a, int64_t pa, const uint16_t *__restrict__ b, int64_t pb,
                  int w, int h, uint8_t *__restrict__ mask, int64_t pm, KrDevStats *st)
{
    uint32_t mna = 0xffffffffu, mxa = 0u, mnb = 0xffffffffu, mxb = 0u;     // packed 2 x u16
    unsigned cnt = 0;
    const int nv = w >> 2;
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const uint2 *ra = reinterpret_cast<const uint2 *>((const char *)a + (int64_t)y * pa);
        const uint2 *rb = reinterpret_cast<const uint2 *>((const char *)b + (int64_t)y * pb);
        uint32_t *rm = reinterpret_cast<uint32_t *>(mask + (int64_t)y * pm);
#pragma unroll 4
        for (int i = threadIdx.x; i < nv; i += 256) {
            const uint2 qa = __ldg(ra + i), qb = __ldg(rb + i);
            mna = __vminu2(mna, __vminu2(qa.x, qa.y)); mxa = __vmaxu2(mxa, __vmaxu2(qa.x, qa.y));
            mnb = __vminu2(mnb, __vminu2(qb.x, qb.y)); mxb = __vmaxu2(mxb, __vmaxu2(qb.x, qb.y));
            // pixel valid <=> a * b != 0 (16 x 16 bit products cannot wrap to zero)
            const uint32_t m0 = ((qa.x & 0xffffu) * (qb.x & 0xffffu)) != 0u;
            const uint32_t m1 = ((qa.x >> 16) * (qb.x >> 16)) != 0u;
            const uint32_t m2 = ((qa.y & 0xffffu) * (qb.y & 0xffffu)) != 0u;
            const uint32_t m3 = ((qa.y >> 16) * (qb.y >> 16)) != 0u;
            cnt += m0 + m1 + m2 + m3;
            rm[i] = m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
        }
        for (int x = (nv << 2) + threadIdx.x; x < w; x += 256) {           // row tail
            const uint32_t va = ((const uint16_t *)ra)[x], vb = ((const uint16_t *)rb)[x];
            mna = __vminu2(mna, va | 0xffff0000u); mxa = __vmaxu2(mxa, va);
            mnb = __vminu2(mnb, vb | 0xffff0000u); mxb = __vmaxu2(mxb, vb);
            const uint32_t m = (va * vb) != 0u;
            cnt += m;
            ((uint8_t *)rm)[x] = (uint8_t)m;
        }
    }
    int mn_a = (int)min(mna & 0xffffu, mna >> 16), mx_a = (int)max(mxa & 0xffffu, mxa >> 16);
    int mn_b = (int)min(mnb & 0xffffu, mnb >> 16), mx_b = (int)max(mxb & 0xffffu, mxb >> 16);
    mn_a = warp_min(mn_a); mx_a = warp_max(mx_a); mn_b = warp_min(mn_b); mx_b = warp_max(mx_b);
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __shared__ int s_v[4][8];
    __shared__ unsigned s_c[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_v[0][wid] = mn_a; s_v[1][wid] = mx_a; s_v[2][wid] = mn_b; s_v[3][wid] = mx_b; s_c[wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long c = 0;
        for (int i = 0; i < 8; i++) {
            mn_a = min(mn_a, s_v[0][i]); mx_a = max(mx_a, s_v[1][i]);
            mn_b = min(mn_b, s_v[2][i]); mx_b = max(mx_b, s_v[3][i]);
            c += s_c[i];
        }
        publish_minmax(st, 0, mn_a, mx_a);
        publish_minmax(st, 1, mn_b, mx_b);
        atomicAdd(&st->valid, c);
    }
}

__global__ void k_reset_stats(KrDevStats *st)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int i = 0; i < 3; i++) {
            st->min_i[i] = INT_MAX; st->max_i[i] = INT_MIN;
            st->minf_enc[i] = 0xffffffffu; st->maxf_enc[i] = 0u;
            st->undecided[i] = 0;
        }
        st->valid = 0ull;
        st->eig_max_enc = KR_ENC_NEG_INF;
        st->n_cand = st->n_thr = st->n_sel = st->n_acc = st->n_corners = st->n_kept = 0;
        st->nms_rounds = st->overflow = st->select_incomplete = 0;
        st->thr_bits = st->cut_bits = st->hist_shift = 0;
        st->barrier[0] = st->barrier[1] = 0;
        st->n_rowkeys = 0;
    }
}

__global__ void k_reset_slot(KrDevStats *st, int slot)
{
    st->min_i[slot] = INT_MAX; st->max_i[slot] = INT_MIN;
    st->minf_enc[slot] = 0xffffffffu; st->maxf_enc[slot] = 0u;
}

// 65 536-entry table of _to_uint8 for a 16-bit (or 8-bit) raster: entry = the
// raw bit pattern of the pixel.  float64: subtract, divide, multiply by 255
// (each correctly rounded, never fused), truncate -- klt.py:47-48.
__global__ void k_build_lut(const KrDevStats *st, int slot, int dtype, int invert, uint8_t *lut)
{
    int bits = blockIdx.x * blockDim.x + threadIdx.x;
    if (bits >= 65536) return;
    int r;
    if (dtype == KR_U8) {
        r = bits & 255;                                 // uint8 input: _to_uint8 is a no-op
    } else {
        int v = (dtype == KR_I16) ? (int)(int16_t)(uint16_t)bits : bits;
        int mn = st->min_i[slot], mx = st->max_i[slot];
        r = 0;
        if (mx > mn) {
            double q = __ddiv_rn(__dsub_rn((double)v, (double)mn), (double)(mx - mn));
            q = __dmul_rn(q, 255.0);
            if (q >= 0.0 && q < 256.0) r = (int)q;      // values outside [mn, mx] never occur
        }
    }
    if (invert) r = 255 - r;
    lut[bits] = (uint8_t)r;
}

// uint8 value of one raw pixel: table lookup (16/8-bit rasters, L1-resident
// 64 KB table) or the float32 expression NumPy evaluates for float32 rasters.
template <typename T>
__device__ __forceinline__ int to_u8(T v, const uint8_t *__restrict__ lut, float fmn, float frange,
                                     int invert)
{
    if (PixTraits<T>::is_float) {
        int r = 0;
        if (frange > 0.f) {
            float q = __fmul_rn(__fdiv_rn(__fsub_rn((float)v, fmn), frange), 255.0f);
            if (q >= 0.f && q < 256.f) r = (int)q;
        }
        return invert ? 255 - r : r;
    }
    return __ldg(lut + (uint16_t)v);       // invert is folded into the table
}

constexpr int LAP_WARPS = 8, LAP_ROWS = 64;

// K2.  cv2.Laplacian's k x k kernel factorises exactly (integers, no rounding):
//   d2 (x) s + s (x) d2 = B^(k-3)_x B^(k-3)_y M3,   B = [1,1],
//   M3 = [[2,0,2],[0,-8,0],[2,0,2]]  (= the k = 3 kernel),
// because d2 = [1,-2,1] * B^(k-3) and s = [1,2,1] * B^(k-3).  So the result is
// a binomial smoothing S of the uint8 image followed by
//   out(x,y) = 2 (E(y-1) + E(y+1)) - 8 S(x,y),   E(y) = S(x-1,y) + S(x+1,y).
// One warp walks down a strip of 32 - 2R columns (R = k/2): lane <-> column, one
// coalesced raw-pixel load per row, the horizontal [1,1] cascade through warp
// shuffles, the vertical cascade and the 3-row window in registers.  No shared
// memory, no block synchronisation; the normalisation is the L1-resident table.
template <int K, typename T>
__global__ void __launch_bounds__(LAP_WARPS * 32)
k_laplacian(const T *__restrict__ img, int64_t pitch, int w, int h, const uint8_t *__restrict__ lut,
            const KrDevStats *__restrict__ st, int slot, int invert, uint8_t *__restrict__ out,
            int64_t out_pitch)
{
    constexpr int R = (K <= 3) ? 1 : K / 2;
    constexpr int NS = (K <= 3) ? 0 : K - 3;          // [1,1] stages per dimension (even)
    constexpr int VALID = 32 - 2 * R;
    constexpr unsigned FULL = 0xffffffffu;

    float fmn = 0.f, frange = 0.f;
    if (PixTraits<T>::is_float) {
        float mn = kr_f32_dec_bits(st->minf_enc[slot], 0), mx = kr_f32_dec_bits(st->maxf_enc[slot], 0);
        fmn = mn;
        frange = (mx > mn) ? (float)((double)mx - (double)mn) : 0.f;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int xs = (blockIdx.x * LAP_WARPS + wid) * VALID;     // first output column of the strip
    if (xs >= w) return;                                        // whole warp leaves together
    const int ys = blockIdx.y * LAP_ROWS;
    const int ye = min(ys + LAP_ROWS, h);
    const int cx = xs + lane - R;
    const int col = kr_reflect101(cx, w);
    const bool store_lane = lane >= R && lane < 32 - R && cx < w;

    int vs[NS > 0 ? NS : 1];                  // vertical cascade state
#pragma unroll
    for (int i = 0; i < (NS > 0 ? NS : 1); i++) vs[i] = 0;
    int e_m2 = 0, e_m1 = 0, s_m1 = 0, s_m2 = 0;

    // software pipeline: the raw pixel is loaded two rows ahead and normalised one
    // row ahead, so the two dependent loads (pixel, table) are off the critical path
    auto load_raw = [&](int r) -> T {
        int tr = r;
        if ((unsigned)tr >= (unsigned)h) tr = kr_reflect101(r, h);
        return *((const T *)((const char *)img + (int64_t)tr * pitch) + col);
    };
    const int r_first = ys - R, r_end = ye + R;
    int v_next = to_u8<T>(load_raw(r_first), lut, fmn, frange, invert);
    T raw_next = load_raw(r_first + 1);
#pragma unroll 2
    for (int r = r_first; r < r_end; r++) {
        int v = v_next;
        v_next = to_u8<T>(raw_next, lut, fmn, frange, invert);
        raw_next = load_raw(r + 2);
        // horizontal binomial cascade, alternating direction to stay centred
#pragma unroll
        for (int i = 0; i < NS; i++) {
            int o = (i & 1) ? __shfl_up_sync(FULL, v, 1) : __shfl_down_sync(FULL, v, 1);
            v += o;
        }
        // vertical cascade: after NS stages the value is centred NS/2 rows back
#pragma unroll
        for (int i = 0; i < NS; i++) {
            int t = v + vs[i];
            vs[i] = v;
            v = t;
        }
        const int s_cur = v;                                  // S at row r - NS/2
        int e_cur;
        if (K == 1) e_cur = 0;
        else e_cur = __shfl_up_sync(FULL, s_cur, 1) + __shfl_down_sync(FULL, s_cur, 1);
        // output row m = (r - NS/2) - 1 = r - R
        const int m = r - R;
        if (m >= ys) {
            int acc;
            if (K == 1) {
                int e1 = __shfl_up_sync(FULL, s_m1, 1) + __shfl_down_sync(FULL, s_m1, 1);
                acc = s_m2 + s_cur + e1 - 4 * s_m1;
            } else {
                acc = 2 * (e_m2 + e_cur) - 8 * s_m1;
            }
            if (store_lane) out[(int64_t)m * out_pitch + cx] = (uint8_t)min(255, max(0, acc));
        }
        e_m2 = e_m1; e_m1 = e_cur;
        s_m2 = s_m1; s_m1 = s_cur;
    }
}

// K2, packed form for k = 3, 5, 7: two adjacent pixels per lane in one 32-bit
// register (16-bit fields: the smoothed value never exceeds 255 * 2^(2(k-3)) <=
// 65280, so packed adds cannot carry across fields).  A warp covers 64 columns;
// every cascade stage is one shuffle + one funnel shift + one add for two pixels.
template <typename T> struct alignas(sizeof(T) * 2) Vec2 { T a, b; };

// FAST: every column of the warp and every row it touches (incl. the prefetch
// distance) lies inside the image and the planes are aligned: vector loads /
// stores through running pointers, no border arithmetic at all.
template <int K, typename T, bool FAST>
__device__ __forceinline__ void lap2_body(const T *__restrict__ img, int64_t pitch, int w, int h,
                                          const uint8_t *__restrict__ lut, float fmn, float frange,
                                          int invert, uint8_t *__restrict__ out, int64_t out_pitch,
                                          int xs, int ys, int ye, int lane)
{
    constexpr int R = K / 2;                          // 1, 2, 3
    constexpr int NS = K - 3;                         // [1,1] stages per dimension
    constexpr int HL = (R + 1) / 2;                   // halo in lanes (2 columns each)
    constexpr int PF = 4;                             // rows of load prefetch
    constexpr unsigned FULL = 0xffffffffu;
    const int cx = xs + 2 * (lane - HL);              // even column of the lane
    const int c0 = FAST ? cx : kr_reflect101(cx, w), c1 = FAST ? cx + 1 : kr_reflect101(cx + 1, w);
    const bool lane_ok = lane >= HL && lane < 32 - HL;
    const bool st0 = lane_ok && cx < w, st1 = lane_ok && cx + 1 < w;
    const int r_first = ys - R, r_end = ye + R;

    const char *pl = (const char *)img + (int64_t)r_first * pitch + (int64_t)cx * sizeof(T);   // FAST only
    int r_load = r_first;
    auto load_next = [&]() -> Vec2<T> {
        Vec2<T> q;
        if (FAST) {
            q = *reinterpret_cast<const Vec2<T> *>(pl);
            pl += pitch;
        } else {
            int tr = r_load;
            if ((unsigned)tr >= (unsigned)h) tr = kr_reflect101(r_load, h);
            const T *row = (const T *)((const char *)img + (int64_t)tr * pitch);
            q.a = row[c0]; q.b = row[c1];
            r_load++;
        }
        return q;
    };
    auto norm2 = [&](const Vec2<T> &q) -> uint32_t {
        return (uint32_t)to_u8<T>(q.a, lut, fmn, frange, invert) |
               ((uint32_t)to_u8<T>(q.b, lut, fmn, frange, invert) << 16);
    };

    uint32_t vs[NS > 0 ? NS : 1];
#pragma unroll
    for (int i = 0; i < (NS > 0 ? NS : 1); i++) vs[i] = 0;
    int e0_m2 = 0, e1_m2 = 0, e0_m1 = 0, e1_m1 = 0, s0_m1 = 0, s1_m1 = 0;
    uint8_t *po = out + (int64_t)ys * out_pitch + cx;           // output row m = ys first

    Vec2<T> raw[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) raw[u] = load_next();
    uint32_t vn = norm2(raw[0]);                       // table lookups run one row ahead too
    for (int rb = r_first; rb < r_end; rb += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int r = rb + u;
            if (r >= r_end) break;
            uint32_t v = vn;
            vn = norm2(raw[(u + 1) % PF]);
            raw[u] = load_next();
            // horizontal [1,1] cascade on the packed pair, alternating direction
#pragma unroll
            for (int i = 0; i < NS; i++) {
                if (i & 1) {
                    const uint32_t pv = __shfl_up_sync(FULL, v, 1);
                    v += __funnelshift_r(pv, v, 16);          // (p(x-1), p(x))
                } else {
                    const uint32_t nv = __shfl_down_sync(FULL, v, 1);
                    v += __funnelshift_r(v, nv, 16);          // (p(x+1), p(x+2))
                }
            }
#pragma unroll
            for (int i = 0; i < NS; i++) {                    // vertical cascade
                const uint32_t t = v + vs[i];
                vs[i] = v;
                v = t;
            }
            // S (packed) at row r - NS/2; unpack, E(x) = S(x-1) + S(x+1)
            const uint32_t sp = __shfl_up_sync(FULL, v, 1), sn = __shfl_down_sync(FULL, v, 1);
            const int s0 = (int)(v & 0xffffu), s1 = (int)(v >> 16);
            const int e0 = (int)(sp >> 16) + s1, e1 = s0 + (int)(sn & 0xffffu);
            if (r - R >= ys) {
                const int a0 = 2 * (e0_m2 + e0) - 8 * s0_m1, a1 = 2 * (e1_m2 + e1) - 8 * s1_m1;
                const int o0 = min(255, max(0, a0)), o1 = min(255, max(0, a1));
                if (FAST) {
                    if (lane_ok) *reinterpret_cast<uint16_t *>(po) = (uint16_t)(o0 | (o1 << 8));
                } else {
                    if (st0) po[0] = (uint8_t)o0;
                    if (st1) po[1] = (uint8_t)o1;
                }
                po += out_pitch;
            }
            e0_m2 = e0_m1; e1_m2 = e1_m1; e0_m1 = e0; e1_m1 = e1;
            s0_m1 = s0; s1_m1 = s1;
        }
    }
}

template <int K, typename T>
__global__ void __launch_bounds__(LAP_WARPS * 32)
k_laplacian2(const T *__restrict__ img, int64_t pitch, int w, int h, const uint8_t *__restrict__ lut,
             const KrDevStats *__restrict__ st, int slot, int invert, uint8_t *__restrict__ out,
             int64_t out_pitch, int aligned)
{
    constexpr int R = K / 2, HL = (R + 1) / 2, VALID = 64 - 4 * HL;
    float fmn = 0.f, frange = 0.f;
    if (PixTraits<T>::is_float) {
        float mn = kr_f32_dec_bits(st->minf_enc[slot], 0), mx = kr_f32_dec_bits(st->maxf_enc[slot], 0);
        fmn = mn;
        frange = (mx > mn) ? (float)((double)mx - (double)mn) : 0.f;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int xs = (blockIdx.x * LAP_WARPS + wid) * VALID;
    if (xs >= w) return;
    const int ys = blockIdx.y * LAP_ROWS, ye = min(ys + LAP_ROWS, h);
    const bool fast = aligned && (xs - 2 * HL >= 0) && (xs - 2 * HL + 64 <= w) && (ys - R >= 0) &&
                      (ye + R + 4 <= h);
    if (fast)
        lap2_body<K, T, true>(img, pitch, w, h, lut, fmn, frange, invert, out, out_pitch, xs, ys, ye, lane);
    else
        lap2_body<K, T, false>(img, pitch, w, h, lut, fmn, frange, invert, out, out_pitch, xs, ys, ye, lane);
}

// K2, four pixels per lane (k = 3, 5, 7; 8/16-bit rasters with vector-aligned rows,
// w % 4 == 0, w >= 256, h >= 64): a warp covers 128 columns (120 outputs), a lane holds
// its pixels as two packed 16-bit pairs A = (p0, p1), B = (p2, p3).  One cascade stage
// is ONE shuffle and two funnel-shift-adds for four pixels, the 3 x 3 step keeps the
// running partial E(m-1) - 4 S(m) per pixel, and the clamp + byte pack of four results
// is two cvt.pack.sat (I2IP) instructions.  The normalisation table of the raster's
// value range sits in shared memory when that range is below 32 K values (one LDS per
// pixel instead of a 64-bit address and a global load).  The image border needs no
// scalar path: the first warp starts 4 columns left of the image and its lane 0 takes
// lane 1's word with the pixels mirrored (REFLECT_101), the last warp is shifted left to
// end 4 columns right of the image and its lane 31 mirrors lane 30's word; rows are
// reflected by index in the top / bottom segments only (ROWFAST elsewhere).
constexpr int L4_WARPS = 4, L4_VALID = 120, L4_PF = 4;

struct L4Raw { uint32_t x, y; };
template <bool V> struct L4Tag { static constexpr bool value = V; };

// ARITH: _to_uint8 of a 16-bit integer raster in integer arithmetic instead of the table.  The
// reference's float64 expression ((v - mn) / (mx - mn) * 255).astype(uint8) (klt.py:47-48) equals
// floor((v - mn) * 255 / R), R = mx - mn, for every 16-bit range: where the quotient is an exact
// integer k the two roundings give fl(fl(k / 255) * 255) = k for all k <= 255, elsewhere it is at
// least 1 / R >= 1.5e-5 away from an integer, against 6e-14 of float64 error (checked for every R
// and every value, tests/test_oracle.py).  And floor(n * 255 / R) = (n * M) >> 32 with
// M = ceil(2^32 * 255 / R) for n <= R <= 65535 when R >= 256 (M < 2^32; the excess n * (M R -
// 255 * 2^32) stays below 2^32 / R of a step): ONE multiply-high per pixel, no table, no shared
// memory, no bank conflicts (the table look-ups had the kernel at 62 % of the shared-memory pipe).
template <int K, typename T, bool ROWFAST, bool ARITH, bool EDGE>
__device__ __forceinline__ void lap4_body(const T *__restrict__ img, int64_t pitch, int h,
                                          const uint8_t *__restrict__ lut, int mn,
                                          uint32_t base2, uint32_t magic, int invert,
                                          uint8_t *__restrict__ out, int64_t out_pitch, int lc, int oc,
                                          int edge, bool warp_edge, bool store_lane, int ys, int ye)
{
    constexpr int R = K / 2;                          // 1, 2, 3
    constexpr int NS = K - 3;                         // [1,1] stages per dimension: 0, 2, 4
    constexpr int PF = L4_PF;
    constexpr bool U8 = sizeof(T) == 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int r_first = ys - R;
    const int n_main = ye - ys;

    // ---- raw rows, PF ahead ----------------------------------------------------
    const char *pl = (const char *)img + (int64_t)r_first * pitch + (int64_t)lc * sizeof(T);   // ROWFAST
    const char *pc = (const char *)img + (int64_t)lc * sizeof(T);
    int r_load = r_first;
    auto load_next = [&]() -> L4Raw {
        const char *p;
        if (ROWFAST) {
            p = pl;
            pl += pitch;
        } else {
            int tr = r_load;
            if (tr < 0) tr = -tr;
            if (tr >= h) tr = 2 * (h - 1) - tr;             // h >= 64: one reflection is enough
            r_load++;
            p = pc + (int64_t)tr * pitch;
        }
        L4Raw q;
        if (U8) {
            q.x = __ldg(reinterpret_cast<const uint32_t *>(p));
            q.y = 0u;
            if (EDGE && warp_edge) {
                if (edge == 1) q.x = __byte_perm(q.x, q.x, 0x1233);       // (., b3, b2, b1)
                if (edge == 2) q.x = __byte_perm(q.x, q.x, 0x0012);       // (b2, b1, b0, .)
            }
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
            q.x = v.x; q.y = v.y;
            if (EDGE && warp_edge) {
                const uint32_t m = __byte_perm(v.y, v.x, 0x7610);          // (p2, p1)
                if (edge == 1) { q.x = v.y; q.y = m; }                     // (., p3), (p2, p1)
                if (edge == 2) { q.x = m; q.y = v.x; }                     // (p2, p1), (p0, .)
            }
        }
        return q;
    };
    // ---- _to_uint8 of a raw row -> packed pairs ---------------------------------
    auto lut2 = [&](uint32_t pr) -> uint32_t {
        uint32_t t0, t1;
        if (ARITH) {
            uint32_t n0, n1;
            if ((T)-1 > (T)0) {                             // unsigned: every field >= mn, no borrow
                const uint32_t d = pr - base2;
                n0 = d & 0xffffu;
                n1 = d >> 16;
            } else {
                n0 = (uint32_t)((int)(int16_t)(pr & 0xffffu) - mn);
                n1 = (uint32_t)(((int)pr >> 16) - mn);
            }
            const uint32_t p = __byte_perm(__umulhi(n0, magic), __umulhi(n1, magic), 0x5410);
            return invert ? 0x00ff00ffu - p : p;
        } else {
            t0 = __ldg(lut + (pr & 0xffffu));
            t1 = __ldg(lut + (pr >> 16));
        }
        return __byte_perm(t0, t1, 0x5410);
    };
    auto norm = [&](const L4Raw &q, uint32_t &A, uint32_t &B) {
        if (U8) {
            const uint32_t v = invert ? ~q.x : q.x;
            A = __byte_perm(v, 0u, 0x4140);
            B = __byte_perm(v, 0u, 0x4342);
        } else {
            A = lut2(q.x);
            B = lut2(q.y);
        }
    };

    uint32_t vsA[NS > 0 ? NS : 1], vsB[NS > 0 ? NS : 1];
#pragma unroll
    for (int i = 0; i < (NS > 0 ? NS : 1); i++) vsA[i] = vsB[i] = 0u;
    int part[4] = {0, 0, 0, 0}, eprev[4] = {0, 0, 0, 0};
    uint8_t *po = out + (int64_t)ys * out_pitch + oc;

    L4Raw raw[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) raw[u] = load_next();
    uint32_t nA, nB;
    norm(raw[0], nA, nB);

    // one input row; ph = its (compile-time) slot in the prefetch ring
    auto step = [&](auto store_tag, int ph) {
        constexpr bool STORE = decltype(store_tag)::value;
        uint32_t A = nA, B = nB;
        norm(raw[(ph + 1) % PF], nA, nB);
        raw[ph % PF] = load_next();
#pragma unroll
        for (int i = 0; i < NS; i++) {
            if (i & 1) {                                                   // + (p(x-1), p(x))
                const uint32_t pb = __shfl_up_sync(FULL, B, 1);
                const uint32_t tA = A + __funnelshift_r(pb, A, 16);
                B += __funnelshift_r(A, B, 16);
                A = tA;
            } else {                                                       // + (p(x+1), p(x+2))
                const uint32_t na = __shfl_down_sync(FULL, A, 1);
                const uint32_t tA = A + __funnelshift_r(A, B, 16);
                B += __funnelshift_r(B, na, 16);
                A = tA;
            }
        }
#pragma unroll
        for (int i = 0; i < NS; i++) {                                     // vertical cascade
            const uint32_t tA = A + vsA[i], tB = B + vsB[i];
            vsA[i] = A; vsB[i] = B;
            A = tA; B = tB;
        }
        // S (packed) of row c = r - NS/2; E(x) = S(x-1) + S(x+1)
        const uint32_t sp = __shfl_up_sync(FULL, B, 1), sn = __shfl_down_sync(FULL, A, 1);
        const int s0 = (int)(A & 0xffffu), s1 = (int)(A >> 16), s2 = (int)(B & 0xffffu), s3 = (int)(B >> 16);
        int e[4];
        e[0] = (int)(sp >> 16) + s1;
        e[1] = s0 + s2;
        e[2] = s1 + s3;
        e[3] = s2 + (int)(sn & 0xffffu);
        if (STORE) {
            // out(c-1) = 2 (E(c-2) + E(c)) - 8 S(c-1), saturated to uint8
            const int o0 = 2 * (part[0] + e[0]), o1 = 2 * (part[1] + e[1]), o2 = 2 * (part[2] + e[2]),
                      o3 = 2 * (part[3] + e[3]);
            uint32_t hi, wv;
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(o3), "r"(o2), "r"(0));
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(wv) : "r"(o1), "r"(o0), "r"(hi));
            if (store_lane) *reinterpret_cast<uint32_t *>(po) = wv;
            po += out_pitch;
        }
        part[0] = eprev[0] - 4 * s0; part[1] = eprev[1] - 4 * s1;
        part[2] = eprev[2] - 4 * s2; part[3] = eprev[3] - 4 * s3;
#pragma unroll
        for (int j = 0; j < 4; j++) eprev[j] = e[j];
    };
    const L4Tag<false> quiet;
    const L4Tag<true> storing;
#pragma unroll
    for (int i = 0; i < 2 * R; i++) step(quiet, i);
    constexpr int P0 = (2 * R) % PF;
    int i = 0;
    for (; i + PF <= n_main; i += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) step(storing, P0 + u);
    }
#pragma unroll
    for (int u = 0; u < PF - 1; u++)
        if (i + u < n_main) step(storing, P0 + u);
}

template <int K, typename T>
__global__ void __launch_bounds__(L4_WARPS * 32)
k_laplacian4(const T *__restrict__ img, int64_t pitch, int w, int h, const uint8_t *__restrict__ lut,
             const KrDevStats *__restrict__ st, int slot, int invert, uint8_t *__restrict__ out,
             int64_t out_pitch, int seg)
{
    constexpr int R = K / 2;
    constexpr bool U8 = sizeof(T) == 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // 16-bit integer rasters with a value range of at least 256: arithmetic _to_uint8 (lap4_body);
    // narrower ranges keep the (global, L1-resident) table
    bool arith = false;                                                     // block-uniform
    uint32_t base2 = 0u, magic = 0u;
    int mn = 0;
    if (!U8) {
        mn = st->min_i[slot];
        const int range = st->max_i[slot] - mn;
        if (range >= 256 && range <= 65535) {
            arith = true;
            magic = (uint32_t)((((unsigned long long)255 << 32) + (unsigned)range - 1ull) / (unsigned)range);
            base2 = ((uint32_t)mn & 0xffffu) * 0x00010001u;
        }
    }
    const int wx = blockIdx.x * L4_WARPS + wid;
    int x0 = wx * L4_VALID - 4;                       // first (virtual) column of the warp
    if (x0 + 4 >= w) return;
    if (x0 + 128 > w + 4) x0 = w - 124;               // last warp: shifted left, columns overlap
    const bool left = x0 < 0, right = x0 + 128 > w;
    int lc = x0 + 4 * lane, edge = 0;
    if (left && lane == 0) { lc = 0; edge = 1; }
    if (right && lane == 31) { lc = w - 4; edge = 2; }
    const bool store_lane = lane >= 1 && lane <= 30;
    const int ys = blockIdx.y * seg, ye = min(ys + seg, h);
    const bool rowfast = (ys - R >= 0) && (ye + R + L4_PF <= h);
    // the first and the last block of a row hold the mirrored lanes (block-uniform choice)
    const bool edge_block = blockIdx.x == 0 || blockIdx.x == gridDim.x - 1;
#define L4_BODY(RF, SL, ED)                                                                          \
    lap4_body<K, T, RF, SL, ED>(img, pitch, h, lut, mn, base2, magic, invert, out, out_pitch, lc, \
                                x0 + 4 * lane, edge, left || right, store_lane, ys, ye)
#define L4_PICK(RF, SL)                                                                              \
    do { if (edge_block) L4_BODY(RF, SL, true); else L4_BODY(RF, SL, false); } while (0)
    if (arith) {
        if (rowfast) L4_PICK(true, true); else L4_PICK(false, true);
    } else {
        if (rowfast) L4_PICK(true, false); else L4_PICK(false, false);
    }
#undef L4_PICK
#undef L4_BODY
}

template <typename T>
int launch_minmax(kr_ctx *ctx, const void *a, int64_t pa, const void *b, int64_t pb, int w, int h,
                  int slot_a, int slot_b, int has_nd_a, double nd_a, int has_nd_b, double nd_b,
                  uint8_t *mask, int64_t pm, cudaStream_t s)
{
    const size_t va = sizeof(T) * 4;
    bool vec = ((uintptr_t)a % va == 0) && (pa % (int64_t)va == 0);
    if (b) vec = vec && ((uintptr_t)b % va == 0) && (pb % (int64_t)va == 0);
    if (mask) vec = vec && ((uintptr_t)mask % 4 == 0) && (pm % 4 == 0);
    int grid = ctx->num_sms * 8;
    if (grid > h) grid = h;
    if (vec && sizeof(T) == 2 && !PixTraits<T>::is_float && ((T)-1 > (T)0) && b && mask && !has_nd_a &&
        !has_nd_b && slot_a == 0 && slot_b == 1) {
        k_minmax_mask_u16<<<grid, 256, 0, s>>>((const uint16_t *)a, pa, (const uint16_t *)b, pb, w, h, mask,
                                              pm, ctx->d_stats);
        KR_LAUNCH_CHECK();
        return KR_OK;
    }
    if (vec)
        k_minmax_mask<T, true><<<grid, 256, 0, s>>>((const T *)a, pa, (const T *)b, pb, w, h, slot_a,
                                                   slot_b, has_nd_a, nd_a, has_nd_b, nd_b, mask, pm,
                                                   ctx->d_stats);
    else
        k_minmax_mask<T, false><<<grid, 256, 0, s>>>((const T *)a, pa, (const T *)b, pb, w, h, slot_a,
                                                    slot_b, has_nd_a, nd_a, has_nd_b, nd_b, mask, pm,
                                                    ctx->d_stats);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

int dispatch_minmax(kr_ctx *ctx, const void *a, int64_t pa, const void *b, int64_t pb, int dtype, int w,
                    int h, int slot_a, int slot_b, int has_nd_a, double nd_a, int has_nd_b,
                    double nd_b, uint8_t *mask, int64_t pm, cudaStream_t s)
{
    switch (dtype) {
    case KR_U8: return launch_minmax<uint8_t>(ctx, a, pa, b, pb, w, h, slot_a, slot_b, has_nd_a, nd_a, has_nd_b, nd_b, mask, pm, s);
    case KR_U16: return launch_minmax<uint16_t>(ctx, a, pa, b, pb, w, h, slot_a, slot_b, has_nd_a, nd_a, has_nd_b, nd_b, mask, pm, s);
    case KR_I16: return launch_minmax<int16_t>(ctx, a, pa, b, pb, w, h, slot_a, slot_b, has_nd_a, nd_a, has_nd_b, nd_b, mask, pm, s);
    case KR_F32: return launch_minmax<float>(ctx, a, pa, b, pb, w, h, slot_a, slot_b, has_nd_a, nd_a, has_nd_b, nd_b, mask, pm, s);
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    }
}

// rows per block of k_laplacian4: whole waves of co-resident blocks, each segment
// pays 2R warm-up rows and the table copy (KR_LAP4_SEG overrides, for tuning)
template <int K, typename T>
int launch_lap4(kr_ctx *ctx, const void *img, int64_t pitch, int w, int h, int slot, int invert,
                uint8_t *out, int64_t out_pitch, cudaStream_t s)
{
    constexpr int R = K / 2;
    const size_t smem = 0;
    struct Cfg { int bps, seg; cudaError_t err; };
    static const Cfg cfg = [smem] {
        Cfg c;
        c.bps = 0;
        c.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.bps, k_laplacian4<K, T>, L4_WARPS * 32, smem);
        if (c.bps < 1) c.bps = 1;
        const char *e = getenv("KR_LAP4_SEG");
        c.seg = e ? atoi(e) : 0;
        return c;
    }();
    KR_CUDA(cfg.err);
    const int bps = cfg.bps, seg_env = cfg.seg;
    const int nwx = (w + L4_VALID - 1) / L4_VALID, bx = (nwx + L4_WARPS - 1) / L4_WARPS;
    // 96-row segments: several waves of blocks in different phases (load / cascade / store)
    // balance better than one wave of long segments (measured per S2 plane: 0.175 ms at 244
    // rows, 0.144 at 128, 0.120 at 96, 0.124 at 64, 0.135 at 32), for 2R / 96 extra warm-up rows;
    // small images get at least one block per SM slot
    (void)bps;
    int seg = seg_env > 0 ? ((seg_env + 3) & ~3) : 96;
    if (seg_env <= 0) {
        const int slots = ctx->num_sms * bps;
        while (seg > 32 && (int64_t)bx * ((h + seg - 1) / seg) < slots) seg -= 16;
    }
    (void)R;
    dim3 grid(bx, (h + seg - 1) / seg);
    k_laplacian4<K, T><<<grid, L4_WARPS * 32, smem, s>>>((const T *)img, pitch, w, h, ctx->d_lut[slot],
                                                        ctx->d_stats, slot, invert, out, out_pitch, seg);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

template <int K, typename T>
int launch_lap(kr_ctx *ctx, const void *img, int64_t pitch, int w, int h, int slot, int invert,
               uint8_t *out, int64_t out_pitch, cudaStream_t s)
{
    if (K >= 3 && K <= 7 && !PixTraits<T>::is_float) {
        constexpr int K4 = (K >= 3 && K <= 7) ? K : 3;
        typedef typename std::conditional<PixTraits<T>::is_float, uint16_t, T>::type T4;
        const size_t va = sizeof(T) * 4;
        static const bool off = getenv("KR_NO_LAP4") != nullptr;
        if (!off && w % 4 == 0 && w >= 256 && h >= 64 && (uintptr_t)img % va == 0 && pitch % (int64_t)va == 0 &&
            (uintptr_t)out % 4 == 0 && out_pitch % 4 == 0)
            return launch_lap4<K4, T4>(ctx, img, pitch, w, h, slot, invert, out, out_pitch, s);
    }
    if (K >= 3 && K <= 7) {
        constexpr int K2 = (K >= 3 && K <= 7) ? K : 3;
        constexpr int HL = (K2 / 2 + 1) / 2, VALID = 64 - 4 * HL;
        const size_t va = sizeof(T) * 2;
        const int aligned = ((uintptr_t)img % va == 0) && (pitch % (int64_t)va == 0) &&
                            ((uintptr_t)out % 2 == 0) && (out_pitch % 2 == 0);
        dim3 grid((w + LAP_WARPS * VALID - 1) / (LAP_WARPS * VALID), (h + LAP_ROWS - 1) / LAP_ROWS);
        k_laplacian2<K2, T><<<grid, LAP_WARPS * 32, 0, s>>>((const T *)img, pitch, w, h, ctx->d_lut[slot],
                                                           ctx->d_stats, slot, invert, out, out_pitch, aligned);
        KR_LAUNCH_CHECK();
        return KR_OK;
    }
    constexpr int R = (K <= 3) ? 1 : K / 2;
    constexpr int VALID = 32 - 2 * R;
    dim3 grid((w + LAP_WARPS * VALID - 1) / (LAP_WARPS * VALID), (h + LAP_ROWS - 1) / LAP_ROWS);
    k_laplacian<K, T><<<grid, LAP_WARPS * 32, 0, s>>>((const T *)img, pitch, w, h, ctx->d_lut[slot],
                                                     ctx->d_stats, slot, invert, out, out_pitch);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

template <typename T>
int dispatch_lap_k(kr_ctx *ctx, const void *img, int64_t pitch, int w, int h, int slot, int ksize,
                   int invert, uint8_t *out, int64_t out_pitch, cudaStream_t s)
{
    switch (ksize) {
    case 1: return launch_lap<1, T>(ctx, img, pitch, w, h, slot, invert, out, out_pitch, s);
    case 3: return launch_lap<3, T>(ctx, img, pitch, w, h, slot, invert, out, out_pitch, s);
    case 5: return launch_lap<5, T>(ctx, img, pitch, w, h, slot, invert, out, out_pitch, s);
    case 7: return launch_lap<7, T>(ctx, img, pitch, w, h, slot, invert, out, out_pitch, s);
    case 9: return launch_lap<9, T>(ctx, img, pitch, w, h, slot, invert, out, out_pitch, s);
    case 11: return launch_lap<11, T>(ctx, img, pitch, w, h, slot, invert, out, out_pitch, s);
    default:
        return kr_set_error(KR_ERR_UNSUPPORTED,
                            "Laplacian ksize %d not supported (1,3,5,7,9,11 are)", ksize);
    }
}

}  // namespace

int krl_reset_stats(kr_ctx *ctx, cudaStream_t s)
{
    k_reset_stats<<<1, 32, 0, s>>>(ctx->d_stats);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

int krl_minmax_mask(kr_ctx *ctx, const void *a, int64_t pa, const void *b, int64_t pb, int dtype, int w,
                    int h, int has_nd_a, double nd_a, int has_nd_b, double nd_b, uint8_t *mask,
                    int64_t pm, cudaStream_t s)
{
    return dispatch_minmax(ctx, a, pa, b, pb, dtype, w, h, 0, 1, has_nd_a, nd_a, has_nd_b, nd_b, mask,
                           pm, s);
}

int krl_minmax_single(kr_ctx *ctx, const void *img, int64_t pitch, int dtype, int w, int h, int slot,
                      cudaStream_t s)
{
    k_reset_slot<<<1, 1, 0, s>>>(ctx->d_stats, slot);
    KR_LAUNCH_CHECK();
    return dispatch_minmax(ctx, img, pitch, nullptr, 0, dtype, w, h, slot, slot, 0, 0.0, 0, 0.0,
                           nullptr, 0, s);
}

int krl_laplacian(kr_ctx *ctx, const void *img, int64_t pitch, int dtype, int w, int h, int slot,
                  int ksize, int invert, uint8_t *out, int64_t out_pitch, cudaStream_t s)
{
    if (slot < 0 || slot > 2) return kr_set_error(KR_ERR_INVALID, "bad min/max slot %d", slot);
    if (dtype != KR_F32) {
        k_build_lut<<<65536 / 256, 256, 0, s>>>(ctx->d_stats, slot, dtype, invert, ctx->d_lut[slot]);
        KR_LAUNCH_CHECK();
    }
    switch (dtype) {
    case KR_U8: return dispatch_lap_k<uint8_t>(ctx, img, pitch, w, h, slot, ksize, invert, out, out_pitch, s);
    case KR_U16: return dispatch_lap_k<uint16_t>(ctx, img, pitch, w, h, slot, ksize, invert, out, out_pitch, s);
    case KR_I16: return dispatch_lap_k<int16_t>(ctx, img, pitch, w, h, slot, ksize, invert, out, out_pitch, s);
    case KR_F32: return dispatch_lap_k<float>(ctx, img, pitch, w, h, slot, ksize, invert, out, out_pitch, s);
    default: return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    }
}

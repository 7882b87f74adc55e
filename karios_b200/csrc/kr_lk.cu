// kr_lk.cu -- K5 (pyrDown) and K6 (pyramidal Lucas-Kanade, forward + backward +
// back-check), plus the compaction of the surviving rows.
//
// Replaces cv2.calcOpticalFlowPyrLK(prev, next, p0, None, winSize=(w,w),
// maxLevel=1, criteria=(EPS|COUNT, 30, 0.03)) called twice at
// karios/matcher/klt.py:134-140 and the back-check of klt.py:142-159.
// Semantics follow SURVEY.md A.5 (verified against cv2 4.13 by
// oracle/klt_oracle.c: status identical, positions within ~1e-4 px):
//   pyramid level l+1 = 5-tap [1,4,6,4,1] pyrDown, (sum+128)>>8, REFLECT_101;
//   Scharr derivatives of the previous image (zero outside the image);
//   14-bit fixed-point bilinear weights, patch values with 5 fractional bits;
//   normal equations from exact integer sums (OpenCV: float32 SIMD sums);
//   termination |delta|^2 <= eps^2 or the 0.01 oscillation test.
//
// One warp tracks one point.  Lane l owns window column l: a window row is one
// coalesced byte load per lane, horizontal neighbours come from warp shuffles,
// vertical neighbours from the previous row kept in registers.  The template
// patch (Iw, Ix, Iy) lives in shared memory as one 64-bit entry per pixel in rows
// of 32 (the entries beyond the window are zero, so the lanes beyond it need no
// masking); the 2x2 system is reduced with two REDUX per sum.
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include "kr_internal.cuh"

namespace {

// ---------------------------------------------------------------- K5 pyrDown
constexpr int PD_ROWS = 32, PD_VALID = 60;

struct PyrPair {
    const uint8_t *src[2];
    uint8_t *dst[2];
    int64_t src_pitch[2], dst_pitch[2];
};

// cv::pyrDown: [1,4,6,4,1] x [1,4,6,4,1], REFLECT_101, (sum + 128) >> 8, output
// ((w+1)/2, (h+1)/2).  One warp walks down a strip of 60 output columns: lane <->
// two output columns = four input columns (one aligned 32-bit load per lane and
// input row), the taps that belong to the neighbour lanes arrive by two warp
// shuffles of the packed word, the horizontal 5-tap sums are two dp4a, the five
// vertical taps slide through registers as packed 16-bit pairs (two new input
// rows per output row).  No shared memory.
// horizontal 5-tap sums of the lane's two outputs from its word (a0..a3) and the
// neighbour lanes' words, packed as two 16-bit fields (each <= 4080)
__device__ __forceinline__ uint32_t pd_hsum(uint32_t cur)
{
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t lw = __shfl_up_sync(FULL, cur, 1), rw = __shfl_down_sync(FULL, cur, 1);
    // output 0 (input centre a0): L.a2 + 4 L.a3 + 6 a0 + 4 a1 + a2
    const uint32_t w0 = __byte_perm(lw, cur, 0x5432);            // bytes (L.a2, L.a3, a0, a1)
    const uint32_t h0 = __dp4a(w0, 0x04060401u, (cur >> 16) & 255u);
    // output 1 (input centre a2): a0 + 4 a1 + 6 a2 + 4 a3 + R.a0
    const uint32_t h1 = __dp4a(cur, 0x04060401u, rw & 255u);
    return h0 | (h1 << 16);
}

// FAST: all 128 input columns of the warp and all input rows it touches are inside
// the image and the planes are aligned (running pointers, 32-bit loads, 16-bit stores).
template <bool FAST>
__device__ __forceinline__ void pyr_down_body(const uint8_t *__restrict__ src, int64_t sp,
                                              uint8_t *__restrict__ dst, int64_t dp, int w, int h, int dw,
                                              int xs, int oys, int oye, int lane)
{
    const int ox = xs + 2 * (lane - 1);                            // two outputs: ox, ox + 1
    const int cx = 2 * ox;                                         // four inputs: cx .. cx + 3
    int tc[4];
#pragma unroll
    for (int j = 0; j < 4; j++) tc[j] = FAST ? cx + j : kr_reflect101(cx + j, w);
    const bool lane_ok = lane >= 1 && lane <= 30;
    const bool st0 = lane_ok && ox < dw, st1 = lane_ok && ox + 1 < dw;
    const int r0 = 2 * oys - 2;
    const uint8_t *pl = src + (int64_t)r0 * sp + cx;               // FAST only
    int r_load = r0;
    auto load_next = [&]() -> uint32_t {
        if (FAST) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(pl));
            pl += sp;
            return v;
        }
        int tr = r_load;
        if ((unsigned)tr >= (unsigned)h) tr = kr_reflect101(r_load, h);
        r_load++;
        const uint8_t *row = src + (int64_t)tr * sp;
        return (uint32_t)__ldg(row + tc[0]) | ((uint32_t)__ldg(row + tc[1]) << 8) |
               ((uint32_t)__ldg(row + tc[2]) << 16) | ((uint32_t)__ldg(row + tc[3]) << 24);
    };
    uint8_t *po = dst + (int64_t)oys * dp + ox;
    auto store_next = [&](uint32_t acc) {
        const uint32_t o0 = (acc >> 8) & 255u, o1 = acc >> 24;
        if (FAST) {
            if (lane_ok) *reinterpret_cast<uint16_t *>(po) = (uint16_t)(o0 | (o1 << 8));
        } else {
            if (st0) po[0] = (uint8_t)o0;
            if (st1) po[1] = (uint8_t)o1;
        }
        po += dp;
    };

    uint32_t h0 = pd_hsum(load_next()), h1 = pd_hsum(load_next()), h2 = pd_hsum(load_next());
    // four input rows (two output rows) per iteration, requested two iterations ahead
    uint32_t n0 = load_next(), n1 = load_next(), n2 = load_next(), n3 = load_next();
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (oys + 2 < oye) { m0 = load_next(); m1 = load_next(); m2 = load_next(); m3 = load_next(); }
    for (int oy = oys; oy < oye; oy += 2) {
        const uint32_t c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        n0 = m0; n1 = m1; n2 = m2; n3 = m3;
        if (oy + 4 < oye) { m0 = load_next(); m1 = load_next(); m2 = load_next(); m3 = load_next(); }
        // packed 16-bit pairs: (h0 + h4) + 4 (h1 + h3) + 6 h2 <= 16 * 4080 = 65280 per field
        const uint32_t h3 = pd_hsum(c0), h4 = pd_hsum(c1);
        store_next(h0 + h4 + 4u * (h1 + h3) + 6u * h2 + 0x00800080u);
        const uint32_t h5 = pd_hsum(c2), h6 = pd_hsum(c3);       // shuffles: all lanes take part
        if (oy + 1 < oye) store_next(h2 + h6 + 4u * (h3 + h5) + 6u * h4 + 0x00800080u);
        h0 = h4; h1 = h5; h2 = h6;
    }
}

// Control flow depends on blockIdx only (warp-uniform as far as the compiler can tell, so the
// shuffles need no convergence guards): the fast / general choice is made per block, and a warp
// whose strip starts right of the image recomputes the last strip instead of leaving early.
template <int PD_WARPS>
__global__ void __launch_bounds__(PD_WARPS * 32) k_pyr_down(PyrPair pp, int w, int h, int aligned)
{
    const uint8_t *__restrict__ src = pp.src[blockIdx.z];
    uint8_t *__restrict__ dst = pp.dst[blockIdx.z];
    const int64_t sp = pp.src_pitch[blockIdx.z], dp = pp.dst_pitch[blockIdx.z];
    const int dw = (w + 1) >> 1, dh = (h + 1) >> 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int last_xs = ((dw - 1) / PD_VALID) * PD_VALID;         // first output column of the last strip
    const int xs = min((blockIdx.x * PD_WARPS + wid) * PD_VALID, last_xs);
    const int oys = blockIdx.y * PD_ROWS, oye = min(oys + PD_ROWS, dh);
    const int bx0 = blockIdx.x * PD_WARPS * PD_VALID;             // block: outputs [bx0, bx0 + PD_WARPS * 60)
    const bool fast = aligned && (2 * bx0 - 4 >= 0) && (2 * (bx0 + (PD_WARPS - 1) * PD_VALID) - 4 + 128 <= w) &&
                      (bx0 + (PD_WARPS - 1) * PD_VALID <= last_xs) && (2 * oys - 2 >= 0) && (2 * oye + 6 <= h);
    if (fast) pyr_down_body<true>(src, sp, dst, dp, w, h, dw, xs, oys, oye, lane);
    else pyr_down_body<false>(src, sp, dst, dp, w, h, dw, xs, oys, oye, lane);
}

// K5, four outputs per lane (w % 4 == 0, w >= 256, h >= 8, 4-byte aligned planes): a warp
// covers 256 input columns / 120 output columns (lanes 1..30), a lane loads its 8 input bytes of
// a row as two 32-bit words, the taps of the neighbour lanes arrive by two shuffles per row (for
// four outputs), four dp4a form the horizontal sums, the vertical taps slide through registers
// as packed 16-bit pairs, and the four results leave with one byte permute and one 32-bit store.
// No scalar border path: the first warp starts 8 columns left of the image and its lane 0
// mirrors lane 1's bytes (REFLECT_101), the last warp is shifted left to end at the image edge
// and its lane 31 mirrors lane 30's; rows are reflected by index in the top / bottom blocks.
constexpr int P4_VALID = 120, P4_ROWS = 32;

struct P4Row { uint32_t lo, hi; };

template <bool ROWFAST, bool EDGE>
__device__ __forceinline__ void pyr4_body(const uint8_t *__restrict__ src, int64_t sp,
                                          uint8_t *__restrict__ dst, int64_t dp, int h, int lc, int edge,
                                          bool warp_edge, bool store_lane, bool lane31, int oc, int oys,
                                          int oye)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int r0 = 2 * oys - 2;
    const uint8_t *pl = src + (int64_t)r0 * sp + lc;               // ROWFAST
    const uint8_t *pc = src + lc;
    int r_load = r0;
    auto load_next = [&]() -> P4Row {
        const uint8_t *p;
        if (ROWFAST) {
            p = pl;
            pl += sp;
        } else {
            int tr = r_load;
            if (tr < 0) tr = -tr;
            if (tr >= h) tr = 2 * (h - 1) - tr;
            r_load++;
            p = pc + (int64_t)tr * sp;
        }
        P4Row q;
        q.lo = __ldg(reinterpret_cast<const uint32_t *>(p));
        q.hi = (lane31 && edge == 0) ? 0u : __ldg(reinterpret_cast<const uint32_t *>(p + 4));
        if (EDGE && warp_edge) {
            if (edge == 1) q.hi = __byte_perm(q.lo, q.lo, 0x1200);       // columns -2, -1 = p2, p1
            if (edge == 2) q.lo = __byte_perm(q.hi, q.hi, 0x3212);       // column w = p(w - 2)
        }
        return q;
    };
    // horizontal [1 4 6 4 1] sums of the lane's four outputs, packed (h0 | h1 << 16), (h2 | h3 << 16)
    auto hsum = [&](const P4Row q, uint32_t &HA, uint32_t &HB) {
        const uint32_t lw = __shfl_up_sync(FULL, q.hi, 1), rw = __shfl_down_sync(FULL, q.lo, 1);
        const uint32_t w0 = __byte_perm(lw, q.lo, 0x5432);               // L.a6 L.a7 a0 a1
        const uint32_t w2 = __byte_perm(q.lo, q.hi, 0x5432);             // a2 a3 a4 a5
        const uint32_t h0 = __dp4a(w0, 0x04060401u, (q.lo >> 16) & 255u);
        const uint32_t h1 = __dp4a(q.lo, 0x04060401u, q.hi & 255u);
        const uint32_t h2 = __dp4a(w2, 0x04060401u, (q.hi >> 16) & 255u);
        const uint32_t h3 = __dp4a(q.hi, 0x04060401u, rw & 255u);
        HA = h0 | (h1 << 16);
        HB = h2 | (h3 << 16);
    };
    uint8_t *po = dst + (int64_t)oys * dp + oc;
    auto store_next = [&](uint32_t accA, uint32_t accB) {
        const uint32_t wv = __byte_perm(accA, accB, 0x7531);             // (acc >> 8) & 255 of the four fields
        if (store_lane) {
            if (EDGE) {
                *reinterpret_cast<uint16_t *>(po) = (uint16_t)wv;
                *reinterpret_cast<uint16_t *>(po + 2) = (uint16_t)(wv >> 16);
            } else {
                *reinterpret_cast<uint32_t *>(po) = wv;
            }
        }
        po += dp;
    };
    uint32_t a0, b0, a1, b1, a2, b2;
    hsum(load_next(), a0, b0);
    hsum(load_next(), a1, b1);
    hsum(load_next(), a2, b2);
    P4Row n0 = load_next(), n1 = load_next(), n2 = load_next(), n3 = load_next();
    for (int oy = oys; oy < oye; oy += 2) {
        const P4Row c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        if (oy + 2 < oye) { n0 = load_next(); n1 = load_next(); n2 = load_next(); n3 = load_next(); }
        uint32_t a3, b3, a4, b4, a5, b5, a6, b6;
        hsum(c0, a3, b3);
        hsum(c1, a4, b4);
        // packed 16-bit pairs: (h0 + h4) + 4 (h1 + h3) + 6 h2 + 128 <= 65408 per field
        store_next(a0 + a4 + 4u * (a1 + a3) + 6u * a2 + 0x00800080u, b0 + b4 + 4u * (b1 + b3) + 6u * b2 + 0x00800080u);
        hsum(c2, a5, b5);
        hsum(c3, a6, b6);
        if (oy + 1 < oye)
            store_next(a2 + a6 + 4u * (a3 + a5) + 6u * a4 + 0x00800080u, b2 + b6 + 4u * (b3 + b5) + 6u * b4 + 0x00800080u);
        a0 = a4; b0 = b4; a1 = a5; b1 = b5; a2 = a6; b2 = b6;
    }
}

template <int PD_WARPS>
__global__ void __launch_bounds__(PD_WARPS * 32) k_pyr_down4(PyrPair pp, int w, int h)
{
    const uint8_t *__restrict__ src = pp.src[blockIdx.z];
    uint8_t *__restrict__ dst = pp.dst[blockIdx.z];
    const int64_t sp = pp.src_pitch[blockIdx.z], dp = pp.dst_pitch[blockIdx.z];
    const int dw = (w + 1) >> 1, dh = (h + 1) >> 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwx = (dw + P4_VALID - 1) / P4_VALID;
    const int wx = min((int)blockIdx.x * PD_WARPS + wid, nwx - 1);   // spare warps recompute the last strip
    int xs = wx * P4_VALID;
    if (xs + P4_VALID > dw) xs = dw - P4_VALID;                      // last warp: shifted left, outputs overlap
    const int c0 = 2 * xs - 8;                                       // first (virtual) input column of the warp
    // right: lane 31's first column is the first one past the image (only the shifted last warp);
    // otherwise lane 31 supplies one true pixel (its byte 0) and never touches its upper word,
    // which may lie past the end of the row
    const bool left = c0 < 0, right = c0 + 248 >= w;
    int lc = c0 + 8 * lane, edge = 0;
    if (left && lane == 0) { lc = 0; edge = 1; }
    if (right && lane == 31) { lc = c0 + 8 * 30; edge = 2; }
    const bool store_lane = lane >= 1 && lane <= 30;
    const int oc = xs + 4 * (lane - 1);
    const int oys = blockIdx.y * P4_ROWS, oye = min(oys + P4_ROWS, dh);
    const bool rowfast = (2 * oys - 2 >= 0) && (2 * oye + 6 <= h);
    const bool edge_block = blockIdx.x == 0 || blockIdx.x == gridDim.x - 1;      // block-uniform
#define P4_BODY(RF, ED) \
    pyr4_body<RF, ED>(src, sp, dst, dp, h, lc, edge, left || right, store_lane, lane == 31, oc, oys, oye)
    if (edge_block) { if (rowfast) P4_BODY(true, true); else P4_BODY(false, true); }
    else { if (rowfast) P4_BODY(true, false); else P4_BODY(false, false); }
#undef P4_BODY
}

// warps per block of k_pyr_down: few, so that the blocks holding an image edge (general path
// for all their warps) stay a small share; KR_PYR_WARPS overrides (1, 2, 4, 8)
int pyr_warps()
{
    static const int v = [] {
        const char *e = getenv("KR_PYR_WARPS");
        const int k = e ? atoi(e) : 2;
        return (k == 1 || k == 2 || k == 4 || k == 8) ? k : 2;
    }();
    return v;
}

int launch_pyr_down(const PyrPair &q, int w, int h, int nw, int nh, int nz, int aligned, cudaStream_t s)
{
    const int pw = pyr_warps();
    static const bool no4 = getenv("KR_NO_PYR4") != nullptr;
    bool ok4 = !no4 && w % 4 == 0 && w >= 256 && h >= 8;
    for (int k = 0; k < nz; k++)
        ok4 = ok4 && ((uintptr_t)q.src[k] % 4 == 0) && (q.src_pitch[k] % 4 == 0) && ((uintptr_t)q.dst[k] % 4 == 0) &&
              (q.dst_pitch[k] % 4 == 0);
    if (ok4) {
        const int nwx = (nw + P4_VALID - 1) / P4_VALID;
        dim3 g4((nwx + pw - 1) / pw, (nh + P4_ROWS - 1) / P4_ROWS, nz);
        switch (pw) {
        case 1: k_pyr_down4<1><<<g4, 32, 0, s>>>(q, w, h); break;
        case 4: k_pyr_down4<4><<<g4, 128, 0, s>>>(q, w, h); break;
        case 8: k_pyr_down4<8><<<g4, 256, 0, s>>>(q, w, h); break;
        default: k_pyr_down4<2><<<g4, 64, 0, s>>>(q, w, h); break;
        }
        KR_LAUNCH_CHECK();
        return KR_OK;
    }
    dim3 grid((nw + pw * PD_VALID - 1) / (pw * PD_VALID), (nh + PD_ROWS - 1) / PD_ROWS, nz);
    switch (pw) {
    case 1: k_pyr_down<1><<<grid, 32, 0, s>>>(q, w, h, aligned); break;
    case 4: k_pyr_down<4><<<grid, 128, 0, s>>>(q, w, h, aligned); break;
    case 8: k_pyr_down<8><<<grid, 256, 0, s>>>(q, w, h, aligned); break;
    default: k_pyr_down<2><<<grid, 64, 0, s>>>(q, w, h, aligned); break;
    }
    KR_LAUNCH_CHECK();
    return KR_OK;
}

// --------------------------------------------------------------------- K6 LK
// One warp per block (and per point): the point index and every trip count derive from
// blockIdx, which the compiler knows to be warp-uniform -- no convergence guards around the
// shuffles (with 8 warps per block BRA.DIV / BSSY / BSYNC / UMOV were 9 % of the instructions).
constexpr int LK_WARPS = 1;
constexpr int W_BITS = 14;

// Sum of one int32 per lane as a 64-bit total (the total of 25 lanes can exceed 32 bits):
// two REDUX instructions on the signed high and the unsigned low half.  The results land in
// uniform registers, so the convergence tests that follow them are known to be warp-uniform.
__device__ __forceinline__ long long warp_sum_i32(int v)
{
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
    return ((long long)hi << 16) + (long long)lo;
}

__device__ __forceinline__ void lk_weights(float a, float b, int &w00, int &w01, int &w10, int &w11)
{
    const float one_a = __fsub_rn(1.f, a), one_b = __fsub_rn(1.f, b);
    const float sc = (float)(1 << W_BITS);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, one_b), sc));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, one_b), sc));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, b), sc));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// Sum over the window of |J(pos) - Iw| (flag ABS) or of diff*Ix, diff*Iy, at
// integer origin (inx, iny) with the given bilinear weights.  All lanes get the
// totals.  FAST: the (win+1)^2 footprint lies inside the image -- no border
// arithmetic, running row pointer, next row prefetched.  WIN > 0 fixes the
// window at compile time (unrolled rows).
template <bool ABS, bool FAST, int WIN>
__device__ __forceinline__ void lk_residual(const uint8_t *__restrict__ J, int64_t pJ, int w, int h,
                                            int inx, int iny, int win_rt, int w00, int w01, int w10,
                                            int w11, const uint2 *sT, int lane, long long &o1,
                                            long long &o2)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int win = WIN ? WIN : win_rt;
    const bool act = lane < win;
    const int lc = act ? lane : win;
    const int ts = win + 1;                  // template row stride; column `win` is a zero pad
    const uint2 *tp = sT + lc;               // lanes beyond the window read the pad
    int b1 = 0, b2 = 0;
    if (FAST) {
        const uint8_t *p = J + (int64_t)iny * pJ + (inx + lc);
        int tv = __ldg(p);
        int vn = __ldg(p + pJ);
        int tvr = __shfl_down_sync(FULL, tv, 1);
#pragma unroll 5
        for (int r = 1; r <= win; r++) {
            const int v = vn;
            p += pJ;
            if (r < win) vn = __ldg(p + pJ);
            const int vr = __shfl_down_sync(FULL, v, 1);
            const int val = (tv * w00 + tvr * w01 + v * w10 + vr * w11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
            const uint2 t = tp[(r - 1) * ts];
            int diff = val - (int)(int16_t)(t.x & 0xffffu);
            if (ABS) {
                if (!act) diff = 0;
                b1 += abs(diff);
            } else {                                       // zero gradients beyond the window
                b1 += diff * ((int)t.x >> 16);
                b2 += diff * (int)t.y;
            }
            tv = v;
            tvr = vr;
        }
    } else {
        const int col = kr_reflect101(inx + lc, w);
        int tv = 0, tvr = 0;
        for (int r = 0; r <= win; r++) {
            const uint8_t *rowp = J + (int64_t)kr_reflect101(iny + r, h) * pJ;
            const int v = __ldg(rowp + col);
            const int vr = __shfl_down_sync(FULL, v, 1);
            if (r >= 1) {
                const int val = (tv * w00 + tvr * w01 + v * w10 + vr * w11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
                const uint2 t = tp[(r - 1) * ts];
                int diff = val - (int)(int16_t)(t.x & 0xffffu);
                if (ABS) {
                    if (!act) diff = 0;
                    b1 += abs(diff);
                } else {
                    b1 += diff * ((int)t.x >> 16);
                    b2 += diff * (int)t.y;
                }
            }
            tv = v;
            tvr = vr;
        }
    }
    o1 = warp_sum_i32(b1);
    o2 = ABS ? 0ll : warp_sum_i32(b2);
}

// The window of J around the current position, cached in shared memory: the iterations of one
// level move it by a fraction of a pixel, so after the first (cold) fill every iteration reads
// its win + 1 rows from shared memory -- with 30 warps of windows per SM the L1 keeps under a
// third of the re-read sectors.  Region: (win + 1 + 2 LK_M)^2 bytes, row stride 32, origin
// (cx0, cy0); used while the integer origin of the window stays within the margin.
constexpr int LK_M = 2;

template <int WIN>
__device__ __forceinline__ void lk_fill_cache(const uint8_t *__restrict__ J, int64_t pJ, int cx0, int cy0,
                                              int win_rt, uint8_t *sJ, int lane)
{
    const int win = WIN ? WIN : win_rt;
    const int side = win + 1 + 2 * LK_M;               // <= 32
    const uint8_t *p = J + (int64_t)cy0 * pJ + cx0 + min(lane, side - 1);
    __syncwarp();
#pragma unroll 6
    for (int r = 0; r < side; r++) {
        sJ[r * 32 + lane] = __ldg(p);
        p += pJ;
    }
    __syncwarp();
}

template <int WIN>
__device__ __forceinline__ void lk_residual_cached(const uint8_t *sJw, int win_rt, int w00, int w01, int w10,
                                                   int w11, const uint2 *sT, int lane, long long &o1,
                                                   long long &o2)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int win = WIN ? WIN : win_rt;
    const int lc = lane < win ? lane : win;
    const int ts = win + 1;
    const uint2 *tp = sT + lc;
    const uint8_t *p = sJw + lc;                        // window origin inside the cache
    int b1 = 0, b2 = 0;
    int tv = p[0];
    int tvr = __shfl_down_sync(FULL, tv, 1);
#pragma unroll 5
    for (int r = 1; r <= win; r++) {
        const int v = p[r * 32];
        const int vr = __shfl_down_sync(FULL, v, 1);
        const int val = (tv * w00 + tvr * w01 + v * w10 + vr * w11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
        const uint2 t = tp[(r - 1) * ts];
        const int diff = val - (int)(int16_t)(t.x & 0xffffu);
        b1 += diff * ((int)t.x >> 16);
        b2 += diff * (int)t.y;
        tv = v;
        tvr = vr;
    }
    o1 = warp_sum_i32(b1);
    o2 = warp_sum_i32(b2);
}

// Template patch of one level: Iw (5 fractional bits), Ix, Iy (Scharr, bilinear at
// the sub-pixel origin) into shared memory, and the sums of Ix^2, IxIy, Iy^2.
// lane <-> image column ipx - 1 + lane; rows stream through registers.
template <bool FAST, int WIN>
__device__ __forceinline__ void lk_patch(const uint8_t *__restrict__ I, int64_t pI, int w, int h, int ipx,
                                         int ipy, int win_rt, int w00, int w01, int w10, int w11,
                                         uint2 *sT, int lane, long long &S11, long long &S12,
                                         long long &S22)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int win = WIN ? WIN : win_rt;
    const int cx = ipx - 1 + min(lane, win + 2);
    const int col = FAST ? cx : kr_reflect101(cx, w);
    const bool col_in = FAST || (cx >= 0 && cx < w);
    const bool act = lane >= 1 && lane <= win;
    const bool wr = lane >= 1 && lane <= win + 1;      // lane l -> column l - 1; lane win + 1 -> the zero pad
    const int ts = win + 1;
    uint2 *tp = sT + (wr ? lane - 1 : 0);
    int a = 0, b = 0, c = 0;
    int t_dx = 0, t_dxr = 0, t_dy = 0, t_dyr = 0, t_pv = 0, t_pvr = 0;
    int s11 = 0, s12 = 0, s22 = 0;
    const uint8_t *p = I + (FAST ? ((int64_t)(ipy - 1) * pI + col) : (int64_t)col);
    int cn = FAST ? (int)__ldg(p) : 0;                     // prefetched next row (FAST)
    __syncwarp();
#pragma unroll 4
    for (int r = 0; r < win + 3; r++) {
        const int ry = ipy - 1 + r;
        a = b; b = c;
        if (FAST) {
            c = cn;
            p += pI;
            if (r < win + 2) cn = __ldg(p);
        } else {
            c = __ldg(p + (int64_t)kr_reflect101(ry, h) * pI);
        }
        if (r < 2) continue;
        const int Y = ry - 1;                   // row of the derivative being formed
        const int S = 3 * (a + c) + 10 * b, D = c - a;
        const int Sl = __shfl_up_sync(FULL, S, 1), Sr = __shfl_down_sync(FULL, S, 1);
        const int Dl = __shfl_up_sync(FULL, D, 1), Dr = __shfl_down_sync(FULL, D, 1);
        int dxv = Sr - Sl, dyv = 3 * (Dl + Dr) + 10 * D;
        if (!FAST && !(col_in && Y >= 0 && Y < h)) { dxv = 0; dyv = 0; }    // derivative border = 0
        const int pv = b;
        const int dxr = __shfl_down_sync(FULL, dxv, 1);
        const int dyr = __shfl_down_sync(FULL, dyv, 1);
        const int pvr = __shfl_down_sync(FULL, pv, 1);
        if (r >= 3) {
            int ix = (t_dx * w00 + t_dxr * w01 + dxv * w10 + dxr * w11 + (1 << (W_BITS - 1))) >> W_BITS;
            int iy = (t_dy * w00 + t_dyr * w01 + dyv * w10 + dyr * w11 + (1 << (W_BITS - 1))) >> W_BITS;
            int iv = (t_pv * w00 + t_pvr * w01 + pv * w10 + pvr * w11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
            if (!act) { ix = 0; iy = 0; iv = 0; }             // the pad column
            // one 64-bit entry per window pixel: Iw (low half) | Ix (high half), Iy
            if (wr) tp[(r - 3) * ts] = make_uint2((uint32_t)(iv & 0xffff) | ((uint32_t)ix << 16), (uint32_t)iy);
            s11 += ix * ix; s12 += ix * iy; s22 += iy * iy;
        }
        t_dx = dxv; t_dxr = dxr; t_dy = dyv; t_dyr = dyr; t_pv = pv; t_pvr = pvr;
    }
    __syncwarp();
    S11 = warp_sum_i32(s11);
    S12 = warp_sum_i32(s12);
    S22 = warp_sum_i32(s22);
}

// calcOpticalFlowPyrLK for one point, all levels, direction dir (0: img[0] is
// the previous image, 1: img[1] is).  Uniform across the warp.
// WANT_ERR: also the mean absolute residual at the final position (OpenCV's err output;
// the round trip of klt_tracker never reads it, klt.py:134-144).
template <int WIN, bool WANT_ERR>
__device__ void lk_track(const KrLkArgs &A, int dir, float ptx, float pty, float &ox, float &oy,
                         uint8_t &status, float &err, uint2 *sT, uint8_t *sJ, int lane)
{
    const int win = WIN ? WIN : A.win;
    const float half = __fmul_rn((float)(win - 1), 0.5f);
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    status = 1;
    err = 0.f;
    ox = 0.f;
    oy = 0.f;
    for (int l = A.levels; l >= 0; l--) {
        const uint8_t *__restrict__ I = A.img[dir][l];
        const uint8_t *__restrict__ J = A.img[dir ^ 1][l];
        const int64_t pI = A.pitch[dir][l], pJ = A.pitch[dir ^ 1][l];
        const int w = A.w[l], h = A.h[l];
        const float scale = (float)(1.0 / (double)(1 << l));
        float px = __fmul_rn(ptx, scale), py = __fmul_rn(pty, scale);
        float nx, ny;
        if (l == A.levels) { nx = px; ny = py; }
        else { nx = __fmul_rn(ox, 2.f); ny = __fmul_rn(oy, 2.f); }
        ox = nx; oy = ny;
        px = __fsub_rn(px, half); py = __fsub_rn(py, half);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) {
            if (l == 0) { status = 0; err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        long long S11, S12, S22;
        if (ipx >= 1 && ipy >= 1 && ipx + win + 1 < w && ipy + win + 1 < h)
            lk_patch<true, WIN>(I, pI, w, h, ipx, ipy, win, w00, w01, w10, w11, sT, lane, S11, S12, S22);
        else
            lk_patch<false, WIN>(I, pI, w, h, ipx, ipy, win, w00, w01, w10, w11, sT, lane, S11, S12, S22);
        const float A11 = __fmul_rn((float)S11, FLT_SCALE), A12 = __fmul_rn((float)S12, FLT_SCALE),
                    A22 = __fmul_rn((float)S22, FLT_SCALE);
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)),
                                        (float)(2 * win * win));
        if (min_eig < A.min_eig_thr || Dt < FLT_EPSILON) {
            if (l == 0) status = 0;
            continue;
        }
        Dt = __fdiv_rn(1.f, Dt);

        // ---- iterations on J ----------------------------------------------
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
        bool cached = false;                       // J window cache of this level / direction
        int cx0 = 0, cy0 = 0;
        for (int j = 0; j < A.max_count; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= w || iny < -win || iny >= h) {
                if (l == 0) status = 0;
                break;
            }
            lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            long long sb1, sb2;
            if (inx >= 0 && iny >= 0 && inx + win < w && iny + win < h) {
                bool hit = cached && (unsigned)(inx - cx0) <= 2u * LK_M && (unsigned)(iny - cy0) <= 2u * LK_M;
                if (!hit && A.use_cache) {
                    const int ncx = inx - LK_M, ncy = iny - LK_M, side = win + 1 + 2 * LK_M;
                    if (ncx >= 0 && ncy >= 0 && ncx + side <= w && ncy + side <= h) {
                        lk_fill_cache<WIN>(J, pJ, ncx, ncy, win, sJ, lane);
                        cx0 = ncx; cy0 = ncy;
                        cached = hit = true;
                    }
                }
                if (hit)
                    lk_residual_cached<WIN>(sJ + (iny - cy0) * 32 + (inx - cx0), win, w00, w01, w10, w11, sT, lane,
                                            sb1, sb2);
                else
                    lk_residual<false, true, WIN>(J, pJ, w, h, inx, iny, win, w00, w01, w10, w11, sT, lane, sb1, sb2);
            } else {
                lk_residual<false, false, WIN>(J, pJ, w, h, inx, iny, win, w00, w01, w10, w11, sT, lane, sb1, sb2);
            }
            const float b1 = __fmul_rn((float)sb1, FLT_SCALE), b2 = __fmul_rn((float)sb2, FLT_SCALE);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt);
            nx = __fadd_rn(nx, ddx); ny = __fadd_rn(ny, ddy);
            ox = __fadd_rn(nx, half); oy = __fadd_rn(ny, half);
            if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= A.eps2) break;
            if (j > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 &&
                fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
                ox = __fsub_rn(ox, __fmul_rn(ddx, 0.5f));
                oy = __fsub_rn(oy, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx; pdy = ddy;
        }
        if (WANT_ERR && status && l == 0) {
            const float ex = __fsub_rn(ox, half), ey = __fsub_rn(oy, half);
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -win || iex >= w || iey < -win || iey >= h) {
                status = 0;
            } else {
                lk_weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey), w00, w01, w10, w11);
                long long sabs, dummy;
                if (iex >= 0 && iey >= 0 && iex + win < w && iey + win < h)
                    lk_residual<true, true, WIN>(J, pJ, w, h, iex, iey, win, w00, w01, w10, w11, sT, lane, sabs, dummy);
                else
                    lk_residual<true, false, WIN>(J, pJ, w, h, iex, iey, win, w00, w01, w10, w11, sT, lane, sabs, dummy);
                err = __fdiv_rn(__fmul_rn((float)sabs, 1.f), (float)(32 * win * win));
            }
        }
    }
}

__device__ __forceinline__ int lk_count(int n, const int32_t *d_count)
{
    if (d_count) { int c = *d_count; return c < n ? c : n; }
    return n;
}

template <int WIN>
__global__ void __launch_bounds__(LK_WARPS * 32)
k_lk_single(KrLkArgs A, const float *__restrict__ p0, int n, const int32_t *d_count,
            float *__restrict__ p1, uint8_t *__restrict__ status, float *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char lk_smem[];
    const int lane = threadIdx.x, wid = 0;                       // LK_WARPS == 1
    uint2 *sT = reinterpret_cast<uint2 *>(lk_smem);                        // [win][win + 1] template entries
    uint8_t *sJ = lk_smem + (size_t)A.win * (A.win + 1) * 8;               // [win + 1 + 2 LK_M][32] window cache
    (void)wid;
    const int cnt = lk_count(n, d_count);
    for (int i = blockIdx.x * LK_WARPS + wid; i < cnt; i += gridDim.x * LK_WARPS) {
        float ox, oy, e;
        uint8_t st;
        lk_track<WIN, true>(A, 0, p0[2 * i], p0[2 * i + 1], ox, oy, st, e, sT, sJ, lane);
        if (lane == 0) {
            p1[2 * i] = ox; p1[2 * i + 1] = oy;
            status[i] = st;
            err[i] = e;
        }
    }
}

// The windows a point is going to read -> L2, one row per lane (0.423 -> 0.416 ms on the S2 pair;
// an L2 hit instead of a DRAM access when the template build reaches the row).  Rows of a window are
// 26 - 30 bytes wide and start at any byte: two sectors per row.
__device__ __forceinline__ void lk_prefetch(const uint8_t *__restrict__ img, int64_t pitch, int w, int h, float cx,
                                            float cy, int half_w, int lane)
{
    const int x = (int)cx - half_w, y = (int)cy - half_w + lane;
    if (lane <= 2 * half_w && y >= 0 && y < h && x >= 0 && x + 2 * half_w < w) {
        const uint8_t *p = img + (int64_t)y * pitch + x;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 2 * half_w));
    }
}

// forward (ref -> mon), backward (mon -> ref) from the forward result, then the
// back-check of klt.py:142-144: d = max|p0 - p0r|, keep = d < 0.1 (float32).
template <int WIN>
__global__ void __launch_bounds__(LK_WARPS * 32)
k_lk_roundtrip(KrLkArgs A, const float *__restrict__ p0, int n_cap, const uint32_t *d_count,
               float back_thr, float *__restrict__ p1, float *__restrict__ dist,
               uint8_t *__restrict__ keep)
{
    extern __shared__ __align__(16) unsigned char lk_smem[];
    const int lane = threadIdx.x, wid = 0;                       // LK_WARPS == 1
    uint2 *sT = reinterpret_cast<uint2 *>(lk_smem);                        // [win][win + 1] template entries
    uint8_t *sJ = lk_smem + (size_t)A.win * (A.win + 1) * 8;               // [win + 1 + 2 LK_M][32] window cache
    (void)wid;
    int cnt = (int)min(*d_count, (uint32_t)n_cap);
    for (int i = blockIdx.x * LK_WARPS + wid; i < cnt; i += gridDim.x * LK_WARPS) {
        const float x0 = p0[2 * i], y0 = p0[2 * i + 1];
        float x1, y1, xr, yr, e;
        uint8_t st;
        // the forward pass reads the reference and the monitored window at every level around the
        // point (the displacement is a fraction of the window); the backward pass finds them cached
        for (int l = A.levels; l >= 0; l--) {
            const float sc = 1.f / (float)(1 << l);
            lk_prefetch(A.img[0][l], A.pitch[0][l], A.w[l], A.h[l], x0 * sc, y0 * sc, 15, lane);
            lk_prefetch(A.img[1][l], A.pitch[1][l], A.w[l], A.h[l], x0 * sc, y0 * sc, 15, lane);
        }
        lk_track<WIN, false>(A, 0, x0, y0, x1, y1, st, e, sT, sJ, lane);
        lk_track<WIN, false>(A, 1, x1, y1, xr, yr, st, e, sT, sJ, lane);
        if (lane == 0) {
            float d = fmaxf(fabsf(__fsub_rn(x0, xr)), fabsf(__fsub_rn(y0, yr)));
            p1[2 * i] = x1; p1[2 * i + 1] = y1;
            dist[i] = d;
            keep[i] = (d < back_thr) ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------ rows (a9, a11)
__global__ void __launch_bounds__(256)
k_row_keys(const float *__restrict__ p0, const uint8_t *__restrict__ keep, int n_cap,
           const uint32_t *d_count, int sort_xy, uint64_t *__restrict__ keys, KrDevStats *st)
{
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_base;
    const uint32_t n = min(*d_count, (uint32_t)n_cap);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        uint32_t i = i0 + threadIdx.x;
        bool k = (i < n) && keep[i];
        unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_cnt[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int q = 0; q < 8; q++) { uint32_t c = s_cnt[q]; s_cnt[q] = tot; tot += c; }
            s_base = tot ? atomicAdd(&st->n_rowkeys, tot) : 0;
        }
        __syncthreads();
        if (k) {
            uint64_t hi = 0;
            if (sort_xy) hi = ((uint64_t)(uint32_t)p0[2 * i] << 16) | (uint64_t)(uint32_t)p0[2 * i + 1];
            keys[s_base + s_cnt[wid] + __popc(bal & ((1u << lane) - 1))] = (hi << 32) | i;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_gather_rows(const uint64_t *__restrict__ sorted, const float *__restrict__ p0,
              const float *__restrict__ p1, const float *__restrict__ dist, float back_thr,
              float x_off, float y_off, kr_rows rows, KrDevStats *st)
{
    uint32_t n = st->n_rowkeys;
    if (n > (uint32_t)rows.capacity) n = (uint32_t)rows.capacity;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        uint32_t i = (uint32_t)sorted[r];
        float x0 = p0[2 * i], y0 = p0[2 * i + 1];
        rows.x0[r] = __fadd_rn(x0, x_off);
        rows.y0[r] = __fadd_rn(y0, y_off);
        rows.dx[r] = __fsub_rn(p1[2 * i], x0);
        rows.dy[r] = __fsub_rn(p1[2 * i + 1], y0);
        rows.score[r] = __fsub_rn(1.f, __fdiv_rn(dist[i], back_thr));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) st->n_kept = n;
}

__global__ void k_clear_rowkeys(KrDevStats *st) { st->n_rowkeys = 0; st->n_kept = 0; }

bool lk_cache_enabled(int win)
{
    static const bool off = getenv("KR_LK_NOCACHE") != nullptr;
    return !off && win + 1 + 2 * LK_M <= 32;
}
size_t lk_smem_bytes(int win)
{
    // KR_LK_SMEM_PAD: unused shared memory per block (caps the resident warps per SM, for tuning)
    static const size_t pad = getenv("KR_LK_SMEM_PAD") ? (size_t)atoi(getenv("KR_LK_SMEM_PAD")) : 0;
    return (size_t)LK_WARPS * ((size_t)win * (win + 1) * 8 + (lk_cache_enabled(win) ? (size_t)(win + 1 + 2 * LK_M) * 32 : 0)) + pad;
}

int lk_prepare(KrLkArgs &a, size_t *smem)
{
    a.use_cache = lk_cache_enabled(a.win) ? 1 : 0;
    if (a.win < 3 || a.win > 29)
        return kr_set_error(KR_ERR_UNSUPPORTED, "LK window %d not supported (3..29)", a.win);
    *smem = lk_smem_bytes(a.win);
    static size_t set1 = 0;
    if (*smem > set1 && *smem > 48 * 1024) {
        KR_CUDA(cudaFuncSetAttribute(k_lk_single<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
        KR_CUDA(cudaFuncSetAttribute(k_lk_roundtrip<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
        set1 = *smem;
    }
    return KR_OK;
}

}  // namespace

int krl_pyr_down(const uint8_t *src, int64_t pitch, int w, int h, uint8_t *dst, int64_t dst_pitch,
                 cudaStream_t s)
{
    PyrPair pp;
    pp.src[0] = pp.src[1] = src; pp.dst[0] = pp.dst[1] = dst;
    pp.src_pitch[0] = pp.src_pitch[1] = pitch; pp.dst_pitch[0] = pp.dst_pitch[1] = dst_pitch;
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    const int aligned = ((uintptr_t)src % 4 == 0) && (pitch % 4 == 0) && ((uintptr_t)dst % 2 == 0) &&
                        (dst_pitch % 2 == 0);
    return launch_pyr_down(pp, w, h, dw, dh, 1, aligned, s);
}

// buildOpticalFlowPyramid: a level is added while both halved sizes exceed the window.
int krl_build_pyramids(kr_ctx *ctx, const uint8_t *prev, int64_t pp, const uint8_t *next, int64_t np_,
                       int w, int h, int win, int max_level, KrLkArgs *a, cudaStream_t s)
{
    if (max_level >= KR_MAX_LEVELS) max_level = KR_MAX_LEVELS - 1;
    if (max_level < 0) max_level = 0;
    if ((int64_t)w * h > (int64_t)ctx->max_w * ctx->max_h || w > 65535 || h > 65535)
        return kr_set_error(KR_ERR_CAPACITY, "image %dx%d larger than the context (%dx%d)", w, h,
                            ctx->max_w, ctx->max_h);
    a->img[0][0] = prev; a->img[1][0] = next;
    a->pitch[0][0] = pp; a->pitch[1][0] = np_;
    a->w[0] = w; a->h[0] = h;
    a->win = win;
    int levels = 0;
    for (int l = 0; l < max_level; l++) {
        int nw = (a->w[l] + 1) / 2, nh = (a->h[l] + 1) / 2;
        if (nw <= win || nh <= win) break;
        // planes of one level are packed with a 128-byte aligned pitch
        int64_t pitch = ((int64_t)nw + 127) & ~(int64_t)127;
        a->w[l + 1] = nw; a->h[l + 1] = nh;
        a->img[0][l + 1] = ctx->d_pyr[0][l + 1]; a->img[1][l + 1] = ctx->d_pyr[1][l + 1];
        a->pitch[0][l + 1] = a->pitch[1][l + 1] = pitch;
        PyrPair q;
        q.src[0] = a->img[0][l]; q.src[1] = a->img[1][l];
        q.dst[0] = ctx->d_pyr[0][l + 1]; q.dst[1] = ctx->d_pyr[1][l + 1];
        q.src_pitch[0] = a->pitch[0][l]; q.src_pitch[1] = a->pitch[1][l];
        q.dst_pitch[0] = q.dst_pitch[1] = pitch;
        int aligned = (pitch % 2 == 0);
        for (int k = 0; k < 2; k++)
            aligned = aligned && ((uintptr_t)q.src[k] % 4 == 0) && (q.src_pitch[k] % 4 == 0) &&
                      ((uintptr_t)q.dst[k] % 2 == 0);
        KR_TRY(launch_pyr_down(q, a->w[l], a->h[l], nw, nh, 2, aligned, s));
        levels = l + 1;
    }
    a->levels = levels;
    return KR_OK;
}

// Level geometry of buildOpticalFlowPyramid for a w x h plane: level l >= 1 has
// ((w_{l-1} + 1) / 2, (h_{l-1} + 1) / 2), a 128-byte aligned pitch, and is added while both
// halved sizes exceed the window.  wl / hl / pl: [KR_MAX_LEVELS], entry 0 = the plane itself.
int krl_pyramid_geometry(int w, int h, int64_t pitch, int win, int max_level, int *levels, int *wl,
                         int *hl, int64_t *pl)
{
    if (max_level >= KR_MAX_LEVELS) max_level = KR_MAX_LEVELS - 1;
    if (max_level < 0) max_level = 0;
    wl[0] = w; hl[0] = h; pl[0] = pitch;
    int n = 0;
    for (int l = 0; l < max_level; l++) {
        const int nw = (wl[l] + 1) / 2, nh = (hl[l] + 1) / 2;
        if (nw <= win || nh <= win) break;
        wl[l + 1] = nw; hl[l + 1] = nh;
        pl[l + 1] = ((int64_t)nw + 127) & ~(int64_t)127;
        n = l + 1;
    }
    *levels = n;
    return KR_OK;
}

// Levels 1 .. levels of ONE plane into caller-owned planes dst[l] (pitch pl[l]).
int krl_pyramid_plane(const uint8_t *src, int levels, const int *wl, const int *hl, const int64_t *pl,
                      uint8_t *const *dst, cudaStream_t s)
{
    const uint8_t *cur = src;
    for (int l = 0; l < levels; l++) {
        PyrPair q;
        q.src[0] = q.src[1] = cur;
        q.dst[0] = q.dst[1] = dst[l + 1];
        q.src_pitch[0] = q.src_pitch[1] = pl[l];
        q.dst_pitch[0] = q.dst_pitch[1] = pl[l + 1];
        const int nw = wl[l + 1], nh = hl[l + 1];
        const int aligned = ((uintptr_t)cur % 4 == 0) && (pl[l] % 4 == 0) && ((uintptr_t)dst[l + 1] % 2 == 0) &&
                            (pl[l + 1] % 2 == 0);
        KR_TRY(launch_pyr_down(q, wl[l], hl[l], nw, nh, 1, aligned, s));
        cur = dst[l + 1];
    }
    return KR_OK;
}

int krl_lk_single(const KrLkArgs &a_in, const float *p0, int n, const int32_t *d_count, float *p1,
                  uint8_t *status, float *err, cudaStream_t s)
{
    KrLkArgs a = a_in;
    size_t smem;
    KR_TRY(lk_prepare(a, &smem));
    if (n <= 0) return KR_OK;
    int grid = (n + LK_WARPS - 1) / LK_WARPS;
    if (a.win == 25) k_lk_single<25><<<grid, LK_WARPS * 32, smem, s>>>(a, p0, n, d_count, p1, status, err);
    else k_lk_single<0><<<grid, LK_WARPS * 32, smem, s>>>(a, p0, n, d_count, p1, status, err);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

int krl_lk_roundtrip(const KrLkArgs &a_in, const float *p0, int n_cap, const uint32_t *d_count,
                     float back_thr, float *p1, float *dist, uint8_t *keep, cudaStream_t s)
{
    KrLkArgs a = a_in;
    size_t smem;
    KR_TRY(lk_prepare(a, &smem));
    if (n_cap <= 0) return KR_OK;
    int grid = (n_cap + LK_WARPS - 1) / LK_WARPS;
    if (grid > 65535 * 8) grid = 65535 * 8;
    if (a.win == 25)
        k_lk_roundtrip<25><<<grid, LK_WARPS * 32, smem, s>>>(a, p0, n_cap, d_count, back_thr, p1, dist, keep);
    else
        k_lk_roundtrip<0><<<grid, LK_WARPS * 32, smem, s>>>(a, p0, n_cap, d_count, back_thr, p1, dist, keep);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

int krl_emit_rows(kr_ctx *ctx, const float *p0, const float *p1, const float *dist,
                  const uint8_t *keep, int n_cap, const uint32_t *d_count, int sort_xy,
                  float back_thr, float x_off, float y_off, kr_rows rows, cudaStream_t s)
{
    k_clear_rowkeys<<<1, 1, 0, s>>>(ctx->d_stats);
    KR_LAUNCH_CHECK();
    int grid = (n_cap + 255) / 256;
    if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
    if (grid < 1) grid = 1;
    k_row_keys<<<grid, 256, 0, s>>>(p0, keep, n_cap, d_count, sort_xy, ctx->d_keys_b, ctx->d_stats);
    KR_LAUNCH_CHECK();
    // ascending (x0, y0, index): DataFrame.sort_values(["x0", "y0"]) (klt.py:348),
    // or plain index order (sort_xy == 0) = boolean-mask filtering (klt.py:150-153)
    KR_TRY(krl_sort_u64(ctx, ctx->d_keys_b, ctx->d_keys_a, &ctx->d_stats->n_rowkeys,
                        (int64_t)n_cap, 0, s));
    k_gather_rows<<<grid, 256, 0, s>>>(ctx->d_keys_a, p0, p1, dist, back_thr, x_off, y_off, rows,
                                       ctx->d_stats);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

// kr_corners.cu -- K3 (min-eigenvalue response + local-maximum candidates) and
// K4 (threshold, pre-selection, parallel min-distance NMS, ordering).
//
// Replaces cv2.goodFeaturesToTrack(img, maxCorners, qualityLevel, minDistance,
// mask, blockSize) as called at karios/matcher/klt.py:120 and :494.  Semantics
// follow SURVEY.md A.3 / A.4 (verified against cv2 4.13 by oracle/klt_oracle.c):
//   Sobel3 scaled by 1/(4*block*255) with OpenCV's FMA placement -> products ->
//   block x block unnormalised box sum in float64 -> (a+c) - sqrt((a-c)^2 + b^2)
//   -> masked max -> TOZERO at max*quality -> 3x3 local maxima (1-px border
//   excluded) -> order by (value desc, address desc) -> greedy min-distance.
//
// The greedy pass is sequential in OpenCV.  Here it is a fixed point: a
// candidate is ACCEPTED once every higher-priority candidate closer than
// minDistance is rejected, and REJECTED once one of them is accepted.  Decisions
// are final and depend only on higher-priority neighbours, so iterating to the
// fixed point (asynchronously, any order) yields exactly the sequential result;
// truncating the accepted set, in priority order, at maxCorners equals OpenCV's
// early exit.
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include "kr_internal.cuh"

namespace {

constexpr int EG_TW = 64, EG_TH = 32, EG_THREADS = 512;
constexpr int EG_RUN = 6;      // eig columns per horizontal sliding run
constexpr int EG_SEG = 9;      // eig rows per vertical sliding segment
constexpr int HIST_BINS = 4096;

struct EigGeom {
    int block, r0;             // window = [p - r0, p - r0 + block - 1]
    int PH, PW, PWS;           // product region (rows, cols, padded row stride)
    int XH, XW, XWS;           // pixel region
    int EW, EWS;               // eig region cols (+ padded stride); rows = EG_TH + 2
};

__host__ __device__ inline EigGeom eig_geom(int block)
{
    EigGeom g;
    g.block = block;
    g.r0 = block / 2;
    g.PH = EG_TH + 1 + block;
    g.PW = EG_TW + 1 + block;
    g.PWS = g.PW | 1;
    g.XH = g.PH + 2;
    g.XW = g.PW + 2;
    g.XWS = (g.XW + 3) & ~3;
    g.EW = EG_TW + 2;
    g.EWS = g.EW | 1;
    return g;
}

__host__ inline size_t eig_smem_bytes(const EigGeom &g)
{
    size_t hs = (size_t)3 * g.PH * g.EWS * sizeof(double);
    size_t pr = (size_t)3 * g.PH * g.PWS * sizeof(float);
    size_t eg = (size_t)(EG_TH + 2) * (g.EW + 1) * sizeof(float);
    size_t px = (size_t)g.XH * g.XWS;
    return hs + pr + eg + px + 64;
}

// K3.  One block = one EG_TW x EG_TH tile of the tile image.
//  P0 stage pixels (REFLECT_101)            P1 Sobel products (float32)
//  P2 horizontal box sums (float64 sliding)  P3 vertical box sums + eigenvalue
//  P4 masked max, 3x3 local maxima -> candidate keys (value bits << 32 | y*W+x)
__global__ void __launch_bounds__(EG_THREADS)
k_eig_candidates(const uint8_t *__restrict__ img, int64_t pitch, const uint8_t *__restrict__ mask,
                 int64_t mpitch, int w, int h, int block, float s, int tail_start,
                 float *__restrict__ eig_out, int64_t eig_pitch, uint64_t *__restrict__ cand,
                 uint32_t cand_cap, KrDevStats *st, int emit)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EigGeom g = eig_geom(block);
    double *hsum = reinterpret_cast<double *>(smem_raw);                         // [3][PH][EWS]
    float *prod = reinterpret_cast<float *>(hsum + (size_t)3 * g.PH * g.EWS);    // [3][PH][PWS]
    float *seig = prod + (size_t)3 * g.PH * g.PWS;                               // [TH+2][EW+1]
    uint8_t *pix = reinterpret_cast<uint8_t *>(seig + (size_t)(EG_TH + 2) * (g.EW + 1));  // [XH][XWS]
    __shared__ uint32_t s_warp_cnt[EG_THREADS / 32];
    __shared__ uint32_t s_base, s_maxenc;

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * EG_TW, y0 = blockIdx.y * EG_TH;
    const int px0 = x0 - 1 - g.r0, py0 = y0 - 1 - g.r0;     // global coords of product (0,0)
    if (tid == 0) s_maxenc = KR_ENC_NEG_INF;

    // ---- P0: pixels of the product region +- 1 ------------------------------
    for (int i = tid; i < g.XH * g.XW; i += EG_THREADS) {
        int ty = i / g.XW, tx = i - ty * g.XW;
        int gy = kr_reflect101(py0 - 1 + ty, h), gx = kr_reflect101(px0 - 1 + tx, w);
        pix[ty * g.XWS + tx] = __ldg(img + (int64_t)gy * pitch + gx);
    }
    __syncthreads();

    // ---- P1: Sobel products -------------------------------------------------
    const int plane = g.PH * g.PWS;
    for (int i = tid; i < g.PH * g.PW; i += EG_THREADS) {
        int py = i / g.PW, px = i - py * g.PW;
        int gy = py0 + py, gx = px0 + px;
        float xx, xy, yy;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
            const uint8_t *c = pix + (py + 1) * g.XWS + (px + 1);
            sobel_products(c[-g.XWS - 1], c[-g.XWS], c[-g.XWS + 1], c[-1], c[1], c[g.XWS - 1],
                           c[g.XWS], c[g.XWS + 1], s, gx >= tail_start, xx, xy, yy);
        } else {
            // box-filter border: the product AT the reflected position (not the
            // derivative of the reflected image: dx*dy would change sign)
            int ry = kr_reflect101(gy, h), rx = kr_reflect101(gx, w);
            int ym = kr_reflect101(ry - 1, h), yp = kr_reflect101(ry + 1, h);
            int xm = kr_reflect101(rx - 1, w), xp = kr_reflect101(rx + 1, w);
            const uint8_t *a = img + (int64_t)ym * pitch, *b = img + (int64_t)ry * pitch,
                          *c = img + (int64_t)yp * pitch;
            sobel_products(__ldg(a + xm), __ldg(a + rx), __ldg(a + xp), __ldg(b + xm), __ldg(b + xp),
                           __ldg(c + xm), __ldg(c + rx), __ldg(c + xp), s, rx >= tail_start, xx, xy, yy);
        }
        int o = py * g.PWS + px;
        prod[o] = xx;
        prod[plane + o] = xy;
        prod[2 * plane + o] = yy;
    }
    __syncthreads();

    // ---- P2: horizontal sums over `block` products, float64 ------------------
    const int hplane = g.PH * g.EWS;
    const int nrun = (g.EW + EG_RUN - 1) / EG_RUN;
    for (int i = tid; i < g.PH * nrun; i += EG_THREADS) {
        int r = i % g.PH, j = i / g.PH;
        int e0 = j * EG_RUN;
        int e1 = min(e0 + EG_RUN, g.EW);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float *p = prod + ch * plane + r * g.PWS;
            double *hrow = hsum + ch * hplane + r * g.EWS;
            double acc = 0.0;
            for (int k = 0; k < block; k++) acc += (double)p[e0 + k];
            hrow[e0] = acc;
            for (int e = e0 + 1; e < e1; e++) {
                acc += (double)p[e + block - 1];
                acc -= (double)p[e - 1];
                hrow[e] = acc;
            }
        }
    }
    __syncthreads();

    // ---- P3: vertical sums, eigenvalue --------------------------------------
    const int ER = EG_TH + 2;
    const int nseg = (ER + EG_SEG - 1) / EG_SEG;
    for (int i = tid; i < g.EW * nseg; i += EG_THREADS) {
        int c = i % g.EW, sg = i / g.EW;
        int r_beg = sg * EG_SEG, r_end = min(r_beg + EG_SEG, ER);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        const double *h0 = hsum + c, *h1 = hsum + hplane + c, *h2 = hsum + 2 * hplane + c;
        for (int k = 0; k < block; k++) {
            int o = (r_beg + k) * g.EWS;
            a0 += h0[o]; a1 += h1[o]; a2 += h2[o];
        }
        for (int r = r_beg; r < r_end; r++) {
            if (r > r_beg) {
                int on = (r + block - 1) * g.EWS, oo = (r - 1) * g.EWS;
                a0 += h0[on]; a0 -= h0[oo];
                a1 += h1[on]; a1 -= h1[oo];
                a2 += h2[on]; a2 -= h2[oo];
            }
            // calcMinEigenVal: plain float32, no contraction
            float a = __fmul_rn((float)a0, 0.5f), b = (float)a1, cc = __fmul_rn((float)a2, 0.5f);
            float t = __fsub_rn(a, cc);
            float e = __fsub_rn(__fadd_rn(a, cc),
                                __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
            seig[r * (g.EW + 1) + c] = e;
        }
    }
    __syncthreads();

    // ---- P4: masked max, local maxima, candidates ----------------------------
    uint32_t my_max = KR_ENC_NEG_INF;
    for (int it = 0; it < (EG_TW * EG_TH) / EG_THREADS; it++) {
        int i = it * EG_THREADS + tid;
        int ty = i / EG_TW, tx = i - ty * EG_TW;
        int gx = x0 + tx, gy = y0 + ty;
        bool inimg = gx < w && gy < h;
        bool is_cand = false;
        float v = 0.f;
        if (inimg) {
            const float *e = seig + (ty + 1) * (g.EW + 1) + (tx + 1);
            v = *e;
            if (eig_out) *(float *)((char *)eig_out + (int64_t)gy * eig_pitch + (int64_t)gx * 4) = v;
            bool mok = mask ? (__ldg(mask + (int64_t)gy * mpitch + gx) != 0) : true;
            if (mok) my_max = max(my_max, kr_f32_enc(v));
            if (emit && mok && v > 0.f && gx >= 1 && gy >= 1 && gx <= w - 2 && gy <= h - 2) {
                const int S = g.EW + 1;
                is_cand = v >= e[-S - 1] && v >= e[-S] && v >= e[-S + 1] && v >= e[-1] && v >= e[1] &&
                          v >= e[S - 1] && v >= e[S] && v >= e[S + 1];
            }
        }
        // block-aggregated append: one global atomic per block and iteration
        unsigned bal = __ballot_sync(0xffffffffu, is_cand);
        int lane = tid & 31, wid = tid >> 5;
        if (lane == 0) s_warp_cnt[wid] = __popc(bal);
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int k = 0; k < EG_THREADS / 32; k++) { uint32_t c = s_warp_cnt[k]; s_warp_cnt[k] = tot; tot += c; }
            s_base = tot ? atomicAdd(&st->n_cand, tot) : 0;
        }
        __syncthreads();
        if (is_cand) {
            uint32_t pos = s_base + s_warp_cnt[wid] + __popc(bal & ((1u << lane) - 1));
            if (pos < cand_cap)
                cand[pos] = ((uint64_t)__float_as_uint(v) << 32) | (uint32_t)(gy * w + gx);
            else
                st->overflow = 1;
        }
        __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
    if ((tid & 31) == 0 && my_max != KR_ENC_NEG_INF) atomicMax(&s_maxenc, my_max);
    __syncthreads();
    if (tid == 0 && s_maxenc != KR_ENC_NEG_INF) atomicMax(&st->eig_max_enc, s_maxenc);
}

// ---------------------------------------------------------------------------
// K3, streaming form (blockSize 15, images >= 16 x 16): no block synchronisation.
//
// One warp owns a strip of 128 image columns (4 per lane, one 32-bit load per
// lane and row) and walks down `seg` rows.  Per row and lane:
//   Sobel row/column filters from the new pixel row and two rows of history
//   (neighbour pixels across lanes by warp shuffle) -> 3 products x 4 columns;
//   vertical 15-row sums are running float64 sums per column (add the new
//   product row, subtract the one 15 rows back, kept as float32 in a per-warp
//   shared-memory ring); horizontal 15-column sums combine per-lane prefix /
//   suffix sums of the 4 columns with those of lanes l-2 .. l+2 (float64
//   shuffles); eigenvalue; 3x3 local maximum against two rows of history;
//   candidates are buffered per warp and appended with one atomic per flush.
// Out-of-image rows / columns of the box filter are REFLECT_101 of the product
// plane: the stream simply visits the reflected row / column with the roles of
// the two neighbours swapped, which reproduces the product AT the reflected
// position.  float64 sums of these float32 products are exact in all but
// ~1e-7 of the pixels, so the summation order is free.
constexpr int EC_WARPS = 4, EC_BLOCKS_PER_SM = 3, EC_OUTW = 104, EC_LEFT = 12, EC_CBUF = 256;
constexpr int EC_RING_F4 = 16 * 2 * 32;          // float4 per warp: 16 rows x (dx, dy) x 32 lanes

// exact int -> float for |i| < 2^22 on the ALU / FMA pipes (the XU pipe that I2F
// uses is the busiest unit of this kernel)
__device__ __forceinline__ float i2f_small(int i)
{
    return __fsub_rn(__int_as_float(0x4B400000 + i), 12582912.0f);
}

// BORDER = the strip touches the left / right image border, the SIMD-tail columns
// or unaligned planes: per-column reflection, tail rounding and bounds checks.
// Interior strips (the vast majority) take the lean path.
template <bool BORDER, bool HAS_MASK, bool DEBUG_EIG>
__device__ __forceinline__ void eig_stream_body(
    const uint8_t *__restrict__ img, int64_t pitch, const uint8_t *__restrict__ mask, int64_t mpitch,
    int w, int h, float s, int tail_start, float *__restrict__ eig_out, int64_t eig_pitch,
    uint64_t *__restrict__ cand, uint32_t cand_cap, KrDevStats *st, int emit, int xs, int ys, int ye,
    float4 *ring, uint64_t *cbuf, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int xb = xs - EC_LEFT + 4 * lane;                        // first (virtual) column of the lane
    const float s2 = 2.0f * s;
    int tc[4];
    bool crefl[4], ctail[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int vx = xb + j;
        tc[j] = BORDER ? kr_reflect101(vx, w) : vx;
        crefl[j] = BORDER && (vx < 0 || vx >= w);
        ctail[j] = BORDER && (tc[j] >= tail_start);
    }
    const bool out_lane = lane >= 3 && lane <= 28;

    auto load_row = [&](int r) -> uint32_t {
        const int tr = (r < 0) ? -r : ((r >= h) ? 2 * (h - 1) - r : r);     // single reflection (h >= 16)
        const uint8_t *prow = img + (int64_t)tr * pitch;
        if (!BORDER) return __ldg(reinterpret_cast<const uint32_t *>(prow + xb));
        return (uint32_t)__ldg(prow + tc[0]) | ((uint32_t)__ldg(prow + tc[1]) << 8) |
               ((uint32_t)__ldg(prow + tc[2]) << 16) | ((uint32_t)__ldg(prow + tc[3]) << 24);
    };
    auto load_mask = [&](int m) -> uint32_t {                                // 1 byte per column
        if (!HAS_MASK) return 0x01010101u;
        if (m < 0 || m >= h) return 0u;
        const uint8_t *mrow = mask + (int64_t)m * mpitch;
        if (!BORDER) return __ldg(reinterpret_cast<const uint32_t *>(mrow + xb));
        uint32_t mk = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int x = xb + j;
            if (x >= 0 && x < w) mk |= (uint32_t)__ldg(mrow + x) << (8 * j);
        }
        return mk;
    };

    int R1[4] = {0, 0, 0, 0}, R2[4] = {0, 0, 0, 0};        // rows r-1, r-2
    float T1[4] = {0, 0, 0, 0}, T2[4] = {0, 0, 0, 0};
    double cs[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) cs[c][j] = 0.0;
    float E1[6] = {0, 0, 0, 0, 0, 0};                  // eig row q-1: [L, 0..3, R]
    float H1[4] = {0, 0, 0, 0}, H2[4] = {0, 0, 0, 0};  // 3-wide maxima of eig rows q-1, q-2
    uint32_t my_max = KR_ENC_NEG_INF;
    int ccount = 0;                                                    // warp-uniform
    int nprod = 0;

    const int r_first = ys - 9, r_last = ye + 8;
    uint32_t pk_next = load_row(r_first);
    uint32_t mk_next = load_mask(r_first - 9);
    for (int r = r_first; r <= r_last; r++) {
        const uint32_t pk = pk_next;
        const uint32_t mk = mk_next;
        if (r < r_last) {                       // prefetch the next pixel / mask words
            pk_next = load_row(r + 1);
            mk_next = load_mask(r + 1 - 9);
        }
        // ---- pixel row r (virtual) -> row filters ---------------------------
        const uint32_t wl = __shfl_up_sync(FULL, pk, 1), wr = __shfl_down_sync(FULL, pk, 1);
        int q[6];
        q[0] = (int)(wl >> 24);
        q[1] = (int)(pk & 255u); q[2] = (int)((pk >> 8) & 255u); q[3] = (int)((pk >> 16) & 255u);
        q[4] = (int)(pk >> 24);
        q[5] = (int)(wr & 255u);
        float fq[6];
#pragma unroll
        for (int j = 0; j < 6; j++) fq[j] = i2f_small(q[j]);
        int R0[4];
        float T0[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int ia = q[j], ic = q[j + 2];                    // p[x-1], p[x+1] in virtual order
            float fa = fq[j], fc = fq[j + 2];
            if (BORDER && crefl[j]) {                        // reflected column: true neighbours swap
                int ti = ia; ia = ic; ic = ti;
                float tf = fa; fa = fc; fc = tf;
            }
            R0[j] = ic - ia;
            if (BORDER && ctail[j])
                T0[j] = __fadd_rn(__fadd_rn(__fmul_rn(s, fa), __fmul_rn(s2, fq[j + 1])), __fmul_rn(s, fc));
            else
                T0[j] = __fmaf_rn(s, fc, __fmaf_rn(s2, fq[j + 1], __fmul_rn(s, fa)));
        }
        if (r >= r_first + 2) {
            // ---- products of (virtual) row r-1 ------------------------------
            const bool rrefl = (r - 1) < 0 || (r - 1) >= h;
            float dxv[4], dyv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                dxv[j] = __fmaf_rn(s, i2f_small(R2[j] + R0[j]), __fmul_rn(s2, i2f_small(R1[j])));
                const float d0 = __fsub_rn(T0[j], T2[j]);
                dyv[j] = rrefl ? -d0 : d0;                   // exact negation
            }
            // ---- vertical running sums (float64); the ring keeps (dx, dy) of the
            //      last 15 product rows as float32, products are re-formed on exit
            nprod++;
            const int rd_slot = r & 15, wr_slot = (r - 1) & 15;
            if (nprod > 15) {
                const float4 ox = ring[(rd_slot * 2 + 0) * 32], oy = ring[(rd_slot * 2 + 1) * 32];
                const float odx[4] = {ox.x, ox.y, ox.z, ox.w}, ody[4] = {oy.x, oy.y, oy.z, oy.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    cs[0][j] -= (double)__fmul_rn(odx[j], odx[j]);
                    cs[1][j] -= (double)__fmul_rn(odx[j], ody[j]);
                    cs[2][j] -= (double)__fmul_rn(ody[j], ody[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                cs[0][j] += (double)__fmul_rn(dxv[j], dxv[j]);
                cs[1][j] += (double)__fmul_rn(dxv[j], dyv[j]);
                cs[2][j] += (double)__fmul_rn(dyv[j], dyv[j]);
            }
            ring[(wr_slot * 2 + 0) * 32] = make_float4(dxv[0], dxv[1], dxv[2], dxv[3]);
            ring[(wr_slot * 2 + 1) * 32] = make_float4(dyv[0], dyv[1], dyv[2], dyv[3]);
            if (nprod >= 15) {
                // ---- horizontal 15-column sums -> eigenvalue row r - 8 --------
                float E0[6];
                {
                    double bx[3][4];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double c0 = cs[c][0], c1 = cs[c][1], c2 = cs[c][2], c3 = cs[c][3];
                        const double P2 = c0 + c1, P3 = P2 + c2, qq = P3 + c3;
                        const double S2 = c2 + c3, S3 = S2 + c1;
                        const double mid = __shfl_up_sync(FULL, qq, 1) + qq + __shfl_down_sync(FULL, qq, 1);
                        const double a3 = __shfl_up_sync(FULL, S3, 2), a2 = __shfl_up_sync(FULL, S2, 2),
                                     a1 = __shfl_up_sync(FULL, c3, 2);
                        const double b1 = __shfl_down_sync(FULL, c0, 2), b2 = __shfl_down_sync(FULL, P2, 2),
                                     b3 = __shfl_down_sync(FULL, P3, 2);
                        bx[c][0] = mid + a3;
                        bx[c][1] = (mid + a2) + b1;
                        bx[c][2] = (mid + a1) + b2;
                        bx[c][3] = mid + b3;
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) E0[j + 1] = eig_from_sums(bx[0][j], bx[1][j], bx[2][j]);
                }
                E0[0] = __shfl_up_sync(FULL, E0[4], 1);
                E0[5] = __shfl_down_sync(FULL, E0[1], 1);
                float H0[4];                                   // 3-wide row maxima of the new row
#pragma unroll
                for (int j = 0; j < 4; j++) H0[j] = fmaxf(fmaxf(E0[j], E0[j + 1]), E0[j + 2]);
                // ---- 3x3 local maxima of row m = r - 9: eig == dilate(eig) ------
                const int m = r - 9;
                if (m >= ys && m < ye) {                       // warp-uniform
                    const bool row_ok = emit && m >= 1 && m <= h - 2;
                    unsigned cm = 0;                            // candidate bits of the 4 columns
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int x = xb + j;
                        const float v = E1[j + 1];
                        const bool inimg = out_lane && (!BORDER || x < w);
                        const bool mok = !HAS_MASK || ((mk >> (8 * j)) & 255u) != 0;
                        if (inimg) {
                            if (DEBUG_EIG) *(float *)((char *)eig_out + (int64_t)m * eig_pitch + (int64_t)x * 4) = v;
                            if (mok) my_max = max(my_max, kr_f32_enc(v));
                        }
                        const float dil = fmaxf(fmaxf(H2[j], H1[j]), H0[j]);
                        const bool is_c = row_ok && inimg && mok && v > 0.f && v == dil &&
                                          (!BORDER || (x >= 1 && x <= w - 2));
                        if (is_c) cm |= 1u << j;
                    }
                    if (__any_sync(FULL, cm != 0)) {
                        if (ccount > EC_CBUF - 128) {              // flush the warp buffer
                            uint32_t base = 0;
                            if (lane == 0) base = atomicAdd(&st->n_cand, (uint32_t)ccount);
                            base = __shfl_sync(FULL, base, 0);
                            __syncwarp();
                            for (int k = lane; k < ccount; k += 32) {
                                if (base + k < cand_cap) cand[base + k] = cbuf[k]; else st->overflow = 1;
                            }
                            __syncwarp();
                            ccount = 0;
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const bool is_c = (cm >> j) & 1u;
                            const unsigned bal = __ballot_sync(FULL, is_c);
                            if (is_c)
                                cbuf[ccount + __popc(bal & ((1u << lane) - 1))] =
                                    ((uint64_t)__float_as_uint(E1[j + 1]) << 32) | (uint32_t)(m * w + xb + j);
                            ccount += __popc(bal);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 6; j++) E1[j] = E0[j];
#pragma unroll
                for (int j = 0; j < 4; j++) { H2[j] = H1[j]; H1[j] = H0[j]; }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) { R2[j] = R1[j]; R1[j] = R0[j]; T2[j] = T1[j]; T1[j] = T0[j]; }
    }
    // final flush, masked maximum
    __syncwarp();
    if (ccount > 0) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&st->n_cand, (uint32_t)ccount);
        base = __shfl_sync(FULL, base, 0);
        for (int k = lane; k < ccount; k += 32) {
            if (base + k < cand_cap) cand[base + k] = cbuf[k]; else st->overflow = 1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(FULL, my_max, o));
    if (lane == 0 && my_max != KR_ENC_NEG_INF) atomicMax(&st->eig_max_enc, my_max);
}

template <bool HAS_MASK, bool DEBUG_EIG>
__global__ void __launch_bounds__(EC_WARPS * 32, EC_BLOCKS_PER_SM)
k_eig_stream(const uint8_t *__restrict__ img, int64_t pitch, const uint8_t *__restrict__ mask,
             int64_t mpitch, int w, int h, float s, int tail_start, float *__restrict__ eig_out,
             int64_t eig_pitch, uint64_t *__restrict__ cand, uint32_t cand_cap, KrDevStats *st, int emit,
             int seg, int aligned)
{
    extern __shared__ __align__(16) unsigned char ec_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4 *ring = reinterpret_cast<float4 *>(ec_smem) + (size_t)wid * EC_RING_F4 + lane;     // [16][2][32]
    uint64_t *cbuf = reinterpret_cast<uint64_t *>(ec_smem + (size_t)EC_WARPS * EC_RING_F4 * 16) +
                     (size_t)wid * EC_CBUF;
    const int xs = (blockIdx.x * EC_WARPS + wid) * EC_OUTW;        // first output column
    if (xs >= w) return;
    const int ys = blockIdx.y * seg, ye = min(ys + seg, h);
    // lean path: all 128 columns inside the image, none in the SIMD tail, aligned planes
    const bool interior = aligned && (xs - EC_LEFT >= 0) && (xs - EC_LEFT + 128 <= w) &&
                          (xs - EC_LEFT + 128 <= tail_start);
    if (interior)
        eig_stream_body<false, HAS_MASK, DEBUG_EIG>(img, pitch, mask, mpitch, w, h, s, tail_start, eig_out,
                                                    eig_pitch, cand, cand_cap, st, emit, xs, ys, ye, ring,
                                                    cbuf, lane);
    else
        eig_stream_body<true, HAS_MASK, DEBUG_EIG>(img, pitch, mask, mpitch, w, h, s, tail_start, eig_out,
                                                   eig_pitch, cand, cand_cap, st, emit, xs, ys, ye, ring,
                                                   cbuf, lane);
}

// threshold = float(maxVal * qualityLevel) (cv::threshold takes a double and
// narrows it); candidates carry v > 0 only, so a non-positive threshold keeps all.
__device__ __forceinline__ void threshold_bits(const KrDevStats *st, double quality, uint32_t &thr_bits,
                                               uint32_t &max_bits)
{
    uint32_t enc = st->eig_max_enc;
    float maxv = (enc == KR_ENC_NEG_INF) ? 0.f : kr_f32_dec_bits(enc, 0);
    float thr = (float)((double)maxv * quality);
    thr_bits = (thr > 0.f) ? __float_as_uint(thr) : 0u;
    max_bits = (maxv > 0.f) ? __float_as_uint(maxv) : 0u;
}

__device__ __forceinline__ uint32_t hist_shift_of(uint32_t thr_bits, uint32_t max_bits)
{
    uint32_t span = (max_bits > thr_bits) ? (max_bits - thr_bits) : 0u;
    int bits = 32 - __clz(span);            // span < 2^bits
    return (bits > 12) ? (uint32_t)(bits - 12) : 0u;
}

__global__ void __launch_bounds__(256)
k_cand_hist(const uint64_t *__restrict__ cand, KrDevStats *st, uint32_t *__restrict__ hist,
            double quality, uint32_t cap)
{
    __shared__ uint32_t sh[HIST_BINS];
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) sh[i] = 0;
    uint32_t thr_bits, max_bits;
    threshold_bits(st, quality, thr_bits, max_bits);
    uint32_t shift = hist_shift_of(thr_bits, max_bits);
    uint32_t n = min(st->n_cand, cap);
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t v = (uint32_t)(cand[i] >> 32);
        if (v > thr_bits) atomicAdd(&sh[min((v - thr_bits) >> shift, (uint32_t)HIST_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->thr_bits = thr_bits;
        st->hist_shift = shift;
    }
}

// Pick the value cut-off: the smallest set of top histogram bins holding at
// least `target` candidates (or everything above the threshold).
__global__ void __launch_bounds__(1024)
k_cutoff(KrDevStats *st, uint32_t *hist, uint32_t target, int select_all)
{
    __shared__ uint32_t suf[1024];
    __shared__ int best;
    const int t = threadIdx.x;
    uint32_t h0 = hist[4 * t], h1 = hist[4 * t + 1], h2 = hist[4 * t + 2], h3 = hist[4 * t + 3];
    hist[4 * t] = hist[4 * t + 1] = hist[4 * t + 2] = hist[4 * t + 3] = 0;     // ready for the next call
    suf[t] = h0 + h1 + h2 + h3;
    if (t == 0) best = 0;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                    // inclusive suffix sums
        uint32_t add = (t + o < 1024) ? suf[t + o] : 0;
        __syncthreads();
        suf[t] += add;
        __syncthreads();
    }
    uint32_t total = suf[0];
    uint32_t above = (t + 1 < 1024) ? suf[t + 1] : 0;       // candidates in bins > 4t+3
    // suffix count from bin b: s3 = above+h3, s2 = s3+h2, ...
    uint32_t s3 = above + h3, s2 = s3 + h2, s1 = s2 + h1, s0 = s1 + h0;
    int b = -1;
    if (s3 >= target) b = 4 * t + 3;
    else if (s2 >= target) b = 4 * t + 2;
    else if (s1 >= target) b = 4 * t + 1;
    else if (s0 >= target) b = 4 * t;
    if (b >= 0) atomicMax(&best, b);
    __syncthreads();
    if (t == 0) {
        uint32_t cut = st->thr_bits + 1;
        if (!select_all && total > target && best > 0) {
            uint32_t c2 = st->thr_bits + ((uint32_t)best << st->hist_shift);
            if (c2 > cut) cut = c2;
        }
        st->cut_bits = cut;
        // tier 1 dropped whole rows below its running estimate of this cut-off: the selection is
        // complete only if the cut-off chosen here is not below that estimate
        if (st->cut_est_bits && cut < st->cut_est_bits) st->fast_fallback = 1;
        st->cut_applied = (cut != st->thr_bits + 1) ? 1u : 0u;
        st->n_thr = total;
    }
}

__global__ void __launch_bounds__(256)
k_select(const uint64_t *__restrict__ cand, uint64_t *__restrict__ keys, KrDevStats *st, uint32_t cap,
         uint32_t out_cap)
{
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_base;
    const uint32_t n = min(st->n_cand, cap);
    const uint32_t cut = st->cut_bits;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        uint32_t i = i0 + threadIdx.x;
        uint64_t key = (i < n) ? cand[i] : 0ull;
        bool keep = (i < n) && ((uint32_t)(key >> 32) >= cut);
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int k = 0; k < 8; k++) { uint32_t c = s_cnt[k]; s_cnt[k] = tot; tot += c; }
            s_base = tot ? atomicAdd(&st->n_sel, tot) : 0;
        }
        __syncthreads();
        if (keep) {
            uint32_t pos = s_base + s_cnt[wid] + __popc(bal & ((1u << lane) - 1));
            if (pos < out_cap) keys[pos] = key; else st->overflow = 1;
        }
        __syncthreads();
    }
}

// ---- grid-wide barrier for the persistent NMS kernel -------------------------
__device__ __forceinline__ void grid_sync(uint32_t *bar, uint32_t nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        volatile uint32_t *gen = bar + 1;
        uint32_t g = *gen;
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}

// K4.  Persistent, co-resident grid (cooperative launch).  state: 0 undecided,
// 1 accepted, 2 rejected.  Cell lists (linked through `next`) hold every
// selected candidate, cell size = round(minDistance) as in OpenCV.
__global__ void __launch_bounds__(256)
k_nms(const uint64_t *__restrict__ keys, uint32_t *__restrict__ xy, uint8_t *state,
      int32_t *__restrict__ next, int32_t *head, uint64_t *__restrict__ accepted, int w, int cell,
      int gw, int gh, double md2, KrDevStats *st, uint32_t key_cap, uint32_t max_corners,
      uint32_t max_rounds)
{
    const uint32_t n = min(st->n_sel, key_cap);
    const uint32_t nb = gridDim.x;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = nb * blockDim.x;
    uint32_t *bar = st->barrier;

    for (uint32_t i = gtid; i < n; i += gstride) {
        uint32_t idx = (uint32_t)keys[i];
        uint32_t y = idx / (uint32_t)w, x = idx - y * (uint32_t)w;
        xy[i] = x | (y << 16);
        state[i] = 0;
        int c = (int)(y / cell) * gw + (int)(x / cell);
        next[i] = atomicExch(&head[c], (int32_t)i);
    }
    grid_sync(bar, nb);

    volatile uint8_t *vstate = state;
    uint32_t round = 0;
    for (; round < max_rounds; round++) {
        uint32_t pending_local = 0;
        for (uint32_t i = gtid; i < n; i += gstride) {
            if (vstate[i] != 0) continue;
            const uint64_t ki = keys[i];
            const uint32_t p = xy[i];
            const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
            const int xc = x / cell, yc = y / cell;
            const int x1 = max(xc - 1, 0), y1 = max(yc - 1, 0);
            const int x2 = min(xc + 1, gw - 1), y2 = min(yc + 1, gh - 1);
            bool rejected = false, pending = false;
            for (int yy = y1; yy <= y2 && !rejected; yy++)
                for (int xx = x1; xx <= x2 && !rejected; xx++)
                    for (int32_t j = head[yy * gw + xx]; j >= 0; j = next[j]) {
                        if ((uint32_t)j == i) continue;
                        if (keys[j] < ki) continue;                 // lower priority: irrelevant
                        uint32_t q = xy[j];
                        int ddx = x - (int)(q & 0xffffu), ddy = y - (int)(q >> 16);
                        if ((double)(ddx * ddx + ddy * ddy) >= md2) continue;
                        uint8_t sj = vstate[j];
                        if (sj == 1) { rejected = true; break; }
                        if (sj == 0) pending = true;
                    }
            if (rejected) vstate[i] = 2;
            else if (!pending) vstate[i] = 1;
            else pending_local++;
        }
        // any thread with pending work bumps this round's counter
        unsigned any = __ballot_sync(0xffffffffu, pending_local != 0);
        if ((threadIdx.x & 31) == 0 && any) atomicAdd(&st->undecided[round % 3], 1u);
        if (gtid == 0) st->undecided[(round + 1) % 3] = 0;
        grid_sync(bar, nb);
        if (*((volatile uint32_t *)&st->undecided[round % 3]) == 0) break;
    }

    // unordered compaction of the accepted keys (sorted afterwards)
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gstride) {
        uint32_t i = i0 + threadIdx.x;
        bool keep = (i < n) && (vstate[i] == 1);
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int k = 0; k < 8; k++) { uint32_t c = s_cnt[k]; s_cnt[k] = tot; tot += c; }
            s_base = tot ? atomicAdd(&st->n_acc, tot) : 0;
        }
        __syncthreads();
        if (keep) accepted[s_base + s_cnt[wid] + __popc(bal & ((1u << lane) - 1))] = keys[i];
        __syncthreads();
    }
    grid_sync(bar, nb);
    if (gtid == 0) {
        st->nms_rounds = round + 1;
        uint32_t nacc = st->n_acc;
        bool enough = (max_corners > 0) && (nacc >= max_corners);
        // the selection was cut short: in the two-tier path n_sel counts exact survivors
        // and n_thr possible candidates, so the cut itself is the criterion there
        const bool cut_short = st->fast_mode ? (st->cut_applied != 0) : (st->n_sel < st->n_thr);
        if (!enough && cut_short) st->select_incomplete = 1;
        st->undecided[0] = st->undecided[1] = st->undecided[2] = 0;
    }
}

// K4 without co-residency (KR_NMS_PLAIN=1; measured slower than the cooperative kernel, 0.139
// vs 0.088 ms, kept as an alternative): the same fixed-point iteration as k_nms, one launch per
// round; a round that finds nothing undecided left by its predecessor returns at once.  When
// NMS_ROUNDS launches do not reach the fixed point (long chains of mutually close candidates),
// select_incomplete is raised and the caller's exact re-run uses the cooperative kernel.
constexpr int NMS_ROUNDS = 16;

__global__ void __launch_bounds__(256)
k_nms_init(const uint64_t *__restrict__ keys, uint32_t *__restrict__ xy, uint8_t *state,
           int32_t *__restrict__ next, int32_t *head, int w, int cell, int gw, KrDevStats *st,
           uint32_t key_cap)
{
    const uint32_t n = min(st->n_sel, key_cap);
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    for (uint32_t i = gtid; i < n; i += gstride) {
        uint32_t idx = (uint32_t)keys[i];
        uint32_t y = idx / (uint32_t)w, x = idx - y * (uint32_t)w;
        xy[i] = x | (y << 16);
        state[i] = 0;
        int c = (int)(y / cell) * gw + (int)(x / cell);
        next[i] = atomicExch(&head[c], (int32_t)i);
    }
    if (gtid < NMS_ROUNDS) st->nms_pending[gtid] = 0;
}

// round r: reads the states its predecessors left (plus whatever this round already decided:
// decisions are final, so a fresher state only saves a round), counts what stays undecided
__global__ void __launch_bounds__(256)
k_nms_round(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ xy, uint8_t *state,
            const int32_t *__restrict__ next, const int32_t *__restrict__ head, int cell, int gw, int gh,
            double md2, KrDevStats *st, uint32_t key_cap, uint32_t round)
{
    if (round > 0 && *((volatile uint32_t *)&st->nms_pending[round - 1]) == 0) return;   // fixed point reached
    const uint32_t n = min(st->n_sel, key_cap);
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    volatile uint8_t *vstate = state;
    uint32_t pending_local = 0;
    for (uint32_t i = gtid; i < n; i += gstride) {
        if (vstate[i] != 0) continue;
        const uint64_t ki = keys[i];
        const uint32_t p = xy[i];
        const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
        const int xc = x / cell, yc = y / cell;
        const int x1 = max(xc - 1, 0), y1 = max(yc - 1, 0);
        const int x2 = min(xc + 1, gw - 1), y2 = min(yc + 1, gh - 1);
        bool rejected = false, pending = false;
        for (int yy = y1; yy <= y2 && !rejected; yy++)
            for (int xx = x1; xx <= x2 && !rejected; xx++)
                for (int32_t j = head[yy * gw + xx]; j >= 0; j = next[j]) {
                    if ((uint32_t)j == i) continue;
                    if (keys[j] < ki) continue;                 // lower priority: irrelevant
                    uint32_t q = xy[j];
                    int ddx = x - (int)(q & 0xffffu), ddy = y - (int)(q >> 16);
                    if ((double)(ddx * ddx + ddy * ddy) >= md2) continue;
                    uint8_t sj = vstate[j];
                    if (sj == 1) { rejected = true; break; }
                    if (sj == 0) pending = true;
                }
        if (rejected) vstate[i] = 2;
        else if (!pending) vstate[i] = 1;
        else pending_local++;
    }
    unsigned any = __ballot_sync(0xffffffffu, pending_local != 0);
    if ((threadIdx.x & 31) == 0 && any) atomicAdd(&st->nms_pending[round], 1u);
    if (gtid == 0) st->nms_rounds = round + 1;
}

__global__ void __launch_bounds__(256)
k_nms_compact(const uint64_t *__restrict__ keys, const uint8_t *__restrict__ state,
              uint64_t *__restrict__ accepted, KrDevStats *st, uint32_t key_cap)
{
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_base;
    const uint32_t n = min(st->n_sel, key_cap);
    const uint32_t gstride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gstride) {
        uint32_t i = i0 + threadIdx.x;
        bool keep = (i < n) && (state[i] == 1);
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int k = 0; k < 8; k++) { uint32_t c = s_cnt[k]; s_cnt[k] = tot; tot += c; }
            s_base = tot ? atomicAdd(&st->n_acc, tot) : 0;
        }
        __syncthreads();
        if (keep) accepted[s_base + s_cnt[wid] + __popc(bal & ((1u << lane) - 1))] = keys[i];
        __syncthreads();
    }
}

__global__ void k_nms_final(KrDevStats *st, uint32_t max_corners)
{
    const uint32_t nacc = st->n_acc;
    const bool enough = (max_corners > 0) && (nacc >= max_corners);
    const bool cut_short = st->fast_mode ? (st->cut_applied != 0) : (st->n_sel < st->n_thr);
    if (!enough && cut_short) st->select_incomplete = 1;
    // no fixed point within NMS_ROUNDS launches: undecided candidates were left out
    if (st->nms_pending[NMS_ROUNDS - 1] != 0) st->select_incomplete = 1;
}

// minDistance < 1: OpenCV skips the grid and takes the sorted list as is.
__global__ void k_accept_all(const uint64_t *__restrict__ keys, uint64_t *__restrict__ accepted,
                             KrDevStats *st, uint32_t key_cap)
{
    const uint32_t n = min(st->n_sel, key_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        accepted[i] = keys[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_acc = n;
        if (st->fast_mode ? (st->cut_applied != 0) : (st->n_sel < st->n_thr)) st->select_incomplete = 1;
    }
}

__global__ void k_emit_corners(const uint64_t *__restrict__ sorted, KrDevStats *st, int w,
                               uint32_t max_corners, uint32_t capacity, float *__restrict__ out_xy,
                               int32_t *d_count, uint32_t key_cap)
{
    uint32_t n = min(st->n_acc, key_cap);
    if (max_corners > 0 && n > max_corners) n = max_corners;
    if (n > capacity) n = capacity;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t idx = (uint32_t)sorted[i];
        uint32_t y = idx / (uint32_t)w, x = idx - y * (uint32_t)w;
        out_xy[2 * i] = (float)x;
        out_xy[2 * i + 1] = (float)y;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_corners = n;
        if (d_count) *d_count = (int32_t)n;
    }
}

__global__ void k_clear_counts(KrDevStats *st)
{
    st->n_cand = st->n_thr = st->n_sel = st->n_acc = st->n_corners = 0;
    st->eig_max_enc = KR_ENC_NEG_INF;
    st->select_incomplete = 0;
    st->nms_rounds = 0;
    st->lmax_enc = st->umax_enc = KR_ENC_NEG_INF;
    st->n_maxlist = st->n_exact = 0;
    st->cut_applied = st->fast_mode = st->fast_fallback = 0;
    st->cut_est_bits = st->fa_rows = st->fa_skipped = st->fa_done = st->fa_rows_in = 0;
}

__global__ void k_set_fast_mode(KrDevStats *st) { st->fast_mode = 1; }

}  // namespace

int krl_good_features(kr_ctx *ctx, const uint8_t *img, int64_t pitch, const uint8_t *mask,
                      int64_t mask_pitch, int w, int h, int max_corners, double quality,
                      double min_distance, int block, int tail_mode, int select_all, float *eig_out,
                      int64_t eig_pitch, float *out_xy, int capacity, int32_t *d_count,
                      cudaStream_t s)
{
    if (w < 1 || h < 1 || w > 65535 || h > 65535)
        return kr_set_error(KR_ERR_INVALID, "image size %dx%d out of range", w, h);
    if (block < 1 || block > 31)
        return kr_set_error(KR_ERR_UNSUPPORTED, "blockSize %d not supported (1..31)", block);
    if ((int64_t)w * h > (int64_t)ctx->max_w * ctx->max_h)
        return kr_set_error(KR_ERR_CAPACITY, "image %dx%d larger than the context (%dx%d)", w, h,
                            ctx->max_w, ctx->max_h);
    const EigGeom g = eig_geom(block);
    const size_t smem = eig_smem_bytes(g);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        KR_CUDA(cudaFuncSetAttribute(k_eig_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        smem_set = smem;
    }
    const float scale = (float)(1.0 / (4.0 * (double)block * 255.0));
    int tail_start = w;
    if (tail_mode > 0) tail_start = tail_mode * (w / tail_mode);
    const int emit = out_xy != nullptr;

    k_clear_counts<<<1, 1, 0, s>>>(ctx->d_stats);
    KR_LAUNCH_CHECK();
    // two-tier response: integer bounds for every pixel, OpenCV's arithmetic only where
    // the answer depends on it (kr_corner_fast.cu); needs a bounded maxCorners
    const bool fast = block == 15 && w >= 16 && h >= 16 && emit && !eig_out && !select_all &&
                      max_corners > 0 && !ctx->no_fast_corners && pitch < (1ll << 31) &&
                      mask_pitch < (1ll << 31);
    if (fast) {
        k_set_fast_mode<<<1, 1, 0, s>>>(ctx->d_stats);
        KR_LAUNCH_CHECK();
        KR_TRY(krl_eig_fast(ctx, img, pitch, mask, mask_pitch, w, h, scale, tail_start,
                            2u * (uint32_t)max_corners + 4096u, s));
    } else if (block == 15 && w >= 16 && h >= 16) {
        const size_t esm = (size_t)EC_WARPS * (EC_RING_F4 * 16 + EC_CBUF * 8);
        static bool ec_set = false;
        if (!ec_set) {
            KR_CUDA(cudaFuncSetAttribute(k_eig_stream<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));
            KR_CUDA(cudaFuncSetAttribute(k_eig_stream<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));
            KR_CUDA(cudaFuncSetAttribute(k_eig_stream<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));
            KR_CUDA(cudaFuncSetAttribute(k_eig_stream<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));
            ec_set = true;
        }
        int aligned = ((uintptr_t)img % 4 == 0) && (pitch % 4 == 0);
        if (mask) aligned = aligned && ((uintptr_t)mask % 4 == 0) && (mask_pitch % 4 == 0);
        // rows per warp: whole waves of co-resident blocks (EC_BLOCKS_PER_SM per SM), each
        // segment pays 18 warm-up rows -- pick the cheaper of two wave counts
        const int sb = (w + EC_WARPS * EC_OUTW - 1) / (EC_WARPS * EC_OUTW);
        int best_seg = h, best_cost = INT_MAX;
        const int slots = ctx->num_sms * EC_BLOCKS_PER_SM;
        const int w0 = (int)(((int64_t)sb * ((h + 255) / 256) + slots - 1) / slots);
        for (int waves = (w0 > 1 ? w0 - 1 : 1); waves <= w0 + 1; waves++) {
            int segs = waves * slots / sb;
            if (segs < 1) segs = 1;
            int sg = (h + segs - 1) / segs;
            if (sg < 32) sg = 32;
            int nseg = (h + sg - 1) / sg;
            int wv = (sb * nseg + slots - 1) / slots;
            int cost = wv * (sg + 18);
            if (cost < best_cost) { best_cost = cost; best_seg = sg; }
        }
        const int seg = best_seg;
        dim3 grid(sb, (h + seg - 1) / seg);
#define KR_EIG_LAUNCH(M, D)                                                                        \
    k_eig_stream<M, D><<<grid, EC_WARPS * 32, esm, s>>>(img, pitch, mask, mask_pitch, w, h, scale,     \
                                                        tail_start, eig_out, eig_pitch, ctx->d_cand,   \
                                                        (uint32_t)ctx->cand_cap, ctx->d_stats, emit,   \
                                                        seg, aligned)
        if (mask && eig_out) KR_EIG_LAUNCH(true, true);
        else if (mask) KR_EIG_LAUNCH(true, false);
        else if (eig_out) KR_EIG_LAUNCH(false, true);
        else KR_EIG_LAUNCH(false, false);
#undef KR_EIG_LAUNCH
    } else {
        // generic tile kernel: any blockSize <= 31, any image size
        dim3 grid((w + EG_TW - 1) / EG_TW, (h + EG_TH - 1) / EG_TH);
        k_eig_candidates<<<grid, EG_THREADS, smem, s>>>(img, pitch, mask, mask_pitch, w, h, block, scale,
                                                       tail_start, eig_out, eig_pitch, ctx->d_cand,
                                                       (uint32_t)ctx->cand_cap, ctx->d_stats, emit);
    }
    KR_LAUNCH_CHECK();
    KR_MARK(ctx, 4, s);
    if (!emit) return KR_OK;

    const uint32_t cap = (uint32_t)ctx->cand_cap;
    const int sgrid = ctx->num_sms * 4;
    if (fast) {
        KR_TRY(krl_cand_hist_fast(ctx, quality, scale, s));
    } else {
        k_cand_hist<<<sgrid, 256, 0, s>>>(ctx->d_cand, ctx->d_stats, ctx->d_hist, quality, cap);
        KR_LAUNCH_CHECK();
    }
    uint32_t target = 0;
    if (max_corners <= 0) select_all = 1;
    else target = 2u * (uint32_t)max_corners + 4096u;
    k_cutoff<<<1, 1024, 0, s>>>(ctx->d_stats, ctx->d_hist, target, select_all);
    KR_LAUNCH_CHECK();
    if (fast) {
        // possible candidates above the cut -> exact verdicts and exact keys
        k_select<<<sgrid, 256, 0, s>>>(ctx->d_cand, ctx->d_keys_b, ctx->d_stats, cap, cap);
        KR_LAUNCH_CHECK();
        KR_TRY(krl_exact_cands(ctx, img, pitch, w, h, scale, tail_start, quality, (int)(target + target / 4),
                               ctx->d_keys_b, ctx->d_keys_a, s));
    } else {
        k_select<<<sgrid, 256, 0, s>>>(ctx->d_cand, ctx->d_keys_a, ctx->d_stats, cap, cap);
        KR_LAUNCH_CHECK();
    }
    KR_MARK(ctx, 5, s);

    if (min_distance >= 1.0) {
        int cell = (int)rint(min_distance);
        if (cell < 1) cell = 1;
        int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
        if ((int64_t)gw * gh > ctx->cell_cap) {      // rare (tiny minDistance): grow, synchronously
            KR_CUDA(cudaStreamSynchronize(s));
            cudaFree(ctx->d_cell_head);
            ctx->d_cell_head = nullptr;
            ctx->cell_cap = 0;
            void *q = nullptr;
            if (cudaMalloc(&q, (size_t)gw * gh * sizeof(int32_t)) != cudaSuccess) {
                cudaGetLastError();
                return kr_set_error(KR_ERR_NOMEM, "NMS grid %dx%d does not fit in device memory", gw, gh);
            }
            ctx->d_cell_head = (int32_t *)q;
            ctx->cell_cap = (int64_t)gw * gh;
        }
        KR_CUDA(cudaMemsetAsync(ctx->d_cell_head, 0xff, (size_t)gw * gh * sizeof(int32_t), s));
        const uint64_t *keys = ctx->d_keys_a;
        uint32_t *xy = ctx->d_xy;
        uint8_t *state = ctx->d_state;
        int32_t *next = ctx->d_next, *head = ctx->d_cell_head;
        uint64_t *acc = ctx->d_keys_b;
        double md2 = min_distance * min_distance;
        KrDevStats *st = ctx->d_stats;
        uint32_t key_cap = cap, mc = (max_corners > 0) ? (uint32_t)max_corners : 0u;
        uint32_t max_rounds = 1u << 20;
        static const bool plain_nms = getenv("KR_NMS_PLAIN") != nullptr;
        if (!plain_nms || select_all || ctx->force_select_all) {
            // default, and always for the exact re-run / unlimited corners: any number of rounds,
            // co-resident grid
            void *args[] = {&keys, &xy, &state, &next, &head, &acc, &w, &cell, &gw, &gh,
                            &md2, &st, &key_cap, &mc, &max_rounds};
            KR_CUDA(cudaLaunchCooperativeKernel((const void *)k_nms, dim3(ctx->nms_grid), dim3(256), args,
                                                0, s));
            kr_note_launch();
        } else {
            const int ng = ctx->num_sms * 2;
            k_nms_init<<<ng, 256, 0, s>>>(keys, xy, state, next, head, w, cell, gw, st, key_cap);
            KR_LAUNCH_CHECK();
            for (int r = 0; r < NMS_ROUNDS; r++) {
                k_nms_round<<<ng, 256, 0, s>>>(keys, xy, state, next, head, cell, gw, gh, md2, st, key_cap,
                                              (uint32_t)r);
                KR_LAUNCH_CHECK();
            }
            k_nms_compact<<<ng, 256, 0, s>>>(keys, state, acc, st, key_cap);
            KR_LAUNCH_CHECK();
            k_nms_final<<<1, 1, 0, s>>>(st, mc);
            KR_LAUNCH_CHECK();
        }
    } else {
        k_accept_all<<<sgrid, 256, 0, s>>>(ctx->d_keys_a, ctx->d_keys_b, ctx->d_stats, cap);
        KR_LAUNCH_CHECK();
    }
    KR_MARK(ctx, 6, s);
    // order by (value desc, address desc) = descending 64-bit key
    KR_TRY(krl_sort_u64(ctx, ctx->d_keys_b, ctx->d_keys_a, &ctx->d_stats->n_acc, ctx->cand_cap, 1, s));
    k_emit_corners<<<sgrid, 256, 0, s>>>(ctx->d_keys_a, ctx->d_stats, w,
                                        (max_corners > 0) ? (uint32_t)max_corners : 0u,
                                        (uint32_t)capacity, out_xy, d_count, cap);
    KR_LAUNCH_CHECK();
    KR_MARK(ctx, 7, s);
    return KR_OK;
}

int kr_nms_occupancy(int *blocks_per_sm)
{
    KR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_nms, 256, 0));
    return KR_OK;
}

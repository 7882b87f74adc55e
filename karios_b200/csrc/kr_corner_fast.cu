// kr_corner_fast.cu -- K3 in two tiers (blockSize 15, bounded maxCorners).
//
// cv2.goodFeaturesToTrack needs the min-eigenvalue response in OpenCV's exact
// arithmetic (float32 products, float64 box sums) at only a few pixels: the
// masked maximum and the strongest local maxima.  Everywhere else the response
// only has to be known well enough to rule a pixel out.
//
// Tier 1, k_eig_approx (every pixel, integer arithmetic): the unscaled Sobel
//   derivatives Dx, Dy are integers (|D| <= 1020), the 15 x 15 sums A = sum Dx^2,
//   B = sum Dx Dy, C = sum Dy^2 fit int32 and are exact; lam~ = (A + C)/2 -
//   sqrt(((A - C)/2)^2 + B^2) in float32.  OpenCV's value differs from s^2 lam~
//   (s = the float32 Sobel scale) by at most
//       E = K1 sqrt(X) + K2 X + K0,   X = A + C        (derivation: DESIGN.md 4.1)
//   so with L = lam~ - E, U = lam~ + E (integer units):
//     - a pixel whose U is below the L of one of its 8 neighbours cannot be a
//       3 x 3 local maximum; all others are emitted as possible candidates with
//       key (bits(U) << 32 | y*W + x);
//     - a masked-in pixel whose U is below a running lower bound of the masked
//       maximum cannot be the maximum; the others go to a small "max list".
// Tier 2, k_eig_exact (one warp per listed pixel): OpenCV's arithmetic,
//   restated literally -- float32 Sobel products with the build's FMA placement,
//   float64 sums of the 225 products, calcMinEigenVal in float32 -- for the
//   pixel (max list) or its 3 x 3 neighbourhood (candidates).  Run on the max
//   list entries that can still reach the lower bound (-> exact masked maximum,
//   hence the exact quality threshold) and, after the value cut-off has kept the
//   ~2 maxCorners + 4096 strongest possible candidates, on those: a candidate
//   survives when it is an exact 3 x 3 maximum, above the exact threshold and not
//   below the cut-off; survivors carry exact keys into the NMS.
// Completeness: a possible candidate that was cut has U < cut, hence an exact
// value below the cut-off; the survivors are exactly the candidates with exact
// value >= cut-off, the same set the one-tier kernel selects.  When the image
// is too flat for the bound to separate anything (q * Lmax <= K0) or a list
// overflows, select_incomplete is raised and the caller re-runs the exact
// one-tier kernel (kr_set_select_all), as it does for a short pre-selection.
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include "kr_internal.cuh"

namespace {

// One warp per block: every branch of the kernel then depends on blockIdx only, which the
// compiler knows to be warp-uniform -- with several warps per block the strip index came from
// threadIdx.x >> 5 and every shuffle was guarded by a convergence check (BRA.DIV + UMOV, 7 %
// of the issued instructions).  FA_BLOCKS_PER_SM = resident warps per SM: 16 (128 registers
// per thread, no spills) or 20 (96 registers, a few spills); KR_EIG_BPS=4 / 5 selects.  20 was
// the faster one before the running cut; with it 16 is (same box: 0.461 against 0.472 ms alone,
// 1.280 against 1.326 ms per pair with six pairs in flight), so 16 is the default.
constexpr int FA_WARPS = 1, FA_BLOCKS_PER_SM = 16, FA_OUTW = 104, FA_LEFT = 12, FA_CBUF = 256;
constexpr int FA_RING_I4 = 16 * 32;              // uint4 per warp: 16 rows x 32 lanes of (Dx | Dy << 16) x 4
constexpr float FA_K1 = 0.04f;                   // >= 1.5 x 64 d, d = 7000 * 2^-24 (Sobel rounding)
constexpr float FA_K2 = 1.9073486328125e-6f;     // 2^-19 >= 13 * 2^-24 (products, formula, tier-1 float32)
constexpr float FA_K0 = 0.001f;                  // second-order terms (225 d^2 ...)
constexpr float FA_NEG_INF = -3.0e38f;   // 'nothing yet' for values that go through kr_f32_enc
// Running cut (see approx_body): bins of the estimate histogram = float bits >> FA_GSHIFT (5 mantissa
// bits, 3 % wide; est_bin), the estimate aims at FA_SAFETY x the number of candidates the selection needs,
// a row piece is dropped when no pixel can reach the estimate minus FA_CUT_MARGIN (integer units;
// >= 3 E(X) at the largest possible X = 2 * 225 * 1020^2, E = 1760 there), warps feed the histogram
// during their first FA_EST_ROWS rows.
// U is between K0 = 1e-3 and 2 * 225 * 1020^2 < 2^29 (integer units): biased exponents 117 .. 156,
// FA_GGROUPS groups of 32 bins (one exponent each) from FA_GEXP0 on.
constexpr int FA_GSHIFT = 18, FA_GEXP0 = 117, FA_GGROUPS = 40, FA_GBINS = 32 * FA_GGROUPS, FA_EST_ROWS = 8;
__device__ __forceinline__ uint32_t est_bin(uint32_t ubits)
{
    const int b = (int)(ubits >> FA_GSHIFT) - (FA_GEXP0 << 5);
    return (uint32_t)min(max(b, 0), FA_GBINS - 1);
}
constexpr float FA_SAFETY = 2.0f, FA_CUT_MARGIN = 8000.f;

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b_s8x4)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_s8x4), "r"(0));
    return d;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Row filters of one pixel row, 4 columns per lane.
struct RowF { int hd[4], hs[4]; };
// One response row: U (NaN where the pixel cannot be a candidate / the maximum:
// masked out, outside the image, halo lane, X == 0) and the 3-wide maxima of the
// masked lower bounds L' (-inf where masked out).
struct RespRow { float U[4], H[4]; bool live; };     // live: the row was evaluated (warp-uniform)

template <bool V> struct BoolTag { static constexpr bool value = V; };

// Running cut, contribution of one warp (see approx_body): its candidates so far go into the estimate
// histogram, then the row pieces they stand for are counted.  Out of line and free of warp
// collectives on purpose: anything more in here (a shuffle, a lane-dependent branch that survives
// inlining) makes ptxas guard every shuffle of the caller's row loop against divergence.
__device__ __noinline__ void est_contribute(uint32_t *__restrict__ ghist, KrDevStats *st, const uint64_t *cbuf,
                                            int n, int rows, int lane)
{
    // rows first, candidates second; the scanner reads the other way round (histogram, then rows):
    // whatever it sees in the histogram is covered by the rows it counts, so a race can only make the
    // estimated density -- and with it the estimate of the cut-off -- lower, which is the safe side
    if (lane == 0) atomicAdd(&st->fa_rows, (uint32_t)rows);
    __threadfence();
    __syncwarp();
    for (int k = lane; k < n; k += 32)
        atomicAdd(&ghist[est_bin((uint32_t)(cbuf[k] >> 32))], 1u);
    // ... and a second count once they have landed: the scanner starts when enough rows are IN
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(&st->fa_rows_in, (uint32_t)rows);
}

// Running cut, the scanner: one extra warp of the grid (block (0, 0), dispatched first) waits until
// the histogram stands for `trigger_rows` row pieces -- or until every worker block is done, on
// images too small to get there -- and turns it into the estimate of the cut-off: the lower edge of
// the highest bin above which the sample holds need_per_row x rows candidates.  Nobody waits for it.
__device__ __noinline__ void est_scanner(const uint32_t *ghist, KrDevStats *st, uint32_t trigger_rows,
                                         float need_per_row, uint32_t n_workers, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    // wait until the histogram holds the candidates of `trigger_rows` row pieces, then (the workers
    // of a wave contribute in a burst) until the contributions under way have landed too, within
    // reason; give up only when every worker is done and the sample is still too small
    uint32_t rows = 0;
    for (int spins = 0;;) {
        uint32_t in = 0, done = 0;
        if (lane == 0) {
            rows = *((volatile uint32_t *)&st->fa_rows);
            in = *((volatile uint32_t *)&st->fa_rows_in);
            done = *((volatile uint32_t *)&st->fa_done);
        }
        rows = __shfl_sync(FULL, rows, 0);
        in = __shfl_sync(FULL, in, 0);
        done = __shfl_sync(FULL, done, 0);
        if (in >= trigger_rows && (in == rows || ++spins > 12)) break;
        if (done >= n_workers) {
            if (in < trigger_rows) return;                   // small image: no estimate
            break;
        }
        __nanosleep(400);
    }
    uint32_t bits = 0;
    for (int attempt = 0; attempt < 64; attempt++) {
        const float need = need_per_row * (float)rows;
        // bin 32 g + lane in v[g]: 40 independent, coalesced loads
        uint32_t v[FA_GGROUPS];
#pragma unroll
        for (int g = 0; g < FA_GGROUPS; g++) v[g] = __ldcg(ghist + 32 * g + lane);
        // groups from the top: the first one whose suffix count reaches the target
        uint32_t suffix = 0, above = 0, mine = 0;
        int gstar = -1;
#pragma unroll
        for (int g = FA_GGROUPS - 1; g >= 0; g--) {
            const uint32_t tot = __reduce_add_sync(FULL, v[g]);
            if (gstar < 0 && (float)(suffix + tot) >= need) { gstar = g; above = suffix; mine = v[g]; }
            suffix += tot;
        }
        bits = 0;
        uint32_t found = 0;
        if (gstar >= 0) {
            uint32_t incl = mine;                           // this lane's bin and the higher ones of the group
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(FULL, incl, o);
                if (lane + o < 32) incl += t;
            }
            const unsigned ok = __ballot_sync(FULL, (float)(above + incl) >= need);     // lane 0 is always set
            const int top = 31 - __clz(ok);
            found = above + __shfl_sync(FULL, incl, top);
            bits = (uint32_t)(32 * gstar + top + (FA_GEXP0 << 5)) << FA_GSHIFT;
        }
        // rows counted by now cover everything the histogram showed: the estimate stands if the
        // count found still reaches the target at that row count, else look again
        __threadfence();
        uint32_t rows_now = 0;
        if (lane == 0) rows_now = *((volatile uint32_t *)&st->fa_rows);
        rows_now = __shfl_sync(FULL, rows_now, 0);
        if (bits != 0 && (float)found >= need_per_row * (float)rows_now) break;
        // too few candidates in sight for the rows counted: contributions still under way (look
        // again) -- or, with nothing under way, an image without that many candidates (no estimate)
        uint32_t in_now = 0;
        if (lane == 0) in_now = *((volatile uint32_t *)&st->fa_rows_in);
        in_now = __shfl_sync(FULL, in_now, 0);
        if (bits == 0 && in_now == rows_now) break;
        rows = rows_now;
        bits = 0;
        __nanosleep(400);
    }
    if (lane == 0 && bits) atomicExch(&st->cut_est_bits, bits);
}

template <bool BORDER, bool HAS_MASK>
__device__ __forceinline__ void approx_body(
    const uint8_t *__restrict__ img, uint32_t pitch, const uint8_t *__restrict__ mask, uint32_t mpitch,
    int w, int h, uint64_t *__restrict__ cand, uint32_t cand_cap, uint64_t *__restrict__ maxlist,
    uint32_t maxlist_cap, KrDevStats *st, int xs, int ys, int ye, uint4 *ring, uint64_t *cbuf, int lane,
    uint32_t *__restrict__ ghist)
{
    constexpr unsigned FULL = 0xffffffffu;
    const float NEG_INF = __int_as_float(0xff800000), QNAN = __int_as_float(0x7fc00000);
    const int xb = xs - FA_LEFT + 4 * lane;                        // first (virtual) column of the lane
    int tc[4];
    bool crefl[4], col_in[4], col_ok[4], col_int[4];
    const bool out_lane = lane >= 3 && lane <= 28;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int vx = xb + j;
        tc[j] = BORDER ? kr_reflect101(vx, w) : vx;
        crefl[j] = BORDER && (vx < 0 || vx >= w);
        col_in[j] = !BORDER || (vx >= 0 && vx < w);
        col_ok[j] = out_lane && (!BORDER || (vx >= 0 && vx < w));
        col_int[j] = out_lane && (!BORDER || (vx >= 1 && vx <= w - 2));
    }
    const unsigned lt_mask = (1u << lane) - 1u;

    // row r of the image (REFLECT_101 outside unless the caller knows it is inside)
    auto load_row = [&](int r, bool inside) -> uint32_t {
        const int tr = inside ? r : ((r < 0) ? -r : ((r >= h) ? 2 * (h - 1) - r : r));   // h >= 16
        const uint8_t *prow = img + (size_t)((uint32_t)tr * (uint64_t)pitch);
        if (!BORDER) return __ldg(reinterpret_cast<const uint32_t *>(prow + xb));
        return (uint32_t)__ldg(prow + tc[0]) | ((uint32_t)__ldg(prow + tc[1]) << 8) |
               ((uint32_t)__ldg(prow + tc[2]) << 16) | ((uint32_t)__ldg(prow + tc[3]) << 24);
    };
    auto load_mask = [&](int m, bool inside) -> uint32_t {                   // 1 byte per column
        if (!HAS_MASK) return 0x01010101u;
        if (!inside && (m < 0 || m >= h)) return 0u;
        const uint8_t *mrow = mask + (size_t)((uint32_t)m * (uint64_t)mpitch);
        if (!BORDER) return __ldg(reinterpret_cast<const uint32_t *>(mrow + xb));
        uint32_t mk = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int x = xb + j;
            if (x >= 0 && x < w) mk |= (uint32_t)__ldg(mrow + x) << (8 * j);
        }
        return mk;
    };

    RowF fa, fb, fc;                                   // rotating: rows r-2, r-1, r
    RespRow ra, rb, rc;                                // rotating: response rows q-2, q-1, q
#pragma unroll
    for (int j = 0; j < 4; j++) {
        fa.hd[j] = fa.hs[j] = fb.hd[j] = fb.hs[j] = fc.hd[j] = fc.hs[j] = 0;
        ra.U[j] = rb.U[j] = rc.U[j] = QNAN;
        ra.H[j] = rb.H[j] = rc.H[j] = NEG_INF;
    }
    ra.live = rb.live = rc.live = false;
    // ---- running cut ----------------------------------------------------------------------
    // Only the strongest ~2 maxCorners candidates survive the selection that follows, so most of
    // the image cannot matter.  While a warp works through its first rows its candidates feed a
    // global histogram; once enough rows of the whole grid are in (a stratified sample: the
    // segments tile the image), the warp that crosses the mark turns the histogram into an
    // estimate of the cut-off (aimed FA_SAFETY x too low) and publishes it.  From then on a row
    // piece whose exact integer sums show that no pixel can reach the estimate -- lambda~ <=
    // min(A, C) and lambda~ <= X/2 - |B| -- is dropped before the float bound, the local-maximum
    // test and the emission.  Dropped pixels have U below the estimate, so they can neither be
    // selected candidates nor the masked maximum, and they cannot veto a neighbour that matters
    // (their L is below that neighbour's U); k_cutoff checks that the final cut-off is not below
    // the estimate, else the call is re-run exactly (select_incomplete).  Which rows get dropped
    // depends on timing; the corners do not.
    int cut_i = ghist ? -1 : 0;                        // integer drop threshold; 0: none, < 0: not published yet
    float cut_f = NEG_INF;                             // the same for single pixels: U below it is of no use
    int n_dropped = 0;                                 // (both warp-uniform)
    // (the estimate arrives through a warp reduction, not a shuffle: the compiler then knows the
    // value -- and every branch on it -- to be warp-uniform and emits no convergence guards)
    auto take_cut = [&]() {
        uint32_t ce = 0;
        if (lane == 0) ce = *((volatile const uint32_t *)&st->cut_est_bits);
        ce = __reduce_max_sync(FULL, ce);
        if (ce) {
            const float c = __uint_as_float(ce) - FA_CUT_MARGIN;
            cut_i = (c > FA_CUT_MARGIN) ? (int)c : 0;
            if (cut_i > 0) cut_f = c;
        }
    };
    if (ghist) take_cut();
    int cs[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) cs[c][j] = 0;
    float run_l = NEG_INF;                             // running lower bound of the masked maximum
    float pushed_l = NEG_INF;
    float my_umax = NEG_INF;
    int ccount = 0;                                                    // warp-uniform
    int nprod = 0;

    const int r_first = ys - 9, r_last = ye + 8;
    // pixel / mask words are fetched two rows ahead of their use
    uint32_t pk_next = load_row(r_first, false), pk_next2 = load_row(r_first + 1, false);
    uint32_t mk_next = load_mask(r_first - 8, false), mk_next2 = load_mask(r_first - 7, false);

    // One pixel row: F2/F1 = row filters of rows r-2 / r-1 (read), F0 = row r (written);
    // P/Q = response rows q-2 / q-1 (read), N = response row q = r - 8 (written).
    // STEADY: rows r-1, r+2, r-8, r-6 inside the image, r-9 in [max(ys,1), min(ye,h-1)),
    // 15 product rows already summed, r + 2 <= r_last -- nothing to test.
    auto step = [&](auto steady_tag, int r, const RowF &F2, const RowF &F1, RowF &F0, const RespRow &P,
                    const RespRow &Q, RespRow &N) {
        constexpr bool STEADY = decltype(steady_tag)::value;
        const uint32_t pk = pk_next;
        const uint32_t mkq = mk_next;                       // mask of response row r - 8
        pk_next = pk_next2;
        mk_next = mk_next2;
        if (STEADY || r + 2 <= r_last) {
            pk_next2 = load_row(r + 2, STEADY);
            mk_next2 = load_mask(r + 2 - 8, STEADY);
        }
        // ---- row filters of pixel row r: bytes q0..q5 = columns xb-1 .. xb+4 ----
        const uint32_t wl = __shfl_up_sync(FULL, pk, 1), wr = __shfl_down_sync(FULL, pk, 1);
        const uint32_t F = __funnelshift_l(wl, pk, 8);              // q0 q1 q2 q3
        const uint32_t G = __funnelshift_r(pk, wr, 8);              // q2 q3 q4 q5
        F0.hs[0] = dp4a_us(F, 0x00010201u);  F0.hd[0] = dp4a_us(F, 0x000100ffu);
        F0.hs[1] = dp4a_us(pk, 0x00010201u); F0.hd[1] = dp4a_us(pk, 0x000100ffu);
        F0.hs[2] = dp4a_us(G, 0x00010201u);  F0.hd[2] = dp4a_us(G, 0x000100ffu);
        F0.hs[3] = dp4a_us(G, 0x01020100u);  F0.hd[3] = dp4a_us(G, 0x0100ff00u);
        if (BORDER) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (crefl[j]) F0.hd[j] = -F0.hd[j];      // mirrored column: true neighbours swap
        }
        if (!STEADY && r < r_first + 2) return;
        // ---- integer Sobel derivatives of (virtual) row r-1, packed Dx | Dy << 16 ----
        const bool rrefl = !STEADY && ((r - 1) < 0 || (r - 1) >= h);
        int Dx[4], Dy[4];
        uint32_t pkd[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            Dx[j] = F2.hd[j] + 2 * F1.hd[j] + F0.hd[j];
            Dy[j] = F0.hs[j] - F2.hs[j];
            if (rrefl) Dy[j] = -Dy[j];
            pkd[j] = __byte_perm((uint32_t)Dx[j], (uint32_t)Dy[j], 0x5410);
        }
        nprod++;
        // the ring keeps the derivatives of the last 15 product rows; the leaving row's
        // products are re-formed from them
        if (STEADY || nprod > 15) {
            const uint4 old = ring[(r & 15) * 32];          // row r-16 leaves
            const uint32_t o[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int ox = (int)(int16_t)(o[j] & 0xffffu), oy = (int)o[j] >> 16;   // sign-extended halves
                cs[0][j] += Dx[j] * Dx[j] - ox * ox;
                cs[1][j] += Dx[j] * Dy[j] - ox * oy;
                cs[2][j] += Dy[j] * Dy[j] - oy * oy;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                cs[0][j] += Dx[j] * Dx[j]; cs[1][j] += Dx[j] * Dy[j]; cs[2][j] += Dy[j] * Dy[j];
            }
        }
        ring[((r - 1) & 15) * 32] = make_uint4(pkd[0], pkd[1], pkd[2], pkd[3]);   // row r-1 enters
        if (!STEADY && nprod < 15) return;
        // ---- horizontal 15-column sums -> response row q = r - 8 ----------------
        int bx[3][4];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int c0 = cs[c][0], c1 = cs[c][1], c2 = cs[c][2], c3 = cs[c][3];
            const int P2 = c0 + c1, P3 = P2 + c2, qq = P3 + c3;
            const int S2 = c2 + c3, S3 = S2 + c1;
            const int mid = __shfl_up_sync(FULL, qq, 1) + qq + __shfl_down_sync(FULL, qq, 1);
            const int a3 = __shfl_up_sync(FULL, S3, 2), a2 = __shfl_up_sync(FULL, S2, 2),
                      a1 = __shfl_up_sync(FULL, c3, 2);
            const int b1 = __shfl_down_sync(FULL, c0, 2), b2 = __shfl_down_sync(FULL, P2, 2),
                      b3 = __shfl_down_sync(FULL, P3, 2);
            bx[c][0] = mid + a3;
            bx[c][1] = mid + a2 + b1;
            bx[c][2] = mid + a1 + b2;
            bx[c][3] = mid + b3;
        }
        bool live = true;
        if (cut_i > 0) {
            int v = INT_MIN;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int a = bx[0][j], b = bx[1][j], c = bx[2][j];
                v = max(v, min(min(a, c), ((a + c) >> 1) - abs(b)));
            }
            if (!out_lane) v = INT_MIN;                 // incomplete sums on the halo lanes
            live = __any_sync(FULL, v >= cut_i);
        }
        N.live = live;
        if (live) {
            // rows / columns outside the image hold mirrored data: no pixel there
            const bool row_in = STEADY || ((r - 8) >= 0 && (r - 8) < h);
            float L0[6];
    #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int X = bx[0][j] + bx[2][j], T = bx[0][j] - bx[2][j];
                const float fX = (float)X, fT = (float)T, fB = (float)bx[1][j];
                const float rad = sqrt_approx(fmaf(fT * fT, 0.25f, fB * fB));
                const float lam = fmaf(0.5f, fX, -rad);
                // upper bound of sqrt(fX) (<= 6 % above): halve the exponent, keep the mantissa
                const float sq = __int_as_float((__float_as_int(fX) >> 1) + 0x1fc00000);
                const float E = fmaf(FA_K1, sq, fmaf(FA_K2, fX, FA_K0));
                const float u = lam + E;
                const bool mq = row_in && col_in[j] && (!HAS_MASK || ((mkq >> (8 * j)) & 255u) != 0);
                L0[j + 1] = mq ? lam - E : NEG_INF;
                // X == 0: every product of the window vanishes and so does OpenCV's value
                N.U[j] = (mq && col_ok[j] && X != 0) ? u : QNAN;      // X >= 1 => u >= K0 > 0
            }
            // (the outermost lanes have no neighbours for the 15-column sums: not pixels of this warp)
    #pragma unroll
            for (int j = 0; j < 4; j++)
                if (col_ok[j]) run_l = fmaxf(run_l, L0[j + 1]);
            my_umax = fmaxf(my_umax, max3f(fmaxf(N.U[0], N.U[1]), N.U[2], N.U[3]));   // fmaxf drops NaN
            L0[0] = __shfl_up_sync(FULL, L0[4], 1);
            L0[5] = __shfl_down_sync(FULL, L0[1], 1);
    #pragma unroll
            for (int j = 0; j < 4; j++) N.H[j] = max3f(L0[j], L0[j + 1], L0[j + 2]);
        } else {
            n_dropped++;
#pragma unroll
            for (int j = 0; j < 4; j++) { N.U[j] = QNAN; N.H[j] = NEG_INF; }
        }
        // ---- row m = r - 9: restricted 3 x 3 maxima (NaN compares false) ----------
        const int m = r - 9;
        if (!STEADY && (m < ys || m >= ye)) return;         // warp-uniform
        const bool row_ok = STEADY || (m >= 1 && m <= h - 2);
        if (!Q.live) return;                                // dropped row: no candidate in it
        bool rl[4];
#pragma unroll
        for (int j = 0; j < 4; j++) rl[j] = Q.U[j] >= fmaxf(max3f(P.H[j], Q.H[j], N.H[j]), cut_f);
        if constexpr (STEADY && !BORDER) {
            // Steady interior rows: Q.U is NaN on the halo lanes and every row / column is inside the
            // image, so rl[] already marks exactly this warp's possible candidates.  The first one of
            // each lane is chosen without branches; further ones in the same lane (bounds of
            // neighbouring pixels overlap) take the per-column ballots below.
            const bool r0 = rl[0], r1 = rl[1], r2 = rl[2], r3 = rl[3];
            const bool any_l = r0 || r1 || r2 || r3;
            const unsigned cb = __ballot_sync(FULL, any_l);
            if (cb == 0) return;
            if (ccount > FA_CBUF - 128) {                   // flush the warp buffer
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
            const uint32_t idx0 = (uint32_t)(m * w + xb);
            const float u = r0 ? Q.U[0] : r1 ? Q.U[1] : r2 ? Q.U[2] : Q.U[3];
            const uint32_t jj = r0 ? 0u : r1 ? 1u : r2 ? 2u : 3u;
            if (any_l) cbuf[ccount + __popc(cb & lt_mask)] = ((uint64_t)__float_as_uint(u) << 32) | (idx0 + jj);
            ccount += __popc(cb);
            const bool more[3] = {r1 && r0, r2 && (r0 || r1), r3 && (r0 || r1 || r2)};
            if (__any_sync(FULL, more[0] || more[1] || more[2])) {
#pragma unroll
                for (int j = 1; j < 4; j++) {
                    const unsigned bal = __ballot_sync(FULL, more[j - 1]);
                    if (more[j - 1])
                        cbuf[ccount + __popc(bal & lt_mask)] = ((uint64_t)__float_as_uint(Q.U[j]) << 32) | (idx0 + j);
                    ccount += __popc(bal);
                }
            }
            return;
        }
        const unsigned any_b = __ballot_sync(FULL, rl[0] || rl[1] || rl[2] || rl[3]);
        if (any_b == 0) return;
        unsigned cm = 0, lm = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool cj = rl[j] && row_ok && col_int[j];
            if (cj) cm |= 1u << j;
            // may hold the masked maximum without being a candidate (image border)
            if ((!STEADY || BORDER) && rl[j] && !cj && Q.U[j] >= run_l) lm |= 1u << j;
        }
        if ((!STEADY || BORDER) && __any_sync(FULL, lm != 0)) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool on = (lm >> j) & 1u;
                const unsigned bal = __ballot_sync(FULL, on);
                if (bal) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&st->n_maxlist, (uint32_t)__popc(bal));
                    base = __shfl_sync(FULL, base, 0);
                    if (on) {
                        const uint32_t pos = base + __popc(bal & lt_mask);
                        if (pos < maxlist_cap)
                            maxlist[pos] = ((uint64_t)__float_as_uint(Q.U[j]) << 32) | (uint32_t)(m * w + xb + j);
                        else
                            st->fast_fallback = 1;
                    }
                }
            }
        }
        const unsigned cb = __ballot_sync(FULL, cm != 0);
        if (cb == 0) return;
        if (ccount > FA_CBUF - 128) {                   // flush the warp buffer
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
        const uint32_t idx0 = (uint32_t)(m * w + xb);
        if (!__any_sync(FULL, (cm & (cm - 1)) != 0)) {
            // usual case: at most one possible maximum among a lane's 4 columns
            if (cm) {
                const int j = __ffs(cm) - 1;
                const float u = (j == 0) ? Q.U[0] : (j == 1) ? Q.U[1] : (j == 2) ? Q.U[2] : Q.U[3];
                cbuf[ccount + __popc(cb & lt_mask)] = ((uint64_t)__float_as_uint(u) << 32) | (idx0 + j);
            }
            ccount += __popc(cb);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool is_c = (cm >> j) & 1u;
                const unsigned bal = __ballot_sync(FULL, is_c);
                if (is_c)
                    cbuf[ccount + __popc(bal & lt_mask)] = ((uint64_t)__float_as_uint(Q.U[j]) << 32) | (idx0 + j);
                ccount += __popc(bal);
            }
        }
    };

    // Every few rows: share the lower bound of the masked maximum between warps, pick the estimate
    // of the cut-off up.
    auto periodic = [&](int) {
        float wl_ = run_l;
        for (int o = 16; o > 0; o >>= 1) wl_ = fmaxf(wl_, __shfl_xor_sync(FULL, wl_, o));
        uint32_t genc = 0;
        if (lane == 0) {
            const uint32_t mine = kr_f32_enc(wl_);
            if (wl_ > pushed_l) genc = max(atomicMax(&st->lmax_enc, mine), mine);
            else genc = *((volatile uint32_t *)&st->lmax_enc);
        }
        genc = __shfl_sync(FULL, genc, 0);
        pushed_l = fmaxf(wl_, pushed_l);
        run_l = fmaxf(wl_, kr_f32_dec_bits(genc, 0));
        if (cut_i < 0) take_cut();
    };
    // general rows until the steady range, whole triples of steady rows, general rows after it
    const int s_lo = max(ys + 9, 10), s_hi = min(r_last - 2, h - 3);       // steady: s_lo <= r <= s_hi
    int r = r_first;
    const BoolTag<false> general;
    const BoolTag<true> steady;
    for (; r <= r_last && r < s_lo; r += 3) {
        periodic(r);
        step(general, r, fa, fb, fc, ra, rb, rc);
        if (r + 1 <= r_last) step(general, r + 1, fb, fc, fa, rb, rc, ra);
        if (r + 2 <= r_last) step(general, r + 2, fc, fa, fb, rc, ra, rb);
    }
    // Steady rows, in two stretches of code.  While the estimate of the cut-off is not known
    // (cut_i < 0): rounds of three triples with the periodic step in between, and the warp's one
    // contribution to the estimate histogram after FA_EST_ROWS response rows (cut_i: -1 before it,
    // -2 after).  Then the main loop, with nothing but the rows in it: interior strips have no more
    // use for the periodic step (their lower bound of the maximum is pushed once, at the end), the
    // strips on the left / right border keep it, it prunes their list of border-only maxima.
    const int r_contrib = ys + 9 + FA_EST_ROWS;
    while (cut_i < 0 && r + 2 <= s_hi) {
        if (cut_i == -1 && r >= r_contrib) {                // (entries flushed meanwhile are lost to the
            est_contribute(ghist, st, cbuf, ccount, min(r - 9 - ys, ye - ys), lane);   // estimate: safer)
            cut_i = -2;
        }
        periodic(r);
        const int r_stop = min(r + 9, s_hi - 1);
        for (; r < r_stop; r += 3) {
            step(steady, r, fa, fb, fc, ra, rb, rc);
            step(steady, r + 1, fb, fc, fa, rb, rc, ra);
            step(steady, r + 2, fc, fa, fb, rc, ra, rb);
        }
    }
    if (cut_i == -1) {                                      // short segment: everything it had
        est_contribute(ghist, st, cbuf, ccount, min(max(r - 9 - ys, 0), ye - ys), lane);
        cut_i = -2;
    }
    if (BORDER) {
        while (r + 2 <= s_hi) {
            periodic(r);
            const int r_stop = min(r + 9, s_hi - 1);
            for (; r < r_stop; r += 3) {
                step(steady, r, fa, fb, fc, ra, rb, rc);
                step(steady, r + 1, fb, fc, fa, rb, rc, ra);
                step(steady, r + 2, fc, fa, fb, rc, ra, rb);
            }
        }
    } else {
        for (; r + 2 <= s_hi; r += 3) {
            step(steady, r, fa, fb, fc, ra, rb, rc);
            step(steady, r + 1, fb, fc, fa, rb, rc, ra);
            step(steady, r + 2, fc, fa, fb, rc, ra, rb);
        }
    }
    for (; r <= r_last; r += 3) {
        periodic(r);
        step(general, r, fa, fb, fc, ra, rb, rc);
        if (r + 1 <= r_last) step(general, r + 1, fb, fc, fa, rb, rc, ra);
        if (r + 2 <= r_last) step(general, r + 2, fc, fa, fb, rc, ra, rb);
    }
    __syncwarp();
    if (lane == 0 && n_dropped) atomicAdd(&st->fa_skipped, (uint32_t)n_dropped);
    if (lane == 0 && ghist) atomicAdd(&st->fa_done, 1u);
    if (ccount > 0) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&st->n_cand, (uint32_t)ccount);
        base = __shfl_sync(FULL, base, 0);
        for (int k = lane; k < ccount; k += 32) {
            if (base + k < cand_cap) cand[base + k] = cbuf[k]; else st->overflow = 1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        run_l = fmaxf(run_l, __shfl_xor_sync(FULL, run_l, o));
        my_umax = fmaxf(my_umax, __shfl_xor_sync(FULL, my_umax, o));
    }
    if (lane == 0) {
        if (run_l > pushed_l) atomicMax(&st->lmax_enc, kr_f32_enc(run_l));
        if (my_umax > NEG_INF) atomicMax(&st->umax_enc, kr_f32_enc(my_umax));
    }
}

template <bool HAS_MASK, int BPS>
__global__ void __launch_bounds__(FA_WARPS * 32, BPS)
k_eig_approx(const uint8_t *__restrict__ img, uint32_t pitch, const uint8_t *__restrict__ mask,
             uint32_t mpitch, int w, int h, uint64_t *__restrict__ cand, uint32_t cand_cap,
             uint64_t *__restrict__ maxlist, uint32_t maxlist_cap, KrDevStats *st, int seg, int aligned,
             const unsigned long long *valid_count, uint32_t *__restrict__ ghist, uint32_t est_trigger_rows,
             float est_need_per_row, uint32_t n_bands, uint32_t band_stride)
{
    extern __shared__ __align__(16) unsigned char fa_smem[];
    const int lane = threadIdx.x, wid = 0;                       // FA_WARPS == 1
    uint4 *ring = reinterpret_cast<uint4 *>(fa_smem) + (size_t)wid * FA_RING_I4 + lane;   // [16][32]
    uint64_t *cbuf = reinterpret_cast<uint64_t *>(fa_smem + (size_t)FA_WARPS * FA_RING_I4 * 16) +
                     (size_t)wid * FA_CBUF;
    // with the running cut the first row of the grid holds the scanner (block (0, 0)) and nothing else
    const int by = ghist ? (int)blockIdx.y - 1 : (int)blockIdx.y;
    if (by < 0) {
        if (blockIdx.x == 0)
            est_scanner(ghist, st, est_trigger_rows, est_need_per_row, gridDim.x * (gridDim.y - 1), lane);
        return;
    }
    const int xs = (blockIdx.x * FA_WARPS + wid) * FA_OUTW;
    if (xs >= w) return;
    // bands are handed out with a stride (coprime to their number) instead of top to bottom, so
    // that the blocks dispatched first -- the sample behind the estimate -- spread over the image
    const int band = (int)(((unsigned)by * band_stride) % n_bands);
    const int ys = band * seg, ye = min(ys + seg, h);
    const bool interior = aligned && (xs - FA_LEFT >= 0) && (xs - FA_LEFT + 128 <= w);
    // the context's auto mask with every pixel valid (the usual case) is no mask
    const bool use_mask = HAS_MASK && !(valid_count && *valid_count == (unsigned long long)w * h);
#define FA_BODY(B, M)                                                                                  \
    approx_body<B, M>(img, pitch, mask, mpitch, w, h, cand, cand_cap, maxlist, maxlist_cap, st, xs, ys, \
                      ye, ring, cbuf, lane, ghist)
    if (HAS_MASK && use_mask) {
        if (interior) FA_BODY(false, HAS_MASK); else FA_BODY(true, HAS_MASK);
    } else {
        if (interior) FA_BODY(false, false); else FA_BODY(true, false);
    }
#undef FA_BODY
}

// ---------------------------------------------------------------------------
// Tier 2: OpenCV's arithmetic at single pixels.  One warp per list entry.
// NB = 1: the pixel itself (max list); NB = 3: its 3 x 3 neighbourhood.
constexpr int EX_WARPS = 1;           // one warp per block: warp-uniform control flow (see FA_WARPS)

template <int NB> struct ExSmem {
    double col[3 * NB * (14 + NB)];
    float prod[3 * (14 + NB) * (14 + NB)];
    float pix[(16 + NB) * (16 + NB) + 3];
    int rowoff[16 + NB], coloff[16 + NB];        // REFLECT_101 pixel coordinates of the patch
    int rsy[14 + NB], csx[14 + NB];              // neighbour strides (sign = mirrored), bit 30 of csx = SIMD tail
};

__device__ __forceinline__ double shfl_f64(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}

template <int NB>
__device__ __forceinline__ void exact_values(const uint8_t *__restrict__ img, int64_t pitch, int w, int h,
                                             float s, int tail_start, int x, int y, ExSmem<NB> &sm, int lane,
                                             float &v_centre, float &v_nbmax)
{
    constexpr int R = 14 + NB, PW = R + 2;                      // product / pixel region side
    const int oy = y - 7 - NB / 2, ox = x - 7 - NB / 2;         // product region origin (virtual)
    float *prod = sm.prod;
    double *col = sm.col;                                        // [3][NB][R]
    float *pix = sm.pix;                                         // virtual patch, REFLECT_101 pixels
    if (lane < PW) {
        sm.rowoff[lane] = kr_reflect101(oy - 1 + lane, h);
        sm.coloff[lane] = kr_reflect101(ox - 1 + lane, w);
    }
    if (lane < R) {
        // box-filter border: the product AT the REFLECT_101 position, i.e. with the
        // neighbour roles of the mirrored patch swapped back
        const int gy = oy + lane, gx = ox + lane;
        sm.rsy[lane] = (gy < 0 || gy >= h) ? -PW : PW;
        sm.csx[lane] = ((gx < 0 || gx >= w) ? -1 : 1) * ((kr_reflect101(gx, w) >= tail_start) ? 2 : 1);
    }
    __syncwarp();
    for (int t = lane; t < PW * PW; t += 32) {
        const int py = t / PW, px = t - py * PW;
        pix[t] = (float)__ldg(img + (int64_t)sm.rowoff[py] * pitch + sm.coloff[px]);
    }
    __syncwarp();
    for (int t = lane; t < R * R; t += 32) {
        const int ty = t / R, tx = t - ty * R;
        const int sy = sm.rsy[ty], cx2 = sm.csx[tx];
        const bool tail = cx2 == 2 || cx2 == -2;
        const int sx = cx2 > 0 ? 1 : -1;
        const float *c = pix + (ty + 1) * PW + (tx + 1);
        float xx, xy, yy;
        sobel_products(c[-sy - sx], c[-sy], c[-sy + sx], c[-sx], c[sx], c[sy - sx], c[sy], c[sy + sx], s,
                       tail, xx, xy, yy);
        prod[t] = xx; prod[R * R + t] = xy; prod[2 * R * R + t] = yy;
    }
    __syncwarp();
    // column sums of 15 rows for every (channel, vertical offset, column), then 15
    // columns per window; float64 sums of float32 products are exact
    for (int t = lane; t < 3 * NB * R; t += 32) {
        const int ch = t / (NB * R), rem = t - ch * NB * R, dy = rem / R, cx = rem - dy * R;
        const float *p = prod + ch * R * R + dy * R + cx;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 15; k++) acc += (double)p[k * R];
        col[t] = acc;
    }
    __syncwarp();
    double sum = 0.0;                                            // lane = channel * NB^2 + window
    if (lane < 3 * NB * NB) {
        const int ch = lane / (NB * NB), wnd = lane - ch * NB * NB, dy = wnd / NB, dx = wnd - dy * NB;
        const double *c = col + (ch * NB + dy) * R + dx;
#pragma unroll
        for (int k = 0; k < 15; k++) sum += c[k];
    }
    const double sxy = shfl_f64(sum, lane + NB * NB), syy = shfl_f64(sum, lane + 2 * NB * NB);
    float v = FA_NEG_INF;
    if (lane < NB * NB) v = eig_from_sums(sum, sxy, syy);
    v_centre = __shfl_sync(0xffffffffu, v, (NB * NB) / 2);
    float mx = v;
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    v_nbmax = mx;
    __syncwarp();
}

// Exact masked maximum: every list entry (candidates and border-only pixels) whose
// upper bound reaches the lower bound of the maximum is evaluated exactly.
template <int DUMMY>
__device__ __forceinline__ uint32_t exact_max_of_list(const uint8_t *__restrict__ img, int64_t pitch, int w,
                                                      int h, float s, int tail_start,
                                                      const uint64_t *__restrict__ list, uint32_t n,
                                                      float lmax, ExSmem<1> &sm, int lane, int warp,
                                                      int warps)
{
    uint32_t best = KR_ENC_NEG_INF;
    for (uint32_t base = (uint32_t)warp * 32u; base < n; base += (uint32_t)warps * 32u) {
        const uint32_t i = base + lane;
        const uint64_t e = (i < n) ? list[i] : 0ull;
        const bool hit = (i < n) && __uint_as_float((uint32_t)(e >> 32)) >= lmax;
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const uint32_t idx = __shfl_sync(0xffffffffu, (uint32_t)e, src);
            const int y = (int)(idx / (uint32_t)w), x = (int)(idx - (uint32_t)y * (uint32_t)w);
            float vc, vm;
            exact_values<1>(img, pitch, w, h, s, tail_start, x, y, sm, lane, vc, vm);
            best = max(best, kr_f32_enc(vc));
        }
    }
    return best;
}

__global__ void __launch_bounds__(EX_WARPS * 32)
k_exact_max(const uint8_t *__restrict__ img, int64_t pitch, int w, int h, float s, int tail_start,
            const uint64_t *__restrict__ cand, uint32_t cand_cap, const uint64_t *__restrict__ maxlist,
            uint32_t maxlist_cap, KrDevStats *st)
{
    __shared__ ExSmem<1> sm[EX_WARPS];
    const int lane = threadIdx.x, wib = 0;                     // EX_WARPS == 1
    const int warp = blockIdx.x, warps = gridDim.x;
    const uint32_t lenc = st->lmax_enc;
    if (lenc == KR_ENC_NEG_INF) return;                       // no live pixel: fallback decides
    const float lmax = kr_f32_dec_bits(lenc, 0);
    uint32_t best = exact_max_of_list<0>(img, pitch, w, h, s, tail_start, cand, min(st->n_cand, cand_cap),
                                         lmax, sm[wib], lane, warp, warps);
    best = max(best, exact_max_of_list<0>(img, pitch, w, h, s, tail_start, maxlist,
                                          min(st->n_maxlist, maxlist_cap), lmax, sm[wib], lane, warp,
                                          warps));
    if (lane == 0 && best != KR_ENC_NEG_INF) atomicMax(&st->eig_max_enc, best);
}

// Histogram of the possible candidates' U (integer units) above the threshold
// expressed in the same units; also decides whether the bound is usable at all.
__global__ void __launch_bounds__(256)
k_cand_hist_fast(const uint64_t *__restrict__ cand, KrDevStats *st, uint32_t *__restrict__ hist,
                 double quality, float s, uint32_t cap)
{
    __shared__ uint32_t sh[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
    const uint32_t enc = st->eig_max_enc;
    const float maxv = (enc == KR_ENC_NEG_INF) ? 0.f : kr_f32_dec_bits(enc, 0);
    const float thr = (float)((double)maxv * quality);                 // exact threshold, real units
    const double s2 = (double)s * (double)s;
    // U < thr_u  =>  s^2 U < thr  =>  the exact value is below the threshold
    float thr_u = (thr > 0.f) ? __double2float_rd((double)thr / s2 * (1.0 - 1e-6)) : 0.f;
    const uint32_t thr_bits = (thr_u > 0.f) ? __float_as_uint(thr_u) : 0u;
    const uint32_t uenc = st->umax_enc;
    const float umax = (uenc == KR_ENC_NEG_INF) ? 0.f : kr_f32_dec_bits(uenc, 0);
    const uint32_t max_bits = (umax > 0.f) ? __float_as_uint(umax) : 0u;
    const uint32_t span = (max_bits > thr_bits) ? (max_bits - thr_bits) : 0u;
    const int bits = 32 - __clz(span);
    const uint32_t shift = (bits > 12) ? (uint32_t)(bits - 12) : 0u;
    const uint32_t n = min(st->n_cand, cap);
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t v = (uint32_t)(cand[i] >> 32);
        if (v > thr_bits) atomicAdd(&sh[min((v - thr_bits) >> shift, 4095u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->thr_bits = thr_bits;
        st->hist_shift = shift;
        // pixels with X == 0 were dropped in tier 1: their exact value is <= K0 (integer
        // units), which is below the threshold only if quality * Lmax > K0
        const uint32_t lenc = st->lmax_enc;
        const float lmax = (lenc == KR_ENC_NEG_INF) ? 0.f : kr_f32_dec_bits(lenc, 0);
        if (!(lmax > 2.f * FA_K0) || !((double)lmax * quality > 2.0 * (double)FA_K0) || !(maxv > 0.f))
            st->fast_fallback = 1;
    }
}

// Exact verdict on the selected possible candidates -> exact keys for the NMS.
__global__ void __launch_bounds__(EX_WARPS * 32)
k_exact_cands(const uint8_t *__restrict__ img, int64_t pitch, int w, int h, float s, int tail_start,
              double quality, const uint64_t *__restrict__ sel, uint64_t *__restrict__ keys,
              KrDevStats *st, uint32_t key_cap)
{
    __shared__ ExSmem<3> sm[EX_WARPS];
    const int lane = threadIdx.x, wib = 0;                     // EX_WARPS == 1
    const uint32_t n = min(st->n_sel, key_cap);
    const int warps = gridDim.x;
    const uint32_t enc = st->eig_max_enc;
    const float maxv = (enc == KR_ENC_NEG_INF) ? 0.f : kr_f32_dec_bits(enc, 0);
    float thr = (float)((double)maxv * quality);
    if (!(thr > 0.f)) thr = 0.f;
    // smallest float32 not below s^2 * cut (cut in integer units): every possible
    // candidate that was NOT selected has an exact value below it
    float cut_real = 0.f;
    if (st->cut_applied)
        cut_real = __double2float_ru((double)__uint_as_float(st->cut_bits) * (double)s * (double)s *
                                     (1.0 + 1e-9));
    for (uint32_t i = blockIdx.x; i < n; i += warps) {
        const uint32_t idx = (uint32_t)sel[i];
        const int y = (int)(idx / (uint32_t)w), x = (int)(idx - (uint32_t)y * (uint32_t)w);
        float vc, vm;
        exact_values<3>(img, pitch, w, h, s, tail_start, x, y, sm[wib], lane, vc, vm);
        if (lane == 0 && vc > thr && vc >= vm && vc >= cut_real) {
            const uint32_t pos = atomicAdd(&st->n_exact, 1u);
            keys[pos] = ((uint64_t)__float_as_uint(vc) << 32) | idx;
        }
    }
}

__global__ void k_commit_exact(KrDevStats *st)
{
    st->n_sel = st->n_exact;
    if (st->fast_fallback) st->select_incomplete = 1;
}

}  // namespace

// Tier 1 + exact maximum.  Leaves the possible candidates in ctx->d_cand (keys in
// integer units) and the exact masked maximum in d_stats->eig_max_enc.
int krl_eig_fast(kr_ctx *ctx, const uint8_t *img, int64_t pitch, const uint8_t *mask, int64_t mask_pitch,
                 int w, int h, float scale, int tail_start, uint32_t target, cudaStream_t s)
{
    // resident one-warp blocks per SM: 16 (128 registers per thread) or 20 (96), KR_EIG_BPS=4 / 5;
    // KR_EIG_SMEM_PAD adds unused shared memory per block (caps the residency, for tuning)
    struct Cfg { int bps, occ, seg; size_t smem; cudaError_t err; };
    static const Cfg cfg = [] {
        Cfg c;
        const char *e = getenv("KR_EIG_BPS");
        c.bps = (e && atoi(e) == 5) ? 20 : ((e && atoi(e) == 4) ? 16 : FA_BLOCKS_PER_SM);
        const char *p = getenv("KR_EIG_SMEM_PAD");
        c.smem = (size_t)FA_WARPS * (FA_RING_I4 * 16 + FA_CBUF * 8) + (p ? (size_t)atoi(p) : 0);
        c.err = cudaFuncSetAttribute(k_eig_approx<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (c.err == cudaSuccess)
            c.err = cudaFuncSetAttribute(k_eig_approx<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (c.err == cudaSuccess)
            c.err = cudaFuncSetAttribute(k_eig_approx<true, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (c.err == cudaSuccess)
            c.err = cudaFuncSetAttribute(k_eig_approx<false, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        c.occ = c.bps;
        if (c.err == cudaSuccess) {
            int o = 0;
            c.err = (c.bps == 20)
                ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_eig_approx<false, 20>, FA_WARPS * 32, c.smem)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_eig_approx<false, 16>, FA_WARPS * 32, c.smem);
            if (o >= 1) c.occ = o;
        }
        const char *sg = getenv("KR_EIG_SEG");
        c.seg = sg ? atoi(sg) : 0;
        return c;
    }();
    KR_CUDA(cfg.err);
    const int bps = cfg.bps;
    const size_t smem = cfg.smem;
    int aligned = ((uintptr_t)img % 4 == 0) && (pitch % 4 == 0);
    if (mask) aligned = aligned && ((uintptr_t)mask % 4 == 0) && (mask_pitch % 4 == 0);
    // rows per warp: whole waves of co-resident blocks; each segment pays 18 warm-up rows
    const int sb = (w + FA_WARPS * FA_OUTW - 1) / (FA_WARPS * FA_OUTW);
    int best_seg = h, best_cost = INT_MAX;
    const int slots = ctx->num_sms * cfg.occ;
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
    const int seg = cfg.seg >= 32 ? cfg.seg : best_seg;
    // the context's own auto mask: all-valid is known on the device (K1's count)
    const unsigned long long *valid = (mask && mask == ctx->d_mask) ? &ctx->d_stats->valid : nullptr;
    // running cut (approx_body): the estimate is made once 1 % of the image's row pieces (at least
    // 512 of them: small images never get there and run without a cut) are in the histogram; it
    // aims at FA_SAFETY x target candidates, i.e. est_need_per_row per row piece of the sample
    static const bool no_cut = getenv("KR_EIG_NOCUT") != nullptr;
    const double total_rows = (double)sb * (double)h;
    uint32_t *ghist = no_cut ? nullptr : ctx->d_ghist;
    uint32_t trigger = (uint32_t)(total_rows / 100.0);
    if (trigger < 512u) trigger = 512u;
    const float need_per_row = (float)((double)FA_SAFETY * (double)target / total_rows);
    if (ghist) KR_CUDA(cudaMemsetAsync(ghist, 0, FA_GBINS * sizeof(uint32_t), s));
    const uint32_t n_bands = (uint32_t)((h + seg - 1) / seg);
    uint32_t band_stride = 1;
    if (ghist && n_bands > 4) {                     // ~0.38 n_bands, made coprime to n_bands
        auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
        band_stride = (uint32_t)(0.381966 * n_bands) | 1u;
        while (gcd(band_stride, n_bands) != 1) band_stride += 2;
    }
    dim3 grid(sb, n_bands + (ghist ? 1 : 0));
#define FA_LAUNCH(M, B)                                                                                    \
    k_eig_approx<M, B><<<grid, FA_WARPS * 32, smem, s>>>(img, (uint32_t)pitch, mask, (uint32_t)mask_pitch, w, \
                                                        h, ctx->d_cand, (uint32_t)ctx->cand_cap,             \
                                                        ctx->d_maxlist, (uint32_t)ctx->maxlist_cap,          \
                                                        ctx->d_stats, seg, aligned, valid, ghist, trigger,     \
                                                        need_per_row, n_bands, band_stride)
    if (mask) { if (bps == 20) FA_LAUNCH(true, 20); else FA_LAUNCH(true, 16); }
    else { if (bps == 20) FA_LAUNCH(false, 20); else FA_LAUNCH(false, 16); }
#undef FA_LAUNCH
    KR_LAUNCH_CHECK();
    k_exact_max<<<ctx->num_sms * 32, EX_WARPS * 32, 0, s>>>(img, pitch, w, h, scale, tail_start, ctx->d_cand,
                                                           (uint32_t)ctx->cand_cap, ctx->d_maxlist,
                                                           (uint32_t)ctx->maxlist_cap, ctx->d_stats);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

int krl_cand_hist_fast(kr_ctx *ctx, double quality, float scale, cudaStream_t s)
{
    k_cand_hist_fast<<<ctx->num_sms * 4, 256, 0, s>>>(ctx->d_cand, ctx->d_stats, ctx->d_hist, quality, scale,
                                                      (uint32_t)ctx->cand_cap);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

// selected possible candidates (sel, st->n_sel of them) -> exact keys, st->n_sel updated
int krl_exact_cands(kr_ctx *ctx, const uint8_t *img, int64_t pitch, int w, int h, float scale,
                    int tail_start, double quality, int expected, const uint64_t *sel, uint64_t *keys,
                    cudaStream_t s)
{
    // one warp per selected candidate when the count is as expected (the warps stride otherwise)
    int grid_blocks = (expected + EX_WARPS - 1) / EX_WARPS;
    if (grid_blocks < ctx->num_sms * 8) grid_blocks = ctx->num_sms * 8;
    if (grid_blocks > ctx->num_sms * 128) grid_blocks = ctx->num_sms * 128;
    k_exact_cands<<<grid_blocks, EX_WARPS * 32, 0, s>>>(img, pitch, w, h, scale, tail_start, quality, sel,
                                                       keys, ctx->d_stats, (uint32_t)ctx->cand_cap);
    KR_LAUNCH_CHECK();
    k_commit_exact<<<1, 1, 0, s>>>(ctx->d_stats);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

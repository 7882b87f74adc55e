// kr_api.cu -- context, error reporting and the C ABI of include/karios_b200.h.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <atomic>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include "kr_internal.cuh"

static std::atomic<unsigned long long> g_launches{0};
void kr_note_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";

int kr_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

__global__ void k_set_u32(uint32_t *p, uint32_t v) { *p = v; }

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---- automatic kernel-size search (kr_auto_ksize) ---------------------------------
struct AutoKs { int32_t k[KR_AUTO_MAX_K]; };

__global__ void k_auto_begin(kr_auto_result *r, int n_k)
{
    r->best_mon = r->best_ref = 0;
    r->n_init = r->n_kept = 0;
    r->redo = 0;
    r->n_k = n_k;
    r->valid = 0;
    for (int i = 0; i < KR_AUTO_MAX_K * KR_AUTO_MAX_K; i++) r->counts[i][0] = r->counts[i][1] = 0;
}
__global__ void k_auto_valid(kr_auto_result *r, const KrDevStats *st) { r->valid = st->valid; }
__global__ void k_auto_note_corners(kr_auto_result *r, const KrDevStats *st)
{
    if (st->select_incomplete || st->overflow) r->redo = 1;
}
__global__ void k_auto_record(kr_auto_result *r, int combo, const int32_t *n_init, const KrDevStats *st)
{
    r->counts[combo][0] = *n_init;
    r->counts[combo][1] = (int32_t)st->n_kept;
}
// klt.py:529-539: ratio = len(points) / ninit (Python float division = IEEE double), a pair
// whose reference plane gave no corners has no result, strict > keeps the first maximum
__global__ void k_auto_pick(kr_auto_result *r, AutoKs ks, KrDevStats *st)
{
    const int n_k = r->n_k;
    int best = -1;
    double best_ratio = -1.0;
    for (int c = 0; c < n_k * n_k; c++) {
        const int ni = r->counts[c][0], nk = r->counts[c][1];
        if (ni <= 0) continue;
        const double ratio = (double)nk / (double)ni;
        if (ratio > best_ratio) { best_ratio = ratio; best = c; }
    }
    if (best >= 0) {
        r->best_mon = ks.k[best / n_k];
        r->best_ref = ks.k[best % n_k];
        r->n_init = r->counts[best][0];
        r->n_kept = r->counts[best][1];
    }
    st->n_corners = (uint32_t)r->n_init;
    st->n_kept = (uint32_t)r->n_kept;
}
__global__ void k_auto_copy(const kr_auto_result *r, AutoKs ks, const float *all_rows, int cap, kr_rows out)
{
    const int n_k = r->n_k;
    if (r->best_mon == 0) return;
    int im = 0, ir = 0;
    for (int i = 0; i < n_k; i++) {
        if (ks.k[i] == r->best_mon) im = i;
        if (ks.k[i] == r->best_ref) ir = i;
    }
    const float *src = all_rows + (size_t)(im * n_k + ir) * 5 * cap;
    int n = r->n_kept;
    if (n > out.capacity) n = out.capacity;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        out.x0[i] = src[i];
        out.y0[i] = src[cap + i];
        out.dx[i] = src[2 * cap + i];
        out.dy[i] = src[3 * cap + i];
        out.score[i] = src[4 * cap + i];
    }
}


static_assert(sizeof(kr_unit_header) == 128, "kr_unit_header is 128 bytes (karios_b200/sharding.py)");
// count + dx / dy moments of the rows of the unit just matched (one block of 1024 threads:
// float32 min / max are exact for float32 columns, the sums run in float64)
__global__ void __launch_bounds__(1024) k_unit_header(const KrDevStats *st, kr_rows rows, kr_unit_header *out)
{
    __shared__ double ssum[32][4];
    __shared__ float sext[32][4];
    uint32_t n = st->n_kept;
    if (n > (uint32_t)rows.capacity) n = rows.capacity;
    double sx = 0, sy = 0, sxx = 0, syy = 0;
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float fx = rows.dx[i], fy = rows.dy[i];
        const double dx = fx, dy = fy;
        sx += dx; sy += dy; sxx = fma(dx, dx, sxx); syy = fma(dy, dy, syy);
        mnx = fminf(mnx, fx); mny = fminf(mny, fy); mxx = fmaxf(mxx, fx); mxy = fmaxf(mxy, fy);
    }
    for (int o = 16; o; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sxx += __shfl_xor_sync(0xffffffffu, sxx, o); syy += __shfl_xor_sync(0xffffffffu, syy, o);
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        ssum[warp][0] = sx; ssum[warp][1] = sy; ssum[warp][2] = sxx; ssum[warp][3] = syy;
        sext[warp][0] = mnx; sext[warp][1] = mny; sext[warp][2] = mxx; sext[warp][3] = mxy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; w++) {
            for (int k = 0; k < 4; k++) ssum[0][k] += ssum[w][k];
            sext[0][0] = fminf(sext[0][0], sext[w][0]); sext[0][1] = fminf(sext[0][1], sext[w][1]);
            sext[0][2] = fmaxf(sext[0][2], sext[w][2]); sext[0][3] = fmaxf(sext[0][3], sext[w][3]);
        }
        out->n_rows = (int32_t)n;
        out->flags = (st->select_incomplete ? 1 : 0) | (st->overflow ? 2 : 0);
        out->n = (double)n;
        out->sum_dx = ssum[0][0]; out->sum_dy = ssum[0][1]; out->sum_dx2 = ssum[0][2]; out->sum_dy2 = ssum[0][3];
        out->min_dx = sext[0][0]; out->min_dy = sext[0][1]; out->max_dx = sext[0][2]; out->max_dy = sext[0][3];
        for (int k = 0; k < 6; k++) out->reserved[k] = 0.0;
    }
}

template <typename T> int dev_alloc(T **p, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return kr_set_error(KR_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T),
                            cudaGetErrorString(e));
    }
    *p = (T *)q;
    return KR_OK;
}

int elem_size(int dtype)
{
    switch (dtype) {
    case KR_U8: return 1;
    case KR_U16: case KR_I16: return 2;
    case KR_F32: return 4;
    default: return 0;
    }
}

int check_conf(const kr_klt_conf *c)
{
    if (!c) return kr_set_error(KR_ERR_INVALID, "conf is NULL");
    if (c->win_size < 3 || c->win_size > 29)
        return kr_set_error(KR_ERR_UNSUPPORTED, "matching_winsize %d not supported (3..29)", c->win_size);
    if (c->block_size < 1 || c->block_size > 31)
        return kr_set_error(KR_ERR_UNSUPPORTED, "blocksize %d not supported (1..31)", c->block_size);
    return KR_OK;
}

int check_dims(kr_ctx *ctx, int w, int h)
{
    if (!ctx) return kr_set_error(KR_ERR_INVALID, "ctx is NULL");
    if (w < 1 || h < 1) return kr_set_error(KR_ERR_INVALID, "empty image %dx%d", w, h);
    if (w > ctx->max_w || h > ctx->max_h)
        return kr_set_error(KR_ERR_CAPACITY, "image %dx%d larger than the context (%dx%d)", w, h,
                            ctx->max_w, ctx->max_h);
    return KR_OK;
}

// corners (unless given) -> pyramids -> LK round trip -> rows
int track_common(kr_ctx *ctx, const uint8_t *ref, int64_t ref_pitch, const uint8_t *mon,
                 int64_t mon_pitch, const uint8_t *mask, int64_t mask_pitch, int w, int h,
                 const kr_klt_conf *c, const float *p0, int n_p0, int sort_xy, float x_off,
                 float y_off, kr_rows rows, cudaStream_t s)
{
    int n_cap;
    if (p0) {
        if (n_p0 < 0 || n_p0 > ctx->corner_cap)
            return kr_set_error(KR_ERR_CAPACITY, "%d points exceed the context (%lld)", n_p0,
                                (long long)ctx->corner_cap);
        if (n_p0 > 0)
            KR_CUDA(cudaMemcpyAsync(ctx->d_p0, p0, (size_t)n_p0 * 2 * sizeof(float),
                                    cudaMemcpyDeviceToDevice, s));
        k_set_u32<<<1, 1, 0, s>>>(&ctx->d_stats->n_corners, (uint32_t)n_p0);
        KR_LAUNCH_CHECK();
        n_cap = n_p0;
    } else {
        n_cap = (c->max_corners > 0 && c->max_corners < ctx->corner_cap) ? c->max_corners
                                                                         : (int)ctx->corner_cap;
        KR_TRY(krl_good_features(ctx, ref, ref_pitch, mask, mask_pitch, w, h, c->max_corners,
                                 c->quality_level, c->min_distance, c->block_size, c->tail_mode,
                                 (ctx->force_select_all || c->max_corners <= 0) ? 1 : 0, nullptr, 0,
                                 ctx->d_p0, n_cap, nullptr, s));
    }
    if (rows.capacity < n_cap)
        return kr_set_error(KR_ERR_CAPACITY, "rows.capacity %d < %d possible rows", rows.capacity, n_cap);
    KrLkArgs a;
    memset(&a, 0, sizeof(a));
    KR_TRY(krl_build_pyramids(ctx, ref, ref_pitch, mon, mon_pitch, w, h, c->win_size, c->max_level, &a, s));
    KR_MARK(ctx, 8, s);
    a.max_count = c->max_count;
    a.eps2 = c->eps * c->eps;
    a.min_eig_thr = (float)c->min_eig_threshold;
    const float back_thr = (float)c->back_threshold;
    KR_TRY(krl_lk_roundtrip(a, ctx->d_p0, n_cap, &ctx->d_stats->n_corners, back_thr, ctx->d_p1,
                            ctx->d_d, ctx->d_keep, s));
    KR_MARK(ctx, 9, s);
    KR_TRY(krl_emit_rows(ctx, ctx->d_p0, ctx->d_p1, ctx->d_d, ctx->d_keep, n_cap,
                         &ctx->d_stats->n_corners, sort_xy, back_thr, x_off, y_off, rows, s));
    KR_MARK(ctx, 10, s);
    return KR_OK;
}

}  // namespace

extern "C" {

KR_API int kr_version(void) { return 100; }
KR_API const char *kr_last_error(void) { return g_err; }
KR_API uint64_t kr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

KR_API int kr_ctx_create(int device, int max_w, int max_h, int max_corners, kr_ctx **out)
{
    if (!out) return kr_set_error(KR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (max_w < 1 || max_h < 1 || max_w > 65535 || max_h > 65535)
        return kr_set_error(KR_ERR_INVALID, "context size %dx%d out of range (1..65535)", max_w, max_h);
    KR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    KR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return kr_set_error(KR_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a",
                            device, prop.major, prop.minor);
    kr_ctx *c = new (std::nothrow) kr_ctx();
    if (!c) return kr_set_error(KR_ERR_NOMEM, "host allocation failed");
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->max_w = max_w; c->max_h = max_h; c->max_corners = max_corners;
    const int64_t P = (int64_t)max_w * max_h;
    c->cand_cap = P / 8 > 65536 ? P / 8 : 65536;
    if (c->cand_cap > P) c->cand_cap = P > 1024 ? P : 1024;
    c->corner_cap = (max_corners > 0) ? max_corners : c->cand_cap;
    if (c->corner_cap > c->cand_cap) c->cand_cap = c->corner_cap;
    c->cell_cap = P / 16 + 1024;
    c->plane_pitch = align_up(max_w, 128);
    int rc = KR_OK;
#define A(expr) if (rc == KR_OK) rc = (expr)
    A(dev_alloc(&c->d_stats, 1));
    for (int i = 0; i < 3; i++) A(dev_alloc(&c->d_lut[i], 65536));
    A(dev_alloc(&c->d_cand, c->cand_cap));
    A(dev_alloc(&c->d_keys_a, c->cand_cap));
    A(dev_alloc(&c->d_keys_b, c->cand_cap));
    c->maxlist_cap = c->cand_cap / 4 > 65536 ? c->cand_cap / 4 : 65536;
    A(dev_alloc(&c->d_maxlist, c->maxlist_cap));
    A(dev_alloc(&c->d_hist, 4096));
    A(dev_alloc(&c->d_ghist, 2048));            // >= FA_GBINS of kr_corner_fast.cu
    A(dev_alloc(&c->d_xy, c->cand_cap));
    A(dev_alloc(&c->d_state, c->cand_cap));
    A(dev_alloc(&c->d_next, c->cand_cap));
    A(dev_alloc(&c->d_cell_head, c->cell_cap));
    A(dev_alloc(&c->d_mask, c->plane_pitch * max_h));
    A(dev_alloc(&c->d_lap[0], c->plane_pitch * max_h));
    A(dev_alloc(&c->d_lap[1], c->plane_pitch * max_h));
    for (int l = 1; l < KR_MAX_LEVELS; l++) {
        int64_t lw = ((int64_t)max_w >> l) + 2, lh = ((int64_t)max_h >> l) + 2;
        c->pyr_pitch[l] = align_up(lw, 128);
        A(dev_alloc(&c->d_pyr[0][l], c->pyr_pitch[l] * lh));
        A(dev_alloc(&c->d_pyr[1][l], c->pyr_pitch[l] * lh));
    }
    A(dev_alloc(&c->d_p0, c->corner_cap * 2));
    A(dev_alloc(&c->d_p1, c->corner_cap * 2));
    A(dev_alloc(&c->d_d, c->corner_cap));
    A(dev_alloc(&c->d_keep, c->corner_cap));
#undef A
    if (rc == KR_OK) {
        cudaError_t e = cudaMemset(c->d_hist, 0, 4096 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(c->d_stats, 0, sizeof(KrDevStats));
        if (e != cudaSuccess) rc = kr_set_error(KR_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e));
    }
    if (rc == KR_OK) {
        int bps = 0;
        rc = kr_nms_occupancy(&bps);
        if (rc == KR_OK && bps < 1) rc = kr_set_error(KR_ERR_CUDA, "NMS kernel does not fit on an SM");
        // blocks per SM of the persistent NMS grid (KR_NMS_BPS = 1 | 2; default 2)
        const char *nb = getenv("KR_NMS_BPS");
        const int want = (nb && atoi(nb) == 1) ? 1 : 2;
        c->nms_grid = c->num_sms * (bps > want ? want : bps);
    }
    if (rc == KR_OK) rc = krl_reset_stats(c, 0);
    if (rc == KR_OK) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = kr_set_error(KR_ERR_CUDA, "context init: %s", cudaGetErrorString(e));
    }
    if (rc != KR_OK) {
        kr_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return KR_OK;
}

KR_API void kr_ctx_destroy(kr_ctx *c)
{
    if (!c) return;
    cudaFree(c->d_stats);
    for (int i = 0; i < 3; i++) cudaFree(c->d_lut[i]);
    cudaFree(c->d_cand); cudaFree(c->d_keys_a); cudaFree(c->d_keys_b); cudaFree(c->d_hist); cudaFree(c->d_ghist);
    cudaFree(c->d_maxlist);
    cudaFree(c->d_xy); cudaFree(c->d_state); cudaFree(c->d_next); cudaFree(c->d_cell_head);
    cudaFree(c->d_mask); cudaFree(c->d_lap[0]); cudaFree(c->d_lap[1]);
    for (int l = 1; l < KR_MAX_LEVELS; l++) { cudaFree(c->d_pyr[0][l]); cudaFree(c->d_pyr[1][l]); }
    cudaFree(c->d_p0); cudaFree(c->d_p1); cudaFree(c->d_d); cudaFree(c->d_keep);
    if (c->ev[0]) for (int i = 0; i <= KR_NUM_STAGES; i++) cudaEventDestroy(c->ev[i]);
    delete c;
}

KR_API int kr_unit_header_write(kr_ctx *ctx, kr_rows rows, kr_unit_header *d_out, void *stream)
{
    if (!ctx || !d_out) return kr_set_error(KR_ERR_INVALID, "NULL argument");
    if (!rows.dx || !rows.dy || rows.capacity < 1) return kr_set_error(KR_ERR_INVALID, "incomplete rows");
    k_unit_header<<<1, 1024, 0, (cudaStream_t)stream>>>(ctx->d_stats, rows, d_out);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

// ---- host rasters in ordinary (pageable) memory -> device ---------------------------------
// cudaMemcpy from pageable memory goes through one staging copy on one thread (about 10 GB/s, a
// 241 MB Sentinel-2 band in 20+ ms).  Here KR_UP_THREADS host threads each own a stream and two
// pinned staging buffers and work through the chunks of the raster in turn: memcpy into a free
// staging buffer, cudaMemcpyAsync from it, next chunk -- the staging copies of the threads run in
// parallel and overlap the DMA.  Returns when the host memory has been read completely; the
// caller's stream is ordered after the last chunk.
namespace {
constexpr int KR_UP_THREADS = 6;
constexpr size_t KR_UP_CHUNK = 4u << 20;
struct UpSlot { cudaStream_t s; void *buf[2]; cudaEvent_t ev[2]; cudaEvent_t done; };
constexpr int KR_UP_MAX_DEVICES = 32;
struct UpState { bool ok = false; UpSlot slot[KR_UP_THREADS]; std::mutex mu; };
UpState g_up_dev[KR_UP_MAX_DEVICES];            // one set of streams / staging buffers per device

int up_init(UpState &g_up, int device)
{
    if (g_up.ok) return KR_OK;
    KR_CUDA(cudaSetDevice(device));
    for (int t = 0; t < KR_UP_THREADS; t++) {
        UpSlot &u = g_up.slot[t];
        KR_CUDA(cudaStreamCreateWithFlags(&u.s, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            KR_CUDA(cudaHostAlloc(&u.buf[b], KR_UP_CHUNK, cudaHostAllocDefault));
            KR_CUDA(cudaEventCreateWithFlags(&u.ev[b], cudaEventDisableTiming));
        }
        KR_CUDA(cudaEventCreateWithFlags(&u.done, cudaEventDisableTiming));
    }
    g_up.ok = true;
    return KR_OK;
}
}  // namespace

KR_API int kr_upload_pageable(void *dst_device, const void *src_host, int64_t bytes, int device, void *stream)
{
    if (!dst_device || !src_host || bytes < 0) return kr_set_error(KR_ERR_INVALID, "bad upload arguments");
    if (bytes == 0) return KR_OK;
    if (device < 0 || device >= KR_UP_MAX_DEVICES) return kr_set_error(KR_ERR_INVALID, "bad device %d", device);
    UpState &g_up = g_up_dev[device];
    std::lock_guard<std::mutex> lock(g_up.mu);
    KR_TRY(up_init(g_up, device));
    cudaStream_t s = (cudaStream_t)stream;
    // the destination may still be in use by work queued on the caller's stream
    cudaEvent_t start;
    KR_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    KR_CUDA(cudaEventRecord(start, s));
    const int64_t n_chunks = (bytes + (int64_t)KR_UP_CHUNK - 1) / (int64_t)KR_UP_CHUNK;
    const int n_thr = (int)(n_chunks < KR_UP_THREADS ? n_chunks : KR_UP_THREADS);
    std::atomic<int> failed{0};
    auto work = [&](int t) {
        if (cudaSetDevice(device) != cudaSuccess) { failed = 1; return; }
        UpSlot &u = g_up.slot[t];
        if (cudaStreamWaitEvent(u.s, start, 0) != cudaSuccess) { failed = 1; return; }
        int b = 0;
        for (int64_t c = t; c < n_chunks; c += n_thr, b ^= 1) {
            const int64_t off = c * (int64_t)KR_UP_CHUNK;
            const size_t n = (size_t)((bytes - off) < (int64_t)KR_UP_CHUNK ? (bytes - off) : (int64_t)KR_UP_CHUNK);
            if (cudaEventSynchronize(u.ev[b]) != cudaSuccess) { failed = 1; return; }     // staging buffer free again
            memcpy(u.buf[b], (const char *)src_host + off, n);
            if (cudaMemcpyAsync((char *)dst_device + off, u.buf[b], n, cudaMemcpyHostToDevice, u.s) != cudaSuccess ||
                cudaEventRecord(u.ev[b], u.s) != cudaSuccess) { failed = 1; return; }
        }
        if (cudaEventRecord(u.done, u.s) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_thr; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    cudaEventDestroy(start);
    if (failed) {
        cudaGetLastError();
        return kr_set_error(KR_ERR_CUDA, "kr_upload_pageable: a staging copy failed");
    }
    for (int t = 0; t < n_thr; t++) KR_CUDA(cudaStreamWaitEvent(s, g_up.slot[t].done, 0));
    return KR_OK;
}

KR_API int kr_set_select_all(kr_ctx *ctx, int on)
{
    if (!ctx) return kr_set_error(KR_ERR_INVALID, "ctx is NULL");
    ctx->force_select_all = on ? 1 : 0;
    return KR_OK;
}

KR_API int kr_set_corner_mode(kr_ctx *ctx, int mode)
{
    if (!ctx) return kr_set_error(KR_ERR_INVALID, "ctx is NULL");
    if (mode != 0 && mode != 1) return kr_set_error(KR_ERR_INVALID, "corner mode must be 0 or 1");
    ctx->no_fast_corners = mode;
    return KR_OK;
}

KR_API int kr_set_profiling(kr_ctx *ctx, int on)
{
    if (!ctx) return kr_set_error(KR_ERR_INVALID, "ctx is NULL");
    if (on && !ctx->ev[0])
        for (int i = 0; i <= KR_NUM_STAGES; i++) KR_CUDA(cudaEventCreate(&ctx->ev[i]));
    ctx->prof_on = on ? 1 : 0;
    return KR_OK;
}

KR_API int kr_read_stage_ms(kr_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return kr_set_error(KR_ERR_INVALID, "NULL argument");
    if (!ctx->ev[0]) return kr_set_error(KR_ERR_INVALID, "profiling was never enabled");
    for (int i = 0; i < KR_NUM_STAGES; i++) {
        ms[i] = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]);
        if (e != cudaSuccess) { cudaGetLastError(); ms[i] = -1.f; }
    }
    return KR_OK;
}

KR_API int kr_read_stats(kr_ctx *ctx, void *stream, kr_stats *o)
{
    if (!ctx || !o) return kr_set_error(KR_ERR_INVALID, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    KrDevStats h;
    KR_CUDA(cudaMemcpyAsync(&h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost, s));
    KR_CUDA(cudaStreamSynchronize(s));
    memset(o, 0, sizeof(*o));
    if (ctx->last_dtype == KR_F32) {
        o->min_a = kr_f32_dec_bits(h.minf_enc[0], 0); o->max_a = kr_f32_dec_bits(h.maxf_enc[0], 0);
        o->min_b = kr_f32_dec_bits(h.minf_enc[1], 0); o->max_b = kr_f32_dec_bits(h.maxf_enc[1], 0);
    } else {
        o->min_a = h.min_i[0]; o->max_a = h.max_i[0];
        o->min_b = h.min_i[1]; o->max_b = h.max_i[1];
    }
    o->valid = h.valid;
    o->eig_max = (h.eig_max_enc == KR_ENC_NEG_INF) ? 0.f : kr_f32_dec_bits(h.eig_max_enc, 0);
    o->n_candidates = h.n_cand;
    o->n_above_threshold = h.n_thr;
    o->n_sorted = h.n_sel;
    o->n_corners = h.n_corners;
    o->n_kept = h.n_kept;
    o->nms_rounds = h.nms_rounds;
    o->overflow = h.overflow;
    o->select_incomplete = h.select_incomplete;
    o->two_tier = h.fast_mode;
    o->two_tier_fallback = h.fast_fallback;
    o->n_border_maxima = h.n_maxlist;
    o->n_exact = h.n_exact;
    o->est_cut_bits = h.cut_est_bits;
    o->rows_skipped = h.fa_skipped;
    return KR_OK;
}

KR_API int kr_minmax_mask(kr_ctx *ctx, const void *img_a, int64_t pitch_a, const void *img_b,
                          int64_t pitch_b, int dtype, int w, int h, int has_nodata_a, double nodata_a,
                          int has_nodata_b, double nodata_b, uint8_t *mask_out, int64_t mask_pitch,
                          void *stream)
{
    KR_TRY(check_dims(ctx, w, h));
    if (!img_a) return kr_set_error(KR_ERR_INVALID, "img_a is NULL");
    if (mask_out && !img_b) return kr_set_error(KR_ERR_INVALID, "the auto mask needs both images");
    cudaStream_t s = (cudaStream_t)stream;
    ctx->last_dtype = dtype;
    KR_TRY(krl_reset_stats(ctx, s));
    return krl_minmax_mask(ctx, img_a, pitch_a, img_b, pitch_b, dtype, w, h, has_nodata_a, nodata_a,
                           has_nodata_b, nodata_b, mask_out, mask_pitch, s);
}

KR_API int kr_u8_laplacian(kr_ctx *ctx, const void *img, int64_t pitch, int dtype, int w, int h,
                           int slot, int ksize, int invert, uint8_t *out, int64_t out_pitch,
                           void *stream)
{
    KR_TRY(check_dims(ctx, w, h));
    if (!img || !out) return kr_set_error(KR_ERR_INVALID, "NULL image");
    cudaStream_t s = (cudaStream_t)stream;
    if (slot < 0) {
        slot = 2;
        if (dtype != KR_U8) KR_TRY(krl_minmax_single(ctx, img, pitch, dtype, w, h, slot, s));
    } else if (slot > 1) {
        return kr_set_error(KR_ERR_INVALID, "slot must be -1, 0 or 1");
    }
    return krl_laplacian(ctx, img, pitch, dtype, w, h, slot, ksize, invert, out, out_pitch, s);
}

KR_API int kr_corner_min_eigen_val(kr_ctx *ctx, const uint8_t *img, int64_t pitch, int w, int h,
                                   int block_size, int tail_mode, float *eig, int64_t eig_pitch,
                                   void *stream)
{
    KR_TRY(check_dims(ctx, w, h));
    if (!img || !eig) return kr_set_error(KR_ERR_INVALID, "NULL image");
    return krl_good_features(ctx, img, pitch, nullptr, 0, w, h, 0, 0.0, 0.0, block_size, tail_mode, 0,
                             eig, eig_pitch, nullptr, 0, nullptr, (cudaStream_t)stream);
}

KR_API int kr_good_features(kr_ctx *ctx, const uint8_t *img, int64_t pitch, const uint8_t *mask,
                            int64_t mask_pitch, int w, int h, int max_corners, double quality_level,
                            double min_distance, int block_size, int tail_mode, float *out_xy,
                            int capacity, int32_t *d_count, void *stream)
{
    KR_TRY(check_dims(ctx, w, h));
    if (!img || !out_xy || capacity < 1) return kr_set_error(KR_ERR_INVALID, "NULL image / output");
    int sel_all = ctx->force_select_all || max_corners <= 0;
    return krl_good_features(ctx, img, pitch, mask, mask_pitch, w, h, max_corners, quality_level,
                             min_distance, block_size, tail_mode, sel_all, nullptr, 0, out_xy,
                             capacity, d_count, (cudaStream_t)stream);
}

KR_API int kr_pyr_down(kr_ctx *ctx, const uint8_t *src, int64_t pitch, int w, int h, uint8_t *dst,
                       int64_t dst_pitch, void *stream)
{
    (void)ctx;
    if (!src || !dst || w < 1 || h < 1) return kr_set_error(KR_ERR_INVALID, "bad pyrDown arguments");
    return krl_pyr_down(src, pitch, w, h, dst, dst_pitch, (cudaStream_t)stream);
}

KR_API int kr_pyr_lk(kr_ctx *ctx, const uint8_t *prev, int64_t prev_pitch, const uint8_t *next,
                     int64_t next_pitch, int w, int h, const float *p0, int n, const int32_t *d_count,
                     int win, int max_level, int max_count, double eps, double min_eig_threshold,
                     float *p1, uint8_t *status, float *err, void *stream)
{
    KR_TRY(check_dims(ctx, w, h));
    if (!prev || !next || !p0 || !p1 || !status || !err)
        return kr_set_error(KR_ERR_INVALID, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    KrLkArgs a;
    memset(&a, 0, sizeof(a));
    KR_TRY(krl_build_pyramids(ctx, prev, prev_pitch, next, next_pitch, w, h, win, max_level, &a, s));
    a.max_count = max_count;
    a.eps2 = eps * eps;
    a.min_eig_thr = (float)min_eig_threshold;
    return krl_lk_single(a, p0, n, d_count, p1, status, err, s);
}

KR_API int kr_klt_track(kr_ctx *ctx, const uint8_t *ref, int64_t ref_pitch, const uint8_t *mon,
                        int64_t mon_pitch, const uint8_t *mask, int64_t mask_pitch, int w, int h,
                        const kr_klt_conf *conf, const float *p0, int n_p0, kr_rows rows, void *stream)
{
    KR_TRY(check_dims(ctx, w, h));
    KR_TRY(check_conf(conf));
    if (!ref || !mon) return kr_set_error(KR_ERR_INVALID, "NULL image");
    if (!rows.x0 || !rows.y0 || !rows.dx || !rows.dy || !rows.score || rows.capacity < 1)
        return kr_set_error(KR_ERR_INVALID, "incomplete rows");
    return track_common(ctx, ref, ref_pitch, mon, mon_pitch, mask, mask_pitch, w, h, conf, p0, n_p0, 0,
                        0.f, 0.f, rows, (cudaStream_t)stream);
}

KR_API int kr_zncc(kr_ctx *ctx, const void *ref, int64_t ref_pitch, int ref_w, int ref_h,
                   const void *mon, int64_t mon_pitch, int mon_w, int mon_h, int dtype, const float *x0,
                   const float *y0, const float *dx, const float *dy, int n, const int32_t *d_count,
                   double *out, void *stream)
{
    (void)ctx;
    if (!ref || !mon || !x0 || !y0 || !dx || !dy || !out)
        return kr_set_error(KR_ERR_INVALID, "NULL argument");
    return krl_zncc(ref, ref_pitch, ref_w, ref_h, mon, mon_pitch, mon_w, mon_h, dtype, x0, y0, dx, dy,
                    nullptr, 0.f, n, (const uint32_t *)d_count, out, (cudaStream_t)stream);
}

KR_API int kr_mutual_info(kr_ctx *ctx, const void *ref, int64_t ref_pitch, int ref_w, int ref_h,
                          const void *mon, int64_t mon_pitch, int mon_w, int mon_h, int dtype,
                          const float *x0, const float *y0, const float *dx, const float *dy, int n,
                          const int32_t *d_count, double *out_studholme, double *out_nmi, void *stream)
{
    (void)ctx;
    if (!ref || !mon || !x0 || !y0 || !dx || !dy || (!out_studholme && !out_nmi))
        return kr_set_error(KR_ERR_INVALID, "NULL argument");
    return krl_mutual_info(ref, ref_pitch, ref_w, ref_h, mon, mon_pitch, mon_w, mon_h, dtype, x0, y0,
                           dx, dy, nullptr, 0.f, n, (const uint32_t *)d_count, out_studholme, out_nmi,
                           (cudaStream_t)stream);
}

KR_API int64_t kr_auto_ksize_scratch_bytes(int w, int h, int n_k, int max_corners, int win_size,
                                           int max_level)
{
    if (w < 1 || h < 1 || n_k < 1 || n_k > KR_AUTO_MAX_K || max_corners < 1) return 0;
    int levels, wl[KR_MAX_LEVELS], hl[KR_MAX_LEVELS];
    int64_t pl[KR_MAX_LEVELS];
    krl_pyramid_geometry(w, h, align_up(w, 128), win_size, max_level, &levels, wl, hl, pl);
    int64_t plane = 0;
    for (int l = 0; l <= levels; l++) plane += align_up(pl[l] * hl[l], 256);
    const int64_t cap = max_corners;
    int64_t total = 2ll * n_k * plane;                          // Laplacian planes + pyramids
    total += (int64_t)n_k * align_up(cap * 2 * 4, 256);         // corners per reference kernel size
    total += align_up((int64_t)n_k * 4, 256);                   // their counts
    total += (int64_t)n_k * n_k * align_up(cap * 5 * 4, 256);   // rows per pair
    return total + 1024;
}

KR_API int kr_auto_ksize(kr_ctx *ctx, const void *mon, int64_t mon_pitch, const void *ref,
                         int64_t ref_pitch, int dtype, int w, int h, const uint8_t *mask,
                         int64_t mask_pitch, int has_nodata_mon, double nodata_mon,
                         int has_nodata_ref, double nodata_ref, const kr_klt_conf *conf,
                         const int32_t *ksizes, int n_k, void *scratch, int64_t scratch_bytes,
                         kr_rows rows, kr_auto_result *d_result, void *stream)
{
    KR_TRY(check_dims(ctx, w, h));
    KR_TRY(check_conf(conf));
    if (!elem_size(dtype)) return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    if (!mon || !ref || !ksizes || !scratch || !d_result) return kr_set_error(KR_ERR_INVALID, "NULL argument");
    if (n_k < 1 || n_k > KR_AUTO_MAX_K) return kr_set_error(KR_ERR_INVALID, "n_k %d out of range", n_k);
    if (conf->max_corners < 1 || conf->max_corners > ctx->corner_cap)
        return kr_set_error(KR_ERR_UNSUPPORTED, "kr_auto_ksize needs 1 <= maxCorners <= the context's");
    if (!rows.x0 || !rows.y0 || !rows.dx || !rows.dy || !rows.score || rows.capacity < conf->max_corners)
        return kr_set_error(KR_ERR_INVALID, "incomplete rows");
    const int64_t need = kr_auto_ksize_scratch_bytes(w, h, n_k, conf->max_corners, conf->win_size, conf->max_level);
    if (scratch_bytes < need || (uintptr_t)scratch % 256 != 0)
        return kr_set_error(KR_ERR_CAPACITY, "scratch of %lld bytes (256-aligned) needed", (long long)need);
    cudaStream_t s = (cudaStream_t)stream;
    ctx->last_dtype = dtype;
    const int cap = conf->max_corners;
    AutoKs ks;
    memset(&ks, 0, sizeof(ks));
    for (int i = 0; i < n_k; i++) ks.k[i] = ksizes[i];

    // ---- scratch layout ----------------------------------------------------------
    int levels, wl[KR_MAX_LEVELS], hl[KR_MAX_LEVELS];
    int64_t pl[KR_MAX_LEVELS];
    krl_pyramid_geometry(w, h, align_up(w, 128), conf->win_size, conf->max_level, &levels, wl, hl, pl);
    char *cur = (char *)scratch;
    auto take = [&](int64_t bytes) { char *p = cur; cur += align_up(bytes, 256); return p; };
    uint8_t *pm[KR_AUTO_MAX_K][KR_MAX_LEVELS], *pr[KR_AUTO_MAX_K][KR_MAX_LEVELS];
    for (int i = 0; i < n_k; i++)
        for (int l = 0; l <= levels; l++) pm[i][l] = (uint8_t *)take(pl[l] * hl[l]);
    for (int i = 0; i < n_k; i++)
        for (int l = 0; l <= levels; l++) pr[i][l] = (uint8_t *)take(pl[l] * hl[l]);
    float *p0[KR_AUTO_MAX_K];
    for (int i = 0; i < n_k; i++) p0[i] = (float *)take((int64_t)cap * 2 * 4);
    int32_t *cnt = (int32_t *)take((int64_t)n_k * 4);
    float *all_rows = (float *)cur;
    // (the rows of the pairs are contiguous: k_auto_copy indexes them as [pair][5][cap])

    k_auto_begin<<<1, 1, 0, s>>>(d_result, n_k);
    KR_LAUNCH_CHECK();
    KR_TRY(krl_reset_stats(ctx, s));
    KR_TRY(krl_minmax_mask(ctx, mon, mon_pitch, ref, ref_pitch, dtype, w, h, has_nodata_mon, nodata_mon,
                           has_nodata_ref, nodata_ref, mask ? nullptr : ctx->d_mask, ctx->plane_pitch, s));
    k_auto_valid<<<1, 1, 0, s>>>(d_result, ctx->d_stats);
    KR_LAUNCH_CHECK();
    const uint8_t *m = mask ? mask : ctx->d_mask;
    const int64_t mp = mask ? mask_pitch : ctx->plane_pitch;
    for (int i = 0; i < n_k; i++) {
        KR_TRY(krl_laplacian(ctx, mon, mon_pitch, dtype, w, h, 0, ks.k[i], conf->invert_mon, pm[i][0], pl[0], s));
        KR_TRY(krl_pyramid_plane(pm[i][0], levels, wl, hl, pl, pm[i], s));
        KR_TRY(krl_laplacian(ctx, ref, ref_pitch, dtype, w, h, 1, ks.k[i], 0, pr[i][0], pl[0], s));
        KR_TRY(krl_pyramid_plane(pr[i][0], levels, wl, hl, pl, pr[i], s));
    }
    for (int i = 0; i < n_k; i++) {
        KR_TRY(krl_good_features(ctx, pr[i][0], pl[0], m, mp, w, h, conf->max_corners, conf->quality_level,
                                 conf->min_distance, conf->block_size, conf->tail_mode,
                                 ctx->force_select_all ? 1 : 0, nullptr, 0, p0[i], cap, cnt + i, s));
        k_auto_note_corners<<<1, 1, 0, s>>>(d_result, ctx->d_stats);
        KR_LAUNCH_CHECK();
    }
    const float back_thr = (float)conf->back_threshold;
    for (int im = 0; im < n_k; im++) {
        for (int ir = 0; ir < n_k; ir++) {
            KrLkArgs a;
            memset(&a, 0, sizeof(a));
            for (int l = 0; l <= levels; l++) {
                a.img[0][l] = pr[ir][l]; a.img[1][l] = pm[im][l];
                a.pitch[0][l] = a.pitch[1][l] = pl[l];
                a.w[l] = wl[l]; a.h[l] = hl[l];
            }
            a.levels = levels;
            a.win = conf->win_size;
            a.max_count = conf->max_count;
            a.eps2 = conf->eps * conf->eps;
            a.min_eig_thr = (float)conf->min_eig_threshold;
            const uint32_t *n_init = (const uint32_t *)(cnt + ir);
            KR_TRY(krl_lk_roundtrip(a, p0[ir], cap, n_init, back_thr, ctx->d_p1, ctx->d_d, ctx->d_keep, s));
            const int combo = im * n_k + ir;
            float *base = all_rows + (size_t)combo * 5 * cap;
            kr_rows rc;
            memset(&rc, 0, sizeof(rc));
            rc.x0 = base; rc.y0 = base + cap; rc.dx = base + 2 * (size_t)cap; rc.dy = base + 3 * (size_t)cap;
            rc.score = base + 4 * (size_t)cap;
            rc.capacity = cap;
            KR_TRY(krl_emit_rows(ctx, p0[ir], ctx->d_p1, ctx->d_d, ctx->d_keep, cap, n_init, 0, back_thr, 0.f,
                                 0.f, rc, s));
            k_auto_record<<<1, 1, 0, s>>>(d_result, combo, cnt + ir, ctx->d_stats);
            KR_LAUNCH_CHECK();
        }
    }
    k_auto_pick<<<1, 1, 0, s>>>(d_result, ks, ctx->d_stats);
    KR_LAUNCH_CHECK();
    k_auto_copy<<<(cap + 255) / 256, 256, 0, s>>>(d_result, ks, all_rows, cap, rows);
    KR_LAUNCH_CHECK();
    return KR_OK;
}

KR_API int kr_match_tile(kr_ctx *ctx, const void *mon, int64_t mon_pitch, const void *ref,
                         int64_t ref_pitch, int dtype, int img_w, int img_h, const uint8_t *mask,
                         int64_t mask_pitch, int x_off, int y_off, int tile_w, int tile_h,
                         int has_nodata_mon, double nodata_mon, int has_nodata_ref, double nodata_ref,
                         const kr_klt_conf *conf, kr_rows rows, void *stream)
{
    KR_TRY(check_dims(ctx, tile_w, tile_h));
    KR_TRY(check_conf(conf));
    const int es = elem_size(dtype);
    if (!es) return kr_set_error(KR_ERR_UNSUPPORTED, "unsupported raster dtype %d", dtype);
    if (!mon || !ref) return kr_set_error(KR_ERR_INVALID, "NULL image");
    if (x_off < 0 || y_off < 0 || x_off + tile_w > img_w || y_off + tile_h > img_h)
        return kr_set_error(KR_ERR_INVALID, "tile window outside the raster");
    if (!rows.x0 || !rows.y0 || !rows.dx || !rows.dy || !rows.score || rows.capacity < 1)
        return kr_set_error(KR_ERR_INVALID, "incomplete rows");
    cudaStream_t s = (cudaStream_t)stream;
    ctx->last_dtype = dtype;
    const char *mon_t = (const char *)mon + (int64_t)y_off * mon_pitch + (int64_t)x_off * es;
    const char *ref_t = (const char *)ref + (int64_t)y_off * ref_pitch + (int64_t)x_off * es;
    const uint8_t *mask_t = mask ? mask + (int64_t)y_off * mask_pitch + x_off : nullptr;

    KR_TRY(krl_reset_stats(ctx, s));
    KR_MARK(ctx, 0, s);
    // a = monitored (slot 0), b = reference (slot 1); the auto mask only without a user mask
    KR_TRY(krl_minmax_mask(ctx, mon_t, mon_pitch, ref_t, ref_pitch, dtype, tile_w, tile_h,
                           has_nodata_mon, nodata_mon, has_nodata_ref, nodata_ref,
                           mask ? nullptr : ctx->d_mask, ctx->plane_pitch, s));
    KR_MARK(ctx, 1, s);
    KR_TRY(krl_laplacian(ctx, mon_t, mon_pitch, dtype, tile_w, tile_h, 0, conf->ksize_mon,
                         conf->invert_mon, ctx->d_lap[0], ctx->plane_pitch, s));
    KR_MARK(ctx, 2, s);
    KR_TRY(krl_laplacian(ctx, ref_t, ref_pitch, dtype, tile_w, tile_h, 1, conf->ksize_ref, 0,
                         ctx->d_lap[1], ctx->plane_pitch, s));
    KR_MARK(ctx, 3, s);
    const uint8_t *m = mask ? mask_t : ctx->d_mask;
    const int64_t mp = mask ? mask_pitch : ctx->plane_pitch;
    KR_TRY(track_common(ctx, ctx->d_lap[1], ctx->plane_pitch, ctx->d_lap[0], ctx->plane_pitch, m, mp,
                        tile_w, tile_h, conf, nullptr, 0, 1, (float)x_off, (float)y_off, rows, s));
    if (conf->compute_zncc && rows.zncc)
        KR_TRY(krl_zncc(ref, ref_pitch, img_w, img_h, mon, mon_pitch, img_w, img_h, dtype, rows.x0,
                        rows.y0, rows.dx, rows.dy, rows.score, (float)conf->zncc_min_score,
                        rows.capacity, &ctx->d_stats->n_kept, rows.zncc, s));
    KR_MARK(ctx, 11, s);
    if (conf->compute_mi && (rows.mutual_info || rows.mi))
        KR_TRY(krl_mutual_info(ref, ref_pitch, img_w, img_h, mon, mon_pitch, img_w, img_h, dtype,
                               rows.x0, rows.y0, rows.dx, rows.dy, rows.score,
                               (float)conf->zncc_min_score, rows.capacity, &ctx->d_stats->n_kept,
                               rows.mutual_info, rows.mi, s));
    KR_MARK(ctx, 12, s);
    return KR_OK;
}

}  // extern "C"

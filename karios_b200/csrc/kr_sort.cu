// kr_sort.cu -- sort of unique 64-bit keys whose count lives in device memory.
//
// Used for (a) the accepted corners, descending by (response, address) -- the
// order cv2.goodFeaturesToTrack returns (klt.py:120) -- and (b) the result rows,
// ascending by (x0, y0) -- DataFrame.sort_values(["x0","y0"]) (klt.py:348).
// Typical sizes are <= a few 10^4 keys, so the design is two launches:
//   1. every chunk of KR_SORT_CHUNK keys is sorted in shared memory (bitonic);
//   2. every key finds its final rank = rank in its chunk + number of keys that
//      precede it in each other chunk (binary search), and scatters itself.
// Keys are unique, hence ranks are a permutation and the result deterministic.
#include "kr_internal.cuh"

namespace {

constexpr int CH = KR_SORT_CHUNK;
// chunk size for a list of n keys: small lists (the usual <= 2 maxCorners + 4096 keys) use
// 2048-key chunks -- four times the blocks and a third of the compare-exchange passes of one
// 8192-key block, for a few more binary searches per key in the rank step
__device__ __forceinline__ int sort_chunk(int64_t n) { return n <= 65536 ? 2048 : CH; }

template <bool DESC>
__global__ void __launch_bounds__(1024) k_chunk_sort(uint64_t *keys, const uint32_t *d_n, int64_t cap)
{
    extern __shared__ uint64_t sk[];
    int64_t n = *d_n;
    if (n > cap) n = cap;
    const uint64_t pad = DESC ? 0ull : ~0ull;
    const int chunk = sort_chunk(n);
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < n; base += (int64_t)gridDim.x * chunk) {
        int m = (int)((n - base < chunk) ? (n - base) : chunk);
        int P = 1;
        while (P < m) P <<= 1;
        for (int i = threadIdx.x; i < P; i += blockDim.x) sk[i] = (i < m) ? keys[base + i] : pad;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    int l = i + j;
                    uint64_t a = sk[i], b = sk[l];
                    bool up = ((i & k) == 0);           // this run ends in final order
                    bool sw = DESC ? (up ? (a < b) : (a > b)) : (up ? (a > b) : (a < b));
                    if (sw) { sk[i] = b; sk[l] = a; }
                }
                __syncthreads();
            }
        }
        for (int i = threadIdx.x; i < m; i += blockDim.x) keys[base + i] = sk[i];
        __syncthreads();
    }
}

template <bool DESC>
__global__ void __launch_bounds__(256) k_rank_merge(const uint64_t *__restrict__ keys,
                                                   uint64_t *__restrict__ out, const uint32_t *d_n,
                                                   int64_t cap)
{
    int64_t n = *d_n;
    if (n > cap) n = cap;
    const int chunk = sort_chunk(n);
    int nchunks = (int)((n + chunk - 1) / chunk);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t key = keys[i];
        int c = (int)(i / chunk);
        int64_t rank = i - (int64_t)c * chunk;
        for (int cc = 0; cc < nchunks; cc++) {
            if (cc == c) continue;
            int64_t lo = (int64_t)cc * chunk;
            int m = (int)((n - lo < chunk) ? (n - lo) : chunk);
            // number of keys of chunk cc that come before `key` in the final order
            int a = 0, b = m;
            while (a < b) {
                int mid = (a + b) >> 1;
                uint64_t v = keys[lo + mid];
                bool before = DESC ? (v > key) : (v < key);
                if (before) a = mid + 1; else b = mid;
            }
            rank += a;
        }
        out[rank] = key;
    }
}

}  // namespace

int krl_sort_u64(kr_ctx *ctx, uint64_t *keys, uint64_t *out, const uint32_t *d_n, int64_t cap,
                 int descending, cudaStream_t s)
{
    static bool attr_set = false;
    const int smem = CH * (int)sizeof(uint64_t);
    if (!attr_set) {
        KR_CUDA(cudaFuncSetAttribute(k_chunk_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        KR_CUDA(cudaFuncSetAttribute(k_chunk_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    int64_t nchunks = (cap + 2047) / 2048;
    int grid1 = (int)((nchunks < 2 * ctx->num_sms) ? nchunks : 2 * ctx->num_sms);
    if (grid1 < 1) grid1 = 1;
    int64_t nb = (cap + 255) / 256;
    int grid2 = (int)((nb < 8 * ctx->num_sms) ? nb : 8 * ctx->num_sms);
    if (grid2 < 1) grid2 = 1;
    if (descending) {
        k_chunk_sort<true><<<grid1, 1024, smem, s>>>(keys, d_n, cap);
        KR_LAUNCH_CHECK();
        k_rank_merge<true><<<grid2, 256, 0, s>>>(keys, out, d_n, cap);
    } else {
        k_chunk_sort<false><<<grid1, 1024, smem, s>>>(keys, d_n, cap);
        KR_LAUNCH_CHECK();
        k_rank_merge<false><<<grid2, 256, 0, s>>>(keys, out, d_n, cap);
    }
    KR_LAUNCH_CHECK();
    return KR_OK;
}

// jax.lax.top_k along the last axis (sm_100a): rows of `len` float32 scores -> the k largest, descending, ties -> lower
// index.  Replaces `jax.lax.top_k(action.flatten(), n)` of the action wrappers (breedgym/vector/vec_wrappers.py:101,
// breedgym/vector/breeding_programs_env.py:27) -- the only O(E n^2) piece of their index math: PairScores ranks
// 64 x 370^2 = 8.8 M pair scores per step.
//
// One CTA per row.  Every element gets a unique 64-bit key (order-preserving image of the float32 value in the high
// word, inverted index in the low word), so "descending, ties -> lower index" is plain descending key order.  The k-th
// largest key is found by an MSB-first radix select (8 bits per pass over a 256-bin shared-memory histogram; the four
// passes over the index half run only when the value at the threshold is tied), the k survivors are collected and
// bitonic-sorted in shared memory.  The scores are read 4-9 times, from L2.
#include "bg_internal.h"

namespace {

constexpr int TK_THREADS = 1024;
constexpr int TK_MAX_K = 1024;

__device__ __forceinline__ uint32_t ordered_bits(float x)
{
    const uint32_t b = __float_as_uint(x + 0.0f);  // -0.0 -> +0.0: the two compare equal, so they get one key
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float from_ordered(uint32_t o)
{
    const uint32_t b = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
    return __uint_as_float(b);
}

__global__ void __launch_bounds__(TK_THREADS) topk_rows_kernel(const float *__restrict__ scores, int64_t len, int k,
                                                              float *__restrict__ vals_out, int32_t *__restrict__ idx_out)
{
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long cand[TK_MAX_K];
    __shared__ uint32_t sh_prefix_hi, sh_prefix_lo, sh_need, sh_count;
    const int tid = threadIdx.x;
    const float *row = scores + (int64_t)blockIdx.x * len;
    if (tid == 0) {
        sh_prefix_hi = 0;
        sh_prefix_lo = 0;
        sh_need = (uint32_t)k;
        sh_count = 0;
    }
    __syncthreads();
    // ---- radix select of the k-th largest 64-bit key, most significant byte first
    bool tied = true;  // (uniform) the low half matters only if several elements share the threshold value
    for (int pass = 0; pass < 8 && tied; ++pass) {
        const bool hi = pass < 4;
        const int shift = 24 - 8 * (pass & 3);
        for (int b = tid; b < 256; b += TK_THREADS) hist[b] = 0;
        __syncthreads();
        const uint32_t p_hi = sh_prefix_hi, p_lo = sh_prefix_lo;
        const uint32_t mask_done = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));  // bits already fixed in this half
        for (int64_t i = tid; i < len; i += TK_THREADS) {
            const uint32_t o = ordered_bits(__ldg(row + i));
            if (hi) {
                if ((o & mask_done) == (p_hi & mask_done)) atomicAdd(&hist[(o >> shift) & 255u], 1u);
            } else if (o == p_hi) {
                const uint32_t lo = ~(uint32_t)i;
                if ((lo & mask_done) == (p_lo & mask_done)) atomicAdd(&hist[(lo >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t need = sh_need, d = 255;
            for (;; --d) {  // from the largest digit down: the bucket that holds the need-th largest remaining key
                if (hist[d] >= need || d == 0) break;
                need -= hist[d];
            }
            sh_need = need;
            if (hi) sh_prefix_hi = p_hi | (d << shift);
            else sh_prefix_lo = p_lo | (d << shift);
            sh_count = hist[d];
        }
        __syncthreads();
        if (pass == 3) tied = sh_count != sh_need;  // all elements equal to the threshold value are wanted: done
    }
    const unsigned long long kth = tied ? (((unsigned long long)sh_prefix_hi << 32) | sh_prefix_lo)
                                        : ((unsigned long long)sh_prefix_hi << 32);
    __syncthreads();
    if (tid == 0) sh_count = 0;
    __syncthreads();
    // ---- collect the k keys >= kth
    for (int64_t i = tid; i < len; i += TK_THREADS) {
        const unsigned long long key = ((unsigned long long)ordered_bits(__ldg(row + i)) << 32) | (uint32_t)~(uint32_t)i;
        if (key >= kth) {
            const uint32_t pos = atomicAdd(&sh_count, 1u);
            if (pos < (uint32_t)TK_MAX_K) cand[pos] = key;
        }
    }
    __syncthreads();
    // ---- bitonic sort (descending) of the candidates, padded with zeros (smaller than any real key: index < 2^32 - 1)
    int P = 1;
    while (P < k) P <<= 1;
    for (int i = tid; i < P; i += TK_THREADS)
        if (i >= k) cand[i] = 0ull;
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < P; i += TK_THREADS) {
                const int j = i ^ stride;
                if (j > i) {
                    const unsigned long long a = cand[i], b = cand[j];
                    const bool desc = (i & size) == 0;
                    if (desc ? a < b : a > b) {
                        cand[i] = b;
                        cand[j] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < k; i += TK_THREADS) {
        const unsigned long long key = cand[i];
        vals_out[(int64_t)blockIdx.x * k + i] = from_ordered((uint32_t)(key >> 32));
        idx_out[(int64_t)blockIdx.x * k + i] = (int32_t)~(uint32_t)key;
    }
}

}  // namespace

int bg_launch_topk(const float *scores, int64_t rows, int64_t len, int k, float *vals_out, int32_t *idx_out, cudaStream_t st)
{
    BG_REQUIRE(k >= 1 && k <= TK_MAX_K && (int64_t)k <= len, BG_ELIMIT, "bg_topk: k must be in 1..min(len, 1024)");
    BG_REQUIRE(len < (int64_t(1) << 32) - 1 && rows < (int64_t(1) << 31), BG_ELIMIT, "bg_topk: row too long");
    if (rows == 0) return BG_OK;
    topk_rows_kernel<<<(unsigned)rows, TK_THREADS, 0, st>>>(scores, len, k, vals_out, idx_out);
    BG_LAUNCHED();
    return BG_OK;
}

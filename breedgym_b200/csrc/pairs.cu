// Action-wrapper index math after the top-k (include/breedgym_b200.h, "pair selection"): what PairScores and
// SelectionScores do between `jax.lax.top_k` and the cross.
//
//   pairs_from_topk_kernel  -- PairScores._convert_actions (breedgym/vector/vec_wrappers.py:100-112; WheatBreedGym
//     re-uses it, breeding_programs_env.py:24-36): offspring per pair = ceil(softmax(best values) * k),
//     `jnp.repeat(pairs, counts, total_repeat_length=k)` = output slot s takes the first pair whose running count
//     exceeds s (the last pair pads), pair = (flat index / n, flat index % n).  One CTA per env: max (= first value, the
//     list is sorted), exp, block sum, counts, block prefix sum in shared memory, one binary search per output slot.
//   diallel_pairs_kernel    -- SelectionScores._convert_actions (vec_wrappers.py:60-78): the chosen entries of the
//     upper-triangular pair list of the k best (`Simulator._diallel_indices`), each repeated ceil(n / n_crosses) times,
//     cut / padded to n.  The (row, column) of linear index p comes from the closed form of np.triu_indices(k, 1).
#include "bg_internal.h"

namespace {

constexpr int PT = 256;
constexpr int PK_MAX = 1024;

__global__ void __launch_bounds__(PT) pairs_from_topk_kernel(const float *__restrict__ vals, const int32_t *__restrict__ idx, int k,
                                                             int64_t row_len, int32_t *__restrict__ out)
{
    __shared__ int ends[PK_MAX];
    __shared__ double wsum[PT / 32];
    __shared__ int wtot[PT / 32];
    __shared__ float inv_shared;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *v = vals + (int64_t)blockIdx.x * k;
    const int32_t *ix = idx + (int64_t)blockIdx.x * k;
    const float vmax = v[0];  // descending list
    // softmax denominator (float64 accumulation, rounded once)
    double s = 0.0;
    for (int i = tid; i < k; i += PT) s += (double)expf(v[i] - vmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < PT / 32; ++w) t += wsum[w];
        inv_shared = (float)t;
    }
    __syncthreads();
    const float denom = inv_shared;
    // counts -> inclusive prefix sums: every thread owns a contiguous run of ceil(k / PT) entries
    const int per = (k + PT - 1) / PT, beg = min(k, tid * per), end = min(k, beg + per);
    int run = 0;
    for (int i = beg; i < end; ++i) {
        const float p = expf(v[i] - vmax) / denom;
        run += (int)ceilf(p * (float)k);
        ends[i] = run;
    }
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    int base = incl - run;
    for (int w = 0; w < warp; ++w) base += wtot[w];
    for (int i = beg; i < end; ++i) ends[i] += base;
    __syncthreads();
    // slot s <- first pair b with ends[b] > s (searchsorted right), the last pair when the counts run out
    for (int sl = tid; sl < k; sl += PT) {
        int lo = 0, hi = k;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ends[mid] <= sl) lo = mid + 1;
            else hi = mid;
        }
        const int b = min(lo, k - 1);
        const int64_t flat = ix[b];
        int32_t *o = out + ((int64_t)blockIdx.x * k + sl) * 2;
        o[0] = (int32_t)(flat / row_len);
        o[1] = (int32_t)(flat % row_len);
    }
}

__global__ void __launch_bounds__(PT) diallel_pairs_kernel(const int32_t *__restrict__ best, const int32_t *__restrict__ perm, int k, int nc,
                                                           int64_t n, int rep, int32_t *__restrict__ out)
{
    const int64_t sl = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (sl >= n) return;
    const int64_t e = blockIdx.y;
    int64_t c = sl / rep;
    if (c > nc - 1) c = nc - 1;  // total_repeat_length pads with the last entry
    const int64_t p = perm[e * nc + c];
    // np.triu_indices(k, 1): row a starts at a (2k - a - 1) / 2
    const double kk = 2.0 * k - 1.0;
    int64_t a = (int64_t)floor((kk - sqrt(kk * kk - 8.0 * (double)p)) * 0.5);
    a = a < 0 ? 0 : (a > k - 2 ? k - 2 : a);
    while (a > 0 && a * (2LL * k - a - 1) / 2 > p) --a;
    while (a < k - 2 && (a + 1) * (2LL * k - a - 2) / 2 <= p) ++a;
    const int64_t b = p - a * (2LL * k - a - 1) / 2 + a + 1;
    out[(e * n + sl) * 2] = best[e * k + a];
    out[(e * n + sl) * 2 + 1] = best[e * k + b];
}

}  // namespace

int bg_launch_pairs_from_topk(const float *vals, const int32_t *idx, int64_t E, int k, int64_t row_len, int32_t *out, cudaStream_t st)
{
    BG_REQUIRE(k >= 1 && k <= PK_MAX, BG_ELIMIT, "bg_pairs_from_topk: k must be in 1..1024");
    BG_REQUIRE(row_len > 0 && E < (int64_t(1) << 31), BG_EINVAL, "bg_pairs_from_topk: bad shape");
    if (E == 0) return BG_OK;
    pairs_from_topk_kernel<<<(unsigned)E, PT, 0, st>>>(vals, idx, k, row_len, out);
    BG_LAUNCHED();
    return BG_OK;
}

int bg_launch_diallel_pairs(const int32_t *best, const int32_t *perm, int64_t E, int k, int nc, int64_t n, int32_t *out, cudaStream_t st)
{
    BG_REQUIRE(k >= 2 && nc >= 1 && n >= 1 && (int64_t)nc <= (int64_t)k * (k - 1) / 2, BG_EINVAL, "bg_diallel_pairs: bad shape");
    BG_REQUIRE(E < 65536, BG_ELIMIT, "bg_diallel_pairs: too many envs");
    if (E == 0) return BG_OK;
    const int rep = (int)((n + nc - 1) / nc);
    dim3 grid((unsigned)((n + PT - 1) / PT), (unsigned)E);
    diallel_pairs_kernel<<<grid, PT, 0, st>>>(best, perm, k, nc, n, rep, out);
    BG_LAUNCHED();
    return BG_OK;
}

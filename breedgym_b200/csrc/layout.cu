// Layout conversion, individual gathers and the reset permutation (sm_100a).
//
// bool[rows][m][2] (the reference's observation layout, breedgym/breedgym.py:47,
// breedgym/vector/vec_env.py:57-62) <-> packed bit planes uint32[rows][2][Wpad];
// `pop[idx]` gathers of whole individuals; and VecBreedGym.reset's
// `_random_selection` (breedgym/vector/vec_env.py:22-27): per env,
// jax.random.choice(key, germplasm, (n,), replace=False) = permutation(key, N)[:n],
// with permutation = repeated stable sort by fresh 32-bit keys (jax _shuffle).
#include "bg_internal.h"
#include "threefry.cuh"

namespace {

constexpr uint32_t FULL = 0xffffffffu;

// warp per (row, word): lane l reads the two allele bytes of marker 32w+l (coalesced 64 B),
// two ballots assemble the plane words.
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, int64_t rows,
                                                   int64_t m, int W, int Wpad)
{
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t total = rows * Wpad;
    if (gw >= total) return;
    const int64_t row = gw / Wpad;
    const int w = (int)(gw % Wpad);
    const int64_t j = (int64_t)w * 32 + lane;
    uint8_t a = 0, b = 0;
    if (w < W && j < m) {
        const uint8_t *p = in + (row * m + j) * 2;
        a = p[0];
        b = p[1];
    }
    const uint32_t wa = __ballot_sync(FULL, a != 0), wb = __ballot_sync(FULL, b != 0);
    if (lane == 0) {
        out[(row * 2) * Wpad + w] = wa;
        out[(row * 2 + 1) * Wpad + w] = wb;
    }
}

// thread per (row, marker): broadcast word read, coalesced 2-byte store
__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, int64_t rows,
                                                     int64_t m, int Wpad)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.y + (int64_t)blockIdx.z * 65535;
    if (j >= m || row >= rows) return;
    const uint32_t a = __ldg(in + (row * 2) * Wpad + (j >> 5)), b = __ldg(in + (row * 2 + 1) * Wpad + (j >> 5));
    uchar2 v;
    v.x = (a >> (j & 31)) & 1u;
    v.y = (b >> (j & 31)) & 1u;
    reinterpret_cast<uchar2 *>(out)[row * m + j] = v;
}

__device__ __forceinline__ int64_t norm_index(int64_t a, int64_t n)
{
    if (a < 0) a += n;
    a = a < 0 ? 0 : a;
    return a > n - 1 ? n - 1 : a;
}

// dst[e][r] = src[e*src_env_rows + idx[e][r]] ; one CTA per destination individual
// src_vals (optional): per-individual values [rows of src][T] (e.g. GEBVs) gathered along with the individuals
__global__ void __launch_bounds__(128) gather_kernel(const uint4 *__restrict__ src, const int32_t *__restrict__ idx,
                                                     uint4 *__restrict__ dst, int64_t n_src, int64_t n,
                                                     int64_t src_env_rows, int V, const float *__restrict__ src_vals,
                                                     float *__restrict__ dst_vals, int T)
{
    const int64_t d = blockIdx.x;  // e*n + r
    const int64_t e = d / n;
    const int64_t s = e * src_env_rows + norm_index(__ldg(idx + d), n_src);
    const uint4 *sp = src + s * V;
    uint4 *dp = dst + d * V;
    for (int v = threadIdx.x; v < V; v += blockDim.x) dp[v] = __ldg(sp + v);
    if (src_vals)
        for (int t = threadIdx.x; t < T; t += blockDim.x) dst_vals[d * T + t] = __ldg(src_vals + s * T + t);
}

// One CTA per env.  smem: sort keys [N], permutation x [N], scratch y [N].
template <int LAYOUT>
__global__ void __launch_bounds__(256) reset_indices_kernel(uint32_t k0, uint32_t k1, int64_t E_total, int64_t env_begin, int N,
                                                            int n, int rounds, int32_t *__restrict__ idx_out)
{
    extern __shared__ uint32_t sm[];
    uint32_t *keys = sm;
    int32_t *x = reinterpret_cast<int32_t *>(sm + N);
    int32_t *y = reinterpret_cast<int32_t *>(sm + 2 * (size_t)N);
    const int tid = threadIdx.x, NT = blockDim.x;
    const int64_t e = blockIdx.x;  // local env; logical env = env_begin + e
    // keys = split(random_key, E_total+1); env g uses keys[1+g]   (vec_env.py:120-128)
    TfKey key = tf_split_at(tf_make_key(k0, k1), (uint64_t)(env_begin + e + 1), (uint64_t)(E_total + 1), LAYOUT);
    for (int j = tid; j < N; j += NT) x[j] = j;
    for (int r = 0; r < rounds; ++r) {
        // key, subkey = split(key); sort_keys = random_bits(subkey, N)
        const TfKey sub = tf_split_at(key, 1, 2, LAYOUT);
        key = tf_split_at(key, 0, 2, LAYOUT);
        for (int j = tid; j < N; j += NT) keys[j] = tf_bits_at(sub, (uint64_t)j, (uint64_t)N, LAYOUT);
        __syncthreads();
        // stable rank sort: rank = #{i : key_i < key_j or (key_i == key_j and i < j)}
        for (int j = tid; j < N; j += NT) {
            const uint32_t kj = keys[j];
            int rank = 0;
            for (int i = 0; i < N; ++i) {
                const uint32_t ki = keys[i];
                rank += (ki < kj) || (ki == kj && i < j);
            }
            y[rank] = x[j];
        }
        __syncthreads();
        for (int j = tid; j < N; j += NT) x[j] = y[j];
        __syncthreads();
    }
    for (int j = tid; j < n; j += NT) idx_out[e * n + j] = x[j];
}

// 16-byte-wide copy between device memory and MAPPED pinned host memory (zero-copy over PCIe).  For the few hundred KB
// the host-facing step moves, a kernel that reads / writes the pinned buffer directly costs less than a copy-engine
// transfer (189 KB H2D: 18.5 us event to event through cudaMemcpyAsync).
__global__ void __launch_bounds__(256) copy16_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int64_t n16,
                                                     const uint32_t *__restrict__ src_tail, uint32_t *__restrict__ dst_tail, int ntail)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) dst[i] = src[i];
    if (i < ntail) dst_tail[i] = src_tail[i];
}

}  // namespace

// bytes % 4 == 0, both pointers 16-byte aligned and device-accessible
int bg_launch_copy_mapped(const void *src, void *dst, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return BG_OK;
    const int64_t n16 = (int64_t)(bytes / 16);
    const int ntail = (int)((bytes % 16) / 4);
    const int64_t threads = n16 > ntail ? n16 : ntail;
    copy16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const uint4 *)src, (uint4 *)dst, n16,
                                                                     (const uint32_t *)src + 4 * n16, (uint32_t *)dst + 4 * n16, ntail);
    BG_LAUNCHED();
    return BG_OK;
}

int bg_launch_pack(const uint8_t *in, uint32_t *out, int64_t rows, int64_t m, int W, int Wpad, cudaStream_t st)
{
    if (rows == 0) return BG_OK;
    const int64_t warps = rows * Wpad;
    const int64_t blocks = (warps + 7) / 8;
    BG_REQUIRE(blocks < (int64_t(1) << 31), BG_ELIMIT, "pack grid too large");
    pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, rows, m, W, Wpad);
    BG_LAUNCHED();
    return BG_OK;
}

int bg_launch_unpack(const uint32_t *in, uint8_t *out, int64_t rows, int64_t m, int Wpad, cudaStream_t st)
{
    if (rows == 0) return BG_OK;
    const int64_t zs = (rows + 65534) / 65535;
    BG_REQUIRE(zs <= 65535, BG_ELIMIT, "unpack grid too large");
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)(rows < 65535 ? rows : 65535), (unsigned)zs);
    unpack_kernel<<<grid, 256, 0, st>>>(in, out, rows, m, Wpad);
    BG_LAUNCHED();
    return BG_OK;
}

int bg_launch_gather(const uint32_t *src, const int32_t *idx, uint32_t *dst, int64_t E, int64_t n_src, int64_t n,
                     int64_t src_env_rows, int Wpad, cudaStream_t st, const float *src_vals, float *dst_vals, int T)
{
    if (E * n == 0) return BG_OK;
    BG_REQUIRE(E * n < (int64_t(1) << 31), BG_ELIMIT, "gather grid too large");
    gather_kernel<<<(unsigned)(E * n), 128, 0, st>>>((const uint4 *)src, idx, (uint4 *)dst, n_src, n, src_env_rows, 2 * Wpad / 4,
                                                     src_vals, dst_vals, T);
    BG_LAUNCHED();
    return BG_OK;
}

int bg_launch_reset_indices(bg_engine *eng, const uint32_t key[2], int64_t E_total, int64_t env_begin, int64_t E, int64_t N,
                            int64_t n, int layout, int32_t *idx_out, cudaStream_t st)
{
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    BG_REQUIRE(n <= N, BG_EINVAL, "Cannot take a larger sample than population when 'replace=False'");
    if (E == 0 || n == 0) return BG_OK;
    const size_t smem = (size_t)N * 12;
    BG_REQUIRE(smem <= (size_t)eng->max_smem_optin, BG_ELIMIT, "germplasm too large for the on-device reset (limit ~19k)");
    BG_REQUIRE(E < (int64_t(1) << 31), BG_ELIMIT, "too many envs");
    // rounds = ceil(3 ln N / ln(2^32 - 1))   (jax _shuffle)
    const int rounds = (int)ceil(3.0 * log((double)(N > 1 ? N : 1)) / log(4294967295.0));
    auto kern = layout == BG_LAYOUT_LEGACY ? reset_indices_kernel<BG_LAYOUT_LEGACY> : reset_indices_kernel<BG_LAYOUT_PARTITIONABLE>;
    if (smem > 48 * 1024) BG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)E, 256, smem, st>>>(key[0], key[1], E_total, env_begin, (int)N, (int)n, rounds, idx_out);
    BG_LAUNCHED();
    return BG_OK;
}

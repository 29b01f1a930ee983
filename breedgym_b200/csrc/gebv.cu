// GEBV kernels (sm_100a): bit-packed dosage x marker-effect reduction.
//
// Replaces chromax TraitModel.__call__ = dot(sum(pop,-1), effects) as reached from
// breedgym/breedgym.py:225-235 (Simulator.GEBV) and breedgym/vector/vec_env.py:
// 132-134 (Simulator.GEBV_model on all envs).
//
// Arithmetic: the float32 effects are converted once (bg_engine_set_map) to 64-bit
// fixed point w_fix = rint(w * 2^s_t), s_t chosen per trait so that 2*sum|w_fix|
// < 2^62.  Sums of w_fix are exact integers, so the result does not depend on the
// summation order (atomics, strips, lanes) and is rounded exactly once, to float32.
//
// Two kernels produce the same integers:
//   direct  : warp per individual, lane <-> marker bit, shuffle-tree reduction
//   byte-LUT: CTA owns a 256-marker strip; a [32][256] table of partial sums of the
//             strip's effects is built in shared memory; every haplotype byte then
//             costs one 64-bit shared load + add; lane <-> individual, so no
//             cross-lane reduction is needed; strips combine with 64-bit atomics.
#include "bg_internal.h"

namespace {

constexpr uint32_t FULL = 0xffffffffu;

__global__ void __launch_bounds__(256) gebv_direct_kernel(const uint32_t *__restrict__ pop, int64_t rows, int W, int Wpad,
                                                          const long long *__restrict__ wfix, int T, int64_t mpad,
                                                          unsigned long long *__restrict__ acc)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= rows) return;
    const uint32_t *h0 = pop + (2 * i) * Wpad, *h1 = h0 + Wpad;
    for (int t = 0; t < T; ++t) {
        const long long *wt = wfix + (int64_t)t * mpad;
        long long s = 0;
        for (int w = 0; w < W; ++w) {
            const uint32_t a = __ldg(h0 + w), b = __ldg(h1 + w);
            const long long d = (long long)(((a >> lane) & 1u) + ((b >> lane) & 1u));
            s += d * __ldg(wt + (int64_t)w * 32 + lane);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (lane == 0) acc[i * T + t] = (unsigned long long)s;
    }
}

// strip = 8 words (256 markers) of both planes; LUT[pos][v] = sum of the effects of the
// set bits of byte value v at byte position pos of the strip.
__global__ void __launch_bounds__(256) gebv_lut_kernel(const uint32_t *__restrict__ pop, int64_t rows, int Wpad,
                                                       const long long *__restrict__ wfix, int64_t mpad,
                                                       unsigned long long *__restrict__ acc, int T, int64_t rows_per_cta)
{
    extern __shared__ __align__(16) long long lut[];  // [32][256]
    __shared__ long long wloc[256];
    const int strip = blockIdx.x, t = blockIdx.z;
    const int tid = threadIdx.x;
    wloc[tid] = __ldg(wfix + (int64_t)t * mpad + (int64_t)strip * 256 + tid);
    __syncthreads();
    for (int e = tid; e < 32 * 256; e += 256) {
        const int pos = e >> 8, v = e & 255;
        long long s = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) s += ((v >> b) & 1) ? wloc[pos * 8 + b] : 0ll;
        lut[e] = s;
    }
    __syncthreads();
    const int w0 = strip * 8;
    const bool second = (w0 + 4) < Wpad;  // Wpad is a multiple of 4: a strip holds 4 or 8 valid words
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int64_t i = r0 + tid; i < r1; i += 256) {
        long long s = 0;
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
            const uint4 *p = reinterpret_cast<const uint4 *>(pop + (2 * i + hp) * Wpad + w0);
            uint4 x = __ldg(p);
            uint4 y = second ? __ldg(p + 1) : make_uint4(0, 0, 0, 0);
            const uint32_t wd[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s += lut[(j * 4 + k) * 256 + ((wd[j] >> (8 * k)) & 255u)];
            }
        }
        atomicAdd(acc + i * T + t, (unsigned long long)s);
    }
}

__global__ void gebv_finalize_kernel(unsigned long long *__restrict__ acc, const double *__restrict__ inv_scale, int T,
                                     int64_t total, float *__restrict__ out, bool clear)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long v = (long long)acc[i];
    out[i] = (float)((double)v * inv_scale[i % T]);
    if (clear) acc[i] = 0ull;
}

// op 0: max, op 1: mean (float64 accumulation)
__global__ void __launch_bounds__(256) reduce_env_kernel(const float *__restrict__ in, int64_t per_env,
                                                         float *__restrict__ out, int op)
{
    __shared__ double sh[8];
    const float *x = in + (int64_t)blockIdx.x * per_env;
    double v = op == 0 ? -INFINITY : 0.0;
    for (int64_t i = threadIdx.x; i < per_env; i += blockDim.x) {
        const double f = (double)x[i];
        v = op == 0 ? fmax(v, f) : v + f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double u = __shfl_xor_sync(FULL, v, o);
        v = op == 0 ? fmax(v, u) : v + u;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = op == 0 ? fmax(v, sh[w]) : v + sh[w];
        out[blockIdx.x] = op == 0 ? (float)v : (float)(v / (double)per_env);
    }
}

}  // namespace

int bg_launch_gebv(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, int algo, cudaStream_t st)
{
    BG_REQUIRE(eng && eng->d_wfix, BG_ESTATE, "engine has no map (call bg_engine_set_map)");
    BG_REQUIRE(algo >= 0 && algo <= 3, BG_EINVAL, "bad GEBV algorithm id");
    if (rows == 0 || eng->T == 0) return BG_OK;
    const int T = eng->T;
    if (algo == 0) {
        algo = eng->opt.gebv_algo;
        if (algo < 1 || algo > 3) algo = eng->tc_N ? 3 : 2;
    }
    if (algo == 3) {
        BG_REQUIRE(eng->tc_N, BG_ELIMIT, "too many traits for the tensor-core GEBV (digits x traits <= 256)");
        return bg_launch_gebv_tc2(eng, pop, rows, out, st);
    }
    const int64_t total = rows * T, mpad = (int64_t)((eng->Wpad + 7) / 8) * 256;  // = bg_engine wfix row stride
    int rc = bg_reserve_acc(eng, (size_t)total);
    if (rc) return rc;

    if (algo == 1) {
        const int wpb = 8;
        gebv_direct_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(pop, rows, eng->W, eng->Wpad, eng->d_wfix, T,
                                                                                  mpad, eng->d_acc);
        BG_LAUNCHED();
        gebv_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(eng->d_acc, eng->d_inv_scale, T, total, out, false);
        BG_LAUNCHED();
        return BG_OK;
    }
    BG_REQUIRE(T <= 65535, BG_ELIMIT, "too many traits for the LUT kernel grid");
    const int strips = (eng->W + 7) / 8;
    // row split: aim at ~4 CTAs per SM overall, but keep >= 1024 rows per CTA so the
    // 64 KB table build is amortised
    int64_t want = (4LL * eng->sm_count + (int64_t)strips * T - 1) / ((int64_t)strips * T);
    int64_t maxsplit = (rows + 1023) / 1024;
    int64_t ysplit = want < 1 ? 1 : want;
    if (ysplit > maxsplit) ysplit = maxsplit;
    if (ysplit > 65535) ysplit = 65535;
    const int64_t rows_per_cta = (rows + ysplit - 1) / ysplit;
    const size_t smem = 32 * 256 * sizeof(long long);
    BG_CUDA(cudaFuncSetAttribute(gebv_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BG_CUDA(cudaMemsetAsync(eng->d_acc, 0, (size_t)total * sizeof(unsigned long long), st));
    dim3 grid((unsigned)strips, (unsigned)ysplit, (unsigned)T);
    gebv_lut_kernel<<<grid, 256, smem, st>>>(pop, rows, eng->Wpad, eng->d_wfix, mpad, eng->d_acc, T, rows_per_cta);
    BG_LAUNCHED();
    gebv_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(eng->d_acc, eng->d_inv_scale, T, total, out, false);
    BG_LAUNCHED();
    return BG_OK;
}

int bg_launch_reduce(const float *in, int64_t E, int64_t per_env, float *out, int op, cudaStream_t st)
{
    if (E == 0) return BG_OK;
    BG_REQUIRE(per_env > 0, BG_EINVAL, "empty reduction");
    BG_REQUIRE(E < (int64_t(1) << 31), BG_ELIMIT, "too many envs");
    reduce_env_kernel<<<(unsigned)E, 256, 0, st>>>(in, per_env, out, op);
    BG_LAUNCHED();
    return BG_OK;
}

// Meiosis / cross kernels (sm_100a).
//
// Replaces chromax functional.cross/_meiosis as called from
// breedgym/breedgym.py:142-143 and breedgym/vector/vec_env.py:75-77,89-91:
//   u = uniform(key,(m,)); s = u < r; mask = cumulative XOR(s); hap[j] = ind[j, mask[j]]
//
// One CTA per gamete row q.  Lanes evaluate Threefry blocks, __ballot_sync packs 32
// compare results into a recombination word held in shared memory (m/8 bytes per
// row, <= 227 KB => m <= ~1.8 M markers), an in-word shift-XOR ladder plus a
// ballot/popc warp scan and a block scan turn it into the crossover mask, and the
// mask then either goes to global memory (vector env: masks are shared by all
// envs, see blend_envs_kernel) or selects alleles from the two bit-plane rows of
// the parent with 128-bit loads/stores (unique-key cross, double haploid).
#include <algorithm>

#include "bg_internal.h"
#include "threefry.cuh"

namespace {

constexpr int ILP = 4;
constexpr uint32_t FULL = 0xffffffffu;

struct RowParams {
    const uint32_t *thr;      // T[j] = clamp(ceil(r_j 2^23), 0, 2^23):  draw <=> (bits >> 9) < T
    const uint32_t *thr_cmp;  // T[j] << 9 (NULL when some T = 2^23 would overflow): draw <=> bits < T << 9, one compare, no shift
    uint32_t mut_thr;
    uint32_t m, W, Wpad;
    uint32_t keys[BG_BATCH_MAX][2];  // mask mode: grid row g <-> gamete row g % rows of key g / rows (one launch per batch of keys)
    uint64_t rows;
    uint64_t total_rows;  // rows x (keys | groups): the grid covers them one per CTA or in a strided loop
    uint32_t same_key;  // 1: every group of `rows` grid rows uses keys[0] (double haploids of E envs under ONE key)
    int schedule;
    int mode;
    uint32_t *mask_out;
    uint32_t *mut_out;
    const uint32_t *pop;
    const int32_t *parents;
    int64_t n_src;
    int64_t dh_offspring;
    uint32_t *out;
    uint32_t one;  // = 1 (see tf2x32_n)
};

// `one` is the constant 1 read from the kernel parameters: x0 = x1 * one + x0 keeps the Threefry
// additions on the FMA pipe (IMAD) while rotate (SHF) and xor (LOP3) use the ALU pipe, so the two
// integer pipes share the ~100 operations of a block instead of all of them queueing on one.
#ifndef BG_TF_MULROT
#define BG_TF_MULROT 0x00000   // bit i: round i rotates through a 64-bit multiply (FMA pipe) instead of a funnel shift (ALU pipe)
#endif
template <int N>
__device__ __forceinline__ void tf2x32_n(const TfKey &k, uint32_t (&x0)[N], uint32_t (&x1)[N], const uint32_t one)
{
    // x * 2^r as a 64-bit product (IMAD.WIDE): lo | hi is the rotation, and the OR folds into the round's XOR (one LOP3)
#define BG_R(i, r)                                                                     \
    _Pragma("unroll") for (int u = 0; u < N; ++u)                                      \
    {                                                                                  \
        x0[u] = x1[u] * one + x0[u];                                                   \
        if ((BG_TF_MULROT >> (i)) & 1) {                                               \
            const unsigned long long p = (unsigned long long)x1[u] * (unsigned long long)(one << (r)); \
            x1[u] = ((uint32_t)p | (uint32_t)(p >> 32)) ^ x0[u];                       \
        } else {                                                                       \
            x1[u] = __funnelshift_l(x1[u], x1[u], r) ^ x0[u];                          \
        }                                                                              \
    }
#define BG_INJ(a, b)                             \
    _Pragma("unroll") for (int u = 0; u < N; ++u) \
    {                                            \
        x0[u] += (a);                            \
        x1[u] += (b);                            \
    }
    // key-schedule words with the round counter folded in (uniform per row: computed once, not per block)
    const uint32_t i1 = k.k2 + 1u, i2 = k.k0 + 2u, i3 = k.k1 + 3u, i4 = k.k2 + 4u, i5 = k.k0 + 5u;
    BG_INJ(k.k0, k.k1)
    BG_R(0, 13) BG_R(1, 15) BG_R(2, 26) BG_R(3, 6) BG_INJ(k.k1, i1)
    BG_R(4, 17) BG_R(5, 29) BG_R(6, 16) BG_R(7, 24) BG_INJ(k.k2, i2)
    BG_R(8, 13) BG_R(9, 15) BG_R(10, 26) BG_R(11, 6) BG_INJ(k.k0, i3)
    BG_R(12, 17) BG_R(13, 29) BG_R(14, 16) BG_R(15, 24) BG_INJ(k.k1, i4)
    BG_R(16, 13) BG_R(17, 15) BG_R(18, 26) BG_R(19, 6) BG_INJ(k.k2, i5)
#undef BG_R
#undef BG_INJ
}

// Draw the m Bernoulli bits `uniform(key)[j] < thr[j]` of one row into the zeroed
// shared bit array S (marker j -> bit j&31 of S[j>>5]).
// The integer (ALU) pipe bounds this loop: a Threefry block is 20 funnel shifts + 20 LOP3 there (the 35 additions go to
// the FMA pipe as IMAD), so everything else on that pipe -- validity tests, selects, address arithmetic, the >> 9 of
// the compare -- is overhead.  Groups of 32 counters whose draws are all inside the row take the FAST path: no
// validity tests, thresholds preshifted (`bits < T << 9`), one uniform branch per ILP groups around the (rare)
// recombination events; only the last, partial groups of a row go through the checked path.
template <int LAYOUT, bool CONST_THR>
__device__ __forceinline__ void draw_bits(uint32_t *S, const TfKey key, const uint32_t *__restrict__ thr,
                                          const uint32_t *__restrict__ thr_cmp, uint32_t cthr, uint32_t m, uint32_t lane,
                                          uint32_t warp, uint32_t NW, const uint32_t one)
{
    const bool fast_ok = CONST_THR ? (cthr < (1u << 23)) : (thr_cmp != nullptr);
    const uint32_t ccmp = cthr << 9;
    if (LAYOUT == BG_LAYOUT_LEGACY) {
        // block c yields draw c (word 0) and draw c+h (word 1): two bit streams, the
        // second starting at bit offset h&31 of word h>>5.
        const uint32_t h = (m + 1) >> 1, G = (h + 31) >> 5, sh = h & 31, wsB = h >> 5;
        const uint32_t Gfull = fast_ok ? ((m & 1) ? (h - 1) >> 5 : h >> 5) : 0u;  // groups with all 64 draws inside the row
        for (uint32_t g0 = warp * ILP; g0 < G; g0 += NW * ILP) {
            uint32_t x0[ILP], x1[ILP], tA[ILP], tB[ILP], bA[ILP], bB[ILP];
            if (g0 + ILP <= Gfull) {
                const uint32_t c0 = g0 * 32 + lane;
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    x0[u] = c0 + 32 * u;
                    x1[u] = c0 + 32 * u + h;
                    tA[u] = CONST_THR ? ccmp : __ldg(thr_cmp + c0 + 32 * u);
                    tB[u] = CONST_THR ? ccmp : __ldg(thr_cmp + c0 + 32 * u + h);
                }
                tf2x32_n<ILP>(key, x0, x1, one);
                uint32_t any = 0;
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    bA[u] = __ballot_sync(FULL, x0[u] < tA[u]);
                    bB[u] = __ballot_sync(FULL, x1[u] < tB[u]);
                    any |= bA[u] | bB[u];
                }
                if (any == 0u) continue;  // recombination events are rare (r ~ 1e-3): most iterations end here
            } else {
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    const uint32_t c = (g0 + u) * 32 + lane, cB = c + h;
                    const bool vA = c < h, vB = vA && cB < m;
                    x0[u] = c;
                    x1[u] = vB ? cB : 0u;  // odd m: the last block's second counter is the zero pad
                    tA[u] = vA ? (CONST_THR ? cthr : __ldg(thr + c)) : 0u;
                    tB[u] = vB ? (CONST_THR ? cthr : __ldg(thr + cB)) : 0u;
                }
                tf2x32_n<ILP>(key, x0, x1, one);
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    bA[u] = __ballot_sync(FULL, (x0[u] >> 9) < tA[u]);
                    bB[u] = __ballot_sync(FULL, (x1[u] >> 9) < tB[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < ILP; ++u) {
                const uint32_t g = g0 + u;
                if ((bA[u] | bB[u]) != 0u && g < G) {  // stitch the two streams (warp-uniform test)
                    if (lane == 0 && bA[u]) atomicOr(&S[g], bA[u]);
                    if (lane == 1 && bB[u]) atomicOr(&S[wsB + g], bB[u] << sh);
                    if (lane == 2 && sh && (bB[u] >> (32 - sh))) atomicOr(&S[wsB + g + 1], bB[u] >> (32 - sh));
                }
            }
        }
    } else {
        const uint32_t G = (m + 31) >> 5;
        const uint32_t Gfull = fast_ok ? m >> 5 : 0u;
        for (uint32_t g0 = warp * ILP; g0 < G; g0 += NW * ILP) {
            uint32_t x0[ILP], x1[ILP], tA[ILP];
            if (g0 + ILP <= Gfull) {
                const uint32_t c0 = g0 * 32 + lane;
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    x0[u] = 0u;
                    x1[u] = c0 + 32 * u;
                    tA[u] = CONST_THR ? ccmp : __ldg(thr_cmp + c0 + 32 * u);
                }
                tf2x32_n<ILP>(key, x0, x1, one);
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    const uint32_t b = __ballot_sync(FULL, (x0[u] ^ x1[u]) < tA[u]);
                    if (lane == 0) S[g0 + u] = b;
                }
            } else {
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    const uint32_t c = (g0 + u) * 32 + lane;
                    x0[u] = 0u;
                    x1[u] = c;
                    tA[u] = (c < m) ? (CONST_THR ? cthr : __ldg(thr + c)) : 0u;
                }
                tf2x32_n<ILP>(key, x0, x1, one);
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    const uint32_t g = g0 + u;
                    const uint32_t b = __ballot_sync(FULL, ((x0[u] ^ x1[u]) >> 9) < tA[u]);
                    if (g < G && lane == 0) S[g] = b;
                }
            }
        }
    }
}

__device__ __forceinline__ uint32_t inword_xor_scan(uint32_t x)
{
    x ^= x << 1;
    x ^= x << 2;
    x ^= x << 4;
    x ^= x << 8;
    x ^= x << 16;
    return x;
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t r;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
    return r;
}

__device__ __forceinline__ int64_t norm_index(int64_t a, int64_t n)
{
    if (a < 0) a += n;
    a = a < 0 ? 0 : a;
    return a > n - 1 ? n - 1 : a;
}

__device__ __forceinline__ uint4 blend4(uint4 h0, uint4 h1, uint4 M)
{
    uint4 o;
    o.x = (h0.x & ~M.x) | (h1.x & M.x);
    o.y = (h0.y & ~M.y) | (h1.y & M.y);
    o.z = (h0.z & ~M.z) | (h1.z & M.z);
    o.w = (h0.w & ~M.w) | (h1.w & M.w);
    return o;
}

template <int LAYOUT, int NT_MAX>
__global__ void __launch_bounds__(NT_MAX, NT_MAX == 256 ? 4 : (NT_MAX == 512 ? 2 : 1)) meiosis_rows_kernel(const RowParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t wtot[32];
    __shared__ uint32_t row_keys[4];  // (rec, mut) keys of the row in progress
    // a dependent kernel launched with programmatic stream serialization (the GEBV of these offspring, gebv_tc2.cu) may
    // run its prologue while this grid drains; kernels launched the ordinary way are not affected
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    uint32_t *S = smem;
    uint32_t *Mu = smem + (P.Wpad + 8);
    const uint32_t tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
    const bool has_mut = P.mut_thr != 0;
    const uint32_t W = P.W, Wpad = P.Wpad, m = P.m;
    // one row per CTA, or (persistent launch, option mask_ctas_per_sm: fewer CTAs than rows) a strided loop over the rows
    for (uint64_t grow = blockIdx.x; grow < P.total_rows; grow += gridDim.x) {  // row of the output arrays
    const uint32_t kb = (uint32_t)(grow / P.rows);      // which key of the batch (0 unless mask mode)
    const uint64_t q = grow - (uint64_t)kb * P.rows;   // gamete row of that key

    for (uint32_t i = tid; i < Wpad + 8; i += NT) {
        S[i] = 0;
        if (has_mut) Mu[i] = 0;
    }
    // per-gamete key: #q of split(k, rows); S2 splits it again into (rec, mut).  Two or three Threefry blocks that
    // every warp used to repeat (a warp instruction costs the same for 1 or 32 lanes, so the redundancy across WARPS
    // is what costs): warp 0 derives them, the others pick them up behind the barrier that follows the zero fill.
    if (warp == 0) {
        const uint32_t ki = P.same_key ? 0u : kb;
        const TfKey kc = tf_make_key(P.keys[ki][0], P.keys[ki][1]);
        const TfKey kq = tf_split_at(kc, q, P.rows, LAYOUT);
        TfKey kr = kq, km = kq;
        if (P.schedule == BG_SCHEDULE_S2) {
            kr = tf_split_at(kq, 0, 2, LAYOUT);
            if (has_mut) km = tf_split_at(kq, 1, 2, LAYOUT);
        }
        if (lane == 0) {
            row_keys[0] = kr.k0;
            row_keys[1] = kr.k1;
            row_keys[2] = km.k0;
            row_keys[3] = km.k1;
        }
    }
    __syncthreads();
    const TfKey krec = tf_make_key(row_keys[0], row_keys[1]), kmut = tf_make_key(row_keys[2], row_keys[3]);

#if defined(BG_FAKE_CONST_THR)   // timing experiment only (wrong masks): what do the threshold loads cost beside the step kernel?
    draw_bits<LAYOUT, true>(S, krec, nullptr, nullptr, 12000u, m, lane, warp, NW, P.one);
#else
    draw_bits<LAYOUT, false>(S, krec, P.thr, P.thr_cmp, 0u, m, lane, warp, NW, P.one);
#endif
    if (has_mut) draw_bits<LAYOUT, true>(Mu, kmut, nullptr, nullptr, P.mut_thr, m, lane, warp, NW, P.one);
    __syncthreads();

    // inclusive prefix-XOR over the whole row, in place: each warp owns a contiguous
    // chunk (multiple of 32 words), then chunk parities are combined across warps.
    const uint32_t CH = ((((W + NW - 1) / NW) + 31) >> 5) << 5;
    const uint32_t wbeg = warp * CH, wend = min(W, wbeg + CH);
    uint32_t carry = 0;
    for (uint32_t w0 = wbeg; w0 < wend; w0 += 32) {
        const uint32_t w = w0 + lane;
        uint32_t x = inword_xor_scan(w < wend ? S[w] : 0u);
        const uint32_t b = __ballot_sync(FULL, x >> 31);
        const uint32_t pre = (__popc(b & lanemask_lt()) & 1u) ^ carry;
        x ^= 0u - pre;
        carry ^= __popc(b) & 1u;
        if (w < wend) S[w] = x;
    }
    if (lane == 0) wtot[warp] = carry;
    __syncthreads();
    {
        const uint32_t v = (lane < warp) ? wtot[lane] : 0u;
        const uint32_t cin = __popc(__ballot_sync(FULL, v)) & 1u;
        const uint32_t tail = m & 31;
        for (uint32_t w = wbeg + lane; w < wend; w += 32) {
            uint32_t x = S[w] ^ (0u - cin);
            if (w == W - 1 && tail) x &= (1u << tail) - 1u;  // keep padding bits zero
            S[w] = x;
        }
    }
    __syncthreads();

    const uint32_t W4 = Wpad >> 2;
    const uint4 *S4 = reinterpret_cast<const uint4 *>(S);
    const uint4 *Mu4 = reinterpret_cast<const uint4 *>(Mu);
    if (P.mode == BG_ROWS_MASK) {
        uint4 *mo = reinterpret_cast<uint4 *>(P.mask_out + grow * Wpad);
        for (uint32_t v = tid; v < W4; v += NT) mo[v] = S4[v];
        if (has_mut) {
            uint4 *uo = reinterpret_cast<uint4 *>(P.mut_out + grow * Wpad);
            for (uint32_t v = tid; v < W4; v += NT) uo[v] = Mu4[v];
        }
        __syncthreads();  // the row buffer is reused by the next row
        continue;
    }
    int64_t src;
    uint4 *dst0, *dst1 = nullptr;
    if (P.mode == BG_ROWS_CROSS) {
        src = norm_index(P.parents[q], P.n_src);
        dst0 = reinterpret_cast<uint4 *>(P.out + q * Wpad);
    } else {  // double haploid: the gamete fills both planes of individual `grow` (env-major: env = grow / rows)
        src = (int64_t)(grow / (uint64_t)P.dh_offspring);
        dst0 = reinterpret_cast<uint4 *>(P.out + (2 * grow) * Wpad);
        dst1 = dst0 + W4;
    }
    const uint4 *h0 = reinterpret_cast<const uint4 *>(P.pop + (uint64_t)(2 * src) * Wpad);
    const uint4 *h1 = h0 + W4;
    for (uint32_t v = tid; v < W4; v += NT) {
        uint4 o = blend4(__ldg(h0 + v), __ldg(h1 + v), S4[v]);
        if (has_mut) {
            const uint4 u = Mu4[v];
            o.x ^= u.x; o.y ^= u.y; o.z ^= u.z; o.w ^= u.w;
        }
        dst0[v] = o;
        if (dst1) dst1[v] = o;
    }
    __syncthreads();
    }  // rows
}

// Vector env: out[e][q] = blend(pop[e][parents[e][q]] planes, mask[q]) for all envs e.
// The mask (and mutation) words are loaded once per thread and reused for every env.
template <bool HAS_MUT>
__global__ void __launch_bounds__(256) blend_envs_kernel(const uint4 *__restrict__ pop, const int32_t *__restrict__ parents,
                                                         const uint4 *__restrict__ mask, const uint4 *__restrict__ mut,
                                                         uint4 *__restrict__ out, int E, int64_t n_src, int64_t rows,
                                                         int W4, int env_chunk)
{
    // let a dependent kernel (the GEBV of this population) start its prologue while this grid drains
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int v = blockIdx.y * blockDim.x + threadIdx.x;
    if (v >= W4) return;
    const int64_t q = blockIdx.x;
    const uint4 M = __ldg(mask + q * W4 + v);
    uint4 U = make_uint4(0, 0, 0, 0);
    if (HAS_MUT) U = __ldg(mut + q * W4 + v);
    const int e0 = blockIdx.z * env_chunk, e1 = min(E, e0 + env_chunk);
    constexpr int UN = 4;
    // running pointers (one 64-bit add per env instead of re-deriving every address)
    const int64_t pop_env = n_src * 2 * W4, out_env = rows * W4;
    const int32_t *pp = parents + (int64_t)e0 * rows + q;
    const uint4 *pe = pop + (int64_t)e0 * pop_env + v;
    uint4 *oe = out + ((int64_t)e0 * rows + q) * W4 + v;
    const int nsrc = (int)n_src, row2 = 2 * W4;
    for (int e = e0; e < e1; e += UN) {
        uint4 a[UN], b[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (e + u < e1) {
                int s = __ldg(pp + (int64_t)u * rows);
                s += s < 0 ? nsrc : 0;  // jnp indexing: negatives wrap once, then clamp
                s = min(max(s, 0), nsrc - 1);
                const uint4 *h0 = pe + u * pop_env + (int64_t)s * row2;
                a[u] = __ldg(h0);
                b[u] = __ldg(h0 + W4);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (e + u < e1) {
                uint4 o = blend4(a[u], b[u], M);
                if (HAS_MUT) {
                    o.x ^= U.x; o.y ^= U.y; o.z ^= U.z; o.w ^= U.w;
                }
                oe[u * out_env] = o;
            }
        }
        pp += (int64_t)UN * rows;
        pe += UN * pop_env;
        oe += UN * out_env;
    }
}

}  // namespace

static int launch_rows(bg_engine *eng, int mode, int64_t rows, int nkeys, const uint32_t (*keys)[2], int layout, int schedule,
                       uint32_t *mask_out, uint32_t *mut_out, const uint32_t *pop, const int32_t *parents, int64_t n_src,
                       int64_t dh_offspring, uint32_t *out, cudaStream_t st, int small_ctas, int64_t groups = 0)
{
    // groups > 0: `groups` x rows grid rows, every group under keys[0] (nkeys == 1)
    const int64_t grid_groups = groups > 0 ? groups : nkeys;
    BG_REQUIRE(eng && eng->d_thr, BG_ESTATE, "engine has no map (call bg_engine_set_map)");
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    BG_REQUIRE(schedule == BG_SCHEDULE_S1 || schedule == BG_SCHEDULE_S2, BG_EINVAL, "bad key schedule");
    BG_REQUIRE(!(schedule == BG_SCHEDULE_S1 && eng->mut_thr), BG_EINVAL, "schedule S1 has no mutation key");
    BG_REQUIRE(nkeys >= 1 && nkeys <= BG_BATCH_MAX, BG_EINVAL, "bad mask batch size");
    BG_REQUIRE(rows >= 0 && rows * grid_groups < (int64_t(1) << 31), BG_ELIMIT, "too many gamete rows");
    if (rows == 0) return BG_OK;
    const bool has_mut = eng->mut_thr != 0;
    const size_t smem = (size_t)(eng->Wpad + 8) * 4 * (has_mut ? 2 : 1);
    BG_REQUIRE(smem <= (size_t)eng->max_smem_optin, BG_ELIMIT,
               "n_markers too large for the shared-memory row buffer (limit ~1.8M markers, half with mutation)");
    RowParams P;
    P.thr = eng->d_thr;
    P.thr_cmp = eng->d_thr_cmp;
    P.mut_thr = eng->mut_thr;
    P.m = (uint32_t)eng->m;
    P.W = (uint32_t)eng->W;
    P.Wpad = (uint32_t)eng->Wpad;
    for (int b = 0; b < BG_BATCH_MAX; ++b) {
        P.keys[b][0] = b < nkeys ? keys[b][0] : 0u;
        P.keys[b][1] = b < nkeys ? keys[b][1] : 0u;
    }
    P.rows = (uint64_t)rows;
    P.total_rows = (uint64_t)(rows * grid_groups);
    P.same_key = groups > 0 ? 1u : 0u;
    P.schedule = schedule;
    P.mode = mode;
    P.mask_out = mask_out;
    P.mut_out = mut_out;
    P.pop = pop;
    P.parents = parents;
    P.n_src = n_src;
    P.dh_offspring = dh_offspring > 0 ? dh_offspring : 1;
    P.out = out;
    P.one = 1u;
    // small_ctas: 128-thread CTAs (8192 registers) for mask kernels that run beside the fused step kernel, see cross_gebv.cu
    // CTA size by row length: short rows 256 threads (4 CTAs per SM; 128 when all rows then fit in ONE wave of 8 per SM, or
    // beside the step kernel), mid rows 512 (2 per SM: one CTA's prologue / scan / store phases overlap the other's draw
    // phase), long rows 1024
    const int64_t total_rows = rows * (groups > 0 ? groups : nkeys);
    int NT;
    if (eng->W <= 1024) {
        NT = small_ctas ? eng->opt.mask_nt : 256;
        if (!small_ctas && total_rows > 4LL * eng->sm_count && total_rows <= 8LL * eng->sm_count) NT = 128;
    } else {
        NT = eng->W <= 8192 ? 512 : 1024;
    }
    if (eng->opt.rows_nt > 0 && !small_ctas) NT = eng->opt.rows_nt;
    void (*kern)(RowParams);
    if (NT <= 256)
        kern = layout == BG_LAYOUT_LEGACY ? meiosis_rows_kernel<BG_LAYOUT_LEGACY, 256> : meiosis_rows_kernel<BG_LAYOUT_PARTITIONABLE, 256>;
    else if (NT <= 512)
        kern = layout == BG_LAYOUT_LEGACY ? meiosis_rows_kernel<BG_LAYOUT_LEGACY, 512> : meiosis_rows_kernel<BG_LAYOUT_PARTITIONABLE, 512>;
    else
        kern = layout == BG_LAYOUT_LEGACY ? meiosis_rows_kernel<BG_LAYOUT_LEGACY, 1024> : meiosis_rows_kernel<BG_LAYOUT_PARTITIONABLE, 1024>;
    if (smem > 48 * 1024) BG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // small_ctas (lookahead batches beside the step kernel): a persistent grid of mask_ctas_per_sm small CTAs per SM
    int64_t grid = rows * grid_groups;
    if (small_ctas && NT <= 256 && eng->opt.mask_ctas_per_sm > 0) grid = std::min<int64_t>(grid, (int64_t)eng->opt.mask_ctas_per_sm * eng->sm_count);
    kern<<<(unsigned)grid, NT, smem, st>>>(P);
    BG_LAUNCHED();
    return BG_OK;
}

int bg_launch_meiosis_rows(bg_engine *eng, int mode, int64_t rows, const uint32_t cross_key[2], int layout, int schedule,
                           uint32_t *mask_out, uint32_t *mut_out, const uint32_t *pop, const int32_t *parents,
                           int64_t n_src, int64_t dh_offspring, uint32_t *out, cudaStream_t st, int small_ctas)
{
    const uint32_t keys[1][2] = {{cross_key[0], cross_key[1]}};
    return launch_rows(eng, mode, rows, 1, keys, layout, schedule, mask_out, mut_out, pop, parents, n_src, dh_offspring, out, st,
                       small_ctas);
}

// double haploids of E populations [E][n][2][Wpad] under ONE key: out [E][n][n_offspring][2][Wpad]
int bg_launch_double_haploid(bg_engine *eng, int64_t E, int64_t n, int64_t n_offspring, const uint32_t cross_key[2], int layout,
                             int schedule, const uint32_t *pop, uint32_t *out, cudaStream_t st)
{
    const uint32_t keys[1][2] = {{cross_key[0], cross_key[1]}};
    return launch_rows(eng, BG_ROWS_DH, n * n_offspring, 1, keys, layout, schedule, nullptr, nullptr, pop, nullptr, n, n_offspring, out,
                       st, 0, E);
}

// masks of `nkeys` cross keys in one launch: mask_out / mut_out [nkeys][rows][Wpad]
int bg_launch_mask_batch(bg_engine *eng, int64_t rows, int nkeys, const uint32_t (*keys)[2], int layout, int schedule,
                         uint32_t *mask_out, uint32_t *mut_out, cudaStream_t st, int small_ctas)
{
    return launch_rows(eng, BG_ROWS_MASK, rows, nkeys, keys, layout, schedule, mask_out, mut_out, nullptr, nullptr, 0, 0, nullptr, st,
                       small_ctas);
}

int bg_launch_blend(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask,
                    const uint32_t *mut, uint32_t *out, int64_t E, int64_t n_src, int64_t n, cudaStream_t st)
{
    const int W4 = eng->Wpad / 4;
    const int64_t rows = 2 * n;
    if (rows == 0 || E == 0) return BG_OK;
    BG_REQUIRE(E < (int64_t(1) << 31) && rows < (int64_t(1) << 31), BG_ELIMIT, "blend grid too large");
    int threads = ((W4 + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    const int tiles = (W4 + threads - 1) / threads;
    BG_REQUIRE(tiles <= 65535, BG_ELIMIT, "n_markers too large for the blend grid");
    // env chunk: enough CTAs for several waves, but >= 4 envs per CTA so the mask load is amortised
    int chunk = eng->opt.blend_env_chunk;
    int64_t zs = (E + chunk - 1) / chunk;
    if (zs > 65535) {
        chunk = (int)((E + 65534) / 65535);
        zs = (E + chunk - 1) / chunk;
    }
    dim3 grid((unsigned)rows, (unsigned)tiles, (unsigned)zs);
    if (mut)
        blend_envs_kernel<true><<<grid, threads, 0, st>>>((const uint4 *)pop, parents, (const uint4 *)mask, (const uint4 *)mut,
                                                          (uint4 *)out, (int)E, n_src, rows, W4, chunk);
    else
        blend_envs_kernel<false><<<grid, threads, 0, st>>>((const uint4 *)pop, parents, (const uint4 *)mask, nullptr,
                                                           (uint4 *)out, (int)E, n_src, rows, W4, chunk);
    BG_LAUNCHED();
    return BG_OK;
}

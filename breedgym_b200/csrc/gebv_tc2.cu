// GEBV on tcgen05, second generation: TMA tile loads + the dosage operand in TENSOR MEMORY.
//
// Same exact int8 GEMM as gebv_tc.cu (D[i, 8t+d] = sum_j dosage[i,j] * digit_d(w_fix[j,t]),
// replaces chromax TraitModel.__call__, breedgym/breedgym.py:233, vec_env.py:132-134), but the
// 128 x K dosage operand never touches shared memory:
//
//   warp 8  (loader, 2 lanes) : lane 0 streams raw bit-plane tiles [256 plane-rows x 64 B = 4 steps] with
//                               cp.async.bulk.tensor.2d (a tensor map over the packed population),
//                               lane 1 streams the digit tiles with 1-D bulk copies; both land on
//                               mbarriers with expect_tx, nothing sits on a thread's scoreboard.
//   warps 0-7 (expanders)     : two groups of 128 threads take alternate 128-marker steps; thread t
//                               <-> individual t of the tile <-> TMEM lane t.  4 words per plane ->
//                               128 dosage bytes by shift/mask/add -> 4 x tcgen05.st.32x32b.x8
//                               straight into the A stage in tensor memory.
//   warp 9  (MMA, 1 lane)     : 4 x tcgen05.mma.kind::i8 per step with A from TMEM, B (digits) from
//                               shared memory, D in TMEM; tcgen05.commit frees the A/B stage.
//   warps 0-3 (epilogue)      : tcgen05.ld, digits -> int64; K-split partials meet in 64-bit integer
//                               atomics and the last CTA of a tile converts to float32 (no finalize launch).
//
// Per CTA: 54 KB of shared memory and 128 TMEM columns for <= 4 traits => 4 CTAs (40 warps) per SM.
#include <cuda.h>
#include <string.h>

#include "bg_internal.h"

namespace {

constexpr int T2_M = 128;
constexpr int T2_KS = 128;          // markers per step (4 words per plane, 32 TMEM columns of int8x4)
#ifndef T2_S_VAL
#define T2_S_VAL 3
#endif
constexpr int T2_S = T2_S_VAL;      // A (TMEM) / B (smem) stages
#ifndef T2_R_VAL
#define T2_R_VAL 4
#endif
constexpr int T2_R = T2_R_VAL;      // raw macro-tile ring
constexpr int T2_SPM = 1;           // steps per raw macro tile (4 = 64-byte TMA rows measured slower: 38 vs 34.6 us at C2)
constexpr int T2_THREADS = 256 + 64;
constexpr int T2F_THREADS = T2_THREADS + 128;   // fused: + 4 blender warps
constexpr int T2F_MSTEPS = 8;                   // fused: steps per blended macro tile (128 B per plane-row)
constexpr uint32_t T2F_TILE_BYTES = 2 * T2_M * 16 * T2F_MSTEPS;  // 128 rows x 2 planes x 128 B = 32 KB, 2 slots
constexpr uint32_t T2_RAW_ROW = 16 * T2_SPM;                 // bytes per plane-row in a macro tile
constexpr uint32_t T2_RAW_BYTES = 2 * T2_M * T2_RAW_ROW;     // 256 plane-rows x 64 B
constexpr uint32_t T2_SPIN_LIMIT = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"  // sleeps up to %3 ns unless the phase completes
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(2000u)
            : "memory");
        if (spin > T2_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&o)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(o[0]), "r"(o[1]),
                 "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                 : "memory");
}

struct T2Bars {
    uint64_t raw_full[T2_R], raw_empty[T2_R];
    uint64_t a_full[T2_S], a_empty[T2_S], b_full[T2_S];
    uint64_t done;
};

// FUSED variant (vector-env step): cross + GEBV in one pass, the offspring are never read back.
// Four extra "blender" warps replace the TMA loader: 8 lanes cover 128 contiguous bytes of one parent
// plane (fully coalesced 128-bit loads of both parent planes and the shared crossover mask), one LOP3
// per word selects the alleles, the offspring words go to HBM (coalesced) and into a shared-memory
// macro tile [128 rows][2 planes][8 x 16 B] whose 16-byte chunks are XOR-swizzled with the row so that
// both the row-contiguous blender stores and the one-row-per-lane expander loads are bank-conflict free.
struct FusedArgs {
    const uint32_t *pop;      // [E][n_src][2][Wpad]
    const int32_t *parents;   // [E][2n]
    const uint32_t *mask;     // [2n][Wpad]
    uint32_t *out_pop;        // [E][n][2][Wpad]
    int64_t n_src, n;
    int Wpad;
};

__device__ __forceinline__ uint4 blend4(const uint4 h0, const uint4 h1, const uint4 M)
{
    uint4 o;
    o.x = (h0.x & ~M.x) | (h1.x & M.x);
    o.y = (h0.y & ~M.y) | (h1.y & M.y);
    o.z = (h0.z & ~M.z) | (h1.z & M.z);
    o.w = (h0.w & ~M.w) | (h1.w & M.w);
    return o;
}

// smem: raw ring [T2_R][256 plane-rows][64 B], then B stages [T2_S][N/8][8 ki][8][16 B]
template <bool FUSED>
__global__ void __launch_bounds__(FUSED ? T2F_THREADS : T2_THREADS, FUSED ? 3 : 4)
    gebv_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const FusedArgs fa, int64_t rows, const int8_t *__restrict__ bdig,
                    int N, int T, int steps_total, int steps_per_split, unsigned long long *__restrict__ acc,
                    unsigned int *__restrict__ tile_cnt, const double *__restrict__ inv_scale, float *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) T2Bars bars;
    __shared__ uint32_t tmem_base_slot;
    __shared__ uint32_t last_cta_flag;
    __shared__ uint32_t row_src[FUSED ? 2 * T2_M : 1], row_msk[FUSED ? 2 * T2_M : 1];  // per plane-row: uint4 offsets

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw_base = smem_u32(smem);
    const uint32_t b_bytes = (uint32_t)N * T2_KS;
    const uint32_t b_base0 = raw_base + (FUSED ? 2 * T2F_TILE_BYTES : T2_R * T2_RAW_BYTES);
    const int64_t row0 = (int64_t)blockIdx.x * T2_M;
    const int s_begin = blockIdx.y * steps_per_split;
    const int nst = min(steps_total, s_begin + steps_per_split) - s_begin;

    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tmem_cols = 32;
    while (tmem_cols < d_cols + T2_S * (T2_KS / 4)) tmem_cols <<= 1;

    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < T2_R; ++i) {
            // plain: expect_tx arrival of the TMA loader / 4 reader warps per step; fused: 4 blender warps / 8 steps x 4 warps
            mbar_init(smem_u32(&bars.raw_full[i]), FUSED ? 4 : 1);
            mbar_init(smem_u32(&bars.raw_empty[i]), FUSED ? 4 * T2F_MSTEPS : 4 * T2_SPM);
        }
        for (int i = 0; i < T2_S; ++i) {
            mbar_init(smem_u32(&bars.a_full[i]), 4);   // the 4 warps of the group that filled the stage
            mbar_init(smem_u32(&bars.a_empty[i]), 1);  // tcgen05.commit
            mbar_init(smem_u32(&bars.b_full[i]), 1);   // expect_tx arrival of the loader
        }
        mbar_init(smem_u32(&bars.done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (FUSED) {
        // plane-row (t, p) of the tile: source = plane 0 of parent p of individual t (plane 1 follows), mask row 2i+p
        const int W4 = fa.Wpad >> 2;
        for (int prow = tid; prow < 2 * T2_M; prow += blockDim.x) {
            const int64_t gi = row0 + (prow >> 1);
            uint32_t src = 0xFFFFFFFFu, msk = 0;
            if (gi < rows) {
                const int64_t e = gi / fa.n, i = gi % fa.n;
                int64_t a = fa.parents[(e * fa.n + i) * 2 + (prow & 1)];
                a += a < 0 ? fa.n_src : 0;  // jnp indexing: negatives wrap once, then clamp
                a = a < 0 ? 0 : (a > fa.n_src - 1 ? fa.n_src - 1 : a);
                src = (uint32_t)(((e * fa.n_src + a) * 2) * W4);
                msk = (uint32_t)((2 * i + (prow & 1)) * W4);
            }
            row_src[prow] = src;
            row_msk[prow] = msk;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // Programmatic dependent launch: everything above (TMEM allocation, barrier init, row table) may overlap the
    // tail of the kernel that produces this population; its memory is visible only past this point.  A no-op when
    // the launch did not ask for programmatic stream serialization.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t tmem_a = tmem_d + d_cols;

    if (warp < 8) {
        // ---------------- expanders: group g takes steps j with j % 2 == g ----------------
        const int g = warp >> 2, r = tid & (T2_M - 1);
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quadrant
        // per-step tail shared by both variants: dosage bytes of 4 words per plane -> A stage in TMEM
        auto expand_step = [&](int j, const uint4 x0, const uint4 x1) {
            const int as = j % T2_S;
            // dosages as 2-bit fields first (even / odd markers: field f of ze <-> marker 2f, of zo <-> 2f+1; a field
            // holds 0..2, no carry) -- done BEFORE waiting for the TMEM stage, so this work hides the MMA's latency
            const uint32_t w0[4] = {x0.x, x0.y, x0.z, x0.w}, w1[4] = {x1.x, x1.y, x1.z, x1.w};
            uint32_t ze[4], zo[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                ze[jj] = (w0[jj] & 0x55555555u) + (w1[jj] & 0x55555555u);
                zo[jj] = ((w0[jj] >> 1) & 0x55555555u) + ((w1[jj] >> 1) & 0x55555555u);
            }
            if (j >= T2_S) mbar_wait(smem_u32(&bars.a_empty[as]), ((j / T2_S) - 1) & 1);  // MMAs of the previous use retired
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tmem_a + lane_sel + (uint32_t)as * (T2_KS / 4);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                // one shift + mask lifts 4 fields into 4 bytes: column q, byte b <-> K index 4q + b <-> marker 8b + q
                // of this word.  22 integer ops per 32 markers in total.
                uint32_t o[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = (((q & 1) ? zo[jj] : ze[jj]) >> (q & ~1)) & 0x03030303u;
                tmem_st8(ta + 8 * jj, o);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.a_full[as]));
        };
        if (FUSED) {
            for (int j = g; j < nst; j += 2) {
                const int mt = j / T2F_MSTEPS, slot = mt & 1;
                mbar_wait(smem_u32(&bars.raw_full[slot]), (mt >> 1) & 1);
                const uint32_t src = raw_base + slot * T2F_TILE_BYTES + (uint32_t)r * 256 +
                                     (uint32_t)(((j % T2F_MSTEPS) ^ (r & 7)) * 16);  // swizzled 16-byte chunk of this step
                const uint4 x0 = lds128(src), x1 = lds128(src + 128);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars.raw_empty[slot]));
                expand_step(j, x0, x1);
            }
        } else {
            for (int j = g; j < nst; j += 2) {
                const int mt = j / T2_SPM, rs = mt % T2_R;
                mbar_wait(smem_u32(&bars.raw_full[rs]), (mt / T2_R) & 1);
                // macro tile [row][plane][64 B]; this step's 16 B of each plane
                const uint32_t src = raw_base + rs * T2_RAW_BYTES + (uint32_t)r * (2 * T2_RAW_ROW) + (uint32_t)(j % T2_SPM) * 16;
                const uint4 x0 = lds128(src), x1 = lds128(src + T2_RAW_ROW);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars.raw_empty[rs]));
                expand_step(j, x0, x1);
            }
        }
    } else if (warp == 8) {
        if (lane == 0 && !FUSED) {
            // ---------------- raw bit-plane tiles (TMA 2-D) ----------------
            const int y = (int)(2 * row0);  // plane-row coordinate
            const int nmt = (nst + T2_SPM - 1) / T2_SPM;
            for (int mt = 0; mt < nmt; ++mt) {
                const int rs = mt % T2_R;
                if (mt >= T2_R) mbar_wait(smem_u32(&bars.raw_empty[rs]), ((mt / T2_R) - 1) & 1);
                const uint32_t full = smem_u32(&bars.raw_full[rs]);
                mbar_arrive_expect_tx(full, T2_RAW_BYTES);  // out-of-bounds parts of the box are zero-filled and counted
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                        raw_base + rs * T2_RAW_BYTES),
                    "l"(reinterpret_cast<uint64_t>(&tmap)), "r"((s_begin + mt * T2_SPM) * 4), "r"(y), "r"(full)
                    : "memory");
            }
        } else if (lane == 1) {
            // ---------------- digit tiles (1-D bulk copies) ----------------
            for (int j = 0; j < nst; ++j) {
                const int bs = j % T2_S;
                if (j >= T2_S) mbar_wait(smem_u32(&bars.a_empty[bs]), ((j / T2_S) - 1) & 1);
                const uint32_t full = smem_u32(&bars.b_full[bs]);
                mbar_arrive_expect_tx(full, b_bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 b_base0 + bs * b_bytes),
                             "l"(bdig + (int64_t)(s_begin + j) * b_bytes), "r"(b_bytes), "r"(full)
                             : "memory");
            }
        }
    } else if (FUSED && warp >= 10) {
        // ---------------- blenders: parents + masks -> offspring (HBM) + swizzled macro tile (smem) ----------------
        const int W4 = fa.Wpad >> 2;
        const uint4 *pop4 = reinterpret_cast<const uint4 *>(fa.pop);
        const uint4 *mask4 = reinterpret_cast<const uint4 *>(fa.mask);
        uint4 *out4 = reinterpret_cast<uint4 *>(fa.out_pop) + (int64_t)(2 * row0) * W4;
        const int bw = warp - 10, c = lane & 7, rsub = lane >> 3;
        const int nmt = (nst + T2F_MSTEPS - 1) / T2F_MSTEPS;
        for (int mt = 0; mt < nmt; ++mt) {
            const int slot = mt & 1;
            if (mt >= 2) mbar_wait(smem_u32(&bars.raw_empty[slot]), ((mt >> 1) - 1) & 1);
            const int w4 = s_begin + mt * T2F_MSTEPS + c;  // uint4 index inside the row = global step index
            const uint32_t tile = raw_base + slot * T2F_TILE_BYTES;
#pragma unroll 2
            for (int pass = bw; pass < (2 * T2_M) / 4; pass += 4) {
                const int prow = pass * 4 + rsub, t = prow >> 1;
                const uint32_t src = row_src[prow];
                uint4 o = make_uint4(0, 0, 0, 0);
                if (src != 0xFFFFFFFFu && w4 < W4) {
                    const uint4 h0 = __ldg(pop4 + src + w4), h1 = __ldg(pop4 + src + W4 + w4);
                    const uint4 M = __ldg(mask4 + row_msk[prow] + w4);
                    o = blend4(h0, h1, M);
                    out4[(int64_t)prow * W4 + w4] = o;
                }
                const uint32_t dst = tile + (uint32_t)t * 256 + (uint32_t)(prow & 1) * 128 + (uint32_t)((c ^ (t & 7)) * 16);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.raw_full[slot]));
        }
    } else if (warp == 9 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(T2_M >> 4) << 24);
        for (int j = 0; j < nst; ++j) {
            const int as = j % T2_S;
            const uint32_t par = (j / T2_S) & 1;
            mbar_wait(smem_u32(&bars.b_full[as]), par);
            mbar_wait(smem_u32(&bars.a_full[as]), par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kk = 0; kk < T2_KS / 32; ++kk) {
                const uint32_t a_taddr = tmem_a + (uint32_t)as * (T2_KS / 4) + 8 * kk;
                const uint64_t bdesc = make_smem_desc(b_base0 + as * b_bytes + kk * 256, 128, 1024);
                const uint32_t acc = (j > 0 || kk > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                    ::"r"(tmem_d), "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(acc)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars.a_empty[as]))
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars.done))
                     : "memory");
    }

    if (warp < 4) {  // epilogue: thread t <-> accumulator row t <-> TMEM lane t
        mbar_wait(smem_u32(&bars.done), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t row = row0 + tid;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        const bool single = gridDim.y == 1;  // no K split: this CTA holds the whole sum
        for (int t = 0; t < T; ++t) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr + (uint32_t)(8 * t))
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            unsigned long long sum = 0;  // modular arithmetic: the true total fits in int64
#pragma unroll
            for (int d = 7; d >= 0; --d) sum = (sum << 8) + (unsigned long long)(long long)(int32_t)v[d];
            if (row < rows) {
                if (single)
                    out[row * T + t] = (float)((double)(long long)sum * inv_scale[t]);
                else
                    atomicAdd(acc + row * T + t, sum);  // integer partial sums: order independent
            }
        }
        if (!single) {
            // the LAST K-split CTA of this tile converts and re-zeroes the accumulators (they are all
            // zero between launches), so no finalize kernel is needed
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) last_cta_flag = atomicAdd(tile_cnt + blockIdx.x, 1u) == gridDim.y - 1 ? 1u : 0u;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (last_cta_flag) {
                __threadfence();
                if (row < rows)
                    for (int t = 0; t < T; ++t) {
                        const unsigned long long tot = atomicExch(acc + row * T + t, 0ull);
                        out[row * T + t] = (float)((double)(long long)tot * inv_scale[t]);
                    }
                if (tid == 0) tile_cnt[blockIdx.x] = 0u;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

// shared launcher: fa == nullptr -> GEBV of the finished population `pop`; else fused cross + GEBV
static int launch_tc2(bg_engine *eng, const uint32_t *pop, const FusedArgs *fa, int64_t rows, float *out, cudaStream_t st,
                      int scratch = 0)
{
    BG_REQUIRE(eng && eng->d_wdig, BG_ESTATE, "engine has no tensor-core digit table");
    const int T = eng->T, N = eng->tc_N;
    const int steps = (int)eng->tc_steps;
    const int64_t tiles = (rows + T2_M - 1) / T2_M;
    BG_REQUIRE(tiles < (int64_t(1) << 31) && 2 * rows < (int64_t(1) << 31), BG_ELIMIT, "too many rows");

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (!fa) {
        EncodeTiledFn enc = encode_tiled();
        BG_REQUIRE(enc, BG_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        // 2-D view of the packed population: [2*rows plane-rows][Wpad words]; box = 16 words x 256 plane-rows
        const cuuint64_t gdim[2] = {(cuuint64_t)eng->Wpad, (cuuint64_t)(2 * rows)};
        const cuuint64_t gstride[1] = {(cuuint64_t)eng->Wpad * 4};
        const cuuint32_t box[2] = {4 * T2_SPM, 2 * T2_M};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t *>(pop), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        BG_REQUIRE(cr == CUDA_SUCCESS, BG_ECUDA, "cuTensorMapEncodeTiled failed");
    }

    const size_t raw_bytes = fa ? (size_t)2 * T2F_TILE_BYTES : (size_t)T2_R * T2_RAW_BYTES;
    const size_t smem = raw_bytes + (size_t)T2_S * N * T2_KS;
    BG_REQUIRE(smem <= (size_t)eng->max_smem_optin, BG_ELIMIT, "too many traits for the tensor-core GEBV tile");
    if (fa)
        BG_REQUIRE((int64_t)eng->Wpad / 4 * 2 * (fa->n_src > fa->n ? fa->n_src : fa->n) * ((rows + fa->n - 1) / fa->n) < (int64_t(1) << 32),
                   BG_ELIMIT, "population too large for the fused kernel's 32-bit row offsets");
    // residency: TMEM columns (512 per SM), shared memory, and registers for the fused variant
    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tcols = 32;
    while (tcols < d_cols + T2_S * (T2_KS / 4)) tcols <<= 1;
    int resident = (int)(512 / tcols);
    const int by_smem = (int)(227 * 1024 / (smem + 1024));
    if (by_smem < resident) resident = by_smem;
    if (resident > 6) resident = 6;  // 10 warps per CTA, 64 per SM
    if (fa && resident > 3) resident = 3;
    if (resident < 1) resident = 1;
    int64_t target = (int64_t)resident * eng->sm_count;  // one full wave
    if (const char *s = getenv("BG_TC_TARGET_CTAS")) target = atoll(s) > 0 ? atoll(s) : target;
    int ksplit = (int)(target / tiles);
    const int max_split = (steps + 7) / 8;
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 65535) ksplit = 65535;
    const int sps = (steps + ksplit - 1) / ksplit;
    ksplit = (steps + sps - 1) / sps;
    const int64_t total = rows * T;
    // zero-invariant scratch: accumulators [rows][T] and one arrival counter per tile
    if (eng->acc2_cap[scratch] < (size_t)total) {
        if (eng->d_acc2[scratch]) BG_CUDA(cudaFree(eng->d_acc2[scratch]));
        eng->d_acc2[scratch] = nullptr;
        eng->acc2_cap[scratch] = 0;
        BG_CUDA(cudaMalloc(&eng->d_acc2[scratch], (size_t)total * sizeof(unsigned long long)));
        BG_CUDA(cudaMemsetAsync(eng->d_acc2[scratch], 0, (size_t)total * sizeof(unsigned long long), st));
        eng->acc2_cap[scratch] = (size_t)total;
    }
    if (eng->tile_cap[scratch] < (size_t)tiles) {
        if (eng->d_tile_cnt[scratch]) BG_CUDA(cudaFree(eng->d_tile_cnt[scratch]));
        eng->d_tile_cnt[scratch] = nullptr;
        eng->tile_cap[scratch] = 0;
        BG_CUDA(cudaMalloc(&eng->d_tile_cnt[scratch], (size_t)tiles * sizeof(unsigned int)));
        BG_CUDA(cudaMemsetAsync(eng->d_tile_cnt[scratch], 0, (size_t)tiles * sizeof(unsigned int), st));
        eng->tile_cap[scratch] = (size_t)tiles;
    }
    dim3 grid((unsigned)tiles, (unsigned)ksplit);
    // largest dynamic smem opted into so far, per engine (= per device: the attribute is per device)
    if (fa) {
        if (smem > eng->tc2_optin[1]) {
            BG_CUDA(cudaFuncSetAttribute(gebv_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            eng->tc2_optin[1] = smem;
        }
        gebv_tc2_kernel<true><<<grid, T2F_THREADS, smem, st>>>(tmap, *fa, rows, eng->d_wdig, N, T, steps, sps, eng->d_acc2[scratch],
                                                              eng->d_tile_cnt[scratch], eng->d_inv_scale, out);
    } else {
        if (smem > eng->tc2_optin[0]) {
            BG_CUDA(cudaFuncSetAttribute(gebv_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            eng->tc2_optin[0] = smem;
        }
        FusedArgs none;
        memset(&none, 0, sizeof(none));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = grid;
        cfg.blockDim = dim3(T2_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const int8_t *bd = eng->d_wdig;
        const double *inv = eng->d_inv_scale;
        unsigned long long *acc = eng->d_acc2[scratch];
        unsigned int *cnt = eng->d_tile_cnt[scratch];
        BG_CUDA(cudaLaunchKernelEx(&cfg, gebv_tc2_kernel<false>, tmap, none, rows, bd, N, T, steps, sps, acc, cnt, inv, out));
    }
    BG_LAUNCHED();
    return BG_OK;
}

// scratch: which of the two zero-invariant accumulator sets to use (two launches may be in flight on two streams)
int bg_launch_gebv_tc2(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, cudaStream_t st, int scratch)
{
    return launch_tc2(eng, pop, nullptr, rows, out, st, scratch);
}

// vector-env step: out_pop[e][i] = cross of pop[e][parents[e][i][0..1]] under mask[2i..2i+1]; gebv[e][i][T]
int bg_launch_cross_gebv_fused(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask, uint32_t *out_pop,
                               int64_t E, int64_t n_src, int64_t n, float *gebv_out, cudaStream_t st)
{
    FusedArgs fa;
    fa.pop = pop;
    fa.parents = parents;
    fa.mask = mask;
    fa.out_pop = out_pop;
    fa.n_src = n_src;
    fa.n = n;
    fa.Wpad = eng->Wpad;
    return launch_tc2(eng, nullptr, &fa, E * n, gebv_out, st);
}

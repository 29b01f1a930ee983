// Vector-env step in ONE kernel: cross (gather parents, blend under the shared crossover masks) + GEBV.
//
// Replaces breedgym/vector/vec_env.py:89-90 (`populations[arange, actions]`), :77 (`vmap(simulator.cross)`)
// and :132-134 (`GEBV_model(populations)`): the offspring are written to HBM once and never read back, the
// 2x-population parent gather is never materialised.  Algorithmic HBM traffic: 0.75 B per offspring-marker
//  -- the separate blend + GEBV pass pays 1.0 B.
//
// One CTA = 128 offspring x a K range of 128-marker steps (same K-split / exact int8 digit GEMM as gebv_tc2.cu).
// The 128 offspring of a tile are taken CHILD-major: tile row R <-> (child i = R / E, env e = R % E).  The crossover
// masks depend on the child slot only (the reference shares one key across envs, vec_env.py:75-77), so with E a
// multiple of 32 all lanes of a warp read the SAME mask words: one broadcast L1/L2 access per warp instead of a
// third of the kernel's gather traffic.  Warp roles:
//
//   warps 10-13 (loaders)   : cp.async (LDGSTS) 16-byte copies, four lanes per 64-byte row segment: the two bit
//                             planes of parent A and of parent B of every offspring -> a ring of stages
//                             [4 chunks][128 rows][64 B = 4 steps] in shared memory (16-byte quarters XOR-swizzled
//                             with the row, so the one-row-per-lane reads below are conflict free).  Nothing
//                             waits on a scoreboard: completion lands on an mbarrier
//                             (cp.async.mbarrier.arrive.noinc), so the bytes in flight are bounded by the ring
//                             (3 x 32 KB per CTA, 2 CTAs per SM), not by registers.  Threads 0-63 also stage the
//                             mask rows of the tile's (<= 8) children.  Before a ring slot is gathered into again,
//                             the same warps drain the offspring words the expanders left in it to HBM, four lanes
//                             per 64-byte row segment (coalesced 128-bit stores).
//   warps 0-7  (expanders)  : thread t <-> offspring t of the tile <-> TMEM lane t; the two groups of 4 warps take
//                             alternate steps.  4 x ld.shared.v4 + the two mask quads (broadcast ld.shared; with
//                             few envs per-lane ld.global.nc, prefetched a step ahead), one LOP3 per word selects
//                             the alleles (h0 & ~M | h1 & M), the offspring words replace parent A's IN PLACE in
//                             the stage (the loader warps store them from there), then 4 words per plane ->
//                             128 prescaled dosage bytes -> tcgen05.st into the A stage in tensor memory.
//   warp 8     (digits)     : 1-D bulk copies (TMA) of the digit tiles, one per pair of steps, after an L2 prefetch
//                             of the CTA's whole digit range.
//   warp 9     (MMA)        : tcgen05.mma.kind::i8, A from TMEM, B from shared memory, D in TMEM; the whole warp runs
//                             the loop and one elected lane issues (elect.sync), one barrier round and one
//                             tcgen05.commit per PAIR of steps.
//   warps 0-3  (epilogue)   : digits -> int64 -> one 64-bit atomic per value carrying the K-split partial sum and
//                             the arrival count -> float32 by the last arrival (tc_common.cuh).
//
// The kernel is capped at 64 registers per thread (launch bounds of 512 threads, 448 launched): two CTAs then leave
// 8192 registers of the SM free, exactly one 128-thread CTA of the mask kernel of the NEXT steps (meiosis.cu, side
// stream), which runs in the issue slots this latency-bound kernel leaves idle.  Measured history and dead ends:
// DESIGN.md section 4; `-DXG_TRACE=1` + scripts/fused_trace.py print one CTA's pipeline timeline.
#include <cuda.h>
#include <string.h>

#include "bg_internal.h"
#include "tc_common.cuh"

using namespace bgtc;

namespace {

#ifndef XG_R_VAL
#define XG_R_VAL 3
#endif
#ifndef XG_S_VAL
#define XG_S_VAL 6
#endif
#ifndef XG_CTAS_VAL
#define XG_CTAS_VAL 2
#endif
#ifndef XG_L2HINT_VAL
#define XG_L2HINT_VAL 0     // L2 prefetch size hint of the cp.async gathers (0 / 128 / 256 bytes): no effect measured at C2
#endif
#ifndef XG_DEBUG_SKIP
#define XG_DEBUG_SKIP 0     // timing experiments only (results are wrong): 2 no offspring stores, 4 no gathers
#endif
constexpr int XG_R = XG_R_VAL;        // stage ring (stages of 4 steps), used IN PLACE: gathered parents -> offspring
constexpr int XG_S = XG_S_VAL;        // A stages in tensor memory (steps); handed over in PAIRS of steps
constexpr int XG_SP = XG_S / 2;       // pair stages: one barrier round and one tcgen05.commit per two steps (the MMA warp's
                                      // fixed costs -- mbarrier wait, commit -- were the pipeline's bottleneck per step)
constexpr int XG_BP_MAX = 8;          // digit ring: up to 8 pairs of steps ahead
constexpr int XG_CTAS = XG_CTAS_VAL;  // CTAs per SM
constexpr int XG_SPS = 4;             // steps per stage: 64 B per row and plane
constexpr int XG_MC = 8;              // children per tile whose mask rows are staged in shared memory
constexpr int XG_LOADER_WARP0 = 10, XG_LOADERS = 128;
constexpr int XG_THREADS = (XG_LOADER_WARP0 + 4) * 32;
#ifndef XG_REGCAP_THREADS
#define XG_REGCAP_THREADS 512
#endif
constexpr uint32_t XG_ROW = 16 * XG_SPS;                // bytes per row and plane in a stage
constexpr uint32_t XG_CHUNK = TILE_M * XG_ROW;          // one plane of a stage: 128 rows x 64 B
constexpr uint32_t XG_IN_BYTES = 4 * XG_CHUNK;          // parent A planes 0/1 (-> offspring planes 0/1), parent B planes 0/1
constexpr uint32_t XG_MASK_BYTES = XG_MC * 2 * XG_ROW;  // mask rows of up to 8 children
constexpr uint32_t XG_NOROW = 0xFFFFFFFFu;

#ifndef XG_TRACE
#define XG_TRACE 0   // 1: CTA (XG_TRACE_CTA, 0) records clock64() stamps of its pipeline events (diagnostics build only)
#endif
#if XG_TRACE
#ifndef XG_TRACE_CTA
#define XG_TRACE_CTA 0
#endif
#ifndef XG_TRACE_Y
#define XG_TRACE_Y 0
#endif
__device__ long long xg_trace_buf[16 * 64];
#define XG_STAMP(slot, idx)                                                                                              \
    do {                                                                                                                 \
        if (blockIdx.x == XG_TRACE_CTA && blockIdx.y == XG_TRACE_Y && (idx) < 64) xg_trace_buf[(slot) * 64 + (idx)] = clock64(); \
    } while (0)
#else
#define XG_STAMP(slot, idx) \
    do {                    \
    } while (0)
#endif

struct XGBars {
    uint64_t raw_full[XG_R], stage_done[XG_R];
    uint64_t a_full[XG_SP], a_empty[XG_SP];
    uint64_t b_full[XG_BP_MAX], b_empty[XG_BP_MAX];
    uint64_t done;
};

struct XGArgs {
    const uint4 *pop;         // [E][n_src][2][W4]
    const int32_t *parents;   // [E][n][2]
    const uint4 *mask;        // [2n][W4]
    uint4 *out_pop;           // [E][n][2][W4]
    int64_t n_src, n, E, rows;  // rows = E * n
    int W4;
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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes)
{
#if XG_L2HINT_VAL == 256
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#elif XG_L2HINT_VAL == 128
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#endif
}

// byte offset of 16-byte quarter q of row t inside a [128 rows][64 B] chunk: quarters XOR-swizzled with the row, so
// that 8 consecutive rows reading the same quarter hit 8 different 16-byte bank groups
__device__ __forceinline__ uint32_t swz(int t, int q) { return (uint32_t)t * XG_ROW + (uint32_t)((q ^ ((t >> 1) & 3)) * 16); }

// smem: stage ring [XG_R][4][128][64 B], mask ring [XG_R][8 children][2][64 B], digit ring [nbp pairs][2 steps][N/8][8 ki][8][16 B]
// launch bounds of 512 threads (448 are launched): caps the kernel at 64 registers per thread, so that two CTAs leave
// 8192 registers of the SM free -- exactly one 128-thread CTA of the mask kernel, which then runs in the issue slots
// this (latency-bound) kernel leaves idle instead of displacing its CTAs
__global__ void __launch_bounds__(XG_REGCAP_THREADS, XG_CTAS)
    cross_gebv_kernel(const XGArgs fa, const int8_t *__restrict__ bdig, int N, int T, int D, int nbp, int steps_total,
                      int steps_per_split, unsigned long long *__restrict__ acc,
                      const double *__restrict__ inv_scale, float *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) XGBars bars;
    __shared__ uint32_t tmem_base_slot;
    // per offspring t of the tile: uint4 offsets of its parents' rows (plane 0; plane 1 follows) and mask rows, and
    // its row index in out_pop / gebv (XG_NOROW past the end)
    __shared__ uint32_t row_src[2 * TILE_M], row_msk[2 * TILE_M], row_out[TILE_M];

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) XG_STAMP(15, 0);
    // Programmatic dependent launch, both ends.  (1) The NEXT step kernel of the stream may be scheduled as soon as every
    // CTA of this grid has started: its CTAs take the slots this grid's last wave leaves idle and run their prologue
    // (tensor-memory allocation, barrier init, row table from the action array) while this grid drains.  (2) This
    // kernel's own prologue below touches nothing an earlier kernel of the stream writes -- the action array is an
    // INPUT of the step, produced before the previous step kernel was launched or by a kernel that does not trigger
    // early -- and everything after `griddepcontrol.wait` (populations, masks, accumulators, outputs) sees the
    // completed, flushed predecessor.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t in_base = smem_u32(smem);
    const uint32_t mask_base = in_base + XG_R * XG_IN_BYTES;
    const uint32_t b_base0 = mask_base + XG_R * XG_MASK_BYTES;
    const uint32_t b_bytes = (uint32_t)N * STEP_K;
    const int64_t row0 = (int64_t)blockIdx.x * TILE_M;
    const int s_begin = blockIdx.y * steps_per_split;                              // a multiple of XG_SPS (launcher)
    const int nst = min(steps_total, s_begin + steps_per_split) - s_begin;          // a multiple of XG_SPS too
    const int nstages = nst / XG_SPS;
    // child-major tile rows: R <-> (child R / E, env R % E); the tile spans children i0 .. i0 + nchild - 1
    const int64_t last_row = min(fa.rows, row0 + TILE_M) - 1;
    const int64_t i0 = row0 / fa.E;
    const int nchild = (int)(last_row / fa.E - i0) + 1;
    const bool mask_smem = nchild <= XG_MC;  // else (few envs): every thread fetches its own mask words from L2

    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tmem_cols = 32;
    while (tmem_cols < d_cols + XG_S * (STEP_K / 4)) tmem_cols <<= 1;

    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 8 * 32) {  // a thread without a row-table entry: the table build below is the prologue's critical path
        for (int i = 0; i < XG_R; ++i) {
            mbar_init(smem_u32(&bars.raw_full[i]), XG_LOADERS);  // one cp.async completion arrival per loader thread
            mbar_init(smem_u32(&bars.stage_done[i]), 8);         // the 8 expander warps: offspring words are in place
        }
        for (int i = 0; i < XG_SP; ++i) {
            mbar_init(smem_u32(&bars.a_full[i]), 8);   // the 8 expander warps: both steps of the pair are in tensor memory
            mbar_init(smem_u32(&bars.a_empty[i]), 1);  // tcgen05.commit
        }
        for (int i = 0; i < XG_BP_MAX; ++i) {
            mbar_init(smem_u32(&bars.b_full[i]), 1);   // expect_tx arrival of the digit loader
            mbar_init(smem_u32(&bars.b_empty[i]), 1);  // tcgen05.commit
        }
        mbar_init(smem_u32(&bars.done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int k = tid; k < 2 * TILE_M; k += blockDim.x) {
        const int t = k >> 1, p = k & 1;
        const int64_t R = row0 + t;
        uint32_t src = XG_NOROW, msk = 0, orow = XG_NOROW;
        if (R < fa.rows) {
            const uint32_t i32 = (uint32_t)R / (uint32_t)fa.E;  // rows < 2^31 (launcher): 32-bit division
            const int64_t i = i32, e = (uint32_t)R - i32 * (uint32_t)fa.E;
            orow = (uint32_t)(e * fa.n + i);
            int64_t a = fa.parents[(int64_t)orow * 2 + p];
            a += a < 0 ? fa.n_src : 0;  // jnp indexing: negatives wrap once, then clamp
            a = a < 0 ? 0 : (a > fa.n_src - 1 ? fa.n_src - 1 : a);
            src = (uint32_t)(((e * fa.n_src + a) * 2) * fa.W4);
            // mask row 2i + p: its offset in the global mask array, or (smem mode) its byte offset in a mask stage
            msk = mask_smem ? (uint32_t)((2 * (i - i0) + p) * XG_ROW) : (uint32_t)((2 * i + p) * fa.W4);
        }
        row_src[k] = src;
        row_msk[k] = msk;
        if (p == 0) row_out[t] = orow;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // a no-op unless launched with programmatic stream serialization
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t tmem_a = tmem_d + d_cols;

    if (warp < 8) {
        // ---------------- expanders: group g takes the steps j with j % 2 == g ----------------
        const int g = warp >> 2, r = tid & (TILE_M - 1);
        const bool stamp = (warp & 3) == 0 && lane == 0;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quadrant
        const uint32_t msk_a = row_msk[2 * r], msk_b = row_msk[2 * r + 1];
        const uint4 *mrow_a = fa.mask + (mask_smem ? 0u : msk_a) + s_begin, *mrow_b = fa.mask + (mask_smem ? 0u : msk_b) + s_begin;
        uint4 ma_next = make_uint4(0, 0, 0, 0), mb_next = ma_next;
        if (!mask_smem) {
            ma_next = __ldg(mrow_a + g);
            mb_next = __ldg(mrow_b + g);
        }
        int ap = 0;  // pair stage of step j = (j / 2) % XG_SP; group g fills half g of it
        uint32_t a_use = 0;
        for (int st = 0; st < nstages; ++st) {
            const int rs = st % XG_R;
            mbar_wait(smem_u32(&bars.raw_full[rs]), (st / XG_R) & 1);
            if (stamp) XG_STAMP(2 + g, st);  // raw_full seen
            const uint32_t stage = in_base + rs * XG_IN_BYTES, mstage = mask_base + rs * XG_MASK_BYTES;
#pragma unroll
            for (int k = 0; k < XG_SPS / 2; ++k) {
                const int q = 2 * k + g, j = XG_SPS * st + q;  // quarter of the stage row, step of this CTA
                const uint32_t off = swz(r, q);
                const uint4 a0 = lds128(stage + off), a1 = lds128(stage + XG_CHUNK + off);
                const uint4 b0 = lds128(stage + 2 * XG_CHUNK + off), b1 = lds128(stage + 3 * XG_CHUNK + off);
                uint4 ma, mb;
                if (mask_smem) {  // all lanes of a warp share a child when E % 32 == 0: broadcast reads
                    ma = lds128(mstage + msk_a + q * 16);
                    mb = lds128(mstage + msk_b + q * 16);
                } else {
                    ma = ma_next;
                    mb = mb_next;
                    if (j + 2 < nst) {  // masks of this thread's next step
                        ma_next = __ldg(mrow_a + j + 2);
                        mb_next = __ldg(mrow_b + j + 2);
                    }
                }
                const uint4 x0 = blend4(a0, a1, ma), x1 = blend4(b0, b1, mb);
                // offspring words replace parent A's in the stage (same thread, same slots): the loader warps store them
                sts128(stage + off, x0);
                sts128(stage + XG_CHUNK + off, x1);
                if (k == XG_SPS / 2 - 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars.stage_done[rs]));
                }
                // dosage bytes -> tensor memory
                const DosageFields f = dosage_fields(x0, x1);
                if (stamp) XG_STAMP(6, j);  // fields done, about to wait a_empty
                if (a_use > 0) mbar_wait(smem_u32(&bars.a_empty[ap]), (a_use - 1) & 1);  // MMAs of the previous use retired
                if (stamp) XG_STAMP(7, j);  // a_empty seen
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                dosage_to_tmem(tmem_a + lane_sel + (uint32_t)(2 * ap + g) * (STEP_K / 4), f);
                if (stamp) XG_STAMP(8, j);  // TMEM stores complete
                if (lane == 0 && g == 0) XG_STAMP(warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 2 ? 4 : 9)), j >> 1);  // per warp of group 0
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars.a_full[ap]));
                if (++ap == XG_SP) {
                    ap = 0;
                    ++a_use;
                }
            }
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ---------------- digit tiles: one bulk copy (TMA) per pair of steps, up to nbp pairs ahead ----------------
            const uint32_t pair_bytes = 2 * b_bytes;
            // the whole digit range of this CTA -> L2 now (the table is usually cold: a step streams more than L2 holds),
            // so that the ring below is fed at L2 latency
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(bdig + (int64_t)s_begin * b_bytes), "r"((uint32_t)nst * b_bytes)
                         : "memory");
            int slot = 0;
            uint32_t use = 0;
            for (int pj = 0; pj < nst / 2; ++pj) {
                if (use > 0) mbar_wait(smem_u32(&bars.b_empty[slot]), (use - 1) & 1);
                const uint32_t full = smem_u32(&bars.b_full[slot]);
                mbar_arrive_expect_tx(full, pair_bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 b_base0 + slot * pair_bytes),
                             "l"(bdig + ((int64_t)s_begin + 2 * pj) * b_bytes), "r"(pair_bytes), "r"(full)
                             : "memory");
                if (++slot == nbp) {
                    slot = 0;
                    ++use;
                }
            }
        }
    } else if (warp == 9) {
        // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues; one round per PAIR of steps ----------------
        const uint32_t idesc = idesc_u8s8(N);
        int ap = 0, slot = 0;
        uint32_t a_par = 0, b_par = 0;
        for (int pj = 0; pj < nst / 2; ++pj) {
            const uint32_t a_taddr = tmem_a + (uint32_t)(2 * ap) * (STEP_K / 4);
            const uint64_t bdesc = make_smem_desc(b_base0 + (uint32_t)slot * 2 * b_bytes, 128, 1024);
            mbar_wait(smem_u32(&bars.b_full[slot]), b_par);
            if (lane == 0) XG_STAMP(10, pj);  // digit pair landed
            mbar_wait(smem_u32(&bars.a_full[ap]), a_par);
            if (lane == 0) XG_STAMP(11, pj);  // A pair full
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int kk = 0; kk < STEP_K / 32; ++kk)  // +16 in the descriptor's address field = +256 bytes
                    mma_i8_ts_warp(tmem_d, a_taddr + h * (STEP_K / 4) + 8 * kk, bdesc + (uint64_t)(h * (b_bytes >> 4)) + 16 * kk, idesc,
                                   (pj > 0 || h > 0 || kk > 0) ? 1u : 0u);
            if (lane == 0) XG_STAMP(5, pj);  // 8 MMAs issued
            mma_commit_warp(smem_u32(&bars.a_empty[ap]));
            mma_commit_warp(smem_u32(&bars.b_empty[slot]));
            if (lane == 0) XG_STAMP(12, pj);  // commits issued
            if (++ap == XG_SP) {
                ap = 0;
                a_par ^= 1;
            }
            if (++slot == nbp) {
                slot = 0;
                b_par ^= 1;
            }
        }
        mma_commit_warp(smem_u32(&bars.done));
    } else {
        // ---------------- loaders / storers: thread u covers quarter q = u & 3 of rows (u >> 2) + 32k ----------------
        // per iteration: drain the offspring words of the stage that used this ring slot XG_R stages ago (coalesced
        // 128-bit stores), then gather the next stage into it (both planes of both parents: 16 cp.async in flight per
        // thread, no register staging) plus, threads u < 64, one 16-byte piece of the tile's mask rows
        const int u = tid - XG_LOADER_WARP0 * 32, q = u & 3;
        uint32_t src[4][2], dst[4], orow[4];
        uint32_t valid = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int t = (u >> 2) + 32 * k;
            src[k][0] = row_src[2 * t];
            src[k][1] = row_src[2 * t + 1];
            orow[k] = row_out[t];
            if (src[k][0] != XG_NOROW) valid |= 1u << k;
            else src[k][0] = src[k][1] = 0;
            dst[k] = swz(t, q);
        }
        // mask piece of this thread: row (2 * (i0 + u / 8) + (u / 4) % 2), quarter q
        const int mrow = u >> 2;  // 0 .. 15 for u < 64
        const bool mask_loader = mask_smem && u < 2 * XG_MC * 4;
        const bool mask_valid = mask_loader && (mrow >> 1) < nchild;
        const uint4 *msrc = fa.mask + (mask_valid ? (2 * i0 + mrow) * fa.W4 : 0);
        for (int it = 0; it < nstages + XG_R; ++it) {
            const int rs = it % XG_R;
            const uint32_t stage = in_base + rs * XG_IN_BYTES;
            if (it >= XG_R) {
                const int so = it - XG_R;
                mbar_wait(smem_u32(&bars.stage_done[rs]), (so / XG_R) & 1);
                if (u == 0) XG_STAMP(14, so);  // stage_done seen
                const int w4 = s_begin + XG_SPS * so + q;
                uint4 v[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[k][0] = lds128(stage + dst[k]);
                    v[k][1] = lds128(stage + XG_CHUNK + dst[k]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (orow[k] != XG_NOROW && !(XG_DEBUG_SKIP & 2)) {
                        uint4 *o = fa.out_pop + (int64_t)orow[k] * 2 * fa.W4 + w4;
                        o[0] = v[k][0];
                        o[fa.W4] = v[k][1];
                    }
            }
            if (it < nstages) {
                // (the slots gathered into below were read by THIS thread's stores above, or by the expanders that
                //  signalled stage_done: no other thread still needs them)
                if (u == 0) XG_STAMP(13, it);  // gathers of stage `it` issued
                const int w4 = s_begin + XG_SPS * it + q;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t nbytes = (((valid >> k) & 1) && !(XG_DEBUG_SKIP & 4)) ? 16u : 0u;  // 0: zero-fill, nothing is read
                    const uint32_t d = stage + dst[k];
                    cp_async16(d, fa.pop + src[k][0] + w4, nbytes);
                    cp_async16(d + XG_CHUNK, fa.pop + src[k][0] + fa.W4 + w4, nbytes);
                    cp_async16(d + 2 * XG_CHUNK, fa.pop + src[k][1] + w4, nbytes);
                    cp_async16(d + 3 * XG_CHUNK, fa.pop + src[k][1] + fa.W4 + w4, nbytes);
                }
                if (mask_loader) cp_async16(mask_base + rs * XG_MASK_BYTES + (uint32_t)mrow * XG_ROW + q * 16, msrc + w4, mask_valid ? 16u : 0u);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars.raw_full[rs])) : "memory");
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }

    if (tid == 0) XG_STAMP(15, 1);
    if (warp < 4) {
        mbar_wait(smem_u32(&bars.done), 0);
        if (tid == 0) XG_STAMP(15, 2);
        digits_epilogue(tmem_d, tid, warp, row0, fa.rows, T, acc, inv_scale, out, gridDim.y, D, row_out[tid] == XG_NOROW ? -1 : (int64_t)row_out[tid]);
    }
    if (tid == 0) XG_STAMP(15, 3);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

}  // namespace

#if XG_TRACE
extern "C" int bg_debug_read_trace(long long *host, int n)
{
    return (int)cudaMemcpyFromSymbol(host, xg_trace_buf, sizeof(long long) * n);
}
#endif

int bg_tc_reserve_scratch(bg_engine *eng, int scratch, int64_t total, int64_t tiles, cudaStream_t st);
void bg_tc_split(int64_t tiles, int steps, int64_t target, int multiple, int *ksplit_out, int *sps_out);

// can the fused kernel take this engine's trait count and this population size?  (else: blend + GEBV kernels)
bool bg_cross_gebv_fused_ok(const bg_engine *eng, int64_t E, int64_t n_src, int64_t n)
{
    if (!eng || !eng->d_wdig || eng->mut_thr) return false;
    const int N = eng->tc_N;
    const size_t smem = (size_t)XG_R * (XG_IN_BYTES + XG_MASK_BYTES) + (size_t)2 * 2 * N * STEP_K;
    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    if (smem + 4096 > (size_t)eng->max_smem_optin || d_cols + XG_S * (STEP_K / 4) > 512) return false;  // + the static arrays
    if (eng->tc_steps % XG_SPS != 0 || E * n >= (int64_t(1) << 31)) return false;
    return (int64_t)eng->Wpad / 4 * 2 * (n_src > n ? n_src : n) * E < (int64_t(1) << 32);
}

// vector-env step: out_pop[e][i] = cross of pop[e][parents[e][i][0..1]] under mask[2i..2i+1]; gebv[e][i][T]
int bg_launch_cross_gebv_fused(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask, uint32_t *out_pop,
                               int64_t E, int64_t n_src, int64_t n, float *gebv_out, cudaStream_t st)
{
    BG_REQUIRE(eng && eng->d_wdig, BG_ESTATE, "engine has no tensor-core digit table");
    static_assert(XG_S % 2 == 0, "the two expander groups alternate over an even number of A stages");
    const int T = eng->T, N = eng->tc_N;
    const int steps = (int)eng->tc_steps;  // a multiple of 8: rows are padded to 32 words
    const int64_t rows = E * n;
    const int64_t tiles = (rows + TILE_M - 1) / TILE_M;
    BG_REQUIRE(steps % XG_SPS == 0, BG_ESTATE, "row pitch is not a multiple of the fused kernel's stage");
    BG_REQUIRE(tiles < (int64_t(1) << 31) && rows < (int64_t(1) << 31), BG_ELIMIT, "too many rows");
    BG_REQUIRE((int64_t)eng->Wpad / 4 * 2 * (n_src > n ? n_src : n) * E < (int64_t(1) << 32), BG_ELIMIT,
               "population too large for the fused kernel's 32-bit row offsets");

    // digit ring: nbp pairs of steps, as deep as fits beside the stage ring with XG_CTAS CTAs per SM
    const size_t rings = (size_t)XG_R * (XG_IN_BYTES + XG_MASK_BYTES), b_bytes = (size_t)N * STEP_K;
    const size_t per_cta = 228 * 1024 / XG_CTAS - 1024 - 3584;  // minus the reserved KB and the static arrays
    int nbp = per_cta > rings ? (int)((per_cta - rings) / (2 * b_bytes)) : 0;
    if (nbp > XG_BP_MAX) nbp = XG_BP_MAX;
    if (nbp < 2) nbp = 2;
    const size_t smem = rings + (size_t)nbp * 2 * b_bytes;
    BG_REQUIRE(smem + 4096 <= (size_t)eng->max_smem_optin, BG_ELIMIT, "too many traits for the fused cross+GEBV tile");
    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tcols = 32;
    while (tcols < d_cols + XG_S * (STEP_K / 4)) tcols <<= 1;
    BG_REQUIRE(tcols <= 512, BG_ELIMIT, "too many traits for the fused kernel's tensor memory");
    int resident = (int)(512 / tcols);
    const int by_smem = (int)(228 * 1024 / (smem + 1024 + 3584));
    if (by_smem < resident) resident = by_smem;
    if (resident > XG_CTAS) resident = XG_CTAS;
    if (resident < 1) resident = 1;
    int ksplit, sps;
    bg_tc_split(tiles, steps, eng->opt.tc_target_ctas > 0 ? eng->opt.tc_target_ctas : 2LL * resident * eng->sm_count /* two waves */, XG_SPS,
                &ksplit, &sps);
    int rc = bg_tc_reserve_scratch(eng, 0, rows * T, tiles, st);
    if (rc) return rc;
    if (smem > eng->tc2_optin[1]) {
        BG_CUDA(cudaFuncSetAttribute(cross_gebv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eng->tc2_optin[1] = smem;
    }
    XGArgs fa;
    fa.pop = reinterpret_cast<const uint4 *>(pop);
    fa.parents = parents;
    fa.mask = reinterpret_cast<const uint4 *>(mask);
    fa.out_pop = reinterpret_cast<uint4 *>(out_pop);
    fa.n_src = n_src;
    fa.n = n;
    fa.E = E;
    fa.rows = rows;
    fa.W4 = eng->Wpad / 4;
    dim3 grid((unsigned)tiles, (unsigned)ksplit);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(XG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = eng->opt.step_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int8_t *bd = eng->d_wdig;
    const int D = eng->tc_D;
    unsigned long long *accp = eng->d_acc2[0];
    const double *inv = eng->d_inv_scale;
    BG_CUDA(cudaLaunchKernelEx(&cfg, cross_gebv_kernel, fa, bd, N, T, D, nbp, steps, sps, accp, inv, gebv_out));
    BG_LAUNCHED();
    return BG_OK;
}

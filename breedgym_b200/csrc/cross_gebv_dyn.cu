// Vector-env step in ONE kernel, PERSISTENT with a dynamic work queue: cross (gather parents, blend under the shared
// crossover masks) + GEBV.
//
// Replaces breedgym/vector/vec_env.py:89-90 (`populations[arange, actions]`), :77 (`vmap(simulator.cross)`)
// and :132-134 (`GEBV_model(populations)`): the offspring are written to HBM once and never read back, the
// 2x-population parent gather is never materialised.  Algorithmic HBM traffic: 0.75 B per offspring-marker
//  -- the separate blend + GEBV pass pays 1.0 B.
//
// Work = ITEMS (tile of 128 offspring) x (a range of stages of 4 x 128 markers).  Every tile's K range is cut into the
// same few parts of DECREASING length (e.g. 8 + 6 + 4 + 2 stages at 10 000 markers), and the items are numbered
// part-major: all tiles' longest parts first, the 2-stage parts last.  The grid is one wave (2 CTAs per SM); CTA c
// starts with item c and takes every further item from a global atomic counter, so a CTA that happens to run slowly
// (L2 / die distance, DRAM contention: finish times of equal static shares differed by 1.6x) simply takes fewer
// items, and the kernel's tail is at most one 2-stage item.  The pipeline runs THROUGH the item boundaries: the next
// item is fetched one item ahead, its parents' rows are derived while the current item streams, the MMA warp
// alternates between two accumulators in tensor memory, and the loader warps drain a finished accumulator between two
// gathers -- so only the FIRST item of a CTA pays the ramp (row table -> first gathers -> first data).  Against one
// CTA per (tile, K range) in 1.9 waves (cross_gebv.cu) this removes the second wave's ramp and the wave tail.
//
// The item queue: 8 entries in shared memory, entry k % 8 = the CTA's k-th item {tile, first stage, stages}; the
// producer is loader thread 0 (it issues the atomic at the start of item k and publishes the answer as item k + 1
// one ring depth later, so the counter's round trip never stalls a gather); every role waits for entry k on the
// mbarrier q_full[k % 8] (phase = (k / 8) & 1) and stops at the entry with 0 stages.  Every item has >= 2 stages, the
// expanders trail the loaders by <= 3 stages, the MMA warp by <= 2 more, an epilogue by <= 2 items (two
// accumulators): the oldest entry still in use is k - 6 when entry k + 1 is published, so 8 entries need no
// "empty" barrier.
//
// The 128 offspring of a tile are taken CHILD-major: tile row R <-> (child i = R / E, env e = R % E).  The crossover
// masks depend on the child slot only (the reference shares one key across envs, vec_env.py:75-77), so with E a
// multiple of 32 all lanes of a warp read the SAME mask words: one broadcast access per warp.  Warp roles:
//
//   warps 10-13 (loaders)   : cp.async (LDGSTS) 16-byte copies, four lanes per 64-byte row segment: the two bit
//                             planes of parent A and of parent B of every offspring -> a ring of stages
//                             [4 chunks][128 rows][64 B = 4 steps] in shared memory (16-byte quarters XOR-swizzled
//                             with the row, so the one-row-per-lane reads below are conflict free).  Nothing
//                             waits on a scoreboard: completion lands on an mbarrier
//                             (cp.async.mbarrier.arrive.noinc), so the bytes in flight are bounded by the ring
//                             (3 x 32 KB per CTA, 2 CTAs per SM), not by registers.  Every loader thread derives the
//                             parents of its own 4 rows straight from the action array: for the first item before
//                             the CTA's setup barrier (overlapping the TMEM allocation), for every later one an item
//                             ahead (parked in a 1.5 KB table).  Threads 0-63 also stage the mask rows of the
//                             tile's (<= 8) children.  Before a ring slot is gathered into again, the same warps drain
//                             the offspring words the expanders left in it to HBM (coalesced 128-bit stores).
//   warps 0-7  (expanders)  : thread t <-> offspring t of the tile <-> TMEM lane t; the two groups of 4 warps take
//                             alternate steps.  4 x ld.shared.v4 + the two mask quads (broadcast ld.shared; with
//                             few envs per-lane ld.global.nc, prefetched a step ahead), one LOP3 per word selects
//                             the alleles (h0 & ~M | h1 & M), the offspring words replace parent A's IN PLACE in
//                             the stage, then 4 words per plane -> 128 prescaled dosage bytes -> tcgen05.st into the
//                             A stage in tensor memory.
//   warp 8     (digits)     : 1-D bulk copies (TMA) of the digit tiles, one per pair of steps, after an L2 prefetch
//                             of the segment's digit range.
//   warp 9     (MMA)        : tcgen05.mma.kind::i8, A from TMEM, B from shared memory, D in TMEM (two accumulators,
//                             alternating per segment); the whole warp runs the loop and one elected lane issues
//                             (elect.sync), one barrier round and one tcgen05.commit per PAIR of steps.
//   warps 10-13 (epilogue)  : the loader warps also own the epilogues (their warp ids cover the four TMEM lane
//                             quadrants): between two stages they TEST the finished-accumulator barrier of the oldest
//                             open segment and, once it has completed, turn its digits into int64 -> one 64-bit atomic
//                             per value carrying the partial sum of the tile's K range and the arrival count -> float32
//                             by the last arrival (tc_common.cuh).  The expanders never stop at a segment boundary.
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
constexpr int XG_Q = 8;               // item queue entries (see the header: no "empty" barrier needed)
constexpr int XG_PARTS_MAX = 8;       // parts a tile's K range is cut into
constexpr int XG_MAX_STEPS = 3000;    // per tile: int32 digit sums (a 128-marker step adds at most 16 * 43520 to one)
#ifndef XG_REGCAP_THREADS
#define XG_REGCAP_THREADS 512
#endif
constexpr uint32_t XG_ROW = 16 * XG_SPS;                // bytes per row and plane in a stage
constexpr uint32_t XG_CHUNK = TILE_M * XG_ROW;          // one plane of a stage: 128 rows x 64 B
constexpr uint32_t XG_IN_BYTES = 4 * XG_CHUNK;          // parent A planes 0/1 (-> offspring planes 0/1), parent B planes 0/1
constexpr uint32_t XG_MASK_BYTES = XG_MC * 2 * XG_ROW;  // mask rows of up to 8 children
constexpr uint32_t XG_NOROW = 0xFFFFFFFFu;

#ifndef XG_TRACE
#define XG_TRACE 0   // 1: CTA XG_TRACE_CTA records clock64() stamps of its pipeline events (diagnostics build only)
#endif
#if XG_TRACE
#ifndef XG_TRACE_CTA
#define XG_TRACE_CTA 0
#endif
__device__ long long xgd_trace_buf[16 * 64];
#define XG_STAMP(slot, idx)                                                                            \
    do {                                                                                               \
        if (blockIdx.x == XG_TRACE_CTA && (idx) < 64) xgd_trace_buf[(slot) * 64 + (idx)] = clock64(); \
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
    uint64_t d_full[2], d_free[2];
    uint64_t q_full[XG_Q];
};

struct XGArgs {
    const uint4 *pop;         // [E][n_src][2][W4]
    const int32_t *parents;   // [E][n][2]
    const uint4 *mask;        // [2n][W4]
    uint4 *out_pop;           // [E][n][2][W4]
    int64_t n_src, n, E, rows;  // rows = E * n
    int W4;
    uint32_t tiles;
    uint32_t items;           // tiles * parts
    int parts;
    uint16_t part_s0[XG_PARTS_MAX], part_len[XG_PARTS_MAX];  // stages; every part has >= 2
    unsigned int *work;       // [0] items handed out beyond the first gridDim.x, [1] CTAs that have finished
};

// an item: stages [s0, s0 + len) of tile `tile`; len == 0 ends the CTA's list
struct Seg {
    uint32_t tile, s0, len;
};
// item w of the part-major numbering (w >= items: the end marker)
__device__ __forceinline__ Seg item_at(const XGArgs &fa, uint32_t w)
{
    Seg s;
    s.tile = s.s0 = s.len = 0;
    if (w < fa.items) {
        const uint32_t part = w / fa.tiles;
        s.tile = w - part * fa.tiles;
        s.s0 = fa.part_s0[part];
        s.len = fa.part_len[part];
    }
    return s;
}

// rows (lu >> 2) + 32 k, k < 4, of tile `tile`: uint4 offsets of the parents' plane-0 rows in `pop` and the output row
struct RowInfo {
    uint32_t src[4][2], orow[4];
};
__device__ __forceinline__ void load_rows(const XGArgs &fa, uint32_t tile, int lu, RowInfo &ri)
{
    const int64_t row0 = (int64_t)tile * TILE_M;
    const uint32_t E32 = (uint32_t)fa.E;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t R = row0 + (lu >> 2) + 32 * k;
        ri.src[k][0] = ri.src[k][1] = 0;
        ri.orow[k] = XG_NOROW;
        if (R < fa.rows) {
            const uint32_t i = (uint32_t)R / E32, e = (uint32_t)R - i * E32;
            ri.orow[k] = e * (uint32_t)fa.n + i;
            const int2 pr = __ldg(reinterpret_cast<const int2 *>(fa.parents) + ri.orow[k]);
            int64_t a = pr.x, b = pr.y;
            a += a < 0 ? fa.n_src : 0;  // jnp indexing: negatives wrap once, then clamp
            a = a < 0 ? 0 : (a > fa.n_src - 1 ? fa.n_src - 1 : a);
            b += b < 0 ? fa.n_src : 0;
            b = b < 0 ? 0 : (b > fa.n_src - 1 ? fa.n_src - 1 : b);
            ri.src[k][0] = (uint32_t)((((int64_t)e * fa.n_src + a) * 2) * fa.W4);
            ri.src[k][1] = (uint32_t)((((int64_t)e * fa.n_src + b) * 2) * fa.W4);
        }
    }
}

// the same in two halves, so that the action loads of the NEXT item are in flight while the current item's stages are
// gathered: rows_issue starts the loads, rows_finish (an item later) turns them into row offsets
struct RowLoads {
    int2 pr[4];
};
__device__ __forceinline__ void rows_issue(const XGArgs &fa, uint32_t tile, int lu, RowLoads &rl)
{
    const int64_t row0 = (int64_t)tile * TILE_M;
    const uint32_t E32 = (uint32_t)fa.E;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t R = row0 + (lu >> 2) + 32 * k;
        rl.pr[k] = make_int2(0, 0);
        if (R < fa.rows) {
            const uint32_t i = (uint32_t)R / E32, e = (uint32_t)R - i * E32;
            rl.pr[k] = __ldg(reinterpret_cast<const int2 *>(fa.parents) + (e * (uint32_t)fa.n + i));
        }
    }
}
__device__ __forceinline__ void rows_finish(const XGArgs &fa, uint32_t tile, int lu, const RowLoads &rl, RowInfo &ri)
{
    const int64_t row0 = (int64_t)tile * TILE_M;
    const uint32_t E32 = (uint32_t)fa.E;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t R = row0 + (lu >> 2) + 32 * k;
        ri.src[k][0] = ri.src[k][1] = 0;
        ri.orow[k] = XG_NOROW;
        if (R < fa.rows) {
            const uint32_t i = (uint32_t)R / E32, e = (uint32_t)R - i * E32;
            ri.orow[k] = e * (uint32_t)fa.n + i;
            int64_t a = rl.pr[k].x, b = rl.pr[k].y;
            a += a < 0 ? fa.n_src : 0;  // jnp indexing: negatives wrap once, then clamp
            a = a < 0 ? 0 : (a > fa.n_src - 1 ? fa.n_src - 1 : a);
            b += b < 0 ? fa.n_src : 0;
            b = b < 0 ? 0 : (b > fa.n_src - 1 ? fa.n_src - 1 : b);
            ri.src[k][0] = (uint32_t)((((int64_t)e * fa.n_src + a) * 2) * fa.W4);
            ri.src[k][1] = (uint32_t)((((int64_t)e * fa.n_src + b) * 2) * fa.W4);
        }
    }
}

// one mbarrier.try_wait (suspends up to BG_MBAR_HINT_NS): true when the phase has completed
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"((uint32_t)BG_MBAR_HINT_NS)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}

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
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// byte offset of 16-byte quarter q of row t inside a [128 rows][64 B] chunk: quarters XOR-swizzled with the row, so
// that 8 consecutive rows reading the same quarter hit 8 different 16-byte bank groups
__device__ __forceinline__ uint32_t swz(int t, int q) { return (uint32_t)t * XG_ROW + (uint32_t)((q ^ ((t >> 1) & 3)) * 16); }

// smem: stage ring [XG_R][4][128][64 B], mask ring [XG_R][8 children][2][64 B], digit ring [nbp pairs][2 steps][N/8][8 ki][8][16 B]
// launch bounds of 512 threads (448 are launched): caps the kernel at 64 registers per thread, so that two CTAs leave
// 8192 registers of the SM free -- exactly one 128-thread CTA of the mask kernel, which then runs in the issue slots
// this (latency-bound) kernel leaves idle instead of displacing its CTAs
__global__ void __launch_bounds__(XG_REGCAP_THREADS, XG_CTAS)
    cross_gebv_dyn_kernel(const __grid_constant__ XGArgs fa, const int8_t *__restrict__ bdig, int N, int T, int D, int nbp,
                      unsigned long long *__restrict__ acc, const double *__restrict__ inv_scale, float *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) XGBars bars;
    __shared__ uint32_t tmem_base_slot;
    // per ring slot: where the offspring words of the stage it holds go (written by the loader lane that gathers row t,
    // read back by the four lanes that drain it: same warp)
    __shared__ uint32_t slot_orow[XG_R][TILE_M];
    __shared__ uint32_t slot_w4[XG_R][4];  // one copy per loader warp
    // the NEXT item's rows, fetched while the current one streams: [row][parent A offset, parent B offset, output row]
    __shared__ uint32_t next_rows[TILE_M][3];
    __shared__ Seg queue[XG_Q];  // the CTA's item list (header comment)

    // warp index through a shuffle: provably warp-uniform, so the role dispatch below and the MMA warp's address
    // arithmetic can live on the uniform datapath
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) XG_STAMP(15, 0);
    // programmatic dependent launch, both ends (see cross_gebv.cu): the next step kernel's CTAs may take this grid's
    // freed slots and run their prologue while it drains; this kernel's prologue reads the action array only
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t in_base = smem_u32(smem);
    const uint32_t mask_base = in_base + XG_R * XG_IN_BYTES;
    const uint32_t b_base0 = mask_base + XG_R * XG_MASK_BYTES;
    const uint32_t b_bytes = (uint32_t)N * STEP_K;
    const Seg first_item = item_at(fa, blockIdx.x);  // (the launcher keeps gridDim.x <= items)
    const uint32_t E32 = (uint32_t)fa.E;

    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2 * d_cols + XG_S * (STEP_K / 4)) tmem_cols <<= 1;

    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 8 * 32) {
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
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars.d_full[i]), 1);   // tcgen05.commit after a segment's last MMA
            mbar_init(smem_u32(&bars.d_free[i]), 4);   // the 4 epilogue warps have read the accumulator
        }
        for (int i = 0; i < XG_Q; ++i) mbar_init(smem_u32(&bars.q_full[i]), 1);  // the producer (loader thread 0)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        queue[0] = first_item;  // entry 0 is static: item blockIdx.x
        mbar_arrive(smem_u32(&bars.q_full[0]));
    }
    // the loaders' first dependent chain (actions -> parent rows -> gathers) starts before the CTA's setup barrier
    RowInfo ri;
    if (warp >= XG_LOADER_WARP0) load_rows(fa, first_item.tile, tid - XG_LOADER_WARP0 * 32, ri);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // (the work counter below is reset by the predecessor's last CTA)
    uint32_t pend = 0;  // loader thread 0: the counter's answer for the item after the next one to be published
    if (tid == XG_LOADER_WARP0 * 32) pend = gridDim.x + atomicAdd(fa.work, 1u);  // item 1 of this CTA
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t tmem_a = tmem_d + 2 * d_cols;
    // entry k of the CTA's item list (blocks until the producer has published it)
    auto item = [&](uint32_t k) -> Seg {
        mbar_wait(smem_u32(&bars.q_full[k % XG_Q]), (k / XG_Q) & 1);
        return queue[k % XG_Q];
    };

    if (warp < 8) {
        // ---------------- expanders: group g takes the steps j with j % 2 == g ----------------
        const int g = warp >> 2, r = tid & (TILE_M - 1);
        const bool stamp = (warp & 3) == 0 && lane == 0;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quadrant
        int ap = 0;  // pair stage of step j = (j / 2) % XG_SP; group g fills half g of it
        uint32_t a_use = 0;
        uint32_t gst = 0;  // stages processed so far by this CTA (ring slot / parity)
        for (uint32_t k = 0;; ++k) {
            const Seg sg = item(k);
            if (sg.len == 0) break;
            if (stamp) XG_STAMP(5 + 2 * g, k);  // expander group g starts item k (slots 5 / 7)
            const int64_t row0 = (int64_t)sg.tile * TILE_M;
            const int64_t last_row = min(fa.rows, row0 + TILE_M) - 1;
            const uint32_t i0 = (uint32_t)row0 / E32;
            const int nchild = (int)((uint32_t)last_row / E32 - i0) + 1;
            const bool mask_smem = nchild <= XG_MC;  // else (few envs): every thread fetches its own mask words from L2
            const int64_t R = row0 + r;
            uint32_t msk_a = 0, msk_b = 0;
            if (R < fa.rows) {
                const uint32_t i = (uint32_t)R / E32;
                // mask rows 2i, 2i + 1: byte offset in a mask stage (smem mode) or uint4 offset in the global mask array
                msk_a = mask_smem ? (2 * (i - i0)) * XG_ROW : (2 * i) * (uint32_t)fa.W4;
                msk_b = mask_smem ? msk_a + XG_ROW : msk_a + (uint32_t)fa.W4;
            }
            const uint4 *mrow_a = fa.mask + (mask_smem ? 0u : msk_a), *mrow_b = fa.mask + (mask_smem ? 0u : msk_b);
            uint4 ma_next = make_uint4(0, 0, 0, 0), mb_next = ma_next;
            if (!mask_smem) {
                ma_next = __ldg(mrow_a + XG_SPS * sg.s0 + g);
                mb_next = __ldg(mrow_b + XG_SPS * sg.s0 + g);
            }
            for (uint32_t st = sg.s0; st < sg.s0 + sg.len; ++st, ++gst) {
                const int rs = gst % XG_R;
                mbar_wait(smem_u32(&bars.raw_full[rs]), (gst / XG_R) & 1);
                if (stamp) XG_STAMP(2 + g, gst);  // raw_full seen
                const uint32_t stage = in_base + rs * XG_IN_BYTES, mstage = mask_base + rs * XG_MASK_BYTES;
#pragma unroll
                for (int k = 0; k < XG_SPS / 2; ++k) {
                    const int q = 2 * k + g;  // quarter of the stage row = step XG_SPS * st + q of the row
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
                        const uint32_t nxt = XG_SPS * st + q + 2;  // this thread's next step
                        if (nxt < XG_SPS * (sg.s0 + sg.len)) {
                            ma_next = __ldg(mrow_a + nxt);
                            mb_next = __ldg(mrow_b + nxt);
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
                    if (a_use > 0) mbar_wait(smem_u32(&bars.a_empty[ap]), (a_use - 1) & 1);  // MMAs of the previous use retired
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    dosage_to_tmem(tmem_a + lane_sel + (uint32_t)(2 * ap + g) * (STEP_K / 4), f);
                    if (stamp) XG_STAMP(8, XG_SPS * gst + q);  // TMEM stores complete
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars.a_full[ap]));
                    if (++ap == XG_SP) {
                        ap = 0;
                        ++a_use;
                    }
                }
            }
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ---------------- digit tiles: one bulk copy (TMA) per pair of steps, up to nbp pairs ahead ----------------
            const uint32_t pair_bytes = 2 * b_bytes;
            int slot = 0;
            uint32_t use = 0;
            for (uint32_t k = 0;; ++k) {
                const Seg sg = item(k);
                if (sg.len == 0) break;
                const int64_t s_begin = (int64_t)XG_SPS * sg.s0;
                const int nst = XG_SPS * (int)sg.len;
                // the segment's digit range -> L2 now (the table is usually cold: a step streams more than L2 holds),
                // so that the ring below is fed at L2 latency
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(bdig + s_begin * b_bytes), "r"((uint32_t)nst * b_bytes)
                             : "memory");
                for (int pj = 0; pj < nst / 2; ++pj) {
                    if (use > 0) mbar_wait(smem_u32(&bars.b_empty[slot]), (use - 1) & 1);
                    const uint32_t full = smem_u32(&bars.b_full[slot]);
                    mbar_arrive_expect_tx(full, pair_bytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     b_base0 + slot * pair_bytes),
                                 "l"(bdig + (s_begin + 2 * pj) * b_bytes), "r"(pair_bytes), "r"(full)
                                 : "memory");
                    if (++slot == nbp) {
                        slot = 0;
                        ++use;
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues; one round per PAIR of steps ----------------
        const uint32_t idesc = idesc_u8s8(N);
        int ap = 0, slot = 0;
        uint32_t a_par = 0, b_par = 0, seg_idx = 0, gp = 0;
        for (;; ++seg_idx) {
            const Seg sg = item(seg_idx);
            if (sg.len == 0) break;
            const uint32_t buf = seg_idx & 1;
            const uint32_t d_taddr = tmem_d + buf * d_cols;
            if (seg_idx >= 2) mbar_wait(smem_u32(&bars.d_free[buf]), ((seg_idx >> 1) - 1) & 1);  // its previous segment has been read
            if (lane == 0) XG_STAMP(6, seg_idx);  // MMA warp starts item seg_idx
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int npairs = (XG_SPS / 2) * (int)sg.len;
            for (int pj = 0; pj < npairs; ++pj, ++gp) {
                const uint32_t a_taddr = tmem_a + (uint32_t)(2 * ap) * (STEP_K / 4);
                const uint64_t bdesc = make_smem_desc(b_base0 + (uint32_t)slot * 2 * b_bytes, 128, 1024);
                mbar_wait(smem_u32(&bars.b_full[slot]), b_par);
                if (lane == 0) XG_STAMP(10, gp);  // digit pair landed
                mbar_wait(smem_u32(&bars.a_full[ap]), a_par);
                if (lane == 0) XG_STAMP(11, gp);  // A pair full
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int kk = 0; kk < STEP_K / 32; ++kk)  // +16 in the descriptor's address field = +256 bytes
                        mma_i8_ts_warp(d_taddr, a_taddr + h * (STEP_K / 4) + 8 * kk, bdesc + (uint64_t)(h * (b_bytes >> 4)) + 16 * kk,
                                       idesc, (pj > 0 || h > 0 || kk > 0) ? 1u : 0u);
                mma_commit_warp(smem_u32(&bars.a_empty[ap]));
                mma_commit_warp(smem_u32(&bars.b_empty[slot]));
                if (lane == 0) XG_STAMP(12, gp);  // commits issued
                if (++ap == XG_SP) {
                    ap = 0;
                    a_par ^= 1;
                }
                if (++slot == nbp) {
                    slot = 0;
                    b_par ^= 1;
                }
            }
            mma_commit_warp(smem_u32(&bars.d_full[buf]));
        }
    } else {
        // ---------------- loaders / storers / epilogues: thread lu covers quarter q = lu & 3 of rows (lu >> 2) + 32k ----------------
        // per stage: drain the offspring words of the stage that used this ring slot XG_R stages ago (coalesced
        // 128-bit stores), then gather the next stage into it (both planes of both parents: 16 cp.async in flight per
        // thread, no register staging) plus, threads lu < 64, one 16-byte piece of the tile's mask rows
        const int lu = tid - XG_LOADER_WARP0 * 32, q = lu & 3;
        uint32_t dst[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[k] = swz((lu >> 2) + 32 * k, q);
        const int mrow = lu >> 2;  // mask piece of this thread: row 2 * i0 + mrow (0 .. 15 for lu < 64), quarter q
        uint32_t gst = 0;  // stages gathered so far

        // epilogues: segment `epi_idx` (units from epi_u on) is the oldest one whose accumulator has not been read
        uint32_t epi_idx = 0, seg_idx = 0;
        const int ew = warp & 3, et = ew * 32 + lane;  // TMEM lane quadrant of this warp, row of the tile of this thread
        auto epilogue = [&](bool block) -> bool {
            const uint32_t buf = epi_idx & 1, par = (epi_idx >> 1) & 1;
            if (block) mbar_wait(smem_u32(&bars.d_full[buf]), par);
            else if (!__shfl_sync(0xffffffffu, (int)mbar_test(smem_u32(&bars.d_full[buf]), par), 0)) return false;  // warp-uniform
            if (lu == 0) XG_STAMP(1, 2 * epi_idx);  // accumulator complete (seen)
            const Seg sg = queue[epi_idx % XG_Q];  // (published long ago: this warp has already gathered the item)
            const int64_t row0 = (int64_t)sg.tile * TILE_M, R = row0 + et;
            int64_t orow = -1;
            if (R < fa.rows) {
                const uint32_t i = (uint32_t)R / E32, e = (uint32_t)R - i * E32;
                orow = (int64_t)e * fa.n + i;
            }
            // the tile's K range arrives in fa.parts partial sums (one per item of the tile)
            digits_epilogue(tmem_d + buf * d_cols, et, ew, row0, fa.rows, T, acc, inv_scale, out, (unsigned)fa.parts, D, orow);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.d_free[buf]));
            if (lu == 0) XG_STAMP(1, 2 * epi_idx + 1);  // epilogue done
            ++epi_idx;
            return true;
        };

        auto drain = [&](uint32_t so) {
            const int rs = so % XG_R;
            const uint32_t stage = in_base + rs * XG_IN_BYTES;
            // (polling: an accumulator that completes meanwhile is drained here -- the MMA warp may be waiting for it, and
            //  the expanders behind the MMA warp are what this wait depends on)
            while (!__all_sync(0xffffffffu, mbar_try(smem_u32(&bars.stage_done[rs]), (so / XG_R) & 1)))
                if (epi_idx < seg_idx) epilogue(false);
            if (lu == 0) XG_STAMP(14, so);  // stage_done seen
            const uint32_t w4 = slot_w4[rs][lu >> 5] + q;
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {  // two rows at a time: 4 x 128 bits in registers
                uint4 v[2][2];
                uint32_t orow[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    orow[k] = slot_orow[rs][(lu >> 2) + 32 * (2 * hk + k)];
                    v[k][0] = lds128(stage + dst[2 * hk + k]);
                    v[k][1] = lds128(stage + XG_CHUNK + dst[2 * hk + k]);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (orow[k] != XG_NOROW && !(XG_DEBUG_SKIP & 2)) {
                        uint4 *o = fa.out_pop + (int64_t)orow[k] * 2 * fa.W4 + w4;
                        o[0] = v[k][0];
                        o[fa.W4] = v[k][1];
                    }
            }
        };

        for (;; ++seg_idx) {
            const Seg sg = item(seg_idx);
            if (sg.len == 0) break;
            if (lu == 0) XG_STAMP(4, seg_idx);  // loaders start item seg_idx
            // the producer publishes the NEXT item (asked for an item ago) and asks for the one after it; the next item's
            // action loads are issued now and consumed after this item's last gather: both round trips hide behind an item
            // (the CTA's FIRST item does all that only after its first XG_R gathers: 2 x 148 CTAs ask the counter at once)
            Seg nxt;
            bool has_next = false;
            RowLoads rl;
            auto announce = [&]() {
                if (lu == 0) {
                    const Seg nx = item_at(fa, pend);  // an end marker once the counter has run past the last item
                    queue[(seg_idx + 1) % XG_Q] = nx;
                    mbar_arrive(smem_u32(&bars.q_full[(seg_idx + 1) % XG_Q]));  // (release: orders the entry before the arrival)
                    if (nx.len != 0) pend = gridDim.x + atomicAdd(fa.work, 1u);
                }
                nxt = item(seg_idx + 1);
                if (lu == 0) XG_STAMP(9, seg_idx);  // next item known
                has_next = nxt.len != 0;
                if (has_next && q == 0) rows_issue(fa, nxt.tile, lu, rl);
            };
            const uint32_t announce_at = seg_idx == 0 ? min(sg.s0 + (uint32_t)XG_R, sg.s0 + sg.len - 1) : sg.s0;
            if (seg_idx != 0) announce();
            const int64_t row0 = (int64_t)sg.tile * TILE_M;
            const int64_t last_row = min(fa.rows, row0 + TILE_M) - 1;
            const uint32_t i0 = (uint32_t)row0 / E32;
            const int nchild = (int)((uint32_t)last_row / E32 - i0) + 1;
            const bool mask_smem = nchild <= XG_MC;
            if (lu == 0) XG_STAMP(0, gst);  // this segment's parents are in registers (index: its first stage)
            const bool mask_loader = mask_smem && lu < 2 * XG_MC * 4;
            const bool mask_valid = mask_loader && (mrow >> 1) < nchild;
            const uint4 *msrc = fa.mask + (mask_valid ? (int64_t)(2 * i0 + mrow) * fa.W4 : 0);
            for (uint32_t st = sg.s0; st < sg.s0 + sg.len; ++st, ++gst) {
                const int rs = gst % XG_R;
                const uint32_t stage = in_base + rs * XG_IN_BYTES;
                if (gst >= XG_R) drain(gst - XG_R);
                if (epi_idx < seg_idx) epilogue(false);  // a finished accumulator? (test, no wait)
                // (the slots gathered into below were read by THIS warp's stores above, or by the expanders that
                //  signalled stage_done: no other thread still needs them)
                if (lu == 0) XG_STAMP(13, gst);  // gathers of stage `gst` issued
                const int w4 = XG_SPS * (int)st + q;
                __syncwarp();  // the drain above has read this slot's destination rows
                if (q == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) slot_orow[rs][(lu >> 2) + 32 * k] = ri.orow[k];
                }
                if ((lu & 31) == 0) slot_w4[rs][lu >> 5] = XG_SPS * st;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t nbytes = (ri.orow[k] != XG_NOROW && !(XG_DEBUG_SKIP & 4)) ? 16u : 0u;  // 0: zero-fill, nothing is read
                    const uint32_t d = stage + dst[k];
                    cp_async16(d, fa.pop + ri.src[k][0] + w4, nbytes);
                    cp_async16(d + XG_CHUNK, fa.pop + ri.src[k][0] + fa.W4 + w4, nbytes);
                    cp_async16(d + 2 * XG_CHUNK, fa.pop + ri.src[k][1] + w4, nbytes);
                    cp_async16(d + 3 * XG_CHUNK, fa.pop + ri.src[k][1] + fa.W4 + w4, nbytes);
                }
                if (mask_loader) cp_async16(mask_base + rs * XG_MASK_BYTES + (uint32_t)mrow * XG_ROW + q * 16, msrc + w4, mask_valid ? 16u : 0u);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars.raw_full[rs])) : "memory");
                if (seg_idx == 0 && st == announce_at) announce();
            }
            if (has_next) {
                if (q == 0) {  // the next item's rows -> the table the four lanes of a row share
                    RowInfo nx;
                    rows_finish(fa, nxt.tile, lu, rl, nx);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t *e = next_rows[(lu >> 2) + 32 * k];
                        e[0] = nx.src[k][0];
                        e[1] = nx.src[k][1];
                        e[2] = nx.orow[k];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t *e = next_rows[(lu >> 2) + 32 * k];
                    ri.src[k][0] = e[0];
                    ri.src[k][1] = e[1];
                    ri.orow[k] = e[2];
                }
                __syncwarp();  // (the table is rewritten during the next segment's first stage)
            }
        }
        // the last XG_R stages, then the accumulators still open
        __syncwarp();
        for (uint32_t so = gst > XG_R ? gst - XG_R : 0; so < gst; ++so) {
            drain(so);
            if (epi_idx < seg_idx) epilogue(false);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        while (epi_idx < seg_idx) epilogue(true);
    }

    if (tid == 0) XG_STAMP(15, 3);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    // the last CTA to finish puts the launch's work counter back (every CTA's item requests precede this barrier)
    if (tid == 0 && atomicAdd(fa.work + 1, 1u) == gridDim.x - 1) {
        fa.work[0] = 0;
        fa.work[1] = 0;
    }
}

}  // namespace

#if XG_TRACE
extern "C" int bg_debug_read_trace_dyn(long long *host, int n)
{
    return (int)cudaMemcpyFromSymbol(host, xgd_trace_buf, sizeof(long long) * n);
}
#endif

int bg_tc_reserve_scratch(bg_engine *eng, int scratch, int64_t total, int64_t tiles, cudaStream_t st);

static size_t xgd_smem_bytes(int N, int *nbp_out)
{
    // digit ring: nbp pairs of steps, as deep as fits beside the stage ring with XG_CTAS CTAs per SM
    const size_t rings = (size_t)XG_R * (XG_IN_BYTES + XG_MASK_BYTES), b_bytes = (size_t)N * STEP_K;
    const size_t per_cta = 228 * 1024 / XG_CTAS - 1024 - 4096;  // minus the reserved KB and the static arrays
    int nbp = per_cta > rings ? (int)((per_cta - rings) / (2 * b_bytes)) : 0;
    if (nbp > XG_BP_MAX) nbp = XG_BP_MAX;
    if (nbp < 2) nbp = 2;
    if (nbp_out) *nbp_out = nbp;
    return rings + (size_t)nbp * 2 * b_bytes;
}

// can the persistent fused kernel take this engine's trait count and this population size?
bool bg_cross_gebv_dyn_ok(const bg_engine *eng, int64_t E, int64_t n_src, int64_t n)
{
    if (!eng || !eng->d_wdig || eng->mut_thr) return false;
    const int N = eng->tc_N;
    const size_t smem = xgd_smem_bytes(N, nullptr);
    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    if (smem + 4096 > (size_t)eng->max_smem_optin || 2 * d_cols + XG_S * (STEP_K / 4) > 512) return false;  // + the static arrays
    if (eng->tc_steps % (2 * XG_SPS) != 0 || eng->tc_steps > XG_MAX_STEPS || E * n >= (int64_t(1) << 31)) return false;
    const int64_t tiles = (E * n + TILE_M - 1) / TILE_M;
    if (tiles * XG_PARTS_MAX >= (int64_t(1) << 31)) return false;
    return (int64_t)eng->Wpad / 4 * 2 * (n_src > n ? n_src : n) * E < (int64_t(1) << 32);
}

// the parts a tile's K range (spt stages) is cut into: decreasing, every part >= 2 stages.  Few tiles per CTA: four
// parts in the proportions 8 : 6 : 4 : 2 (the tail of the kernel is one 2-stage item); many tiles per CTA: the bulk
// in one part (fewer accumulator hand-overs), then 4- and 2-stage parts to level the end.
static int xgd_parts(const bg_engine *eng, int spt, int64_t tiles, int64_t G, uint16_t *s0, uint16_t *len)
{
    int w[XG_PARTS_MAX], P = 0;
    long long code = eng->opt.xg_parts;  // e.g. 8642: proportions, most significant digit first
    if (code <= 0) code = tiles >= 3 * G ? 1422 : 8642;
    int digits[XG_PARTS_MAX], nd = 0;
    for (; code > 0 && nd < XG_PARTS_MAX; code /= 10)
        if (code % 10) digits[nd++] = (int)(code % 10);
    for (int i = nd - 1; i >= 0; --i) w[P++] = digits[i];
    if (P > spt / 2) P = spt / 2;
    if (P < 1) P = 1;
    int wsum = 0;
    for (int i = 0; i < P; ++i) wsum += w[i];
    int used = 0;
    for (int i = P - 1; i >= 1; --i) {  // the small parts first, the first part takes the remainder
        int l = (int)((long long)spt * w[i] / wsum);
        if (l < 2) l = 2;
        len[i] = (uint16_t)l;
        used += l;
    }
    while (P > 1 && spt - used < 2) {  // (cannot happen with P <= spt / 2 and sane proportions; keep every part >= 2)
        used -= len[P - 1];
        --P;
    }
    len[0] = (uint16_t)(spt - used);
    int at = 0;
    for (int i = 0; i < P; ++i) {
        s0[i] = (uint16_t)at;
        at += len[i];
    }
    return P;
}

// vector-env step: out_pop[e][i] = cross of pop[e][parents[e][i][0..1]] under mask[2i..2i+1]; gebv[e][i][T]
int bg_launch_cross_gebv_dyn(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask, uint32_t *out_pop,
                             int64_t E, int64_t n_src, int64_t n, float *gebv_out, cudaStream_t st)
{
    BG_REQUIRE(eng && eng->d_wdig, BG_ESTATE, "engine has no tensor-core digit table");
    static_assert(XG_S % 2 == 0, "the two expander groups alternate over an even number of A stages");
    BG_REQUIRE(bg_cross_gebv_dyn_ok(eng, E, n_src, n), BG_ELIMIT, "shape outside the fused cross+GEBV kernel's limits");
    const int T = eng->T, N = eng->tc_N;
    const int steps = (int)eng->tc_steps;  // a multiple of 8: rows are padded to 32 words
    const int64_t rows = E * n;
    const int64_t tiles = (rows + TILE_M - 1) / TILE_M;
    const int spt = steps / XG_SPS;

    int nbp = 2;
    const size_t smem = xgd_smem_bytes(N, &nbp);
    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tcols = 32;
    while (tcols < 2 * d_cols + XG_S * (STEP_K / 4)) tcols <<= 1;
    int resident = (int)(512 / tcols);
    const int by_smem = (int)(228 * 1024 / (smem + 1024 + 4096));
    if (by_smem < resident) resident = by_smem;
    if (resident > XG_CTAS) resident = XG_CTAS;
    if (resident < 1) resident = 1;
    XGArgs fa;
    int64_t G = eng->opt.tc_target_ctas > 0 ? eng->opt.tc_target_ctas : (int64_t)resident * eng->sm_count;
    if (G > (int64_t)resident * eng->sm_count) G = (int64_t)resident * eng->sm_count;  // one wave: every CTA must be resident
    fa.parts = xgd_parts(eng, spt, tiles, G, fa.part_s0, fa.part_len);
    const int64_t items = tiles * fa.parts;
    if (G > items) G = items;
    int rc = bg_tc_reserve_scratch(eng, 0, rows * T, tiles, st);
    if (rc) return rc;
    if (!eng->d_xg_work) {  // one work-counter pair per launch in flight (64 slots, each reset by its launch's last CTA)
        BG_CUDA(cudaMalloc(&eng->d_xg_work, 64 * 2 * sizeof(unsigned int)));
        BG_CUDA(cudaMemsetAsync(eng->d_xg_work, 0, 64 * 2 * sizeof(unsigned int), st));
    }
    if (smem > eng->tc2_optin[3]) {
        BG_CUDA(cudaFuncSetAttribute(cross_gebv_dyn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eng->tc2_optin[3] = smem;
    }
    fa.pop = reinterpret_cast<const uint4 *>(pop);
    fa.parents = parents;
    fa.mask = reinterpret_cast<const uint4 *>(mask);
    fa.out_pop = reinterpret_cast<uint4 *>(out_pop);
    fa.n_src = n_src;
    fa.n = n;
    fa.E = E;
    fa.rows = rows;
    fa.W4 = eng->Wpad / 4;
    fa.tiles = (uint32_t)tiles;
    fa.items = (uint32_t)items;
    fa.work = eng->d_xg_work + 2 * (eng->xg_seq++ % 64);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)G);
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
    BG_CUDA(cudaLaunchKernelEx(&cfg, cross_gebv_dyn_kernel, fa, bd, N, T, D, nbp, accp, inv, gebv_out));
    BG_LAUNCHED();
    return BG_OK;
}

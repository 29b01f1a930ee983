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

#include <algorithm>
#include <string.h>

#include "bg_internal.h"
#include "tc_common.cuh"

using namespace bgtc;

namespace {

constexpr int T2_M = TILE_M;
constexpr int T2_KS = STEP_K;       // markers per step (4 words per plane, 32 TMEM columns of int8x4)
#ifndef T2_R_VAL
#define T2_R_VAL 4
#endif
constexpr int T2_R = T2_R_VAL;      // raw macro-tile ring
constexpr int T2_THREADS = 256 + 64;
// Two pipeline shapes, chosen per launch (measured at BASELINE C2 / C4, B200):
//   <SPM 1, S 3>  one 128-marker step per TMA box (16-byte rows), 3 A/B stages: short K loops (C2: 30.7 us vs 35.5)
//   <SPM 2, S 4>  two steps per box (32-byte rows: whole sectors), 4 stages: long K loops (C4: 0.86 ms vs 1.18)
// SPM = steps per raw macro tile, S = A (TMEM) / B (smem) stages.
#ifndef T2_LONG_SPM
#define T2_LONG_SPM 2
#endif
#ifndef T2_LONG_S
#define T2_LONG_S 4
#endif
constexpr int T2_LONG_K_STEPS = 256;  // steps per CTA from which the second shape is used

template <int S>
struct T2Bars {
    uint64_t raw_full[T2_R], raw_empty[T2_R];
    uint64_t a_full[S], a_empty[S], b_full[S];
    uint64_t done;
};

// smem: raw ring [T2_R][256 plane-rows][16 B], then B stages [T2_S][N/8][8 ki][8][16 B]
template <int SPM, int S>
__global__ void __launch_bounds__(T2_THREADS, 4)
    gebv_tc2_kernel(const __grid_constant__ CUtensorMap tmap, int64_t rows, const int8_t *__restrict__ bdig, int N, int T, int D,
                    int steps_total, int steps_per_split, unsigned long long *__restrict__ acc,
                    const double *__restrict__ inv_scale, float *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) T2Bars<S> bars;
    constexpr uint32_t T2_RAW_ROW = 16 * SPM;              // bytes per plane-row in a macro tile
    constexpr uint32_t T2_RAW_BYTES = 2 * T2_M * T2_RAW_ROW;  // 256 plane-rows
    __shared__ uint32_t tmem_base_slot;

    // warp index through a shuffle: provably warp-uniform, so the MMA warp's loop lives on the uniform datapath
    // (UTCIMMA back to back instead of an ELECT + R2UR sequence per instruction: ~70 -> ~10 cycles per MMA issued)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const uint32_t raw_base = smem_u32(smem);
    const uint32_t b_bytes = (uint32_t)N * T2_KS;
    const uint32_t b_base0 = raw_base + T2_R * T2_RAW_BYTES;
    const int64_t row0 = (int64_t)blockIdx.x * T2_M;
    const int s_begin = blockIdx.y * steps_per_split;
    const int nst = min(steps_total, s_begin + steps_per_split) - s_begin;

    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tmem_cols = 32;
    while (tmem_cols < d_cols + S * (T2_KS / 4)) tmem_cols <<= 1;

    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < T2_R; ++i) {
            mbar_init(smem_u32(&bars.raw_full[i]), 1);            // expect_tx arrival of the TMA loader
            mbar_init(smem_u32(&bars.raw_empty[i]), 4 * SPM);  // 4 reader warps per step
        }
        for (int i = 0; i < S; ++i) {
            mbar_init(smem_u32(&bars.a_full[i]), 4);   // the 4 warps of the group that filled the stage
            mbar_init(smem_u32(&bars.a_empty[i]), 1);  // tcgen05.commit
            mbar_init(smem_u32(&bars.b_full[i]), 1);   // expect_tx arrival of the loader
        }
        mbar_init(smem_u32(&bars.done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // Programmatic dependent launch: everything above (TMEM allocation, barrier init) may overlap the tail of the
    // kernel that produces this population; its memory is visible only past this point.  A no-op when the launch
    // did not ask for programmatic stream serialization.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t tmem_a = tmem_d + d_cols;

    if (warp < 8) {
        // ---------------- expanders: group g takes steps j with j % 2 == g ----------------
        const int g = warp >> 2, r = tid & (T2_M - 1);
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quadrant
        int pending = -1;  // A stage whose stores are issued but not yet published to the MMA warp
        for (int j = g; j < nst; j += 2) {
            const int mt = j / SPM, rs = mt % T2_R;
            mbar_wait(smem_u32(&bars.raw_full[rs]), (mt / T2_R) & 1);
            // macro tile [row][plane][16 B x SPM]; this step's 16 B of each plane
            const uint32_t src = raw_base + rs * T2_RAW_BYTES + (uint32_t)r * (2 * T2_RAW_ROW) + (uint32_t)(j % SPM) * 16;
            const uint4 x0 = lds128(src), x1 = lds128(src + T2_RAW_ROW);
            // cross-proxy WAR: these generic-proxy reads must be ordered before the async-proxy (TMA) write that refills
            // the slot.  Without the fence the 32-byte-row shape returned wrong sums in half of the runs at 1 M markers
            // (rows of the tile's first warp read the NEXT macro tile), the 16-byte-row shape in 1 of 24.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.raw_empty[rs]));
            // the field arithmetic happens BEFORE waiting for the TMEM stage, so it hides the MMA's latency
            const DosageFields f = dosage_fields(x0, x1);
            // the stores of this group's PREVIOUS step have had a whole iteration to land: publish them now instead of
            // stalling on tcgen05.wait::st right behind the stores (22 % of the kernel's stall samples at C4)
            if (pending >= 0) {
                tmem_st_publish();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars.a_full[pending]));
            }
            const int as = j % S;
            if (j >= S) mbar_wait(smem_u32(&bars.a_empty[as]), ((j / S) - 1) & 1);  // MMAs of the previous use retired
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            dosage_to_tmem_issue(tmem_a + lane_sel + (uint32_t)as * (T2_KS / 4), f);
            pending = as;
        }
        if (pending >= 0) {
            tmem_st_publish();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.a_full[pending]));
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ---------------- raw bit-plane tiles (TMA 2-D) ----------------
            const int y = (int)(2 * row0);  // plane-row coordinate
            const int nmt = (nst + SPM - 1) / SPM;
            for (int mt = 0; mt < nmt; ++mt) {
                const int rs = mt % T2_R;
                if (mt >= T2_R) mbar_wait(smem_u32(&bars.raw_empty[rs]), ((mt / T2_R) - 1) & 1);
                const uint32_t full = smem_u32(&bars.raw_full[rs]);
                mbar_arrive_expect_tx(full, T2_RAW_BYTES);  // out-of-bounds parts of the box are zero-filled and counted
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                        raw_base + rs * T2_RAW_BYTES),
                    "l"(reinterpret_cast<uint64_t>(&tmap)), "r"((s_begin + mt * SPM) * 4), "r"(y), "r"(full)
                    : "memory");
            }
        } else if (lane == 1) {
            // ---------------- digit tiles (1-D bulk copies) ----------------
            for (int j = 0; j < nst; ++j) {
                const int bs = j % S;
                if (j >= S) mbar_wait(smem_u32(&bars.a_empty[bs]), ((j / S) - 1) & 1);
                const uint32_t full = smem_u32(&bars.b_full[bs]);
                mbar_arrive_expect_tx(full, b_bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 b_base0 + bs * b_bytes),
                             "l"(bdig + (int64_t)(s_begin + j) * b_bytes), "r"(b_bytes), "r"(full)
                             : "memory");
            }
        }
    } else if (warp == 9) {
        // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues ----------------
        const uint32_t idesc = idesc_u8s8(N);
        for (int j = 0; j < nst; ++j) {
            const int as = j % S;
            const uint32_t par = (j / S) & 1;
            const uint32_t a_taddr = tmem_a + (uint32_t)as * (T2_KS / 4);
            const uint64_t bdesc = make_smem_desc(b_base0 + as * b_bytes, 128, 1024);
            mbar_wait(smem_u32(&bars.b_full[as]), par);
            mbar_wait(smem_u32(&bars.a_full[as]), par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int kk = 0; kk < T2_KS / 32; ++kk)  // +16 in the descriptor's address field = +256 bytes
                mma_i8_ts_warp(tmem_d, a_taddr + 8 * kk, bdesc + 16 * kk, idesc, (j > 0 || kk > 0) ? 1u : 0u);
            mma_commit_warp(smem_u32(&bars.a_empty[as]));
        }
        mma_commit_warp(smem_u32(&bars.done));
    }

    if (warp < 4) {
        mbar_wait(smem_u32(&bars.done), 0);
        digits_epilogue(tmem_d, tid, warp, row0, rows, T, acc, inv_scale, out, gridDim.y, D);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

}  // namespace

// zero-invariant scratch of the K-split kernels: accumulators [rows][T] (sum in the low 56 bits, arrival count on top)
int bg_tc_reserve_scratch(bg_engine *eng, int scratch, int64_t total, int64_t tiles, cudaStream_t st)
{
    if (eng->acc2_cap[scratch] < (size_t)total) {
        if (eng->d_acc2[scratch]) BG_CUDA(cudaFree(eng->d_acc2[scratch]));
        eng->d_acc2[scratch] = nullptr;
        eng->acc2_cap[scratch] = 0;
        BG_CUDA(cudaMalloc(&eng->d_acc2[scratch], (size_t)total * sizeof(unsigned long long)));
        BG_CUDA(cudaMemsetAsync(eng->d_acc2[scratch], 0, (size_t)total * sizeof(unsigned long long), st));
        eng->acc2_cap[scratch] = (size_t)total;
    }
    return BG_OK;
}

// K split: about `target` CTAs, >= 8 steps each, <= 3000 steps each, a multiple of `multiple` steps per split
void bg_tc_split(int64_t tiles, int steps, int64_t target, int multiple, int *ksplit_out, int *sps_out)
{
    int ksplit = (int)(target / tiles);
    const int max_split = (steps + 7) / 8;
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
    // int32 accumulators: a 128-marker step adds at most 16 * 43520 to a digit sum (prescaled bytes <= 128,
    // |digit| <= 128), so at most 3000 steps per CTA keeps every digit sum below 2^31
    if (ksplit < (steps + 2999) / 3000) ksplit = (steps + 2999) / 3000;
    if (ksplit > KSPLIT_MAX) ksplit = KSPLIT_MAX;  // 8-bit arrival count in the accumulators (tc_common.cuh)
    int sps = (steps + ksplit - 1) / ksplit;
    sps = (sps + multiple - 1) / multiple * multiple;
    *ksplit_out = (steps + sps - 1) / sps;
    *sps_out = sps;
}

// scratch: which of the two zero-invariant accumulator sets to use (two launches may be in flight on two streams)
int bg_launch_gebv_tc2(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, cudaStream_t st, int scratch)
{
    BG_REQUIRE(eng && eng->d_wdig, BG_ESTATE, "engine has no tensor-core digit table");
    const int T = eng->T, N = eng->tc_N, D = eng->tc_D;
    const int steps = (int)eng->tc_steps;
    const int64_t tiles = (rows + T2_M - 1) / T2_M;
    BG_REQUIRE(tiles < (int64_t(1) << 31) && 2 * rows < (int64_t(1) << 31), BG_ELIMIT, "too many rows");

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    EncodeTiledFn enc = encode_tiled();
    BG_REQUIRE(enc, BG_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    // 2-D view of the packed population: [2*rows plane-rows][Wpad words]; box = 4 words x 256 plane-rows
    const cuuint64_t gdim[2] = {(cuuint64_t)eng->Wpad, (cuuint64_t)(2 * rows)};
    const cuuint64_t gstride[1] = {(cuuint64_t)eng->Wpad * 4};
    // steps one CTA would run with the short-K shape's split (4 CTAs per SM, >= 8 steps each)
    int64_t ks = 4LL * eng->sm_count / std::max<int64_t>(tiles, 1);
    ks = std::max<int64_t>(1, std::min<int64_t>(ks, (steps + 7) / 8));
    const bool long_k = eng->opt.gebv_shape ? eng->opt.gebv_shape == 2 : steps / ks >= T2_LONG_K_STEPS;
    const int SPM = long_k ? T2_LONG_SPM : 1, S = long_k ? T2_LONG_S : 3;
    const cuuint32_t box[2] = {(cuuint32_t)(4 * SPM), 2 * T2_M};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t *>(pop), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BG_REQUIRE(cr == CUDA_SUCCESS, BG_ECUDA, "cuTensorMapEncodeTiled failed");

    const size_t smem = (size_t)T2_R * (2 * T2_M * 16 * SPM) + (size_t)S * N * T2_KS;
    BG_REQUIRE(smem <= (size_t)eng->max_smem_optin, BG_ELIMIT, "too many traits for the tensor-core GEBV tile");
    // residency: TMEM columns (512 per SM) and shared memory
    uint32_t d_cols = 32;
    while ((int)d_cols < N) d_cols <<= 1;
    uint32_t tcols = 32;
    while (tcols < d_cols + S * (T2_KS / 4)) tcols <<= 1;
    int resident = (int)(512 / tcols);
    const int by_smem = (int)(227 * 1024 / (smem + 1024));
    if (by_smem < resident) resident = by_smem;
    if (resident > 4) resident = 4;  // __launch_bounds__
    if (resident < 1) resident = 1;
    int ksplit, sps;
    bg_tc_split(tiles, steps, eng->opt.tc_target_ctas > 0 ? eng->opt.tc_target_ctas : (int64_t)resident * eng->sm_count /* one full wave */, 1,
                &ksplit, &sps);
    int rc = bg_tc_reserve_scratch(eng, scratch, rows * T, tiles, st);
    if (rc) return rc;
    dim3 grid((unsigned)tiles, (unsigned)ksplit);
    // largest dynamic smem opted into so far, per engine (= per device: the attribute is per device)
    auto kern = long_k ? gebv_tc2_kernel<T2_LONG_SPM, T2_LONG_S> : gebv_tc2_kernel<1, 3>;
    size_t &optin = long_k ? eng->tc2_optin[2] : eng->tc2_optin[0];
    if (smem > optin) {
        BG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin = smem;
    }
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
    BG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, rows, bd, N, T, D, steps, sps, acc, inv, out));
    BG_LAUNCHED();
    return BG_OK;
}

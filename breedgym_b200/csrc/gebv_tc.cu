// GEBV on the 5th-generation tensor cores (tcgen05, sm_100a): an exact int8 GEMM.
//
// Replaces chromax TraitModel.__call__ = dot(sum(pop,-1), effects[m,T]) (called from
// breedgym/breedgym.py:233 and breedgym/vector/vec_env.py:132-134) for any number of
// traits.  D[i, 8t+d] = sum_j dosage[i,j] * digit_d(w_fix[j,t]) where w_fix is the 64-bit
// fixed-point effect of gebv.cu split into 8 balanced base-256 digits (int8), dosage in
// {0,1,2} (int8) and the accumulation int32 in tensor memory: every product and sum is an
// exact integer, so recombining the digits reproduces the SAME int64 as the CUDA-core
// kernels, bit for bit, for T = 1 .. 32 traits at no extra ALU cost per trait.
//
// One CTA owns 128 individuals x a contiguous range of 256-marker K chunks and runs a
// warp-specialised pipeline:
//   warps 0-7 (producers): cp.async ring (depth 4) streams the two bit planes of "their" row
//       into shared memory; each thread expands 4 words (128 markers) per chunk into dosage
//       bytes with shift/mask/add (K is permuted inside each 32-marker word; the digit matrix
//       is permuted identically on the host) and stores them, 128 bits at a time and
//       bank-conflict free, in the canonical no-swizzle K-major core-matrix layout; a
//       per-warp mbarrier arrival marks the stage full.
//   warp 8 (one elected lane): TMA bulk copy (cp.async.bulk) of the chunk's digit tile, then
//       8 x tcgen05.mma.kind::i8 (M=128, N=8T padded to 16, K=32), accumulator in TMEM;
//       tcgen05.commit releases the stage to the producers.
//   warps 0-3 (epilogue): tcgen05.ld (row <-> TMEM lane), digits -> int64, partial sums out.
// K-split partials are summed by a tiny finalize kernel (deterministic, no atomics).
#include "bg_internal.h"

namespace {

constexpr int TC_M = 128;
constexpr int TC_KC = 256;         // markers (= int8 K elements) per chunk: 8 words per plane
constexpr int TC_PRODUCERS = 256;  // 8 warps: thread <-> (row, 128-marker half of the chunk)
constexpr int TC_THREADS = TC_PRODUCERS + 32;
constexpr int TC_STAGES = 2;       // operand (A/B) stages
constexpr int TC_RING = 4;         // cp.async prefetch depth (chunks)
constexpr uint32_t TC_SPIN_LIMIT = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: core matrix = 8 rows x 16 bytes stored
// contiguously; LBO = byte stride between core matrices adjacent in K, SBO = adjacent in M/N.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;  // descriptor version (sm_100)
    return d;         // base offset 0, layout type 0 (no swizzle)
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
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > TC_SPIN_LIMIT) __trap();  // never hang the GPU on a lost arrival
    }
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// 16-byte async copy global -> shared; src_bytes = 0 zero-fills (rows / words outside the population)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// smem: TC_STAGES x { A [16 mi][16 ki][8 rows][16 B] = 32 KB, B [N/8 ni][16 ki][8][16 B] = N*256 B },
//       (B = two consecutive step tiles [N/8][8 ki][8][16 B]), then the cp.async ring [TC_RING][2 planes][256 threads][16 B]
__global__ void __launch_bounds__(TC_THREADS) gebv_tc_kernel(const uint32_t *__restrict__ pop, int64_t rows, int Wpad,
                                                             const int8_t *__restrict__ bdig, int N, int T,
                                                             int chunks_total, int chunks_per_split,
                                                             long long *__restrict__ partial)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_done;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t stage_bytes = (uint32_t)(TC_M * TC_KC + N * TC_KC);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t ring_base = smem_base + TC_STAGES * stage_bytes;
    const int64_t row0 = (int64_t)blockIdx.x * TC_M;
    const int c_begin = blockIdx.y * chunks_per_split;
    const int c_end = min(chunks_total, c_begin + chunks_per_split);
    const int nch = c_end - c_begin;

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < N) tmem_cols <<= 1;

    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 8 + 1);  // 8 producer warps + the digit-tile expect_tx arrival
            mbar_init(smem_u32(&bar_empty[s]), 1);     // tcgen05.commit
        }
        mbar_init(smem_u32(&bar_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp < 8) {
        // ---------------- producers ----------------
        const int r = tid & (TC_M - 1), half = tid >> 7;
        const int64_t row = row0 + r;
        const bool row_ok = row < rows;
        const uint32_t *h0 = pop + (row_ok ? row : 0) * 2 * (int64_t)Wpad;
        const uint32_t *h1 = h0 + Wpad;
        const uint32_t a_row_off = (uint32_t)((r >> 3) * 2048 + (r & 7) * 16);
        const uint32_t my_ring = ring_base + (uint32_t)tid * 16;

        auto issue = [&](int f) {  // chunk c_begin + f -> ring slot f % TC_RING; always commits a group
            if (f < nch) {
                const int w = (c_begin + f) * 8 + 4 * half;
                const bool ok = row_ok && w < Wpad;  // Wpad % 4 == 0: a uint4 is entirely inside or outside
                const uint32_t dst = my_ring + (uint32_t)(f % TC_RING) * (2 * TC_PRODUCERS * 16);
                cp_async16(dst, h0 + (ok ? w : 0), ok ? 16u : 0u);
                cp_async16(dst + TC_PRODUCERS * 16, h1 + (ok ? w : 0), ok ? 16u : 0u);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int f = 0; f < TC_RING; ++f) issue(f);

        for (int it = 0; it < nch; ++it) {
            const int s = it % TC_STAGES, use = it / TC_STAGES;
            asm volatile("cp.async.wait_group %0;" ::"n"(TC_RING - 1) : "memory");
            const uint32_t src = my_ring + (uint32_t)(it % TC_RING) * (2 * TC_PRODUCERS * 16);
            const uint4 x0 = lds128(src), x1 = lds128(src + TC_PRODUCERS * 16);
            if (it >= TC_STAGES) mbar_wait(smem_u32(&bar_empty[s]), (use - 1) & 1);  // MMAs of the previous use are done

            const uint32_t a_base = smem_base + s * stage_bytes;
            const uint32_t w0[4] = {x0.x, x0.y, x0.z, x0.w}, w1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                uint32_t o[8];
#pragma unroll
                for (int sft = 0; sft < 8; ++sft)  // byte b of o[sft] <-> marker 8b + sft of this word
                    o[sft] = (((w0[jj] >> sft) & 0x01010101u) + ((w1[jj] >> sft) & 0x01010101u)) << (sft & ~1);  // prescaled: dosage * 4^(sft/2)
                const uint32_t dst = a_base + a_row_off + (uint32_t)(2 * (4 * half + jj)) * 128;
                sts128(dst, o[0], o[1], o[2], o[3]);
                sts128(dst + 128, o[4], o[5], o[6], o[7]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
            issue(it + TC_RING);  // the ring slot consumed above is free again (its words are in registers no more)
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (lane == 0) {
        // ---------------- MMA issuer ----------------
        // instruction descriptor: D = S32, A = B = signed 8-bit, both K-major, N, M = 128
        // A = unsigned 8-bit (prescaled dosage bytes reach 128), B = signed 8-bit digits, D = int32
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        const uint32_t b_bytes = (uint32_t)(N * TC_KC);
        for (int it = 0; it < nch; ++it) {
            const int s = it % TC_STAGES, use = it / TC_STAGES;
            const uint32_t a_base = smem_base + s * stage_bytes, b_base = a_base + TC_M * TC_KC;
            const uint32_t full = smem_u32(&bar_full[s]);
            if (it >= TC_STAGES) mbar_wait(smem_u32(&bar_empty[s]), (use - 1) & 1);
            // digit tile of this chunk: one contiguous bulk copy (already in core-matrix order)
            mbar_arrive_expect_tx(full, b_bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(b_base),
                         "l"(bdig + (int64_t)(c_begin + it) * b_bytes), "r"(b_bytes), "r"(full)
                         : "memory");
            mbar_wait(full, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kk = 0; kk < TC_KC / 32; ++kk) {
                const uint64_t adesc = make_smem_desc(a_base + kk * 256, 128, 2048);
                // the chunk's digit tile = two consecutive 128-marker step tiles [N/8][8][8][16 B]
                const uint64_t bdesc = make_smem_desc(b_base + (kk >> 2) * (b_bytes >> 1) + (kk & 3) * 256, 128, 1024);
                const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_empty[s]))
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done))
                     : "memory");
    }

    if (warp < 4) {  // epilogue: thread t <-> accumulator row t <-> TMEM lane t
        mbar_wait(smem_u32(&bar_done), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t row = row0 + tid;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        long long *dst = partial + ((int64_t)blockIdx.y * rows + row) * T;
        for (int t = 0; t < T; ++t) {
            uint32_t v[8];
            if (nch > 0) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr + (uint32_t)(8 * t))
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int d = 0; d < 8; ++d) v[d] = 0;
            }
            unsigned long long sum = 0;  // modular arithmetic: the true total fits in int64
#pragma unroll
            for (int d = 7; d >= 0; --d) sum = (sum << 8) + (unsigned long long)(long long)(int32_t)v[d];
            if (row < rows) dst[t] = (long long)sum >> 6;  // the prescaled operand makes every sum exactly 64x (api.cu)
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

__global__ void gebv_tc_finalize_kernel(const long long *__restrict__ partial, int ksplit, int64_t total,
                                        const double *__restrict__ inv_scale, int T, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long s = 0;
    for (int k = 0; k < ksplit; ++k) s += partial[(int64_t)k * total + i];
    out[i] = (float)((double)s * inv_scale[i % T]);
}

}  // namespace

int bg_gebv_tc_max_traits(void) { return 32; }

int bg_launch_gebv_tc(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, cudaStream_t st)
{
    BG_REQUIRE(eng && eng->d_wdig, BG_ESTATE, "engine has no tensor-core digit table");
    const int T = eng->T, N = eng->tc_N;
    const int chunks = (int)((eng->tc_steps + 1) / 2);
    const int64_t tiles = (rows + TC_M - 1) / TC_M;
    BG_REQUIRE(tiles < (int64_t(1) << 31), BG_ELIMIT, "too many rows");
    const size_t smem = (size_t)TC_STAGES * (TC_M * TC_KC + N * TC_KC) + (size_t)TC_RING * 2 * TC_PRODUCERS * 16;
    BG_REQUIRE(smem <= (size_t)eng->max_smem_optin, BG_ELIMIT, "too many traits for the tensor-core GEBV tile");
    // K split: about `target` CTAs so every SM holds its resident set a few times over, >= 8 chunks each
    int resident = (int)(227 * 1024 / (smem + 2048));
    if (resident < 1) resident = 1;
    int64_t target = 2LL * resident * eng->sm_count;
    if (const char *s = getenv("BG_TC_TARGET_CTAS")) target = atoll(s) > 0 ? atoll(s) : target;
    int ksplit = (int)((target + tiles - 1) / tiles);
    const int max_split = (chunks + 7) / 8;
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
    // int32 accumulators: a 256-marker chunk adds at most 2 * 16 * 43520 to a digit sum (prescaled bytes <= 128,
    // |digit| <= 128), so at most 1500 chunks per CTA keeps every digit sum below 2^31
    if (ksplit < (chunks + 1499) / 1500) ksplit = (chunks + 1499) / 1500;
    if (ksplit > 65535) ksplit = 65535;
    int cps = (chunks + ksplit - 1) / ksplit;
    ksplit = (chunks + cps - 1) / cps;
    const int64_t total = rows * T;
    int rc = bg_reserve_acc(eng, (size_t)total * ksplit);
    if (rc) return rc;
    BG_CUDA(cudaFuncSetAttribute(gebv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)tiles, (unsigned)ksplit);
    gebv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(pop, rows, eng->Wpad, eng->d_wdig, N, T, chunks, cps,
                                                   reinterpret_cast<long long *>(eng->d_acc));
    BG_LAUNCHED();
    gebv_tc_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const long long *>(eng->d_acc), ksplit,
                                                                            total, eng->d_inv_scale, T, out);
    BG_LAUNCHED();
    return BG_OK;
}

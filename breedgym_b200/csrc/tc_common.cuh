// Device helpers shared by the tcgen05 GEBV kernels (gebv_tc2.cu, cross_gebv.cu): mbarrier / TMEM / descriptor
// wrappers, the bit-plane -> prescaled dosage byte expansion into tensor memory, and the int8-digit epilogue.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace bgtc {

constexpr int TILE_M = 128;     // individuals per tile (TMEM lanes)
constexpr int STEP_K = 128;     // markers per step (4 words per plane, 32 TMEM columns of int8x4)
#ifndef BG_SPIN_LIMIT
#define BG_SPIN_LIMIT (1u << 28)   // mbarrier waits trap after this many try_wait rounds (diagnostic builds lower it)
#endif
constexpr uint32_t SPIN_LIMIT = BG_SPIN_LIMIT;

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
#ifndef BG_MBAR_HINT_NS
#define BG_MBAR_HINT_NS 2000   // suspend-time hint of mbarrier.try_wait (0: plain try_wait spin)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
#if BG_MBAR_HINT_NS > 0
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"  // sleeps up to %3 ns unless the phase completes
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"((uint32_t)BG_MBAR_HINT_NS)
            : "memory");
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
#endif
        if (spin > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&o)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(o[0]), "r"(o[1]),
                 "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                 : "memory");
}

// instruction descriptor of tcgen05.mma.kind::i8: A = unsigned 8-bit (prescaled dosage bytes reach 128),
// B = signed 8-bit digits, D = int32, M = 128, N = n
__device__ __forceinline__ uint32_t idesc_u8s8(int n)
{
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// 2-bit dosage fields of 4 words per plane (even / odd markers: field f of ze <-> marker 2f, of zo <-> 2f+1;
// a field holds 0..2, no carry)
struct DosageFields {
    uint32_t ze[4], zo[4];
};
__device__ __forceinline__ DosageFields dosage_fields(const uint4 x0, const uint4 x1)
{
    const uint32_t w0[4] = {x0.x, x0.y, x0.z, x0.w}, w1[4] = {x1.x, x1.y, x1.z, x1.w};
    DosageFields f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        f.ze[jj] = (w0[jj] & 0x55555555u) + (w1[jj] & 0x55555555u);
        f.zo[jj] = ((w0[jj] >> 1) & 0x55555555u) + ((w1[jj] >> 1) & 0x55555555u);
    }
    return f;
}

// 128 markers of this thread's individual -> 32 TMEM columns of its lane.  One mask keeps 4 fields where they
// sit in their bytes: column q, byte b <-> K index 4q + b <-> marker 8b + q of the word, as the UNSIGNED byte
// dosage * 4^(q/2); the digit table carries the inverse scale (api.cu), so all sums are exactly 64x.
// 16 integer ops per 32 markers in total.
// issue only: the stores are asynchronous; tmem_st_publish() makes them visible to the MMA warp's barrier round
__device__ __forceinline__ void dosage_to_tmem_issue(uint32_t taddr, const DosageFields &f)
{
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = ((q & 1) ? f.zo[jj] : f.ze[jj]) & (0x03030303u << (q & ~1));
        tmem_st8(taddr + 8 * jj, o);
    }
}
__device__ __forceinline__ void tmem_st_publish()
{
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void dosage_to_tmem(uint32_t taddr, const DosageFields &f)
{
    dosage_to_tmem_issue(taddr, f);
    tmem_st_publish();
}

__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t a_taddr, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// warp-uniform variants: called by ALL lanes of a converged warp with identical arguments, one elected lane issues.
// (Issuing from inside an `if (lane == 0)` region makes ptxas wrap every UTCIMMA in an ELECT / BRA.U.ANY loop and
// re-derive its uniform-register operands: ~100 cycles per MMA, 700 per 128-marker step.)
__device__ __forceinline__ void mma_i8_ts_warp(uint32_t tmem_d, uint32_t a_taddr, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_warp(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Epilogue of warps 0-3 (thread t <-> accumulator row t <-> TMEM lane t): digits -> int64 (exactly 64x the
// fixed-point sum, shifted back).  K-split partials meet in ONE 64-bit atomic per value: the high 56 bits carry the sum
// (two's complement, |sum| < 2^55 by construction of the fixed-point scale, api.cu), the low 8 bits count arrivals
// (<= 255, so they never carry into the sum), so the CTA whose atomicAdd returns nsplit - 1 arrivals holds the
// complete sum in the returned value, converts it to float32 and puts
// the zero back (the accumulators are all zero between launches): one L2 round trip, no fence, no barrier, no
// finalize launch.  (The earlier protocol -- partial atomics, __threadfence, tile counter, last CTA re-reads -- cost
// three dependent round trips: ~4300 of the ~36000 cycles a CTA of the fused kernel lives.)
constexpr int KSPLIT_MAX = 255;
__device__ __forceinline__ void digits_epilogue(uint32_t tmem_d, int tid, int warp, int64_t row0, int64_t rows, int T,
                                                unsigned long long *__restrict__ acc, const double *__restrict__ inv_scale,
                                                float *__restrict__ out, unsigned nsplit, int D,
                                                int64_t out_row = -2)
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t row = row0 + tid;  // accumulator slot; the result goes to row `orow` of `out` (-2: the same row)
    const int64_t orow = out_row == -2 ? row : out_row;
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int t = 0; t < T; ++t) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr + (uint32_t)(D * t))  // columns D*t .. D*t+7 (those past D belong to the next trait: ignored)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        unsigned long long sum = 0;  // modular arithmetic: the true total fits in int64
#pragma unroll
        for (int d = 7; d >= 0; --d)
            if (d < D) sum = (sum << 8) + (unsigned long long)(long long)(int32_t)v[d];
        long long total = (long long)sum >> 6;  // prescaled operand: every (partial) sum is exactly 64x
        if (row < rows) {
            bool complete = nsplit == 1;  // no K split: this CTA holds the whole sum
            if (!complete) {
                const unsigned long long mine = ((unsigned long long)total << 8) + 1ull;
                const unsigned long long old = atomicAdd(acc + row * T + t, mine);
                if ((unsigned)(old & 0xFFull) == nsplit - 1) {  // the last of the nsplit partial sums
                    total = (long long)(old + mine) >> 8;  // arithmetic shift: the 56-bit sum, sign-extended
                    acc[row * T + t] = 0ull;
                    complete = true;
                }
            }
            if (complete) out[orow * T + t] = (float)((double)total * inv_scale[t]);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled()
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

}  // namespace bgtc

// Measured INT32 issue rate of this GPU, the roofline of the Threefry-bound kernels (meiosis.cu):
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/int32_peak scripts/int32_peak.cu && scripts/int32_peak
//
// Every thread runs CHAINS independent dependency chains of one instruction kind, 8192 deep, at full occupancy
// (2048 threads per SM); ops/s = threads x chains x depth x ops per link / time, best of 5.  Kinds:
//   iadd    x += y; y += x              (IADD3)
//   lop3    x = (x ^ y) & z | w-ish     (LOP3, ALU pipe)
//   shf     x = funnelshift(x, x, r)    (SHF, ALU pipe)
//   imad    x = y * one + x; y = x * one + y   (IMAD, FMA pipe: `one` is a kernel argument, so it is not folded to IADD3)
//   tf_alu  the Threefry round with the add as IADD3:  x0 += x1; x1 = rotl(x1, r) ^ x0        (3 ops, all ALU)
//   tf_mix  the Threefry round as meiosis.cu issues it: x0 = x1 * one + x0 (IMAD) ; SHF ; LOP3 (3 ops, two pipes)
// Output: one JSON object; `int32_gops` = the tf_mix rate (the most the rounds of a Threefry block can issue at),
// `threefry_int_ops_per_draw` = 37.5: a 20-round block is 60 round ops + 15 key-injection adds and yields 2 draws in the
// legacy jax layout (the compare / ballot / scan work on top is the kernel's own overhead, not credited).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

constexpr int DEPTH = 8192;

template <int KIND, int CHAINS>
__global__ void __launch_bounds__(1024, 2) probe(unsigned *out, unsigned one, unsigned seed)
{
    unsigned x0[CHAINS], x1[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x0[c] = seed + threadIdx.x * 31u + c;
        x1[c] = seed * 7u + blockIdx.x + c * 977u;
    }
#pragma unroll 1
    for (int i = 0; i < DEPTH / 8; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (KIND == 0) {  // two dependent adds per link (a lone x0 += x1 with constant x1 folds into one multiply-add)
                    x0[c] += x1[c];
                    x1[c] += x0[c];
                }
                if (KIND == 1) x0[c] = (x0[c] ^ x1[c]) | (x0[c] & seed);
                if (KIND == 2) x0[c] = __funnelshift_l(x0[c], x0[c], 13);
                if (KIND == 3) {
                    x0[c] = x1[c] * one + x0[c];
                    x1[c] = x0[c] * one + x1[c];
                }
                if (KIND == 4) {
                    x0[c] += x1[c];
                    x1[c] = __funnelshift_l(x1[c], x1[c], 13 + u) ^ x0[c];
                }
                if (KIND == 5) {
                    x0[c] = x1[c] * one + x0[c];
                    x1[c] = __funnelshift_l(x1[c], x1[c], 13 + u) ^ x0[c];
                }
            }
        }
    }
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x0[c] ^ x1[c];
    if (acc == 0x12345678u) out[0] = acc;  // never true in practice: keeps the chains alive
}

template <int KIND, int CHAINS>
double run(int sms, unsigned *d_out)
{
    const int blocks = sms * 2 * 4;
    const double ops_per_link = KIND >= 4 ? 3.0 : ((KIND == 0 || KIND == 3) ? 2.0 : 1.0);
    float best = 1e30f;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        probe<KIND, CHAINS><<<blocks, 1024>>>(d_out, 1u, 12345u + rep);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    return (double)blocks * 1024 * CHAINS * DEPTH * ops_per_link / (best * 1e-3) / 1e9;  // Gops
}

int main()
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    unsigned *d_out;
    cudaMalloc(&d_out, 4);
    // warm the clocks up
    for (int i = 0; i < 20; ++i) run<0, 4>(sms, d_out);
    const double iadd = run<0, 8>(sms, d_out), lop3 = run<1, 8>(sms, d_out), shf = run<2, 8>(sms, d_out), imad = run<3, 8>(sms, d_out);
    const double tf_alu4 = run<4, 4>(sms, d_out), tf_mix4 = run<5, 4>(sms, d_out);
    const double tf_alu8 = run<4, 8>(sms, d_out), tf_mix8 = run<5, 8>(sms, d_out);
    const double tf_alu = tf_alu4 > tf_alu8 ? tf_alu4 : tf_alu8, tf_mix = tf_mix4 > tf_mix8 ? tf_mix4 : tf_mix8;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    const double per_clk = 1e9 / ((double)sms * khz * 1e3);
    printf("{\"sms\": %d, \"sm_clock_mhz\": %.0f, \"iadd3_gops\": %.0f, \"lop3_gops\": %.0f, \"shf_gops\": %.0f, \"imad_gops\": %.0f, "
           "\"threefry_round_alu_only_gops\": %.0f, \"threefry_round_imad_shf_lop3_gops\": %.0f, "
           "\"iadd3_per_sm_clk\": %.1f, \"lop3_per_sm_clk\": %.1f, \"shf_per_sm_clk\": %.1f, \"imad_per_sm_clk\": %.1f, "
           "\"threefry_mix_per_sm_clk\": %.1f, \"int32_gops\": %.0f, \"threefry_int_ops_per_draw\": 37.5, "
           "\"how\": \"scripts/int32_peak.cu: 8 (4) independent chains per thread, 8192 links, 2048 threads per SM, best of 5; "
           "int32_gops = the Threefry round issued as IMAD + SHF + LOP3 (two pipes), the form meiosis.cu uses\"}\n",
           sms, khz / 1e3, iadd, lop3, shf, imad, tf_alu, tf_mix, iadd * per_clk, lop3 * per_clk, shf * per_clk, imad * per_clk,
           tf_mix * per_clk, tf_mix > tf_alu ? tf_mix : tf_alu);
    return 0;
}

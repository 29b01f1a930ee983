// Internal declarations shared by the translation units of libbreedgym_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/breedgym_b200.h"

// Crossover masks of up to BG_BATCH_MAX CONSECUTIVE cross keys of the simulator's key chain (vector env), generated
// by ONE launch of the mask kernel.  Masks depend on the key chain only, so while the steps of one batch run, the
// batch that continues the chain is generated on a side stream; three buffers rotate (in use / being generated / its
// readers long gone).  A lookup is by key, so a reseed or an explicit-key call simply misses and regenerates.
constexpr int BG_BATCH_MAX = 8;
struct bg_mask_batch {
    uint32_t *mask = nullptr, *mut = nullptr;   // [count][rows][Wpad]
    size_t cap = 0, mut_cap = 0;                // words
    bool valid = false;
    int count = 0;
    uint32_t keys[BG_BATCH_MAX][2] = {};
    bool has_state = false;                     // state_after = the chain state after keys[count-1] was drawn
    uint32_t state_after[2] = {0, 0};
    int layout = -1, schedule = -1;
    int64_t rows = 0;
    cudaEvent_t ready = nullptr;                // recorded after the generating kernel
    cudaEvent_t freed = nullptr;                // recorded behind the batch's readers when their stream changes / before a refill
    bool freed_set = false;
    cudaStream_t gen_stream = nullptr;          // stream the generating kernel ran on
    cudaStream_t synced_stream = nullptr;       // stream that has already been made to wait for `ready`
    cudaStream_t use_stream = nullptr;          // stream of the readers since the last `freed` record
    bool used = false;
    unsigned long long stamp = 0;               // last use (LRU choice of the buffer to refill)
    cudaEvent_t t0 = nullptr, t1 = nullptr;     // option timing: around the generating kernel
    bool timed = false;
};

constexpr int BG_MASK_BATCHES = 3;

// behaviour switches, read ONCE (bg_engine_create: environment; bg_engine_set_option afterwards) -- nothing on the
// step path calls getenv
struct bg_options {
    int fuse = 1;              // 0: blend + GEBV kernels instead of the fused step kernel (cross-checks)
    int gebv_algo = 0;         // 0 auto, 1..3 as bg_gebv_algo
    int lookahead = BG_BATCH_MAX;  // steps of masks generated ahead on the side stream (0: none)
    int mask_nt = 128;         // threads of the small mask CTAs that run beside the step kernel
    int rows_nt = 0;           // threads per CTA of the unique-key cross / mask kernel (0: by row length; 128..1024)
    int mask_big_ctas = 0;     // diagnostics: full-size mask CTAs on the side stream
    int mask_ctas_per_sm = 0;  // lookahead batches: 0 = one small CTA per row; k > 0 = persistent grid of k small CTAs per SM
    int blend_env_chunk = 8;
    int copy_engine = 0;       // 1: never use the mapped-memory copy kernels
    long long mapped_d2h_max = 32 * 1024;
    long long tc_target_ctas = 0;  // 0: default
    int timing = 0;
    int gebv_shape = 0;        // tensor-core GEBV pipeline shape: 0 auto, 1 short-K <SPM 1, S 3>, 2 long-K <SPM 2, S 4>
    int gebv_digits = 0;       // 0: as many base-256 digits as the map needs; 8: always 8 (cross-checks)
    int step_pdl = 1;          // programmatic dependent launch of the step kernels (their prologue overlaps the previous step's tail)
    int fused_dyn = -1;        // the persistent fused step kernel with a dynamic work queue (cross_gebv_dyn.cu): 1 always, 0 never,
                               // -1 when a launch has >= 4 tiles per resident CTA (measured: 256 vs 275 us at 512 envs, 43 vs 40 us at 64)
    long long xg_parts = 0;    // its split of a tile's K range, proportions as decimal digits (8642 = 8:6:4:2); 0: auto
};

struct bg_engine {
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 0;
    // map constants
    int64_t m = 0;       // markers
    int32_t W = 0;       // ceil(m/32)
    int32_t Wpad = 0;    // W rounded up to a multiple of 32 (128-byte rows)
    int32_t T = 0;       // traits
    uint32_t *d_thr = nullptr;        // [Wpad*32 + 32] recombination thresholds (zero padded)
    uint32_t *d_thr_cmp = nullptr;    // thresholds << 9 for the one-compare fast path (NULL when some threshold is 2^23)
    uint32_t mut_thr = 0;             // mutation threshold (0 = off)
    long long *d_wfix = nullptr;      // [T][Wpad*32] fixed-point effects (zero padded)
    double *d_inv_scale = nullptr;    // [T] 2^-s_t
    signed char *d_wdig = nullptr;    // [tc_steps][tc_N/8][8][8][16] int8 base-256 digits in core-matrix order
    int64_t tc_steps = 0;             // 128-marker K steps (= Wpad / 4)
    int32_t tc_D = 0;                 // base-256 digits per effect (4..8, chosen per map by bg_engine_set_map)
    int32_t tc_N = 0;                 // tc_D*T rounded up to a multiple of 16 (0: tensor-core path unavailable)
    // grow-only scratch
    bg_options opt;
    bg_mask_batch batches[BG_MASK_BATCHES];
    unsigned long long use_clock = 0;
    cudaStream_t side = nullptr;        // lookahead stream
    cudaEvent_t tmp_event = nullptr;
    cudaEvent_t reset_ready = nullptr;  // recorded behind a prefetched reset on the side stream
    cudaEvent_t join_event = nullptr;   // bg_engine_join
    bool reset_pending = false;
    uint32_t *d_mut = nullptr;          // mutation scratch of bg_meiosis_masks
    size_t mut_cap = 0;                 // words
    unsigned long long *d_acc = nullptr;
    size_t acc_cap = 0;                 // elements
    unsigned long long *d_acc2[2] = {nullptr, nullptr};  // gebv_tc2: all-zero between launches (one set per stream)
    size_t acc2_cap[2] = {0, 0};
    struct bg_peer *peer = nullptr;     // reward exchange over peer memory (peer.cu): bg_vec_step publishes its rewards through it
    unsigned int *d_xg_work = nullptr;  // work counters of the persistent fused step kernel (cross_gebv_dyn.cu)
    unsigned xg_seq = 0;
    size_t tc2_optin[4] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};  // dynamic smem already opted into (GEBV short-K, fused kernel, GEBV long-K, persistent fused kernel)
};

void bg_set_error(const std::string &msg);
int bg_cuda_fail(cudaError_t e, const char *what);

#define BG_CUDA(call)                                       \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return bg_cuda_fail(e__, #call); \
    } while (0)

// after every kernel launch: count it (bg_kernel_launches) and surface launch errors
extern std::atomic<long long> bg_launch_counter;
#define BG_LAUNCHED()                                     \
    do {                                                  \
        bg_launch_counter.fetch_add(1, std::memory_order_relaxed); \
        BG_CUDA(cudaGetLastError());                      \
    } while (0)

#define BG_REQUIRE(cond, code, msg) \
    do {                            \
        if (!(cond)) {              \
            bg_set_error(msg);      \
            return (code);          \
        }                           \
    } while (0)

int bg_reserve_u32(uint32_t **p, size_t *cap, size_t words);
int bg_reserve_acc(bg_engine *eng, size_t elems);

// meiosis.cu
enum { BG_ROWS_MASK = 0, BG_ROWS_CROSS = 1, BG_ROWS_DH = 2 };
int bg_launch_meiosis_rows(bg_engine *eng, int mode, int64_t rows, const uint32_t cross_key[2], int layout, int schedule,
                           uint32_t *mask_out, uint32_t *mut_out, const uint32_t *pop, const int32_t *parents,
                           int64_t n_src, int64_t dh_offspring, uint32_t *out, cudaStream_t st, int small_ctas = 0);
int bg_launch_double_haploid(bg_engine *eng, int64_t E, int64_t n, int64_t n_offspring, const uint32_t cross_key[2], int layout,
                             int schedule, const uint32_t *pop, uint32_t *out, cudaStream_t st);
int bg_launch_mask_batch(bg_engine *eng, int64_t rows, int nkeys, const uint32_t (*keys)[2], int layout, int schedule,
                         uint32_t *mask_out, uint32_t *mut_out, cudaStream_t st, int small_ctas);
int bg_launch_blend(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask,
                    const uint32_t *mut, uint32_t *out, int64_t E, int64_t n_src, int64_t n, cudaStream_t st);

// gebv.cu
int bg_launch_gebv(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, int algo, cudaStream_t st);
int bg_launch_reduce(const float *in, int64_t E, int64_t per_env, float *out, int op, cudaStream_t st);

// gebv_tc2.cu: TMA tile loads + operand A in tensor memory
int bg_launch_gebv_tc2(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, cudaStream_t st, int scratch = 0);
bool bg_cross_gebv_fused_ok(const bg_engine *eng, int64_t E, int64_t n_src, int64_t n);
bool bg_cross_gebv_dyn_ok(const bg_engine *eng, int64_t E, int64_t n_src, int64_t n);
int bg_launch_cross_gebv_dyn(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask, uint32_t *out_pop,
                             int64_t E, int64_t n_src, int64_t n, float *gebv_out, cudaStream_t st);
int bg_launch_cross_gebv_fused(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask, uint32_t *out_pop,
                               int64_t E, int64_t n_src, int64_t n, float *gebv_out, cudaStream_t st);

// peer.cu
int bg_launch_reduce_publish(const float *in, int64_t E, int64_t per_env, float *out, bg_peer *peer, cudaStream_t st);
void bg_peer_engine_gone(bg_engine *eng);

// topk.cu
int bg_launch_topk(const float *scores, int64_t rows, int64_t len, int k, float *vals_out, int32_t *idx_out, cudaStream_t st);

// pairs.cu
int bg_launch_pairs_from_topk(const float *vals, const int32_t *idx, int64_t E, int k, int64_t row_len, int32_t *out, cudaStream_t st);
int bg_launch_diallel_pairs(const int32_t *best, const int32_t *perm, int64_t E, int k, int nc, int64_t n, int32_t *out, cudaStream_t st);

// layout.cu
int bg_launch_copy_mapped(const void *src, void *dst, size_t bytes, cudaStream_t st);
int bg_launch_pack(const uint8_t *in, uint32_t *out, int64_t rows, int64_t m, int W, int Wpad, cudaStream_t st);
int bg_launch_unpack(const uint32_t *in, uint8_t *out, int64_t rows, int64_t m, int Wpad, cudaStream_t st);
int bg_launch_gather(const uint32_t *src, const int32_t *idx, uint32_t *dst, int64_t E, int64_t n_src, int64_t n,
                     int64_t src_env_rows, int Wpad, cudaStream_t st, const float *src_vals = nullptr, float *dst_vals = nullptr,
                     int T = 0);
int bg_launch_reset_indices(bg_engine *eng, const uint32_t key[2], int64_t E_total, int64_t env_begin, int64_t E, int64_t N,
                            int64_t n, int layout, int32_t *idx_out, cudaStream_t st);

// C-ABI entry points of libbreedgym_b200 (see include/breedgym_b200.h).
#include <math.h>
#include <string.h>

#include <chrono>
#include <cstdio>
#include <vector>

#include "bg_internal.h"
#include "threefry.cuh"

static thread_local std::string g_err;
std::atomic<long long> bg_launch_counter{0};

void bg_set_error(const std::string &msg) { g_err = msg; }

int bg_cuda_fail(cudaError_t e, const char *what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? BG_ENOMEM : BG_ECUDA;
}

int bg_reserve_u32(uint32_t **p, size_t *cap, size_t words)
{
    if (*cap >= words) return BG_OK;
    if (*p) BG_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    BG_CUDA(cudaMalloc(p, words * sizeof(uint32_t)));
    *cap = words;
    return BG_OK;
}

int bg_reserve_acc(bg_engine *eng, size_t elems)
{
    if (eng->acc_cap >= elems) return BG_OK;
    if (eng->d_acc) BG_CUDA(cudaFree(eng->d_acc));
    eng->d_acc = nullptr;
    eng->acc_cap = 0;
    BG_CUDA(cudaMalloc(&eng->d_acc, elems * sizeof(unsigned long long)));
    eng->acc_cap = elems;
    return BG_OK;
}

namespace {
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace

#define BG_ENTER(eng)                                                \
    BG_REQUIRE((eng) != nullptr, BG_EINVAL, "null engine");          \
    DeviceGuard guard__((eng)->device);                              \
    BG_REQUIRE(guard__.ok, BG_ECUDA, "cudaSetDevice failed")

extern "C" {

int bg_version(void) { return BG_VERSION; }
int64_t bg_kernel_launches(void) { return (int64_t)bg_launch_counter.load(std::memory_order_relaxed); }
const char *bg_last_error(void) { return g_err.c_str(); }

void bg_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t out[2])
{
    tf2x32(tf_make_key(k0, k1), x0, x1);
    out[0] = x0;
    out[1] = x1;
}

int bg_key_split(const uint32_t key[2], int64_t num, int layout, uint32_t *out)
{
    BG_REQUIRE(key && out && num >= 0, BG_EINVAL, "bg_key_split: bad argument");
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    const TfKey k = tf_make_key(key[0], key[1]);
    for (int64_t q = 0; q < num; ++q) {
        const TfKey s = tf_split_at(k, (uint64_t)q, (uint64_t)num, layout);
        out[2 * q] = s.k0;
        out[2 * q + 1] = s.k1;
    }
    return BG_OK;
}

int bg_key_chain_next(uint32_t state[2], int layout, uint32_t out[6])
{
    BG_REQUIRE(state && out, BG_EINVAL, "bg_key_chain_next: bad argument");
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    const TfKey cur = tf_make_key(state[0], state[1]);
    const TfKey after = tf_split_at(cur, 0, 2, layout), k = tf_split_at(cur, 1, 2, layout);
    const TfKey after2 = tf_split_at(after, 0, 2, layout), next_k = tf_split_at(after, 1, 2, layout);
    const TfKey next2_k = tf_split_at(after2, 1, 2, layout);
    state[0] = after.k0;
    state[1] = after.k1;
    out[0] = k.k0;
    out[1] = k.k1;
    out[2] = next_k.k0;
    out[3] = next_k.k1;
    out[4] = next2_k.k0;
    out[5] = next2_k.k1;
    return BG_OK;
}

int bg_key_split_at(const uint32_t key[2], int64_t index, int64_t num, int layout, uint32_t out[2])
{
    BG_REQUIRE(key && out && num > 0 && index >= 0 && index < num, BG_EINVAL, "bg_key_split_at: bad argument");
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    const TfKey s = tf_split_at(tf_make_key(key[0], key[1]), (uint64_t)index, (uint64_t)num, layout);
    out[0] = s.k0;
    out[1] = s.k1;
    return BG_OK;
}

int bg_shuffle_sort_keys(const uint32_t *keys, int64_t E, int64_t n, int layout, int rounds, uint32_t *out)
{
    BG_REQUIRE(keys && out && E >= 0 && n >= 0 && rounds >= 0, BG_EINVAL, "bg_shuffle_sort_keys: bad argument");
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    for (int64_t e = 0; e < E; ++e) {
        TfKey key = tf_make_key(keys[2 * e], keys[2 * e + 1]);
        for (int r = 0; r < rounds; ++r) {
            // jax _shuffle: key, subkey = split(key); sort_keys = random_bits(subkey, n)
            const TfKey sub = tf_split_at(key, 1, 2, layout);
            key = tf_split_at(key, 0, 2, layout);
            uint32_t *o = out + ((int64_t)r * E + e) * n;
            for (int64_t j = 0; j < n; ++j) o[j] = tf_bits_at(sub, (uint64_t)j, (uint64_t)n, layout);
        }
    }
    return BG_OK;
}

int bg_random_bits(const uint32_t key[2], int64_t n, int layout, uint32_t *out)
{
    BG_REQUIRE(key && out && n >= 0, BG_EINVAL, "bg_random_bits: bad argument");
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    const TfKey k = tf_make_key(key[0], key[1]);
    for (int64_t j = 0; j < n; ++j) out[j] = tf_bits_at(k, (uint64_t)j, (uint64_t)n, layout);
    return BG_OK;
}

static inline uint32_t threshold_of(float r)
{
    // u = (bits>>9) * 2^-23 exactly, so  u < r  <=>  (bits>>9) < ceil(r * 2^23)
    if (!(r > 0.0f)) return 0u;  // also NaN
    const double t = ceil((double)r * 8388608.0);
    return t >= 8388608.0 ? 8388608u : (uint32_t)t;
}

int bg_thresholds(const float *r, int64_t m, uint32_t *out)
{
    BG_REQUIRE(r && out && m >= 0, BG_EINVAL, "bg_thresholds: bad argument");
    for (int64_t j = 0; j < m; ++j) out[j] = threshold_of(r[j]);
    return BG_OK;
}

// row pitch: a multiple of 32 words, so every bit-plane row starts on a 128-byte line (whole sectors for the
// 32-byte gathers of the fused step kernel, whole lines for the coalesced blend and the TMA boxes)
int64_t bg_words_per_row(int64_t n_markers) { return ((n_markers + 31) / 32 + 31) / 32 * 32; }

int bg_engine_create(int device, bg_engine **out)
{
    BG_REQUIRE(out, BG_EINVAL, "null out pointer");
    int count = 0;
    BG_CUDA(cudaGetDeviceCount(&count));
    BG_REQUIRE(device >= 0 && device < count, BG_EINVAL, "no such CUDA device");
    DeviceGuard g(device);
    BG_REQUIRE(g.ok, BG_ECUDA, "cudaSetDevice failed");
    bg_engine *e = new (std::nothrow) bg_engine();
    BG_REQUIRE(e, BG_ENOMEM, "out of host memory");
    e->device = device;
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&e->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    // tuning switches: the environment is read HERE, once (BG_OPT_<NAME>); nothing on the step path calls getenv
    static const char *const names[] = {"fuse", "gebv_algo", "lookahead", "mask_nt", "mask_big_ctas", "mask_ctas_per_sm", "blend_env_chunk",
                                        "copy_engine", "mapped_d2h_max", "tc_target_ctas", "timing", "gebv_digits", "gebv_shape", "rows_nt", "fused_dyn", "step_pdl",
                                        "xg_parts"};
    for (const char *name : names) {
        std::string env = "BG_OPT_";
        for (const char *c = name; *c; ++c) env += (char)toupper(*c);
        if (const char *v = getenv(env.c_str())) {
            if (bg_engine_set_option(e, name, atoll(v)) != BG_OK) {
                delete e;
                return BG_EINVAL;
            }
        }
    }
    *out = e;
    return BG_OK;
}

int bg_engine_set_option(bg_engine *eng, const char *name, int64_t value)
{
    BG_REQUIRE(eng && name, BG_EINVAL, "bg_engine_set_option: null argument");
    bg_options &o = eng->opt;
    const std::string n(name);
    if (n == "fuse") o.fuse = value != 0;
    else if (n == "gebv_algo") {
        BG_REQUIRE(value >= 0 && value <= 3, BG_EINVAL, "gebv_algo must be 0..3");
        o.gebv_algo = (int)value;
    } else if (n == "lookahead") {
        BG_REQUIRE(value >= 0, BG_EINVAL, "lookahead must be >= 0");
        o.lookahead = (int)(value > BG_BATCH_MAX ? BG_BATCH_MAX : value);
    } else if (n == "mask_nt") {
        BG_REQUIRE(value >= 32 && value <= 256, BG_EINVAL, "mask_nt must be 32..256");
        o.mask_nt = (int)value / 32 * 32;
    } else if (n == "mask_big_ctas") o.mask_big_ctas = value != 0;
    else if (n == "mask_ctas_per_sm") {
        BG_REQUIRE(value >= 0 && value <= 16, BG_EINVAL, "mask_ctas_per_sm must be 0..16");
        o.mask_ctas_per_sm = (int)value;
    }
    else if (n == "blend_env_chunk") {
        BG_REQUIRE(value >= 1, BG_EINVAL, "blend_env_chunk must be >= 1");
        o.blend_env_chunk = (int)value;
    } else if (n == "copy_engine") o.copy_engine = value != 0;
    else if (n == "mapped_d2h_max") o.mapped_d2h_max = value;
    else if (n == "tc_target_ctas") o.tc_target_ctas = value;
    else if (n == "timing") o.timing = value != 0;
    else if (n == "fused_dyn") o.fused_dyn = value < 0 ? -1 : (value != 0);
    else if (n == "step_pdl") o.step_pdl = value != 0;
    else if (n == "xg_parts") {
        BG_REQUIRE(value >= 0 && value < 100000000, BG_EINVAL, "xg_parts: up to 8 decimal digits");
        o.xg_parts = value;
    }
    else if (n == "rows_nt") {
        BG_REQUIRE(value == 0 || (value >= 64 && value <= 1024 && value % 32 == 0), BG_EINVAL, "rows_nt must be 0 or a multiple of 32 in 64..1024");
        o.rows_nt = (int)value;
    } else if (n == "gebv_shape") {
        BG_REQUIRE(value >= 0 && value <= 2, BG_EINVAL, "gebv_shape must be 0 (auto), 1 (short K) or 2 (long K)");
        o.gebv_shape = (int)value;
    } else if (n == "gebv_digits") {
        BG_REQUIRE(value == 0 || (value >= 4 && value <= 8), BG_EINVAL, "gebv_digits must be 0 (auto) or 4..8");
        BG_REQUIRE(eng->m == 0, BG_ESTATE, "gebv_digits must be set before bg_engine_set_map");
        o.gebv_digits = (int)value;
    } else {
        bg_set_error("unknown option '" + n + "'");
        return BG_EINVAL;
    }
    return BG_OK;
}

static void free_map(bg_engine *e)
{
    cudaFree(e->d_thr);
    cudaFree(e->d_thr_cmp);
    e->d_thr_cmp = nullptr;
    cudaFree(e->d_wfix);
    cudaFree(e->d_inv_scale);
    cudaFree(e->d_wdig);
    e->d_thr = nullptr;
    e->d_wfix = nullptr;
    e->d_inv_scale = nullptr;
    e->d_wdig = nullptr;
}

int bg_engine_destroy(bg_engine *eng)
{
    if (!eng) return BG_OK;
    bg_peer_engine_gone(eng);
    {
        DeviceGuard g(eng->device);
        free_map(eng);
        if (eng->side) {
            cudaStreamSynchronize(eng->side);
            cudaStreamDestroy(eng->side);
        }
        for (auto &b : eng->batches) {
            cudaFree(b.mask);
            cudaFree(b.mut);
            if (b.ready) cudaEventDestroy(b.ready);
            if (b.freed) cudaEventDestroy(b.freed);
            if (b.t0) cudaEventDestroy(b.t0);
            if (b.t1) cudaEventDestroy(b.t1);
        }
        if (eng->tmp_event) cudaEventDestroy(eng->tmp_event);
        if (eng->reset_ready) cudaEventDestroy(eng->reset_ready);
        if (eng->join_event) cudaEventDestroy(eng->join_event);
        cudaFree(eng->d_mut);
        cudaFree(eng->d_acc);
        cudaFree(eng->d_xg_work);
        for (int i = 0; i < 2; ++i) {
            cudaFree(eng->d_acc2[i]);
        }
    }
    delete eng;
    return BG_OK;
}

int bg_engine_set_map(bg_engine *eng, const float *recomb, const float *effects, int64_t n_markers, int32_t n_traits,
                      float mutation)
{
    BG_ENTER(eng);
    BG_REQUIRE(recomb && n_markers > 0 && n_markers < (int64_t(1) << 31) - 64, BG_EINVAL, "bad recombination vector");
    BG_REQUIRE(n_traits >= 0 && (n_traits == 0 || effects), BG_EINVAL, "bad marker effects");
    free_map(eng);
    if (eng->side) BG_CUDA(cudaStreamSynchronize(eng->side));
    for (auto &b : eng->batches) b.valid = false;  // masks depend on the thresholds
    eng->m = n_markers;
    eng->W = (int32_t)((n_markers + 31) / 32);
    eng->Wpad = (int32_t)bg_words_per_row(n_markers);
    eng->T = n_traits;
    eng->mut_thr = threshold_of(mutation);

    const size_t nthr = (size_t)eng->Wpad * 32 + 32;
    std::vector<uint32_t> thr(nthr, 0u);
    for (int64_t j = 0; j < n_markers; ++j) thr[j] = threshold_of(recomb[j]);
    BG_CUDA(cudaMalloc(&eng->d_thr, nthr * sizeof(uint32_t)));
    BG_CUDA(cudaMemcpy(eng->d_thr, thr.data(), nthr * sizeof(uint32_t), cudaMemcpyHostToDevice));
    // (bits >> 9) < T  <=>  bits < T << 9 as long as T << 9 fits 32 bits, i.e. T < 2^23 (r < 1): the kernels' fast path
    bool fits = true;
    for (size_t j = 0; j < nthr; ++j) fits = fits && thr[j] < (1u << 23);
    if (fits) {
        for (size_t j = 0; j < nthr; ++j) thr[j] <<= 9;
        BG_CUDA(cudaMalloc(&eng->d_thr_cmp, nthr * sizeof(uint32_t)));
        BG_CUDA(cudaMemcpy(eng->d_thr_cmp, thr.data(), nthr * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }

    if (n_traits > 0) {
        // Fixed point: w_fix = rint(w * 2^s), one scale s per trait.  Sums of w_fix are exact integers, so a GEBV does
        // not depend on the summation order and is rounded once, to float32.  The tensor-core kernels feed w_fix as D
        // balanced base-256 int8 digits per effect; D (and with it s) is chosen PER MAP:
        //   * worst-case quantisation error of a GEBV: m * 2^-s (dosage <= 2, |w - w_fix 2^-s| <= 2^-(s+1)); required
        //     <= 2^-25 * sum|w|, i.e. a quarter of float32's epsilon relative to the GEBV's range 2 sum|w| -- the
        //     reference's own float32 dot (TraitModel.__call__) carries 2^-24 * sum|w d| in the worst case;
        //   * D = the digits that takes (>= 4), raised to the most that costs nothing (the GEMM's N = D * T is padded
        //     to a multiple of 16: one trait always gets all 8); option gebv_digits fixes it instead;
        //   * s = the largest scale D digits hold (64 |w_fix| <= 2^(8D-2): the prescaled operand, below) that keeps any
        //     sum, rounding included, below 2^55 (K-split accumulators: 56-bit sums + an 8-bit arrival count).
        const size_t stride = (size_t)((eng->Wpad + 7) / 8) * 256;
        std::vector<long long> wfix(stride * n_traits, 0ll);
        std::vector<double> inv(n_traits, 1.0);
        std::vector<double> sums(n_traits, 0.0), maxs(n_traits, 0.0);
        int need_bits = 0;  // max over traits of: bits of the largest |w_fix| at the minimum scale
        for (int t = 0; t < n_traits; ++t) {
            double sum = 0.0, mx = 0.0;
            for (int64_t j = 0; j < n_markers; ++j) {
                const double w = (double)effects[j * n_traits + t];
                BG_REQUIRE(isfinite(w), BG_EINVAL, "marker effects must be finite");
                sum += fabs(w);
                if (fabs(w) > mx) mx = fabs(w);
            }
            sums[t] = sum;
            maxs[t] = mx;
            if (sum > 0.0) {
                int e_mean, e_max;
                frexp(sum / (double)n_markers, &e_mean);  // mean|w| >= 2^(e_mean-1)
                frexp(mx, &e_max);                        // max|w| < 2^e_max
                const int s_min = 25 - (e_mean - 1);      // 2^-s <= 2^-25 * mean|w|
                const int bits = s_min + e_max;           // |w_fix| < 2^bits at s_min
                if (bits > need_bits) need_bits = bits;
            }
        }
        int D = eng->opt.gebv_digits;
        if (D == 0) {
            D = (need_bits + 8 + 7) / 8;  // |w_fix| <= 2^(8D-8)  <=>  bits <= 8D - 8
            if (D < 4) D = 4;
            if (D > 8) D = 8;
            while (D < 8 && ((D + 1) * n_traits + 15) / 16 == (D * n_traits + 15) / 16) ++D;  // digits that cost nothing
        }
        for (int t = 0; t < n_traits; ++t) {
            int s = 0;
            if (sums[t] > 0.0) {
                int ex_sum, ex_max;
                frexp(2.0 * sums[t], &ex_sum);  // 2*sum < 2^ex_sum
                frexp(maxs[t], &ex_max);
                s = 54 - ex_sum;                                       // any sum stays below 2^55
                if (s > 8 * D - 8 - ex_max) s = 8 * D - 8 - ex_max;    // |w_fix| <= 2^(8D-8): 64x it fits D balanced digits
                if (s > 1000) s = 1000;
                if (s < -1000) s = -1000;
            }
            for (int64_t j = 0; j < n_markers; ++j)
                wfix[t * stride + j] = llrint(ldexp((double)effects[j * n_traits + t], s));
            inv[t] = ldexp(1.0, -s);
        }
        BG_CUDA(cudaMalloc(&eng->d_wfix, wfix.size() * sizeof(long long)));
        BG_CUDA(cudaMemcpy(eng->d_wfix, wfix.data(), wfix.size() * sizeof(long long), cudaMemcpyHostToDevice));
        BG_CUDA(cudaMalloc(&eng->d_inv_scale, n_traits * sizeof(double)));
        BG_CUDA(cudaMemcpy(eng->d_inv_scale, inv.data(), n_traits * sizeof(double), cudaMemcpyHostToDevice));

        // tensor-core digit table: w_fix = sum_d digit_d * 256^d with balanced digits in [-128,127], D digits per
        // effect (column D*t + d), stored per 128-marker step as [N/8][8][8 rows][16 B] core matrices
        // (gebv_tc2.cu, cross_gebv.cu).  Inside a 32-marker word, K index 4*s + b holds marker 8*b + s (matches the
        // kernels' expansion).  PRESCALED operand: the kernels feed the dosage of marker 8*b + s as the unsigned byte
        // dosage * 4^(s/2) (a masked 2-bit field left where it sits in its byte: one LOP3, no shift),
        // so the digits here are those of w_fix * 4^(3 - s/2) and every tensor-core sum is exactly
        // 64x the plain fixed-point sum (shifted back in the kernels' epilogues).
        eng->tc_N = 0;
        eng->tc_D = D;
        eng->tc_steps = 0;
        if (D * n_traits <= 256) {
            const int N = ((D * n_traits + 15) / 16) * 16;
            const int64_t steps = eng->Wpad / 4;
            std::vector<signed char> dig((size_t)(steps + 1) * N * 128, 0);  // one zero step of slack
            for (int t = 0; t < n_traits; ++t)
                for (int64_t j = 0; j < n_markers; ++j) {
                    long long w = wfix[t * stride + j];
                    if (w == 0) continue;
                    const int64_t st = j / 128;
                    const int word = (int)(j % 128) / 32, bit = (int)(j % 32);
                    w *= 1ll << (2 * (3 - (bit % 8) / 2));
                    const int k = word * 32 + 4 * (bit % 8) + bit / 8;
                    for (int d = 0; d < D; ++d) {
                        const int dgt = (int)(((w + 128) & 255) - 128);
                        w = (w - dgt) >> 8;
                        const int n = D * t + d;
                        const size_t off = (((size_t)st * (N / 8) + n / 8) * 8 + k / 16) * 128 + (n % 8) * 16 + k % 16;
                        dig[off] = (signed char)dgt;
                    }
                    BG_REQUIRE(w == 0, BG_ESTATE, "internal: a fixed-point effect does not fit its digits");
                }
            BG_CUDA(cudaMalloc(&eng->d_wdig, dig.size()));
            BG_CUDA(cudaMemcpy(eng->d_wdig, dig.data(), dig.size(), cudaMemcpyHostToDevice));
            eng->tc_N = N;
            eng->tc_steps = steps;
        }
    }
    return BG_OK;
}

int bg_gebv_digits(bg_engine *eng) { return eng ? eng->tc_D : 0; }

int bg_pack(bg_engine *eng, const uint8_t *bool_in, uint32_t *packed_out, int64_t rows, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(eng->m > 0, BG_ESTATE, "engine has no map (call bg_engine_set_map)");
    BG_REQUIRE(rows >= 0 && (rows == 0 || (bool_in && packed_out)), BG_EINVAL, "bg_pack: bad argument");
    return bg_launch_pack(bool_in, packed_out, rows, eng->m, eng->W, eng->Wpad, (cudaStream_t)stream);
}

int bg_unpack(bg_engine *eng, const uint32_t *packed_in, uint8_t *bool_out, int64_t rows, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(eng->m > 0, BG_ESTATE, "engine has no map (call bg_engine_set_map)");
    BG_REQUIRE(rows >= 0 && (rows == 0 || (packed_in && bool_out)), BG_EINVAL, "bg_unpack: bad argument");
    return bg_launch_unpack(packed_in, bool_out, rows, eng->m, eng->Wpad, (cudaStream_t)stream);
}

int bg_gather_individuals(bg_engine *eng, const uint32_t *src, const int32_t *idx, uint32_t *dst, int64_t E, int64_t n_src,
                          int64_t n, int64_t src_env_rows, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(eng->m > 0, BG_ESTATE, "engine has no map (call bg_engine_set_map)");
    BG_REQUIRE(E >= 0 && n >= 0 && n_src > 0, BG_EINVAL, "bg_gather_individuals: bad shape");
    BG_REQUIRE(E * n == 0 || (src && idx && dst), BG_EINVAL, "bg_gather_individuals: null buffer");
    return bg_launch_gather(src, idx, dst, E, n_src, n, src_env_rows, eng->Wpad, (cudaStream_t)stream);
}

// ---- crossover-mask batches (vector env) -----------------------------------------------------
// key chain of chromax's Simulator.cross: `random_key, k = split(random_key)`
static inline void chain_next(uint32_t state[2], int layout, uint32_t k[2])
{
    const TfKey cur = tf_make_key(state[0], state[1]);
    const TfKey after = tf_split_at(cur, 0, 2, layout), kk = tf_split_at(cur, 1, 2, layout);
    state[0] = after.k0;
    state[1] = after.k1;
    k[0] = kk.k0;
    k[1] = kk.k1;
}

// how many keys' masks one batch buffer may hold (64 MB per buffer at most; 1 for giant maps)
static int batch_capacity(const bg_engine *eng, int64_t rows)
{
    const size_t per_key = (size_t)rows * eng->Wpad * sizeof(uint32_t) * (eng->mut_thr ? 2 : 1);
    size_t c = per_key ? (size_t(64) << 20) / per_key : BG_BATCH_MAX;
    if (c > (size_t)BG_BATCH_MAX) c = BG_BATCH_MAX;
    return c < 1 ? 1 : (int)c;
}

static bool batch_lookup(bg_engine *eng, const uint32_t key[2], int layout, int schedule, int64_t rows, int *bi, int *pos)
{
    for (int i = 0; i < BG_MASK_BATCHES; ++i) {
        const bg_mask_batch &b = eng->batches[i];
        if (!b.valid || b.layout != layout || b.schedule != schedule || b.rows != rows) continue;
        for (int p = 0; p < b.count; ++p)
            if (b.keys[p][0] == key[0] && b.keys[p][1] == key[1]) {
                *bi = i;
                *pos = p;
                return true;
            }
    }
    return false;
}

// generate the masks of keys[0..count) into batch `bi` on stream `on` (after the buffer's readers are done)
static int batch_generate(bg_engine *eng, int bi, int count, const uint32_t (*keys)[2], const uint32_t *state_after, int layout,
                          int schedule, int64_t rows, cudaStream_t on, int small_ctas)
{
    bg_mask_batch &b = eng->batches[bi];
    b.valid = false;
    if (!b.ready) BG_CUDA(cudaEventCreateWithFlags(&b.ready, cudaEventDisableTiming));
    if (!b.freed) BG_CUDA(cudaEventCreateWithFlags(&b.freed, cudaEventDisableTiming));
    if (!eng->tmp_event) BG_CUDA(cudaEventCreateWithFlags(&eng->tmp_event, cudaEventDisableTiming));
    // readers of the old contents: everything enqueued so far on the stream that used it
    if (b.used && b.use_stream != on) {
        BG_CUDA(cudaEventRecord(eng->tmp_event, b.use_stream));
        BG_CUDA(cudaStreamWaitEvent(on, eng->tmp_event, 0));
    }
    if (b.freed_set) BG_CUDA(cudaStreamWaitEvent(on, b.freed, 0));
    if (b.gen_stream && b.gen_stream != on) BG_CUDA(cudaStreamWaitEvent(on, b.ready, 0));  // a stale generation may still be writing
    b.used = false;
    b.freed_set = false;
    b.use_stream = nullptr;
    const size_t words = (size_t)count * rows * eng->Wpad;
    if (b.cap < words || (eng->mut_thr && b.mut_cap < words)) {
        // growing the buffer frees the old one: cudaFree synchronises the device, so nothing can still be reading it
        const size_t want = (size_t)batch_capacity(eng, rows) * rows * eng->Wpad;
        int rc = bg_reserve_u32(&b.mask, &b.cap, want > words ? want : words);
        if (rc) return rc;
        if (eng->mut_thr) {
            rc = bg_reserve_u32(&b.mut, &b.mut_cap, want > words ? want : words);
            if (rc) return rc;
        }
    }
    if (eng->opt.timing) {  // diagnostics: device time of the previous generation into this buffer
        if (!b.t0) {
            BG_CUDA(cudaEventCreate(&b.t0));
            BG_CUDA(cudaEventCreate(&b.t1));
        }
        if (b.timed) {
            float ms = 0.f;
            if (cudaEventSynchronize(b.t1) == cudaSuccess && cudaEventElapsedTime(&ms, b.t0, b.t1) == cudaSuccess)
                fprintf(stderr, "[bg mask batch] %d keys x %lld rows generated in %.1f us\n", b.count, (long long)b.rows, 1e3 * ms);
        }
        BG_CUDA(cudaEventRecord(b.t0, on));
    }
    int rc = bg_launch_mask_batch(eng, rows, count, keys, layout, schedule, b.mask, eng->mut_thr ? b.mut : nullptr, on, small_ctas);
    if (rc) return rc;
    if (eng->opt.timing) {
        BG_CUDA(cudaEventRecord(b.t1, on));
        b.timed = true;
    }
    BG_CUDA(cudaEventRecord(b.ready, on));
    b.gen_stream = on;
    b.synced_stream = on;  // work enqueued on `on` later is ordered behind the kernel anyway
    b.count = count;
    for (int p = 0; p < count; ++p) {
        b.keys[p][0] = keys[p][0];
        b.keys[p][1] = keys[p][1];
    }
    b.has_state = state_after != nullptr;
    if (state_after) {
        b.state_after[0] = state_after[0];
        b.state_after[1] = state_after[1];
    }
    b.layout = layout;
    b.schedule = schedule;
    b.rows = rows;
    b.valid = true;
    b.stamp = ++eng->use_clock;
    return BG_OK;
}

// least recently used batch buffer other than `keep`
static int batch_victim(const bg_engine *eng, int keep)
{
    int v = -1;
    for (int i = 0; i < BG_MASK_BATCHES; ++i) {
        if (i == keep) continue;
        if (!eng->batches[i].valid) return i;
        if (v < 0 || eng->batches[i].stamp < eng->batches[v].stamp) v = i;
    }
    return v;
}

// Masks of `key` for work enqueued on `st` after this returns.  `state_after` (may be NULL): the key-chain state
// after `key` was drawn; when given, the batch that CONTINUES the chain is generated on the side stream as soon as the
// first key of a batch is used, doubling in length up to the lookahead option (1, 2, 4, 8, 8, ...), so that a steady
// stream of steps never waits for masks and a short run does not pay for masks it never uses.
static int masks_acquire(bg_engine *eng, const uint32_t key[2], const uint32_t *state_after, int layout, int schedule, int64_t rows,
                         cudaStream_t st, const uint32_t **mask, const uint32_t **mut, int *batch, bool *first)
{
    int bi = -1, pos = 0;
    if (!batch_lookup(eng, key, layout, schedule, rows, &bi, &pos)) {
        bi = batch_victim(eng, -1);
        const uint32_t keys[1][2] = {{key[0], key[1]}};
        const int rc = batch_generate(eng, bi, 1, keys, state_after, layout, schedule, rows, st, 0);
        if (rc) return rc;
        pos = 0;
    }
    bg_mask_batch &b = eng->batches[bi];
    if (b.synced_stream != st) {
        BG_CUDA(cudaStreamWaitEvent(st, b.ready, 0));
        b.synced_stream = st;
    }
    if (b.used && b.use_stream != st) {  // readers move to another stream: leave a marker behind the old ones
        BG_CUDA(cudaEventRecord(b.freed, b.use_stream));
        b.freed_set = true;
    }
    b.used = true;
    b.use_stream = st;
    b.stamp = ++eng->use_clock;
    const size_t off = (size_t)pos * rows * eng->Wpad;
    *mask = b.mask + off;
    *mut = eng->mut_thr ? b.mut + off : nullptr;

    *batch = bi;
    *first = pos == 0;
    return BG_OK;
}

// lookahead: called after the step that used the FIRST key of chain batch `bi` has been enqueued -- make sure the batch
// that continues the chain exists or is being generated (on the side stream, behind nothing: the mask kernel's small
// CTAs run beside the step kernel)
static int masks_continue(bg_engine *eng, int bi, cudaStream_t st)
{
    const bg_mask_batch &b = eng->batches[bi];
    if (!b.has_state || eng->opt.lookahead <= 0) return BG_OK;
    const int layout = b.layout, schedule = b.schedule;
    const int64_t rows = b.rows;
    uint32_t state[2] = {b.state_after[0], b.state_after[1]};
    uint32_t keys[BG_BATCH_MAX][2];
    chain_next(state, layout, keys[0]);
    int ci, cp;
    if (batch_lookup(eng, keys[0], layout, schedule, rows, &ci, &cp)) return BG_OK;
    int count = 2 * b.count;
    const int cap = batch_capacity(eng, rows);
    if (count > eng->opt.lookahead) count = eng->opt.lookahead;
    if (count > cap) count = cap;
    for (int p = 1; p < count; ++p) chain_next(state, layout, keys[p]);
    if (!eng->side) BG_CUDA(cudaStreamCreateWithFlags(&eng->side, cudaStreamNonBlocking));
    (void)st;
    // small CTAs that fit beside the step kernel's two CTAs per SM (cross_gebv.cu)
    return batch_generate(eng, batch_victim(eng, bi), count, keys, state, layout, schedule, rows, eng->side,
                          eng->opt.mask_big_ctas ? 0 : 1);
}

// gebv_out != nullptr: also score the offspring; fused into one kernel when the tensor-core path applies
static int cross_envs_impl(bg_engine *eng, const uint32_t *pop, const int32_t *parents, uint32_t *out, int64_t E, int64_t n_src,
                           int64_t n, const uint32_t cross_key[2], const uint32_t *state_after, int layout, int schedule,
                           float *gebv_out, cudaStream_t st)
{
    const uint32_t *mask = nullptr, *mut = nullptr;
    int batch = -1;
    bool first = false;
    int rc = masks_acquire(eng, cross_key, state_after, layout, schedule, 2 * n, st, &mask, &mut, &batch, &first);
    if (rc) return rc;
    // One fused cross+GEBV kernel (cross_gebv.cu) whenever the tensor-core path applies; option fuse=0 selects the
    // two-kernel path (cross-checks, tuning).
    if (gebv_out && eng->opt.fuse && bg_cross_gebv_fused_ok(eng, E, n_src, n)) {
        const int64_t tiles = (E * n + 127) / 128;
        const bool dyn = eng->opt.fused_dyn > 0 || (eng->opt.fused_dyn < 0 && tiles >= 8LL * eng->sm_count);
        rc = (dyn && bg_cross_gebv_dyn_ok(eng, E, n_src, n))
                 ? bg_launch_cross_gebv_dyn(eng, pop, parents, mask, out, E, n_src, n, gebv_out, st)
                 : bg_launch_cross_gebv_fused(eng, pop, parents, mask, out, E, n_src, n, gebv_out, st);
    } else {
        // blend and GEBV are adjacent in the stream (no event between them) so that the GEBV kernel's
        // programmatic dependent launch can overlap its prologue with the blend's tail
        rc = bg_launch_blend(eng, pop, parents, mask, mut, out, E, n_src, n, st);
        if (!rc && gebv_out) rc = bg_launch_gebv(eng, out, E * n, gebv_out, 0, st);
    }
    if (!rc && first) rc = masks_continue(eng, batch, st);
    return rc;
}

int bg_cross(bg_engine *eng, const uint32_t *pop, const int32_t *parents, uint32_t *out, int64_t E, int64_t n_src, int64_t n,
             const uint32_t cross_key[2], int layout, int schedule, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(E >= 0 && n >= 0 && n_src > 0 && cross_key, BG_EINVAL, "bg_cross: bad shape");
    BG_REQUIRE(E * n == 0 || (pop && parents && out), BG_EINVAL, "bg_cross: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    if (E * n == 0) return BG_OK;
    if (E == 1)
        return bg_launch_meiosis_rows(eng, BG_ROWS_CROSS, 2 * n, cross_key, layout, schedule, nullptr, nullptr, pop, parents, n_src,
                                      0, out, st);
    return cross_envs_impl(eng, pop, parents, out, E, n_src, n, cross_key, nullptr, layout, schedule, nullptr, st);
}

int bg_cross_gebv(bg_engine *eng, const uint32_t *pop, const int32_t *parents, uint32_t *out, int64_t E, int64_t n_src, int64_t n,
                  const uint32_t cross_key[2], int layout, int schedule, float *gebv_out, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(E >= 0 && n >= 0 && n_src > 0 && cross_key, BG_EINVAL, "bg_cross_gebv: bad shape");
    BG_REQUIRE(E * n == 0 || (pop && parents && out && gebv_out), BG_EINVAL, "bg_cross_gebv: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    if (E * n == 0) return BG_OK;
    if (E == 1) {
        const int rc = bg_launch_meiosis_rows(eng, BG_ROWS_CROSS, 2 * n, cross_key, layout, schedule, nullptr, nullptr, pop,
                                              parents, n_src, 0, out, st);
        return rc ? rc : bg_launch_gebv(eng, out, n, gebv_out, 0, st);
    }
    return cross_envs_impl(eng, pop, parents, out, E, n_src, n, cross_key, nullptr, layout, schedule, gebv_out, st);
}

int bg_blend_envs(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask, const uint32_t *mut,
                  uint32_t *out, int64_t E, int64_t n_src, int64_t n, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(eng->m > 0, BG_ESTATE, "engine has no map (call bg_engine_set_map)");
    BG_REQUIRE(E >= 0 && n >= 0 && n_src > 0, BG_EINVAL, "bg_blend_envs: bad shape");
    BG_REQUIRE(E * n == 0 || (pop && parents && mask && out), BG_EINVAL, "bg_blend_envs: null buffer");
    return bg_launch_blend(eng, pop, parents, mask, mut, out, E, n_src, n, (cudaStream_t)stream);
}

int bg_double_haploid(bg_engine *eng, const uint32_t *pop, uint32_t *out, int64_t E, int64_t n, int64_t n_offspring,
                      const uint32_t cross_key[2], int layout, int schedule, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(E >= 0 && n >= 0 && n_offspring >= 0 && cross_key, BG_EINVAL, "bg_double_haploid: bad shape");
    if (E * n * n_offspring == 0) return BG_OK;
    BG_REQUIRE(pop && out, BG_EINVAL, "bg_double_haploid: null buffer");
    return bg_launch_double_haploid(eng, E, n, n_offspring, cross_key, layout, schedule, pop, out, (cudaStream_t)stream);
}

int bg_meiosis_masks(bg_engine *eng, uint32_t *mask_out, int64_t rows, const uint32_t cross_key[2], int layout, int schedule,
                     void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(rows >= 0 && cross_key && (rows == 0 || mask_out), BG_EINVAL, "bg_meiosis_masks: bad argument");
    int rc = BG_OK;
    if (eng->mut_thr) {
        rc = bg_reserve_u32(&eng->d_mut, &eng->mut_cap, (size_t)rows * eng->Wpad);
        if (rc) return rc;
    }
    return bg_launch_meiosis_rows(eng, BG_ROWS_MASK, rows, cross_key, layout, schedule, mask_out,
                                  eng->mut_thr ? eng->d_mut : nullptr, nullptr, nullptr, 0, 0, nullptr, (cudaStream_t)stream);
}

int bg_gebv_algo(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, int algo, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(rows >= 0 && (rows == 0 || (pop && out)), BG_EINVAL, "bg_gebv: bad argument");
    return bg_launch_gebv(eng, pop, rows, out, algo, (cudaStream_t)stream);
}

int bg_gebv(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, void *stream)
{
    return bg_gebv_algo(eng, pop, rows, out, 0, stream);
}

int bg_reduce_max(bg_engine *eng, const float *gebv, int64_t E, int64_t per_env, float *out, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(E >= 0 && (E == 0 || (gebv && out)), BG_EINVAL, "bg_reduce_max: bad argument");
    return bg_launch_reduce(gebv, E, per_env, out, 0, (cudaStream_t)stream);
}

int bg_reduce_mean(bg_engine *eng, const float *gebv, int64_t E, int64_t per_env, float *out, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(E >= 0 && (E == 0 || (gebv && out)), BG_EINVAL, "bg_reduce_mean: bad argument");
    return bg_launch_reduce(gebv, E, per_env, out, 1, (cudaStream_t)stream);
}

int bg_topk(bg_engine *eng, const float *scores, int64_t rows, int64_t len, int32_t k, float *vals_out, int32_t *idx_out, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(rows >= 0 && len >= 0 && (rows == 0 || (scores && vals_out && idx_out)), BG_EINVAL, "bg_topk: bad argument");
    return bg_launch_topk(scores, rows, len, k, vals_out, idx_out, (cudaStream_t)stream);
}

int bg_pairs_from_topk(bg_engine *eng, const float *vals, const int32_t *idx, int64_t E, int32_t k, int64_t row_len, int32_t *pairs_out,
                       void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(E >= 0 && (E == 0 || (vals && idx && pairs_out)), BG_EINVAL, "bg_pairs_from_topk: bad argument");
    return bg_launch_pairs_from_topk(vals, idx, E, k, row_len, pairs_out, (cudaStream_t)stream);
}

int bg_diallel_pairs(bg_engine *eng, const int32_t *best, const int32_t *perm, int64_t E, int32_t k, int32_t nc, int64_t n,
                     int32_t *pairs_out, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(E >= 0 && (E == 0 || (best && perm && pairs_out)), BG_EINVAL, "bg_diallel_pairs: bad argument");
    return bg_launch_diallel_pairs(best, perm, E, k, nc, n, pairs_out, (cudaStream_t)stream);
}

int bg_reset_indices(bg_engine *eng, const uint32_t random_key[2], int64_t E_total, int64_t env_begin, int64_t E, int64_t n_germ,
                     int64_t n, int layout, int32_t *idx_out, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(random_key && E >= 0 && n >= 0 && n_germ > 0 && (E * n == 0 || idx_out), BG_EINVAL, "bg_reset_indices: bad argument");
    BG_REQUIRE(env_begin >= 0 && env_begin + E <= E_total, BG_EINVAL, "bg_reset_indices: env range outside [0, E_total)");
    return bg_launch_reset_indices(eng, random_key, E_total, env_begin, E, n_germ, n, layout, idx_out, (cudaStream_t)stream);
}

int bg_vec_reset(bg_engine *eng, const uint32_t *germplasm, int64_t n_germ, const uint32_t random_key[2], int64_t E_total,
                 int64_t env_begin, int64_t E, int64_t n, int layout, int32_t *idx_dev, uint32_t *pop_out, float *gebv_dev,
                 float *gebv_host, const float *germ_gebv, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(germplasm && random_key && idx_dev && pop_out && E > 0 && n > 0 && n_germ > 0, BG_EINVAL, "bg_vec_reset: bad argument");
    BG_REQUIRE(env_begin >= 0 && env_begin + E <= E_total, BG_EINVAL, "bg_vec_reset: env range outside [0, E_total)");
    BG_REQUIRE(!gebv_host || gebv_dev, BG_EINVAL, "bg_vec_reset: gebv_host needs gebv_dev");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = bg_launch_reset_indices(eng, random_key, E_total, env_begin, E, n_germ, n, layout, idx_dev, st);
    if (rc) return rc;
    // a GEBV is a function of the individual alone: with the germplasm's GEBVs at hand the reset infos are gathered
    // along with the individuals (bit-identical: every GEBV kernel sums the same integers) instead of recomputed
    const bool gather_gebv = gebv_dev && germ_gebv;
    rc = bg_launch_gather(germplasm, idx_dev, pop_out, E, n_germ, n, 0, eng->Wpad, st, gather_gebv ? germ_gebv : nullptr,
                          gather_gebv ? gebv_dev : nullptr, eng->T);
    if (rc) return rc;
    if (gebv_dev && !gather_gebv) {
        rc = bg_launch_gebv(eng, pop_out, E * n, gebv_dev, 0, st);
        if (rc) return rc;
    }
    if (gebv_host) {
        BG_CUDA(cudaMemcpyAsync(gebv_host, gebv_dev, (size_t)E * n * eng->T * sizeof(float), cudaMemcpyDeviceToHost, st));
        BG_CUDA(cudaStreamSynchronize(st));
    }
    return BG_OK;
}

int bg_vec_reset_prefetch(bg_engine *eng, const uint32_t *germplasm, int64_t n_germ, const uint32_t random_key[2], int64_t E_total,
                          int64_t env_begin, int64_t E, int64_t n, int layout, int32_t *idx_dev, uint32_t *pop_out, float *gebv_dev,
                          const float *germ_gebv, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(germ_gebv && gebv_dev, BG_EINVAL, "bg_vec_reset_prefetch: needs the germplasm's GEBVs (no GEBV kernel off the step stream)");
    cudaStream_t st = (cudaStream_t)stream;
    if (!eng->side) BG_CUDA(cudaStreamCreateWithFlags(&eng->side, cudaStreamNonBlocking));
    if (!eng->tmp_event) BG_CUDA(cudaEventCreateWithFlags(&eng->tmp_event, cudaEventDisableTiming));
    if (!eng->reset_ready) BG_CUDA(cudaEventCreateWithFlags(&eng->reset_ready, cudaEventDisableTiming));
    // behind everything enqueued so far on the caller's stream: the buffers' previous readers
    BG_CUDA(cudaEventRecord(eng->tmp_event, st));
    BG_CUDA(cudaStreamWaitEvent(eng->side, eng->tmp_event, 0));
    const int rc = bg_vec_reset(eng, germplasm, n_germ, random_key, E_total, env_begin, E, n, layout, idx_dev, pop_out, gebv_dev, nullptr,
                                germ_gebv, eng->side);
    if (rc) return rc;
    BG_CUDA(cudaEventRecord(eng->reset_ready, eng->side));
    eng->reset_pending = true;
    return BG_OK;
}

int bg_engine_join(bg_engine *eng, void *stream)
{
    BG_ENTER(eng);
    if (!eng->side) return BG_OK;
    if (!eng->join_event) BG_CUDA(cudaEventCreateWithFlags(&eng->join_event, cudaEventDisableTiming));
    BG_CUDA(cudaEventRecord(eng->join_event, eng->side));
    BG_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, eng->join_event, 0));
    return BG_OK;
}

int bg_vec_reset_adopt(bg_engine *eng, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(eng->reset_pending, BG_ESTATE, "bg_vec_reset_adopt: no prefetched reset");
    BG_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, eng->reset_ready, 0));
    eng->reset_pending = false;
    return BG_OK;
}

// Host <-> device transfers of the host-facing step.  SMALL ones (single-env actions / GEBVs, per-env rewards) go through
// a copy kernel on the mapped pinned buffer when the host pointer is device-accessible (pinned + 16-byte aligned): no
// copy-engine hand-off (C1 step 100.7 -> 92.8 us).  Larger ones use the copy engine: at the 189 KB + 95 KB of the 64-env
// step the zero-copy kernels measured slower (105 vs 98 us per step).  Option copy_engine=1: always the copy engine.
constexpr size_t BG_MAPPED_COPY_MAX = 32 * 1024;
static void *mapped_device_pointer(const bg_engine *eng, const void *host)
{
    if (eng->opt.copy_engine || ((uintptr_t)host & 15)) return nullptr;
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, const_cast<void *>(host), 0) != cudaSuccess) {
        cudaGetLastError();  // pageable memory: not an error for us
        dp = nullptr;
    }
    return dp;
}

static int copy_h2d(const bg_engine *eng, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st)
{
    if (void *dp = (bytes % 4 == 0 && bytes <= BG_MAPPED_COPY_MAX) ? mapped_device_pointer(eng, src_host) : nullptr)
        return bg_launch_copy_mapped(dp, dst_dev, bytes, st);
    BG_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
    return BG_OK;
}

static int copy_d2h(const bg_engine *eng, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st)
{
    if (void *dp = (bytes % 4 == 0 && (long long)bytes <= eng->opt.mapped_d2h_max) ? mapped_device_pointer(eng, dst_host) : nullptr)
        return bg_launch_copy_mapped(src_dev, dp, bytes, st);
    BG_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    return BG_OK;
}

// option timing=1: host-side wall time of each phase of bg_vec_step, printed every 1000 calls (diagnostics)
struct StepTimer {
    double acc[4] = {0, 0, 0, 0};
    long calls = 0;
    std::chrono::steady_clock::time_point t;
    void start(bool on)
    {
        if (on) t = std::chrono::steady_clock::now();
    }
    void lap(bool on, int i)
    {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        acc[i] += std::chrono::duration<double, std::micro>(n - t).count();
        t = n;
    }
    void done(bool on)
    {
        if (!on || ++calls % 1000) return;
        fprintf(stderr, "[bg_vec_step us/call] h2d %.2f step %.2f reduce+d2h %.2f sync %.2f\n", acc[0] / calls, acc[1] / calls,
                acc[2] / calls, acc[3] / calls);
    }
};
static StepTimer g_step_timer;

int bg_vec_step(bg_engine *eng, const uint32_t *pop, uint32_t *out, const int32_t *actions_host, int32_t *actions_dev, int64_t E,
                int64_t n_src, int64_t n, uint32_t key_state[2], int layout, int schedule, float *gebv_dev, float *reward_dev,
                float *gebv_host, float *reward_host, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(actions_dev && gebv_dev, BG_EINVAL, "bg_vec_step: null device buffer");
    BG_REQUIRE(!reward_host || reward_dev, BG_EINVAL, "bg_vec_step: reward_host needs reward_dev");
    BG_REQUIRE(E > 0 && n > 0 && n_src > 0 && key_state && pop && out, BG_EINVAL, "bg_vec_step: bad argument");
    BG_REQUIRE(layout == BG_LAYOUT_LEGACY || layout == BG_LAYOUT_PARTITIONABLE, BG_EINVAL, "bad PRNG layout");
    cudaStream_t st = (cudaStream_t)stream;
    const bool tm_on = eng->opt.timing != 0;
    StepTimer &tm = g_step_timer;
    tm.start(tm_on);
    int rc;
    if (actions_host) {
        rc = copy_h2d(eng, actions_dev, actions_host, (size_t)E * n * 2 * sizeof(int32_t), st);
        if (rc) return rc;
    }
    tm.lap(tm_on, 0);
    // random_key, k = split(random_key)   (chromax Simulator.cross)
    uint32_t state[2] = {key_state[0], key_state[1]}, k[2];
    chain_next(state, layout, k);
    if (E == 1) {
        rc = bg_launch_meiosis_rows(eng, BG_ROWS_CROSS, 2 * n, k, layout, schedule, nullptr, nullptr, pop, actions_dev, n_src, 0, out,
                                    st);
        if (!rc) rc = bg_launch_gebv(eng, out, n, gebv_dev, 0, st);
    } else {
        rc = cross_envs_impl(eng, pop, actions_dev, out, E, n_src, n, k, state, layout, schedule, gebv_dev, st);
    }
    if (rc) return rc;
    key_state[0] = state[0];  // the chain advances only when the step has been enqueued
    key_state[1] = state[1];
    tm.lap(tm_on, 1);
    if (reward_dev) {
        // with a peer exchange attached (env-sharded run) the reduction also stores the rewards into every rank's window
        rc = eng->peer ? bg_launch_reduce_publish(gebv_dev, E, n * eng->T, reward_dev, eng->peer, st)
                       : bg_launch_reduce(gebv_dev, E, n * eng->T, reward_dev, 0, st);
        if (rc) return rc;
    }
    bool sync = false;
    if (gebv_host) {
        rc = copy_d2h(eng, gebv_host, gebv_dev, (size_t)E * n * eng->T * sizeof(float), st);
        if (rc) return rc;
        sync = true;
    }
    if (reward_host) {
        rc = copy_d2h(eng, reward_host, reward_dev, (size_t)E * sizeof(float), st);
        if (rc) return rc;
        sync = true;
    }
    tm.lap(tm_on, 2);
    if (sync) BG_CUDA(cudaStreamSynchronize(st));
    tm.lap(tm_on, 3);
    tm.done(tm_on);
    return BG_OK;
}

}  // extern "C"

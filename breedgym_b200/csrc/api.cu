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
    *out = e;
    return BG_OK;
}

static void free_map(bg_engine *e)
{
    cudaFree(e->d_thr);
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
    {
        DeviceGuard g(eng->device);
        free_map(eng);
        if (eng->side) {
            cudaStreamSynchronize(eng->side);
            cudaStreamDestroy(eng->side);
        }
        for (auto &sl : eng->slots) {
            cudaFree(sl.mask);
            cudaFree(sl.mut);
            if (sl.ready) cudaEventDestroy(sl.ready);
            if (sl.freed) cudaEventDestroy(sl.freed);
        }
        cudaFree(eng->d_mut);
        cudaFree(eng->d_acc);
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
    for (auto &sl : eng->slots) sl.valid = false;  // masks depend on the thresholds
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

    if (n_traits > 0) {
        // fixed point: w_fix = rint(w * 2^s), s = largest shift with 2*sum|w| * 2^s < 2^54, so that any GEBV sum, rounding of
        // the w_fix included (+ at most one unit per marker), stays below 2^55 in magnitude (the tensor-core kernels
        // sum 64x these integers, see the digit table below, and must stay inside int64)
        const size_t stride = (size_t)((eng->Wpad + 7) / 8) * 256;
        std::vector<long long> wfix(stride * n_traits, 0ll);
        std::vector<double> inv(n_traits, 1.0);
        for (int t = 0; t < n_traits; ++t) {
            double sum = 0.0;
            for (int64_t j = 0; j < n_markers; ++j) {
                const double w = (double)effects[j * n_traits + t];
                BG_REQUIRE(isfinite(w), BG_EINVAL, "marker effects must be finite");
                sum += fabs(w);
            }
            int s = 0;
            if (sum > 0.0) {
                int ex;
                frexp(2.0 * sum, &ex);  // 2*sum < 2^ex
                s = 54 - ex;
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

        // tensor-core digit table: w_fix = sum_d digit_d * 256^d with balanced digits in [-128,127],
        // stored per 128-marker step as [N/8][8][8 rows][16 B] core matrices (gebv_tc*.cu).  Inside
        // a 32-marker word, K index 4*s + b holds marker 8*b + s (matches the kernels' expansion).
        // PRESCALED operand: the kernels feed the dosage of marker 8*b + s as the unsigned byte
        // dosage * 4^(s/2) (a masked 2-bit field left where it sits in its byte: one LOP3, no shift),
        // so the digits here are those of w_fix * 4^(3 - s/2) and every tensor-core sum is exactly
        // 64x the plain fixed-point sum (shifted back in the kernels' epilogues).
        eng->tc_N = 0;
        eng->tc_steps = 0;
        if (n_traits <= bg_gebv_tc_max_traits()) {
            const int N = ((8 * n_traits + 15) / 16) * 16;
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
                    for (int d = 0; d < 8; ++d) {
                        const int dgt = (int)(((w + 128) & 255) - 128);
                        w = (w - dgt) >> 8;
                        const int n = 8 * t + d;
                        const size_t off = (((size_t)st * (N / 8) + n / 8) * 8 + k / 16) * 128 + (n % 8) * 16 + k % 16;
                        dig[off] = (signed char)dgt;
                    }
                }
            BG_CUDA(cudaMalloc(&eng->d_wdig, dig.size()));
            BG_CUDA(cudaMemcpy(eng->d_wdig, dig.data(), dig.size(), cudaMemcpyHostToDevice));
            eng->tc_N = N;
            eng->tc_steps = steps;
        }
    }
    return BG_OK;
}

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

// ---- crossover-mask slots (vector env) ------------------------------------------------------
static bool slot_matches(const bg_mask_slot &sl, const uint32_t key[2], int layout, int schedule, int64_t rows)
{
    // timing experiments only (results are wrong): any generated slot serves any key -> no mask kernel in the step
    static const bool reuse = getenv("BG_DEBUG_REUSE_MASKS") != nullptr;
    if (reuse && sl.valid && sl.rows == rows) return true;
    return sl.valid && sl.key[0] == key[0] && sl.key[1] == key[1] && sl.layout == layout && sl.schedule == schedule &&
           sl.rows == rows;
}

static int slot_prepare(bg_engine *eng, bg_mask_slot &sl, int64_t rows)
{
    const size_t words = (size_t)rows * eng->Wpad;
    int rc = bg_reserve_u32(&sl.mask, &sl.cap, words);
    if (rc) return rc;
    if (eng->mut_thr) {
        rc = bg_reserve_u32(&sl.mut, &sl.mut_cap, words);
        if (rc) return rc;
    }
    if (!sl.ready) BG_CUDA(cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming));
    if (!sl.freed) BG_CUDA(cudaEventCreateWithFlags(&sl.freed, cudaEventDisableTiming));
    return BG_OK;
}

static int slot_generate(bg_engine *eng, bg_mask_slot &sl, const uint32_t key[2], int layout, int schedule, int64_t rows,
                         cudaStream_t on, int small_ctas = 0)
{
    sl.valid = false;
    int rc = slot_prepare(eng, sl, rows);
    if (rc) return rc;
    rc = bg_launch_meiosis_rows(eng, BG_ROWS_MASK, rows, key, layout, schedule, sl.mask, eng->mut_thr ? sl.mut : nullptr, nullptr,
                                nullptr, 0, 0, nullptr, on, small_ctas);
    if (rc) return rc;
    BG_CUDA(cudaEventRecord(sl.ready, on));
    sl.ready_set = true;
    sl.valid = true;
    sl.key[0] = key[0];
    sl.key[1] = key[1];
    sl.layout = layout;
    sl.schedule = schedule;
    sl.rows = rows;
    return BG_OK;
}

// masks for `key`, usable by work enqueued on `st` after this returns
static int masks_acquire(bg_engine *eng, const uint32_t key[2], int layout, int schedule, int64_t rows, cudaStream_t st, int *slot)
{
    for (int i = 0; i < BG_MASK_SLOTS; ++i)
        if (slot_matches(eng->slots[i], key, layout, schedule, rows)) {
            BG_CUDA(cudaStreamWaitEvent(st, eng->slots[i].ready, 0));  // generated ahead of time (or earlier on st)
            *slot = i;
            return BG_OK;
        }
    const int i = (eng->last_slot + 1) % BG_MASK_SLOTS;
    bg_mask_slot &sl = eng->slots[i];
    if (sl.ready_set) BG_CUDA(cudaStreamWaitEvent(st, sl.ready, 0));  // a stale lookahead may still be writing the slot
    if (sl.freed_set) BG_CUDA(cudaStreamWaitEvent(st, sl.freed, 0));
    const int rc = slot_generate(eng, sl, key, layout, schedule, rows, st);
    if (rc) return rc;
    *slot = i;
    return BG_OK;
}

static int masks_release(bg_engine *eng, int slot, cudaStream_t st)
{
    bg_mask_slot &sl = eng->slots[slot];
    BG_CUDA(cudaEventRecord(sl.freed, st));
    sl.freed_set = true;
    eng->last_slot = slot;
    return BG_OK;
}

// start generating the masks of an UPCOMING cross key on the side stream, in a slot that holds neither the current
// masks (`cur`) nor those of another upcoming key (`keep`, -1: none); returns the slot used (or already holding them)
// `after_step`: start only once the kernels of the current step are done (their `freed` record).  The host-facing
// step synchronises and leaves the GPU idle while the host works: the integer-bound mask kernel then runs in that gap
// instead of competing with the step kernel for issue slots.  The device-resident pipeline has no gap: there the mask
// kernels run TWO steps ahead, so that they only fill the slots the step kernels leave free and never delay their start.
static int masks_lookahead(bg_engine *eng, const uint32_t key[2], int layout, int schedule, int64_t rows, int cur, int keep,
                           bool after_step, int *used)
{
    for (int i = 0; i < BG_MASK_SLOTS; ++i)
        if (slot_matches(eng->slots[i], key, layout, schedule, rows)) {
            *used = i;
            return BG_OK;
        }
    if (!eng->side) BG_CUDA(cudaStreamCreateWithFlags(&eng->side, cudaStreamNonBlocking));
    int v = -1;
    for (int i = 0; i < BG_MASK_SLOTS; ++i)
        if (i != cur && i != keep) v = i;
    bg_mask_slot &sl = eng->slots[v];
    if (after_step && eng->slots[cur].freed_set) BG_CUDA(cudaStreamWaitEvent(eng->side, eng->slots[cur].freed, 0));
    if (sl.freed_set) BG_CUDA(cudaStreamWaitEvent(eng->side, sl.freed, 0));  // its last reader (an earlier step) is done
    *used = v;
    // overlapping the step kernel: small CTAs that fit beside its two CTAs per SM
    static const bool big = getenv("BG_MASK_BIG_CTAS") != nullptr;  // diagnostics
    return slot_generate(eng, sl, key, layout, schedule, rows, eng->side, (after_step || big) ? 0 : 1);
}

// gebv_out != nullptr: also score the offspring; fused into one kernel when the tensor-core path applies
static int cross_envs_impl(bg_engine *eng, const uint32_t *pop, const int32_t *parents, uint32_t *out, int64_t E, int64_t n_src,
                           int64_t n, const uint32_t cross_key[2], const uint32_t *next_key, int layout, int schedule,
                           float *gebv_out, cudaStream_t st, bool lookahead_after_step = false)
{
    // One fused cross+GEBV kernel (cross_gebv.cu: 47 us at C2 against 33 + 32 us for blend + GEBV) whenever the
    // tensor-core path applies; BG_NO_FUSE=1 selects the two-kernel path (cross-checks, tuning).
    const bool no_fuse = getenv("BG_NO_FUSE") != nullptr;
    int slot = 0;
    int rc = masks_acquire(eng, cross_key, layout, schedule, 2 * n, st, &slot);
    if (rc) return rc;
    const bg_mask_slot &sl = eng->slots[slot];
    if (gebv_out && !no_fuse && bg_cross_gebv_fused_ok(eng, E, n_src, n)) {
        rc = bg_launch_cross_gebv_fused(eng, pop, parents, sl.mask, out, E, n_src, n, gebv_out, st);
    } else {
        // blend and GEBV are adjacent in the stream (no event between them) so that the GEBV kernel's
        // programmatic dependent launch can overlap its prologue with the blend's tail
        rc = bg_launch_blend(eng, pop, parents, sl.mask, eng->mut_thr ? sl.mut : nullptr, out, E, n_src, n, st);
        if (!rc && gebv_out) rc = bg_launch_gebv(eng, out, E * n, gebv_out, 0, st);
    }
    if (rc) return rc;
    rc = masks_release(eng, slot, st);
    if (rc) return rc;
    static const bool no_lookahead = getenv("BG_NO_LOOKAHEAD") != nullptr;  // diagnostics
    if (next_key && !no_lookahead) {
        // next_key[0..1]: the next step's key, next_key[2..3]: the one after it
        int s1 = -1, s2 = -1;
        rc = masks_lookahead(eng, next_key, layout, schedule, 2 * n, slot, -1, lookahead_after_step, &s1);
        static const bool one_ahead = getenv("BG_LOOKAHEAD1") != nullptr;  // diagnostics
        if (!rc && !lookahead_after_step && !one_ahead)
            rc = masks_lookahead(eng, next_key + 2, layout, schedule, 2 * n, slot, s1, false, &s2);
    }
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

int bg_double_haploid(bg_engine *eng, const uint32_t *pop, uint32_t *out, int64_t n, int64_t n_offspring,
                      const uint32_t cross_key[2], int layout, int schedule, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(n >= 0 && n_offspring >= 0 && cross_key, BG_EINVAL, "bg_double_haploid: bad shape");
    if (n * n_offspring == 0) return BG_OK;
    BG_REQUIRE(pop && out, BG_EINVAL, "bg_double_haploid: null buffer");
    return bg_launch_meiosis_rows(eng, BG_ROWS_DH, n * n_offspring, cross_key, layout, schedule, nullptr, nullptr, pop, nullptr, n,
                                  n_offspring, out, (cudaStream_t)stream);
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

// Host <-> device transfers of the host-facing step.  SMALL ones (single-env actions / GEBVs, per-env rewards) go through
// a copy kernel on the mapped pinned buffer when the host pointer is device-accessible (pinned + 16-byte aligned): no
// copy-engine hand-off (C1 step 100.7 -> 92.8 us).  Larger ones use the copy engine: at the 189 KB + 95 KB of the 64-env
// step the zero-copy kernels measured slower (105 vs 98 us per step).  BG_COPY_ENGINE=1: always the copy engine.
constexpr size_t BG_MAPPED_COPY_MAX = 32 * 1024;
static void *mapped_device_pointer(const void *host)
{
    static const bool off = getenv("BG_COPY_ENGINE") != nullptr;
    if (off || ((uintptr_t)host & 15)) return nullptr;
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, const_cast<void *>(host), 0) != cudaSuccess) {
        cudaGetLastError();  // pageable memory: not an error for us
        dp = nullptr;
    }
    return dp;
}

static int copy_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st)
{
    if (void *dp = (bytes % 4 == 0 && bytes <= BG_MAPPED_COPY_MAX) ? mapped_device_pointer(src_host) : nullptr)
        return bg_launch_copy_mapped(dp, dst_dev, bytes, st);
    BG_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
    return BG_OK;
}

static int copy_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st)
{
    static const size_t d2h_max = getenv("BG_MAPPED_D2H_MAX") ? (size_t)atoll(getenv("BG_MAPPED_D2H_MAX")) : BG_MAPPED_COPY_MAX;  // tuning
    if (void *dp = (bytes % 4 == 0 && bytes <= d2h_max) ? mapped_device_pointer(dst_host) : nullptr)
        return bg_launch_copy_mapped(src_dev, dp, bytes, st);
    BG_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    return BG_OK;
}

// BG_TIMING=1: host-side wall time of each phase of bg_vec_step, printed every 1000 calls (diagnostics)
struct StepTimer {
    bool on;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    long calls = 0;
    std::chrono::steady_clock::time_point t;
    StepTimer() : on(getenv("BG_TIMING") != nullptr) {}
    void start()
    {
        if (on) t = std::chrono::steady_clock::now();
    }
    void lap(int i)
    {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        acc[i] += std::chrono::duration<double, std::micro>(n - t).count();
        t = n;
    }
    void done()
    {
        if (!on || ++calls % 1000) return;
        fprintf(stderr, "[bg_vec_step us/call] h2d %.1f cross %.1f gebv %.1f reduce %.1f d2h %.1f sync %.1f\n", acc[0] / calls,
                acc[1] / calls, acc[2] / calls, acc[3] / calls, acc[4] / calls, acc[5] / calls);
    }
};
static StepTimer g_step_timer;

int bg_vec_step(bg_engine *eng, const uint32_t *pop, uint32_t *out, const int32_t *actions_host, int32_t *actions_dev, int64_t E,
                int64_t n_src, int64_t n, const uint32_t cross_key[2], const uint32_t *next_cross_key, int layout, int schedule,
                float *gebv_dev, float *reward_dev, float *gebv_host, float *reward_host, void *stream)
{
    BG_ENTER(eng);
    BG_REQUIRE(actions_dev && gebv_dev, BG_EINVAL, "bg_vec_step: null device buffer");
    BG_REQUIRE(!reward_host || reward_dev, BG_EINVAL, "bg_vec_step: reward_host needs reward_dev");
    BG_REQUIRE(E > 0 && n > 0 && n_src > 0 && cross_key && pop && out, BG_EINVAL, "bg_vec_step: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    StepTimer &tm = g_step_timer;
    tm.start();
    int rc;
    if (actions_host) {
        rc = copy_h2d(actions_dev, actions_host, (size_t)E * n * 2 * sizeof(int32_t), st);
        if (rc) return rc;
    }
    tm.lap(0);
    if (E == 1) {
        rc = bg_launch_meiosis_rows(eng, BG_ROWS_CROSS, 2 * n, cross_key, layout, schedule, nullptr, nullptr, pop, actions_dev,
                                    n_src, 0, out, st);
        if (!rc) rc = bg_launch_gebv(eng, out, n, gebv_dev, 0, st);
    } else {
        rc = cross_envs_impl(eng, pop, actions_dev, out, E, n_src, n, cross_key, next_cross_key, layout, schedule, gebv_dev, st,
                             gebv_host != nullptr || reward_host != nullptr);
    }
    if (rc) return rc;
    tm.lap(1);
    tm.lap(2);
    if (reward_dev) {
        rc = bg_launch_reduce(gebv_dev, E, n * eng->T, reward_dev, 0, st);
        if (rc) return rc;
    }
    tm.lap(3);
    bool sync = false;
    if (gebv_host) {
        rc = copy_d2h(gebv_host, gebv_dev, (size_t)E * n * eng->T * sizeof(float), st);
        if (rc) return rc;
        sync = true;
    }
    if (reward_host) {
        rc = copy_d2h(reward_host, reward_dev, (size_t)E * sizeof(float), st);
        if (rc) return rc;
        sync = true;
    }
    tm.lap(4);
    if (sync) BG_CUDA(cudaStreamSynchronize(st));
    tm.lap(5);
    tm.done();
    return BG_OK;
}

}  // extern "C"

/* breedgym_b200 -- C ABI of the B200-native breeding-simulation engine.
 *
 * Drop-in boundary for the ONE hot path of younik/breedgym: meiosis/cross ->
 * GEBV scoring -> vector-env step.  The reference has no FFI layer: its operator
 * API is the Python object `chromax.Simulator` as BreedGym calls it.  Each entry
 * point below names the reference interface it replaces (paths relative to
 * /root/reference; "chromax:" = the un-vendored PyPI dependency).
 * The ctypes binding a maintainer would add is shown in INTEGRATION.md and is
 * what breedgym_b200/_lib.py does.
 *
 * Conventions
 *   - every function returns 0 on success, a negative BG_E* code on failure;
 *     bg_last_error() gives the message (thread local).  No C++ exception crosses.
 *   - all population / action / output buffers are CALLER-allocated device memory
 *     (e.g. torch CUDA tensors' data_ptr()); the library never frees or retains
 *     them.  The engine owns only its copies of the map constants and scratch.
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as
 *     void*); nothing synchronises unless stated.  One engine per (device,thread).
 *   - packed population layout: uint32 words [rows][2][Wpad], haplotype-major bit
 *     planes, marker j <-> bit (j & 31) of word (j >> 5), Wpad = bg_words_per_row(m)
 *     (ceil(m/32) rounded up to a multiple of 32 => every row starts on a 128-byte line);
 *     padding bits are zero.  The reference's byte layout `bool[rows][m][2]`
 *     (breedgym/breedgym.py:47, vec_env.py:57-62) appears only at bg_pack/bg_unpack.
 */
#ifndef BREEDGYM_B200_H
#define BREEDGYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_VERSION 210

#define BG_OK 0
#define BG_EINVAL (-1)   /* bad argument */
#define BG_ECUDA (-2)    /* CUDA runtime error */
#define BG_ENOMEM (-3)   /* allocation failure */
#define BG_ELIMIT (-4)   /* shape beyond a documented kernel limit */
#define BG_ESTATE (-5)   /* engine not configured (bg_engine_set_map missing) */

/* jax PRNG bit layout */
#define BG_LAYOUT_LEGACY 0        /* jax_threefry_partitionable=False (jax < 0.5) */
#define BG_LAYOUT_PARTITIONABLE 1 /* jax_threefry_partitionable=True  (jax >= 0.5) */
/* chromax per-gamete key schedule */
#define BG_SCHEDULE_S1 1 /* gamete key drives the recombination draw directly */
#define BG_SCHEDULE_S2 2 /* gamete key is split into (recombination, mutation) keys */

typedef struct bg_engine bg_engine;

int bg_version(void);
/* number of CUDA kernels this library has launched in this process (monotonic) */
int64_t bg_kernel_launches(void);
const char *bg_last_error(void);

/* ---- host-side PRNG helpers (no GPU touched) ------------------------------
 * Replace jax.random.key/split/bits as used for the KEY CHAIN only
 * (chromax: Simulator.cross `random_key, k = split(random_key)`;
 *  breedgym/vector/vec_env.py:115,120; breedgym/vector/vec_wrappers.py:82). */
void bg_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t out[2]);
int bg_key_split(const uint32_t key[2], int64_t num, int layout, uint32_t *out /* [num][2] */);
/* element `index` of split(key, num) without computing the others (vec_env.py:120-121 keeps keys[0] only) */
int bg_key_split_at(const uint32_t key[2], int64_t index, int64_t num, int layout, uint32_t out[2]);
int bg_random_bits(const uint32_t key[2], int64_t n, int layout, uint32_t *out /* [n] */);
/* sort keys of jax.random.permutation / choice(replace=False) for E independent keys at once
 * (breedgym/vector/vec_wrappers.py:68-70): per key and round, key, sub = split(key);
 * out[r][e][:] = random_bits(sub, n).  keys: [E][2], out: [rounds][E][n] */
int bg_shuffle_sort_keys(const uint32_t *keys, int64_t E, int64_t n, int layout, int rounds, uint32_t *out);
/* one link of chromax's chain `random_key, k = split(random_key)`: state <- split(state)[0],
 * out[0..1] = k, out[2..3] / out[4..5] = the k the NEXT two calls will return (lookahead for bg_vec_step) */
int bg_key_chain_next(uint32_t state[2], int layout, uint32_t out[6]);
/* integer form of `uniform(key) < r`: (bits >> 9) < T,  T = clamp(ceil(r * 2^23), 0, 2^23) */
int bg_thresholds(const float *r, int64_t m, uint32_t *out /* [m] */);

/* ---- engine ---------------------------------------------------------------- */
int64_t bg_words_per_row(int64_t n_markers);

/* chromax: Simulator.__init__ (device placement of recombination_vec and
 * GEBV_model.marker_effects; breedgym/breedgym.py:36, vec_env.py:45). */
int bg_engine_create(int device, bg_engine **out);
int bg_engine_destroy(bg_engine *eng);
/* Tuning / cross-check switches (none changes results).  bg_engine_create reads each once from the
 * environment as BG_OPT_<NAME>; nothing on the step path calls getenv.
 *   fuse (1)            0: bg_cross_gebv / bg_vec_step run blend + GEBV kernels instead of the fused one
 *   fused_dyn (-1)      which fused step kernel: 0 = one CTA per (tile, K range), 1 = persistent CTAs with a dynamic
 *                       work queue, -1 = the latter from 4 tiles per resident CTA on (512 envs x 370 per GPU)
 *   xg_parts (0)        the persistent kernel's split of a tile's K range, proportions as decimal digits
 *                       (8642 = 8 : 6 : 4 : 2); 0 = by tiles per CTA
 *   step_pdl (1)        programmatic dependent launch of consecutive step kernels (prologue overlaps the previous tail)
 *   gebv_algo (0)       default algorithm of bg_gebv, see bg_gebv_algo
 *   gebv_digits (0)     base-256 digits of the tensor-core GEBV operand: 0 = as many as the map needs,
 *                       4..8 fixed (set BEFORE bg_engine_set_map)
 *   gebv_shape (0)      pipeline shape of the tensor-core GEBV: 0 = by K-loop length, 1 = short K, 2 = long K
 *   lookahead (8)       steps of crossover masks generated ahead of bg_vec_step on a side stream
 *   mask_ctas_per_sm (0) k > 0: the lookahead mask kernel runs as a persistent grid of k small CTAs per SM (a fixed
 *                       footprint beside the step kernel; measured slower than one CTA per row on a lower-priority stream)
 *   rows_nt (0 = by row length), mask_nt (128), mask_big_ctas (0), blend_env_chunk (8), tc_target_ctas (0 = auto),
 *   copy_engine (0), mapped_d2h_max (32768), timing (0)      launch-shape / transfer tuning */
int bg_engine_set_option(bg_engine *eng, const char *name, int64_t value);
/* recomb: host float32[m] (already shifted / chromosome starts = 0.5);
 * effects: host float32[m][n_traits] row-major; mutation: chromax `mutation`. */
int bg_engine_set_map(bg_engine *eng, const float *recomb, const float *effects, int64_t n_markers,
                      int32_t n_traits, float mutation);

/* ---- layout conversion ------------------------------------------------------
 * chromax: Simulator.load_population -> device array (breedgym/breedgym.py:38-42);
 * bg_unpack materialises the observation `bool[rows][m][2]` on request. */
int bg_pack(bg_engine *eng, const uint8_t *bool_in, uint32_t *packed_out, int64_t rows, void *stream);
int bg_unpack(bg_engine *eng, const uint32_t *packed_in, uint8_t *bool_out, int64_t rows, void *stream);
/* dst[e][r] = src[e * src_env_rows + idx[e][r]]  (whole individuals, both planes).
 * `populations[arange, idx]` style gathers: germplasm[selected] (breedgym.py:125,
 * vec_env.py:126-128 after the permutation), Simulator.select's pop[idx].
 * src_env_rows = 0 broadcasts one source population to every env. */
int bg_gather_individuals(bg_engine *eng, const uint32_t *src, const int32_t *idx /* dev [E][n] */, uint32_t *dst,
                          int64_t E, int64_t n_src, int64_t n, int64_t src_env_rows, void *stream);

/* ---- meiosis / cross --------------------------------------------------------
 * Replaces `parents = population[action]; simulator.cross(parents)`
 * (breedgym/breedgym.py:142-143) and the vmapped, SHARED-KEY version
 * (breedgym/vector/vec_env.py:75-77,89-91) -- chromax: functional.cross/_meiosis.
 * pop:     packed [E][n_src][2][Wpad]
 * parents: device int32 [E][n][2], jnp indexing semantics (negatives wrap once,
 *          then clamp)
 * out:     packed [E][n][2][Wpad]; out[e][i][p] = gamete of pop[e][parents[e][i][p]]
 * cross_key: the `k` of `random_key, k = split(random_key)`; gamete (i,p) uses key
 *          #(2i+p) of split(k, 2n) for EVERY env (the reference's vmap shares it).
 * E == 1 runs the fused unique-key kernel; E > 1 generates the 2n masks once and
 * blends all envs against them. */
int bg_cross(bg_engine *eng, const uint32_t *pop, const int32_t *parents, uint32_t *out, int64_t E, int64_t n_src,
             int64_t n, const uint32_t cross_key[2], int layout, int schedule, void *stream);

/* bg_cross followed by the GEBV of the offspring (gebv_out float32 [E][n][n_traits]).  For E > 1
 * (vector env) the two run as ONE kernel (cross_gebv.cu): the parents' bit planes are gathered into
 * shared memory, the offspring words are stored once and scored on the tensor cores before they
 * leave the SM (breedgym/vector/vec_env.py:89-94: cross, then get_info's GEBV_model(populations)).
 * With mutation > 0 or more than 32 traits it falls back to bg_cross + bg_gebv; BG_NO_FUSE=1 in
 * the environment forces that path (cross-checks). */
int bg_cross_gebv(bg_engine *eng, const uint32_t *pop, const int32_t *parents, uint32_t *out, int64_t E, int64_t n_src,
                  int64_t n, const uint32_t cross_key[2], int layout, int schedule, float *gebv_out, void *stream);

/* second half of the E > 1 path of bg_cross on its own: blend every env against
 * precomputed masks (mask / mut: packed [2n][Wpad] from bg_meiosis_masks; mut may
 * be NULL).  Same reference interface as bg_cross (vec_env.py:75-77,89-91). */
int bg_blend_envs(bg_engine *eng, const uint32_t *pop, const int32_t *parents, const uint32_t *mask, const uint32_t *mut,
                  uint32_t *out, int64_t E, int64_t n_src, int64_t n, void *stream);

/* chromax: Simulator.double_haploid / functional.double_haploid, and its vmap over envs under ONE key
 * (breedgym/vector/breeding_programs_env.py:39-41: `vmap(simulator.double_haploid, in_axes=(None, 0))`).
 * pop packed [E][n][2][Wpad] -> out packed [E][n][n_offspring][2][Wpad], both planes = the gamete of key
 * #(i*n_offspring+o) of split(k, n*n_offspring), the same keys for every env; E = 1: one population. */
int bg_double_haploid(bg_engine *eng, const uint32_t *pop, uint32_t *out, int64_t E, int64_t n, int64_t n_offspring,
                      const uint32_t cross_key[2], int layout, int schedule, void *stream);

/* crossover masks only (tests / diagnostics): mask_out [rows][Wpad], inclusive
 * prefix-XOR of the recombination draws of key #q of split(k, rows). */
int bg_meiosis_masks(bg_engine *eng, uint32_t *mask_out, int64_t rows, const uint32_t cross_key[2], int layout,
                     int schedule, void *stream);

/* ---- GEBV -------------------------------------------------------------------
 * chromax: TraitModel.__call__ = dot(sum(pop,-1), effects) (+0 offset) as called
 * from Simulator.GEBV / GEBV_model (breedgym/breedgym.py:233, vec_env.py:132-134).
 * pop packed [rows][2][Wpad] -> out float32 [rows][n_traits].
 * Arithmetic: 64-bit fixed point of the float32 effects (scale chosen per map, see bg_gebv_digits),
 * exact integer sums, one final rounding to float32: deterministic, order independent, identical
 * across kernels; within 1 float32 ulp of a float64 dot for single-trait maps, within
 * 2^-25 * sum|effects| + 1 ulp in general. */
int bg_gebv(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, void *stream);
/* bg_gebv with an explicit kernel choice (cross-checks, tuning, more traits than the tensor-core tile holds);
 * algorithm id: 0 auto, 1 direct bit-test, 2 byte-LUT, 3 tcgen05 int8 GEMM
 * (TMA loads, dosage operand in tensor memory).  All produce the same 64-bit integers. */
int bg_gebv_algo(bg_engine *eng, const uint32_t *pop, int64_t rows, float *out, int algo, void *stream);

/* base-256 int8 digits per marker effect in the tensor-core GEBV operand (4..8), chosen per map by
 * bg_engine_set_map from the effects' range (TraitModel.marker_effects): the float32 effects are held
 * in fixed point with a worst-case GEBV error <= 2^-25 * sum|effects| (one trait always gets all 8). */
int bg_gebv_digits(bg_engine *eng);

/* rews = np.max(infos["GEBV"], axis=(1,2))  (breedgym/vector/vec_env.py:97):
 * gebv float32 [E][per_env] -> out float32 [E] */
int bg_reduce_max(bg_engine *eng, const float *gebv, int64_t E, int64_t per_env, float *out, void *stream);
/* np.mean(GEBV.to_numpy()) (breedgym/breedgym.py:153), accumulated in float64 */
int bg_reduce_mean(bg_engine *eng, const float *gebv, int64_t E, int64_t per_env, float *out, void *stream);

/* ---- top-k -------------------------------------------------------------------
 * jax.lax.top_k along the last axis, as the action wrappers call it on flattened pair scores
 * (breedgym/vector/vec_wrappers.py:101, breedgym/vector/breeding_programs_env.py:27): scores device float32
 * [rows][len] -> the k largest per row, DESCENDING, ties -> lower index (-0.0 == +0.0); vals_out float32 [rows][k],
 * idx_out int32 [rows][k].  k <= 1024.  Radix select + bitonic sort, one CTA per row. */
int bg_topk(bg_engine *eng, const float *scores, int64_t rows, int64_t len, int32_t k, float *vals_out, int32_t *idx_out,
            void *stream);

/* ---- pair selection after the top-k -----------------------------------------------
 * PairScores._convert_actions (breedgym/vector/vec_wrappers.py:100-112; WheatBreedGym's conversion,
 * breeding_programs_env.py:24-36): vals / idx = the k best flattened pair scores of every env (bg_topk's output,
 * descending) -> pairs int32 [E][k][2]: pair b = (idx / row_len, idx % row_len) fills ceil(softmax(vals)[b] * k)
 * output slots, `jnp.repeat(..., total_repeat_length=k)` (cut, or padded with the last pair).  k <= 1024. */
int bg_pairs_from_topk(bg_engine *eng, const float *vals, const int32_t *idx, int64_t E, int32_t k, int64_t row_len,
                       int32_t *pairs_out, void *stream);
/* SelectionScores._convert_actions (vec_wrappers.py:60-78): best int32 [E][k] (the k best individuals, bg_topk) and
 * perm int32 [E][nc] (the chosen entries of the C(k,2) upper-triangular pair list of `Simulator._diallel_indices`,
 * = jax.random.choice(replace=False): bg_reset_indices) -> pairs int32 [E][n][2], every chosen pair repeated
 * ceil(n / nc) times, `total_repeat_length=n`. */
int bg_diallel_pairs(bg_engine *eng, const int32_t *best, const int32_t *perm, int64_t E, int32_t k, int32_t nc, int64_t n,
                     int32_t *pairs_out, void *stream);

/* ---- reset ------------------------------------------------------------------
 * VecBreedGym.reset's `_random_selection` (breedgym/vector/vec_env.py:22-27,
 * 120-128): env e draws permutation(keys[1+e], n_germ)[:n] where
 * keys = split(random_key, E_total+1).  A shard computes envs
 * [env_begin, env_begin+E) of the E_total logical envs; idx_out device int32 [E][n]
 * (then bg_gather_individuals(germplasm, idx_out, ..., src_env_rows = 0)). */
int bg_reset_indices(bg_engine *eng, const uint32_t random_key[2], int64_t E_total, int64_t env_begin, int64_t E,
                     int64_t n_germ, int64_t n, int layout, int32_t *idx_out, void *stream);

/* VecBreedGym.reset in one call (breedgym/vector/vec_env.py:109-130): bg_reset_indices, the gather
 * of the drawn individuals from the germplasm (packed [n_germ][2][Wpad]) into pop_out
 * ([E][n][2][Wpad]) and, when gebv_dev is non-NULL, the reset infos GEBV_model(populations)
 * ([E][n][T]); gebv_host non-NULL copies them to the host and synchronises the stream.
 * germ_gebv (device float32 [n_germ][T] = bg_gebv of the germplasm, or NULL): when given, the
 * reset infos are gathered from it along with the individuals instead of recomputed. */
int bg_vec_reset(bg_engine *eng, const uint32_t *germplasm, int64_t n_germ, const uint32_t random_key[2], int64_t E_total,
                 int64_t env_begin, int64_t E, int64_t n, int layout, int32_t *idx_dev, uint32_t *pop_out, float *gebv_dev,
                 float *gebv_host, const float *germ_gebv, void *stream);

/* The NEXT autoreset, ahead of time: it depends on the reset key chain and the germplasm only
 * (breedgym/vector/vec_env.py:109-130), so it can be drawn while the episode runs.  bg_vec_reset_prefetch enqueues the
 * work of bg_vec_reset (germ_gebv required, no host copy) on the engine's internal side stream, behind everything
 * enqueued so far on `stream` (the previous readers of the buffers); bg_vec_reset_adopt makes `stream` wait for it.
 * The caller keeps pop_out / gebv_dev / idx_dev untouched in between. */
int bg_vec_reset_prefetch(bg_engine *eng, const uint32_t *germplasm, int64_t n_germ, const uint32_t random_key[2],
                          int64_t E_total, int64_t env_begin, int64_t E, int64_t n, int layout, int32_t *idx_dev,
                          uint32_t *pop_out, float *gebv_dev, const float *germ_gebv, void *stream);
int bg_vec_reset_adopt(bg_engine *eng, void *stream);
/* Order `stream` behind everything the engine has enqueued so far on its internal side stream (the crossover masks of
 * FOLLOWING steps, a prefetched reset): a timing harness calls it before its closing event so that a measured window
 * pays for all the work its steps started, not only for what the step stream itself ran. */
int bg_engine_join(bg_engine *eng, void *stream);

/* ---- one-call vector-env step ------------------------------------------------
 * VecBreedGym.step hot path (breedgym/vector/vec_env.py:88-100) with host
 * buffers at the boundary: copies actions_host (int32 [E][n][2], pinned or
 * pageable; NULL = the actions are already in actions_dev) to `actions_dev`,
 * advances the simulator's key chain IN PLACE (`key_state`: host uint32[2] =
 * Simulator.random_key; chromax: `random_key, k = split(random_key)`), runs
 * cross -> GEBV (-> max reward when reward_dev is non-NULL; with a peer exchange attached to the engine,
 * bg_engine_set_peer, the same reduction also stores the rewards into every rank's window), copies gebv/reward
 * back to the host buffers when non-NULL, and synchronises the stream iff any
 * device->host copy was requested.  The crossover masks depend on the key chain
 * only: the masks of the following steps (option `lookahead`) are generated by
 * ONE kernel launch per batch of steps on an internal side stream while the
 * current steps run; a reseed simply misses and regenerates. */
int bg_vec_step(bg_engine *eng, const uint32_t *pop, uint32_t *out, const int32_t *actions_host, int32_t *actions_dev,
                int64_t E, int64_t n_src, int64_t n, uint32_t key_state[2], int layout, int schedule,
                float *gebv_dev /* [E][n][T] */, float *reward_dev /* [E] or NULL */, float *gebv_host /* or NULL */,
                float *reward_host /* or NULL */, void *stream);

/* ---- reward all-gather (multi-GPU) ---------------------------------------------
 * The ONE collective of the env-sharded path: every rank contributes float32[count]
 * rewards, every rank receives float32[world * count] in rank order.  B200-native
 * replacement for the observation/reward pipes of DistributedBreedGym
 * (breedgym/vector/vec_env.py:197-219: step_async / step_wait through host
 * subprocess pipes).  ncclAllGather over NVLink on the caller's stream: no host
 * round trip, capturable, ordered behind the step that produced the rewards.
 * NCCL (libnccl.so.2, the copy the process already has loaded -- torch's -- or the
 * system one) is bound at run time with dlopen; without it bg_comm_* return BG_ESTATE.
 *   rank 0: bg_comm_unique_id(id) -> broadcast the 128 bytes (any channel) -> every
 *   rank: bg_comm_create(eng, id, world, rank, &comm)  (collective call).
 * bg_comm_create runs one warm-up all-gather and synchronises, so the first timed
 * collective does not pay NCCL's lazy channel setup. */
typedef struct bg_comm bg_comm;
#define BG_COMM_ID_BYTES 128
int bg_comm_unique_id(uint8_t id_out[BG_COMM_ID_BYTES]);
int bg_comm_create(bg_engine *eng, const uint8_t id[BG_COMM_ID_BYTES], int world, int rank, bg_comm **out);
int bg_comm_destroy(bg_comm *comm);
int bg_allgather_f32(bg_comm *comm, const float *send_dev, float *recv_dev, int64_t count, void *stream);

/* ---- reward exchange over peer memory (multi-GPU, one NVSwitch box) ---------------
 * The same exchange WITHOUT a collective launch: the kernel that reduces a step's
 * GEBVs to per-env rewards (breedgym/vector/vec_env.py:95-100) stores each reward
 * straight into a small receive window on EVERY rank (peer-to-peer stores over
 * NVLink) and, once all of this rank's rewards are out, raises the rank's epoch flag
 * in every window.  No NCCL kernel, no extra launch, no host work at an episode's
 * end; a consumer orders itself behind the exchange with bg_peer_wait (a one-warp
 * kernel that spins on the local flags).  Replaces the reward half of
 * DistributedBreedGym's step_wait (vec_env.py:197-219).
 *   every rank: bg_peer_create(eng, world, rank, total, offset, &peer)  (`total` rewards over all
 *   ranks, this rank owns [offset, offset + its E): ragged shards are fine); bg_peer_handle(peer, h)
 *   -> all-gather the BG_PEER_HANDLE_BYTES of every rank on the host (any channel)
 *   -> every rank: bg_peer_connect(peer, handles)   (CUDA IPC; ranks that live in the
 *   same process are connected through their raw pointers)
 *   -> bg_engine_set_peer(eng, peer): from now on every bg_vec_step of that engine
 *   with reward_dev != NULL publishes its E rewards as the next EPOCH.
 * Window of epoch k: float32[total] (env order) at bg_peer_result(peer, k & 1).
 * Flow control: epoch k is written only after every rank has published epoch k - 1,
 * so the rewards of an episode stay valid on a rank until that rank ends its next
 * episode.  All spins are bounded (option below): a rank that never arrives is
 * counted in bg_peer_timeouts instead of hanging the GPU.  The window itself is
 * never freed before the process exits (late peers must not fault). */
typedef struct bg_peer bg_peer;
#define BG_PEER_HANDLE_BYTES 96
#define BG_PEER_MAX_WORLD 16
int bg_peer_create(bg_engine *eng, int world, int rank, int64_t total, int64_t offset, bg_peer **out);
int bg_peer_handle(bg_peer *peer, uint8_t handle_out[BG_PEER_HANDLE_BYTES]);
int bg_peer_connect(bg_peer *peer, const uint8_t *handles /* [world][BG_PEER_HANDLE_BYTES] */);
int bg_engine_set_peer(bg_engine *eng, bg_peer *peer /* NULL: detach */);
/* standalone publish of this rank's float32[count] as the next epoch, for rewards that
 * were not reduced by bg_vec_step */
int bg_peer_publish_f32(bg_peer *peer, const float *send_dev, int64_t count, void *stream);
/* make `stream` wait (on the device) until every rank's slice of the LAST published epoch has landed */
int bg_peer_wait(bg_peer *peer, void *stream);
int64_t bg_peer_epoch(bg_peer *peer);                 /* epochs published so far by this rank */
float *bg_peer_result(bg_peer *peer, int parity);     /* device pointer: float32[total] of epochs with k & 1 == parity */
int bg_peer_set_timeout_ms(bg_peer *peer, int64_t ms); /* bound of every device-side spin (default 20 000) */
int64_t bg_peer_timeouts(bg_peer *peer);              /* spins that hit the bound so far (synchronises the device) */
int bg_peer_destroy(bg_peer *peer);

#ifdef __cplusplus
}
#endif
#endif /* BREEDGYM_B200_H */

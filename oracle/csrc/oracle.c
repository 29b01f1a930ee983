/* CPU oracle, C restatement -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference hot path (jax.random Threefry-2x32 ->
 * uniform < r -> cumulative XOR -> allele select; dosage x effects dot), byte
 * wide and float-compare based exactly like the reference's XLA program, so it
 * shares no bit tricks with the CUDA kernels.  Used (a) as a second,
 * structurally independent checker next to oracle/chromax_ref.py and (b) as the
 * timed CPU baseline in bench.py (cpu_baseline kind "port", and --impl reference
 * because neither jax nor chromax is installable on this image).
 *
 * Follows: breedgym/breedgym.py:142-143 (gather + cross), breedgym/vector/
 * vec_env.py:75-94 (shared-key vmap cross, GEBV of all envs), chromax
 * functional.cross/_meiosis and TraitModel.__call__,
 * jax/_src/prng.py threefry_random_bits / threefry_split.
 *
 * PARITY UNPINNED against real chromax/jax (not installable here); pinned
 * against Random123 / documented JAX vectors in tests/test_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LAYOUT_LEGACY 0
#define LAYOUT_PARTITIONABLE 1
#define SCHED_S1 1
#define SCHED_S2 2

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *o0, uint32_t *o1)
{
    static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    x0 += ks[0];
    x1 += ks[1];
    for (int g = 0; g < 5; ++g) {
        for (int i = 0; i < 4; ++i) {
            x0 += x1;
            x1 = rotl32(x1, R[g & 1][i]);
            x1 ^= x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    *o0 = x0;
    *o1 = x1;
}

/* jax.random.bits(key,(n,),uint32) */
void orc_random_bits(const uint32_t key[2], int64_t n, int layout, uint32_t *out)
{
    if (layout == LAYOUT_LEGACY) {
        int64_t h = (n + 1) / 2;
        for (int64_t c = 0; c < h; ++c) {
            uint32_t c1 = (c + h < n) ? (uint32_t)(c + h) : 0u; /* odd n: one zero pad */
            uint32_t a, b;
            orc_threefry2x32(key[0], key[1], (uint32_t)c, c1, &a, &b);
            out[c] = a;
            if (c + h < n) out[c + h] = b;
        }
    } else {
        for (int64_t j = 0; j < n; ++j) {
            uint32_t a, b;
            orc_threefry2x32(key[0], key[1], (uint32_t)((uint64_t)j >> 32), (uint32_t)j, &a, &b);
            out[j] = a ^ b;
        }
    }
}

/* jax.random.split(key,num) -> out[num][2] */
void orc_split(const uint32_t key[2], int64_t num, int layout, uint32_t *out)
{
    if (layout == LAYOUT_LEGACY) {
        orc_random_bits(key, 2 * num, LAYOUT_LEGACY, out);
    } else {
        for (int64_t j = 0; j < num; ++j)
            orc_threefry2x32(key[0], key[1], (uint32_t)((uint64_t)j >> 32), (uint32_t)j, &out[2 * j], &out[2 * j + 1]);
    }
}

static inline float bits_to_uniform(uint32_t bits)
{
    uint32_t u = (bits >> 9) | 0x3F800000u;
    float f;
    memcpy(&f, &u, 4);
    f -= 1.0f;
    return f < 0.0f ? 0.0f : f;
}

/* chromax _meiosis: individual bool[m][2] (byte interleaved) -> hap, written with stride 2 */
static void meiosis_row(const uint8_t *ind, const float *r, int64_t m, const uint32_t key[2], float mutation,
                        int schedule, int layout, uint32_t *bits_scratch, uint8_t *out_stride2)
{
    uint32_t krec[2] = {key[0], key[1]}, kmut[2] = {0, 0};
    if (schedule == SCHED_S2) {
        uint32_t ks[4];
        orc_split(key, 2, layout, ks);
        krec[0] = ks[0]; krec[1] = ks[1];
        kmut[0] = ks[2]; kmut[1] = ks[3];
    }
    orc_random_bits(krec, m, layout, bits_scratch);
    uint8_t mask = 0;
    for (int64_t j = 0; j < m; ++j) {
        mask ^= (uint8_t)(bits_to_uniform(bits_scratch[j]) < r[j]);
        out_stride2[2 * j] = ind[2 * j + mask];
    }
    if (mutation > 0.0f && schedule == SCHED_S2) {
        orc_random_bits(kmut, m, layout, bits_scratch);
        for (int64_t j = 0; j < m; ++j)
            out_stride2[2 * j] ^= (uint8_t)(bits_to_uniform(bits_scratch[j]) < mutation);
    }
}

static inline int64_t norm_index(int64_t a, int64_t n)
{
    if (a < 0) a += n;
    if (a < 0) a = 0;
    if (a > n - 1) a = n - 1;
    return a;
}

/* population[action] gather + Simulator.cross for E envs sharing ONE cross key
 * (E = 1: BreedGym.step; E > 1: VecBreedGym.step's vmap).  pops bool[E][N][m][2],
 * actions int32[E][n][2], out bool[E][n][m][2].  Returns 0, or -1 on alloc failure. */
int orc_cross_envs(const uint8_t *pops, const int32_t *actions, const float *r, int64_t E, int64_t N, int64_t n,
                   int64_t m, const uint32_t cross_key[2], float mutation, int schedule, int layout, uint8_t *out)
{
    uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)n);
    if (!keys) return -1;
    orc_split(cross_key, 2 * n, layout, keys);
    int err = 0;
#pragma omp parallel
    {
        uint32_t *scratch = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(m + 1));
        if (!scratch) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(dynamic, 4) collapse(2)
            for (int64_t e = 0; e < E; ++e)
                for (int64_t q = 0; q < 2 * n; ++q) {
                    int64_t i = q >> 1, p = q & 1;
                    int64_t a = norm_index(actions[(e * n + i) * 2 + p], N);
                    const uint8_t *ind = pops + ((e * N + a) * m) * 2;
                    uint8_t *o = out + ((e * n + i) * m) * 2 + p;
                    meiosis_row(ind, r, m, &keys[2 * q], mutation, schedule, layout, scratch, o);
                }
            free(scratch);
        }
    }
    free(keys);
    return err;
}

/* The same step the way the reference's vmap executes it (breedgym/vector/vec_env.py:75-77: `jax.vmap(
 * simulator.cross, in_axes=(None, 0))` traces the key split and the uniform draws ONCE, outside the env axis): the
 * 2n crossover masks are drawn once per step (byte per marker, as XLA materialises `cumxor(u < r)`), then every env
 * only gathers and selects.  Same results as orc_cross_envs; this is the honest CPU baseline of the vector env
 * (bench.py: cpu_baseline / --impl reference), orc_cross_envs re-draws per env (E x the Threefry work). */
int orc_cross_envs_shared(const uint8_t *pops, const int32_t *actions, const float *r, int64_t E, int64_t N, int64_t n,
                          int64_t m, const uint32_t cross_key[2], float mutation, int schedule, int layout, uint8_t *out)
{
    const int64_t rows = 2 * n;
    uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (size_t)rows);
    uint8_t *mask = (uint8_t *)malloc((size_t)rows * (size_t)m);
    uint8_t *mut = (mutation > 0.0f && schedule == SCHED_S2) ? (uint8_t *)malloc((size_t)rows * (size_t)m) : NULL;
    if (!keys || !mask || ((mutation > 0.0f && schedule == SCHED_S2) && !mut)) {
        free(keys); free(mask); free(mut);
        return -1;
    }
    orc_split(cross_key, rows, layout, keys);
    int err = 0;
#pragma omp parallel
    {
        uint32_t *scratch = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(m + 1));
        if (!scratch) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(dynamic, 4)
            for (int64_t q = 0; q < rows; ++q) {
                uint32_t krec[2] = {keys[2 * q], keys[2 * q + 1]}, kmut[2] = {0, 0};
                if (schedule == SCHED_S2) {
                    uint32_t ks[4];
                    orc_split(&keys[2 * q], 2, layout, ks);
                    krec[0] = ks[0]; krec[1] = ks[1];
                    kmut[0] = ks[2]; kmut[1] = ks[3];
                }
                orc_random_bits(krec, m, layout, scratch);
                uint8_t c = 0;
                for (int64_t j = 0; j < m; ++j) {
                    c ^= (uint8_t)(bits_to_uniform(scratch[j]) < r[j]);
                    mask[q * m + j] = c;
                }
                if (mut) {
                    orc_random_bits(kmut, m, layout, scratch);
                    for (int64_t j = 0; j < m; ++j) mut[q * m + j] = (uint8_t)(bits_to_uniform(scratch[j]) < mutation);
                }
            }
#pragma omp for schedule(static) collapse(2)
            for (int64_t e = 0; e < E; ++e)
                for (int64_t q = 0; q < rows; ++q) {
                    const int64_t i = q >> 1, p = q & 1;
                    const int64_t a = norm_index(actions[(e * n + i) * 2 + p], N);
                    const uint8_t *ind = pops + ((e * N + a) * m) * 2;
                    uint8_t *o = out + ((e * n + i) * m) * 2 + p;
                    const uint8_t *mk = mask + q * m;
                    if (mut) {
                        const uint8_t *mu = mut + q * m;
                        for (int64_t j = 0; j < m; ++j) o[2 * j] = ind[2 * j + mk[j]] ^ mu[j];
                    } else {
                        for (int64_t j = 0; j < m; ++j) o[2 * j] = ind[2 * j + mk[j]];
                    }
                }
            free(scratch);
        }
    }
    free(keys); free(mask); free(mut);
    return err;
}

/* TraitModel.__call__ in float64: pop bool[rows][m][2], effects f32[m][T] -> out f64[rows][T] */
void orc_gebv(const uint8_t *pop, const float *effects, int64_t rows, int64_t m, int64_t T, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i) {
        const uint8_t *g = pop + i * m * 2;
        for (int64_t t = 0; t < T; ++t) out[i * T + t] = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            int d = g[2 * j] + g[2 * j + 1];
            if (d)
                for (int64_t t = 0; t < T; ++t) out[i * T + t] += (double)d * (double)effects[j * T + t];
        }
    }
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Threefry-2x32 (20 rounds) and the jax.random key/bit layouts, host + device.
//
// Replaces jax/_src/prng.py threefry_2x32 / threefry_split / threefry_random_bits
// as reached from the reference through chromax.  Legacy
// layout: random_bits(key, n) pairs counter c with counter c + ceil(n/2) in ONE
// block (word 0 -> draw c, word 1 -> draw c + ceil(n/2); odd n pads a zero
// counter).  Partitionable layout: one block (0, j) per draw, output x0 ^ x1.
#pragma once
#include <stdint.h>

#include "../../include/breedgym_b200.h"

#if defined(__CUDACC__)
#define BG_HD __host__ __device__ __forceinline__
#else
#define BG_HD inline
#endif

struct TfKey {
    uint32_t k0, k1, k2;  // k2 = k0 ^ k1 ^ 0x1BD11BDA
};

BG_HD TfKey tf_make_key(uint32_t k0, uint32_t k1)
{
    TfKey k;
    k.k0 = k0;
    k.k1 = k1;
    k.k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    return k;
}

BG_HD uint32_t tf_rotl(uint32_t x, int r)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, x, r);
#else
    return (x << r) | (x >> (32 - r));
#endif
}

#define BG_TF_ROUND(x0, x1, r) \
    do {                       \
        x0 += x1;              \
        x1 = tf_rotl(x1, r);   \
        x1 ^= x0;              \
    } while (0)

BG_HD void tf2x32(const TfKey &k, uint32_t &x0, uint32_t &x1)
{
    x0 += k.k0;
    x1 += k.k1;
    BG_TF_ROUND(x0, x1, 13); BG_TF_ROUND(x0, x1, 15); BG_TF_ROUND(x0, x1, 26); BG_TF_ROUND(x0, x1, 6);
    x0 += k.k1; x1 += k.k2 + 1u;
    BG_TF_ROUND(x0, x1, 17); BG_TF_ROUND(x0, x1, 29); BG_TF_ROUND(x0, x1, 16); BG_TF_ROUND(x0, x1, 24);
    x0 += k.k2; x1 += k.k0 + 2u;
    BG_TF_ROUND(x0, x1, 13); BG_TF_ROUND(x0, x1, 15); BG_TF_ROUND(x0, x1, 26); BG_TF_ROUND(x0, x1, 6);
    x0 += k.k0; x1 += k.k1 + 3u;
    BG_TF_ROUND(x0, x1, 17); BG_TF_ROUND(x0, x1, 29); BG_TF_ROUND(x0, x1, 16); BG_TF_ROUND(x0, x1, 24);
    x0 += k.k1; x1 += k.k2 + 4u;
    BG_TF_ROUND(x0, x1, 13); BG_TF_ROUND(x0, x1, 15); BG_TF_ROUND(x0, x1, 26); BG_TF_ROUND(x0, x1, 6);
    x0 += k.k2; x1 += k.k0 + 5u;
}

// word j of random_bits(key, n)
BG_HD uint32_t tf_bits_at(const TfKey &k, uint64_t j, uint64_t n, int layout)
{
    uint32_t x0, x1;
    if (layout == BG_LAYOUT_PARTITIONABLE) {
        x0 = (uint32_t)(j >> 32);
        x1 = (uint32_t)j;
        tf2x32(k, x0, x1);
        return x0 ^ x1;
    }
    const uint64_t h = (n + 1) >> 1;
    if (j < h) {
        x0 = (uint32_t)j;
        x1 = (j + h < n) ? (uint32_t)(j + h) : 0u;
        tf2x32(k, x0, x1);
        return x0;
    }
    x0 = (uint32_t)(j - h);
    x1 = (uint32_t)j;
    tf2x32(k, x0, x1);
    return x1;
}

// key #q of split(key, num)
BG_HD TfKey tf_split_at(const TfKey &k, uint64_t q, uint64_t num, int layout)
{
    if (layout == BG_LAYOUT_PARTITIONABLE) {
        uint32_t x0 = (uint32_t)(q >> 32), x1 = (uint32_t)q;
        tf2x32(k, x0, x1);
        return tf_make_key(x0, x1);
    }
    const uint32_t a = tf_bits_at(k, 2 * q, 2 * num, BG_LAYOUT_LEGACY);
    const uint32_t b = tf_bits_at(k, 2 * q + 1, 2 * num, BG_LAYOUT_LEGACY);
    return tf_make_key(a, b);
}

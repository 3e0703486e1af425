// svr_rng.cuh -- the two random streams of the path tracer.
//
//  * XorwowCompat reproduces what the reference draws: curand_init(wangHash(frameNo) + pixelOffset,
//    0, 0) followed by curand_uniform (pathtracer.cu:70-79, 205-206, 302).  With subsequence 0 and
//    offset 0 cuRAND performs no skip-ahead, so the state is a closed form of the seed; only the
//    six state words are kept (cuRAND's 48-byte state also carries Box-Muller fields).
//  * Philox is the product stream: Philox2x32-7 (Salmon et al., SC'11), counter = (sample index |
//    draw block, f(pixel, seed)), fixed key.  A path is a pure function of (seed, pixel, sample), so an
//    image does not depend on launch shape, on how samples are batched, or on how they are split
//    across GPUs.
#pragma once

#include "svr_math.cuh"

namespace svr {

SVR_HD uint32_t wang_hash(uint32_t a)  // pathtracer.cu:70-79
{
    a = (a ^ 61u) ^ (a >> 16);
    a = a + (a << 3);
    a = a ^ (a >> 4);
    a = a * 0x27d4eb2du;
    a = a ^ (a >> 15);
    return a;
}

struct XorwowCompat {
    uint32_t v0, v1, v2, v3, v4, d;

    // stream for (frame, pixel): seed = wangHash(frameNo) + offset, as kernel_pathtracer seeds it
    SVR_DEV void init(uint32_t /*seedKey*/, uint32_t pixel, uint32_t sample)
    {
        uint32_t seed = wang_hash(sample) + pixel;
        uint32_t s0 = seed ^ 0xaad26b49u;
        uint32_t s1 = 0xf7dcefddu;  // high seed word is zero
        uint32_t t0 = 1099087573u * s0;
        uint32_t t1 = 2591861531u * s1;
        d = 6615241u + t1 + t0;
        v0 = 123456789u + t0;
        v1 = 362436069u ^ t0;
        v2 = 521288629u + t1;
        v3 = 88675123u ^ t1;
        v4 = 5783321u + t0;
    }
    SVR_DEV uint32_t next_u32()
    {
        uint32_t t = v0 ^ (v0 >> 2);
        v0 = v1;
        v1 = v2;
        v2 = v3;
        v3 = v4;
        v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
        d += 362437u;
        return v4 + d;
    }
    // curand_uniform: (0, 1]
    SVR_DEV float next() { return (float)next_u32() * 2.3283064e-10f + (2.3283064e-10f / 2.0f); }
    // 1 - u as the reference computes it before logf
    SVR_DEV float next_one_minus() { return 1.f - next(); }
};

struct Philox {
    uint32_t c0;   // low counter word: low 18 bits of the sample index in the high 18 bits, draw-block number in the low 14
    uint32_t c1;   // high counter word: pixel, seed and the sample index's bits 18..31
    uint32_t r1;   // second word of the current block
    uint32_t have; // 1 = r1 not yet consumed

    static constexpr uint32_t M = 0xD256D193u;  // Philox2x32 multiplier
    static constexpr uint32_t W = 0x9E3779B9u;  // Weyl key increment
    static constexpr uint32_t K = 0x5EED5EEDu;  // the (fixed) key
    static constexpr int BLOCK_BITS = 14;
// Rounds: 7.  Salmon et al. (SC'11) ship 10 rounds as a safety margin and name 7 rounds as the point where the 32-bit
// Philox family (Philox4x32-7) already passes TestU01's BigCrush; a renderer needs decorrelated streams, not a margin
// against future test batteries, and the three rounds are 5 % of the C3 kernel (9.56 -> 9.08 ms, round 2).  The
// evidence this library relies on is its own: every statistical parity test against the reference's kernels (per-tile
// Welch statistics, RMSE against the reference-vs-reference noise floor, image means) runs on this round count.
#ifndef SVR_PHILOX_ROUNDS
#define SVR_PHILOX_ROUNDS 7
#endif
    static constexpr int ROUNDS = SVR_PHILOX_ROUNDS;

    // Philox2x32 is a keyed bijection of the 64-bit counter.  The stream identity (pixel, seed, sample)
    // lives in the COUNTER and the key is a compile-time constant, so the round keys are immediates
    // instead of registers.  A path that draws more than 2^14 blocks runs on into the counter range
    // of the next sample index -- still deterministic, and far beyond what a path consumes.
    // Sample indices are 32 bits wide (frameNo of a progressive render, first_sample + rank * spp of a split): the low
    // 18 bits sit in c0, bits 18..31 are folded into c1 -- zero for the first 262144 samples, so those streams are what
    // they always were, and later samples get streams of their own instead of replaying sample mod 2^18.
    SVR_DEV void init(uint32_t seedKey, uint32_t pixel, uint32_t sample_)
    {
        c1 = (pixel * 0x9E3779B1u + seedKey) ^ ((sample_ >> (32 - BLOCK_BITS)) * 0x85EBCA6Bu);  // bijective in pixel for a fixed seed and epoch
        // keep the key in its register: under register pressure the compiler otherwise re-derives it from (pixel, seed, sample)
        // inside every generate() -- three extra instructions per block, +3 % on the whole kernel (ncu, C3 close view)
        asm volatile("" : "+r"(c1));
        c0 = sample_ << BLOCK_BITS;
        have = 0;
        r1 = 0;
    }
    SVR_DEV void generate(uint32_t& o0, uint32_t& o1)
    {
        uint32_t a = c0++, b = c1;
#pragma unroll
        for (int i = 0; i < ROUNDS; ++i) {
            uint32_t hi = __umulhi(M, a);
            uint32_t lo = M * a;
            a = hi ^ (K + (uint32_t)i * W) ^ b;
            b = lo;
        }
        o0 = a;
        o1 = b;
    }
    SVR_DEV uint32_t next_u32()
    {
        if (have) {
            have = 0;
            return r1;
        }
        uint32_t a;
        generate(a, r1);
        have = 1;
        return a;
    }
    // [0, 1): 24 random mantissa bits
    SVR_DEV float next() { return (float)(next_u32() >> 8) * 5.9604645e-8f; }
    // (0, 1]
    SVR_DEV float next_one_minus() { return 1.f - next(); }
    // two fresh uniforms from one block (drops a buffered word so call sites stay convergent)
    SVR_DEV void next2(float& a, float& b)
    {
        uint32_t x, y;
        generate(x, y);
        a = (float)(x >> 8) * 5.9604645e-8f;
        b = (float)(y >> 8) * 5.9604645e-8f;
    }
};

}  // namespace svr

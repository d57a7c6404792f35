// rng.h -- random numbers, host/device.
//
// (1) The reference's generators, kept for the primary-ray parity hook (udpt.cl:1131-1147, seeding :175-187).
// (2) The product's generator: Philox4x32-10 used as a COUNTER-BASED stream.  A draw is addressed by
//     (seed ; pixel, sample, path vertex, purpose) -- never by "how many numbers were drawn before" -- so
//     a sample's value does not depend on wavefront scheduling, on the spp split between calls, or on the
//     GPU that renders it.  The oracle port (oracle/yune_oracle.cpp) addresses its draws identically.
// Uniforms are mapped like the reference does it, u = (float)word / 2^32, i.e. u in [0, 1] INCLUSIVE
// (udpt.cl:185: 'seed / (float) UINT_MAX'; appendix B#6 of SURVEY.md).
#ifndef YUNE_RNG_H
#define YUNE_RNG_H

#include "strict_math.h"

namespace yune {

YUNE_HD uint32_t wang_hash(uint32_t seed)
{
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
YUNE_HD uint32_t xor_shift(uint32_t seed)
{
    seed ^= seed << 13; seed ^= seed >> 17; seed ^= seed << 5;
    return seed;
}
// udpt.cl:185 divides by (float)UINT_MAX = 2^32; x / 2^32 == x * 2^-32 bit for bit (power of two, no underflow for integer-valued x)
YUNE_HD float u01(uint32_t w) { return YF_MUL((float)w, 2.3283064365386963e-10f); }

YUNE_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

struct U4 { uint32_t x, y, z, w; };

// Philox4x32-10 (Salmon et al., SC'11), constants as published.
YUNE_HD_LEAF U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
    #pragma unroll
#endif
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
        U4 n; n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n; k0 += W0; k1 += W1;
    }
    return c;
}

#define YUNE_RNG_KEY1 0x59554e45u            /* "YUNE" */
#define YUNE_VERTEX_CAMERA 0xFFFFFFFFu       /* pseudo-vertex that owns the pixel jitter */
// purpose blocks of one path vertex (4 words each):
#define YUNE_BLK_BOUNCE 0u   /* x lobe selection, y,z direction sample, w Fresnel choice            */
#define YUNE_BLK_NEE    1u   /* x Russian roulette, y NEE lobe selection, z,w MIS direction sample   */
#define YUNE_BLK_LIGHT  2u   /* + light index: x,y point on the light; block 2 also: z light pick    */

YUNE_HD U4 draw4(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t vertex, uint32_t block)
{
    U4 c; c.x = pixel; c.y = sample; c.z = vertex; c.w = block;
    return philox4x32_10(c, seed, YUNE_RNG_KEY1);
}

} // namespace yune
#endif

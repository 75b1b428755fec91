// rcg_philox.h -- Philox4x32-10 phase source (stands in for rand::thread_rng at
// reference fft.rs:64, which is OS-seeded and irreproducible).  One block per bin
// quadruple {k, W-k, H-k, H+k}; counter = (k, frame_lo, frame_hi, channel), key = seed.
#pragma once
#include "rcg_common.h"

namespace rcg {

struct PhiloxKeys { uint32_t k0[10], k1[10]; };  // round keys, precomputed on the host

inline PhiloxKeys make_philox_keys(uint64_t seed) {
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { k.k0[r] = a; k.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    return k;
}

RCG_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

RCG_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          const PhiloxKeys& key, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c0, h0, l0);
        mulhilo(0xCD9E8D57u, c2, h1, l1);
        uint32_t n0 = h1 ^ c1 ^ key.k0[r], n2 = h0 ^ c3 ^ key.k1[r];
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// rand 0.8.5 gen_range(0.0..PI) on f32 (fft.rs:13: the constant named TWO_PI is PI):
// ((u >> 9) * 2^-23) * PI + 0.0
// The scaling by 2^-23 is exact, so one product with the constant 2^-23 * PI (itself exact) rounds to the same value.
// On the device the way rand itself gets there: v = u >> 9 becomes the mantissa of x = 1 + v 2^-23 in [1, 2) -- ONE funnel
// shift, (0x7F : u) >> 9 = 0x3F800000 | v -- and x PI - PI as one fused multiply-add is the exact product v 2^-23 PI
// rounded once: the same value bit for bit (tests/test_oracle_crosscheck.py), two instructions instead of three
// (shift, int-to-float, multiply) for each of the W phases of a frame.
RCG_HD float phase_from_u32(uint32_t u) {
#if defined(RCG_PHASE_TWO_PRODUCTS)
    return (float)(u >> 9) * 1.1920928955078125e-07f * 3.14159274101257324f;
#elif defined(__CUDA_ARCH__)
    return __fmaf_rn(__uint_as_float(__funnelshift_r(u, 0x7Fu, 9)), 3.14159274101257324f, -3.14159274101257324f);
#else
    return (float)(u >> 9) * (1.1920928955078125e-07f * 3.14159274101257324f);
#endif
}

}  // namespace rcg

// rcg_dft.h -- fully unrolled radix-2/4/8/16 DFTs on register arrays (SoA re/im).
// Forward sign (e^{-2 pi i ak/R}); the inverse is obtained by calling with re/im swapped.
#pragma once
#include "rcg_common.h"

namespace rcg {

// multiply (re, im) by w_16^K16 = exp(-2 pi i K16/16), K16 in 0..7
template <int K16>
RCG_HD void mul_w16(float& re, float& im) {
    constexpr float C1 = 0.92387953251128674f;  // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;  // sin(pi/8)
    constexpr float R2 = 0.70710678118654752f;  // sqrt(1/2)
    if constexpr (K16 == 0) {
    } else if constexpr (K16 == 4) {  // -j
        float t = re; re = im; im = -t;
    } else if constexpr (K16 == 2) {  // (1 - j)/sqrt2
        float a = re + im, b = im - re; re = a * R2; im = b * R2;
    } else if constexpr (K16 == 6) {  // (-1 - j)/sqrt2
        float a = im - re, b = -(re + im); re = a * R2; im = b * R2;
    } else {
        constexpr float c = (K16 == 1) ? C1 : (K16 == 3) ? S1 : (K16 == 5) ? -S1 : -C1;
        constexpr float s = (K16 == 1 || K16 == 7) ? S1 : C1;  // sin(K16*pi/8) >= 0
        // (re + j im)(c - j s)
        float nr = re * c + im * s, ni = im * c - re * s;
        re = nr; im = ni;
    }
}

template <int R>
struct Dft {
    // y[k] = sum_a x[a] w_R^{ak}; x is consumed
    static RCG_HD void run(float (&xr)[R], float (&xi)[R]) {
        constexpr int Hh = R / 2;
        float er[Hh], ei[Hh], orr[Hh], oi[Hh];
#pragma unroll
        for (int a = 0; a < Hh; a++) { er[a] = xr[2 * a]; ei[a] = xi[2 * a]; orr[a] = xr[2 * a + 1]; oi[a] = xi[2 * a + 1]; }
        Dft<Hh>::run(er, ei);
        Dft<Hh>::run(orr, oi);
        twiddle_and_combine<0>(er, ei, orr, oi, xr, xi);
    }
    template <int K>
    static RCG_HD void twiddle_and_combine(float (&er)[R / 2], float (&ei)[R / 2], float (&orr)[R / 2],
                                           float (&oi)[R / 2], float (&yr)[R], float (&yi)[R]) {
        if constexpr (K < R / 2) {
            float tr = orr[K], ti = oi[K];
            mul_w16<K * (16 / R)>(tr, ti);
            yr[K] = er[K] + tr; yi[K] = ei[K] + ti;
            yr[K + R / 2] = er[K] - tr; yi[K + R / 2] = ei[K] - ti;
            twiddle_and_combine<K + 1>(er, ei, orr, oi, yr, yi);
        }
    }
};
template <>
struct Dft<1> {
    static RCG_HD void run(float (&)[1], float (&)[1]) {}
};

}  // namespace rcg

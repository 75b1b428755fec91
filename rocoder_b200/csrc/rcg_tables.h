// rcg_tables.h -- host-side builders for the constant tables the kernels read:
// Hann window (windows.rs:4-9), crossfade envelope (crossfade.rs:4-10), FFT twiddles.
// Window and envelope use the reference's f32 expressions evaluated with the host libm, so
// the values match what rocoder computes on the same machine; twiddles are rounded from
// double.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "rcg_common.h"

namespace rcg {

inline void host_hanning(float* out, size_t len) {
    const float two_pi = 3.14159274101257324f * 2.0f;  // f32::consts::PI * 2.0
    for (size_t i = 0; i < len; i++)
        out[i] = 0.5f - (cosf(((float)i * two_pi) / (float)(len - 1)) * 0.5f);
}

inline void host_crossfade_compensation(float* out, size_t len) {
    const float two_pi = 3.14159274101257324f * 2.0f;
    const float hinv_sqrt2 = (1.0f + sqrtf(sqrtf(0.5f))) * 0.5f;
    for (size_t i = 0; i < len; i++)
        out[i] = 0.5f - ((1.0f - hinv_sqrt2) * cosf(((float)i * two_pi) / (float)(len - 1)));
}

inline int rt_pos_to_bin(const PlanDesc& d, int pos) {
    int k = 0;
    for (int p = 0; p < d.npasses; p++)
        k |= ((pos >> d.mlog2[p]) & ((1 << d.rlog2[p]) - 1)) << d.klo[p];
    return k;
}
inline int rt_bin_to_pos(const PlanDesc& d, int k) {
    int pos = 0;
    for (int p = 0; p < d.npasses; p++)
        pos |= ((k >> d.klo[p]) & ((1 << d.rlog2[p]) - 1)) << d.mlog2[p];
    return pos;
}

// packed per-pass twiddles in the layout rcg_common.h's tw_scheme() prescribes: the full plan's
// tables followed by the half-rate synthesis plan's
inline std::vector<cf> host_pass_twiddles(int log2h) {
    std::vector<cf> tw((size_t)std::max(tw_total(log2h), 1));
    const double two_pi = 6.283185307179586476925286766559;
    auto w = [&](double num, double den) {
        double a = -two_pi * num / den;
        return make_cf((float)cos(a), (float)sin(a));
    };
    for (int hr = 0; hr < 2; hr++) {
        const PlanDesc d = make_plan(log2h, hr != 0);
        for (int p = 0; p < d.npasses; p++) {
            const int M = 1 << d.mlog2[p], R = 1 << d.rlog2[p];
            const double N = (double)R * (double)M;
            cf* t = tw.data() + tw_offset(log2h, p, hr != 0);
            switch (tw_scheme(log2h, p, hr != 0)) {
                case kTwFull:
                    for (int j = 1; j < R; j++)
                        for (int m = 0; m < M; m++) t[(j - 1) * M + m] = w((double)m * j, N);
                    break;
                case kTwPow:
                    for (int i = 0; i < d.rlog2[p]; i++)
                        for (int m = 0; m < M; m++) t[i * M + m] = w((double)m * (1 << i), N);
                    break;
                case kTwTwoLevel:
                    for (int hi = 0; hi < M / 64; hi++) t[hi] = w((double)hi * 64, N);
                    for (int lo = 0; lo < 64; lo++) t[M / 64 + lo] = w((double)lo, N);
                    break;
                default: break;
            }
        }
    }
    return tw;
}

// exp(-2 pi i k / W) for the bin k the spectral step's pair slot c owns
inline std::vector<cf> host_spectral_twiddles(int log2h) {
    const PlanDesc d = make_plan(log2h);
    const int H = 1 << log2h;
    std::vector<cf> tw((size_t)H / 2);
    const double two_pi = 6.283185307179586476925286766559;
    for (int c = 0; c < H / 2; c++) {
        int k = rt_pos_to_bin(d, spectral_slot_pos(log2h, c));
        double a = -two_pi * (double)k / (double)(2 * H);
        tw[c] = make_cf((float)cos(a), (float)sin(a));
    }
    return tw;
}

// half-rate synthesis: exp(-2 pi i k / W) for the bin k the quad slot c owns (k < W/8 ... in the
// spectral step's order: c -> (rho, kl), kl < 4, position rho*16 + kl)
inline std::vector<cf> host_spectral_twiddles_hr(int log2h) {
    const PlanDesc d = make_plan(log2h);
    const int H = 1 << log2h;
    std::vector<cf> tw((size_t)std::max(H / 4, 1));
    const double two_pi = 6.283185307179586476925286766559;
    for (int c = 0; c < H / 4; c++) {
        int k = rt_pos_to_bin(d, spectral_slot_pos_hr(log2h, c));
        double a = -two_pi * (double)k / (double)(2 * H);
        tw[c] = make_cf((float)cos(a), (float)sin(a));
    }
    return tw;
}

// ---- W = 65536 (CTA pair, rcg_frame_pair.h) ---------------------------------------------------
// two-level table for w_H^n = exp(-2 pi i n / H), n < H/2: A[n >> 6] then B[n & 63]
inline std::vector<cf> host_pair_cross_twiddles(int log2h) {
    const int H = 1 << log2h, HL = H / 2;
    std::vector<cf> tw((size_t)HL / 64 + 64);
    const double two_pi = 6.283185307179586476925286766559;
    for (int a = 0; a < HL / 64; a++) { double x = -two_pi * (double)(a * 64) / H; tw[a] = make_cf((float)cos(x), (float)sin(x)); }
    for (int b = 0; b < 64; b++) { double x = -two_pi * (double)b / H; tw[HL / 64 + b] = make_cf((float)cos(x), (float)sin(x)); }
    return tw;
}
// exp(-2 pi i (2j + odd) / W) for the local bin j the pair slot c of the half-size plan owns
inline std::vector<cf> host_pair_spectral_twiddles(int log2h, int odd) {
    const int log2hl = log2h - 1;
    const PlanDesc d = make_plan(log2hl);
    const int HL = 1 << log2hl;
    std::vector<cf> tw((size_t)HL / 2);
    const double two_pi = 6.283185307179586476925286766559;
    for (int c = 0; c < HL / 2; c++) {
        int j = rt_pos_to_bin(d, spectral_slot_pos(log2hl, c));
        double a = -two_pi * (double)(2 * j + odd) / (double)(4 * HL);
        tw[c] = make_cf((float)cos(a), (float)sin(a));
    }
    return tw;
}

// ---- windows that are not a power of two: Bluestein's chirp tables (rcg_frame_bluestein.h) ----------------------
// smallest LM with 2^LM >= 2W - 1
inline int bluestein_log2m(int W) {
    int lm = 0;
    while ((1 << lm) < 2 * W - 1) lm++;
    return lm;
}
// c[n] = exp(i pi n^2 / W): the exponent is reduced exactly (n^2 mod 2W) before it meets floating point
inline std::vector<cf> host_bluestein_chirp(int W) {
    std::vector<cf> c((size_t)W);
    const double pi = 3.14159265358979323846264338327950288;
    for (long long n = 0; n < W; n++) {
        const long long r = (n * n) % (2LL * W);
        const double a = pi * (double)r / (double)W;
        c[(size_t)n] = make_cf((float)cos(a), (float)sin(a));
    }
    return c;
}
// FFT_M(b) with b[n] = b[M - n] = c[n] for n < W (0 elsewhere), natural bin order, in double
inline void host_bluestein_filter_spectrum(int W, int lm, std::vector<double>& re, std::vector<double>& im) {
    const size_t M = (size_t)1 << lm;
    const double pi = 3.14159265358979323846264338327950288;
    re.assign(M, 0.0); im.assign(M, 0.0);
    for (long long n = 0; n < W; n++) {
        const long long r = (n * n) % (2LL * W);
        const double a = pi * (double)r / (double)W;
        re[(size_t)n] = cos(a); im[(size_t)n] = sin(a);
        if (n) { re[M - (size_t)n] = cos(a); im[M - (size_t)n] = sin(a); }
    }
    // in-place radix-2 decimation-in-time FFT in double (a table built once per engine)
    for (size_t i = 1, j = 0; i < M; i++) {
        size_t bit = M >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    std::vector<double> wr(M / 2), wi(M / 2);
    for (size_t k = 0; k < M / 2; k++) {
        const double a = -2.0 * pi * (double)k / (double)M;
        wr[k] = cos(a); wi[k] = sin(a);
    }
    for (size_t len = 2; len <= M; len <<= 1) {
        const size_t stride = M / len;
        for (size_t i = 0; i < M; i += len)
            for (size_t k = 0; k < len / 2; k++) {
                const double cr = wr[k * stride], ci = wi[k * stride];
                const size_t u = i + k, v = i + k + len / 2;
                const double tr = re[v] * cr - im[v] * ci, ti = re[v] * ci + im[v] * cr;
                re[v] = re[u] - tr; im[v] = im[u] - ti;
                re[u] += tr; im[u] += ti;
            }
    }
}
// ... / M, in the position order the DIF passes of plan LM leave (in-SM chirp-z kernels, rcg_frame_bluestein.h)
inline std::vector<cf> host_bluestein_filter(int W, int lm) {
    const int M = 1 << lm;
    std::vector<double> re, im;
    host_bluestein_filter_spectrum(W, lm, re, im);
    const PlanDesc d = make_plan(lm);
    std::vector<cf> f((size_t)M);
    for (int pos = 0; pos < M; pos++) {
        const int k = rt_pos_to_bin(d, pos);
        f[(size_t)pos] = make_cf((float)(re[(size_t)k] / M), (float)(im[(size_t)k] / M));
    }
    return f;
}

// frame-invariant index tables of the spectral steps: padded shared-memory slot indices of the bins
// a pair / quad thread touches, and the pair's bin number
inline std::vector<uint2> host_pair_table(int log2h) {
    const PlanDesc d = make_plan(log2h);
    const int H = 1 << log2h;
    std::vector<uint2> tab((size_t)H / 2);
    for (int c = 0; c < H / 2; c++) {
        const int pos = spectral_slot_pos(log2h, c), k = rt_pos_to_bin(d, pos);
        const unsigned pa = (unsigned)phys(pos), pb = (unsigned)phys(rt_bin_to_pos(d, (H - k) % H));
        tab[c].x = pa | (pb << 16);
        tab[c].y = (unsigned)k;
    }
    return tab;
}
inline std::vector<uint4> host_quad_table(int log2h) {
    const PlanDesc d = make_plan(log2h);
    const int H = 1 << log2h, Q = H / 2;
    std::vector<uint4> tab((size_t)std::max(H / 4, 1));
    for (int c = 0; c < H / 4; c++) {
        const int pos = spectral_slot_pos_hr(log2h, c), k = rt_pos_to_bin(d, pos);
        const unsigned pa = (unsigned)phys(pos), pb = (unsigned)phys(rt_bin_to_pos(d, (H - k) % H));
        const unsigned pc = (unsigned)phys(rt_bin_to_pos(d, (Q - k + H) % H)), pd = (unsigned)phys(rt_bin_to_pos(d, Q + k));
        tab[c].x = pa | (pb << 16);
        tab[c].y = pc | (pd << 16);
        tab[c].z = (unsigned)k;
        tab[c].w = 0;
    }
    return tab;
}

// ---- pass twiddles as Tensor Memory rows (tw_in_tmem, rcg_common.h): image[column][thread] ------------------------
// half_rate: which spectral-step table entries follow the twiddles (spectral_in_tmem): the quads of the half-rate path
// or the pairs of the full-rate one
// window / env (full-rate image only, window_in_tmem): the engine's analysis / synthesis window [W] and envelope [H]
inline std::vector<float> host_tmem_image(int log2h, bool half_rate = false, const float* window = nullptr,
                                          const float* env = nullptr) {
    const int NT = threads_per_frame(log2h), ncol = tmem_image_cols(log2h);
    std::vector<float> img((size_t)ncol * NT, 0.0f);
    const double two_pi = 6.283185307179586476925286766559;
    for (int hr = 0; hr < 2; hr++) {
        if (!tw_in_tmem(log2h, hr != 0)) continue;
        const PlanDesc d = make_plan(log2h, hr != 0);
        for (int p = 0; p < d.npasses; p++) {
            const int R = 1 << d.rlog2[p], sc = tmem_set_cols(log2h, p, hr != 0), off = tmem_col_offset(log2h, p, hr != 0);
            if (sc == 0) continue;
            const double n = (double)R * (double)(1 << d.mlog2[p]);
            for (int t = 0; t < NT; t++)
                for (int set = 0; set < tmem_sets(log2h, p, hr != 0); set++) {
                    const int m = tmem_m_of_thread(log2h, p, hr != 0, t, set);
                    for (int j = (R == 2 ? 1 : 0); j < R; j++) {
                        const double a = -two_pi * (double)m * (double)j / n;
                        const int col = off + set * sc + (R == 2 ? 0 : 2 * j);
                        img[(size_t)col * NT + t] = (float)cos(a);
                        img[(size_t)(col + 1) * NT + t] = (float)sin(a);
                    }
                }
        }
    }
    if (window_in_tmem(log2h) && !half_rate && window && env) {
        // over the half-rate plan's columns (a full-rate launch never reads them): iteration i of the load / overlap-add
        // loops of thread t works on packed index m = t + NT i
        const int H = 1 << log2h;
        for (int t = 0; t < NT; t++)
            for (int i = 0; i < tmem_window_iters(log2h); i++) {
                const int m = t + NT * i, wc = tmem_window_offset(log2h) + 4 * i, ec = tmem_env_col(log2h, i);
                img[(size_t)(wc + 0) * NT + t] = window[2 * m];
                img[(size_t)(wc + 1) * NT + t] = window[2 * m + 1];
                img[(size_t)(wc + 2) * NT + t] = window[H + 2 * m];
                img[(size_t)(wc + 3) * NT + t] = window[H + 2 * m + 1];
                img[(size_t)(ec + 0) * NT + t] = env[2 * m];
                img[(size_t)(ec + 1) * NT + t] = env[2 * m + 1];
            }
    }
    if (spectral_in_tmem(log2h)) {
        auto put = [&](int col, int t, uint32_t bits) { memcpy(&img[(size_t)col * NT + t], &bits, sizeof(float)); };
        auto putf = [&](int col, int t, float f) { img[(size_t)col * NT + t] = f; };
        const int off = tmem_spectral_offset(log2h), H = 1 << log2h;
        if (half_rate) {
            const std::vector<uint4> qt = host_quad_table(log2h);
            const std::vector<cf> tw = host_spectral_twiddles_hr(log2h);
            for (int c = 0; c < H / 4; c++) {
                const int t = c % NT, col = off + 8 * (c / NT);
                put(col, t, qt[c].x); put(col + 1, t, qt[c].y); put(col + 2, t, qt[c].z);
                putf(col + 3, t, tw[c].x); putf(col + 4, t, tw[c].y);
            }
        } else {
            const std::vector<uint2> pt = host_pair_table(log2h);
            const std::vector<cf> tw = host_spectral_twiddles(log2h);
            for (int c = 0; c < H / 2; c++) {
                const int t = c % NT, col = off + 4 * (c / NT);
                put(col, t, pt[c].x); put(col + 1, t, pt[c].y);
                putf(col + 2, t, tw[c].x); putf(col + 3, t, tw[c].y);
            }
        }
    }
    return img;
}

// even samples of a table (half-rate synthesis reads window[2i], env[2i])
inline std::vector<float> host_decimate2(const std::vector<float>& v) {
    std::vector<float> o((v.size() + 1) / 2);
    for (size_t i = 0; i < o.size(); i++) o[i] = v[2 * i];
    return o;
}

}  // namespace rcg

// rcg_frame_bluestein.h -- windows that are not a power of two (reference README.md:100: "powers of 2 are
// recommended", i.e. any length is legal; fft.rs:27-29 plans whatever length it is given).
//
// The reference runs a W-point complex FFT, |X| e^{i phi} for all W bins and a W-point inverse (fft.rs:50-74).  For a W
// that is not a power of two both transforms are done with Bluestein's chirp-z identity on the power-of-two
// in-shared-memory passes of rcg_frame.h:
//     nk = (n^2 + k^2 - (k - n)^2) / 2,   c[n] = exp(i pi n^2 / W)
//     X[k] = conj(c[k]) * sum_n (x[n] conj(c[n])) c[k - n]
// i.e. one circular convolution of length M = 2^LM >= 2W - 1 with the fixed chirp filter: DIF passes (natural ->
// digit-reversed order), a pointwise product with the filter's spectrum stored in that same order, DIT passes back to
// natural order.  The inverse transform is the same routine on conjugated input (IDFT(Y) = conj(DFT(conj Y))).
// One frame costs four M-point complex transforms, M ~ 2..4 W: this path is about correctness for every legal
// window, not speed.  Everything after the transforms -- /W * window, overlap-add with the carried half window,
// envelope * amp, integer resample -- is the reference's arithmetic in the reference's order (stretcher.rs:96-112,
// resampler.rs:3-35), with H = W / 2 a run-time value.  Even W only: for an odd W the reference appends W - H = H + 1
// samples per frame while advancing by H (stretcher.rs:101-103), a drift nobody can want; rejected with RCG_ERR_WINDOW.
#pragma once
#include "rcg_frame.h"

namespace rcg {

struct BluesteinTables {
    const cf* chirp;     // [W]  c[n] = exp(i pi n^2 / W)
    const cf* filter;    // [M]  FFT_M of b (b[n] = b[M - n] = c[n], n < W; 0 elsewhere), in the position order the
                         //      DIF passes leave, pre-divided by M
    int w;               // window length W (even, not a power of two)
};

// A[n] *= filter[n] over the whole transform (position order on both sides)
template <int LM>
RCG_HD void bs_phase_filter(cf* S, int t, const BluesteinTables& bt) {
    using C = FrameCfg<LM>;
    for (int i = t; i < C::H; i += C::NT) {
        const cf a = S[phys(i)], b = ld_cf(bt.filter + i);
        S[phys(i)] = make_cf(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
}

// all passes of the M-point transform, one barrier-separated phase each
template <int LM, bool INV, int PASS, class Exec>
RCG_HD void bs_fft(Exec& ex, cf* S, const StretchParams& prm, bool work) {
    constexpr int NP = FrameCfg<LM>::NP;
    if constexpr (PASS >= 0 && PASS < NP) {
        ex.par([&](int t) { if (work) phase_fft_pass<LM, PASS, INV, false>(S, t, prm.tw); });
        bs_fft<LM, INV, INV ? PASS - 1 : PASS + 1>(ex, S, prm, work);
    }
}

// circular convolution of the M-point sequence in S with the chirp filter (result in natural order, scaled by 1)
template <int LM, class Exec>
RCG_HD void bs_convolve(Exec& ex, cf* S, const StretchParams& prm, const BluesteinTables& bt, bool work) {
    bs_fft<LM, false, 0>(ex, S, prm, work);
    ex.par([&](int t) { if (work) bs_phase_filter<LM>(S, t, bt); });
    bs_fft<LM, true, FrameCfg<LM>::NP - 1>(ex, S, prm, work);
}

// One frame.  Leaves y[n] = (Re(IFFT(|X| e^{i phi}))[n] / W) * w[n] (fft.rs:70-73) in S[phys(n)].x, n < W.
template <int LM, class Exec>
RCG_HD void bs_resynth(Exec& ex, cf* S, const StretchParams& prm, const BluesteinTables& bt, int channel, long long frame,
                       bool work) {
    using C = FrameCfg<LM>;
    const int W = bt.w;
    const float* in = prm.in + (long long)channel * prm.in_stride;
    const long long base = frame * (long long)prm.step - prm.in_origin;
    const bool bypass = prm.flags & kFlagBypass;
    if (bypass) {   // test hook: resynth(s)[n] = (s[n] w[n]) w[n]
        ex.par([&](int t) {
            if (!work) return;
            for (int n = t; n < W; n += C::NT) {
                const long long i = base + n;
                const float s = (i >= 0 && i < prm.in_valid) ? ld_stream(in + i) : 0.0f;
                const float w = prm.window[n];
                S[phys(n)] = make_cf(RCG_MUL(RCG_MUL(s, w), w), 0.0f);
            }
        });
        return;
    }
    // a[n] = (s[n] w[n]) conj(c[n]), zero padded to M (fft.rs:51-58)
    ex.par([&](int t) {
        if (!work) return;
        for (int n = t; n < C::H; n += C::NT) {
            cf v = make_cf(0.0f, 0.0f);
            if (n < W) {
                const long long i = base + n;
                const float s = (i >= 0 && i < prm.in_valid) ? ld_stream(in + i) : 0.0f;
                const float x = RCG_MUL(s, prm.window[n]);
                const cf c = ld_cf(bt.chirp + n);
                v = make_cf(x * c.x, -(x * c.y));
            }
            S[phys(n)] = v;
        }
    });
    bs_convolve<LM>(ex, S, prm, bt, work);
    // X[k] = conj(c[k]) conv[k]; only |X[k]| = |conv[k]| survives (fft.rs:67).  Y[k] = |X[k]| e^{i phi_k}; the inverse
    // transform runs on conj(Y): a'[k] = conj(Y[k]) conj(c[k]), zero padded
    ex.par([&](int t) {
        if (!work) return;
        const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
        const uint32_t chan = (uint32_t)(channel + prm.channel_offset);
        const int H = W / 2, Q = H / 2;
        for (int k = t; k < C::H; k += C::NT) {
            cf v = make_cf(0.0f, 0.0f);
            if (k < W) {
                const cf x = S[phys(k)];
                const float mag = fast_sqrt(x.x * x.x + x.y * x.y);
                // the Philox block / word that serves bin k: the same assignment as the power-of-two kernels
                // (rcg_phase): pair = k, H-k, k-H, W-k -> words 0, 2, 3, 1
                uint32_t pair, word;
                if (k <= Q) { pair = (uint32_t)k; word = 0; }
                else if (k <= H) { pair = (uint32_t)(H - k); word = 2; }
                else if (k < H + Q) { pair = (uint32_t)(k - H); word = 3; }
                else { pair = (uint32_t)(W - k); word = 1; }
                uint32_t u[4];
                philox4x32_10(pair, f_lo, f_hi, chan, prm.keys, u);
                const cf e = cis_phase(u[word]);
                const cf c = ld_cf(bt.chirp + k);
                // conj(Y) conj(c) = mag conj(e c)
                v = make_cf(mag * (e.x * c.x - e.y * c.y), -(mag * (e.x * c.y + e.y * c.x)));
            }
            S[phys(k)] = v;
        }
    });
    bs_convolve<LM>(ex, S, prm, bt, work);
    // y[n] = Re(conj(conj(c[n]) conv[n])) = Re(c[n] conj(conv[n])) = c.x conv.x + c.y conv.y; then / W * w[n]
    ex.par([&](int t) {
        if (!work) return;
        const float wl = (float)W;
        for (int n = t; n < W; n += C::NT) {
            const cf v = S[phys(n)], c = ld_cf(bt.chirp + n);
            const float re = v.x * c.x + v.y * c.y;
            S[phys(n)] = make_cf(RCG_MUL(RCG_DIV(re, wl), prm.window[n]), 0.0f);
        }
    });
}

// integer resample of the finished block z[0 .. S) held as floats (resampler.rs:3-35); frame-relative indices as in
// phase_resample, H a run-time value
template <int LM>
RCG_HD void bs_resample(const float* z, int t, const StretchParams& prm, int H, int channel, long long frame) {
    using C = FrameCfg<LM>;
    float* out = prm.out + (long long)channel * prm.out_stride;
    const int p = prm.pitch;
    if (p == 1) {
        for (int i = t; i < H; i += C::NT) out[frame * (long long)H + i - prm.out_origin] = z[i];
    } else if (p > 1) {
        const long long g0 = frame * (long long)H;
        const long long o_begin = (g0 + p - 1) / p, o_end = (g0 + H + p - 1) / p;
        for (long long o = o_begin + t; o < o_end; o += C::NT) out[o - prm.out_origin] = z[(int)(o * p - g0)];
    } else {
        const int q = -p;
        const int n = (prm.samples_needed - 1) * q;
        float* dst = out + (frame * (long long)n - prm.out_origin);
        for (int tt = t; tt < n; tt += C::NT) {
            const int i = tt / q, j = tt - i * q;
            const float cur = z[i], nxt = z[i + 1];
            dst[tt] = RCG_ADD(cur, RCG_MUL(RCG_SUB(nxt, cur), prm.lerp_ratio[j]));
        }
    }
}

// the frame loop of one work item: same roles as run_work_item (halo frame, tail_in / tail_out, deferred first block)
template <int LM, class Exec>
RCG_HD void run_work_item_bluestein(Exec& ex, cf* S, float* zbuf, const StretchParams& prm, const BluesteinTables& bt,
                                    int item, float* tail, bool active) {
    using C = FrameCfg<LM>;
    const int W = bt.w, H = W / 2;
    const int channel = active ? item / prm.n_runs : 0;
    const int run = active ? item - channel * prm.n_runs : 0;
    const long long f0 = prm.frame_begin + (long long)run * prm.run_len;
    long long f1 = f0 + prm.run_len;
    if (f1 > prm.frame_end) f1 = prm.frame_end;
    const bool first_run = (run == 0);
    const bool defer = first_run && (prm.flags & kFlagDeferFirst) && f0 > 0;
    const bool have_tail_in = first_run && prm.tail_in != nullptr && f0 > 0;
    const bool halo = active && f0 > 0 && !defer && !have_tail_in;
    ex.par([&](int t) {
        if (!active) return;
        const float* src = have_tail_in ? prm.tail_in + (long long)channel * H : nullptr;
        for (int i = t; i < H; i += C::NT) tail[i] = src ? src[i] : 0.0f;
    });
    const int n_iter = prm.run_len + 1;
    for (int it = 0; it < n_iter; it++) {
        const long long f = f0 - 1 + it;
        const bool is_halo = (it == 0);
        const bool work = active && (is_halo ? halo : f < f1);
        bs_resynth<LM>(ex, S, prm, bt, channel, f, work);
        const bool head = work && defer && f == f0;
        const bool emit = work && !is_halo && !head;
        // z[i] = ((y[i] + carried[i]) * env[i]) * amp, carried <- y[H + i]  (stretcher.rs:96-104)
        ex.par([&](int t) {
            if (!work) return;
            for (int i = t; i < H; i += C::NT) {
                const float y = S[phys(i)].x, tl = S[phys(H + i)].x;
                const float old = tail[i];
                tail[i] = tl;
                if (head) prm.head_out[(long long)channel * H + i] = y;
                if (emit) zbuf[i] = RCG_MUL(RCG_MUL(RCG_ADD(y, old), prm.env[i]), prm.amp);
            }
        });
        ex.par([&](int t) { if (emit) bs_resample<LM>(zbuf, t, prm, H, channel, f); });
    }
    if (prm.tail_out != nullptr) {
        ex.par([&](int t) {
            if (!active || f1 != prm.frame_end) return;
            float* dst = prm.tail_out + (long long)channel * H;
            for (int i = t; i < H; i += C::NT) dst[i] = tail[i];
        });
    }
}

// boundary block of a time shard (rcg_finish_boundary): ((head + tail_prev) * env) * amp, resampled
template <int LM, class Exec>
RCG_HD void run_boundary_bluestein(Exec& ex, float* zbuf, const StretchParams& prm, const BluesteinTables& bt, int channel,
                                   const float* head, const float* tail_prev) {
    using C = FrameCfg<LM>;
    const int H = bt.w / 2;
    ex.par([&](int t) {
        const float* h = head + (long long)channel * H;
        const float* tp = tail_prev + (long long)channel * H;
        for (int i = t; i < H; i += C::NT) zbuf[i] = RCG_MUL(RCG_MUL(RCG_ADD(h[i], tp[i]), prm.env[i]), prm.amp);
    });
    ex.par([&](int t) { bs_resample<LM>(zbuf, t, prm, H, channel, prm.frame_begin); });
}

}  // namespace rcg

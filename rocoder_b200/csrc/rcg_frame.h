// rcg_frame.h -- the per-frame phases of the fused stretch kernel, written once and
// executed either by a CUDA thread (rcg_kernels.cu) or, thread by thread, by the
// host-side emulator the CPU tests use to check indexing without a GPU
// (tests/emul/rcg_emul.cpp; test infrastructure, never on a product path).
//
// One frame = the reference's ReFFT::resynth (fft.rs:42-74) followed by the half-window
// overlap-add, envelope, amplitude and integer resample of Stretcher::next_window
// (stretcher.rs:96-112, resampler.rs:3-35).
#pragma once
#include "rcg_common.h"
#include "rcg_dft.h"
#include "rcg_philox.h"

#if defined(__CUDA_ARCH__)
#define RCG_MUL(a, b) __fmul_rn((a), (b))
#define RCG_ADD(a, b) __fadd_rn((a), (b))
#define RCG_SUB(a, b) __fsub_rn((a), (b))
#define RCG_DIV(a, b) __fdiv_rn((a), (b))
#else
#define RCG_MUL(a, b) ((a) * (b))
#define RCG_ADD(a, b) ((a) + (b))
#define RCG_SUB(a, b) ((a) - (b))
#define RCG_DIV(a, b) ((a) / (b))
#endif

namespace rcg {

enum : uint32_t {
    kFlagBypass = 1u,      // test hook: resynth(s)[n] = (s[n]*w[n])*w[n]  (no FFT, no phases)
    kFlagDeferFirst = 2u,  // do not emit block frame_begin; write its raw first half to head_out
};

// Everything one launch needs.  Global indices (samples, frames, output positions) are
// absolute within the channel so that segments of a long input computed by different
// launches / GPUs produce identical values.
struct StretchParams {
    const float* in;        // [channels][in_stride]; in[c][0] is global sample in_origin
    long long in_stride;
    long long in_origin;
    long long in_valid;     // samples present from in_origin on; reads beyond are zero padding
    float* out;             // [channels][out_stride]; out[c][0] is global output index out_origin
    long long out_stride;
    long long out_origin;
    const float* window;    // [W]   windows::hanning (host-computed, windows.rs:4-9)
    const float* env;       // [H]   crossfade::hanning_crossfade_compensation (crossfade.rs:4-10)
    const cf* tw;           // packed per-pass twiddles
    const cf* spec_tw;      // [H/2] exp(-2 pi i k/W) in the spectral step's pair order
    float* tail_scratch;    // [work lanes][H] carried second half window (stretcher.rs:101-102)
    const float* tail_in;   // optional [channels][H]: second half of frame frame_begin-1
    float* tail_out;        // optional [channels][H]: second half of frame frame_end-1
    float* head_out;        // optional [channels][H]: raw first half of frame frame_begin
    long long frame_begin;  // global frame range [begin, end) whose blocks are emitted
    long long frame_end;
    int run_len;            // frames per work item
    int n_runs;             // work items per channel
    int channels;
    int channel_offset;     // channel index of row 0 (phase counter)
    int step;               // sample_step_len (stretcher.rs:55)
    int pitch;              // pitch_multiple
    int samples_needed;     // S (stretcher.rs:47-51)
    float amp;              // corrected_amp_factor (stretcher.rs:53)
    float inv_w;            // 1 / W (exact: W is a power of two)
    uint32_t flags;
    PhiloxKeys keys;
};

template <int LOG2H>
struct FrameCfg {
    static constexpr int NP = make_plan(LOG2H).npasses;
    static constexpr int H = 1 << LOG2H;
    static constexpr int W = 2 * H;
    static constexpr int NT = threads_per_frame(LOG2H);
    static constexpr int FPC = frames_per_cta(LOG2H);
    static constexpr int SMEM_CF = padded_len(H);  // per frame lane
    static constexpr int RLAST = 1 << make_plan(LOG2H).rlog2[NP - 1];
};

RCG_HD float ld_stream(const float* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
RCG_HD cf ld_cf(const cf* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- LOAD: samples * window -> even/odd packed complex, natural order (fft.rs:51-58) ----
template <int LOG2H>
RCG_HD void phase_load(cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    const float* in = prm.in + (long long)channel * prm.in_stride;
    const long long base = frame * (long long)prm.step - prm.in_origin;
    for (int m = t; m < C::H; m += C::NT) {
        long long i0 = base + 2 * m;
        float x0 = (i0 >= 0 && i0 < prm.in_valid) ? ld_stream(in + i0) : 0.0f;
        float x1 = (i0 + 1 >= 0 && i0 + 1 < prm.in_valid) ? ld_stream(in + i0 + 1) : 0.0f;
        float w0 = ld_stream(prm.window + 2 * m), w1 = ld_stream(prm.window + 2 * m + 1);
        S[phys(m)] = make_cf(RCG_MUL(x0, w0), RCG_MUL(x1, w1));
    }
}

// ---- one in-place FFT pass ------------------------------------------------------------
template <int LOG2H, int PASS, bool INV>
RCG_HD void phase_fft_pass(cf* S, int t, const cf* tw_all) {
    using C = FrameCfg<LOG2H>;
    constexpr int RL = make_plan(LOG2H).rlog2[PASS], R = 1 << RL, ML = make_plan(LOG2H).mlog2[PASS], M = 1 << ML;
    constexpr int NB = C::H / R;
    constexpr int TWOFF = tw_offset(LOG2H, PASS);
    const cf* tw = tw_all + TWOFF;
    for (int b = t; b < NB; b += C::NT) {
        const int m = b & (M - 1), blk = b >> ML;
        const int base = (blk << (ML + RL)) + m;
        float xr[R], xi[R];
#pragma unroll
        for (int a = 0; a < R; a++) {
            cf v = S[phys(base + (a << ML))];
            xr[a] = v.x; xi[a] = v.y;
        }
        if constexpr (INV && M > 1) {
#pragma unroll
            for (int a = 1; a < R; a++) {  // conj twiddle on the inputs (DIT)
                cf w = ld_cf(tw + (a - 1) * M + m);
                float nr = xr[a] * w.x + xi[a] * w.y, ni = xi[a] * w.x - xr[a] * w.y;
                xr[a] = nr; xi[a] = ni;
            }
        }
        if constexpr (INV) Dft<R>::run(xi, xr); else Dft<R>::run(xr, xi);
        if constexpr (!INV && M > 1) {
#pragma unroll
            for (int j = 1; j < R; j++) {  // twiddle on the outputs (DIF)
                cf w = ld_cf(tw + (j - 1) * M + m);
                float nr = xr[j] * w.x - xi[j] * w.y, ni = xi[j] * w.x + xr[j] * w.y;
                xr[j] = nr; xi[j] = ni;
            }
        }
#pragma unroll
        for (int j = 0; j < R; j++) S[phys(base + (j << ML))] = make_cf(xr[j], xi[j]);
    }
}

RCG_HD void sincos_phase(float phi, float& s, float& c) {
#if defined(__CUDA_ARCH__)
    sincosf(phi, &s, &c);
#else
    s = sinf(phi); c = cosf(phi);
#endif
}

// ---- SPECTRAL: real-FFT split -> |X[k]| -> random phases for all W bins -> Hermitian
// fold -> pre-pass of the real inverse (fft.rs:63-69).  One thread owns the bin pair
// (k, H-k) of the half-size transform, i.e. bins {k, W-k, H-k, H+k} of the W-point one.
template <int LOG2H>
RCG_HD void phase_spectral(cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    constexpr int HALF = C::RLAST / 2;
    const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
    for (int c = t; c < C::H / 2; c += C::NT) {
        const int rho = c / HALF, kl = c - rho * HALF;
        const int pos = rho * C::RLAST + kl;
        const int k = pos_to_bin<LOG2H>(pos);
        uint32_t u[4];
        if (c == 0) {
            // k = 0: bins 0 and H are real: X[0] = a + b, X[H] = a - b
            cf z0 = S[0];
            philox4x32_10(0u, f_lo, f_hi, (uint32_t)(channel + prm.channel_offset), prm.keys, u);
            float m0 = fabsf(z0.x + z0.y), mh = fabsf(z0.x - z0.y);
            float s, c0, ch;
            sincos_phase(phase_from_u32(u[0]), s, c0);
            sincos_phase(phase_from_u32(u[2]), s, ch);
            float y0 = m0 * c0, yh = mh * ch;  // Re(B_0), Re(B_H)
            S[0] = make_cf(y0 + yh, y0 - yh);
            // k = H/2 pairs with itself: X[H/2] = conj(Z[H/2]); Z'[H/2] = 2 conj(Y[H/2])
            const int pq = phys(bin_to_pos<LOG2H>(C::H / 2));
            cf zq = S[pq];
            philox4x32_10((uint32_t)(C::H / 2), f_lo, f_hi, (uint32_t)(channel + prm.channel_offset), prm.keys, u);
            float mq = sqrtf(zq.x * zq.x + zq.y * zq.y);
            float sa, ca, sb, cb;
            sincos_phase(phase_from_u32(u[0]), sa, ca);
            sincos_phase(phase_from_u32(u[1]), sb, cb);
            S[pq] = make_cf(mq * (ca + cb), -(mq * (sa - sb)));
            continue;
        }
        const int pa = phys(pos), pb = phys(bin_to_pos<LOG2H>(C::H - k));
        const cf za = S[pa], zb = S[pb];
        const cf tw = ld_cf(prm.spec_tw + c);  // exp(-2 pi i k / W)
        philox4x32_10((uint32_t)k, f_lo, f_hi, (uint32_t)(channel + prm.channel_offset), prm.keys, u);
        // P = za + conj(zb), Q = za - conj(zb), U = -i t Q
        float pr = za.x + zb.x, pi = za.y - zb.y, qr = za.x - zb.x, qi = za.y + zb.y;
        float tqr = tw.x * qr - tw.y * qi, tqi = tw.x * qi + tw.y * qr;
        float ur = tqi, ui = -tqr;
        // X[k] = (P + U)/2, conj(X[H-k]) = (P - U)/2
        float xkr = 0.5f * (pr + ur), xki = 0.5f * (pi + ui);
        float xbr = 0.5f * (pr - ur), xbi = 0.5f * (pi - ui);
        float mk = sqrtf(xkr * xkr + xki * xki), mb = sqrtf(xbr * xbr + xbi * xbi);
        float s0, c0, s1, c1, s2, c2, s3, c3;
        sincos_phase(phase_from_u32(u[0]), s0, c0);  // bin k
        sincos_phase(phase_from_u32(u[1]), s1, c1);  // bin W - k
        sincos_phase(phase_from_u32(u[2]), s2, c2);  // bin H - k
        sincos_phase(phase_from_u32(u[3]), s3, c3);  // bin H + k
        // Hermitian part of the W-bin random-phase spectrum: Y[k] = |X[k]| (e^{i phi_k} + e^{-i phi_{W-k}}) / 2
        float hk = 0.5f * mk, hb = 0.5f * mb;
        float ykr = hk * (c0 + c1), yki = hk * (s0 - s1);
        float ybr = hb * (c2 + c3), ybi = hb * (s2 - s3);
        // A = Y[k] + conj(Y[H-k]), B = Y[k] - conj(Y[H-k]), Cc = i conj(t) B
        float ar = ykr + ybr, ai = yki - ybi, br = ykr - ybr, bi = yki + ybi;
        float cr = -(tw.x * bi - tw.y * br), ci = tw.x * br + tw.y * bi;
        S[pa] = make_cf(ar + cr, ai + ci);      // Z'[k]   = A + Cc
        S[pb] = make_cf(ar - cr, -(ai - ci));   // Z'[H-k] = conj(A - Cc)
    }
}

// value i (0..W) of the frame held as packed pairs
RCG_HD float frame_value(const cf* S, int i) {
    cf v = S[phys(i >> 1)];
    return (i & 1) ? v.y : v.x;
}

// ---- EMIT part 1: scale + synthesis window (fft.rs:70-73), overlap-add with the carried
// half window, envelope, amplitude (stretcher.rs:96-102).  Arithmetic order is the
// reference's: ((re / W) * w), then ((y + tail) * env) * amp.
template <int LOG2H>
RCG_HD void phase_ola(cf* S, int t, const StretchParams& prm, int channel, long long frame,
                      float* tail, bool emit, bool write_head) {
    using C = FrameCfg<LOG2H>;
    const bool bypass = prm.flags & kFlagBypass;
    float* out = prm.out + (long long)channel * prm.out_stride;
    for (int m = t; m < C::H / 2; m += C::NT) {
        const cf a = S[phys(m)], b = S[phys(m + C::H / 2)];
        const float wa0 = ld_stream(prm.window + 2 * m), wa1 = ld_stream(prm.window + 2 * m + 1);
        const float wb0 = ld_stream(prm.window + C::H + 2 * m), wb1 = ld_stream(prm.window + C::H + 2 * m + 1);
        float y0, y1, t0, t1;
        if (bypass) {
            y0 = RCG_MUL(a.x, wa0); y1 = RCG_MUL(a.y, wa1);
            t0 = RCG_MUL(b.x, wb0); t1 = RCG_MUL(b.y, wb1);
        } else {
            y0 = RCG_MUL(RCG_MUL(a.x, prm.inv_w), wa0); y1 = RCG_MUL(RCG_MUL(a.y, prm.inv_w), wa1);
            t0 = RCG_MUL(RCG_MUL(b.x, prm.inv_w), wb0); t1 = RCG_MUL(RCG_MUL(b.y, prm.inv_w), wb1);
        }
        const float o0 = tail[2 * m], o1 = tail[2 * m + 1];
        tail[2 * m] = t0; tail[2 * m + 1] = t1;
        if (write_head) {
            float* head = prm.head_out + (long long)channel * C::H;
            head[2 * m] = y0; head[2 * m + 1] = y1;
        }
        if (emit) {
            const float e0 = ld_stream(prm.env + 2 * m), e1 = ld_stream(prm.env + 2 * m + 1);
            const float z0 = RCG_MUL(RCG_MUL(RCG_ADD(y0, o0), e0), prm.amp);
            const float z1 = RCG_MUL(RCG_MUL(RCG_ADD(y1, o1), e1), prm.amp);
            if (prm.pitch == 1) {
                const long long o = frame * (long long)C::H + 2 * m - prm.out_origin;
                out[o] = z0; out[o + 1] = z1;
            } else {
                S[phys(m)] = make_cf(z0, z1);  // staged for the resampling gather
            }
        }
    }
}

// ---- EMIT part 2: integer resample of the finished half window held in S[0..H/2)
// (resampler.rs:3-35).  p > 1: out[g / p] = z[g] for g % p == 0 (g = absolute index into the
// overlap-added stream; calls tile it exactly since S = W*p).  p < -1: one frame per call;
// the first S samples are linearly interpolated |p| times, S-1 intervals (math.rs:28-30).
template <int LOG2H>
RCG_HD void phase_resample(const cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    float* out = prm.out + (long long)channel * prm.out_stride;
    const int p = prm.pitch;
    if (p > 1) {
        const long long g0 = frame * (long long)C::H;
        const long long o_begin = (g0 + p - 1) / p, o_end = (g0 + C::H + p - 1) / p;
        for (long long o = o_begin + t; o < o_end; o += C::NT) {
            const int i = (int)(o * p - g0);
            out[o - prm.out_origin] = frame_value(S, i);
        }
    } else if (p < -1) {
        const int q = -p;
        const int n = (prm.samples_needed - 1) * q;
        float* dst = out + (frame * (long long)n - prm.out_origin);
        for (int tt = t; tt < n; tt += C::NT) {
            const int i = tt / q, j = tt - i * q;
            const float cur = frame_value(S, i), nxt = frame_value(S, i + 1);
            const float ratio = RCG_DIV((float)j, (float)q);
            dst[tt] = RCG_ADD(cur, RCG_MUL(RCG_SUB(nxt, cur), ratio));
        }
    }
}

template <int LOG2H, bool INV, int PASS, class Exec>
RCG_HD void fft_all(Exec& ex, cf* S, const StretchParams& prm, bool work) {
    using C = FrameCfg<LOG2H>;
    if constexpr (PASS >= 0 && PASS < C::NP) {
        ex.par([&](int t) { if (work) phase_fft_pass<LOG2H, PASS, INV>(S, t, prm.tw); });
        fft_all<LOG2H, INV, INV ? PASS - 1 : PASS + 1>(ex, S, prm, work);
    }
}

// ---- the frame loop of one work item (a run of consecutive frames of one channel) --------
// Exec::par(f) runs f(t) for every thread of the frame lane followed by a barrier.
template <int LOG2H, class Exec>
RCG_HD void run_work_item(Exec& ex, cf* S, const StretchParams& prm, int item, float* tail, bool active) {
    using C = FrameCfg<LOG2H>;
    const int channel = active ? item / prm.n_runs : 0;
    const int run = active ? item - channel * prm.n_runs : 0;
    const long long f0 = prm.frame_begin + (long long)run * prm.run_len;
    long long f1 = f0 + prm.run_len;
    if (f1 > prm.frame_end) f1 = prm.frame_end;
    // How the half window carried into block f0 is obtained:
    //  - f0 == 0: zeros (stretcher.rs:58-59)
    //  - first run with tail_in: copy it;  with kFlagDeferFirst: block f0 is finished later
    //  - otherwise recompute frame f0-1 (halo frame; its block is not emitted)
    const bool first_run = (run == 0);
    const bool defer = first_run && (prm.flags & kFlagDeferFirst) && f0 > 0;
    const bool have_tail_in = first_run && prm.tail_in != nullptr && f0 > 0;
    const bool halo = active && f0 > 0 && !defer && !have_tail_in;
    ex.par([&](int t) {
        if (!active) return;
        const float* src = have_tail_in ? prm.tail_in + (long long)channel * C::H : nullptr;
        for (int i = t; i < C::H; i += C::NT) tail[i] = src ? src[i] : 0.0f;
    });
    const int n_iter = prm.run_len + 1;  // uniform trip count for every lane of the CTA
    for (int it = 0; it < n_iter; it++) {
        const long long f = f0 - 1 + it;
        const bool is_halo = (it == 0);
        const bool work = active && (is_halo ? halo : f < f1);
        const bool bypass = prm.flags & kFlagBypass;
        ex.par([&](int t) { if (work) phase_load<LOG2H>(S, t, prm, channel, f); });
        if (!bypass) {
            fft_all<LOG2H, false, 0>(ex, S, prm, work);
            ex.par([&](int t) { if (work) phase_spectral<LOG2H>(S, t, prm, channel, f); });
            fft_all<LOG2H, true, C::NP - 1>(ex, S, prm, work);
        }
        const bool head = work && defer && f == f0;
        const bool emit = work && !is_halo && !head;
        ex.par([&](int t) { if (work) phase_ola<LOG2H>(S, t, prm, channel, f, tail, emit, head); });
        if (prm.pitch != 1)
            ex.par([&](int t) { if (emit) phase_resample<LOG2H>(S, t, prm, channel, f); });
    }
    if (prm.tail_out != nullptr) {
        ex.par([&](int t) {
            if (!active || f1 != prm.frame_end) return;
            float* dst = prm.tail_out + (long long)channel * C::H;
            for (int i = t; i < C::H; i += C::NT) dst[i] = tail[i];
        });
    }
}

// ---- boundary block: finish block frame_begin of a segment from the deferred raw first
// half and the previous segment's carried half window (multi-GPU time sharding).
template <int LOG2H, class Exec>
RCG_HD void run_boundary(Exec& ex, cf* S, const StretchParams& prm, int channel, const float* head,
                         const float* tail_prev) {
    using C = FrameCfg<LOG2H>;
    float* out = prm.out + (long long)channel * prm.out_stride;
    const long long frame = prm.frame_begin;
    ex.par([&](int t) {
        for (int m = t; m < C::H / 2; m += C::NT) {
            const float* h = head + (long long)channel * C::H;
            const float* tp = tail_prev + (long long)channel * C::H;
            const float e0 = prm.env[2 * m], e1 = prm.env[2 * m + 1];
            const float z0 = RCG_MUL(RCG_MUL(RCG_ADD(h[2 * m], tp[2 * m]), e0), prm.amp);
            const float z1 = RCG_MUL(RCG_MUL(RCG_ADD(h[2 * m + 1], tp[2 * m + 1]), e1), prm.amp);
            if (prm.pitch == 1) {
                const long long o = frame * (long long)C::H + 2 * m - prm.out_origin;
                out[o] = z0; out[o + 1] = z1;
            } else {
                S[phys(m)] = make_cf(z0, z1);
            }
        }
    });
    if (prm.pitch != 1) ex.par([&](int t) { phase_resample<LOG2H>(S, t, prm, channel, frame); });
}

}  // namespace rcg

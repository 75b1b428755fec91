// rcg_frame_quad.h -- W = 65536, fused path: one frame across a 4-CTA thread-block cluster.
//
// The packed H = 32768 point complex transform is split by a radix-4 decimation-in-frequency stage taken
// across the cluster: CTA r (r = 0..3) forms y_r[n] = (sum_q z[n + q H/4] (-i)^{qr}) w_H^{rn}, n < H/4,
// straight from global memory and runs the ordinary 8192-point in-place transform of rcg_frame.h on it,
// which leaves the bins k = 4j + r.  Each CTA is then exactly a W = 16384 frame lane (256 threads, 69.6 KB),
// so three of them fit on an SM: 111 frames in flight with 24 warps per SM, where the CTA pair of
// rcg_frame_pair.h (one 139 KB CTA per SM) holds 74 with 16.
//
// Bin k pairs with H - k = 4 (H/4 - 1 - j) + (4 - r): classes 0 and 2 pair within their CTA, classes 1 and 3
// with each other -- CTA 1 and CTA 3 each own the pairs whose smaller bin is theirs and reach the partner
// bin through distributed shared memory.  After the local inverses each CTA multiplies its quarter by
// conj(w_H^{rn}); the radix-4 decimation-in-time stage then reads all four quarters through DSMEM: CTAs 0, 1
// produce the first half window (emitted), CTAs 2, 3 the second half (carried through the cluster's
// double-buffered global scratch).  Same reference arithmetic as rcg_frame.h (fft.rs:42-74,
// stretcher.rs:96-112, resampler.rs:3-35).  The frequency-kernel path at this size stays on the CTA pair.
//
// STATUS: opt-in (environment RCG_QUAD=1 at rcg_create), parity-tested on the host emulator and on the GPU.
// Measured on B200, config 3 (W = 65536, f = 4): 6.07 ms against 5.15 ms for the CTA pair -- 104 resident
// clusters instead of 74, but each CTA loads the whole frame and window (4x the L2 reads), the last stage reads
// three remote quarters per output, and a frame crosses four cluster barriers (28 % of the stall samples;
// long-scoreboard stalls on global and DSMEM loads another 35 %; profiles/r1q_quad_kernel.md).
#pragma once
#include "rcg_frame_pair.h"

namespace rcg {

constexpr int kQuadLog2HL = 13;               // local transform: 8192 points
constexpr int kQuadHL = 1 << kQuadLog2HL;     // H/4
constexpr int kQuadTwLen = kQuadHL / 64 + 64; // two-level table per r
static_assert(4 * kQuadHL == kPairH, "same window as the CTA pair");

struct QuadTables {
    const cf* cross_tw;    // [3][kQuadTwLen], r = 1..3: exp(-2 pi i r 64 a / H), a < HL/64, then exp(-2 pi i r b / H), b < 64
    const cf* spec_tw;     // [4][HL/2]: exp(-2 pi i (4 j + r) / W) in the local pair order of the 8192-point plan
    const cf* tw;          // pass twiddles of the 8192-point plan (StretchParams::tw of a W = 16384 engine)
    const uint2* pair_tab; // its pair table (StretchParams::pair_tab)
    float* scratch;        // per cluster: tail[2][H] then stage[H]  (floats; H = kPairH)
};

RCG_HD cf quad_cross_tw(const cf* tab, int r, int n) {   // w_H^{r n}, 1 <= r <= 3
    const cf* t = tab + (r - 1) * kQuadTwLen;
    const cf a = ld_cf(t + (n >> 6)), b = ld_cf(t + kQuadHL / 64 + (n & 63));
    cf w;
    cmul(a.x, a.y, b.x, b.y, w.x, w.y);
    return w;
}

// radix-4 butterfly output `o` of inputs v0..v3 with sign s = -1 (forward, (-i)^{qo}) or +1 (inverse, (+i)^{qo})
template <int SIGN>
RCG_HD cf radix4_out(int o, cf v0, cf v1, cf v2, cf v3) {
    const float ax = v0.x + v2.x, ay = v0.y + v2.y, bx = v0.x - v2.x, by = v0.y - v2.y;
    const float cx = v1.x + v3.x, cy = v1.y + v3.y, dx = v1.x - v3.x, dy = v1.y - v3.y;
    if (o == 0) return make_cf(ax + cx, ay + cy);
    if (o == 2) return make_cf(ax - cx, ay - cy);
    // o = 1: b + SIGN i d;  o = 3: b - SIGN i d;   i d = (-dy, dx)
    const bool plus = (o == 1) == (SIGN > 0);
    return plus ? make_cf(bx - dy, by + dx) : make_cf(bx + dy, by - dx);
}

// ---- LOAD + cross-CTA radix-4 DIF stage -------------------------------------------------------------
template <bool FAST>
RCG_HD void quad_phase_load_impl(cf* S, int r, int t, const StretchParams& prm, const QuadTables& qt, const float* in,
                                 long long base, bool bypass) {
    using C = FrameCfg<kQuadLog2HL>;
    auto sample = [&](long long i) { return (i >= 0 && i < prm.in_valid) ? ld_stream(in + i) : 0.0f; };
RCG_UNROLL(2)
    for (int n = t; n < kQuadHL; n += C::NT) {
        cf z[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int m = 2 * (n + q * kQuadHL);
            const cf w = ld_stream2(prm.window + m);
            cf x;
            if constexpr (FAST) x = ld_stream2(in + base + m);
            else x = make_cf(sample(base + m), sample(base + m + 1));
            z[q] = make_cf(RCG_MUL(x.x, w.x), RCG_MUL(x.y, w.y));
        }
        cf v;
        if (bypass) {
            v = z[r];
        } else {
            v = radix4_out<-1>(r, z[0], z[1], z[2], z[3]);
            if (r != 0) {
                const cf w = quad_cross_tw(qt.cross_tw, r, n);
                v = make_cf(v.x * w.x - v.y * w.y, v.y * w.x + v.x * w.y);
            }
        }
        S[phys(n)] = v;
    }
}
RCG_HD void quad_phase_load(cf* S, int r, int t, const StretchParams& prm, const QuadTables& qt, int channel,
                            long long frame) {
    const float* in = prm.in + (long long)channel * prm.in_stride;
    const long long base = frame * (long long)prm.step - prm.in_origin;
    const bool bypass = prm.flags & kFlagBypass;
    const bool fast = base >= 0 && base + kPairW <= prm.in_valid &&
                      (reinterpret_cast<unsigned long long>(in + base) & 7ull) == 0;
    if (fast) quad_phase_load_impl<true>(S, r, t, prm, qt, in, base, bypass);
    else quad_phase_load_impl<false>(S, r, t, prm, qt, in, base, bypass);
}

// ---- SPECTRAL, bin classes 1, 2, 3: k = 4j + r (j < HL/2, the smaller bin of its pair) lives in `own`,
// its partner H - k = 4 (HL-1-j) + (4-r) in `other` (= own for r = 2, the peer CTA 4-r otherwise) ----
RCG_HD void quad_phase_spectral_rest(cf* own, cf* other, int r, int t, const StretchParams& prm, const QuadTables& qt,
                                     int channel, long long frame) {
    using C = FrameCfg<kQuadLog2HL>;
    const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
    const uint32_t chan = (uint32_t)(channel + prm.channel_offset);
    const cf* twt = qt.spec_tw + r * (kQuadHL / 2);
    for (int c = t; c < kQuadHL / 2; c += C::NT) {
        const int pos = spectral_slot_pos(kQuadLog2HL, c);
        const int j = pos_to_bin<kQuadLog2HL>(pos);
        const int k = 4 * j + r;
        const int pa = phys(pos), pb = phys(bin_to_pos<kQuadLog2HL>(kQuadHL - 1 - j));
        const cf za = own[pa], zb = other[pb];
        const cf tw = ld_cf(twt + c);
        uint32_t u[4];
        philox4x32_10((uint32_t)k, f_lo, f_hi, chan, prm.keys, u);
        float ar, ai, br, bi;
        pair_fold(za, zb, tw, u, ar, ai, br, bi, prm, (uint32_t)k, (uint32_t)kPairW);
        const float cr = -(tw.x * bi - tw.y * br), ci = tw.x * br + tw.y * bi;
        own[pa] = make_cf(ar + cr, ai + ci);
        other[pb] = make_cf(ar - cr, -(ai - ci));
    }
}

// ---- after the local inverse: quarter r times conj(w_H^{rn}), in place ----------------------------------
RCG_HD void quad_phase_untwist(cf* S, int r, int t, const QuadTables& qt) {
    using C = FrameCfg<kQuadLog2HL>;
    if (r == 0) return;
RCG_UNROLL(4)
    for (int n = t; n < kQuadHL; n += C::NT) {
        const cf v = S[phys(n)];
        const cf w = quad_cross_tw(qt.cross_tw, r, n);
        S[phys(n)] = make_cf(v.x * w.x + v.y * w.y, v.y * w.x - v.x * w.y);
    }
}

// ---- cross-CTA radix-4 DIT stage + overlap-add: CTA q produces quarter q of the window -------------------
RCG_HD void quad_phase_ola(cf* const (&bufs)[4], int q, int t, const StretchParams& prm, int channel, long long frame,
                           const float* tail_old, float* tail_new, float* stage, bool emit, bool write_head) {
    using C = FrameCfg<kQuadLog2HL>;
    const bool bypass = prm.flags & kFlagBypass;
    float* out = prm.out + (long long)channel * prm.out_stride;
    const float scale = bypass ? 1.0f : prm.inv_w;
    const bool first_half = q < 2;
    const int half_off = (q & 1) * kQuadHL;   // complex index of this quarter inside its half window
RCG_UNROLL(2)
    for (int n = t; n < kQuadHL; n += C::NT) {
        cf z;
        if (bypass) {
            z = bufs[q][phys(n)];
        } else {
            const cf y0 = bufs[0][phys(n)], y1 = bufs[1][phys(n)], y2 = bufs[2][phys(n)], y3 = bufs[3][phys(n)];
            z = radix4_out<+1>(q, y0, y1, y2, y3);
        }
        const int m = 2 * (half_off + n);     // float index inside the half window
        if (first_half) {
            const cf wa = ld_stream2(prm.window + m);
            const float y0 = RCG_MUL(RCG_MUL(z.x, scale), wa.x), y1 = RCG_MUL(RCG_MUL(z.y, scale), wa.y);
            const cf old = ld_pair(tail_old + m);
            const cf e = ld_stream2(prm.env + m);
            if (write_head) st_pair(prm.head_out + (long long)channel * kPairH + m, y0, y1);
            if (emit) {
                const float z0 = RCG_MUL(RCG_MUL(RCG_ADD(y0, old.x), e.x), prm.amp);
                const float z1 = RCG_MUL(RCG_MUL(RCG_ADD(y1, old.y), e.y), prm.amp);
                if (prm.pitch == 1) st_pair(out + (frame * (long long)kPairH + m - prm.out_origin), z0, z1);
                else st_pair(stage + m, z0, z1);
            }
        } else {
            const cf wb = ld_stream2(prm.window + kPairH + m);
            st_pair(tail_new + m, RCG_MUL(RCG_MUL(z.x, scale), wb.x), RCG_MUL(RCG_MUL(z.y, scale), wb.y));
        }
    }
}

// resample from the staged half window in global memory, all four CTAs (tg = thread index in the cluster)
RCG_HD void quad_phase_resample(const float* stage, int tg, int nthreads, const StretchParams& prm, int channel,
                                long long frame) {
    float* out = prm.out + (long long)channel * prm.out_stride;
    const int p = prm.pitch;
    if (p > 1) {
        const long long g0 = frame * (long long)kPairH;
        const long long o_begin = (g0 + p - 1) / p, o_end = (g0 + kPairH + p - 1) / p;
        for (long long o = o_begin + tg; o < o_end; o += nthreads) out[o - prm.out_origin] = stage[(int)(o * p - g0)];
    } else if (p < -1) {
        const int q = -p;
        const int n = (prm.samples_needed - 1) * q;
        float* dst = out + (frame * (long long)n - prm.out_origin);
        const int di = nthreads / q, dj = nthreads - di * q;
        int i = tg / q, j = tg - i * q;
        for (int tt = tg; tt < n; tt += nthreads) {
            const float cur = stage[i], nxt = stage[i + 1];
            dst[tt] = RCG_ADD(cur, RCG_MUL(RCG_SUB(nxt, cur), prm.lerp_ratio[j]));
            i += di; j += dj;
            if (j >= q) { j -= q; i++; }
        }
    }
}

// Exec for the cluster: par(f) runs f(r, t) for every thread of the four CTAs, then a CTA barrier;
// par_cluster(f) ends in a cluster barrier (release/acquire) instead; cluster_sync() is a bare one.
template <int PASS, int LAST, bool INV, class Exec>
RCG_HD void quad_fft_range(Exec& ex, cf* const (&bufs)[4], const StretchParams& prm, bool work) {
    if constexpr (INV ? PASS >= LAST : PASS <= LAST) {
        ex.par([&](int r, int t) { if (work) phase_fft_pass<kQuadLog2HL, PASS, INV>(bufs[r], t, prm.tw); });
        quad_fft_range<INV ? PASS - 1 : PASS + 1, LAST, INV>(ex, bufs, prm, work);
    }
}
template <bool INV, class Exec>
RCG_HD void quad_fft_all(Exec& ex, cf* const (&bufs)[4], const StretchParams& prm, bool work) {
    constexpr int NP = FrameCfg<kQuadLog2HL>::NP;
    if constexpr (!INV) quad_fft_range<0, NP - 3, false>(ex, bufs, prm, work);
    ex.par_steps(fft_tail_steps<kQuadLog2HL>(),
                 [&](int step, int r, int t) { if (work) fft_tail_step<kQuadLog2HL, INV>(bufs[r], t, step, prm.tw); });
    if constexpr (INV) quad_fft_range<NP - 3, 0, true>(ex, bufs, prm, work);
}

template <class Exec>
RCG_HD void run_work_item_quad(Exec& ex, cf* const (&bufs)[4], const StretchParams& prm_in, const QuadTables& qt,
                               int item, float* scratch, bool active) {
    using C = FrameCfg<kQuadLog2HL>;
    StretchParams prm = prm_in;   // the local transforms use the tables of the 8192-point plan
    prm.tw = qt.tw; prm.pair_tab = qt.pair_tab; prm.spec_tw = qt.spec_tw;
    const int channel = active ? item / prm.n_runs : 0;
    const int run = active ? item - channel * prm.n_runs : 0;
    const long long f0 = prm.frame_begin + (long long)run * prm.run_len;
    long long f1 = f0 + prm.run_len;
    if (f1 > prm.frame_end) f1 = prm.frame_end;
    const bool first_run = (run == 0);
    const bool defer = first_run && (prm.flags & kFlagDeferFirst) && f0 > 0;
    const bool have_tail_in = first_run && prm.tail_in != nullptr && f0 > 0;
    const bool halo = active && f0 > 0 && !defer && !have_tail_in;
    const bool bypass = prm.flags & kFlagBypass;
    float* tails[2] = {scratch, scratch + kPairH};
    float* stage = scratch + 2 * kPairH;
    // carried half window for the first iteration lives in tails[0]
    ex.par_cluster([&](int r, int t) {
        if (!active) return;
        const float* src = have_tail_in ? prm.tail_in + (long long)channel * kPairH : nullptr;
        for (int i = r * C::NT + t; i < kPairH; i += 4 * C::NT) tails[0][i] = src ? src[i] : 0.0f;
    });
    const int n_iter = prm.run_len + 1;
    int cnt = 0;  // frames done so far: the carried half window is in tails[cnt & 1]
    for (int it = 0; it < n_iter; it++) {
        const long long f = f0 - 1 + it;
        const bool is_halo = (it == 0);
        const bool work = active && (is_halo ? halo : f < f1);
        const bool head = work && defer && f == f0;
        const bool emit = work && !is_halo && !head;
        ex.par([&](int r, int t) { if (work) quad_phase_load(bufs[r], r, t, prm, qt, channel, f); });
        if (!bypass) {
            quad_fft_all<false>(ex, bufs, prm, work);
            ex.cluster_sync();   // CTAs 1 and 3 are about to read each other's bins
            ex.par_cluster([&](int r, int t) {   // ends in a cluster barrier: the partners' writes have landed
                if (!work) return;
                if (r == 0) {
                    phase_spectral<kQuadLog2HL, 4>(bufs[0], t, prm, channel, f);
                } else {
                    quad_phase_spectral_rest(bufs[r], bufs[(4 - r) & 3], r, t, prm, qt, channel, f);
                }
            });
            quad_fft_all<true>(ex, bufs, prm, work);
            ex.par([&](int r, int t) { if (work) quad_phase_untwist(bufs[r], r, t, qt); });
        }
        ex.cluster_sync();   // all four quarters of the inverse are complete
        ex.par_cluster([&](int r, int t) {   // ends in a cluster barrier: peers are done reading, tails are visible
            if (work) quad_phase_ola(bufs, r, t, prm, channel, f, tails[cnt & 1], tails[(cnt & 1) ^ 1], stage, emit, head);
        });
        if (work) cnt++;
        if (prm.pitch != 1)
            ex.par([&](int r, int t) { if (emit) quad_phase_resample(stage, r * C::NT + t, 4 * C::NT, prm, channel, f); });
    }
    if (prm.tail_out != nullptr) {
        ex.par([&](int r, int t) {
            if (!active || f1 != prm.frame_end) return;
            const float* last = tails[cnt & 1];
            float* dst = prm.tail_out + (long long)channel * kPairH;
            for (int i = r * C::NT + t; i < kPairH; i += 4 * C::NT) dst[i] = last[i];
        });
    }
}

}  // namespace rcg

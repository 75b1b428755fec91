// rcg_frame_pair.h -- W = 65536: one frame across a 2-CTA thread-block cluster.
//
// The packed H = 32768 point complex transform (256 KB) does not fit one SM's shared memory, so
// the first DIF radix-2 stage is taken across the CTA pair: CTA 0 keeps u[n] = z[n] + z[n+H/2]
// (-> even bins), CTA 1 keeps v[n] = (z[n] - z[n+H/2]) w_H^n (-> odd bins); each then runs the
// ordinary 16384-point in-place transform of rcg_frame.h on its half.  The spectral step stays
// local because bin k pairs with H-k of the same parity.  After the local inverses the last DIT
// radix-2 stage reads the peer's half through distributed shared memory: CTA 0 produces the
// first half window (emitted), CTA 1 the second half (carried to the next frame through a
// double-buffered global scratch).  Same reference arithmetic as rcg_frame.h (fft.rs:42-74,
// stretcher.rs:96-112, resampler.rs:3-35).
#pragma once
#include "rcg_frame.h"

// unroll factors of the two phases that wait on global / distributed-shared loads (512 threads per CTA, one
// CTA per SM: 128 registers per thread are available)
// (measured after the kernel's shared-memory accesses became LDS / STS, profiles/r3p_*, r3r_*: load 2 / overlap-add 8
// is 7 % faster than the 4 / 2 of before; load 1, 3, 4, 8 and overlap-add 1, 2, 4 are all slower)
#ifndef RCG_PAIR_LOAD_UNROLL
#define RCG_PAIR_LOAD_UNROLL 2
#endif
#ifndef RCG_PAIR_OLA_UNROLL
#define RCG_PAIR_OLA_UNROLL 8
#endif
// 1: the stretch path reads a frame once per pair (pair_phase_load_split); 0: every CTA reads the whole frame
#ifndef RCG_PAIR_SPLIT_LOAD
#define RCG_PAIR_SPLIT_LOAD 1
#endif

namespace rcg {

constexpr int kPairLog2HL = 14;              // local transform: 16384 points
constexpr int kPairHL = 1 << kPairLog2HL;    // H/2
constexpr int kPairH = 2 * kPairHL;          // 32768
constexpr int kPairW = 2 * kPairH;           // 65536
constexpr int kPairTwLen = kPairHL / 64 + 64;  // two-level table for w_H^n, n < H/2

struct PairTables {
    const cf* cross_tw;      // [kPairTwLen]: A[n >> 6] (exp(-2 pi i 64 a / H)) then B[n & 63]
    const cf* spec_tw_even;  // [HL/2]: exp(-2 pi i (2j) / W) in the local pair order
    const cf* spec_tw_odd;   // [HL/2]: exp(-2 pi i (2j+1) / W)
    float* scratch;          // per cluster: tail[2][H] then stage[H]
};

RCG_HD cf pair_cross_tw(const cf* tab, int n) {   // tab: global memory, or the CTA's shared-memory copy
    const cf a = tab[n >> 6], b = tab[kPairHL / 64 + (n & 63)];
    return tw_mul<false>(a, make_tw(b));
}

// ---- LOAD + cross-CTA radix-2 DIF stage ---------------------------------------------------------
// FAST: the whole frame lies inside the input and is 8-byte aligned -> branch-free float2 loads that
// the compiler batches across unrolled iterations
template <bool FAST>
RCG_HD void pair_phase_load_impl(cf* S, int r, int t, const StretchParams& prm, const PairTables& pt, const float* in,
                                 long long base, bool bypass) {
    using C = FrameCfg<kPairLog2HL>;
    auto sample = [&](long long i) { return (i >= 0 && i < prm.in_valid) ? ld_stream(in + i) : 0.0f; };
RCG_UNROLL(RCG_PAIR_LOAD_UNROLL)
    for (int n = t; n < kPairHL; n += C::NT) {
        const cf wa = ld_stream2(prm.window + 2 * n), wb = ld_stream2(prm.window + kPairH + 2 * n);
        cf xa, xb;
        if constexpr (FAST) {
            xa = ld_stream2(in + base + 2 * n);
            xb = ld_stream2(in + base + kPairH + 2 * n);
        } else {
            xa = make_cf(sample(base + 2 * n), sample(base + 2 * n + 1));
            xb = make_cf(sample(base + kPairH + 2 * n), sample(base + kPairH + 2 * n + 1));
        }
        const float a0 = RCG_MUL(xa.x, wa.x), a1 = RCG_MUL(xa.y, wa.y);
        const float b0 = RCG_MUL(xb.x, wb.x), b1 = RCG_MUL(xb.y, wb.y);
        cf v;
        if (bypass) {
            v = r == 0 ? make_cf(a0, a1) : make_cf(b0, b1);
        } else if (r == 0) {
            v = make_cf(a0 + b0, a1 + b1);
        } else {
            const cf w = pair_cross_tw(pt.cross_tw, n);
            const float dr = a0 - b0, di = a1 - b1;
            v = make_cf(dr * w.x - di * w.y, di * w.x + dr * w.y);
        }
        S[phys(n)] = v;
    }
}
RCG_HD void pair_phase_load(cf* S, int r, int t, const StretchParams& prm, const PairTables& pt, int channel,
                            long long frame) {
    const float* in = prm.in + (long long)channel * prm.in_stride;
    const long long base = frame * (long long)prm.step - prm.in_origin;
    const bool bypass = prm.flags & kFlagBypass;
    const bool fast = base >= 0 && base + kPairW <= prm.in_valid &&
                      (reinterpret_cast<unsigned long long>(in + base) & 7ull) == 0;
    if (fast) pair_phase_load_impl<true>(S, r, t, prm, pt, in, base, bypass);
    else pair_phase_load_impl<false>(S, r, t, prm, pt, in, base, bypass);
}

// ---- SPECTRAL, odd bins (CTA 1): k = 2j+1 pairs with H-k = 2(HL-1-j)+1; no self-paired bins ----
RCG_HD void pair_phase_spectral_odd(cf* S, int t, const StretchParams& prm, const PairTables& pt, int channel,
                                    long long frame) {
    using C = FrameCfg<kPairLog2HL>;
    const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
    const uint32_t chan = (uint32_t)(channel + prm.channel_offset);
    for (int c = t; c < kPairHL / 2; c += C::NT) {
        const int pos = spectral_slot_pos(kPairLog2HL, c);
        const int j = pos_to_bin<kPairLog2HL>(pos);
        const int k = 2 * j + 1;
        const int pa = phys(pos), pb = phys(bin_to_pos<kPairLog2HL>(kPairHL - 1 - j));
        const cf za = S[pa], zb = S[pb];
        const Tw tw = make_tw<(RCG_PK_SPEC != 0)>(ld_cf(pt.spec_tw_odd + c));
        uint32_t u[4];
        philox4x32_10((uint32_t)k, f_lo, f_hi, chan, prm.keys, u);
        cf A, B;
        pair_fold(za, zb, tw, u, A, B, prm, (uint32_t)k, (uint32_t)kPairW, frame_clock(prm, frame));
        const cf Cc = mul_i_conj(B, tw);
        using SA = Ar<(RCG_PK_SPEC != 0)>;
        S[pa] = SA::add(A, Cc);
        S[pb] = SA::mul(SA::sub(A, Cc), make_cf(1.0f, -1.0f));
    }
}

// ---- frequency-kernel path (fft.rs:76-108) at W = 65536 ------------------------------------------------
// Same two-phase scheme as rcg_frame.h: bins k and H-k have the same parity, so the real-FFT split, the
// copy-out of X[k] / X[W-k] and the magnitude copy-in all stay inside the CTA that owns the parity.
// CTA r owns global bins k = 2b + r, b = local bin.
RCG_HD void pair_phase_spectrum_split_odd(cf* S, int t, const PairTables& pt) {
    using C = FrameCfg<kPairLog2HL>;
    for (int c = t; c < kPairHL / 2; c += C::NT) {
        const int pos = spectral_slot_pos(kPairLog2HL, c);
        const int j = pos_to_bin<kPairLog2HL>(pos);
        const int pa = phys(pos), pb = phys(bin_to_pos<kPairLog2HL>(kPairHL - 1 - j));
        const cf za = S[pa], zb = S[pb];
        const cf tw = ld_cf(pt.spec_tw_odd + c);
        float pr = za.x + zb.x, pi = za.y - zb.y, qr = za.x - zb.x, qi = za.y + zb.y;
        float tqr = tw.x * qr - tw.y * qi, tqi = tw.x * qi + tw.y * qr;
        float ur = tqi, ui = -tqr;
        S[pa] = make_cf(0.5f * (pr + ur), 0.5f * (pi + ui));          // X[k], k = 2j+1
        S[pb] = make_cf(0.5f * (pr - ur), -(0.5f * (pi - ui)));       // X[H-k]
    }
}
RCG_HD void pair_phase_spectrum_copyout(const cf* S, int r, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<kPairLog2HL>;
    cf* X = prm.spectrum + ((long long)channel * prm.spec_frames + (frame - prm.spec_frame0)) * kPairW;
    for (int b = t; b < kPairHL; b += C::NT) {
        const cf v = S[phys(bin_to_pos<kPairLog2HL>(b))];
        const int k = 2 * b + r;
        if (k == 0) {
            X[0] = make_cf(v.x, 0.0f);
            X[kPairH] = make_cf(v.y, 0.0f);
        } else {
            X[k] = v;
            X[kPairW - k] = make_cf(v.x, -v.y);
        }
    }
}
RCG_HD void pair_phase_mags_copyin(cf* S, int r, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<kPairLog2HL>;
    const cf* X = prm.spectrum + ((long long)channel * prm.spec_frames + (frame - prm.spec_frame0)) * kPairW;
    for (int b = t; b < kPairHL; b += C::NT) {
        const int k = 2 * b + r;
        const cf lo = ld_once(X + k), hi = ld_once(X + (k == 0 ? kPairH : kPairW - k));
        S[phys(bin_to_pos<kPairLog2HL>(b))] = make_cf(cf_norm(lo), cf_norm(hi));
    }
}
// odd bins, magnitudes from the kernel's output: slot of k holds (|P[k]|, |P[W-k]|)
RCG_HD void pair_phase_spectral_in_odd(cf* S, int t, const StretchParams& prm, const PairTables& pt, int channel,
                                       long long frame) {
    using C = FrameCfg<kPairLog2HL>;
    const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
    const uint32_t chan = (uint32_t)(channel + prm.channel_offset);
    for (int c = t; c < kPairHL / 2; c += C::NT) {
        const int pos = spectral_slot_pos(kPairLog2HL, c);
        const int j = pos_to_bin<kPairLog2HL>(pos);
        const int k = 2 * j + 1;
        const int pa = phys(pos), pb = phys(bin_to_pos<kPairLog2HL>(kPairHL - 1 - j));
        const cf ma = S[pa], mb = S[pb];
        const cf tw = ld_cf(pt.spec_tw_odd + c);
        uint32_t u[4];
        philox4x32_10((uint32_t)k, f_lo, f_hi, chan, prm.keys, u);
        float s0, c0, s1, c1, s2, c2, s3, c3;
        sincos_phase(phase_from_u32(u[0]), s0, c0);
        sincos_phase(phase_from_u32(u[1]), s1, c1);
        sincos_phase(phase_from_u32(u[2]), s2, c2);
        sincos_phase(phase_from_u32(u[3]), s3, c3);
        float ykr = 0.5f * (ma.x * c0 + ma.y * c1), yki = 0.5f * (ma.x * s0 - ma.y * s1);
        float ybr = 0.5f * (mb.x * c2 + mb.y * c3), ybi = 0.5f * (mb.x * s2 - mb.y * s3);
        float ar = ykr + ybr, ai = yki - ybi, br = ykr - ybr, bi = yki + ybi;
        float cr = -(tw.x * bi - tw.y * br), ci = tw.x * br + tw.y * bi;
        S[pa] = make_cf(ar + cr, ai + ci);
        S[pb] = make_cf(ar - cr, -(ai - ci));
    }
}

// ---- cross-CTA radix-2 DIT stage + overlap-add -------------------------------------------------------
// After the local inverses CTA 0 holds u[n], CTA 1 holds v[n], n < HL; the last DIT stage is
//     first half window [n] = u[n] + conj(w^n) v[n]      second half window [n] = u[n] - conj(w^n) v[n].
// CTA r finishes BOTH half windows for the packed indices n in [r HL/2, (r+1) HL/2): it needs its own buffer and only
// half of the peer's.  That half is PUSHED: before the phase each CTA writes the part of its buffer the peer needs
// into the peer's landing zone (distributed shared memory, posted stores; one cluster barrier makes them visible),
// so no remote load sits on the critical path, and the carried half window of an index range is read and written
// by the same CTA -- it never crosses the pair.  The push is one bulk asynchronous copy per CTA (cp.async.bulk
// shared::cta -> shared::cluster, issued by one thread, completion counted in bytes on an mbarrier in the receiving
// CTA): no thread moves data, and nobody waits for more than its own landing zone.
constexpr int kPairNR = kPairHL / 2;                       // packed indices per CTA in this phase
constexpr int kPairLandCf = kPairNR + (kPairNR >> 4);      // landing zone: the padded slots of kPairNR elements

// What CTA r hands to its peer: the padded slots of the packed indices the peer will finish.  CTA 0 gives [NR, HL) to
// CTA 1, CTA 1 gives [0, NR) to CTA 0; NR is a multiple of 16, so the padded slots of a range are contiguous
// (phys(n) - phys(base) = phys(n - base)): ONE contiguous block of kPairLandCf values, which the device moves with a
// single bulk asynchronous copy into the peer's shared memory (Exec::exchange, rcg_pair_body.h).
RCG_HD const cf* pair_push_source(const cf* S, int r) { return S + phys((1 - r) * kPairNR); }

// first half window: /W * window, + carried half, * envelope * amp, emit (pitch 1) or stage in global memory for the
// resampling gather; second half window -> carried buffer (same index, same thread: in place)
RCG_HD void pair_phase_ola(const cf* S, const cf* land, int r, int t, const StretchParams& prm, const PairTables& pt,
                           int channel, long long frame, float* tail, float* stage, bool emit, bool write_head) {
    using C = FrameCfg<kPairLog2HL>;
    const bool bypass = prm.flags & kFlagBypass;
    float* out = prm.out + (long long)channel * prm.out_stride;
    const float scale = bypass ? 1.0f : prm.inv_w;
    // the carried half window is read and written in place by the same thread; the next iteration's value is
    // fetched before this iteration's store so that the load does not wait behind it
    cf old_next = ld_pair(tail + 2 * (r * kPairNR + t));
RCG_UNROLL(RCG_PAIR_OLA_UNROLL)
    for (int i = t; i < kPairNR; i += C::NT) {
        const int n = r * kPairNR + i;
        const cf old = old_next;
        if (i + C::NT < kPairNR) old_next = ld_pair(tail + 2 * (n + C::NT));
        const cf own = S[phys(n)], got = land[phys(i)];
        const cf u = r == 0 ? own : got;      // CTA 0's buffer
        const cf v = r == 0 ? got : own;      // CTA 1's buffer
        cf za, zb;                            // first / second half window, before scaling
        if (bypass) {
            za = u; zb = v;                   // the load phase left the two half windows in the two buffers
        } else {
            const cf w = pair_cross_tw(pt.cross_tw, n);  // conj(w) v
            const float pr = v.x * w.x + v.y * w.y, pi = v.y * w.x - v.x * w.y;
            za = make_cf(u.x + pr, u.y + pi);
            zb = make_cf(u.x - pr, u.y - pi);
        }
        const cf wa = ld_stream2(prm.window + 2 * n), wb = ld_stream2(prm.window + kPairH + 2 * n);
        const float y0 = RCG_MUL(RCG_MUL(za.x, scale), wa.x), y1 = RCG_MUL(RCG_MUL(za.y, scale), wa.y);
        st_pair(tail + 2 * n, RCG_MUL(RCG_MUL(zb.x, scale), wb.x), RCG_MUL(RCG_MUL(zb.y, scale), wb.y));
        const cf e = ld_stream2(prm.env + 2 * n);
        if (write_head) st_pair(prm.head_out + (long long)channel * kPairH + 2 * n, y0, y1);
        if (emit) {
            const float z0 = RCG_MUL(RCG_MUL(RCG_ADD(y0, old.x), e.x), prm.amp);
            const float z1 = RCG_MUL(RCG_MUL(RCG_ADD(y1, old.y), e.y), prm.amp);
            if (prm.pitch == 1) st_pair(out + (frame * (long long)kPairH + 2 * n - prm.out_origin), z0, z1);
            else st_pair(stage + 2 * n, z0, z1);
        }
    }
}

// resample from the staged half window in global memory: thread tg of nthr (both CTAs of the pair share the block)
RCG_HD void pair_phase_resample(const float* stage, int tg, int nthr, const StretchParams& prm, int channel, long long frame) {
    float* out = prm.out + (long long)channel * prm.out_stride;
    const int p = prm.pitch;
    if (p > 1) {
        const long long g0 = frame * (long long)kPairH;
        const long long o_begin = (g0 + p - 1) / p, o_end = (g0 + kPairH + p - 1) / p;
        for (long long o = o_begin + tg; o < o_end; o += nthr) out[o - prm.out_origin] = stage[(int)(o * p - g0)];
    } else if (p < -1) {
        const int q = -p;
        const int n = (prm.samples_needed - 1) * q;
        float* dst = out + (frame * (long long)n - prm.out_origin);
        const int nt_q = (nthr / q) * q;   // a multiple of q threads: a thread's j and ratio never change (see phase_resample)
        if (tg < nt_q) {
            const int di = nt_q / q;
            int i = tg / q;
            const float ratio = prm.lerp_ratio[tg - i * q];
            // running pointers and a trip count known before the loop (see phase_resample)
            const float* sp = stage + i;
            float* dp = dst + tg;
            const int trips = (n - tg + nt_q - 1) / nt_q;
RCG_UNROLL(4)
            for (int k = 0; k < trips; k++, sp += di, dp += nt_q) {
                const float cur = sp[0], nxt = sp[1];
                *dp = RCG_ADD(cur, RCG_MUL(RCG_SUB(nxt, cur), ratio));
            }
        }
    }
}

// CTA r's entry of a per-CTA pointer pair.  On the device both entries are the CTA's own buffer (the host emulator
// runs the two CTAs of a pair with two buffers), and the index must be a constant there: indexed by the runtime rank
// the array lives in local memory, the pointer comes back without its address space, and every shared-memory access
// of the kernel turns into a generic load / store on the long scoreboard (SASS before: 180 LD + 189 ST against 44 LDS
// + 9 STS, 32 bytes of stack; ncu: long scoreboard 3.25 warps per issue cycle, lg_throttle in the butterflies).
RCG_HD cf* pair_buf(cf* const (&b)[2], int r) {
#if defined(__CUDA_ARCH__)
    (void)r;
    return b[0];
#else
    return b[r];
#endif
}

// The same stage with the frame read ONCE per pair (the stretch path; the analysis path above reads it twice and never
// synchronises the pair): CTA r loads the packed indices n in [r NR, (r+1) NR) of both half windows, forms BOTH outputs
// of the stage, keeps its own and stores the peer's straight into the peer's transform buffer at the same slot
// (Exec::store_peer: a posted distributed-shared-memory store on the device).  The cluster barrier that ends the phase
// makes those stores visible.  Half the bytes per CTA (the phase ran at ~47 of the ~64 bytes per clock an SM gets from
// L2: 15 % of the frame time), and the twiddle products of the stage are shared between the CTAs instead of all
// falling on CTA 1.  Same operations on the same operands as pair_phase_load_impl: bit-identical results.
template <bool FAST, class Exec>
RCG_HD void pair_phase_load_split_impl(Exec& ex, cf* const (&bufs)[2], int r, int t, const StretchParams& prm,
                                       const PairTables& pt, const float* in, long long base, bool bypass) {
    using C = FrameCfg<kPairLog2HL>;
    cf* S = pair_buf(bufs, r);
    auto sample = [&](long long i) { return (i >= 0 && i < prm.in_valid) ? ld_stream(in + i) : 0.0f; };
RCG_UNROLL(RCG_PAIR_LOAD_UNROLL)
    for (int i = t; i < kPairNR; i += C::NT) {
        const int n = r * kPairNR + i;
        const cf wa = ld_stream2(prm.window + 2 * n), wb = ld_stream2(prm.window + kPairH + 2 * n);
        cf xa, xb;
        if constexpr (FAST) {
            xa = ld_stream2(in + base + 2 * n);
            xb = ld_stream2(in + base + kPairH + 2 * n);
        } else {
            xa = make_cf(sample(base + 2 * n), sample(base + 2 * n + 1));
            xb = make_cf(sample(base + kPairH + 2 * n), sample(base + kPairH + 2 * n + 1));
        }
        const float a0 = RCG_MUL(xa.x, wa.x), a1 = RCG_MUL(xa.y, wa.y);
        const float b0 = RCG_MUL(xb.x, wb.x), b1 = RCG_MUL(xb.y, wb.y);
        cf v0, v1;   // CTA 0's / CTA 1's value
        if (bypass) {
            v0 = make_cf(a0, a1);
            v1 = make_cf(b0, b1);
        } else {
            v0 = make_cf(a0 + b0, a1 + b1);
            const cf w = pair_cross_tw(pt.cross_tw, n);
            const float dr = a0 - b0, di = a1 - b1;
            v1 = make_cf(dr * w.x - di * w.y, di * w.x + dr * w.y);
        }
        const int slot = phys(n);
        S[slot] = r == 0 ? v0 : v1;
        ex.store_peer(bufs, r, slot, r == 0 ? v1 : v0);
    }
}
template <class Exec>
RCG_HD void pair_phase_load_split(Exec& ex, cf* const (&bufs)[2], int r, int t, const StretchParams& prm,
                                  const PairTables& pt, int channel, long long frame) {
    const float* in = prm.in + (long long)channel * prm.in_stride;
    const long long base = frame * (long long)prm.step - prm.in_origin;
    const bool bypass = prm.flags & kFlagBypass;
    const bool fast = base >= 0 && base + kPairW <= prm.in_valid &&
                      (reinterpret_cast<unsigned long long>(in + base) & 7ull) == 0;
    if (fast) pair_phase_load_split_impl<true>(ex, bufs, r, t, prm, pt, in, base, bypass);
    else pair_phase_load_split_impl<false>(ex, bufs, r, t, prm, pt, in, base, bypass);
}

// Exec for the pair: par(f) runs f(r, t) for every thread of both CTAs, then a CTA barrier;
// par_cluster(f) ends in a cluster barrier (release/acquire) instead; cluster_arrive() / cluster_wait() are the two
// halves of one more cluster barrier, issued apart.  bufs[r] is CTA r's buffer.
// passes FIRST..LAST of the local transform one by one (each ends in a CTA barrier)
template <int PASS, int LAST, bool INV, class Exec>
RCG_HD void pair_fft_range(Exec& ex, cf* const (&bufs)[2], const StretchParams& prm, bool work) {
    if constexpr (INV ? PASS >= LAST : PASS <= LAST) {
        ex.par([&](int r, int t) {
            if (!work) return;
            if constexpr (FrameCfg<kPairLog2HL>::TW_TMEM) phase_fft_pass<kPairLog2HL, PASS, INV>(pair_buf(bufs, r), t, tw_src(prm));
            else phase_fft_pass<kPairLog2HL, PASS, INV>(pair_buf(bufs, r), t, PASS == 0 ? prm.tw0 : prm.tw);
        });
        pair_fft_range<INV ? PASS - 1 : PASS + 1, LAST, INV>(ex, bufs, prm, work);
    }
}
// all passes 0..NP-1 (forward) / NP-1..0 (inverse); the last two run as one warp-synchronous phase
template <bool INV, class Exec>
RCG_HD void pair_fft_all(Exec& ex, cf* const (&bufs)[2], const StretchParams& prm, bool work) {
    constexpr int NP = FrameCfg<kPairLog2HL>::NP;
    if constexpr (!INV) pair_fft_range<0, NP - 3, false>(ex, bufs, prm, work);
    ex.par_steps(fft_tail_steps<kPairLog2HL>(),
                 [&](int step, int r, int t) { if (work) fft_tail_step<kPairLog2HL, INV>(pair_buf(bufs, r), t, step, tw_src(prm)); });
    if constexpr (INV) pair_fft_range<NP - 3, 0, true>(ex, bufs, prm, work);
}

// bufs[r] = CTA r's transform buffer, land[r] = CTA r's landing zone.  One cluster barrier per frame, issued in two
// halves: `arrive` after the overlap-add (I am done with my landing zone, and -- my zone was complete -- the copy out
// of the PEER's buffer has been consumed), `wait` at the top of the next frame (so the peer's zone is free for my next
// push and my own buffer, the source of my last push, may be overwritten).  The exchange itself synchronises on the
// receiving CTA's mbarrier only.
template <class Exec>
RCG_HD void run_work_item_pair(Exec& ex, cf* const (&bufs)[2], cf* const (&land)[2], const StretchParams& prm,
                               const PairTables& pt, int item, float* scratch, bool active) {
    using C = FrameCfg<kPairLog2HL>;
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
    const bool spectrum_in = prm.flags & kFlagSpectrumIn;
    float* tail = scratch;                 // [H]: carried half window; CTA r owns the samples of its index range
    float* stage = scratch + 2 * kPairH;   // [H]: finished half window for the resampling gather
    ex.par([&](int r, int t) {
        if (!active) return;
        const float* src = have_tail_in ? prm.tail_in + (long long)channel * kPairH : nullptr;
        for (int i = 2 * r * kPairNR + t; i < 2 * (r + 1) * kPairNR; i += C::NT) tail[i] = src ? src[i] : 0.0f;
    });
    ex.cluster_arrive();   // pairs with the first frame's wait
    const int n_iter = prm.run_len + 1;
    for (int it = 0; it < n_iter; it++) {
        ex.cluster_wait();   // the peer is through with the previous frame: its zone is free, my buffer has been read
        const long long f = f0 - 1 + it;
        const bool is_halo = (it == 0);
        const bool work = active && (is_halo ? halo : f < f1);
        const bool head = work && defer && f == f0;
        const bool emit = work && !is_halo && !head;
        if (spectrum_in) {
            ex.par([&](int r, int t) { if (work) pair_phase_mags_copyin(pair_buf(bufs, r), r, t, prm, channel, f); });
            ex.par([&](int r, int t) {
                if (!work) return;
                if (r == 0) {
                    StretchParams pe = prm;
                    pe.spec_tw = pt.spec_tw_even;
                    phase_spectral_in<kPairLog2HL, 2>(bufs[0], t, pe, channel, f);
                } else {
                    pair_phase_spectral_in_odd(bufs[1], t, prm, pt, channel, f);
                }
            });
            pair_fft_all<true>(ex, bufs, prm, work);
        } else {
        // (the wait above: the peer has finished the last frame's overlap-add, so its buffer may be written)
#if RCG_PAIR_SPLIT_LOAD
        ex.par_pair([&](int r, int t) { if (work) pair_phase_load_split(ex, bufs, r, t, prm, pt, channel, f); });
#else
        ex.par([&](int r, int t) { if (work) pair_phase_load(pair_buf(bufs, r), r, t, prm, pt, channel, f); });
#endif
        if (!bypass) {
            pair_fft_all<false>(ex, bufs, prm, work);
            ex.par([&](int r, int t) {
                if (!work) return;
                if (r == 0) {
                    StretchParams pe = prm;
                    pe.spec_tw = pt.spec_tw_even;
                    phase_spectral<kPairLog2HL, 2>(bufs[0], t, pe, channel, f);
                } else {
                    pair_phase_spectral_odd(bufs[1], t, prm, pt, channel, f);
                }
            });
            pair_fft_all<true>(ex, bufs, prm, work);
        }
        }
        if (work) ex.exchange(bufs, land, kPairLandCf);   // each CTA's share of the other buffer lands in its zone
        auto ola = [&](int r, int t) {
            if (work) pair_phase_ola(pair_buf(bufs, r), pair_buf(land, r), r, t, prm, pt, channel, f, tail, stage, emit, head);
        };
        if (prm.pitch != 1) {
            ex.par_cluster(ola);   // the staged half window crosses the pair through global memory
            ex.par([&](int r, int t) { if (emit) pair_phase_resample(stage, r * C::NT + t, 2 * C::NT, prm, channel, f); });
            // (the stage is written again only after the next frame's wait below has seen the peer's arrive)
        } else {
            ex.par(ola);
        }
        ex.cluster_arrive();   // done with my landing zone
    }
    ex.cluster_wait();
    if (prm.tail_out != nullptr) {
        ex.par([&](int r, int t) {
            if (!active || f1 != prm.frame_end) return;
            float* dst = prm.tail_out + (long long)channel * kPairH;
            for (int i = 2 * r * kPairNR + t; i < 2 * (r + 1) * kPairNR; i += C::NT) dst[i] = tail[i];
        });
    }
}

// frequency-kernel path: analysis of a run of frames; the two CTAs never exchange data here
template <class Exec>
RCG_HD void run_analysis_item_pair(Exec& ex, cf* const (&bufs)[2], const StretchParams& prm, const PairTables& pt,
                                   int item, bool active) {
    const int channel = active ? item / prm.n_runs : 0;
    const int run = active ? item - channel * prm.n_runs : 0;
    const long long f0 = prm.frame_begin + (long long)run * prm.run_len;
    for (int it = 0; it < prm.run_len; it++) {
        const long long f = f0 + it;
        const bool work = active && f < prm.frame_end;
        ex.par([&](int r, int t) { if (work) pair_phase_load(pair_buf(bufs, r), r, t, prm, pt, channel, f); });
        pair_fft_all<false>(ex, bufs, prm, work);
        ex.par([&](int r, int t) {
            if (!work) return;
            if (r == 0) {
                StretchParams pe = prm;
                pe.spec_tw = pt.spec_tw_even;
                phase_spectrum_split<kPairLog2HL>(bufs[0], t, pe);
            } else {
                pair_phase_spectrum_split_odd(bufs[1], t, pt);
            }
        });
        ex.par([&](int r, int t) { if (work) pair_phase_spectrum_copyout(pair_buf(bufs, r), r, t, prm, channel, f); });
    }
}

// boundary block at W = 65536 (one CTA per channel, staging in the cluster scratch of slot `channel`)
template <class ExecCta>
RCG_HD void run_boundary_pair(ExecCta& ex, const StretchParams& prm, int channel, const float* head,
                              const float* tail_prev, float* stage) {
    using C = FrameCfg<kPairLog2HL>;
    float* out = prm.out + (long long)channel * prm.out_stride;
    const long long frame = prm.frame_begin;
    ex.par([&](int t) {
        const float* h = head + (long long)channel * kPairH;
        const float* tp = tail_prev + (long long)channel * kPairH;
        for (int i = t; i < kPairH; i += C::NT) {
            const float z = RCG_MUL(RCG_MUL(RCG_ADD(h[i], tp[i]), prm.env[i]), prm.amp);
            if (prm.pitch == 1) out[frame * (long long)kPairH + i - prm.out_origin] = z;
            else stage[i] = z;
        }
    });
    if (prm.pitch != 1) ex.par([&](int t) { pair_phase_resample(stage, t, C::NT, prm, channel, frame); });
}

}  // namespace rcg

// rcg_emul.cpp -- TEST INFRASTRUCTURE.  Executes the kernel's per-frame phases
// (rocoder_b200/csrc/rcg_frame.h) on the host, one emulated thread at a time with a
// barrier between phases, so that the CPU test suite can check the kernel's indexing,
// digit-reversed spectral pairing, Philox slotting, overlap-add and resampling against the
// oracle without a GPU.  Never loaded by the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#if defined(RCG_EMUL_LOCAL_PLUGIN)
// Second build of this file (librcg_emul_local.so): the hook that fuses a bin-local frequency kernel into the
// spectral step (RCG_PLUGIN_LOCAL, rcg_jit_entry.h) executed on the host, with the `tilt` kernel of
// tests/plugins.py standing in for the user's source.
#include "../../rocoder_b200/csrc/rcg_common.h"
typedef cf float2;
static inline float2 make_float2(float x, float y) { return make_cf(x, y); }
static inline float2 apply(size_t, unsigned bin, unsigned n_bins, float2 v) {
    const float g = 0.25f + (float)bin / (float)n_bins;
    return make_float2((v.x * 0.6f - v.y * 0.8f) * g, (v.x * 0.8f + v.y * 0.6f) * g);
}
#define RCG_PLUGIN_LOCAL 1
#endif

#include "../../rocoder_b200/csrc/rcg_frame.h"
#include "../../rocoder_b200/csrc/rcg_frame_pair.h"
#include "../../rocoder_b200/csrc/rcg_frame_bluestein.h"
#include "../../rocoder_b200/csrc/rcg_tables.h"

using namespace rcg;

// Threads of a phase run one after the other.  A phase must not depend on the order (no thread may
// read what another wrote in the same phase): the tests run every case in ascending and descending
// thread order and demand identical results, which exposes intra-phase races.
static int g_reverse_threads = 0;
struct HostExec {
    int nt;
    template <class F>
    void par(F f) {
        if (g_reverse_threads) for (int t = nt - 1; t >= 0; t--) f(t);
        else for (int t = 0; t < nt; t++) f(t);
    }
    // warp by warp: all steps of one warp before the next warp starts, so any dependency that
    // crosses a warp boundary between steps shows up as a wrong result
    template <class F>
    void par_steps(int n, F f) {
        const int nw = (nt + 31) / 32;
        for (int wi = 0; wi < nw; wi++) {
            const int w = g_reverse_threads ? nw - 1 - wi : wi;
            for (int s = 0; s < n; s++)
                for (int li = 0; li < 32; li++) {
                    const int l = g_reverse_threads ? 31 - li : li;
                    if (w * 32 + l < nt) f(s, w * 32 + l);
                }
        }
    }
};

static int g_disable_half_rate = 0;
static std::vector<float> lerp_table(int pitch) {
    const int q = pitch < 0 ? -pitch : 1;
    std::vector<float> r(128, 0.0f);
    for (int j = 0; j < q && j < 128; j++) r[j] = (float)j / (float)q;
    return r;
}

template <int LOG2H>
static int emul_run(StretchParams prm) {
    using C = FrameCfg<LOG2H>;
    const int W = C::W, H = C::H;
    std::vector<float> window(W), env(H);
    host_hanning(window.data(), W);
    host_crossfade_compensation(env.data(), H);
    std::vector<cf> tw = host_pass_twiddles(LOG2H), stw = host_spectral_twiddles(LOG2H);
    std::vector<cf> stw_hr = host_spectral_twiddles_hr(LOG2H);
    std::vector<float> we = host_decimate2(window), ee = host_decimate2(env);
    prm.window = window.data(); prm.env = env.data(); prm.tw = tw.data(); prm.tw0 = tw.data(); prm.spec_tw = stw.data();
    prm.spec_tw_hr = stw_hr.data(); prm.window_even = we.data(); prm.env_even = ee.data();
    std::vector<float> twi;   // the Tensor Memory rows of the device kernel, read straight from their image here
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    std::vector<uint2> ptab = host_pair_table(LOG2H);
    std::vector<uint4> qtab = host_quad_table(LOG2H);
    prm.pair_tab = ptab.data(); prm.quad_tab = qtab.data();
    if (use_half_rate(prm.pitch, prm.flags) && !g_disable_half_rate) prm.flags |= kFlagHalfRate;
    if (C::TW_TMEM) { twi = host_tmem_image(LOG2H, (prm.flags & kFlagHalfRate) != 0, window.data(), env.data()); prm.tw_image = twi.data(); }
    // what with_mode_constants does on the device for the full-rate modes: window / envelope from the image
    if (window_in_tmem(LOG2H) && stretch_mode_of(prm.pitch, prm.flags) >= 2) prm.flags |= kFlagWinTmem;
    std::vector<cf> S(C::SMEM_CF);
    std::vector<float> tail(H);
    HostExec ex{C::NT};
    const int items = prm.channels * prm.n_runs;
    for (int item = 0; item < items; item++)
        run_work_item<LOG2H>(ex, S.data(), prm, item, tail.data(), true);
    return 0;
}

// W = 65536: both CTAs of the pair, phase by phase
struct PairHostExec {
    int nt;
    template <class F>
    void par(F f) {
        for (int r = 0; r < 2; r++) {
            const int rr = g_reverse_threads ? 1 - r : r;
            if (g_reverse_threads) for (int t = nt - 1; t >= 0; t--) f(rr, t);
            else for (int t = 0; t < nt; t++) f(rr, t);
        }
    }
    template <class F>
    void par_steps(int n, F f) {   // warp by warp, like HostExec::par_steps
        const int nw = (nt + 31) / 32;
        for (int r = 0; r < 2; r++)
            for (int wi = 0; wi < nw; wi++) {
                const int w = g_reverse_threads ? nw - 1 - wi : wi;
                for (int s = 0; s < n; s++)
                    for (int l = 0; l < 32; l++)
                        if (w * 32 + l < nt) f(s, r, w * 32 + l);
            }
    }
    template <class F>
    void par_cluster(F f) { par(f); }
    template <class F>
    void par_pair(F f) { par(f); }
    void store_peer(cf* const (&bufs)[2], int r, int slot, cf v) { bufs[r ^ 1][slot] = v; }
    void cluster_sync() {}
    void cluster_arrive() {}
    void cluster_wait() {}
    // what the device does with one bulk copy per CTA: each CTA's share of its buffer lands in the other's zone
    void exchange(cf* const (&bufs)[2], cf* const (&land)[2], int n_cf) {
        for (int r = 0; r < 2; r++) memcpy(land[r ^ 1], pair_push_source(bufs[r], r), sizeof(cf) * (size_t)n_cf);
    }
};

static int emul_run_pair(StretchParams prm) {
    using C = FrameCfg<kPairLog2HL>;
    std::vector<float> window(kPairW), env(kPairH);
    host_hanning(window.data(), kPairW);
    host_crossfade_compensation(env.data(), kPairH);
    std::vector<cf> tw = host_pass_twiddles(kPairLog2HL);
    std::vector<float> twi;   // the Tensor Memory rows of the device kernel, read straight from their image here
    if (FrameCfg<kPairLog2HL>::TW_TMEM) { twi = host_tmem_image(kPairLog2HL); prm.tw_image = twi.data(); }
    std::vector<cf> cross = host_pair_cross_twiddles(kPairLog2HL + 1);
    std::vector<cf> se = host_pair_spectral_twiddles(kPairLog2HL + 1, 0), so = host_pair_spectral_twiddles(kPairLog2HL + 1, 1);
    prm.window = window.data(); prm.env = env.data(); prm.tw = tw.data(); prm.tw0 = tw.data();
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    std::vector<uint2> ptab = host_pair_table(kPairLog2HL);
    prm.pair_tab = ptab.data();
    std::vector<float> scratch(3 * (size_t)kPairH);
    PairTables pt{cross.data(), se.data(), so.data(), scratch.data()};
    std::vector<cf> S0(C::SMEM_CF), S1(C::SMEM_CF);
    cf* bufs[2] = {S0.data(), S1.data()};
    std::vector<cf> L0(kPairLandCf), L1(kPairLandCf);
    cf* land[2] = {L0.data(), L1.data()};
    PairHostExec ex{C::NT};
    const int items = prm.channels * prm.n_runs;
    for (int item = 0; item < items; item++) run_work_item_pair(ex, bufs, land, prm, pt, item, scratch.data(), true);
    return 0;
}

static int emul_boundary_pair(StretchParams prm, const float* head, const float* tail_prev) {
    using C = FrameCfg<kPairLog2HL>;
    std::vector<float> env(kPairH), stage(kPairH);
    host_crossfade_compensation(env.data(), kPairH);
    prm.env = env.data();
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    HostExec ex{C::NT};
    for (int c = 0; c < prm.channels; c++) run_boundary_pair(ex, prm, c, head, tail_prev, stage.data());
    return 0;
}

template <int LOG2H>
static int emul_boundary(StretchParams prm, const float* head, const float* tail_prev) {
    using C = FrameCfg<LOG2H>;
    std::vector<float> env(C::H);
    host_crossfade_compensation(env.data(), C::H);
    std::vector<float> ee = host_decimate2(env);
    prm.env = env.data(); prm.env_even = ee.data();
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    if (use_half_rate(prm.pitch, prm.flags) && !g_disable_half_rate) prm.flags |= kFlagHalfRate;
    std::vector<cf> S(C::SMEM_CF);
    HostExec ex{C::NT};
    for (int c = 0; c < prm.channels; c++) run_boundary<LOG2H>(ex, S.data(), prm, c, head, tail_prev);
    return 0;
}

// windows that are not a power of two: the chirp-z path (rcg_frame_bluestein.h), one CTA's threads phase by phase
template <int LM>
static int emul_run_bluestein(StretchParams prm, int W) {
    using C = FrameCfg<LM>;
    const int H = W / 2;
    std::vector<float> window(W), env(H);
    host_hanning(window.data(), W);
    host_crossfade_compensation(env.data(), H);
    std::vector<cf> tw = host_pass_twiddles(LM), chirp = host_bluestein_chirp(W), filt = host_bluestein_filter(W, LM);
    prm.window = window.data(); prm.env = env.data(); prm.tw = tw.data(); prm.tw0 = tw.data();
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    BluesteinTables bt{chirp.data(), filt.data(), W};
    std::vector<cf> S(C::SMEM_CF);
    std::vector<float> zbuf(H), tail(H);
    HostExec ex{C::NT};
    const int items = prm.channels * prm.n_runs;
    for (int item = 0; item < items; item++) run_work_item_bluestein<LM>(ex, S.data(), zbuf.data(), prm, bt, item, tail.data(), true);
    return 0;
}
template <int LM>
static int emul_boundary_bluestein(StretchParams prm, int W, const float* head, const float* tail_prev) {
    using C = FrameCfg<LM>;
    const int H = W / 2;
    std::vector<float> env(H), zbuf(H);
    host_crossfade_compensation(env.data(), H);
    prm.env = env.data();
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    BluesteinTables bt{nullptr, nullptr, W};
    HostExec ex{C::NT};
    for (int c = 0; c < prm.channels; c++) run_boundary_bluestein<LM>(ex, zbuf.data(), prm, bt, c, head, tail_prev);
    return 0;
}

extern "C" {

// the same entry points for a window that is not a power of two (W even, 64..8192)
int rcg_emul_stretch_any(int W, const float* in, long long in_stride, long long in_origin, long long in_valid, float* out,
                         long long out_stride, long long out_origin, long long frame_begin, long long frame_end, int run_len,
                         int channels, int step, int pitch, int samples_needed, float amp, unsigned long long seed,
                         unsigned flags, int first_channel, const float* tail_in, float* tail_out, float* head_out) {
    StretchParams p;
    memset(&p, 0, sizeof(p));
    p.in = in; p.in_stride = in_stride; p.in_origin = in_origin; p.in_valid = in_valid;
    p.out = out; p.out_stride = out_stride; p.out_origin = out_origin;
    p.frame_begin = frame_begin; p.frame_end = frame_end; p.run_len = run_len;
    long long nf = frame_end - frame_begin;
    p.n_runs = nf > 0 ? (int)((nf + run_len - 1) / run_len) : 0;
    p.channels = channels; p.channel_offset = first_channel; p.step = step; p.pitch = pitch; p.samples_needed = samples_needed;
    p.amp = amp; p.flags = flags;
    p.keys = make_philox_keys(seed);
    p.tail_in = tail_in; p.tail_out = tail_out; p.head_out = head_out;
    switch (bluestein_log2m(W)) {
#define CASE(L) case L: return emul_run_bluestein<L>(p, W);
        CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
    }
    return -1;
}
int rcg_emul_boundary_any(int W, float* out, long long out_stride, long long out_origin, long long frame_begin, int channels,
                          int pitch, int samples_needed, float amp, const float* head, const float* tail_prev) {
    StretchParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.out_stride = out_stride; p.out_origin = out_origin;
    p.frame_begin = frame_begin; p.channels = channels; p.pitch = pitch;
    p.samples_needed = samples_needed; p.amp = amp;
    switch (bluestein_log2m(W)) {
#define CASE(L) case L: return emul_boundary_bluestein<L>(p, W, head, tail_prev);
        CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
    }
    return -1;
}

// force the full-rate synthesis path even for even pitch multiples (A/B checks)
void rcg_emul_disable_half_rate(int v) { g_disable_half_rate = v; }
void rcg_emul_reverse_threads(int v) { g_reverse_threads = v; }

// Processes frames [frame_begin, frame_end) of every channel.  See StretchParams.
int rcg_emul_stretch(int log2w, const float* in, long long in_stride, long long in_origin,
                     long long in_valid, float* out, long long out_stride, long long out_origin,
                     long long frame_begin, long long frame_end, int run_len, int channels, int step,
                     int pitch, int samples_needed, float amp, unsigned long long seed, unsigned flags,
                     int first_channel, const float* tail_in, float* tail_out, float* head_out) {
    StretchParams p;
    memset(&p, 0, sizeof(p));
    p.in = in; p.in_stride = in_stride; p.in_origin = in_origin; p.in_valid = in_valid;
    p.out = out; p.out_stride = out_stride; p.out_origin = out_origin;
    p.frame_begin = frame_begin; p.frame_end = frame_end; p.run_len = run_len;
    long long nf = frame_end - frame_begin;
    p.n_runs = nf > 0 ? (int)((nf + run_len - 1) / run_len) : 0;
    p.channels = channels; p.channel_offset = first_channel; p.step = step; p.pitch = pitch; p.samples_needed = samples_needed;
    p.amp = amp; p.inv_w = 1.0f / (float)(1 << log2w); p.flags = flags;
    p.keys = make_philox_keys(seed);
    p.tail_in = tail_in; p.tail_out = tail_out; p.head_out = head_out;
    if (log2w == 16) return emul_run_pair(p);
    switch (log2w - 1) {
#define CASE(L) case L: return emul_run<L>(p);
        CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
    }
    return -1;
}

int rcg_emul_boundary(int log2w, float* out, long long out_stride, long long out_origin,
                      long long frame_begin, int channels, int pitch, int samples_needed, float amp,
                      const float* head, const float* tail_prev) {
    StretchParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.out_stride = out_stride; p.out_origin = out_origin;
    p.frame_begin = frame_begin; p.channels = channels; p.pitch = pitch;
    p.samples_needed = samples_needed; p.amp = amp;
    if (log2w == 16) return emul_boundary_pair(p, head, tail_prev);
    switch (log2w - 1) {
#define CASE(L) case L: return emul_boundary<L>(p, head, tail_prev);
        CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
    }
    return -1;
}

}  // extern "C"

// Plugin path (fft.rs:76-108) on the host: analysis phases -> full spectra -> `apply` per frame ->
// synthesis phases with kFlagSpectrumIn.  Whole channels in one go (frames [0, frame_end)).
typedef int (*apply_fn)(size_t, const float*, size_t, float*);

template <int LOG2H>
static int emul_plugin(StretchParams prm, apply_fn fn) {
    using C = FrameCfg<LOG2H>;
    const int W = C::W, H = C::H;
    std::vector<float> window(W), env(H);
    host_hanning(window.data(), W);
    host_crossfade_compensation(env.data(), H);
    std::vector<cf> tw = host_pass_twiddles(LOG2H), stw = host_spectral_twiddles(LOG2H);
    prm.window = window.data(); prm.env = env.data(); prm.tw = tw.data(); prm.tw0 = tw.data(); prm.spec_tw = stw.data();
    std::vector<float> twi;
    if (C::TW_TMEM) { twi = host_tmem_image(LOG2H); prm.tw_image = twi.data(); }
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    std::vector<uint2> ptab = host_pair_table(LOG2H);
    prm.pair_tab = ptab.data();
    const long long nf = prm.frame_end - prm.frame_begin;
    std::vector<cf> spec((size_t)prm.channels * nf * W);
    prm.spectrum = spec.data(); prm.spec_frame0 = prm.frame_begin; prm.spec_frames = nf;
    std::vector<cf> S(C::SMEM_CF);
    std::vector<float> tail(H), tmp(2 * W);
    HostExec ex{C::NT};
    const int items = prm.channels * prm.n_runs;
    for (int item = 0; item < items; item++) run_analysis_item<LOG2H>(ex, S.data(), prm, item, true);
    for (long long f = 0; f < nf; f++)
        for (int c = 0; c < prm.channels; c++) {
            float* x = reinterpret_cast<float*>(spec.data() + ((size_t)c * nf + f) * W);
            if (fn && fn(1234, x, W, tmp.data()) == 0) memcpy(x, tmp.data(), sizeof(float) * 2 * W);
        }
    prm.flags |= kFlagSpectrumIn;
    for (int item = 0; item < items; item++) run_work_item<LOG2H, true>(ex, S.data(), prm, item, tail.data(), true);
    return 0;
}

// the same at W = 65536 (CTA pair)
static int emul_plugin_pair(StretchParams prm, apply_fn fn) {
    using C = FrameCfg<kPairLog2HL>;
    std::vector<float> window(kPairW), env(kPairH);
    host_hanning(window.data(), kPairW);
    host_crossfade_compensation(env.data(), kPairH);
    std::vector<cf> tw = host_pass_twiddles(kPairLog2HL);
    std::vector<float> twi;   // the Tensor Memory rows of the device kernel, read straight from their image here
    if (FrameCfg<kPairLog2HL>::TW_TMEM) { twi = host_tmem_image(kPairLog2HL); prm.tw_image = twi.data(); }
    std::vector<cf> cross = host_pair_cross_twiddles(kPairLog2HL + 1);
    std::vector<cf> se = host_pair_spectral_twiddles(kPairLog2HL + 1, 0), so = host_pair_spectral_twiddles(kPairLog2HL + 1, 1);
    prm.window = window.data(); prm.env = env.data(); prm.tw = tw.data(); prm.tw0 = tw.data();
    std::vector<float> lr = lerp_table(prm.pitch);
    prm.lerp_ratio = lr.data();
    std::vector<uint2> ptab = host_pair_table(kPairLog2HL);
    prm.pair_tab = ptab.data();
    std::vector<float> scratch(3 * (size_t)kPairH);
    PairTables pt{cross.data(), se.data(), so.data(), scratch.data()};
    const long long nf = prm.frame_end - prm.frame_begin;
    std::vector<cf> spec((size_t)prm.channels * nf * kPairW);
    prm.spectrum = spec.data(); prm.spec_frame0 = prm.frame_begin; prm.spec_frames = nf;
    std::vector<cf> S0(C::SMEM_CF), S1(C::SMEM_CF);
    cf* bufs[2] = {S0.data(), S1.data()};
    std::vector<float> tmp(2 * (size_t)kPairW);
    PairHostExec ex{C::NT};
    const int items = prm.channels * prm.n_runs;
    for (int item = 0; item < items; item++) run_analysis_item_pair(ex, bufs, prm, pt, item, true);
    for (long long f = 0; f < nf; f++)
        for (int c = 0; c < prm.channels; c++) {
            float* x = reinterpret_cast<float*>(spec.data() + ((size_t)c * nf + f) * kPairW);
            if (fn && fn(1234, x, kPairW, tmp.data()) == 0) memcpy(x, tmp.data(), sizeof(float) * 2 * kPairW);
        }
    prm.flags |= kFlagSpectrumIn;
    std::vector<cf> L0(kPairLandCf), L1(kPairLandCf);
    cf* land[2] = {L0.data(), L1.data()};
    for (int item = 0; item < items; item++) run_work_item_pair(ex, bufs, land, prm, pt, item, scratch.data(), true);
    return 0;
}

extern "C" {
int rcg_emul_stretch_plugin(int log2w, const float* in, long long in_stride, long long in_valid, float* out,
                            long long out_stride, long long frame_end, int run_len, int channels, int step,
                            int pitch, int samples_needed, float amp, unsigned long long seed, apply_fn fn) {
    StretchParams p;
    memset(&p, 0, sizeof(p));
    p.in = in; p.in_stride = in_stride; p.in_valid = in_valid;
    p.out = out; p.out_stride = out_stride;
    p.frame_begin = 0; p.frame_end = frame_end; p.run_len = run_len;
    p.n_runs = frame_end > 0 ? (int)((frame_end + run_len - 1) / run_len) : 0;
    p.channels = channels; p.step = step; p.pitch = pitch; p.samples_needed = samples_needed;
    p.amp = amp; p.inv_w = 1.0f / (float)(1 << log2w);
    p.keys = make_philox_keys(seed);
    if (log2w == 16) return emul_plugin_pair(p, fn);
    switch (log2w - 1) {
#define CASE(L) case L: return emul_plugin<L>(p, fn);
        CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
    }
    return -1;
}

// bank-conflict audit: for every warp-wide shared-memory access pattern of one frame,
// report how many 128-byte wavefronts the 64-bit accesses need (ideal: 2 per warp).
}

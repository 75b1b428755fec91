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

// unroll factors of the load and overlap-add loops (measured: 1..8 within 1 % of each other)
// columns per Tensor Memory load of a radix-8 / 16 butterfly's twiddles: 16 (8 twiddles) from W = 4096 up, 8 below
// (measured, profiles/r3j_*, r3k_*: -1.5 % at W = 16384 / 8192, -2 % at 32768, +1.3 % at 1024); -DRCG_TMEM_CHUNK=8|16
// forces one value for every window
#ifndef RCG_TMEM_CHUNK
#define RCG_TMEM_CHUNK 0
#endif
#define RCG_LOAD_UNROLL 4
#define RCG_OLA_UNROLL 2
// ... except at 512 threads per frame (W = 32768: one frame per SM, 16 warps), where the loops are latency-bound like
// those of the CTA pair (rcg_frame_pair.h) and the factors matter
#ifndef RCG_LOAD_UNROLL_512
#define RCG_LOAD_UNROLL_512 8   // W = 32768 f 4 p 1: 2.378 -> 2.325 ms, profiles/r3t_*
#endif
#ifndef RCG_OLA_UNROLL_512
#define RCG_OLA_UNROLL_512 2
#endif
#define RCG_PRAGMA(x) _Pragma(#x)
#define RCG_UNROLL(n) RCG_PRAGMA(unroll n)

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
    kFlagSpectrumIn = 4u,  // plugin path: magnitudes come from prm.spectrum (plugin output), not the FFT
    kFlagHalfRate = 8u,    // even pitch multiple > 0: only even output samples survive the decimation
    kFlagWinTmem = 16u,    // in-kernel only (with_mode_constants): window / envelope come from the Tensor Memory row
                           // (resampler.rs:15-18), so the synthesis side runs at half size
};

// the one rule both the engine and the test emulator use to pick the half-rate path
RCG_HD bool use_half_rate(int pitch, uint32_t flags) {
    return pitch > 0 && (pitch & 1) == 0 && !(flags & kFlagSpectrumIn);
}

// which compile-time specialisation of the stretch kernel serves a launch (rcg_kernel_body.h: MODE)
RCG_HD constexpr int stretch_mode_of(int pitch, uint32_t flags) {
    const uint32_t fl = flags & ~(uint32_t)kFlagDeferFirst;
    if (fl == kFlagHalfRate && pitch == 2) return 1;
    if (fl == 0u && pitch == 1) return 2;
    if (fl == 0u && pitch == -2) return 4;
    if (fl == 0u && pitch == -3) return 5;
    if (fl == 0u) return 3;
    return 0;
}

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
    const cf* tw0;          // the table of pass 0 (== tw unless a kernel moved the other passes' tables, see rcg_pair_body.h)
    const float* tw_image;  // [tmem_total_cols][NT] what the threads' Tensor Memory rows are filled from (tw_in_tmem, rcg_common.h)
    uint32_t tmem_addr;     // tw_in_tmem: Tensor Memory address of this thread's row / column half; set by the kernel prologue
    const cf* spec_tw;      // [H/2] exp(-2 pi i k/W) in the spectral step's pair order
    const cf* spec_tw_hr;   // [H/4] the same in the half-rate quad order
    const float* window_even;  // [H]   window[2i]   (half-rate synthesis)
    const float* env_even;     // [H/2] env[2i]
    const float* lerp_ratio;   // [|p|] j / |p| as f32 (resampler.rs:29), for pitch < -1
    const uint4* quad_tab;     // [H/4] half-rate spectral step: {pa | pb<<16, pc | pd<<16, k, 0} (padded slot indices)
    const uint2* pair_tab;     // [H/2] full-rate spectral step: {pa | pb<<16, k}
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
    // plugin path (fft.rs:76-108): full W-bin complex spectra, natural order,
    // [channel][frame - spec_frame0][W]
    cf* spectrum;
    long long spec_frame0;
    long long spec_frames;  // frames per channel in the buffer
    unsigned long long elapsed_ms;  // what a fused bin-local frequency kernel receives (fft.rs:89-92): the clock of
                                    // the launch, or -- RCG_FLAG_FRAME_CLOCK -- of frame 0 of the stream, to which
    unsigned long long clock_num;   // frame * clock_num / clock_den milliseconds are added (0 / 1: one clock per launch)
    unsigned long long clock_den;
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
    static constexpr bool TW_SMEM = tw_in_smem(LOG2H);   // prm.tw of the kernels points into shared memory
    static constexpr bool TW_TMEM = tw_in_tmem(LOG2H);   // pass twiddles come from the thread's Tensor Memory row
    static constexpr bool TW_TMEM_HR = tw_in_tmem(LOG2H, true);   // ... those of the half-rate plan too
    static constexpr int TW_CF = tw_total(LOG2H);
};

// ---- global loads ---------------------------------------------------------------------
// Input samples are re-read by W/step overlapping frames and the window / envelope by every
// frame: all of it must stay L2-resident, so these are plain read-only loads.  (Measured:
// ld.global.nc.L1::no_allocate is issued to L2 as evict-first on sm_100 and sent 7 GB per launch
// back to DRAM -- profiles/r1b_stretch_kernel_v1.md.)
RCG_HD float ld_stream(const float* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
RCG_HD cf ld_stream2(const float* p) {  // p must be 8-byte aligned
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const float2*>(p));
#else
    return make_cf(p[0], p[1]);
#endif
}
RCG_HD cf ld_cf(const cf* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// pass twiddles: read-only global tables, or the CTA's shared-memory copy (FrameCfg::TW_SMEM)
template <int LOG2H>
RCG_HD cf ld_tw(const cf* p) {
    // LOG2H == 14: the W = 65536 pair kernel keeps the tables of passes >= 1 in shared memory, the W = 32768 kernel
    // in global memory: a generic load serves both
    if constexpr (FrameCfg<LOG2H>::TW_SMEM || LOG2H == 14) return *p;
    else return ld_cf(p);
}
RCG_HD uint4 ld_u4(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
RCG_HD uint2 ld_u2(const uint2* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// single-use stream (spectra of the frequency-kernel path): keep it out of L1, where the twiddle tables, window
// and envelope live, and let L2 drop it first (what L1::no_allocate means to L2 on sm_100, see above)
RCG_HD cf ld_once(const cf* p) {
#if defined(__CUDA_ARCH__)
    cf r;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
#else
    return *p;
#endif
}
RCG_HD cf ld_pair(const float* p) { return *reinterpret_cast<const cf*>(p); }
RCG_HD void st_pair(float* p, float a, float b) { *reinterpret_cast<cf*>(p) = make_cf(a, b); }
RCG_HD void st_cf(float* p, cf v) { *reinterpret_cast<cf*>(p) = v; }

// ---- where a pass finds its twiddles -----------------------------------------------------------------------
// The packed per-pass tables (global or shared memory), and -- kernels with FrameCfg::TW_TMEM -- the thread's Tensor
// Memory row.  Code that passes a bare table pointer (CTA-pair and chirp-z kernels) gets the tables.
struct TwSrc {
    const cf* tw;
    uint32_t tmem;        // device: row / column-half base of this thread
    const float* image;   // host execution (tests/emul) reads the image the rows are filled from
    RCG_HD TwSrc(const cf* p) : tw(p), tmem(0u), image(nullptr) {}
    RCG_HD TwSrc(const cf* p, uint32_t a, const float* im) : tw(p), tmem(a), image(im) {}
};
RCG_HD TwSrc tw_src(const StretchParams& prm) { return TwSrc(prm.tw, prm.tmem_addr, prm.tw_image); }

// N = 2 or 8 consecutive columns of the set the butterfly with index m reads, starting at column `col` of the set.
// tcgen05.ld writes the registers asynchronously and nothing may read them before tcgen05.wait::ld, which is why the
// wait sits in the same statement.  (Issuing the next four twiddles before multiplying by the last four -- load and
// wait as separate statements, the registers passed through the wait as in/out operands -- was measured 3 % SLOWER on
// config 2: 6.11 vs 5.93 ms, profiles/r3d_tmem_pipelined_variants.jsonl.  The load's latency is 12 cycles.)
template <int LOG2H, int PASS, bool HR, int N>
RCG_HD void tw_fetch(const TwSrc& s, int m, int col, float (&v)[N]) {
    static_assert(N == 2 || N == 8 || N == 16, "tcgen05.ld shapes used here");
    constexpr int OFF = tmem_col_offset(LOG2H, PASS, HR), SC = tmem_set_cols(LOG2H, PASS, HR);
    constexpr int NT = threads_per_frame(LOG2H), M = 1 << make_plan(LOG2H, HR).mlog2[PASS];
    const int set = (M > NT) ? m / NT : 0;
#if defined(__CUDA_ARCH__)
    const uint32_t a = s.tmem + (uint32_t)(OFF + set * SC + col);
    uint32_t r[N];
    if constexpr (N == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(a));
    } else if constexpr (N == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(a));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]) : "r"(a));
    }
#pragma unroll
    for (int q = 0; q < N; q++) v[q] = __uint_as_float(r[q]);
#else
    const int row = tmem_row_of_m(LOG2H, PASS, HR, m);
    for (int q = 0; q < N; q++) v[q] = s.image[(size_t)(OFF + set * SC + col + q) * NT + row];
#endif
}

// N = 4 or 8 columns of the thread's row starting at absolute column `col`; host execution: row `row` of the image
template <int LOG2H, int N>
RCG_HD void row_fetch(const TwSrc& s, int row, int col, uint32_t (&r)[N]) {
    static_assert(N == 4 || N == 8 || N == 16, "tcgen05.ld shapes used here");
#if defined(__CUDA_ARCH__)
    (void)row;
    const uint32_t a = s.tmem + (uint32_t)col;
    if constexpr (N == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(a));
    } else if constexpr (N == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(a));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
    }
#else
    for (int q = 0; q < N; q++) memcpy(&r[q], &s.image[(size_t)(col + q) * threads_per_frame(LOG2H) + row], sizeof(float));
#endif
}
RCG_HD float bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, sizeof(f)); return f;
#endif
}

// ---- twiddles -------------------------------------------------------------------------
// Multiply element j (1 <= j < R) of a butterfly by w^(m*j) (or its conjugate), w = the pass's
// principal root, using the table layout tw_scheme() selected.
// the stored power-of-two exponents of the pow scheme, fetched ahead of the butterfly so that
// their (L1 / L2 / shared-memory) latency overlaps the data loads and the DFT
struct TwPre { cf w1, w2, w4, w8; };
template <int LOG2H, int PASS, int R, bool HR = false, bool TM = false>
RCG_HD TwPre tw_preload(const cf* tw, int m) {
    constexpr int SCHEME = TM ? (int)kTwNone : tw_scheme(LOG2H, PASS, HR);
    constexpr int M = 1 << make_plan(LOG2H, HR).mlog2[PASS];
    TwPre p;
    p.w1 = p.w2 = p.w4 = p.w8 = make_cf(1.0f, 0.0f);
    if constexpr (SCHEME == kTwPow) {
        p.w1 = ld_tw<LOG2H>(tw + m); p.w2 = ld_tw<LOG2H>(tw + M + m);
        if constexpr (R >= 8) p.w4 = ld_tw<LOG2H>(tw + 2 * M + m);
        if constexpr (R >= 16) p.w8 = ld_tw<LOG2H>(tw + 3 * M + m);
    }
    return p;
}

// Packed arithmetic (rcg_dft.h): a product with a twiddle whose rotated copy is at hand is two instructions.
template <int LOG2H, int PASS, bool CONJ, int R, bool HR = false, bool TM = false>
RCG_HD void twiddle_butterfly(cf (&x)[R], const TwSrc& src, int m, const TwPre& pre) {
    constexpr int SCHEME = tw_scheme(LOG2H, PASS, HR);
    constexpr int M = 1 << make_plan(LOG2H, HR).mlog2[PASS];
    const cf* tw = src.tw + tw_offset(LOG2H, PASS, HR);
    if constexpr (TM && SCHEME != kTwNone) {
        // every w^(m j) straight from the thread's Tensor Memory row, four at a time
        if constexpr (R == 2) {
            float v[2];
            tw_fetch<LOG2H, PASS, HR, 2>(src, m, 0, v);
            x[1] = tw_mul<CONJ>(x[1], make_tw(make_cf(v[0], v[1])));
        } else {
            // twiddles per load: 4 (8 columns), or 8 where the butterfly has that many
            constexpr int CHUNK = RCG_TMEM_CHUNK != 0 ? RCG_TMEM_CHUNK : (LOG2H >= 11 ? 16 : 8);
            constexpr int NCOLS = (CHUNK == 16 && R >= 8) ? 16 : 8, NTW = NCOLS / 2;
#pragma unroll
            for (int c = 0; c < R / NTW; c++) {
                float v[NCOLS];
                tw_fetch<LOG2H, PASS, HR, NCOLS>(src, m, NCOLS * c, v);
#pragma unroll
                for (int q = 0; q < NTW; q++) {
                    const int j = NTW * c + q;
                    if (j >= 1) x[j] = tw_mul<CONJ>(x[j], make_tw(make_cf(v[2 * q], v[2 * q + 1])));
                }
            }
        }
    } else if constexpr (SCHEME == kTwFull) {
#pragma unroll
        for (int j = 1; j < R; j++) x[j] = tw_mul<CONJ>(x[j], make_tw(ld_tw<LOG2H>(tw + (j - 1) * M + m)));
    } else if constexpr (SCHEME == kTwTwoLevel) {
        static_assert(SCHEME != kTwTwoLevel || R == 2, "two-level tables serve radix-2 passes");
        // w^m = A[m >> 6] B[m & 63]: the two factors are applied one after the other
        const Tw a = make_tw(ld_tw<LOG2H>(tw + (m >> 6))), b = make_tw(ld_tw<LOG2H>(tw + M / 64 + (m & 63)));
        x[1] = tw_mul<CONJ>(tw_mul<CONJ>(x[1], a), b);
    } else if constexpr (SCHEME == kTwPow) {
        // stored: w1, w2, w4, w8 (correctly rounded); kept products: w3 = w1 w2, w12 = w8 w4; an exponent with two
        // set bits beyond those is two products applied in sequence (depth <= 2 multiplications either way)
        const Tw t1 = make_tw(pre.w1), t2 = make_tw(pre.w2);
        const Tw t3 = tw_product(t1, t2);
        x[1] = tw_mul<CONJ>(x[1], t1);
        x[2] = tw_mul<CONJ>(x[2], t2);
        x[3] = tw_mul<CONJ>(x[3], t3);
        if constexpr (R >= 8) {
            const Tw t4 = make_tw(pre.w4);
            x[4] = tw_mul<CONJ>(x[4], t4);
            x[5] = tw_mul<CONJ>(tw_mul<CONJ>(x[5], t4), t1);
            x[6] = tw_mul<CONJ>(tw_mul<CONJ>(x[6], t4), t2);
            x[7] = tw_mul<CONJ>(tw_mul<CONJ>(x[7], t4), t3);
            if constexpr (R >= 16) {
                const Tw t8 = make_tw(pre.w8);
                const Tw t12 = tw_product(t8, t4);
                x[8] = tw_mul<CONJ>(x[8], t8);
                x[9] = tw_mul<CONJ>(tw_mul<CONJ>(x[9], t8), t1);
                x[10] = tw_mul<CONJ>(tw_mul<CONJ>(x[10], t8), t2);
                x[11] = tw_mul<CONJ>(tw_mul<CONJ>(x[11], t8), t3);
                x[12] = tw_mul<CONJ>(x[12], t12);
                x[13] = tw_mul<CONJ>(tw_mul<CONJ>(x[13], t12), t1);
                x[14] = tw_mul<CONJ>(tw_mul<CONJ>(x[14], t12), t2);
                x[15] = tw_mul<CONJ>(tw_mul<CONJ>(x[15], t12), t3);
            }
        }
    }
}

template <int LOG2H, int PASS, bool CONJ, int R, bool HR = false, bool TM = false>
RCG_HD void twiddle_butterfly(cf (&x)[R], const TwSrc& src, int m) {
    twiddle_butterfly<LOG2H, PASS, CONJ, R, HR, TM>(x, src, m, tw_preload<LOG2H, PASS, R, HR, TM>(src.tw + tw_offset(LOG2H, PASS, HR), m));
}

// shared-memory offset of butterfly input a relative to input 0: phys(base + a*M) - phys(base)
// is a compile-time constant because M is 1 or a multiple of 16
template <int M>
RCG_HD constexpr int slot_off(int a) {
    static_assert(M == 1 || (M % 16) == 0, "pass strides are 1 or multiples of 16");
    return M == 1 ? a : a * (M + (M >> 4));
}

// ---- LOAD fused with forward pass 0 ---------------------------------------------------------
// samples * window (fft.rs:51-58) -> even/odd packed complex z[n] = (x[2n], x[2n+1]) -> first
// DIF pass straight from global memory -> shared memory.
template <int LOG2H>
struct LoadCtx {
    const float* in;   // channel row
    long long base;    // local index of the frame's first sample
    bool bypass, aligned;
    TwSrc tw;
    RCG_HD LoadCtx() : tw(nullptr) {}
};
template <int LOG2H>
RCG_HD LoadCtx<LOG2H> make_load_ctx(const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    LoadCtx<LOG2H> c;
    c.in = prm.in + (long long)channel * prm.in_stride;
    c.base = frame * (long long)prm.step - prm.in_origin;
    c.bypass = prm.flags & kFlagBypass;
    const bool interior = c.base >= 0 && c.base + C::W <= prm.in_valid;
    c.aligned = interior && ((reinterpret_cast<unsigned long long>(c.in + c.base) & 7ull) == 0);
    c.tw = tw_src(prm);
    return c;
}
// global loads of butterfly m of pass 0 (packed samples n = m + a*M) times the window
// (ALIGNED: the whole frame lies inside the input and is 8-byte aligned -> branch-free float2 loads,
// which lets the compiler batch the loads of several unrolled iterations)
template <int LOG2H, int R, bool ALIGNED>
RCG_HD void load_fetch(const LoadCtx<LOG2H>& c, const StretchParams& prm, int m, cf (&x)[R]) {
    constexpr int M = 1 << make_plan(LOG2H).mlog2[0];
#pragma unroll
    for (int a = 0; a < R; a++) {
        const int n = m + a * M;  // packed index: samples 2n, 2n+1
        cf v;
        if constexpr (ALIGNED) {
            v = ld_stream2(c.in + c.base + 2 * n);
        } else {
            const long long i0 = c.base + 2 * n;
            v.x = (i0 >= 0 && i0 < prm.in_valid) ? ld_stream(c.in + i0) : 0.0f;
            v.y = (i0 + 1 >= 0 && i0 + 1 < prm.in_valid) ? ld_stream(c.in + i0 + 1) : 0.0f;
        }
        x[a] = ex_mul(v, ld_stream2(prm.window + 2 * n));   // samples * window (fft.rs:51-55), rounded before any add
    }
}
// first DIF pass on the fetched values -> shared memory
template <int LOG2H, int R, bool BYPASS = false>
RCG_HD void load_finish(cf* S, const LoadCtx<LOG2H>& c, int m, cf (&x)[R]) {
    constexpr int M = 1 << make_plan(LOG2H).mlog2[0];
    if constexpr (!BYPASS) {
        Dft<R>::run(x);
        twiddle_butterfly<LOG2H, 0, false, R, false, FrameCfg<LOG2H>::TW_TMEM>(x, c.tw, m);
    }
    cf* p = S + phys(m);
#pragma unroll
    for (int j = 0; j < R; j++) p[slot_off<M>(j)] = x[j];
}

template <int LOG2H>
RCG_HD void phase_load(cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    constexpr int R = 1 << make_plan(LOG2H).rlog2[0], M = 1 << make_plan(LOG2H).mlog2[0];
    const LoadCtx<LOG2H> c = make_load_ctx<LOG2H>(prm, channel, frame);
    if (c.bypass) {
        for (int m = t; m < M; m += C::NT) {
            cf x[R];
            load_fetch<LOG2H, R, false>(c, prm, m, x);
            load_finish<LOG2H, R, true>(S, c, m, x);
        }
    } else if (window_in_tmem(LOG2H) && (prm.flags & kFlagWinTmem) && c.aligned) {
        // the window from the thread's Tensor Memory row: 16 columns = the values of four iterations per load
        if constexpr (window_in_tmem(LOG2H)) {
            static_assert(R == 2 && tmem_window_iters(LOG2H) % 4 == 0, "4 window columns per iteration");
            for (int g = 0; g < tmem_window_iters(LOG2H) / 4; g++) {
                uint32_t wv[16];
                row_fetch<LOG2H, 16>(c.tw, t, tmem_window_offset(LOG2H) + 16 * g, wv);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int m = t + C::NT * (4 * g + u);
                    cf x[R];
#pragma unroll
                    for (int a = 0; a < R; a++)
                        x[a] = ex_mul(ld_stream2(c.in + c.base + 2 * (m + a * M)),
                                      make_cf(bits_to_float(wv[4 * u + 2 * a]), bits_to_float(wv[4 * u + 2 * a + 1])));
                    load_finish<LOG2H, R>(S, c, m, x);
                }
            }
        }
    } else if (c.aligned) {
RCG_UNROLL((LOG2H >= 14 ? RCG_LOAD_UNROLL_512 : RCG_LOAD_UNROLL))
        for (int m = t; m < M; m += C::NT) {
            cf x[R];
            load_fetch<LOG2H, R, true>(c, prm, m, x);
            load_finish<LOG2H, R>(S, c, m, x);
        }
    } else {
        for (int m = t; m < M; m += C::NT) {
            cf x[R];
            load_fetch<LOG2H, R, false>(c, prm, m, x);
            load_finish<LOG2H, R>(S, c, m, x);
        }
    }
}

// ---- one in-place butterfly of pass PASS >= 1 (pass 0 is fused into load / overlap-add) ---------
template <int LOG2H, int PASS, bool INV, bool TM = FrameCfg<LOG2H>::TW_TMEM>
RCG_HD void fft_butterfly(cf* S, const TwSrc& tw, int b) {
    constexpr int RL = make_plan(LOG2H).rlog2[PASS], R = 1 << RL, ML = make_plan(LOG2H).mlog2[PASS], M = 1 << ML;
    const int m = b & (M - 1), blk = b >> ML;
    const TwPre pre = tw_preload<LOG2H, PASS, R, false, TM>(tw.tw + tw_offset(LOG2H, PASS), m);
    cf* p = S + phys((blk << (ML + RL)) + m);
    cf x[R];
#pragma unroll
    for (int a = 0; a < R; a++) x[a] = p[slot_off<M>(a)];
    if constexpr (INV) {
        twiddle_butterfly<LOG2H, PASS, true, R, false, TM>(x, tw, m, pre);  // conj twiddle on the inputs (DIT)
        Dft<R, true>::run(x);
    } else {
        Dft<R>::run(x);
        twiddle_butterfly<LOG2H, PASS, false, R, false, TM>(x, tw, m, pre);  // twiddle on the outputs (DIF)
    }
#pragma unroll
    for (int j = 0; j < R; j++) p[slot_off<M>(j)] = x[j];
}

template <int LOG2H, int PASS, bool INV, bool TM = FrameCfg<LOG2H>::TW_TMEM>
RCG_HD void phase_fft_pass(cf* S, int t, const TwSrc& tw_all) {
    using C = FrameCfg<LOG2H>;
    constexpr int NB = C::H >> make_plan(LOG2H).rlog2[PASS];
    for (int b = t; b < NB; b += C::NT) fft_butterfly<LOG2H, PASS, INV, TM>(S, tw_all, b);
}

// The last two passes (strides 16 and 1) of a warp's butterflies b .. b+31 touch the same 512
// elements, so they run as consecutive steps of ONE phase separated by __syncwarp only:
// step 2i and 2i+1 = the two passes for butterfly index t + NT*i.
template <int LOG2H, bool INV>
RCG_HD void fft_tail_step(cf* S, int t, int step, const TwSrc& tw_all) {
    using C = FrameCfg<LOG2H>;
    constexpr int NB = C::H / 16;
    static_assert(make_plan(LOG2H).rlog2[C::NP - 1] == 4 && make_plan(LOG2H).rlog2[C::NP - 2] == 4, "radix-16 tail");
    const int b = t + C::NT * (step >> 1);
    if (b >= NB) return;
    const bool first = (step & 1) == 0;
    if (first != INV) fft_butterfly<LOG2H, C::NP - 2, INV>(S, tw_all, b);
    else fft_butterfly<LOG2H, C::NP - 1, INV>(S, tw_all, b);
}
template <int LOG2H>
constexpr int fft_tail_steps() { return 2 * (((FrameCfg<LOG2H>::H / 16) + FrameCfg<LOG2H>::NT - 1) / FrameCfg<LOG2H>::NT); }

// phases are k * 2^-23 * PI with k < 2^23, i.e. in [0, PI): the SFU's sin/cos (abs. error
// ~2^-21) is accurate far beyond the 1e-4 PCM tolerance and costs 3 instructions instead of ~90
RCG_HD void sincos_phase(float phi, float& s, float& c) {
#if defined(__CUDA_ARCH__)
    __sincosf(phi, &s, &c);
#else
    s = sinf(phi); c = cosf(phi);
#endif
}
RCG_HD float fast_sqrt(float v) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return sqrtf(v);
#endif
}

#if defined(RCG_PLUGIN_LOCAL)
// |apply(X[bin])| of the bin-local frequency kernel compiled into this translation unit (only the norm of a
// kernel's output survives, fft.rs:67)
struct StretchParams;
RCG_HD float plugin_mag_raw(unsigned long long ms, uint32_t bin, uint32_t n_bins, float xr, float xi) {
    const float2 p = apply((size_t)ms, (unsigned)bin, (unsigned)n_bins, make_float2(xr, xi));
    return fast_sqrt(p.x * p.x + p.y * p.y);
}
// `frame_ms` = the clock of the frame being processed, a local of the calling phase (frame_clock below)
#define plugin_mag(prm, bin, n_bins, xr, xi) plugin_mag_raw(frame_ms, (bin), (n_bins), (xr), (xi))
#endif
// elapsed_ms a frequency kernel sees for `frame`: the reference reads the wall clock before every frame
// (fft.rs:89-92); a launch computes thousands of frames at the same moment, so the per-frame clock is the frame's
// position in output time (what a run paced by playback would read), when the engine was asked for one
RCG_HD unsigned long long frame_clock(const StretchParams& prm, long long frame) {
    return prm.elapsed_ms + (prm.clock_num ? (unsigned long long)frame * prm.clock_num / prm.clock_den : 0ull);
}

// (cos, sin) of the phase word u: e^{i phi}
RCG_HD cf cis_phase(uint32_t u) {
    float s, c;
    sincos_phase(phase_from_u32(u), s, c);
    return make_cf(c, s);
}

// One bin pair (kk, H-kk) of the half-size forward transform: real-FFT split -> magnitudes of
// bins kk and H-kk -> the four random phases of bins {kk, W-kk, H-kk, H+kk} (one Philox block)
// -> Hermitian fold.  Returns A = Y[kk] + conj(Y[H-kk]) and B = Y[kk] - conj(Y[H-kk]) where
// Y[k] = |X[k]| (e^{i phi_k} + e^{-i phi_{W-k}}) / 2 is the Hermitian part of the reference's
// W-bin spectrum (fft.rs:65-68): the real part of its inverse FFT is the inverse of Y.
// t = exp(-2 pi i kk / W) with its rotated copy.  Packed arithmetic (rcg_dft.h): a conjugation or a rotation by
// -+i is the (1, -1) / (-1, 1) factor or the swapped operand of the instruction that consumes the value.
// kk = the pair's bin number, n_bins = W: only read when a bin-local frequency kernel is fused in (RCG_PLUGIN_LOCAL,
// rcg_jit_entry.h): the magnitudes are then those of the kernel's output for the four bins.
RCG_HD void pair_fold(cf za, cf zb, const Tw& t, const uint32_t (&u)[4], cf& A, cf& B,
                      const StretchParams& prm, uint32_t kk, uint32_t n_bins, unsigned long long frame_ms) {
    using SA = Ar<(RCG_PK_SPEC != 0)>;
    const cf PM = make_cf(1.0f, -1.0f), MP = make_cf(-1.0f, 1.0f);
    // P = za + conj(zb), Q = za - conj(zb), U = -i t Q;  X[kk] = (P + U)/2, conj(X[H-kk]) = (P - U)/2
    const cf P = SA::fma(zb, PM, za), Q = SA::fma(zb, MP, za);
    const cf tq = tw_mul<false, (RCG_PK_SPEC != 0)>(Q, t);
    const cf xk = SA::fma(pk_swap(tq), PM, P);   // halves folded into the 0.25 below
    const cf xb = SA::fma(pk_swap(tq), MP, P);
    const cf e0 = cis_phase(u[0]);  // bin kk
    const cf e1 = cis_phase(u[1]);  // bin W - kk
    const cf e2 = cis_phase(u[2]);  // bin H - kk
    const cf e3 = cis_phase(u[3]);  // bin H + kk
#if defined(RCG_PLUGIN_LOCAL)
    // X[kk] = (xk)/2, X[W-kk] = conj X[kk]; X[H-kk] = conj(xb)/2, X[H+kk] = conj X[H-kk]  (real input frame)
    const uint32_t hh = n_bins >> 1;
    const float m0 = plugin_mag(prm, kk, n_bins, 0.5f * xk.x, 0.5f * xk.y);
    const float m1 = plugin_mag(prm, n_bins - kk, n_bins, 0.5f * xk.x, -0.5f * xk.y);
    const float m2 = plugin_mag(prm, hh - kk, n_bins, 0.5f * xb.x, -0.5f * xb.y);
    const float m3 = plugin_mag(prm, hh + kk, n_bins, 0.5f * xb.x, 0.5f * xb.y);
    // Y = (m e + m' conj(e')) / 2
    const cf yk = SA::scale(SA::fma(SA::mul(e1, PM), pk_bcast(m1), SA::scale(e0, m0)), 0.5f);
    const cf yb = SA::scale(SA::fma(SA::mul(e3, PM), pk_bcast(m3), SA::scale(e2, m2)), 0.5f);
#else
    (void)prm; (void)kk; (void)n_bins; (void)frame_ms;
    const float mk = fast_sqrt(xk.x * xk.x + xk.y * xk.y), mb = fast_sqrt(xb.x * xb.x + xb.y * xb.y);
    const cf yk = SA::scale(SA::fma(e1, PM, e0), 0.25f * mk);   // |X[kk]| (e0 + conj(e1)) / 2
    const cf yb = SA::scale(SA::fma(e3, PM, e2), 0.25f * mb);
#endif
    A = SA::fma(yb, PM, yk);   // yk + conj(yb)
    B = SA::fma(yb, MP, yk);   // yk - conj(yb)
}
// i conj(t) * b for a twiddle t: the factor is swap(t.w), its rotated copy -swap(t.iw) -- both operand modifiers
RCG_HD cf mul_i_conj(cf b, const Tw& t) {
    using SA = Ar<(RCG_PK_SPEC != 0)>;
    return SA::fma(pk_bcast(b.y), pk_neg(pk_swap(t.iw)), SA::mul(pk_bcast(b.x), pk_swap(t.w)));
}

// ---- SPECTRAL, half-rate synthesis (even pitch multiples) --------------------------------------
// Only y[2n] is ever emitted, and y[2n] = IFFT_H(E)[n] with E[k] = Y[k] + conj(Y[H-k]) (= the A
// of pair_fold), itself Hermitian.  So the inverse is a real transform of length H done as a
// complex one of length Q = H/2: Z''[k] = (E[k] + conj(E[Q-k])) + i e^{2 pi i k/H} (E[k] - conj(E[Q-k])).
// One thread owns the quad {k, H-k, Q-k, Q+k} (two pairs, two Philox blocks) and leaves Z''[k],
// Z''[Q-k] in the slots of bins k and Q-k: the slots of the full layout whose last digit is < 8,
// which is exactly the layout the half-rate inverse passes walk (make_plan(.., hr)).
template <int LOG2H>
RCG_HD void phase_spectral_hr(cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    constexpr int Q = C::H / 2;
    const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
    const uint32_t chan = (uint32_t)(channel + prm.channel_offset);
    const cf PM = make_cf(1.0f, -1.0f), MP = make_cf(-1.0f, 1.0f);
    const unsigned long long frame_ms = frame_clock(prm, frame);
    (void)frame_ms;
    for (int c = t; c < C::H / 4; c += C::NT) {
        // table entries of this iteration from the Tensor Memory row: the load is warp-collective (.sync.aligned), so it
        // sits ahead of the c == 0 branch that thread 0 takes alone
        uint4 qt;
        cf qtw;
        if constexpr (spectral_in_tmem(LOG2H)) {   // {qt.x, qt.y, qt.z, tw.x, tw.y, -, -, -} of iteration c / NT
            uint32_t r[8];
            row_fetch<LOG2H, 8>(tw_src(prm), c % C::NT, tmem_spectral_offset(LOG2H) + 8 * (c / C::NT), r);
            qt.x = r[0]; qt.y = r[1]; qt.z = r[2]; qt.w = 0u;
            qtw = make_cf(bits_to_float(r[3]), bits_to_float(r[4]));
        }
        uint32_t u[4];
        if (c == 0) {
            // E[0] = Y[0] + Y[H] (both real), E[Q] = 2 Re Y[Q]
            cf z0 = S[0];
            philox4x32_10(0u, f_lo, f_hi, chan, prm.keys, u);
#if defined(RCG_PLUGIN_LOCAL)
            float m0 = plugin_mag(prm, 0u, (uint32_t)C::W, z0.x + z0.y, 0.0f), mh = plugin_mag(prm, (uint32_t)C::H, (uint32_t)C::W, z0.x - z0.y, 0.0f);
#else
            float m0 = fabsf(z0.x + z0.y), mh = fabsf(z0.x - z0.y);
#endif
            float s, c0, ch;
            sincos_phase(phase_from_u32(u[0]), s, c0);
            sincos_phase(phase_from_u32(u[2]), s, ch);
            const float e0 = m0 * c0 + mh * ch;
            cf zq = S[phys(bin_to_pos<LOG2H>(Q))];
            philox4x32_10((uint32_t)Q, f_lo, f_hi, chan, prm.keys, u);
            float ca, cb;
            sincos_phase(phase_from_u32(u[0]), s, ca);
            sincos_phase(phase_from_u32(u[1]), s, cb);
#if defined(RCG_PLUGIN_LOCAL)   // X[Q] = conj(Z[Q]), X[W-Q] = Z[Q]
            const float eq = plugin_mag(prm, (uint32_t)Q, (uint32_t)C::W, zq.x, -zq.y) * ca +
                             plugin_mag(prm, (uint32_t)(C::W - Q), (uint32_t)C::W, zq.x, zq.y) * cb;
#else
            float mq = fast_sqrt(zq.x * zq.x + zq.y * zq.y);
            const float eq = mq * (ca + cb);
#endif
            S[0] = make_cf(e0 + eq, e0 - eq);
            // k = Q/2 pairs with itself: Z''[Q/2] = 2 conj(E[Q/2]); its twiddle is exp(-i pi/4)
            const int ph = phys(bin_to_pos<LOG2H>(Q / 2));
            const cf za = S[ph], zb = S[phys(bin_to_pos<LOG2H>(C::H - Q / 2))];
            philox4x32_10((uint32_t)(Q / 2), f_lo, f_hi, chan, prm.keys, u);
            cf A, B;
            pair_fold(za, zb, make_tw<(RCG_PK_SPEC != 0)>(make_cf(0.70710678118654752f, -0.70710678118654752f)), u, A, B, prm, (uint32_t)(Q / 2), (uint32_t)C::W, frame_ms);
            S[ph] = make_cf(2.0f * A.x, -2.0f * A.y);
            continue;
        }
        // slot indices of bins k, H-k, Q-k, Q+k and k itself: frame-invariant, precomputed on the host
        if constexpr (!spectral_in_tmem(LOG2H)) {
            qt = ld_u4(prm.quad_tab + c);
            qtw = ld_cf(prm.spec_tw_hr + c);
        }
        const int pa = (int)(qt.x & 0xFFFFu), pc = (int)(qt.y & 0xFFFFu);
        const cf za = S[pa], zb = S[qt.x >> 16];
        const cf zc = S[pc], zd = S[qt.y >> 16];
        const Tw tw = make_tw<(RCG_PK_SPEC != 0)>(qtw);  // exp(-2 pi i k / W)
        // exp(-2 pi i (Q-k)/W) = -i conj(tw) = -swap(tw); its rotated copy is swap(i tw): no instruction either
        Tw tw_c;
        tw_c.w = pk_neg(pk_swap(tw.w));
        tw_c.iw = pk_swap(tw.iw);
        cf E, F, unused;
        const uint32_t k = qt.z;
        philox4x32_10(k, f_lo, f_hi, chan, prm.keys, u);
        pair_fold(za, zb, tw, u, E, unused, prm, k, (uint32_t)C::W, frame_ms);                           // E[k]
        philox4x32_10((uint32_t)Q - k, f_lo, f_hi, chan, prm.keys, u);
        pair_fold(zc, zd, tw_c, u, F, unused, prm, (uint32_t)Q - k, (uint32_t)C::W, frame_ms);           // E[Q-k]
        // A2 = E[k] + conj(E[Q-k]), B2 = E[k] - conj(E[Q-k]), C2 = i conj(t2) B2, t2 = exp(-2 pi i k/H) = tw^2
        using SA = Ar<(RCG_PK_SPEC != 0)>;
        const cf a2 = SA::fma(F, PM, E), b2 = SA::fma(F, MP, E);
        const cf c2 = mul_i_conj(b2, make_tw<(RCG_PK_SPEC != 0)>(tw_mul<false, (RCG_PK_SPEC != 0)>(tw.w, tw)));
        S[pa] = SA::add(a2, c2);                  // Z''[k]
        S[pc] = SA::mul(SA::sub(a2, c2), PM);     // Z''[Q-k] = conj(A2 - C2)
    }
}

// ---- SPECTRAL: real-FFT split -> |X[k]| -> random phases for all W bins -> Hermitian
// fold -> pre-pass of the real inverse (fft.rs:63-69).  One thread owns the bin pair
// (k, H-k) of the half-size transform, i.e. bins {k, W-k, H-k, H+k} of the W-point one.
// KMUL: the global bin index of local bin j is j*KMUL (2 for the even-bin half of a CTA pair).
template <int LOG2H, int KMUL = 1>
RCG_HD void phase_spectral(cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    static_assert(C::RLAST == 16, "the spectral slot order assumes a radix-16 last pass");
    const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
    const unsigned long long frame_ms = frame_clock(prm, frame);
    (void)frame_ms;
    for (int c = t; c < C::H / 2; c += C::NT) {
        uint2 pt;
        cf ptw;
        if constexpr (spectral_in_tmem(LOG2H) && KMUL == 1) {   // {pt.x, pt.y, tw.x, tw.y} of iteration c / NT; ahead of the branch
            uint32_t r[4];
            row_fetch<LOG2H, 4>(tw_src(prm), c % C::NT, tmem_spectral_offset(LOG2H) + 4 * (c / C::NT), r);
            pt.x = r[0]; pt.y = r[1];
            ptw = make_cf(bits_to_float(r[2]), bits_to_float(r[3]));
        }
        uint32_t u[4];
        if (c == 0) {
            // k = 0: bins 0 and H are real: X[0] = a + b, X[H] = a - b
            cf z0 = S[0];
            philox4x32_10(0u, f_lo, f_hi, (uint32_t)(channel + prm.channel_offset), prm.keys, u);
#if defined(RCG_PLUGIN_LOCAL)
            float m0 = plugin_mag(prm, 0u, (uint32_t)(C::W * KMUL), z0.x + z0.y, 0.0f);
            float mh = plugin_mag(prm, (uint32_t)(C::H * KMUL), (uint32_t)(C::W * KMUL), z0.x - z0.y, 0.0f);
#else
            float m0 = fabsf(z0.x + z0.y), mh = fabsf(z0.x - z0.y);
#endif
            float s, c0, ch;
            sincos_phase(phase_from_u32(u[0]), s, c0);
            sincos_phase(phase_from_u32(u[2]), s, ch);
            float y0 = m0 * c0, yh = mh * ch;  // Re(B_0), Re(B_H)
            S[0] = make_cf(y0 + yh, y0 - yh);
            // k = H/2 pairs with itself: X[H/2] = conj(Z[H/2]); Z'[H/2] = 2 conj(Y[H/2])
            const int pq = phys(bin_to_pos<LOG2H>(C::H / 2));
            cf zq = S[pq];
            philox4x32_10((uint32_t)(C::H / 2 * KMUL), f_lo, f_hi, (uint32_t)(channel + prm.channel_offset), prm.keys, u);
            float sa, ca, sb, cb;
            sincos_phase(phase_from_u32(u[0]), sa, ca);
            sincos_phase(phase_from_u32(u[1]), sb, cb);
#if defined(RCG_PLUGIN_LOCAL)   // X[H/2] = conj(Z[H/2]), X[W - H/2] = Z[H/2]
            const float ma = plugin_mag(prm, (uint32_t)(C::H / 2 * KMUL), (uint32_t)(C::W * KMUL), zq.x, -zq.y);
            const float mb = plugin_mag(prm, (uint32_t)((C::W - C::H / 2) * KMUL), (uint32_t)(C::W * KMUL), zq.x, zq.y);
            S[pq] = make_cf(ma * ca + mb * cb, -(ma * sa - mb * sb));
#else
            float mq = fast_sqrt(zq.x * zq.x + zq.y * zq.y);
            S[pq] = make_cf(mq * (ca + cb), -(mq * (sa - sb)));
#endif
            continue;
        }
        if constexpr (!(spectral_in_tmem(LOG2H) && KMUL == 1)) {
            pt = ld_u2(prm.pair_tab + c);   // frame-invariant slot indices of bins k, H-k and k itself
            ptw = ld_cf(prm.spec_tw + c);
        }
        const int pa = (int)(pt.x & 0xFFFFu), pb = (int)(pt.x >> 16);
        const cf za = S[pa], zb = S[pb];
        const Tw tw = make_tw<(RCG_PK_SPEC != 0)>(ptw);  // exp(-2 pi i k / W)
        philox4x32_10(pt.y * (uint32_t)KMUL, f_lo, f_hi, (uint32_t)(channel + prm.channel_offset), prm.keys, u);
        // A = Y[k] + conj(Y[H-k]), B = Y[k] - conj(Y[H-k]), Cc = i conj(t) B
        cf A, B;
        pair_fold(za, zb, tw, u, A, B, prm, pt.y * (uint32_t)KMUL, (uint32_t)(C::W * KMUL), frame_ms);
        const cf Cc = mul_i_conj(B, tw);
        using SA = Ar<(RCG_PK_SPEC != 0)>;
        S[pa] = SA::add(A, Cc);                              // Z'[k]   = A + Cc
        S[pb] = SA::mul(SA::sub(A, Cc), make_cf(1.0f, -1.0f));  // Z'[H-k] = conj(A - Cc)
    }
}

// ---- plugin path, analysis side: real-FFT split -> the full W-bin spectrum the reference hands
// to `apply` (fft.rs:87: every bin of the complex FFT of the real windowed frame).
// Two phases so that HBM sees coalesced, natural-order writes although the transform leaves the
// bins digit-reversed: (1) pair threads replace Z[k], Z[H-k] by X[k], X[H-k] in place (position
// order, conflict free); (2) after a barrier the spectrum is copied out in natural bin order
// (the digit-reversing gather is paid in shared-memory bank conflicts, not in DRAM sectors).
template <int LOG2H>
RCG_HD void phase_spectrum_split(cf* S, int t, const StretchParams& prm) {
    using C = FrameCfg<LOG2H>;
    for (int c = t; c < C::H / 2; c += C::NT) {
        const int pos = spectral_slot_pos(LOG2H, c);
        const int k = pos_to_bin<LOG2H>(pos);
        if (c == 0) {
            const cf z0 = S[0];
            S[0] = make_cf(z0.x + z0.y, z0.x - z0.y);                 // (X[0], X[H]): both real
            const int pq = phys(bin_to_pos<LOG2H>(C::H / 2));
            const cf zq = S[pq];
            S[pq] = make_cf(zq.x, -zq.y);                             // X[H/2] = conj(Z[H/2])
            continue;
        }
        const int pa = phys(pos), pb = phys(bin_to_pos<LOG2H>(C::H - k));
        const cf za = S[pa], zb = S[pb];
        const cf tw = ld_cf(prm.spec_tw + c);
        float pr = za.x + zb.x, pi = za.y - zb.y, qr = za.x - zb.x, qi = za.y + zb.y;
        float tqr = tw.x * qr - tw.y * qi, tqi = tw.x * qi + tw.y * qr;
        float ur = tqi, ui = -tqr;
        S[pa] = make_cf(0.5f * (pr + ur), 0.5f * (pi + ui));          // X[k]
        S[pb] = make_cf(0.5f * (pr - ur), -(0.5f * (pi - ui)));       // X[H-k]
    }
}
// Which bin the i-th element of a natural-order sweep over the H bin slots stands for.  The transform
// leaves bin k at a position whose low 4 bits are k's top 4 bits, so 32 consecutive bins share one
// shared-memory bank pair (a 32-way conflict).  Instead a warp covers 8 values of k's top bits times 4
// consecutive bins: 8 distinct bank pairs (4 wavefronts instead of 32) while each group of 4 lanes still
// moves one whole 32-byte sector of the spectrum in HBM.
template <int LOG2H>
RCG_HD constexpr int sweep_bin(int i) {
    using C = FrameCfg<LOG2H>;
    if constexpr (LOG2H < 9 || C::NT < 32) {
        return i;
    } else {
        constexpr int HS = LOG2H - 4;   // k >> HS = the digit of the last pass = position bits 0..3
        const int r = i >> 5;
        return (i & 3) | ((r & ((1 << (HS - 2)) - 1)) << 2) | (((i >> 2) & 7) << HS) | ((r >> (HS - 2)) << (LOG2H - 1));
    }
}
template <int LOG2H>
RCG_HD void phase_spectrum_copyout(const cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    cf* X = prm.spectrum + ((long long)channel * prm.spec_frames + (frame - prm.spec_frame0)) * C::W;
RCG_UNROLL(4)
    for (int i = t; i < C::H; i += C::NT) {
        const int b = sweep_bin<LOG2H>(i);
        const cf v = S[phys(bin_to_pos<LOG2H>(b))];
        // bin 0 holds (X[0], X[H]), both real; otherwise the spectrum of a real frame is Hermitian
        X[b] = make_cf(v.x, b == 0 ? 0.0f : v.y);
        X[b == 0 ? C::H : C::W - b] = make_cf(b == 0 ? v.y : v.x, b == 0 ? 0.0f : -v.y);
    }
}

RCG_HD float cf_norm(cf v) { return fast_sqrt(v.x * v.x + v.y * v.y); }

// ---- plugin path, synthesis side: |apply(X)[k]| for each of the W bins (only the norm of the
// plugin's output survives, fft.rs:67) -> random phase -> Hermitian fold -> inverse pre-pass.
// Again two phases: (1) natural-order, coalesced read of the plugin output; the slot of bin b < H
// receives (|P[b]|, |P[W-b]|) (slot 0: (|P[0]|, |P[H]|)); (2) pair threads in position order.
template <int LOG2H>
RCG_HD void phase_mags_copyin(cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    const cf* X = prm.spectrum + ((long long)channel * prm.spec_frames + (frame - prm.spec_frame0)) * C::W;
    // all loads of a batch are issued before the first use: the sweep is bound by HBM latency otherwise
    constexpr int ITERS = C::H / C::NT, U = ITERS % 8 == 0 ? 8 : (ITERS % 4 == 0 ? 4 : 1);
    static_assert(C::H % C::NT == 0, "bins per thread");
    for (int i0 = t; i0 < C::H; i0 += C::NT * U) {
        cf lo[U], hi[U];
RCG_UNROLL(U)
        for (int u = 0; u < U; u++) {
            const int b = sweep_bin<LOG2H>(i0 + u * C::NT);
            lo[u] = ld_once(X + b);
            hi[u] = ld_once(X + (b == 0 ? C::H : C::W - b));
        }
RCG_UNROLL(U)
        for (int u = 0; u < U; u++) {
            const int b = sweep_bin<LOG2H>(i0 + u * C::NT);
            S[phys(bin_to_pos<LOG2H>(b))] = make_cf(cf_norm(lo[u]), cf_norm(hi[u]));
        }
    }
}
// KMUL: as in phase_spectral (2 for the even-bin half of a CTA pair)
template <int LOG2H, int KMUL = 1>
RCG_HD void phase_spectral_in(cf* S, int t, const StretchParams& prm, int channel, long long frame) {
    using C = FrameCfg<LOG2H>;
    const uint32_t f_lo = (uint32_t)(unsigned long long)frame, f_hi = (uint32_t)((unsigned long long)frame >> 32);
    const uint32_t chan = (uint32_t)(channel + prm.channel_offset);
    for (int c = t; c < C::H / 2; c += C::NT) {
        uint32_t u[4];
        if (c == 0) {
            const cf m0h = S[0];                                     // (|P[0]|, |P[H]|)
            philox4x32_10(0u, f_lo, f_hi, chan, prm.keys, u);
            float s, c0, ch;
            sincos_phase(phase_from_u32(u[0]), s, c0);
            sincos_phase(phase_from_u32(u[2]), s, ch);
            float y0 = m0h.x * c0, yh = m0h.y * ch;
            S[0] = make_cf(y0 + yh, y0 - yh);
            const int pq = phys(bin_to_pos<LOG2H>(C::H / 2));
            const cf mq = S[pq];                                     // (|P[H/2]|, |P[3H/2]|)
            philox4x32_10((uint32_t)(C::H / 2 * KMUL), f_lo, f_hi, chan, prm.keys, u);
            float sa, ca, sb, cb;
            sincos_phase(phase_from_u32(u[0]), sa, ca);
            sincos_phase(phase_from_u32(u[1]), sb, cb);
            S[pq] = make_cf(mq.x * ca + mq.y * cb, -(mq.x * sa - mq.y * sb));
            continue;
        }
        const uint2 pt = ld_u2(prm.pair_tab + c);   // slot indices of bins k, H-k and k itself (as in phase_spectral)
        const int pa = (int)(pt.x & 0xFFFFu), pb = (int)(pt.x >> 16), k = (int)pt.y;
        const cf ma = S[pa], mb = S[pb];                             // (|P[k]|, |P[W-k]|), (|P[H-k]|, |P[H+k]|)
        const cf tw = ld_cf(prm.spec_tw + c);
        philox4x32_10((uint32_t)(k * KMUL), f_lo, f_hi, chan, prm.keys, u);
        const float m0 = ma.x, m1 = ma.y, m2 = mb.x, m3 = mb.y;
        float s0, c0, s1, c1, s2, c2, s3, c3;
        sincos_phase(phase_from_u32(u[0]), s0, c0);
        sincos_phase(phase_from_u32(u[1]), s1, c1);
        sincos_phase(phase_from_u32(u[2]), s2, c2);
        sincos_phase(phase_from_u32(u[3]), s3, c3);
        float ykr = 0.5f * (m0 * c0 + m1 * c1), yki = 0.5f * (m0 * s0 - m1 * s1);
        float ybr = 0.5f * (m2 * c2 + m3 * c3), ybi = 0.5f * (m2 * s2 - m3 * s3);
        float ar = ykr + ybr, ai = yki - ybi, br = ykr - ybr, bi = yki + ybi;
        float cr = -(tw.x * bi - tw.y * br), ci = tw.x * br + tw.y * bi;
        S[pa] = make_cf(ar + cr, ai + ci);
        S[pb] = make_cf(ar - cr, -(ai - ci));
    }
}

// value i (0..W) of the frame held as packed pairs: component i & 1 of slot phys(i >> 1), i.e. float number
// 2 ((i >> 1) + (i >> 5)) + (i & 1) = i + 2 (i >> 5) of the buffer -- one scalar load instead of a 64-bit load, a parity
// test and a select (the lerp walk of a negative pitch multiple reads two of them per output sample)
RCG_HD float frame_value(const cf* S, int i) {
    return reinterpret_cast<const float*>(S)[i + ((i >> 5) << 1)];
}
// the same value addressed in bytes from a pointer that already stands at float i0: value i0 + c with c a constant
// of the unrolled loop costs a shift, one shift-add and a load with an immediate offset
RCG_HD float frame_value_at(const char* at_i0, int i0, int c) {
    return *reinterpret_cast<const float*>(at_i0 + 4 * c + (((i0 + c) >> 5) << 3));
}

// ---- inverse pass 0 fused with EMIT part 1: last DIT pass out of shared memory, then scale +
// synthesis window (fft.rs:70-73), overlap-add with the carried half window, envelope,
// amplitude (stretcher.rs:96-102).  Arithmetic order is the reference's: ((re / W) * w), then
// ((y + tail) * env) * amp.  A thread owns packed outputs n = m + a*M, a < R: those with a < R/2
// are the first half window (emitted), the others the second half (carried to the next frame);
// both map onto the same slots of the carried buffer, so no other thread ever touches them.
template <int LOG2H, bool BYPASS>
RCG_HD void phase_ola_impl(cf* S, int t, const StretchParams& prm, int channel, long long frame,
                           float* tail, bool emit, bool write_head) {
    using C = FrameCfg<LOG2H>;
    constexpr int RL = make_plan(LOG2H).rlog2[0], R = 1 << RL, ML = make_plan(LOG2H).mlog2[0], M = 1 << ML;
    constexpr int RH = R / 2;
    constexpr bool bypass = BYPASS;
    constexpr bool IO_PK = (RCG_PK_IO != 0);
    float* out = prm.out + (long long)channel * prm.out_stride;
    const TwSrc tw = tw_src(prm);
    // one iteration: packed index m; TM: window / envelope values of this iteration handed in (Tensor Memory row)
    auto iteration = [&](int m, bool tm, const cf* wtm, cf etm) {
        cf* p = S + phys(m);
        cf x[R];
#pragma unroll
        for (int a = 0; a < R; a++) x[a] = p[slot_off<M>(a)];
        if constexpr (!bypass) {
            twiddle_butterfly<LOG2H, 0, true, R, false, C::TW_TMEM>(x, tw, m);
            Dft<R, true>::run(x);
        }
        // the reference's operation order: (re / W) * w, then ((y + tail) * env) * amp; the multiply and the add
        // around y are scalar so that nothing can fuse them (see ex_mul), the rest may be packed
#pragma unroll
        for (int a = 0; a < RH; a++) {
            const int n = m + a * M;              // packed index inside the first half window
            const cf wa = tm ? wtm[0] : ld_stream2(prm.window + 2 * n), wb = tm ? wtm[1] : ld_stream2(prm.window + C::H + 2 * n);
            cf y, tl;
            if constexpr (bypass) {
                y = ex_mul(x[a], wa); tl = ex_mul(x[a + RH], wb);
            } else {
                y = ex_mul(mul_chain<IO_PK>(x[a], pk_bcast(prm.inv_w)), wa);
                tl = ex_mul(mul_chain<IO_PK>(x[a + RH], pk_bcast(prm.inv_w)), wb);
            }
            const cf old = ld_pair(tail + 2 * n);
            st_cf(tail + 2 * n, tl);
            const cf e = tm ? etm : ld_stream2(prm.env + 2 * n);   // unconditional: keeps the loop body free of branches around loads
            if (write_head) st_cf(prm.head_out + (long long)channel * C::H + 2 * n, y);
            if (emit) {
                const cf z = mul_chain<IO_PK>(mul_chain<IO_PK>(ex_add(y, old), e), pk_bcast(prm.amp));
                if (prm.pitch == 1) {
                    st_cf(out + (frame * (long long)C::H + 2 * n - prm.out_origin), z);
                } else {
                    p[slot_off<M>(a)] = z;  // staged at packed index n for the resampling gather
                }
            }
        }
    };
    if constexpr (window_in_tmem(LOG2H) && !bypass) {
        if (prm.flags & kFlagWinTmem) {
            // window (8 columns) and envelope (4 columns) of two iterations per pair of Tensor Memory loads
            static_assert(R == 2 && tmem_window_iters(LOG2H) % 2 == 0, "4 window + 2 envelope columns per iteration");
            for (int g = 0; g < tmem_window_iters(LOG2H) / 2; g++) {
                uint32_t wv[8], ev[4];
                row_fetch<LOG2H, 8>(tw, t, tmem_window_offset(LOG2H) + 8 * g, wv);
                row_fetch<LOG2H, 4>(tw, t, tmem_env_col(LOG2H, 2 * g), ev);
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const cf w2[2] = {make_cf(bits_to_float(wv[4 * u]), bits_to_float(wv[4 * u + 1])),
                                      make_cf(bits_to_float(wv[4 * u + 2]), bits_to_float(wv[4 * u + 3]))};
                    iteration(t + C::NT * (2 * g + u), true, w2, make_cf(bits_to_float(ev[2 * u]), bits_to_float(ev[2 * u + 1])));
                }
            }
            return;
        }
    }
RCG_UNROLL((LOG2H >= 14 ? RCG_OLA_UNROLL_512 : RCG_OLA_UNROLL))
    for (int m = t; m < M; m += C::NT) iteration(m, false, nullptr, make_cf(0.0f, 0.0f));
}
template <int LOG2H>
RCG_HD void phase_ola(cf* S, int t, const StretchParams& prm, int channel, long long frame,
                      float* tail, bool emit, bool write_head) {
    if (prm.flags & kFlagBypass) phase_ola_impl<LOG2H, true>(S, t, prm, channel, frame, tail, emit, write_head);
    else phase_ola_impl<LOG2H, false>(S, t, prm, channel, frame, tail, emit, write_head);
}

// ---- half-rate synthesis: inverse DIT butterfly of pass PASS >= 1 of the Q-point transform ------
// Logical indices of the hr plan, physical slots of the full layout (expand_hr).  For the long-
// stride pass the butterflies are dealt to lanes with bits 3 and 6 of m swapped, so that the two
// 8-slot rows a 16-lane half warp touches are 8 rows apart (disjoint bank pairs).
template <int LOG2H, int PASS>
RCG_HD void fft_butterfly_hr(cf* S, const TwSrc& tw, int b) {
    constexpr PlanDesc P = make_plan(LOG2H, true);
    constexpr int RL = P.rlog2[PASS], R = 1 << RL, ML = P.mlog2[PASS], M = 1 << ML;
    constexpr int STRIDE = (M == 1) ? 1 : 2 * M;   // physical stride between butterfly inputs
    static_assert(M == 1 || M >= 8, "hr pass strides are 1 or multiples of 8");
    constexpr bool TM = FrameCfg<LOG2H>::TW_TMEM_HR;
    int m = b & (M - 1);
    const int blk = b >> ML;
    if constexpr (M >= 128) m = hr_lane_swizzle(m);
    const TwPre pre = tw_preload<LOG2H, PASS, R, true, TM>(tw.tw + tw_offset(LOG2H, PASS, true), m);
    cf* p = S + phys(expand_hr((blk << (ML + RL)) + m));
    cf x[R];
#pragma unroll
    for (int a = 0; a < R; a++) x[a] = p[slot_off<STRIDE>(a)];
    twiddle_butterfly<LOG2H, PASS, true, R, true, TM>(x, tw, m, pre);
    Dft<R, true>::run(x);
#pragma unroll
    for (int j = 0; j < R; j++) p[slot_off<STRIDE>(j)] = x[j];
}

template <int LOG2H, int PASS>
RCG_HD void phase_fft_pass_hr(cf* S, int t, const TwSrc& tw_all) {
    using C = FrameCfg<LOG2H>;
    constexpr int NB = (C::H / 2) >> make_plan(LOG2H, true).rlog2[PASS];
    for (int b = t; b < NB; b += C::NT) fft_butterfly_hr<LOG2H, PASS>(S, tw_all, b);
}

// hr tail: the radix-8 last pass works on rows of 8 slots; the radix-16 pass before it on blocks
// of 16 rows.  A warp's 32 radix-16 butterflies (index 32w'..32w'+31) cover rows 64w'..64w'+63,
// so step 2i = those 64 radix-8 butterflies (two per lane), step 2i+1 = the radix-16 ones.
template <int LOG2H>
RCG_HD void fft_tail_step_hr(cf* S, int t, int step, const TwSrc& tw_all) {
    using C = FrameCfg<LOG2H>;
    constexpr int NB_PREV = (C::H / 2) / 16, NB_LAST = (C::H / 2) / 8;
    const int b = t + C::NT * (step >> 1);
    if ((step & 1) == 0) {
        const int b1 = 64 * (b >> 5) + (b & 31);
        if (b1 < NB_LAST) fft_butterfly_hr<LOG2H, C::NP - 1>(S, tw_all, b1);
        if (b1 + 32 < NB_LAST) fft_butterfly_hr<LOG2H, C::NP - 1>(S, tw_all, b1 + 32);
    } else if (b < NB_PREV) {
        fft_butterfly_hr<LOG2H, C::NP - 2>(S, tw_all, b);
    }
}
template <int LOG2H>
constexpr int fft_tail_steps_hr() {
    return 2 * ((((FrameCfg<LOG2H>::H / 2) / 16) + FrameCfg<LOG2H>::NT - 1) / FrameCfg<LOG2H>::NT);
}

// value i of the staged half-rate block (packed pairs at the hr slots)
RCG_HD float frame_value_hr(const cf* S, int i) {
    cf v = S[phys(expand_hr(i >> 1))];
    return (i & 1) ? v.y : v.x;
}

// ---- half-rate synthesis: last DIT pass + /W * window + overlap-add + envelope * amp on the even
// samples only.  ye[i] = y[2i], i < H: i < Q is the emitted half window, the rest is carried.
// The carried buffer, head_out / tail_in / tail_out then hold even samples ([Q] per channel).
template <int LOG2H>
RCG_HD void phase_ola_hr(cf* S, int t, const StretchParams& prm, int channel, long long frame,
                         float* tail, bool emit, bool write_head) {
    using C = FrameCfg<LOG2H>;
    constexpr PlanDesc P = make_plan(LOG2H, true);
    constexpr int RL = P.rlog2[0], R = 1 << RL, ML = P.mlog2[0], M = 1 << ML;
    constexpr int RH = R / 2, Q = C::H / 2;
    constexpr int STRIDE = (M == 1) ? 1 : 2 * M;
    constexpr bool IO_PK = (RCG_PK_IO != 0);
    float* out = prm.out + (long long)channel * prm.out_stride;
    if (prm.flags & kFlagBypass) {
        // test hook: y[n] = (s[n] w[n]) w[n]; the load phase left s[2n] w[2n] in S[phys(n)].x
        for (int i = t; i < Q; i += C::NT) {
            const float y = RCG_MUL(S[phys(i)].x, prm.window_even[i]);
            const float tl = RCG_MUL(S[phys(i + Q)].x, prm.window_even[i + Q]);
            const float old = tail[i];
            tail[i] = tl;
            if (write_head) prm.head_out[(long long)channel * C::H + i] = y;
            if (emit) {
                const float z = RCG_MUL(RCG_MUL(RCG_ADD(y, old), prm.env_even[i]), prm.amp);
                if (prm.pitch == 2) out[frame * (long long)Q + i - prm.out_origin] = z;
                else S[phys(i)].x = z;  // staged in the slot only this thread reads
            }
        }
        return;
    }
    const TwSrc tw = tw_src(prm);
RCG_UNROLL((LOG2H >= 14 ? RCG_OLA_UNROLL_512 : RCG_OLA_UNROLL))
    for (int m = t; m < M; m += C::NT) {
        cf* p = S + phys(expand_hr(m));
        cf x[R];
#pragma unroll
        for (int a = 0; a < R; a++) x[a] = p[slot_off<STRIDE>(a)];
        twiddle_butterfly<LOG2H, 0, true, R, true, C::TW_TMEM_HR>(x, tw, m);
        Dft<R, true>::run(x);
#pragma unroll
        for (int a = 0; a < RH; a++) {
            const int n = m + a * M;   // packed index: even samples 2n, 2n+1 of the emitted half
            const cf wa = ld_stream2(prm.window_even + 2 * n), wb = ld_stream2(prm.window_even + Q + 2 * n);
            const cf y = ex_mul(mul_chain<IO_PK>(x[a], pk_bcast(prm.inv_w)), wa);
            const cf tl = ex_mul(mul_chain<IO_PK>(x[a + RH], pk_bcast(prm.inv_w)), wb);
            const cf old = ld_pair(tail + 2 * n);
            st_cf(tail + 2 * n, tl);
            const cf e = ld_stream2(prm.env_even + 2 * n);  // unconditional: no branches around loads
            if (write_head) st_cf(prm.head_out + (long long)channel * C::H + 2 * n, y);
            if (emit) {
                const cf z = mul_chain<IO_PK>(mul_chain<IO_PK>(ex_add(y, old), e), pk_bcast(prm.amp));
                if (prm.pitch == 2) {
                    st_cf(out + (frame * (long long)Q + 2 * n - prm.out_origin), z);
                } else {
                    p[slot_off<STRIDE>(a)] = z;  // staged for the decimating gather
                }
            }
        }
    }
}

// ---- EMIT part 2: integer resample of the finished half window held in S[0..H/2)
// (resampler.rs:3-35).  p > 1: out[g / p] = z[g] for g % p == 0 (g = absolute index into the
// overlap-added stream; calls tile it exactly since S = W*p).  p < -1: one frame per call;
// the first S samples are linearly interpolated |p| times, S-1 intervals (math.rs:28-30).
template <int LOG2H>
RCG_HD void phase_resample(const cf* S, int t, const StretchParams& prm, int channel, long long frame,
                           bool plain_stage = false) {
    using C = FrameCfg<LOG2H>;
    float* out = prm.out + (long long)channel * prm.out_stride;
    const int p = prm.pitch;
    if (p > 1 && (prm.flags & kFlagHalfRate)) {
        // staged block = even samples ze[i] = z[2i], i < Q; out[g/p] = z[g] <=> out[ge/(p/2)] = ze[ge]
        constexpr int Q = C::H / 2;
        const int pe = p / 2;
        const bool bypass = plain_stage || (prm.flags & kFlagBypass);  // staged as S[phys(i)].x
        const long long g0 = frame * (long long)Q;
        const long long o_begin = (g0 + pe - 1) / pe, o_end = (g0 + Q + pe - 1) / pe;
        // 32-bit index walk from the block's first output (the 64-bit products are loop invariants)
        const int cnt = (int)(o_end - o_begin);
        float* dp = out + (o_begin - prm.out_origin);
        int i = (int)(o_begin * pe - g0) + t * pe;
        for (int k = t; k < cnt; k += C::NT, i += C::NT * pe) dp[k] = bypass ? S[phys(i)].x : frame_value_hr(S, i);
    } else if (p > 1) {
        const long long g0 = frame * (long long)C::H;
        const long long o_begin = (g0 + p - 1) / p, o_end = (g0 + C::H + p - 1) / p;
        const int cnt = (int)(o_end - o_begin);
        float* dp = out + (o_begin - prm.out_origin);
        int i = (int)(o_begin * p - g0) + t * p;
        for (int k = t; k < cnt; k += C::NT, i += C::NT * p) dp[k] = frame_value(S, i);
    } else if (p < -1) {
        // resample_slower (resampler.rs:20-35): out[i*q + j] = z[i] + (z[i+1] - z[i]) * (j / q), i < S - 1.
        // Consecutive threads write consecutive outputs (whole sectors per warp store; a thread per input interval
        // writing its q outputs was measured 32 % slower on config 4: partial-sector stores).  The threads used are
        // a multiple of q, so that a thread's j -- and its ratio j/q, from a host-computed table (same IEEE division
        // as the reference) -- never changes and i advances by a constant: no division, no branch in the loop.
        const int q = -p;
        const int n = (prm.samples_needed - 1) * q;
        float* dst = out + (frame * (long long)n - prm.out_origin);
        const int nt_q = (C::NT / q) * q;
        if (nt_q > 0) {
            if (t < nt_q) {
                // a running store pointer and a trip count known before the loop: unrolled, the stores and the index
                // updates use immediate offsets (the loop was 25 instructions per output sample, a fifth of all
                // instructions executed on config 4, with a 64-bit address rebuilt for every store)
                const int di = nt_q / q;
                int i = t / q;
                const float ratio = prm.lerp_ratio[t - i * q];
                float* dp = dst + t;
                const char* sp = reinterpret_cast<const char*>(S) + 4 * i;
                const int trips = (n - t + nt_q - 1) / nt_q;
RCG_UNROLL(4)
                for (int k = 0; k < trips; k++, i += di, sp += 4 * di, dp += nt_q) {
                    const float cur = frame_value_at(sp, i, 0), nxt = frame_value_at(sp, i, 1);
                    *dp = RCG_ADD(cur, RCG_MUL(RCG_SUB(nxt, cur), ratio));
                }
            }
        } else {   // more steps per interval than threads in the lane (tiny windows with a large |p|)
            for (int tt = t; tt < n; tt += C::NT) {
                const int i = tt / q, j = tt - i * q;
                const float cur = frame_value(S, i), nxt = frame_value(S, i + 1);
                dst[tt] = RCG_ADD(cur, RCG_MUL(RCG_SUB(nxt, cur), prm.lerp_ratio[j]));
            }
        }
    }
}

// Exec::par(f) runs f(t) for every thread of the frame lane followed by a block barrier;
// Exec::par_steps(n, f) runs f(step, t) for step = 0..n-1 with a warp barrier after every step
// (only intra-warp dependencies between steps) and a block barrier at the end.
template <int LOG2H, bool INV, int PASS, int LAST, class Exec>
RCG_HD void fft_range(Exec& ex, cf* S, const StretchParams& prm, bool work) {   // passes PASS..LAST one by one
    if constexpr (INV ? PASS >= LAST : PASS <= LAST) {
        ex.par([&](int t) { if (work) phase_fft_pass<LOG2H, PASS, INV>(S, t, tw_src(prm)); });
        fft_range<LOG2H, INV, INV ? PASS - 1 : PASS + 1, LAST>(ex, S, prm, work);
    }
}
template <int LOG2H, int PASS, int LAST, class Exec>
RCG_HD void fft_range_hr(Exec& ex, cf* S, const StretchParams& prm, bool work) {
    if constexpr (PASS >= LAST) {
        ex.par([&](int t) { if (work) phase_fft_pass_hr<LOG2H, PASS>(S, t, tw_src(prm)); });
        fft_range_hr<LOG2H, PASS - 1, LAST>(ex, S, prm, work);
    }
}

// forward passes 1..NP-1 / inverse passes NP-1..1 of the full plan
template <int LOG2H, bool INV, class Exec>
RCG_HD void fft_all(Exec& ex, cf* S, const StretchParams& prm, bool work) {
    using C = FrameCfg<LOG2H>;
    if constexpr (C::NP >= 3) {
        if constexpr (!INV) fft_range<LOG2H, false, 1, C::NP - 3>(ex, S, prm, work);
        ex.par_steps(fft_tail_steps<LOG2H>(), [&](int step, int t) { if (work) fft_tail_step<LOG2H, INV>(S, t, step, tw_src(prm)); });
        if constexpr (INV) fft_range<LOG2H, true, C::NP - 3, 1>(ex, S, prm, work);
    } else {
        fft_range<LOG2H, INV, INV ? C::NP - 1 : 1, INV ? 1 : C::NP - 1>(ex, S, prm, work);
    }
}
// inverse passes NP-1..1 of the half-rate plan
template <int LOG2H, class Exec>
RCG_HD void fft_inv_hr(Exec& ex, cf* S, const StretchParams& prm, bool work) {
    using C = FrameCfg<LOG2H>;
    if constexpr (C::NP >= 3) {
        ex.par_steps(fft_tail_steps_hr<LOG2H>(), [&](int step, int t) { if (work) fft_tail_step_hr<LOG2H>(S, t, step, tw_src(prm)); });
        fft_range_hr<LOG2H, C::NP - 3, 1>(ex, S, prm, work);
    } else {
        fft_range_hr<LOG2H, C::NP - 1, 1>(ex, S, prm, work);
    }
}

// ---- the frame loop of one work item (a run of consecutive frames of one channel) --------
// Exec::par(f) runs f(t) for every thread of the frame lane followed by a barrier.
// SPEC_IN: the frequency-kernel synthesis (kFlagSpectrumIn) as its own instantiation, so that its phases
// do not weigh on the register allocation of the fused kernel.
template <int LOG2H, bool SPEC_IN = false, class Exec>
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
    const bool half_rate = !SPEC_IN && (prm.flags & kFlagHalfRate);
    const bool bypass = prm.flags & kFlagBypass;
    constexpr bool spectrum_in = SPEC_IN;
    const int n_iter = prm.run_len + 1;  // uniform trip count for every lane of the CTA
    for (int it = 0; it < n_iter; it++) {
        const long long f = f0 - 1 + it;
        const bool is_halo = (it == 0);
        const bool work = active && (is_halo ? halo : f < f1);
        if constexpr (spectrum_in) {
            ex.par([&](int t) { if (work) phase_mags_copyin<LOG2H>(S, t, prm, channel, f); });
            ex.par([&](int t) { if (work) phase_spectral_in<LOG2H>(S, t, prm, channel, f); });
            fft_all<LOG2H, true>(ex, S, prm, work);
        } else {
            ex.par([&](int t) { if (work) phase_load<LOG2H>(S, t, prm, channel, f); });
            if (!bypass) {
                fft_all<LOG2H, false>(ex, S, prm, work);
                if (half_rate) {
                    ex.par([&](int t) { if (work) phase_spectral_hr<LOG2H>(S, t, prm, channel, f); });
                    fft_inv_hr<LOG2H>(ex, S, prm, work);
                } else {
                    ex.par([&](int t) { if (work) phase_spectral<LOG2H>(S, t, prm, channel, f); });
                    fft_all<LOG2H, true>(ex, S, prm, work);
                }
            }
        }
        const bool head = work && defer && f == f0;
        const bool emit = work && !is_halo && !head;
        if (half_rate) {
            ex.par([&](int t) { if (work) phase_ola_hr<LOG2H>(S, t, prm, channel, f, tail, emit, head); });
            if (prm.pitch != 2)
                ex.par([&](int t) { if (emit) phase_resample<LOG2H>(S, t, prm, channel, f); });
        } else {
            ex.par([&](int t) { if (work) phase_ola<LOG2H>(S, t, prm, channel, f, tail, emit, head); });
            if (prm.pitch != 1)
                ex.par([&](int t) { if (emit) phase_resample<LOG2H>(S, t, prm, channel, f); });
        }
    }
    if (prm.tail_out != nullptr) {
        ex.par([&](int t) {
            if (!active || f1 != prm.frame_end) return;
            float* dst = prm.tail_out + (long long)channel * C::H;
            for (int i = t; i < C::H; i += C::NT) dst[i] = tail[i];
        });
    }
}

// ---- plugin path: analysis of a run of frames (window, forward FFT, full spectrum to HBM) ----
template <int LOG2H, class Exec>
RCG_HD void run_analysis_item(Exec& ex, cf* S, const StretchParams& prm, int item, bool active) {
    const int channel = active ? item / prm.n_runs : 0;
    const int run = active ? item - channel * prm.n_runs : 0;
    const long long f0 = prm.frame_begin + (long long)run * prm.run_len;
    for (int it = 0; it < prm.run_len; it++) {
        const long long f = f0 + it;
        const bool work = active && f < prm.frame_end;
        ex.par([&](int t) { if (work) phase_load<LOG2H>(S, t, prm, channel, f); });
        fft_all<LOG2H, false>(ex, S, prm, work);
        ex.par([&](int t) { if (work) phase_spectrum_split<LOG2H>(S, t, prm); });
        ex.par([&](int t) { if (work) phase_spectrum_copyout<LOG2H>(S, t, prm, channel, f); });
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
    if (prm.flags & kFlagHalfRate) {  // head / tail hold even samples: ze[i] = ((head + tail) * env[2i]) * amp
        constexpr int Q = C::H / 2;
        ex.par([&](int t) {
            const float* h = head + (long long)channel * C::H;
            const float* tp = tail_prev + (long long)channel * C::H;
            for (int i = t; i < Q; i += C::NT) {
                const float z = RCG_MUL(RCG_MUL(RCG_ADD(h[i], tp[i]), prm.env_even[i]), prm.amp);
                if (prm.pitch == 2) out[frame * (long long)Q + i - prm.out_origin] = z;
                else S[phys(i)].x = z;
            }
        });
        if (prm.pitch != 2) ex.par([&](int t) { phase_resample<LOG2H>(S, t, prm, channel, frame, true); });
        return;
    }
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

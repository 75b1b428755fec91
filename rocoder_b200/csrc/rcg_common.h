// rcg_common.h -- types and the compile-time FFT plan shared by the sm_100a kernels,
// the host table builder and the host-side emulator used by the CPU tests.
#pragma once
#if defined(__CUDACC_RTC__)
// NVRTC (frequency kernels fused into the stretch kernel at run time, rcg_jit.*): no host headers
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
#else
#include <stdint.h>
#include <math.h>
#include <string.h>
#endif

#if defined(__CUDACC__)
#define RCG_HD __host__ __device__ __forceinline__
#define RCG_D __device__ __forceinline__
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>
#endif
typedef float2 cf;
#else
#define RCG_HD inline
#define RCG_D inline
struct alignas(8) cf { float x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
#endif

RCG_HD cf make_cf(float x, float y) { cf r; r.x = x; r.y = y; return r; }

namespace rcg {

// ---- FFT plan ------------------------------------------------------------------------
// A frame of W real samples is transformed as an H = W/2 point complex FFT of the
// even/odd-packed samples (real-FFT split in the spectral step).  The H-point transform
// is an in-place Cooley-Tukey over mixed power-of-two digits: forward = DIF passes
// 0..NP-1 (natural order in, digit-reversed order out), inverse = DIT passes NP-1..0
// (digit-reversed in, natural out), so no reordering pass ever runs.
// Radices: log2(H) = 4*n16 + r -> one leading pass of radix 2^r (if r != 0) followed by
// n16 radix-16 passes.
constexpr int kMaxPasses = 5;

struct PlanDesc {
    int log2h;              // log2 of the transform size this plan describes
    int hr;                 // half-rate synthesis plan (see make_plan)
    int npasses;
    int rlog2[kMaxPasses];  // log2 radix of pass p
    int mlog2[kMaxPasses];  // log2 of M_p = stride between butterfly inputs in pass p
    int klo[kMaxPasses];    // bit offset of digit p inside the bin index k
};

// hr = true describes the half-rate synthesis transform used for even pitch multiples: a
// 2^(log2h-1) point inverse with the same passes but the last radix halved (16 -> 8).  Its
// data live in the slots of the full layout whose last digit is < 8, so its pass strides in
// shared memory are those of the full plan (see expand_hr()), while mlog2/klo are the logical
// ones that index the twiddle tables.
RCG_HD constexpr PlanDesc make_plan(int log2h, bool hr = false) {
    PlanDesc d{};
    const int size_log2 = hr ? log2h - 1 : log2h;
    d.log2h = size_log2;
    d.hr = hr ? 1 : 0;
    int rl[kMaxPasses] = {0, 0, 0, 0, 0};
    int r = log2h % 4, n16 = log2h / 4, np = 0, used = 0;
    if (r != 0) rl[np++] = r;
    for (int i = 0; i < n16; i++) rl[np++] = 4;
    if (hr) rl[np - 1] -= 1;
    for (int p = 0; p < np; p++) { d.rlog2[p] = rl[p]; d.klo[p] = used; used += rl[p]; d.mlog2[p] = size_log2 - used; }
    d.npasses = np;
    return d;
}

// logical index of the half-rate transform -> slot index of the full layout (zero inserted at
// bit 3: the last digit only takes values 0..7 of its 16 slots)
RCG_HD constexpr int expand_hr(int x) { return ((x >> 3) << 4) | (x & 7); }

// position (in natural cf units, before padding) of bin k after the forward passes, and
// its inverse.  Fully unrolled: every shift is a compile-time constant.
template <int LOG2H, int PASS = 0>
RCG_HD constexpr int bin_to_pos(int k) {
    constexpr PlanDesc d = make_plan(LOG2H);
    if constexpr (PASS >= d.npasses) {
        return 0;
    } else {
        return (((k >> d.klo[PASS]) & ((1 << d.rlog2[PASS]) - 1)) << d.mlog2[PASS]) | bin_to_pos<LOG2H, PASS + 1>(k);
    }
}
template <int LOG2H, int PASS = 0>
RCG_HD constexpr int pos_to_bin(int pos) {
    constexpr PlanDesc d = make_plan(LOG2H);
    if constexpr (PASS >= d.npasses) {
        return 0;
    } else {
        return (((pos >> d.mlog2[PASS]) & ((1 << d.rlog2[PASS]) - 1)) << d.klo[PASS]) | pos_to_bin<LOG2H, PASS + 1>(pos);
    }
}

// ---- which bin slot a thread of the spectral step owns ----------------------------------------
// The spectral step walks slots in POSITION order (the forward transform leaves bins digit-reversed):
// slot c -> position pos = 16*row + kl with kl < 8 (pairs: bins k < H/2) or kl < 4 (half-rate quads:
// k < H/4).  Rows are 17 cf apart in the padded buffer, so the rows sharing a 16-lane half warp are
// chosen 8 (resp. 4) apart: their 8 (4) consecutive slots then fall on disjoint 8-byte bank pairs,
// and so do the partner bins' slots (H-k, Q-k, Q+k), whose rows are the digit complements.
RCG_HD constexpr int spectral_slot_pos(int log2h, int c) {
    if (log2h < 8) return (c >> 3) * 16 + (c & 7);   // fewer than 16 rows: plain order
    const int kl = c & 7, rsel = (c >> 3) & 1, s = (c >> 4) & 7, g = c >> 7;
    return (16 * g + 8 * rsel + s) * 16 + kl;
}
RCG_HD constexpr int spectral_slot_pos_hr(int log2h, int c) {
    if (log2h < 8) return (c >> 2) * 16 + (c & 3);
    const int kl = c & 3, rsel = (c >> 2) & 3, s = (c >> 4) & 3, g = c >> 6;
    return (16 * g + 4 * rsel + s) * 16 + kl;
}

// shared-memory padding: one cf of padding every 16 keeps stride-16 and stride-1 access
// patterns of all passes on distinct 8-byte bank pairs.
RCG_HD constexpr int phys(int i) { return i + (i >> 4); }
RCG_HD constexpr int padded_len(int n) { return n + (n >> 4); }

// threads cooperating on one frame, frames side by side in one CTA
RCG_HD constexpr int threads_per_frame(int log2h) {
    int h = 1 << log2h;
    int nt = h / 16;
    if (nt > 256) nt = 256;
    if (log2h >= 14) nt = 512;
    if (nt < 1) nt = 1;
    return nt;
}
// W = 16384: three frame lanes (768 threads) share one CTA per SM: 3 x 69.6 KB of the 227 KB an SM offers.  (With three
// 1-lane CTAs per SM instead, the carve-out leaves 28 KB of L1 and every twiddle load is an L2 round trip -- ncu: 19 %
// L1 hit rate.)  Where the pass twiddles of that CTA live:
//   tw_in_tmem  in Tensor Memory, one row per thread (below) -- the default
//   tw_in_smem  one copy of the packed tables behind the frame buffers (+17 KB: 226 KB), -DRCG_TW_TMEM=0
#ifndef RCG_TW_TMEM
#define RCG_TW_TMEM 1
#endif
RCG_HD constexpr bool three_lanes(int log2h) { return log2h == 13; }
// log2h = 14 (the W = 32768 kernel and the CTA pair of W = 65536): 512 threads, four warps per Tensor Memory row, 128
// columns each -- exactly the full plan's twiddles; the half-rate plan (W = 32768 only) keeps its packed tables
// log2h = 9 .. 11 (W = 1024 .. 4096; one to four warps per frame lane, several lanes and at least three CTAs per SM):
// a row serves thread t = row mod NT of every lane; 128 columns per CTA hold the full plan's twiddles and the spectral
// step's table entries (below).  log2h = 12 (256 threads, two warps per row, three CTAs per SM): 2 x 64 columns, the
// full plan's twiddles.  The half-rate plan keeps its packed tables at all of these sizes (at log2h = 9 its radix-16
// pass has 16 butterflies for a 32-thread lane, and the Tensor Memory load is warp-collective).
#ifndef RCG_TW_TMEM_512
#define RCG_TW_TMEM_512 1
#endif
#ifndef RCG_TW_TMEM_SMALL
#define RCG_TW_TMEM_SMALL 1
#endif
RCG_HD constexpr bool tw_in_tmem(int log2h, bool hr = false) {
    if (RCG_TW_TMEM == 0) return false;
    if (log2h == 13) return true;
    if (log2h == 14) return RCG_TW_TMEM_512 != 0 && !hr;
    return RCG_TW_TMEM_SMALL != 0 && log2h >= 9 && log2h <= 12 && !hr;
}
// Tensor Memory columns a CTA takes (a power of two; every CTA that can share an SM must get its share of the 512)
RCG_HD constexpr int tmem_alloc_cols(int log2h) { return log2h >= 13 ? 512 : 128; }
RCG_HD constexpr bool tw_in_smem(int log2h) { return three_lanes(log2h) && !tw_in_tmem(log2h); }
RCG_HD constexpr int frames_per_cta(int log2h) {
    if (three_lanes(log2h)) return 3;
    int nt = threads_per_frame(log2h);
    int f = 256 / nt;
    return f < 1 ? 1 : f;
}

// ---- pass twiddles in Tensor Memory (tw_in_tmem) ------------------------------------------------------------
// A thread's butterflies see the same twiddles in every frame: butterfly b = t + NT i of a pass has m = b mod M, so
// with M <= NT the thread needs ONE set w^(m j), j < R, for good, and with M > NT (the radix-2 passes fused into the
// load / overlap-add phases) M / NT sets of one value.  Blackwell's Tensor Memory (256 KB per SM: 128 rows x 512
// 32-bit columns, otherwise idle in a kernel without tensor-core work) gives warp w the rows 32 (w mod 4) .. + 31 --
// one row per thread, shared by the warps that are equal mod 4 -- and tcgen05.ld moves 8 columns = 4 twiddles of a
// row into registers per instruction, on a datapath of its own.  So every thread keeps its twiddles in its row
// (176 columns at W = 16384; the two warps of a frame lane that share a row take columns 0.. and 256.., the three
// frame lanes of a CTA read the same ones; 128 columns x four warps per row at 512 threads per frame) and no twiddle load touches the shared-memory pipe, which the butterfly
// data keep 55 % busy: a quarter of that pipe's wavefronts were twiddle loads.  Full tables for every pass also drop
// the products the packed tables needed (w^5 = w^4 w^1 ..., two-level radix-2 twiddles).
// Row layout: [full plan: pass 0 | pass 1 | ...][half-rate plan: pass 0 | ...]; a set of a radix-R pass is 2 R
// columns (twiddle j at columns 2j, 2j+1; j = 0 unused), of a radix-2 pass 2 columns.
RCG_HD constexpr int tmem_set_cols(int log2h, int pass, bool hr) {
    const PlanDesc d = make_plan(log2h, hr);
    if (d.mlog2[pass] == 0) return 0;
    return d.rlog2[pass] == 1 ? 2 : 2 << d.rlog2[pass];
}
RCG_HD constexpr int tmem_sets(int log2h, int pass, bool hr) {
    const PlanDesc d = make_plan(log2h, hr);
    const int M = 1 << d.mlog2[pass], NT = threads_per_frame(log2h);
    return M > NT ? M / NT : 1;
}
RCG_HD constexpr int tmem_col_offset(int log2h, int pass, bool hr) {
    int off = 0;
    const int np = make_plan(log2h).npasses;
    if (hr) for (int p = 0; p < np; p++) off += tmem_sets(log2h, p, false) * tmem_set_cols(log2h, p, false);
    for (int p = 0; p < pass; p++) off += tmem_sets(log2h, p, hr) * tmem_set_cols(log2h, p, hr);
    return off;
}
RCG_HD constexpr int tmem_total_cols(int log2h) {
    return tw_in_tmem(log2h, true) ? tmem_col_offset(log2h, make_plan(log2h).npasses, true)
                                   : tmem_col_offset(log2h, make_plan(log2h).npasses, false);
}
// W = 1024 .. 4096 and 16384 also keep the frame-invariant table entries of the spectral step in the row, behind the
// twiddles: per iteration of the step {slots k | H-k, slots Q-k | Q+k, k, twiddle, -} (8 columns) on the half-rate
// path, {slots, k, twiddle} (4 columns) on the full-rate one -- 64 columns at W = 16384, 32 below (two images; a
// launch takes the one of its path).  The table loads from global memory were the head of a load -> address ->
// shared-memory load chain in every iteration of the step: config 3 at W = 16384 2.07 -> 1.84 ms.
RCG_HD constexpr bool spectral_in_tmem(int log2h) { return tw_in_tmem(log2h) && (log2h == 13 || (log2h >= 9 && log2h <= 11)); }
RCG_HD constexpr int tmem_spectral_offset(int log2h) { return tmem_total_cols(log2h); }
RCG_HD constexpr int tmem_spectral_cols(int log2h) {
    return spectral_in_tmem(log2h) ? 4 * ((1 << log2h) / 2 / threads_per_frame(log2h)) : 0;
}
// W = 16384, full-rate launches (pitch 1, odd and negative multiples: BASELINE configs 3, 4, 5): the columns of the
// half-rate plan's twiddles are free, and the row takes the thread's analysis / synthesis window and crossfade envelope
// instead.  Load and overlap-add deal iteration i of their loops the same packed index m = t + NT i, and both touch
// window[2m], [2m+1], [H+2m], [H+2m+1] (4 columns per iteration, 16 iterations), the overlap-add env[2m], [2m+1] too:
// 64 + 32 columns.  Window: the half-rate plan's 80 columns from `tmem_window_offset` on; envelope: what is left of
// them (iterations 0..7) and the 16 columns behind the spectral entries (8..15).  -DRCG_WIN_TMEM=0: tables stay global.
#ifndef RCG_WIN_TMEM
#define RCG_WIN_TMEM 1
#endif
RCG_HD constexpr bool window_in_tmem(int log2h) { return RCG_WIN_TMEM != 0 && log2h == 13 && tw_in_tmem(log2h, true) && spectral_in_tmem(log2h); }
RCG_HD constexpr int tmem_window_offset(int log2h) { return tmem_col_offset(log2h, 0, true); }
RCG_HD constexpr int tmem_window_iters(int log2h) { return (1 << make_plan(log2h).mlog2[0]) / threads_per_frame(log2h); }
RCG_HD constexpr int tmem_env_col(int log2h, int iter) {
    const int half = tmem_window_iters(log2h) / 2;
    return iter < half ? tmem_window_offset(log2h) + 4 * tmem_window_iters(log2h) + 2 * iter
                       : tmem_total_cols(log2h) + tmem_spectral_cols(log2h) + 2 * (iter - half);
}
RCG_HD constexpr int tmem_image_cols(int log2h) {
    return tmem_total_cols(log2h) + tmem_spectral_cols(log2h) + (window_in_tmem(log2h) ? tmem_window_iters(log2h) : 0);
}
// the warps of a frame lane that share a row (equal mod 4) take one column group each
RCG_HD constexpr int tmem_groups(int log2h) { return threads_per_frame(log2h) > 128 ? threads_per_frame(log2h) / 128 : 1; }
RCG_HD constexpr int tmem_group_cols(int log2h) { return tmem_alloc_cols(log2h) / tmem_groups(log2h); }
// long-stride passes of the half-rate plan deal butterflies to lanes with bits 3 and 6 of m swapped (rcg_frame.h)
RCG_HD constexpr int hr_lane_swizzle(int m) { return (m & ~0x48) | ((m & 0x08) << 3) | ((m & 0x40) >> 3); }
// which m the butterflies of thread t see in set `set` of a pass, and a thread (row) that sees a given m
RCG_HD constexpr int tmem_m_of_thread(int log2h, int pass, bool hr, int t, int set) {
    const PlanDesc d = make_plan(log2h, hr);
    const int M = 1 << d.mlog2[pass], NT = threads_per_frame(log2h);
    if (M > NT) return t + NT * set;
    const int m = t & (M - 1);
    return (hr && pass >= 1 && M >= 128) ? hr_lane_swizzle(m) : m;   // pass 0 is fused into the overlap-add: no swizzle
}
RCG_HD constexpr int tmem_row_of_m(int log2h, int pass, bool hr, int m) {
    const PlanDesc d = make_plan(log2h, hr);
    const int M = 1 << d.mlog2[pass], NT = threads_per_frame(log2h);
    if (M >= NT) return m & (NT - 1);
    return (hr && pass >= 1 && M >= 128) ? hr_lane_swizzle(m) : m;
}

// ---- twiddle tables --------------------------------------------------------------------
// Pass p needs w^(m*j), w = exp(-2 pi i / (R*M)), m < M, 0 < j < R.  How they are stored:
//   none      M == 1
//   full      M <= 16 (or a radix-2 pass with M < 64): [(R-1)][M], L1-resident
//   pow       R >= 4, M > 16: only the power-of-two exponents j = 1, 2, 4, 8 are stored
//             ([log2 R][M]); the others are products of two stored/kept values
//   twolevel  R == 2, M >= 64: w^m = A[m >> 6] * B[m & 63]  (A: M/64 entries, B: 64)
enum TwScheme : int { kTwNone = 0, kTwFull = 1, kTwPow = 2, kTwTwoLevel = 3 };

RCG_HD constexpr int tw_scheme(int log2h, int pass, bool hr = false) {
    const PlanDesc d = make_plan(log2h, hr);
    const int M = 1 << d.mlog2[pass], R = 1 << d.rlog2[pass];
    if (M == 1) return kTwNone;
    if (R == 2) return M >= 64 ? kTwTwoLevel : kTwFull;
    if (M <= 16) return kTwFull;
    return kTwPow;
}
RCG_HD constexpr int tw_size(int log2h, int pass, bool hr = false) {
    const PlanDesc d = make_plan(log2h, hr);
    const int M = 1 << d.mlog2[pass], R = 1 << d.rlog2[pass];
    switch (tw_scheme(log2h, pass, hr)) {
        case kTwFull: return (R - 1) * M;
        case kTwPow: return d.rlog2[pass] * M;
        case kTwTwoLevel: return M / 64 + 64;
        default: return 0;
    }
}
// full-plan tables first, then the half-rate plan's
RCG_HD constexpr int tw_offset(int log2h, int pass, bool hr = false) {
    int off = 0;
    const int np = make_plan(log2h).npasses;
    if (hr) for (int p = 0; p < np; p++) off += tw_size(log2h, p, false);
    for (int p = 0; p < pass; p++) off += tw_size(log2h, p, hr);
    return off;
}
RCG_HD constexpr int tw_total(int log2h) { return tw_offset(log2h, make_plan(log2h).npasses, true); }

}  // namespace rcg

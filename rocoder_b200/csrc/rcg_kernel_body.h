// rcg_kernel_body.h -- the body of the fused stretch kernel: barrier flavours of a frame lane, the shared-memory
// copy of the twiddle tables, the item loop.  Shared by rcg_kernels.cu (nvcc, all window sizes) and by the
// translation unit NVRTC builds at run time when a bin-local frequency kernel is fused into the stretch kernel
// (rcg_jit_entry.h).  Device code only.
#pragma once
#include "rcg_frame.h"

namespace rcg {

// WARP_LOCAL: a frame lane is at most one warp (W <= 1024), so the frames sharing a CTA never
// exchange data and every phase boundary is a warp barrier: the lanes drift apart and overlap their
// memory and compute phases instead of marching in lockstep.
// Otherwise, when several multi-warp frame lanes share a CTA (W = 2048, 4096), each lane
// synchronises on its own named barrier (bar.sync id, NT) for the same reason.
template <bool WARP_LOCAL, int LANE_THREADS = 0>
struct DeviceExecT {
    int t;
    int lane = 0;
    __device__ __forceinline__ void lane_barrier() {
        if constexpr (WARP_LOCAL) __syncwarp();
        else if constexpr (LANE_THREADS > 0) asm volatile("bar.sync %0, %1;" ::"r"(lane + 1), "n"(LANE_THREADS) : "memory");
        else __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void par(F f) {
        f(t);
        lane_barrier();
    }
    // steps that only depend on each other within a warp: warp barriers in between, one block
    // barrier at the end
    template <class F>
    __device__ __forceinline__ void par_steps(int n, F f) {
        for (int s = 0; s < n; s++) {
            f(s, t);
            __syncwarp();
        }
        if constexpr (!WARP_LOCAL) lane_barrier();
    }
};
using DeviceExec = DeviceExecT<false>;


// FrameCfg::TW_SMEM: the CTA copies the pass twiddle tables behind its frame buffers once (the CTAs are
// persistent) and the frame code reads them from there
template <int LOG2H>
__device__ __forceinline__ StretchParams with_smem_twiddles(const StretchParams& prm, cf* smem_base) {
    using C = FrameCfg<LOG2H>;
    StretchParams p = prm;
    if constexpr (C::TW_SMEM) {
        cf* tw = smem_base + C::FPC * C::SMEM_CF;
        for (int i = threadIdx.x; i < C::TW_CF; i += blockDim.x) tw[i] = __ldg(prm.tw + i);
        __syncthreads();
        p.tw = tw;
    }
    return p;
}

// FrameCfg::TW_TMEM (rcg_common.h): the CTA takes its share of the SM's Tensor Memory columns (all 512 where it is
// alone on the SM because the frame buffers fill the shared memory, 128 where three or four CTAs share one), the
// first threads copy the rows from the image in global memory (once per launch), and every thread remembers where
// its row starts: rows 32 (warp mod 4) .., columns 0.. for the first four warps of a frame lane, the next column
// group for the next four.
template <int LOG2H>
__device__ __forceinline__ StretchParams with_tmem_twiddles(const StretchParams& prm, uint32_t* slot) {
    using C = FrameCfg<LOG2H>;
    StretchParams p = prm;
    if constexpr (C::TW_TMEM) {
        constexpr int NCOL = tmem_image_cols(LOG2H), ALLOC = tmem_alloc_cols(LOG2H);
        static_assert(NCOL % 8 == 0 && NCOL <= tmem_group_cols(LOG2H), "a column group per warp that shares the row");
        static_assert(C::NT >= 32, "a row serves one thread of a frame lane");
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "n"(ALLOC) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // hardware rows of this warp: 32 (warp mod 4) ..; frame lanes narrower than four warps repeat every NT rows
        const int warp = (int)(threadIdx.x / 32), t = (int)(threadIdx.x % C::NT);
        const int group = C::NT > 128 ? (t / 32) / 4 : 0;
        const uint32_t mine = *slot + ((uint32_t)(32 * (warp % 4)) << 16) + (uint32_t)(group * tmem_group_cols(LOG2H));
        if (threadIdx.x < (C::NT > 128 ? C::NT : 128)) {   // every row of every group once
            for (int c = 0; c < NCOL; c += 8) {
                uint32_t r[8];
#pragma unroll
                for (int q = 0; q < 8; q++) r[q] = __float_as_uint(__ldg(prm.tw_image + (size_t)(c + q) * C::NT + t));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(mine + (uint32_t)c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
                               "r"(r[6]), "r"(r[7]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        p.tmem_addr = mine;
    }
    return p;
}
// ... and hands the columns back before it exits
template <int LOG2H>
__device__ __forceinline__ void release_tmem_twiddles(const uint32_t* slot) {
    if constexpr (FrameCfg<LOG2H>::TW_TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot), "n"(tmem_alloc_cols(LOG2H)) : "memory");
    }
}
// the word behind the frame buffers (and the shared-memory tables, where there are any) that receives the allocation
template <int LOG2H>
__device__ __forceinline__ uint32_t* tmem_slot(unsigned char* smem_raw) {
    using C = FrameCfg<LOG2H>;
    return reinterpret_cast<uint32_t*>(reinterpret_cast<cf*>(smem_raw) + (size_t)C::FPC * C::SMEM_CF + (C::TW_SMEM ? C::TW_CF : 0));
}
template <int LOG2H>
constexpr size_t kernel_smem_bytes() {
    using C = FrameCfg<LOG2H>;
    return ((size_t)C::FPC * C::SMEM_CF + (C::TW_SMEM ? C::TW_CF : 0)) * sizeof(cf) + (C::TW_TMEM ? 16 : 0);
}
template <int LOG2H>
constexpr int min_ctas() { return (LOG2H >= 14 || three_lanes(LOG2H)) ? 1 : 3; }

// MODE: the launch's scalar configuration as compile-time constants.  The frame code reads pitch and flags in
// many places (which synthesis path, where the overlap-add phase stores, whether a resample phase exists); with
// the values known the compiler drops the other paths: config 2 6.92 -> 6.46 ms.  The arithmetic is the same in
// every mode.
//   0  generic (bypass hook, anything else)     1  pitch 2 on the half-rate path (the headline `-p 2`)
//   2  pitch 1, full rate                       3  any other pitch, full rate
//   4  pitch -2                                 5  pitch -3 (BASELINE config 4): the lerp walk's strides are constants
// (stretch_mode_of, the rule that picks MODE from pitch and flags: rcg_frame.h, shared with the host)
template <int MODE>
__device__ __forceinline__ StretchParams with_mode_constants(const StretchParams& in) {
    StretchParams p = in;
    if constexpr (MODE == 1) { p.pitch = 2; p.flags = (in.flags & kFlagDeferFirst) | kFlagHalfRate; }
    // the full-rate modes read window and envelope from the Tensor Memory row where the image holds them (window_in_tmem)
    if constexpr (MODE == 2) { p.pitch = 1; p.flags = (in.flags & kFlagDeferFirst) | kFlagWinTmem; }
    if constexpr (MODE == 3) { p.flags = (in.flags & kFlagDeferFirst) | kFlagWinTmem; }
    if constexpr (MODE == 4) { p.pitch = -2; p.flags = (in.flags & kFlagDeferFirst) | kFlagWinTmem; }
    if constexpr (MODE == 5) { p.pitch = -3; p.flags = (in.flags & kFlagDeferFirst) | kFlagWinTmem; }
    return p;
}

template <int LOG2H, bool SPEC_IN>
__device__ __forceinline__ void stretch_body(const StretchParams& prm, int n_items, unsigned char* smem_raw) {
    using C = FrameCfg<LOG2H>;
    const int lane = threadIdx.x / C::NT;
    // frame lanes of one CTA are independent: warp barriers (<= 1 warp per lane) or one named barrier per lane
    DeviceExecT<(C::NT <= 32), (C::NT > 32 && C::FPC > 1) ? C::NT : 0> ex{(int)(threadIdx.x % C::NT), lane};
    cf* S = reinterpret_cast<cf*>(smem_raw) + lane * C::SMEM_CF;
    float* tail = prm.tail_scratch + ((long long)blockIdx.x * C::FPC + lane) * C::H;
    // A full round gives CTA c the FPC neighbouring items r * per_round + c * FPC + l (neighbouring runs of a
    // channel share input and tables in L1 / L2: 6.77 vs 6.95 ms on config 2); a partial last round spreads its items
    // over the CTAs first, item r * per_round + l * grid + c, so that a small job keeps every SM busy before any
    // CTA runs a second lane.  The trip count is uniform over the CTA.
    const int per_round = (int)gridDim.x * C::FPC;
    const int rounds = (n_items + per_round - 1) / per_round;
    for (int r = 0; r < rounds; r++) {
        const int first = r * per_round;
        const bool full = n_items - first >= per_round;
        const int item = first + (full ? (int)blockIdx.x * C::FPC + lane : lane * (int)gridDim.x + (int)blockIdx.x);
        run_work_item<LOG2H, SPEC_IN>(ex, S, prm, item, tail, item < n_items);
    }
}
}  // namespace rcg

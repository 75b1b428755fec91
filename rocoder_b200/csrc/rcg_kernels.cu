// rcg_kernels.cu -- sm_100a kernels of the stretch path.
//
// stretch_kernel: one frame lane (NT threads) owns a run of consecutive frames of one channel
// and, per frame, does everything ReFFT::resynth + the body of Stretcher::next_window do
// (reference fft.rs:42-74, stretcher.rs:94-112, resampler.rs:3-35) without leaving the SM:
//   window * samples -> packed real FFT (in-place radix-16 passes in shared memory)
//   -> |X| -> Philox random phases for all W bins -> Hermitian fold -> inverse real FFT
//   -> /W * window -> overlap-add with the carried half window -> envelope * amp -> resample.
// HBM sees each input sample once per overlapping frame (L2 absorbs the overlap) and each
// output sample exactly once.  No tensor cores: nothing here is a dense contraction.
//
// One source, several translation units so that the window sizes compile in parallel (rocoder_b200/build.py):
//   -DRCG_TU_LOG2H=L   the kernels and launchers of W = 2^(L+1), L = 5..14
//   (nothing)          the W = 65536 CTA-pair kernels and the dispatch by window size
#include <algorithm>

#include "rcg_kernels.h"

#include "rcg_frame_pair.h"
#include "rcg_kernel_body.h"
#include "rcg_pair_body.h"

namespace rcg {

#if defined(RCG_TU_LOG2H)

template <int LOG2H, bool SPEC_IN = false, int MODE = 0>
__global__ void __launch_bounds__(FrameCfg<LOG2H>::NT * FrameCfg<LOG2H>::FPC, min_ctas<LOG2H>())
stretch_kernel(const __grid_constant__ StretchParams prm_in, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if constexpr (MODE == 0 && !FrameCfg<LOG2H>::TW_SMEM && !FrameCfg<LOG2H>::TW_TMEM) {
        stretch_body<LOG2H, SPEC_IN>(prm_in, n_items, smem_raw);
    } else {
        const StretchParams prm = with_mode_constants<MODE>(prm_in);
        if constexpr (FrameCfg<LOG2H>::TW_TMEM) {
            stretch_body<LOG2H, SPEC_IN>(with_tmem_twiddles<LOG2H>(prm, tmem_slot<LOG2H>(smem_raw)), n_items, smem_raw);
            release_tmem_twiddles<LOG2H>(tmem_slot<LOG2H>(smem_raw));
        } else if constexpr (FrameCfg<LOG2H>::TW_SMEM) {
            stretch_body<LOG2H, SPEC_IN>(with_smem_twiddles<LOG2H>(prm, reinterpret_cast<cf*>(smem_raw)), n_items, smem_raw);
        } else {
            stretch_body<LOG2H, SPEC_IN>(prm, n_items, smem_raw);
        }
    }
}

// plugin path, analysis half: frames -> full complex spectra in HBM for the host `apply` call
template <int LOG2H>
__device__ __forceinline__ void analysis_body(const StretchParams& prm, int n_items, unsigned char* smem_raw) {
    using C = FrameCfg<LOG2H>;
    const int lane = threadIdx.x / C::NT;
    DeviceExec ex{(int)(threadIdx.x % C::NT)};
    cf* S = reinterpret_cast<cf*>(smem_raw) + lane * C::SMEM_CF;
    const int per_round = (int)gridDim.x * C::FPC;
    const int rounds = (n_items + per_round - 1) / per_round;
    for (int r = 0; r < rounds; r++) {   // same item order as stretch_body
        const int first = r * per_round;
        const bool full = n_items - first >= per_round;
        const int item = first + (full ? (int)blockIdx.x * C::FPC + lane : lane * (int)gridDim.x + (int)blockIdx.x);
        run_analysis_item<LOG2H>(ex, S, prm, item, item < n_items);
    }
}
template <int LOG2H>
__global__ void __launch_bounds__(FrameCfg<LOG2H>::NT * FrameCfg<LOG2H>::FPC, min_ctas<LOG2H>())
analysis_kernel(const __grid_constant__ StretchParams prm, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if constexpr (FrameCfg<LOG2H>::TW_TMEM) {
        analysis_body<LOG2H>(with_tmem_twiddles<LOG2H>(prm, tmem_slot<LOG2H>(smem_raw)), n_items, smem_raw);
        release_tmem_twiddles<LOG2H>(tmem_slot<LOG2H>(smem_raw));
    } else if constexpr (FrameCfg<LOG2H>::TW_SMEM) {
        analysis_body<LOG2H>(with_smem_twiddles<LOG2H>(prm, reinterpret_cast<cf*>(smem_raw)), n_items, smem_raw);
    } else {
        analysis_body<LOG2H>(prm, n_items, smem_raw);
    }
}

template <int LOG2H>
__global__ void __launch_bounds__(FrameCfg<LOG2H>::NT)
boundary_kernel(const __grid_constant__ StretchParams prm, const float* head, const float* tail_prev) {
    using C = FrameCfg<LOG2H>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DeviceExec ex{(int)threadIdx.x};
    run_boundary<LOG2H>(ex, reinterpret_cast<cf*>(smem_raw), prm, (int)blockIdx.x, head, tail_prev);
}

template <int LOG2H>
cudaError_t geometry_t(LaunchGeometry* g) {
    using C = FrameCfg<LOG2H>;
    g->threads = C::NT * C::FPC;
    g->frames_per_cta = C::FPC;
    g->smem_bytes = kernel_smem_bytes<LOG2H>();
    cudaError_t e = cudaSuccess;
    const void* kernels[] = {(const void*)stretch_kernel<LOG2H, false, 0>, (const void*)stretch_kernel<LOG2H, false, 1>,
                             (const void*)stretch_kernel<LOG2H, false, 2>, (const void*)stretch_kernel<LOG2H, false, 3>,
                             (const void*)stretch_kernel<LOG2H, false, 4>, (const void*)stretch_kernel<LOG2H, false, 5>,
                             (const void*)stretch_kernel<LOG2H, true, 0>};
    for (const void* k : kernels) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes);
        if (e != cudaSuccess) return e;
    }
    e = cudaFuncSetAttribute(analysis_kernel<LOG2H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)g->smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(boundary_kernel<LOG2H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(C::SMEM_CF * sizeof(cf)));
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stretch_kernel<LOG2H, false, 0>, g->threads, g->smem_bytes);
    if (e == cudaSuccess && C::TW_TMEM && tmem_alloc_cols(LOG2H) < 512) {
        // A kernel that allocates Tensor Memory is reported as one CTA per SM (ncu: grid 148 instead of 444) although
        // each CTA takes 128 of the 512 columns: count by registers and shared memory, capped by the columns.
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, stretch_kernel<LOG2H, false, 0>);
        if (e != cudaSuccess) return e;
        cudaDeviceProp prop;
        int dev = 0;
        cudaGetDevice(&dev);
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return e;
        const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * g->threads;
        const int by_regs = regs_per_cta > 0 ? prop.regsPerMultiprocessor / regs_per_cta : 1;
        const int by_smem = (int)(prop.sharedMemPerMultiprocessor / (g->smem_bytes + fa.sharedSizeBytes + prop.reservedSharedMemPerBlock));
        const int by_threads = prop.maxThreadsPerMultiProcessor / g->threads;
        n = std::max(n, std::min(std::min(by_regs, by_smem), std::min(by_threads, 512 / tmem_alloc_cols(LOG2H))));
    }
    g->max_ctas_per_sm = n;
    return e;
}

template <int LOG2H>
cudaError_t launch_t(const StretchParams& prm, int grid, cudaStream_t stream) {
    using C = FrameCfg<LOG2H>;
    const int n_items = prm.channels * prm.n_runs;
    if (n_items <= 0) return cudaSuccess;
    const dim3 block(C::NT * C::FPC);
    const size_t smem = kernel_smem_bytes<LOG2H>();
    if (prm.flags & kFlagSpectrumIn) {
        stretch_kernel<LOG2H, true, 0><<<grid, block, smem, stream>>>(prm, n_items);
    } else {
        switch (stretch_mode_of(prm.pitch, prm.flags)) {   // same arithmetic, scalar configuration folded in
            case 1: stretch_kernel<LOG2H, false, 1><<<grid, block, smem, stream>>>(prm, n_items); break;
            case 2: stretch_kernel<LOG2H, false, 2><<<grid, block, smem, stream>>>(prm, n_items); break;
            case 3: stretch_kernel<LOG2H, false, 3><<<grid, block, smem, stream>>>(prm, n_items); break;
            case 4: stretch_kernel<LOG2H, false, 4><<<grid, block, smem, stream>>>(prm, n_items); break;
            case 5: stretch_kernel<LOG2H, false, 5><<<grid, block, smem, stream>>>(prm, n_items); break;
            default: stretch_kernel<LOG2H, false, 0><<<grid, block, smem, stream>>>(prm, n_items); break;
        }
    }
    return cudaGetLastError();
}

template <int LOG2H>
cudaError_t analysis_t(const StretchParams& prm, int grid, cudaStream_t stream) {
    using C = FrameCfg<LOG2H>;
    const int n_items = prm.channels * prm.n_runs;
    if (n_items <= 0) return cudaSuccess;
    analysis_kernel<LOG2H><<<grid, C::NT * C::FPC, kernel_smem_bytes<LOG2H>(), stream>>>(prm, n_items);
    return cudaGetLastError();
}

template <int LOG2H>
cudaError_t boundary_t(const StretchParams& prm, const float* head, const float* tail_prev,
                              cudaStream_t stream) {
    using C = FrameCfg<LOG2H>;
    boundary_kernel<LOG2H><<<prm.channels, C::NT, C::SMEM_CF * sizeof(cf), stream>>>(prm, head, tail_prev);
    return cudaGetLastError();
}

template cudaError_t geometry_t<RCG_TU_LOG2H>(LaunchGeometry*);
template cudaError_t launch_t<RCG_TU_LOG2H>(const StretchParams&, int, cudaStream_t);
template cudaError_t analysis_t<RCG_TU_LOG2H>(const StretchParams&, int, cudaStream_t);
template cudaError_t boundary_t<RCG_TU_LOG2H>(const StretchParams&, const float*, const float*, cudaStream_t);

#else   // ---- main translation unit ---------------------------------------------------------------------------

// defined by the per-size translation units
template <int LOG2H> cudaError_t geometry_t(LaunchGeometry* g);
template <int LOG2H> cudaError_t launch_t(const StretchParams& prm, int grid, cudaStream_t stream);
template <int LOG2H> cudaError_t analysis_t(const StretchParams& prm, int grid, cudaStream_t stream);
template <int LOG2H> cudaError_t boundary_t(const StretchParams& prm, const float* head, const float* tail_prev, cudaStream_t stream);

// ---- W = 65536: a frame across a 2-CTA cluster (rcg_frame_pair.h) -------------------------------
// PairDeviceExec, stretch_pair_body: rcg_pair_body.h (shared with the NVRTC translation unit)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FrameCfg<kPairLog2HL>::NT, 1)
stretch_pair_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ PairTables pt, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    stretch_pair_body(prm, pt, n_items, smem_raw);
}

// frequency-kernel path at W = 65536, analysis half: each CTA of the pair writes the bins of its parity
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FrameCfg<kPairLog2HL>::NT, 1)
analysis_pair_kernel(const __grid_constant__ StretchParams prm_in, const __grid_constant__ PairTables pt, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int r = (int)cluster_cta_rank();
    cf* S = reinterpret_cast<cf*>(smem_raw);
    cf* bufs[2] = {S, S};   // only bufs[r] is touched: no data crosses the pair in this kernel
    PairDeviceExec ex{r, (int)threadIdx.x};
    uint32_t* tm_slot = reinterpret_cast<uint32_t*>(smem_raw + kPairSmemBytes - 16);
    const StretchParams prm = with_tmem_twiddles<kPairLog2HL>(prm_in, tm_slot);
    const int cid = blockIdx.x / 2, ncl = gridDim.x / 2;
    for (int item = cid; item < n_items; item += ncl) run_analysis_item_pair(ex, bufs, prm, pt, item, true);
    release_tmem_twiddles<kPairLog2HL>(tm_slot);
}

__global__ void __launch_bounds__(FrameCfg<kPairLog2HL>::NT)
boundary_pair_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ PairTables pt, const float* head,
                     const float* tail_prev) {
    DeviceExec ex{(int)threadIdx.x};
    float* stage = pt.scratch + (long long)blockIdx.x * 3 * kPairH + 2 * kPairH;
    run_boundary_pair(ex, prm, (int)blockIdx.x, head, tail_prev, stage);
}

cudaError_t pair_geometry(LaunchGeometry* g) {
    using C = FrameCfg<kPairLog2HL>;
    g->threads = C::NT;
    g->frames_per_cta = 1;
    g->smem_bytes = kPairSmemBytes;   // transform buffer + landing zone + small twiddle tables
    g->max_ctas_per_sm = 1;
    cudaError_t e = cudaFuncSetAttribute(analysis_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(stretch_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes);
}
cudaError_t launch_analysis_pair(const StretchParams& prm, const PairTables& pt, int n_clusters, cudaStream_t stream) {
    using C = FrameCfg<kPairLog2HL>;
    const int n_items = prm.channels * prm.n_runs;
    if (n_items <= 0) return cudaSuccess;
    analysis_pair_kernel<<<2 * n_clusters, C::NT, kPairSmemBytes, stream>>>(prm, pt, n_items);
    return cudaGetLastError();
}
cudaError_t launch_stretch_pair(const StretchParams& prm, const PairTables& pt, int n_clusters, cudaStream_t stream) {
    using C = FrameCfg<kPairLog2HL>;
    const int n_items = prm.channels * prm.n_runs;
    if (n_items <= 0) return cudaSuccess;
    stretch_pair_kernel<<<2 * n_clusters, C::NT, kPairSmemBytes, stream>>>(prm, pt, n_items);
    return cudaGetLastError();
}
cudaError_t launch_boundary_pair(const StretchParams& prm, const PairTables& pt, const float* head, const float* tail_prev,
                                 cudaStream_t stream) {
    boundary_pair_kernel<<<prm.channels, FrameCfg<kPairLog2HL>::NT, 0, stream>>>(prm, pt, head, tail_prev);
    return cudaGetLastError();
}

#ifdef RCG_ONLY_LOG2H  /* experiment builds: one window size only, to compile in seconds */
#define RCG_DISPATCH(log2w, CALL)                                    \
    switch ((log2w) - 1) {                                           \
        case RCG_ONLY_LOG2H: return CALL(RCG_ONLY_LOG2H);            \
        default: return cudaErrorInvalidValue;                       \
    }
#else
#define RCG_DISPATCH(log2w, CALL)                                    \
    switch ((log2w) - 1) {                                           \
        case 5: return CALL(5);   case 6: return CALL(6);            \
        case 7: return CALL(7);   case 8: return CALL(8);            \
        case 9: return CALL(9);   case 10: return CALL(10);          \
        case 11: return CALL(11); case 12: return CALL(12);          \
        case 13: return CALL(13); case 14: return CALL(14);          \
        default: return cudaErrorInvalidValue;                       \
    }
#endif

cudaError_t stretch_geometry(int log2w, LaunchGeometry* g) {
#define CALL(L) geometry_t<L>(g)
    RCG_DISPATCH(log2w, CALL)
#undef CALL
}
cudaError_t launch_stretch(int log2w, const StretchParams& prm, int grid, cudaStream_t stream) {
#define CALL(L) launch_t<L>(prm, grid, stream)
    RCG_DISPATCH(log2w, CALL)
#undef CALL
}
cudaError_t launch_analysis(int log2w, const StretchParams& prm, int grid, cudaStream_t stream) {
#define CALL(L) analysis_t<L>(prm, grid, stream)
    RCG_DISPATCH(log2w, CALL)
#undef CALL
}
cudaError_t launch_boundary(int log2w, const StretchParams& prm, const float* head, const float* tail_prev,
                            cudaStream_t stream) {
#define CALL(L) boundary_t<L>(prm, head, tail_prev, stream)
    RCG_DISPATCH(log2w, CALL)
#undef CALL
}

#endif  // RCG_TU_LOG2H

}  // namespace rcg

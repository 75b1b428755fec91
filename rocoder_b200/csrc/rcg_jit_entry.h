// rcg_jit_entry.h -- entry point of the translation unit NVRTC builds when a BIN-LOCAL frequency kernel is fused
// into the stretch kernel (rcg_push_local_device_frequency_kernel).  The engine prepends the user's source, which
// defines
//     __device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x)
// (the reference's `apply`, fft.rs:93, restricted to kernels whose output bin depends on the same input bin only --
// the README's own example, README.md:121-129, is one), and defines RCG_JIT_LOG2H = log2(W) - 1.  Because only
// the norm of a kernel's output survives (fft.rs:67), the fused spectral step evaluates |apply(X[k])| in
// registers: the spectra never exist in memory and the frame costs what it costs without a kernel.
// RCG_JIT_MODE (optional) = the launch's scalar configuration as compile-time constants, exactly as for the library's
// own instantiations (stretch_mode_of / with_mode_constants, rcg_kernel_body.h): the engine passes it when the mode
// cannot change during the engine's life (odd or negative pitch multiples: no half-rate path to fall out of).
#pragma once
#define RCG_PLUGIN_LOCAL 1
#include "rcg_kernel_body.h"
#ifndef RCG_JIT_MODE
#define RCG_JIT_MODE 0
#endif

#if RCG_JIT_LOG2H == 15   // W = 65536: the CTA-pair kernel (rcg_frame_pair.h)
#include "rcg_pair_body.h"
extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(rcg::FrameCfg<rcg::kPairLog2HL>::NT, 1)
rcg_jit_stretch(const __grid_constant__ rcg::StretchParams prm_in, const __grid_constant__ rcg::PairTables pt, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const rcg::StretchParams prm = rcg::with_mode_constants<RCG_JIT_MODE>(prm_in);
    rcg::stretch_pair_body(prm, pt, n_items, smem_raw);
}
#else
extern "C" __global__ void
__launch_bounds__(rcg::FrameCfg<RCG_JIT_LOG2H>::NT * rcg::FrameCfg<RCG_JIT_LOG2H>::FPC, rcg::min_ctas<RCG_JIT_LOG2H>())
rcg_jit_stretch(const __grid_constant__ rcg::StretchParams prm_in, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const rcg::StretchParams prm = rcg::with_mode_constants<RCG_JIT_MODE>(prm_in);
    if constexpr (rcg::FrameCfg<RCG_JIT_LOG2H>::TW_TMEM) {
        rcg::stretch_body<RCG_JIT_LOG2H, false>(rcg::with_tmem_twiddles<RCG_JIT_LOG2H>(prm, rcg::tmem_slot<RCG_JIT_LOG2H>(smem_raw)),
                                                n_items, smem_raw);
        rcg::release_tmem_twiddles<RCG_JIT_LOG2H>(rcg::tmem_slot<RCG_JIT_LOG2H>(smem_raw));
    } else if constexpr (rcg::FrameCfg<RCG_JIT_LOG2H>::TW_SMEM) {
        rcg::stretch_body<RCG_JIT_LOG2H, false>(rcg::with_smem_twiddles<RCG_JIT_LOG2H>(prm, reinterpret_cast<cf*>(smem_raw)),
                                                n_items, smem_raw);
    } else {
        rcg::stretch_body<RCG_JIT_LOG2H, false>(prm, n_items, smem_raw);
    }
}
#endif

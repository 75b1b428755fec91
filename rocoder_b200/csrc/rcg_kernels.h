// rcg_kernels.h -- host-callable launchers of the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include "rcg_frame.h"
#include "rcg_big.h"
#include "rcg_frame_bluestein.h"
#include "rcg_frame_pair.h"

namespace rcg {

struct LaunchGeometry {
    int threads;          // per CTA
    int frames_per_cta;   // work lanes per CTA
    size_t smem_bytes;
    int max_ctas_per_sm;  // occupancy of the stretch kernel
};

// log2w in [6, 15].  Returns cudaErrorInvalidValue for unsupported sizes.
cudaError_t stretch_geometry(int log2w, LaunchGeometry* g);
// Launch the fused stretch kernel over prm.channels * prm.n_runs work items with `grid`
// persistent CTAs.  prm.tail_scratch must hold grid * frames_per_cta * H floats.
cudaError_t launch_stretch(int log2w, const StretchParams& prm, int grid, cudaStream_t stream);
// plugin path: spectra of frames [frame_begin, frame_end) into prm.spectrum
cudaError_t launch_analysis(int log2w, const StretchParams& prm, int grid, cudaStream_t stream);
cudaError_t launch_boundary(int log2w, const StretchParams& prm, const float* head, const float* tail_prev,
                            cudaStream_t stream);

// W = 65536 (log2w == 16): a frame across a 2-CTA cluster; pt.scratch holds 3*H floats per cluster
// (and per channel for the boundary kernel)
cudaError_t pair_geometry(LaunchGeometry* g);
cudaError_t launch_stretch_pair(const StretchParams& prm, const PairTables& pt, int n_clusters, cudaStream_t stream);
cudaError_t launch_analysis_pair(const StretchParams& prm, const PairTables& pt, int n_clusters, cudaStream_t stream);
cudaError_t launch_boundary_pair(const StretchParams& prm, const PairTables& pt, const float* head, const float* tail_prev,
                                 cudaStream_t stream);

// windows that are not a power of two (rcg_frame_bluestein.h): lm = log2 of the convolution length, 7..14;
// prm.tail_scratch must hold grid * W/2 floats
cudaError_t bluestein_geometry(int lm, int W, LaunchGeometry* g);
cudaError_t launch_bluestein(int lm, const StretchParams& prm, const BluesteinTables& bt, int grid, cudaStream_t stream);
cudaError_t launch_boundary_bluestein(int lm, const StretchParams& prm, const BluesteinTables& bt, const float* head,
                                      const float* tail_prev, cudaStream_t stream);

}  // namespace rcg

// rcg_bluestein.cu -- kernels for windows that are not a power of two (rcg_frame_bluestein.h): one CTA per run of
// frames, the M = 2^LM point chirp convolutions in shared memory.  Its own translation unit (builds in parallel).
#include "rcg_kernels.h"

#include "rcg_frame_bluestein.h"
#include "rcg_kernel_body.h"

namespace rcg {

template <int LM>
constexpr size_t bluestein_smem_bytes(int W) { return (size_t)FrameCfg<LM>::SMEM_CF * sizeof(cf) + (size_t)(W / 2) * sizeof(float); }

template <int LM>
__global__ void __launch_bounds__(FrameCfg<LM>::NT)
bluestein_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ BluesteinTables bt, int n_items) {
    using C = FrameCfg<LM>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf* S = reinterpret_cast<cf*>(smem_raw);
    float* zbuf = reinterpret_cast<float*>(S + C::SMEM_CF);
    DeviceExec ex{(int)threadIdx.x};
    float* tail = prm.tail_scratch + (long long)blockIdx.x * (bt.w / 2);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
        run_work_item_bluestein<LM>(ex, S, zbuf, prm, bt, item, tail, true);
}

template <int LM>
__global__ void __launch_bounds__(FrameCfg<LM>::NT)
bluestein_boundary_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ BluesteinTables bt, const float* head,
                          const float* tail_prev) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DeviceExec ex{(int)threadIdx.x};
    run_boundary_bluestein<LM>(ex, reinterpret_cast<float*>(smem_raw), prm, bt, (int)blockIdx.x, head, tail_prev);
}

template <int LM>
static cudaError_t geometry_b(int W, LaunchGeometry* g) {
    using C = FrameCfg<LM>;
    g->threads = C::NT;
    g->frames_per_cta = 1;
    g->smem_bytes = bluestein_smem_bytes<LM>(W);
    cudaError_t e = cudaFuncSetAttribute(bluestein_kernel<LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bluestein_kernel<LM>, g->threads, g->smem_bytes);
    g->max_ctas_per_sm = n;
    return e;
}
template <int LM>
static cudaError_t launch_b(const StretchParams& prm, const BluesteinTables& bt, int grid, cudaStream_t stream) {
    const int n_items = prm.channels * prm.n_runs;
    if (n_items <= 0) return cudaSuccess;
    bluestein_kernel<LM><<<grid, FrameCfg<LM>::NT, bluestein_smem_bytes<LM>(bt.w), stream>>>(prm, bt, n_items);
    return cudaGetLastError();
}
template <int LM>
static cudaError_t boundary_b(const StretchParams& prm, const BluesteinTables& bt, const float* head, const float* tail_prev,
                              cudaStream_t stream) {
    bluestein_boundary_kernel<LM><<<prm.channels, FrameCfg<LM>::NT, (size_t)(bt.w / 2) * sizeof(float), stream>>>(prm, bt, head, tail_prev);
    return cudaGetLastError();
}

#define RCG_BS_DISPATCH(lm, CALL)                                    \
    switch (lm) {                                                    \
        case 7: return CALL(7);   case 8: return CALL(8);            \
        case 9: return CALL(9);   case 10: return CALL(10);          \
        case 11: return CALL(11); case 12: return CALL(12);          \
        case 13: return CALL(13); case 14: return CALL(14);          \
        default: return cudaErrorInvalidValue;                       \
    }

cudaError_t bluestein_geometry(int lm, int W, LaunchGeometry* g) {
#define CALL(L) geometry_b<L>(W, g)
    RCG_BS_DISPATCH(lm, CALL)
#undef CALL
}
cudaError_t launch_bluestein(int lm, const StretchParams& prm, const BluesteinTables& bt, int grid, cudaStream_t stream) {
#define CALL(L) launch_b<L>(prm, bt, grid, stream)
    RCG_BS_DISPATCH(lm, CALL)
#undef CALL
}
cudaError_t launch_boundary_bluestein(int lm, const StretchParams& prm, const BluesteinTables& bt, const float* head,
                                      const float* tail_prev, cudaStream_t stream) {
#define CALL(L) boundary_b<L>(prm, bt, head, tail_prev, stream)
    RCG_BS_DISPATCH(lm, CALL)
#undef CALL
}

}  // namespace rcg

// warp_exchange.cu -- why the FFT passes exchange data through shared memory and not through warp shuffles.
// Between two radix-16 passes a thread hands its 16 complex values to 16 other threads and receives 16 (a 16 x 16
// transpose inside each half warp).  Two ways to do that for a warp's 512 values:
//   smem : 16 STS.64 + __syncwarp + 16 LDS.64 per thread through a padded, conflict-free 512-element block
//          (what rcg_frame.h does in its merged tail passes; the other passes do the same across warps)
//   shfl : the same transpose with shuffles.  Register numbers are static, so element j of lane l cannot be sent to
//          "lane j, register l" directly: four butterfly stages (lane ^ 1, 2, 4, 8), each exchanging half of the
//          registers with SHFL and picking the kept / received halves with selects -- 4 x 16 SHFL.64-equivalents
//          (2 x 32-bit SHFL each) + selects per thread.
// Both kernels run the exchange `iters` times on 24 warps per SM (the residency of the W = 16384 kernel) and report
// cycles per exchange per warp; a checksum keeps the compiler honest and confirms both produce the same transpose.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_exchange warp_exchange.cu ; run: ./warp_exchange
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int phys(int i) { return i + (i >> 4); }

template <bool SHFL>
__global__ void __launch_bounds__(768, 1) k_exchange(float* out, long long* cycles, int iters) {
    extern __shared__ float2 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* blk = smem + warp * 544;                 // 512 + 1/16 padding
    float2 x[16];
#pragma unroll
    for (int a = 0; a < 16; a++) x[a] = make_float2((float)(threadIdx.x * 16 + a), (float)a);
    const int h = lane & 15, half = lane >> 4;       // 16 x 16 transpose inside each half warp
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if constexpr (!SHFL) {
#pragma unroll
            for (int a = 0; a < 16; a++) blk[phys(256 * half + 16 * h + a)] = x[a];     // row h of the half's 16 x 16 tile
            __syncwarp();
#pragma unroll
            for (int a = 0; a < 16; a++) x[a] = blk[phys(256 * half + 16 * a + h)];     // column h
            __syncwarp();
        } else {
            // recursive 2 x 2 block transpose: stage s swaps the off-diagonal blocks of size 2^s
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int bit = 1 << s;
                const bool upper = (h & bit) != 0;
#pragma unroll
                for (int a = 0; a < 16; a++) {
                    if ((a & bit) == 0) {
                        // lanes without the bit send x[a | bit] and keep x[a]; lanes with it send x[a] and keep x[a | bit]
                        const float2 send = upper ? x[a] : x[a | bit];
                        float2 got;
                        got.x = __shfl_xor_sync(0xffffffffu, send.x, bit);
                        got.y = __shfl_xor_sync(0xffffffffu, send.y, bit);
                        if (upper) x[a] = got; else x[a | bit] = got;
                    }
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 16; a++) x[a].x += 1.0f;   // a dependent touch so that iterations do not collapse
    }
    const long long t1 = clock64();
    float acc = 0.0f;
#pragma unroll
    for (int a = 0; a < 16; a++) acc += x[a].x * (float)(a + 1) + x[a].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    const int iters = 2000, grid = 148, threads = 768;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    const size_t smem = sizeof(float2) * 544 * (threads / 32);
    cudaFuncSetAttribute(k_exchange<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_exchange<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* h = (float*)malloc(sizeof(float) * grid * threads);
    double sums[2];
    for (int v = 0; v < 2; v++) {
        for (int rep = 0; rep < 2; rep++) {
            if (v == 0) k_exchange<false><<<grid, threads, smem>>>(out, cyc, iters);
            else k_exchange<true><<<grid, threads, smem>>>(out, cyc, iters);
            cudaDeviceSynchronize();
        }
        long long c[148];
        cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        cudaMemcpy(h, out, sizeof(float) * grid * threads, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < threads; i++) s += h[i];
        sums[v] = s;
        long long mx = 0; for (int i = 0; i < grid; i++) mx = c[i] > mx ? c[i] : mx;
        printf("%-5s 24 warps/SM: %8.1f cycles per 16 x 16 complex transpose per warp-iteration (all 24 warps together: %.1f per SM), checksum %.6e, %s\n",
               v == 0 ? "smem" : "shfl", (double)mx / iters, (double)mx / iters, s, cudaGetErrorString(cudaGetLastError()));
    }
    printf("same transpose: %s\n", sums[0] == sums[1] ? "yes" : "NO");
    return 0;
}

// tmem_ld.cu -- what a Tensor Memory load costs a thread (B200): latency of tcgen05.ld + tcgen05.wait::ld, whether the
// wait is the expensive half (k loads per wait), and the aggregate rate with 1 .. 24 warps per SM and 1 .. 3 CTAs per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu ; run: ./tmem_ld
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int LOADS_PER_WAIT>
__global__ void k_ldtm(float* out, long long* cycles, int iters, int alloc_cols) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        if (alloc_cols == 512) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t mine = slot + ((uint32_t)(32 * ((threadIdx.x / 32) % 4)) << 16);
    if (threadIdx.x < 128) {
        for (int c = 0; c < 128; c += 8) {
            uint32_t v = __float_as_uint(1.0f + 1e-6f * (float)(c + threadIdx.x));
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(mine + c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float acc = 0.0f;
    uint32_t off = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t r[LOADS_PER_WAIT][8];
#pragma unroll
        for (int l = 0; l < LOADS_PER_WAIT; l++)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[l][0]), "=r"(r[l][1]), "=r"(r[l][2]), "=r"(r[l][3]), "=r"(r[l][4]), "=r"(r[l][5]), "=r"(r[l][6]), "=r"(r[l][7])
                         : "r"(mine + ((off + 8 * l) & 127)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int l = 0; l < LOADS_PER_WAIT; l++)
#pragma unroll
            for (int q = 0; q < 8; q++) acc += __uint_as_float(r[l][q]);
        off = (off + 8 * LOADS_PER_WAIT + (__float_as_uint(acc) & 8u)) & 127u;   // the next address depends on the data
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (threadIdx.x < 32) {
        if (alloc_cols == 512) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(slot) : "memory");
    }
}

// the same access pattern out of shared memory (LDS.64 x 4 per "load") and out of L1-cached global memory
template <int LOADS_PER_WAIT, bool GLOBAL>
__global__ void k_lds(float* out, long long* cycles, int iters, const float2* gtab) {
    __shared__ float2 tab[128 * 4];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = make_float2(1.0f + 1e-6f * i, 1.0f);
    __syncthreads();
    const float2* src = GLOBAL ? gtab : tab;
    float acc = 0.0f;
    uint32_t off = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        float2 r[LOADS_PER_WAIT][4];
#pragma unroll
        for (int l = 0; l < LOADS_PER_WAIT; l++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float2* p = src + (((off + 4 * l + q) * 16 + (threadIdx.x & 15)) & 511);
                r[l][q] = GLOBAL ? __ldg(p) : *p;
            }
#pragma unroll
        for (int l = 0; l < LOADS_PER_WAIT; l++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc += r[l][q].x + r[l][q].y;
        off = (off + 4 * LOADS_PER_WAIT + (__float_as_uint(acc) & 4u)) & 31u;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc; float2* gtab;
    cudaMalloc(&out, 148 * 3 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * 3 * sizeof(long long)); cudaMalloc(&gtab, 512 * sizeof(float2));
    cudaMemset(gtab, 0, 512 * sizeof(float2));
    const int iters = 4096;
    auto report = [&](const char* what, int ctas, int threads, int loads) {
        cudaDeviceSynchronize();
        long long h[148 * 3];
        cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < ctas; i++) s += (double)h[i];
        printf("%-28s ctas/SM %d warps/CTA %2d loads/wait %d : %.1f cycles per wait group, %.1f per 8-column load (%s)\n", what,
               ctas / 148, threads / 32, loads, s / ctas / iters, s / ctas / iters / loads, cudaGetErrorString(cudaGetLastError()));
    };
    for (int ctas_per_sm = 1; ctas_per_sm <= 3; ctas_per_sm += 2)
        for (int threads : {32, 256, 768}) {
            if (ctas_per_sm * threads > 2048 || (ctas_per_sm == 3 && threads == 768)) continue;
            const int ctas = 148 * ctas_per_sm, cols = ctas_per_sm == 1 ? 512 : 128;
            k_ldtm<1><<<ctas, threads>>>(out, cyc, iters, cols); report("tmem", ctas, threads, 1);
            k_ldtm<2><<<ctas, threads>>>(out, cyc, iters, cols); report("tmem", ctas, threads, 2);
            k_ldtm<4><<<ctas, threads>>>(out, cyc, iters, cols); report("tmem", ctas, threads, 4);
            k_lds<1, false><<<ctas, threads>>>(out, cyc, iters, gtab); report("shared (4 x LDS.64)", ctas, threads, 1);
            k_lds<4, false><<<ctas, threads>>>(out, cyc, iters, gtab); report("shared (4 x LDS.64)", ctas, threads, 4);
            k_lds<1, true><<<ctas, threads>>>(out, cyc, iters, gtab); report("global L1 (4 x LDG.64)", ctas, threads, 1);
            k_lds<4, true><<<ctas, threads>>>(out, cyc, iters, gtab); report("global L1 (4 x LDG.64)", ctas, threads, 4);
        }
    return 0;
}

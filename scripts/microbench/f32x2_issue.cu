// f32x2_issue.cu -- what packed FP32 (FFMA2 / FADD2 / FMUL2, sm_100a) costs in issue slots and FMA-pipe cycles
// next to scalar FFMA, and what the operand swizzles ptxas folds in (.LO_HI swap, scalar broadcast) cost.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_issue f32x2_issue.cu ; run on a B200.
// Prints warp-instructions per clock per SM for each variant (one 512-thread CTA per SM = 4 warps per SMSP, 8 independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;

template <int V>
__global__ void __launch_bounds__(512) bench(float2* out, float2 seed, long long* cycles) {
    float2 a[CH];
    float s[CH];
    unsigned u[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i); s[i] = seed.x * i + threadIdx.x; u[i] = threadIdx.x * 2654435761u + i; }
    const float2 w = make_float2(seed.y, seed.x);
    const float2 pm = make_float2(1.0f, -1.0f);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if constexpr (V == 0) {        // 2 scalar FFMA (what one FFMA2 replaces)
                a[i].x = fmaf(a[i].x, w.x, w.y); a[i].y = fmaf(a[i].y, w.x, w.y);
            } else if constexpr (V == 1) { // FFMA2 plain
                a[i] = __ffma2_rn(a[i], w, pm);
            } else if constexpr (V == 2) { // FFMA2 with swapped operand + scalar broadcast
                a[i] = __ffma2_rn(make_float2(a[i].y, a[i].x), make_float2(w.x, w.x), pm);
            } else if constexpr (V == 3) { // FADD2
                a[i] = __fadd2_rn(a[i], w);
            } else if constexpr (V == 4) { // 2 scalar FADD
                a[i].x = __fadd_rn(a[i].x, w.x); a[i].y = __fadd_rn(a[i].y, w.y);
            } else if constexpr (V == 5) { // FFMA2 + one ALU op (LOP3) per chain: do they dual-share issue?
                a[i] = __ffma2_rn(a[i], w, pm);
                u[i] = (u[i] ^ (u[i] >> 3)) & 0x7fffffffu;
            } else if constexpr (V == 6) { // 2 FFMA + one ALU op
                a[i].x = fmaf(a[i].x, w.x, w.y); a[i].y = fmaf(a[i].y, w.x, w.y);
                u[i] = (u[i] ^ (u[i] >> 3)) & 0x7fffffffu;
            } else if constexpr (V == 7) { // mulhilo as lo = a*b, hi = umulhi  (2 IMAD)
                unsigned lo = u[i] * 0xD2511F53u, hi = __umulhi(u[i], 0xD2511F53u);
                u[i] = lo ^ hi;
            } else if constexpr (V == 8) { // mul.wide.u32 (IMAD.WIDE)
                unsigned long long p = (unsigned long long)u[i] * 0xD2511F53ull;
                u[i] = (unsigned)p ^ (unsigned)(p >> 32);
            } else if constexpr (V == 9) { // FMUL2 then FADD2 (unfused packed pair)
                a[i] = __fadd2_rn(__fmul2_rn(a[i], w), pm);
            } else if constexpr (V == 10) { // IMAD.WIDE alone (lo feeds back, hi accumulated by the next one's addend)
                unsigned long long p = (unsigned long long)u[i] * 0xD2511F53ull;
                u[i] = (unsigned)(p >> 32) + (unsigned)p;
            } else if constexpr (V == 11) { // 32-bit IMAD alone
                u[i] = u[i] * 0xD2511F53u + 12345u;
            } else if constexpr (V == 12) { // IMAD.HI alone
                u[i] = __umulhi(u[i], 0xD2511F53u) + 77u;
            } else if constexpr (V == 13) { // IMAD.WIDE + 2 FFMA2 on other chains: shared pipe?
                unsigned long long p = (unsigned long long)u[i] * 0xD2511F53ull;
                u[i] = (unsigned)(p >> 32) ^ (unsigned)p;
                a[i] = __ffma2_rn(a[i], w, pm);
                a[i] = __ffma2_rn(a[i], w, pm);
            } else if constexpr (V == 14) { // 2 FFMA2 alone (reference for V == 13)
                a[i] = __ffma2_rn(a[i], w, pm);
                a[i] = __ffma2_rn(a[i], w, pm);
            } else if constexpr (V == 15) { // MUFU.SIN + MUFU.COS
                a[i].x = __sinf(a[i].x); a[i].y = __cosf(a[i].y);
            }
        }
    }
    long long t1 = clock64();
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CH; i++) { r.x += a[i].x + s[i]; r.y += a[i].y + (float)u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int inst_per_chain_iter) {
    int dev_sms = 148;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); dev_sms = p.multiProcessorCount;
    float2* out; long long* cyc;
    const int threads = 512, blocks = dev_sms;   // one CTA per SM: 16 warps per SM = 4 per SMSP
    cudaMalloc(&out, sizeof(float2) * threads * blocks); cudaMalloc(&cyc, sizeof(long long) * blocks);
    bench<V><<<blocks, threads>>>(out, make_float2(1.0001f, 0.5f), cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<V><<<blocks, threads>>>(out, make_float2(1.0001f, 0.5f), cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double warp_inst = (double)ITERS * CH * inst_per_chain_iter * (threads / 32);   // per SM
    printf("%-44s %8.3f ms  %9lld cyc  %6.3f warp-inst/clk/SM (of 4.0 issue)\n", name, ms, h[0], warp_inst / (double)h[0]);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("2x FFMA (scalar)", 2);
    run<1>("FFMA2", 1);
    run<2>("FFMA2 .LO_HI + scalar broadcast", 1);
    run<4>("2x FADD (scalar)", 2);
    run<3>("FADD2", 1);
    run<9>("FMUL2 + FADD2", 2);
    run<6>("2x FFMA + 3 ALU", 5);
    run<5>("FFMA2 + 3 ALU", 4);
    run<7>("IMAD.LO + IMAD.HI + LOP3", 3);
    run<8>("IMAD.WIDE + LOP3", 2);
    run<10>("IMAD.WIDE + IADD", 2);
    run<11>("IMAD (32-bit)", 1);
    run<12>("IMAD.HI + IADD", 2);
    run<14>("2x FFMA2", 2);
    run<13>("IMAD.WIDE + LOP3 + 2x FFMA2", 4);
    run<15>("MUFU.SIN + MUFU.COS (+range ops)", 2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

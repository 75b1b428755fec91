// rcg_aux.cu -- small kernels beside the stretch path:
//  * fade in / out of the reference's live playback (`-x`; mixer.rs:132-190, math.rs:33-40),
//    off in file mode like the reference (main.rs:195-204) unless asked for;
//  * an FP32 FMA microbenchmark: the measured denominator for the path's compute roofline.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "../../include/rocoder_b200.h"

namespace {

// math::sqrt_interp (math.rs:33-40), same operation order
__device__ __forceinline__ float sqrt_interp(float start, float end, float ratio) {
    const bool increasing = start < end;
    const float progress = __fmul_rn(__fsub_rn(__fmul_rn(ratio, 2.0f), 1.0f), increasing ? 1.0f : -1.0f);
    const float factor = __fsqrt_rn(__fmul_rn(0.5f, __fadd_rn(1.0f, fmaxf(progress, -1.0f))));
    const float abs_interval = fabsf(__fsub_rn(end, start));
    return __fadd_rn(increasing ? start : end, __fmul_rn(abs_interval, factor));
}

// Layer::current_amp (mixer.rs:132-145) over keyframes (0, 0), (fade_n, 1), (start_n, 1), (end_n, 0)
__device__ __forceinline__ float fade_gain(unsigned long long pos, unsigned long long fade_n, unsigned long long start_n,
                                           unsigned long long end_n) {
    const unsigned long long kp[4] = {0ull, fade_n, start_n, end_n};
    const float kv[4] = {0.0f, 1.0f, 1.0f, 0.0f};
    int i = 0;
    while (i < 2 && kp[i + 1] < pos) i++;
    if (pos > kp[3]) return kv[3];
    if (kp[i + 1] == kp[i]) return kv[i + 1];
    const float progress = __fdiv_rn((float)(pos - kp[i]), (float)(kp[i + 1] - kp[i]));
    return sqrt_interp(kv[i], kv[i + 1], progress);
}

__global__ void fade_kernel(float* audio, long long stride, int channels, unsigned long long n, unsigned long long first_pos,
                            unsigned long long fade_n, unsigned long long start_n, unsigned long long end_n) {
    const unsigned long long total = n * (unsigned long long)channels;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long c = i / n, k = i - c * n;
        float* p = audio + (long long)c * stride + k;
        *p = __fmul_rn(*p, fade_gain(first_pos + k, fade_n, start_n, end_n));   // sample * current_amp
    }
}

__global__ void fma_peak_kernel(float* sink, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.0f, a2 = a0 + 2.0f, a3 = a0 + 3.0f, a4 = a0 + 4.0f, a5 = a0 + 5.0f,
          a6 = a0 + 6.0f, a7 = a0 + 7.0f;
    const float b = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123.456f) sink[0] = a0;
}

uint64_t f32_to_usize(float v) {
    if (!(v == v) || v <= 0.0f) return 0;
    if (v >= 18446744073709551615.0f) return UINT64_MAX;
    return (uint64_t)v;
}

}  // namespace

// keyframe positions of Mixer::fade_in_out (mixer.rs:179-190) for expected_total_samples output samples
extern "C" int rcg_fade_keyframes(float fade_s, uint32_t sample_rate, uint64_t expected_total_samples, uint64_t* fade_n,
                                  uint64_t* start_n, uint64_t* end_n) {
    if (!fade_n || !start_n || !end_n || sample_rate == 0) return RCG_ERR_INVALID_ARG;
    const float total_s = (float)expected_total_samples / (float)sample_rate;
    // Duration holds integer nanoseconds: from_secs_f32 rounds to them, the subtraction is exact
    const double total_ns = std::floor((double)total_s * 1e9 + 0.5), fade_ns = std::floor((double)fade_s * 1e9 + 0.5);
    double start_ns = total_ns - fade_ns;
    if (start_ns < 0) start_ns = 0;   // the reference panics on Duration underflow
    auto dur_to_sample = [&](double ns) { return f32_to_usize((float)(ns / 1e9) * (float)sample_rate); };  // mixer.rs:147-149
    *fade_n = dur_to_sample(fade_ns);
    *start_n = dur_to_sample(start_ns);
    *end_n = dur_to_sample(start_ns + fade_ns);
    return RCG_OK;
}

extern "C" int rcg_fade_in_out(float* d_audio, int64_t stride, uint32_t channels, uint64_t n, uint64_t first_pos,
                               float fade_s, uint32_t sample_rate, uint64_t expected_total_samples, void* stream) {
    if (!d_audio || channels == 0) return RCG_ERR_INVALID_ARG;
    uint64_t fade_n, start_n, end_n;
    int rc = rcg_fade_keyframes(fade_s, sample_rate, expected_total_samples, &fade_n, &start_n, &end_n);
    if (rc) return rc;
    if (n == 0) return RCG_OK;
    const unsigned long long total = n * channels;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148ull * 16 ? (total + 255) / 256 : 148ull * 16);
    fade_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_audio, stride, (int)channels, n, first_pos, fade_n, start_n, end_n);
    return cudaGetLastError() == cudaSuccess ? RCG_OK : RCG_ERR_CUDA;
}

// FP32 FMA throughput of the current device in TFLOP/s (2 flops per FMA), best of `reps`
extern "C" int rcg_measure_fp32_peak(double* tflops, int reps) {
    if (!tflops) return RCG_ERR_INVALID_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RCG_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* sink = nullptr;
    if (cudaMalloc(&sink, 4) != cudaSuccess) return RCG_ERR_CUDA;
    const int iters = 2048, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < (reps > 0 ? reps : 5) + 1; r++) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<blocks, threads>>>(sink, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        if (r > 0 && ms > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? RCG_OK : RCG_ERR_CUDA;
}

// ---- NUMA placement of the calling thread ---------------------------------------------------------------
// Pinned staging buffers are allocated on the NUMA node the calling thread runs on; on a multi-socket box a
// buffer on the far socket halves the host<->device rate of every copy through it.  This pins the calling
// thread (and hence its later rcg_host_alloc / cudaMallocHost pages) to the CPUs of the node the GPU hangs off.
#include <sched.h>

#include <cctype>
#include <cstdio>
#include <cstring>

static bool read_line(const char* path, char* buf, size_t cap) {
    FILE* f = fopen(path, "r");
    if (!f) return false;
    const bool ok = fgets(buf, (int)cap, f) != nullptr;
    fclose(f);
    return ok;
}

extern "C" int rcg_bind_host_thread(int device, int* numa_node) {
    if (numa_node) *numa_node = -1;
    char bus[64] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        (void)cudaGetLastError();
        return RCG_ERR_NO_DEVICE;
    }
    for (char* c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
    char path[160], line[4096];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (!read_line(path, line, sizeof line) || sscanf(line, "%d", &node) != 1 || node < 0) return RCG_OK;  // no NUMA info: leave as is
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    if (!read_line(path, line, sizeof line)) return RCG_OK;
    cpu_set_t allowed, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return RCG_OK;
    int n_want = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {   // "0-31,64-95"
        int a = 0, b = 0;
        const int got = sscanf(tok, "%d-%d", &a, &b);
        if (got < 1) continue;
        if (got == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); n_want++; }   // never widen a cgroup / taskset mask
    }
    if (n_want == 0) return RCG_OK;
    if (sched_setaffinity(0, sizeof want, &want) != 0) return RCG_OK;
    if (numa_node) *numa_node = node;
    return RCG_OK;
}

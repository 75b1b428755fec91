// rcg_big.cu -- kernels of the HBM-resident chirp-z path for long windows (rcg_big.h).  Its own translation unit.
#include <algorithm>

#include "rcg_kernels.h"

#include "rcg_big.h"

namespace rcg {

namespace {

constexpr int kBigThreads = 256;
constexpr int kTile = 1 << kBigTileLog2;

__device__ __forceinline__ cf cmul(cf a, cf b) { return make_cf(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cf cmul_conj(cf a, cf b) { return make_cf(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a conj(b)
// exp(-2 pi i j / n), n a power of two, j < n: the argument of sincospif is exact
__device__ __forceinline__ cf unit_root(long long j, int log2n) {
    float s, c;
    sincospif(-(float)j * exp2f((float)(1 - log2n)), &s, &c);
    return make_cf(c, s);
}

// a[n] = (s[n] w[n]) conj(c[n]), zero padded to M (fft.rs:51-58); bypass hook: y[n] = (s[n] w[n]) w[n] straight
// into the frame buffer
__global__ void __launch_bounds__(kBigThreads)
big_load_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ BigTables bt, int channel,
                long long frame0, int n_frames) {
    const int W = bt.w, lm = bt.lm;
    const long long M = 1LL << lm, total = (long long)n_frames << lm;
    const float* in = prm.in + (long long)channel * prm.in_stride;
    const bool bypass = prm.flags & kFlagBypass;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i >> lm;
        const int n = (int)(i & (M - 1));
        cf v = make_cf(0.0f, 0.0f);
        if (n < W) {
            const long long src = (frame0 + f) * (long long)prm.step - prm.in_origin + n;
            const float s = (src >= 0 && src < prm.in_valid) ? ld_stream(in + src) : 0.0f;
            const float w = prm.window[n];
            const float x = RCG_MUL(s, w);
            if (bypass) {
                bt.ybuf[f * W + n] = RCG_MUL(x, w);
            } else {
                const cf c = ld_cf(bt.chirp + n);
                v = make_cf(x * c.x, -(x * c.y));
            }
        }
        if (!bypass) bt.work[i] = v;
    }
}

// One in-place radix-R pass over every sub-transform of length 2^ln in the workspace (forward: DIF, twiddles on the
// outputs; inverse: DIT, conjugate twiddles on the inputs).  Thread i owns the butterfly (sub-transform, b):
// consecutive threads, consecutive b -- every access is coalesced.
template <int RL, bool INV>
__global__ void __launch_bounds__(kBigThreads)
big_pass_kernel(cf* __restrict__ data, int ln, long long n_butterflies) {
    constexpr int R = 1 << RL;
    const int lq = ln - RL;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_butterflies; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i & ((1LL << lq) - 1);
        cf* p = data + ((i >> lq) << ln) + b;
        cf x[R];
#pragma unroll
        for (int a = 0; a < R; a++) x[a] = p[(long long)a << lq];
        if constexpr (INV) {
#pragma unroll
            for (int j = 1; j < R; j++) x[j] = cmul_conj(x[j], unit_root(b * j, ln));
            Dft<R, true>::run(x);
        } else {
            Dft<R, false>::run(x);
#pragma unroll
            for (int j = 1; j < R; j++) x[j] = cmul(x[j], unit_root(b * j, ln));
        }
#pragma unroll
        for (int a = 0; a < R; a++) p[(long long)a << lq] = x[a];
    }
}

// The middle of a convolution on one 4096-point tile in shared memory: radix-2 DIF stages (natural -> bit-reversed),
// the product with the filter spectrum (stored in the order the forward passes leave), radix-2 DIT stages back.
__global__ void __launch_bounds__(kBigThreads)
big_tile_kernel(cf* __restrict__ data, const cf* __restrict__ filter, int lm, long long n_tiles) {
    __shared__ cf S[kTile];
    __shared__ cf tw[kTile / 2];
    const int t = threadIdx.x;
    for (int j = t; j < kTile / 2; j += kBigThreads) tw[j] = unit_root(j, kBigTileLog2);
    const long long mask = (1LL << lm) - 1;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        cf* g = data + tile * kTile;
        __syncthreads();
        for (int i = t; i < kTile; i += kBigThreads) S[i] = g[i];
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < kBigTileLog2; s++) {
            const int lh = kBigTileLog2 - 1 - s, half = 1 << lh;
            for (int bf = t; bf < kTile / 2; bf += kBigThreads) {
                const int j = bf & (half - 1), i0 = ((bf >> lh) << (lh + 1)) + j;
                const cf a = S[i0], b = S[i0 + half];
                S[i0] = make_cf(a.x + b.x, a.y + b.y);
                S[i0 + half] = cmul(make_cf(a.x - b.x, a.y - b.y), tw[j << s]);
            }
            __syncthreads();
        }
        const cf* fl = filter + ((tile * kTile) & mask);
        for (int i = t; i < kTile; i += kBigThreads) S[i] = cmul(S[i], ld_cf(fl + i));
        __syncthreads();
#pragma unroll 1
        for (int s = kBigTileLog2 - 1; s >= 0; s--) {
            const int lh = kBigTileLog2 - 1 - s, half = 1 << lh;
            for (int bf = t; bf < kTile / 2; bf += kBigThreads) {
                const int j = bf & (half - 1), i0 = ((bf >> lh) << (lh + 1)) + j;
                const cf a = S[i0], b = cmul_conj(S[i0 + half], tw[j << s]);
                S[i0] = make_cf(a.x + b.x, a.y + b.y);
                S[i0 + half] = make_cf(a.x - b.x, a.y - b.y);
            }
            __syncthreads();
        }
        for (int i = t; i < kTile; i += kBigThreads) g[i] = S[i];
    }
}

// X[k] = conj(c[k]) conv[k]; only |X[k]| = |conv[k]| survives (fft.rs:67).  Y[k] = |X[k]| e^{i phi_k}; the inverse
// transform runs on conj(Y): a'[k] = conj(Y[k]) conj(c[k]), zero padded.  Phase source: the Philox block / word the
// power-of-two kernels give bin k (rcg_frame_bluestein.h).
__global__ void __launch_bounds__(kBigThreads)
big_spectral_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ BigTables bt, int channel,
                    long long frame0, int n_frames) {
    const int W = bt.w, lm = bt.lm, H = W / 2, Q = H / 2;
    const long long M = 1LL << lm, total = (long long)n_frames << lm;
    const uint32_t chan = (uint32_t)(channel + prm.channel_offset);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i & (M - 1));
        cf v = make_cf(0.0f, 0.0f);
        if (k < W) {
            const unsigned long long frame = (unsigned long long)(frame0 + (i >> lm));
            const cf x = bt.work[i];
            const float mag = fast_sqrt(x.x * x.x + x.y * x.y);
            uint32_t pair, word;
            if (k <= Q) { pair = (uint32_t)k; word = 0; }
            else if (k <= H) { pair = (uint32_t)(H - k); word = 2; }
            else if (k < H + Q) { pair = (uint32_t)(k - H); word = 3; }
            else { pair = (uint32_t)(W - k); word = 1; }
            uint32_t u[4];
            philox4x32_10(pair, (uint32_t)frame, (uint32_t)(frame >> 32), chan, prm.keys, u);
            const cf e = cis_phase(u[word]);
            const cf c = ld_cf(bt.chirp + k);
            v = make_cf(mag * (e.x * c.x - e.y * c.y), -(mag * (e.x * c.y + e.y * c.x)));
        }
        bt.work[i] = v;
    }
}

// y[n] = Re(c[n] conj(conv[n])) / W * w[n]  (fft.rs:70-73)
__global__ void __launch_bounds__(kBigThreads)
big_finish_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ BigTables bt, int n_frames) {
    const int W = bt.w, lm = bt.lm;
    const float wl = (float)W;
    const long long total = (long long)n_frames * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / W;
        const int n = (int)(i - f * W);
        const cf v = bt.work[(f << lm) + n], c = ld_cf(bt.chirp + n);
        const float re = v.x * c.x + v.y * c.y;
        bt.ybuf[i] = RCG_MUL(RCG_DIV(re, wl), prm.window[n]);
    }
}

// z[i] of block `frame`: ((y[i] + carried[i]) * env[i]) * amp  (stretcher.rs:96-100)
struct BlockView {
    const float* y;      // first half of the block's own frame
    const float* prev;   // second half of the frame before it
    const float* env;
    float amp;
    __device__ __forceinline__ float z(int i) const { return RCG_MUL(RCG_MUL(RCG_ADD(y[i], prev[i]), env[i]), amp); }
};
// integer resample of one finished block straight to the output (resampler.rs:3-35; index rules of phase_resample)
__device__ __forceinline__ void big_emit_block(const BlockView& b, const StretchParams& prm, int H, int channel,
                                               long long frame) {
    float* out = prm.out + (long long)channel * prm.out_stride;
    const int p = prm.pitch, t = threadIdx.x, nt = blockDim.x;
    if (p == 1) {
        for (int i = t; i < H; i += nt) out[frame * (long long)H + i - prm.out_origin] = b.z(i);
    } else if (p > 1) {
        const long long g0 = frame * (long long)H;
        const long long o_begin = (g0 + p - 1) / p, o_end = (g0 + H + p - 1) / p;
        for (long long o = o_begin + t; o < o_end; o += nt) out[o - prm.out_origin] = b.z((int)(o * p - g0));
    } else {
        const int q = -p;
        const int n = (prm.samples_needed - 1) * q;
        float* dst = out + (frame * (long long)n - prm.out_origin);
        for (int tt = t; tt < n; tt += nt) {
            const int i = tt / q, j = tt - i * q;
            const float cur = b.z(i), nxt = b.z(i + 1);
            dst[tt] = RCG_ADD(cur, RCG_MUL(RCG_SUB(nxt, cur), prm.lerp_ratio[j]));
        }
    }
}

// One CTA per frame of the batch.  Frame `frame0 + blockIdx.x` gives block f its first half; the carried half window
// is the second half of the frame before it: in the batch, or (first frame of the batch) the channel's carry buffer.
// skip_before: frames below it are halo frames (their block is not emitted); head_frame: the block whose raw first
// half goes to head_out instead (deferred first block of a time shard), -1 if none.
__global__ void __launch_bounds__(kBigThreads)
big_emit_kernel(const __grid_constant__ StretchParams prm, const __grid_constant__ BigTables bt, int channel,
                long long frame0, long long skip_before, long long head_frame) {
    const int W = bt.w, H = W / 2;
    const long long frame = frame0 + blockIdx.x;
    if (frame < skip_before) return;
    const float* y = bt.ybuf + (long long)blockIdx.x * W;
    if (frame == head_frame) {
        for (int i = threadIdx.x; i < H; i += blockDim.x) prm.head_out[(long long)channel * H + i] = y[i];
        return;
    }
    const BlockView b{y, blockIdx.x == 0 ? bt.carry + (long long)channel * H : y - H, prm.env, prm.amp};
    big_emit_block(b, prm, H, channel, frame);
}

__global__ void __launch_bounds__(kBigThreads)
big_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {   // src == nullptr: zeros
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src ? src[i] : 0.0f;
}

// boundary block of a time shard (rcg_finish_boundary): ((head + tail_prev) * env) * amp, resampled
__global__ void __launch_bounds__(kBigThreads)
big_boundary_kernel(const __grid_constant__ StretchParams prm, int W, const float* head, const float* tail_prev) {
    const int H = W / 2, channel = blockIdx.x;
    const BlockView b{head + (long long)channel * H, tail_prev + (long long)channel * H, prm.env, prm.amp};
    big_emit_block(b, prm, H, channel, prm.frame_begin);
}

int grid_for(long long work_items, int per_thread = 4) {
    const long long blocks = (work_items + (long long)kBigThreads * per_thread - 1) / ((long long)kBigThreads * per_thread);
    return (int)std::max<long long>(1, std::min<long long>(blocks, 148 * 16));
}

template <bool INV>
cudaError_t top_pass(cf* work, int rl, int ln, long long n_elems, cudaStream_t stream) {
    const long long nb = n_elems >> rl;
    const int grid = grid_for(nb, 1);
    switch (rl) {
        case 1: big_pass_kernel<1, INV><<<grid, kBigThreads, 0, stream>>>(work, ln, nb); break;
        case 2: big_pass_kernel<2, INV><<<grid, kBigThreads, 0, stream>>>(work, ln, nb); break;
        case 3: big_pass_kernel<3, INV><<<grid, kBigThreads, 0, stream>>>(work, ln, nb); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// circular convolution of every M-point row of the workspace with the chirp filter
cudaError_t convolve(const BigTables& bt, int n_frames, cudaStream_t stream) {
    const long long n_elems = (long long)n_frames << bt.lm;
    int ln = bt.lm;
    for (int p = 0; p < bt.n_top; p++) {
        if (cudaError_t e = top_pass<false>(bt.work, bt.top_rlog2[p], ln, n_elems, stream)) return e;
        ln -= bt.top_rlog2[p];
    }
    const long long n_tiles = n_elems >> kBigTileLog2;
    big_tile_kernel<<<(int)std::min<long long>(n_tiles, 148 * 4), kBigThreads, 0, stream>>>(bt.work, bt.filter, bt.lm, n_tiles);
    if (cudaError_t e = cudaGetLastError()) return e;
    for (int p = bt.n_top - 1; p >= 0; p--) {
        ln += bt.top_rlog2[p];
        if (cudaError_t e = top_pass<true>(bt.work, bt.top_rlog2[p], ln, n_elems, stream)) return e;
    }
    return cudaSuccess;
}

}  // namespace

// Frames [frame_begin, frame_end) of every channel; the roles of tail_in / tail_out / head_out / kFlagDeferFirst and of
// the halo frame are those of run_work_item (rcg_frame.h).  prm.run_len / n_runs / tail_scratch are not used.
cudaError_t launch_big(const StretchParams& prm, const BigTables& bt, cudaStream_t stream) {
    const int W = bt.w, H = W / 2;
    const long long f0 = prm.frame_begin, f1 = prm.frame_end;
    if (f1 <= f0) return cudaSuccess;
    const bool defer = (prm.flags & kFlagDeferFirst) && f0 > 0;
    const bool have_tail_in = prm.tail_in != nullptr && f0 > 0;
    const bool halo = f0 > 0 && !defer && !have_tail_in;
    const bool bypass = prm.flags & kFlagBypass;
    for (int ch = 0; ch < prm.channels; ch++) {
        float* carry = bt.carry + (long long)ch * H;
        // the half window carried into block f0: zeros (f0 == 0, stretcher.rs:58-59; also the unused one of a deferred
        // block), tail_in, or -- below -- the second half of the recomputed frame f0 - 1
        if (!halo) {
            big_copy_kernel<<<grid_for(H), kBigThreads, 0, stream>>>(carry, have_tail_in ? prm.tail_in + (long long)ch * H : nullptr, H);
            if (cudaError_t e = cudaGetLastError()) return e;
        }
        for (long long fb = halo ? f0 - 1 : f0; fb < f1; fb += bt.batch) {
            const int nb = (int)std::min<long long>(bt.batch, f1 - fb);
            big_load_kernel<<<grid_for((long long)nb << bt.lm), kBigThreads, 0, stream>>>(prm, bt, ch, fb, nb);
            if (cudaError_t e = cudaGetLastError()) return e;
            if (!bypass) {
                if (cudaError_t e = convolve(bt, nb, stream)) return e;
                big_spectral_kernel<<<grid_for((long long)nb << bt.lm), kBigThreads, 0, stream>>>(prm, bt, ch, fb, nb);
                if (cudaError_t e = cudaGetLastError()) return e;
                if (cudaError_t e = convolve(bt, nb, stream)) return e;
                big_finish_kernel<<<grid_for((long long)nb * W), kBigThreads, 0, stream>>>(prm, bt, nb);
                if (cudaError_t e = cudaGetLastError()) return e;
            }
            big_emit_kernel<<<nb, kBigThreads, 0, stream>>>(prm, bt, ch, fb, f0, defer ? f0 : -1);
            if (cudaError_t e = cudaGetLastError()) return e;
            // the next batch (and tail_out) continue from the second half of this batch's last frame
            big_copy_kernel<<<grid_for(H), kBigThreads, 0, stream>>>(carry, bt.ybuf + (long long)(nb - 1) * W + H, H);
            if (cudaError_t e = cudaGetLastError()) return e;
        }
        if (prm.tail_out != nullptr) {
            big_copy_kernel<<<grid_for(H), kBigThreads, 0, stream>>>(prm.tail_out + (long long)ch * H, carry, H);
            if (cudaError_t e = cudaGetLastError()) return e;
        }
    }
    return cudaSuccess;
}

cudaError_t launch_boundary_big(const StretchParams& prm, const BigTables& bt, const float* head, const float* tail_prev,
                                cudaStream_t stream) {
    big_boundary_kernel<<<prm.channels, kBigThreads, 0, stream>>>(prm, bt.w, head, tail_prev);
    return cudaGetLastError();
}

}  // namespace rcg

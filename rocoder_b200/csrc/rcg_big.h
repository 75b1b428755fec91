// rcg_big.h -- windows too long for one SM (or one CTA pair): even lengths that are not a power of two above 8192,
// and powers of two above 65536, up to 2^20 samples (reference README.md:100: any window length is legal; fft.rs:27-29
// plans whatever length it is given).
//
// Same mathematics as rcg_frame_bluestein.h -- both W-point transforms of a frame as chirp-z convolutions of length
// M = 2^LM >= 2W - 1, window, /W, overlap-add, envelope, amplitude and resample in the reference's order -- but the
// M-point sequences live in HBM and a launch works on a BATCH of frames of one channel at a time:
//     load * window * conj(chirp)            one pass over [batch][M]
//     top DIF passes (radix 8/4/2, in place) until the sub-transforms are 4096 points long
//     tile pass: 4096-point DIF stages, product with the filter spectrum, 4096-point DIT stages -- in shared memory
//     top DIT passes
//     |X| e^{i phi} -> conj(Y) conj(chirp)   one pass
//     the same convolution again
//     Re(chirp conj(conv)) / W * window      one pass, into [batch][W] real frames
//     emit: overlap-add of neighbouring frames of the batch (gather form: block f = first half of frame f + second
//           half of frame f - 1), envelope * amp, integer resample, straight to the output
// The forward passes leave the spectrum in mixed-radix digit-reversed order; the filter spectrum is stored in that
// order and the inverse passes undo it, so no reordering pass exists (the scheme of the in-SM kernels).
// HBM-bound by construction (about 25 transfers of the batch per frame); a correctness path for every legal window.
#pragma once
#include "rcg_frame.h"

namespace rcg {

constexpr int kBigTileLog2 = 12;          // sub-transform length finished in shared memory
constexpr int kBigMaxLog2M = 21;          // M <= 2^21, i.e. W <= 2^20
constexpr int kBigMaxTopPasses = 4;

struct BigTables {
    const cf* chirp;     // [W]  c[n] = exp(i pi n^2 / W)
    const cf* filter;    // [M]  FFT_M of the chirp filter / M, in the position order the forward passes leave
    cf* work;            // [batch][M] convolution workspace
    float* ybuf;         // [batch][W] resynthesised frames: (Re(IFFT) / W) * window  (fft.rs:70-73)
    float* carry;        // [channels][W/2] second half of the last frame seen (stretcher.rs:101-102)
    int w, lm, batch;
    int n_top;                           // global passes before the tile pass
    int top_rlog2[kBigMaxTopPasses];     // their radices (log2), outermost first
};

// radices of the global passes for convolution length 2^lm: greedy radix 8, then 4 / 2 for what is left
inline int big_top_plan(int lm, int (&rlog2)[kBigMaxTopPasses]) {
    int left = lm - kBigTileLog2, n = 0;
    while (left > 0) {
        const int r = left >= 3 ? 3 : left;
        rlog2[n++] = r;
        left -= r;
    }
    return n;
}
// bin index held at position `pos` after the forward passes (digits of the global passes, then bit reversal inside
// the tile): position digit d_1 (most significant) is the least significant digit of the bin
inline long long big_pos_to_bin(int lm, long long pos) {
    int rl[kBigMaxTopPasses];
    const int n_top = big_top_plan(lm, rl);
    long long k = 0;
    int used = 0, shift = lm;
    for (int p = 0; p < n_top; p++) {
        shift -= rl[p];
        k |= ((pos >> shift) & ((1LL << rl[p]) - 1)) << used;
        used += rl[p];
    }
    for (int b = 0; b < shift; b++) k |= ((pos >> (shift - 1 - b)) & 1LL) << (used + b);
    return k;
}

cudaError_t launch_big(const StretchParams& prm, const BigTables& bt, cudaStream_t stream);
cudaError_t launch_boundary_big(const StretchParams& prm, const BigTables& bt, const float* head, const float* tail_prev,
                                cudaStream_t stream);

}  // namespace rcg

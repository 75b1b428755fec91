/*
 * rocoder_oracle.h -- CPU restatement of ajyoon/rocoder's stretch path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * build, load or call it.  The product (rocoder_b200/) never links or imports it.
 *
 * PARITY STATUS: "parity unpinned" for the FFT/phase/OLA stages -- the reference is a
 * Rust crate (no cargo/rustc in this image, third-party crates rustfft 6.3.0,
 * rand 0.8.5, num-complex 0.4.6, slice-deque 0.3.0 are not vendored) and its own tests
 * hold no golden vectors for fft.rs or Stretcher::next_window.  What IS pinned, and
 * checked in tests/test_oracle_golden.py: windows::hanning(32) (windows.rs:30-43),
 * windows::inverse (windows.rs:53-56), resample p=1 / resample_faster
 * (resampler.rs:43-54), math::lerp (math.rs:68-80), sqrt_interp samples
 * (audio.rs:290-313) and ensure_input_samples_available (stretcher.rs:145-161).
 * The FFT stage is cross-checked against numpy/scipy pocketfft in fp64 and the whole
 * pipeline against an independent numpy restatement (oracle/np_oracle.py).
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it
 * follows.  Two builds of the numeric core exist: `_f32` is the restatement (all
 * arithmetic in f32 in the reference's order of operations); `_f64` is the arbiter
 * (same tables, same f32 phases, arithmetic in double) used to bound rounding error.
 */
#ifndef ROCODER_ORACLE_H
#define ROCODER_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- helpers pinned by reference golden vectors -------------------------------- */

/* windows.rs:4-9 */
void ro_hanning(float *out, size_t len);
/* windows.rs:12-14 */
void ro_rectangular(float *out, size_t len);
/* windows.rs:19-21 */
void ro_inverse(const float *in, float *out, size_t len);
/* crossfade.rs:4-10 */
void ro_hanning_crossfade_compensation(float *out, size_t len);
/* math.rs:28-30 */
float ro_lerp(float start, float end, float ratio);
/* math.rs:33-40 */
float ro_sqrt_interp(float start, float end, float ratio);
/* resampler.rs:3-35.  Returns 0, or -1 for the factors on which the reference panics
 * (0 and -1).  *n_out receives the produced length: n (p=1), ceil(n/p) (p>1),
 * (n-1)*|p| (p<-1).  `out` must hold at least ro_resample_len(n, p) floats. */
int ro_resample(const float *in, size_t n, int factor, float *out, size_t *n_out);
size_t ro_resample_len(size_t n, int factor);

/* ---- deterministic phase source (replaces rand::thread_rng at fft.rs:64) --------- *
 * The reference draws every phase from an OS-seeded ChaCha12 stream, so no two runs of
 * the reference agree.  Both this oracle and the GPU engine instead use Philox4x32-10
 * keyed by the 64-bit seed with counter (pair, frame_lo, frame_hi, channel); one
 * Philox block serves the four bins {pair, W-pair, H-pair, H+pair}, H = W/2.
 * u32 -> phase is rand 0.8.5's UniformFloat<f32>::sample_single for 0.0..PI:
 * ((u >> 9) * 2^-23) * PI_f32 + 0.0   (fft.rs:13 TWO_PI is PI; fft.rs:67). */
void ro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float ro_phase_from_u32(uint32_t u);
/* phase of `bin` (0..W-1) of `frame` of `channel`; W must be even. */
float ro_phase(uint64_t seed, uint32_t channel, uint64_t frame, uint32_t bin, uint32_t window_len);
/* fill phases[0..W) for one frame (one Philox call per pair) */
void ro_phase_fill(uint64_t seed, uint32_t channel, uint64_t frame, uint32_t window_len,
                   float *phases);

/* ---- FFT (stands in for rustfft at fft.rs:27-29, 59, 69) ------------------------- *
 * In-place complex transform, interleaved re/im, unnormalised, natural order.
 * inverse=0: e^{-2 pi i nk/N}; inverse=1: e^{+2 pi i nk/N}.  Any n (power of two:
 * Stockham radix-4/2; otherwise an O(n^2) DFT accumulated in double). */
void ro_fft_f32(float *re_im, size_t n, int inverse);
void ro_fft_f64(double *re_im, size_t n, int inverse);
/* Timed CPU baseline only: route every f32 transform of this library through `hook` (NULL restores the built-in
 * one).  bench.py passes scipy's pocketfft so that the baseline's FFT is a SIMD library like the reference's. */
typedef void (*ro_fft_hook_f32)(float *re_im, size_t n, int inverse);
void ro_set_fft_hook_f32(ro_fft_hook_f32 hook);

/* ---- plugin ABI (C form of fft.rs:93 `apply`) ------------------------------------ *
 * in/out: n_bins interleaved (re, im) pairs.  Return 0 on success; non-zero stands for
 * the reference's panic (fft.rs:101-106): the newest kernel is popped and the call is
 * retried with the previous one, else the spectrum passes through. */
typedef int (*ro_apply_fn)(size_t elapsed_ms, const float *in_re_im, size_t n_bins,
                           float *out_re_im);

/* ---- Stretcher (stretcher.rs) ---------------------------------------------------- */
enum {
    RO_PHASE_PHILOX = 0, /* ro_phase(seed, channel, frame, bin) */
    RO_PHASE_TABLE = 1,  /* caller-provided phases, frame-major [frame][W] */
    RO_PHASE_BYPASS = 2  /* test hook: resynth(s)[n] = (s[n]*w[n])*w[n], no FFT */
};

typedef struct ro_stretcher ro_stretcher;

/* stretcher.rs:30-76.  `window` is copied (len = window_len).  Returns NULL when
 * pitch_multiple == 0 (assert at stretcher.rs:40) or out of i8 range. precision: 32|64 */
ro_stretcher *ro_stretcher_new(uint32_t channels, uint32_t sample_rate, float factor,
                               float amplitude, int pitch_multiple, const float *window,
                               size_t window_len, float buffer_dur_s, int precision);
void ro_stretcher_free(ro_stretcher *s);
void ro_stretcher_set_phase_philox(ro_stretcher *s, uint64_t seed, uint32_t channel);
void ro_stretcher_set_phase_table(ro_stretcher *s, const float *phases, size_t n_frames);
void ro_stretcher_set_phase_bypass(ro_stretcher *s);
/* start resynth frame numbering at `frame` (time-sharded runs) */
void ro_stretcher_set_first_frame(ro_stretcher *s, uint64_t frame);
/* push a plugin onto the kernel stack (fft.rs:78-81) */
void ro_stretcher_push_plugin(ro_stretcher *s, ro_apply_fn fn);
/* Receiver<Vec<f32>> side: queue a chunk / drop the sender */
void ro_stretcher_send(ro_stretcher *s, const float *chunk, size_t n);
void ro_stretcher_close(ro_stretcher *s);
/* stretcher.rs:78 */
int ro_stretcher_is_done(const ro_stretcher *s);
/* stretcher.rs:82-85 */
size_t ro_stretcher_channel_bound(const ro_stretcher *s);
/* stretcher.rs:123-135.  Returns 0, or -2 if it would block (no chunk queued and the
 * sender not dropped). */
int ro_stretcher_ensure_input_samples_available(ro_stretcher *s, size_t n);
size_t ro_stretcher_input_buf_len(const ro_stretcher *s);
void ro_stretcher_input_buf_copy(const ro_stretcher *s, float *out);
/* stretcher.rs:87-121.  Output as f32 (precision 32) -- or double via _f64 when the
 * stretcher was built with precision 64.  Returns 0; -1 invalid resample factor
 * (resampler.rs:11); -2 would block; -3 step outside 1..=W (usize underflow at
 * stretcher.rs:105-106).  `out` must hold ro_stretcher_out_len() values. */
int ro_stretcher_next_window(ro_stretcher *s, float *out, size_t *n_out);
int ro_stretcher_next_window_f64(ro_stretcher *s, double *out, size_t *n_out);
size_t ro_stretcher_out_len(const ro_stretcher *s);
/* derived constants (stretcher.rs:40-59) */
typedef struct {
    float pitch_shifted_factor;
    float corrected_amp_factor;
    size_t window_len;
    size_t half_window_len;
    size_t samples_needed_per_window;
    size_t sample_step_len;
    size_t frames_per_call; /* ceil(S / H) */
    size_t out_len_per_call;
} ro_constants;
void ro_stretcher_constants(const ro_stretcher *s, ro_constants *out);
uint64_t ro_stretcher_frames_done(const ro_stretcher *s);

/* ---- StretcherProcessor driver (stretcher_processor.rs:56-71) + AudioBus::into_audio
 * (audio.rs:152-172): each channel receives its whole input as one chunk and the
 * sender is dropped (main.rs:133-153); windows are produced round-robin until the
 * first channel reports done; per-channel outputs are concatenated.
 * out[c] must hold ro_expected_output_len() floats.  Returns calls made per channel,
 * or a negative ro_stretcher_next_window error. */
long ro_process_all(uint32_t channels, uint32_t sample_rate, const float *const *in,
                    size_t n_in, float factor, float amplitude, int pitch_multiple,
                    size_t window_len, uint64_t seed, int phase_mode, ro_apply_fn plugin,
                    float *const *out, size_t out_cap, size_t *n_out);
/* same, arithmetic in double */
long ro_process_all_f64(uint32_t channels, uint32_t sample_rate, const float *const *in,
                        size_t n_in, float factor, float amplitude, int pitch_multiple,
                        size_t window_len, uint64_t seed, int phase_mode,
                        ro_apply_fn plugin, double *const *out, size_t out_cap,
                        size_t *n_out);
/* closed form of the call/frame/output counts; returns 0 or -1/-3 */
int ro_expected_counts(size_t n_in, float factor, int pitch_multiple, size_t window_len,
                       size_t *calls, size_t *frames, size_t *out_len);

/* ---- fade in / out of live playback (`-x`; mixer.rs:132-190, main.rs:154, player_processor.rs) ----
 * Gain the Mixer applies to output sample `pos` after fade_in_out(Some(fade), Some(fade)) with
 * expected_total_samples = (n_in * factor) as usize: keyframes (0, 0.0), (fade_n, 1.0),
 * (start_n, 1.0), (end_n, 0.0) interpolated with math::sqrt_interp. */
float ro_fade_gain(uint64_t pos, float fade_s, uint32_t sample_rate, uint64_t expected_total_samples);
void ro_fade_keyframes(float fade_s, uint32_t sample_rate, uint64_t expected_total_samples,
                       uint64_t *fade_n, uint64_t *start_n, uint64_t *end_n);

/* time a bounded sample: run `max_calls` calls (0 = all) of one channel; returns
 * calls made.  Used by bench.py's cpu_baseline leg only. */
long ro_bench_channel(const float *in, size_t n_in, uint32_t sample_rate, float factor,
                      float amplitude, int pitch_multiple, size_t window_len,
                      uint64_t seed, uint32_t channel, size_t max_calls, float *sink,
                      size_t *n_out_total);

#ifdef __cplusplus
}
#endif
#endif

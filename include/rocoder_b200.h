/*
 * rocoder_b200.h -- C ABI of the B200-native rocoder stretch engine.
 *
 * Drop-in boundary for ONE path of ajyoon/rocoder: Stretcher::next_window ->
 * ReFFT::resynth -> overlap-add -> integer resample (reference src/stretcher.rs,
 * fft.rs, windows.rs, crossfade.rs, resampler.rs).  Plain pointers and sizes only; no
 * exceptions or aborts cross this boundary; every entry point returns an rcg_status
 * (0 = ok) and rcg_last_error() describes the last failure on the calling thread.
 * An engine is single-threaded like the reference's Stretcher (it is moved into the one
 * DSP thread, stretcher_processor.rs:55-71); distinct engines are independent.
 *
 * The shared library (librocoder_b200.so) is CUDA-only: there is no CPU fallback.
 * Entry points marked [host] are pure host arithmetic and work without a GPU; everything
 * else fails with RCG_ERR_CUDA / RCG_ERR_NO_DEVICE when no sm_100 device is present.
 *
 * Each declaration cites the reference interface it replaces as file:line relative to
 * the reference's src/ directory.  INTEGRATION.md shows the Rust FFI binding.
 */
#ifndef ROCODER_B200_H
#define ROCODER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCG_ABI_VERSION 2
#define RCG_MAX_DEVICES 16

typedef enum rcg_status {
    RCG_OK = 0,
    RCG_ERR_INVALID_ARG = 1,
    RCG_ERR_PITCH = 2,       /* pitch_multiple 0 (assert, stretcher.rs:40), -1 (panic, resampler.rs:11), or outside i8 */
    RCG_ERR_WINDOW = 3,      /* window_len: every EVEN length from 64 to 1048576 is supported (README.md:100 allows any length):
                                powers of two up to 65536 on the fast in-SM kernels, other lengths up to 8192 on the in-SM
                                chirp-z kernels (~10x slower per sample), everything longer on chirp-z convolutions through
                                HBM; odd lengths are not (stretcher.rs:101-103 drifts by one sample per frame for them) */
    RCG_ERR_STEP = 4,        /* sample_step_len outside 1..=W: the reference underflows usize at stretcher.rs:105-106 */
    RCG_ERR_CUDA = 5,
    RCG_ERR_NO_DEVICE = 6,
    RCG_ERR_WOULD_BLOCK = 7, /* next_window needs input that has not been sent and the sender is not closed (recv() at stretcher.rs:125 would block) */
    RCG_ERR_CAPACITY = 8,    /* caller buffer too small */
    RCG_ERR_PLUGIN = 9,
    RCG_ERR_STATE = 10
} rcg_status;

/* flags for rcg_config.flags */
#define RCG_FLAG_RANDOM_SEED 1u   /* ignore `seed`, draw one from the OS like rand::thread_rng (fft.rs:64) */
#define RCG_FLAG_BYPASS_RESYNTH 2u /* test hook: resynth(s)[n] = (s[n]*w[n])*w[n]; checks all indexing bit-exactly */
#define RCG_FLAG_SHARD_HALO 4u     /* multi-device time shards recompute the frame before their first one instead of
                                      receiving the carried half window from the neighbour device: zero communication,
                                      bit-identical output; the check of the exchange path (see n_devices) */
#define RCG_FLAG_SHARD_ALWAYS 8u   /* use every listed device even for jobs too small to gain from it (tests) */
#define RCG_FLAG_FRAME_CLOCK 16u   /* device frequency kernels get a per-frame elapsed_ms: the clock base (wall clock at the
                                      engine's first processing call, or rcg_set_clock_base) plus the frame's position in
                                      output time -- what the reference's per-frame clock (fft.rs:89-92) reads in a run
                                      paced by playback.  Without the flag: one wall-clock value per launch. */

/* Arguments of Stretcher::new (stretcher.rs:30-39) plus the phase seed and the device(s).
 * The Hann window of main.rs:131 is built internally from window_len.
 *
 * Several GPUs behind one engine (n_devices > 1): the reference makes ONE call per file for all channels
 * (main.rs:133-156) and its channels are independent Stretchers (stretcher_processor.rs:55-71), so
 * rcg_process_all partitions the job over the listed devices itself -- by channel when the channel count
 * divides evenly (no communication at all), otherwise by time: contiguous call ranges of every channel, each
 * device reading its own input span, and ONE exchange per boundary: the carried half window (channels x W/2
 * floats, stretcher.rs:101-102) goes from device g to g+1 (cudaMemcpyPeerAsync: NVLink peer-to-peer where the
 * devices are peers) and is added into the first block of g+1 -- there is no phase carry to pass (fft.rs:63-68
 * draws every phase independently).  Output is bit-identical for every device list.  The same ordinal may be
 * listed more than once (shards then share that GPU).  Jobs too small to amortise a second device, and jobs with
 * a host `apply` on top of the kernel stack (it must be called from one thread in the reference's order), run on
 * devices[0] alone.  Streaming calls are served by device (channel mod n_devices); rcg_process_segment /
 * rcg_finish_boundary run on the listed device their buffers live on. */
typedef struct rcg_config {
    uint32_t struct_size;     /* sizeof(rcg_config) */
    uint32_t channels;        /* AudioSpec.channels  (audio.rs:32-37); one Stretcher per channel (main.rs:133-153) */
    uint32_t sample_rate;     /* AudioSpec.sample_rate */
    uint32_t window_len;      /* -w, window.len() */
    float factor;             /* -f */
    float amplitude;          /* -a */
    int32_t pitch_multiple;   /* -p (i8 in the reference) */
    float buffer_dur_s;       /* -b, Duration::as_secs_f32 */
    uint64_t seed;            /* Philox key; (seed, channel, frame, bin) fixes every phase */
    int32_t device;           /* CUDA device ordinal */
    uint32_t flags;
    uint32_t max_lookahead_calls; /* streaming: calls computed per launch when input allows (0 = default 64) */
    uint32_t first_channel;   /* channel index of row 0: Stretcher c of main.rs:133-153 is an engine with channels=1, first_channel=c */
    /* -- ABI 2 (a struct_size that ends before this line is accepted and means one device) -- */
    uint32_t n_devices;       /* 0 or 1: `device` alone; > 1: devices[0 .. n_devices) share every rcg_process_all job */
    int32_t devices[RCG_MAX_DEVICES];
} rcg_config;

/* The constants Stretcher::new derives (stretcher.rs:40-59), f32 arithmetic, truncating casts. */
typedef struct rcg_plan_info {
    float pitch_shifted_factor;       /* stretcher.rs:42-46 */
    float corrected_amp_factor;       /* stretcher.rs:53 */
    uint64_t window_len;
    uint64_t half_window_len;         /* stretcher.rs:54 */
    uint64_t samples_needed_per_window; /* stretcher.rs:47-51 */
    uint64_t sample_step_len;         /* stretcher.rs:55 */
    uint64_t frames_per_call;         /* trips of the while loop at stretcher.rs:91 */
    uint64_t out_len_per_call;        /* resample() output length, resampler.rs:3-35 */
} rcg_plan_info;

/* ---- [host] arithmetic that must be bit-exact with the reference ----------------------- */
/* stretcher.rs:40-59.  Validates pitch, window and step. */
int rcg_plan(const rcg_config *cfg, rcg_plan_info *out);
/* Calls / frames / output samples per channel for n_in input samples: the stop rule of
 * stretcher.rs:123-135 + stretcher_processor.rs:64-68. */
int rcg_expected_counts(const rcg_config *cfg, uint64_t n_in, uint64_t *calls, uint64_t *frames,
                        uint64_t *out_len);
/* windows::hanning, windows.rs:4-9 */
int rcg_hanning(float *out, size_t len);
/* crossfade::hanning_crossfade_compensation, crossfade.rs:4-10 */
int rcg_hanning_crossfade_compensation(float *out, size_t len);
/* Stretcher::channel_bound, stretcher.rs:82-85 */
int rcg_channel_bound(const rcg_config *cfg, uint64_t *bound);
/* phase the engine assigns to `bin` of `frame` of `channel` (replaces fft.rs:64) */
float rcg_phase(uint64_t seed, uint32_t channel, uint64_t frame, uint32_t bin, uint32_t window_len);

const char *rcg_last_error(void);
const char *rcg_status_string(int status);
int rcg_abi_version(void);

/* ---- engine ---------------------------------------------------------------------------- */
typedef struct rcg_engine rcg_engine;

/* Stretcher::new for every channel + ReFFT::new (stretcher.rs:30-76, fft.rs:25-40). */
int rcg_create(const rcg_config *cfg, rcg_engine **out);
void rcg_destroy(rcg_engine *e);
int rcg_get_plan(const rcg_engine *e, rcg_plan_info *out);
/* The `window: Vec<f32>` argument of Stretcher::new (stretcher.rs:36) when it is not the Hann window
 * of main.rs:131 (e.g. windows::rectangular, windows.rs:12-14).  len must equal window_len; call
 * before any processing. */
int rcg_set_window(rcg_engine *e, const float *window, size_t len);

/* -- batched, host buffers: what main.rs:133-160 + AudioBus::into_audio (audio.rs:152-172)
 * do for a whole file.  in[c] holds n_in samples of channel c (the one chunk main.rs:148
 * sends, sender then dropped); out[c] receives the concatenated windows.  Host<->device
 * copies happen inside; pinned buffers (rcg_host_alloc) make them asynchronous.  With n_devices > 1 every
 * listed device takes its share (see rcg_config). */
int rcg_process_all(rcg_engine *e, const float *const *in, uint64_t n_in, float *const *out,
                    uint64_t out_cap, uint64_t *n_out);
/* How the last rcg_process_all was partitioned: devices used, "channel" / "time" / "none" split (static string),
 * and the number of carried-half-window exchanges between devices. */
int rcg_last_partition(const rcg_engine *e, uint32_t *devices_used, const char **split, uint32_t *exchanges);

/* -- batched, device-resident: calls [first_call, first_call + n_calls) of every channel.
 * d_in[c * in_stride + i] is input sample in_origin + i of channel c; samples at or past
 * in_origin + in_valid read as the zero padding of stretcher.rs:130.  d_out[c * out_stride]
 * receives the windows of these calls back to back.
 * How the half window carried into the first call is obtained (stretcher.rs:58-59, 112):
 *   first_call == 0           zeros;
 *   d_tail_in != NULL         [channels][W/2] second half of the preceding frame;
 *   RCG_SEG_DEFER_FIRST       the first half-window block is NOT emitted; its raw first half
 *                             goes to d_head_out [channels][W/2] and rcg_finish_boundary()
 *                             completes it once the neighbour's tail has arrived;
 *   otherwise                 the preceding frame is recomputed (needs in_origin <=
 *                             (first_call*frames_per_call - 1) * sample_step_len).
 * d_tail_out (optional) receives the second half of the last frame.  `stream` is a
 * cudaStream_t: NULL = the engine's own non-blocking stream; pass cudaStreamLegacy
 * ((cudaStream_t)0x1) to name the default stream explicitly.  The call is asynchronous on it
 * (with frequency kernels pushed it synchronises: the spectra visit the host).  The tail / head
 * buffers must be 8-byte aligned; for pitch multiples 1 and 2 so must d_out, with out_stride even.  One launch at a time per
 * engine: the carried-half-window scratch is shared by consecutive launches. */
#define RCG_SEG_DEFER_FIRST 1u
typedef struct rcg_segment {
    const float *d_in;
    int64_t in_stride;
    int64_t in_origin;
    int64_t in_valid;
    float *d_out;
    int64_t out_stride;
    uint64_t first_call;
    uint64_t n_calls;
    const float *d_tail_in;
    float *d_tail_out;
    float *d_head_out;
    uint32_t flags;
    uint32_t reserved;
    void *stream;
} rcg_segment;
int rcg_process_segment(rcg_engine *e, const rcg_segment *seg);
/* Finish the deferred first block of a segment: ((head + tail_prev) * env) * amp, resampled,
 * written where rcg_process_segment left the gap. */
int rcg_finish_boundary(rcg_engine *e, float *d_out, int64_t out_stride, uint64_t first_call,
                        const float *d_head, const float *d_tail_prev, void *stream);
int rcg_synchronize(rcg_engine *e);

/* -- streaming, one window at a time: the Stretcher API itself ----------------------------- */
/* Sender<Vec<f32>>::send on the channel's input (main.rs:148, stretcher.rs:125-128) */
int rcg_send(rcg_engine *e, uint32_t channel, const float *chunk, uint64_t n);
/* dropping the Sender (main.rs:148-152): recv() then fails and the tail is zero padded */
int rcg_close_input(rcg_engine *e, uint32_t channel);
/* Stretcher::next_window, stretcher.rs:87-121 */
int rcg_next_window(rcg_engine *e, uint32_t channel, float *out, uint64_t cap, uint64_t *n_out);
/* Stretcher::is_done, stretcher.rs:78 */
int rcg_is_done(rcg_engine *e, uint32_t channel, int *done);

/* -- frequency-kernel plugin (fft.rs:76-108) -------------------------------------------------
 * C form of the reference's Rust-ABI `apply(elapsed_ms, Vec<(f32,f32)>) -> Vec<(f32,f32)>`
 * (fft.rs:93, README.md:108-113): in/out are n_bins interleaved (re, im) pairs, the full
 * W-bin unnormalised spectrum in natural order.  Return 0; non-zero stands for a panic:
 * the newest kernel is popped and the frame retried with the previous one, else passed
 * through (fft.rs:101-106).  Called once per frame per channel, frames in order.
 * Clock: the reference reads SystemTime before every frame (fft.rs:89-92) and polls for a new library before every
 * frame (fft.rs:78-81).  Here a host `apply` gets the wall clock of its own call; a device `apply` gets ONE
 * elapsed_ms per launch -- all frames of a batch (up to 1 GiB of spectra, or a whole fused launch) are computed at
 * the same moment -- unless RCG_FLAG_FRAME_CLOCK asks for a per-frame clock; new libraries are taken between batches.  A segment that starts mid-stream without
 * d_tail_in re-analyses the frame before it, so `apply` sees that frame once more. */
typedef int (*rcg_apply_fn)(size_t elapsed_ms, const float *in_re_im, size_t n_bins,
                            float *out_re_im);
/* what receiving a Library on kernel_recv does (fft.rs:78-81) */
int rcg_push_frequency_kernel(rcg_engine *e, rcg_apply_fn fn);
/* dlopen(path) and push its `apply` symbol (hotswapper.rs:85-86 + fft.rs:93-94) */
int rcg_load_frequency_kernel(rcg_engine *e, const char *so_path);
/* Device form of the same contract: CUDA source defining
 *     __device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2 *spectrum)
 * which returns the output value of `bin` given the frame's whole W-bin input spectrum.  Compiled
 * with NVRTC for sm_100a (the compile-and-swap role of hotswapper.rs:58-86); the spectra then never
 * leave HBM.  A compile error returns RCG_ERR_PLUGIN and keeps the previous kernel
 * (hotswapper.rs:36-42). */
int rcg_push_device_frequency_kernel(rcg_engine *e, const char *cuda_source);
/* Bin-local device form: CUDA source defining
 *     __device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x)
 * for kernels whose output bin depends on the same input bin only (the README's example, README.md:121-129, every
 * bin times a constant, is one).  NVRTC builds the whole stretch kernel around it for this engine's window size
 * (a few seconds): |apply(X[k])| is evaluated in registers inside the fused spectral step -- only the norm of a
 * kernel's output survives, fft.rs:67 -- so no spectrum ever exists in memory and a frame costs what it costs
 * without a kernel.  Same stack / fall-back rules as the other forms. */
int rcg_push_local_device_frequency_kernel(rcg_engine *e, const char *cuda_source);
int rcg_frequency_kernel_count(const rcg_engine *e, uint32_t *n);
/* RCG_FLAG_FRAME_CLOCK: elapsed_ms of frame 0 (instead of the wall clock at the first processing call): makes a
 * render with a time-dependent device kernel reproducible */
int rcg_set_clock_base(rcg_engine *e, uint64_t base_ms);
/* hotswapper::hotswap(path) (hotswapper.rs:14-33, called from ReFFT::new, fft.rs:31-35): build `path` now on
 * the calling thread, then poll it every poll_ms (0 = the reference's 100 ms, hotswapper.rs:12) on a background
 * thread and build it again whenever it is modified.  Each successful build is queued and becomes the newest
 * kernel at the start of the next batch of frames (fft.rs:78-81 polls without blocking before every frame); a
 * failed build changes nothing (hotswapper.rs:36-42) and its compiler output is kept for rcg_watch_log.  Built by
 * suffix: .cu = device `apply` via NVRTC (`*.local.cu` = the bin-local form, fused into the stretch kernel); .c/.cc/.cpp = host `apply` via `$RCG_CXX|gcc|g++ -O3 -shared -fPIC`
 * (the reference shells out to rustc, hotswapper.rs:60-72); .so = dlopen; .rs = a kernel written for the reference
 * (Rust-ABI `apply`, README.md:108-113) plus an appended C-ABI shim, via `$RCG_RUSTC|rustc --crate-type cdylib` when a Rust
 * compiler is installed.  One watched file per engine, as in
 * the reference.  Must be called before processing starts when the pitch multiple is even. */
int rcg_watch_frequency_kernel(rcg_engine *e, const char *path, uint32_t poll_ms);
/* take queued builds now (also done by every process / next_window call); *swapped = kernels installed */
int rcg_poll_frequency_kernels(rcg_engine *e, uint32_t *swapped);
/* builds that succeeded / failed on the watcher so far, and kernels installed from it */
int rcg_watch_stats(const rcg_engine *e, uint32_t *builds_ok, uint32_t *builds_failed, uint32_t *swaps);
/* compiler output of the last failed build ("" after a success) */
int rcg_watch_log(const rcg_engine *e, char *buf, size_t cap);

/* -- fade in / out (`-x`): what the reference's live Mixer applies (mixer.rs:132-190 with
 * math::sqrt_interp, math.rs:33-40); not applied in file mode by the reference (main.rs:195-204) and
 * therefore not by rcg_process_all either.  Multiplies d_audio[c * stride + k], k < n, by the gain of
 * output position first_pos + k: sqrt(r) over the first fade_s seconds, 1, then sqrt(1 - r) over the
 * last fade_s seconds before expected_total_samples (= (n_in * factor) as usize, main.rs:154), 0 after. */
int rcg_fade_in_out(float *d_audio, int64_t stride, uint32_t channels, uint64_t n, uint64_t first_pos,
                    float fade_s, uint32_t sample_rate, uint64_t expected_total_samples, void *stream);
/* [host] the keyframe positions Mixer::fade_in_out computes (mixer.rs:179-190) */
int rcg_fade_keyframes(float fade_s, uint32_t sample_rate, uint64_t expected_total_samples, uint64_t *fade_n,
                       uint64_t *start_n, uint64_t *end_n);

/* -- helpers ------------------------------------------------------------------------------ */
int rcg_host_alloc(void **ptr, size_t bytes); /* pinned host memory */
int rcg_host_free(void *ptr);
/* [host] pin the calling thread to the CPUs of the NUMA node `device` is attached to (never widening the
 * current affinity mask), so that pinned buffers allocated afterwards sit next to the GPU.  *numa_node = the
 * node bound to, or -1 when the box exposes no NUMA information (then nothing changes).  Optional: call it once
 * per rank before rcg_host_alloc / rcg_create on multi-socket, multi-GPU boxes. */
int rcg_bind_host_thread(int device, int *numa_node);
/* number of CUDA devices this library can run on (sm_100); 0 with RCG_OK when there is none */
int rcg_device_count(int *n);
/* FP32 FMA throughput of the current device, TFLOP/s: the measured compute-roofline denominator */
int rcg_measure_fp32_peak(double *tflops, int reps);
/* kernels launched by this engine so far */
int rcg_launch_count(const rcg_engine *e, uint64_t *n);
/* how the engine's persistent stretch kernel is launched on its (first) device: threads per CTA, frame lanes per CTA,
 * dynamic shared memory per CTA, CTAs resident per SM and the largest grid (SMs x CTAs per SM).  Diagnostic: no
 * reference counterpart; tests pin the occupancy of the Tensor Memory kernels with it. */
int rcg_launch_geometry(const rcg_engine *e, uint32_t *threads, uint32_t *lanes_per_cta, uint64_t *smem_bytes,
                        uint32_t *ctas_per_sm, uint32_t *max_grid);

#ifdef __cplusplus
}
#endif
#endif

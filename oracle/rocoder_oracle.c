/*
 * rocoder_oracle.c -- CPU restatement of ajyoon/rocoder's stretch path (see
 * rocoder_oracle.h).  TEST INFRASTRUCTURE ONLY: parity oracle and timed CPU baseline.
 * Plain C99; build with oracle/Makefile (gcc -O2, no -ffast-math: the reference's f32
 * order of operations is part of the contract).
 */
#define _GNU_SOURCE
#include "rocoder_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* Rust `as usize` on a float: saturating, NaN -> 0 */
static size_t ro_f32_to_usize(float v) {
    if (!(v == v) || v <= 0.0f) return 0;
    if (v >= 18446744073709551615.0f) return (size_t)-1;
    return (size_t)v;
}

/* ---------------------------------------------------------------- small helpers -- */

/* windows.rs:4-9 */
void ro_hanning(float *out, size_t len) {
    const float two_pi = 3.14159274101257324f * 2.0f; /* f32::consts::PI * 2.0 */
    for (size_t i = 0; i < len; i++)
        out[i] = 0.5f - (cosf(((float)i * two_pi) / (float)(len - 1)) * 0.5f);
}

/* windows.rs:12-14 */
void ro_rectangular(float *out, size_t len) {
    for (size_t i = 0; i < len; i++) out[i] = 1.0f;
}

/* windows.rs:19-21 */
void ro_inverse(const float *in, float *out, size_t len) {
    for (size_t i = 0; i < len; i++) out[i] = 1.0f / in[i];
}

/* crossfade.rs:4-10 */
void ro_hanning_crossfade_compensation(float *out, size_t len) {
    const float two_pi = 3.14159274101257324f * 2.0f;
    const float hinv_sqrt2 = (1.0f + sqrtf(sqrtf(0.5f))) * 0.5f;
    for (size_t i = 0; i < len; i++)
        out[i] = 0.5f - ((1.0f - hinv_sqrt2) * cosf(((float)i * two_pi) / (float)(len - 1)));
}

/* math.rs:28-30 */
float ro_lerp(float start, float end, float ratio) { return start + (end - start) * ratio; }

/* math.rs:33-40 */
float ro_sqrt_interp(float start, float end, float ratio) {
    int increasing = start < end;
    float progress = ((ratio * 2.0f) - 1.0f) * (increasing ? 1.0f : -1.0f);
    float factor = sqrtf(0.5f * (1.0f + fmaxf(progress, -1.0f)));
    float abs_interval = fabsf(end - start);
    return (increasing ? start : end) + (abs_interval * factor);
}

size_t ro_resample_len(size_t n, int factor) {
    if (factor == 1) return n;
    if (factor > 1) return (n + (size_t)factor - 1) / (size_t)factor;
    if (factor < -1) return n ? (n - 1) * (size_t)(-factor) : 0;
    return 0;
}

/* ---------------------------------------------------------------- fades ---------- */
/* Mixer::dur_to_sample (mixer.rs:147-149): (dur.as_secs_f32() * sample_rate as f32) as usize */
static uint64_t ro_dur_to_sample(float secs, uint32_t sr) { return ro_f32_to_usize(secs * (float)sr); }

/* Mixer::fade_in_out (mixer.rs:179-190): fade in from 0 s; fade out starts at total_dur - fade where
 * total_dur = Duration::from_secs_f32(expected_total_samples as f32 / sample_rate as f32) */
void ro_fade_keyframes(float fade_s, uint32_t sr, uint64_t expected_total, uint64_t *fade_n, uint64_t *start_n,
                       uint64_t *end_n) {
    float total_s = (float)expected_total / (float)sr;
    /* Durations hold integer nanoseconds; from_secs_f32 rounds to them, subtraction is exact */
    double total_ns = floor((double)total_s * 1e9 + 0.5), fade_ns = floor((double)fade_s * 1e9 + 0.5);
    double start_ns = total_ns - fade_ns;
    if (start_ns < 0) start_ns = 0; /* the reference would panic on Duration underflow */
    *fade_n = ro_dur_to_sample((float)(fade_ns / 1e9), sr);
    *start_n = ro_dur_to_sample((float)(start_ns / 1e9), sr);
    *end_n = ro_dur_to_sample((float)((start_ns + fade_ns) / 1e9), sr);
}

/* Layer::current_amp (mixer.rs:132-145) with keyframes pruned as playback passes them (:107-120) */
float ro_fade_gain(uint64_t pos, float fade_s, uint32_t sr, uint64_t expected_total) {
    uint64_t fade_n, start_n, end_n;
    ro_fade_keyframes(fade_s, sr, expected_total, &fade_n, &start_n, &end_n);
    const uint64_t kp[4] = {0, fade_n, start_n, end_n};
    const float kv[4] = {0.0f, 1.0f, 1.0f, 0.0f};
    /* the active pair is the last one whose `next` position has not been passed */
    int i = 0;
    while (i < 2 && kp[i + 1] < pos) i++;
    if (pos > kp[3]) return kv[3];
    if (kp[i + 1] == kp[i]) return kv[i + 1];
    float progress = (float)(pos - kp[i]) / (float)(kp[i + 1] - kp[i]);
    return ro_sqrt_interp(kv[i], kv[i + 1], progress);
}

/* ---------------------------------------------------------------- phase source --- */

void ro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* rand 0.8.5 UniformFloat<f32>::sample_single(0.0, PI): value1_2 = bits(u >> 9) with
 * exponent 0, value0_1 = value1_2 - 1.0 (exact: k * 2^-23), res = value0_1 * scale + low
 * with scale = PI - 0.0, low = 0.0.  res < PI always holds, so no rejection. */
float ro_phase_from_u32(uint32_t u) {
    const float pi = 3.14159274101257324f; /* fft.rs:13: TWO_PI = f32::consts::PI */
    float value0_1 = (float)(u >> 9) * 1.1920928955078125e-07f; /* 2^-23, exact */
    return value0_1 * pi + 0.0f;
}

/* which Philox block / word serves `bin` */
static void ro_phase_slot(uint32_t bin, uint32_t W, uint32_t *pair, uint32_t *word) {
    uint32_t H = W / 2, Q = H / 2;
    if (bin <= Q)            { *pair = bin;     *word = 0; }
    else if (bin <= H)       { *pair = H - bin; *word = 2; }
    else if (bin < H + Q)    { *pair = bin - H; *word = 3; }
    else                     { *pair = W - bin; *word = 1; }
}

float ro_phase(uint64_t seed, uint32_t channel, uint64_t frame, uint32_t bin, uint32_t W) {
    uint32_t pair, word, out[4];
    ro_phase_slot(bin, W, &pair, &word);
    uint32_t ctr[4] = { pair, (uint32_t)frame, (uint32_t)(frame >> 32), channel };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    ro_philox4x32_10(ctr, key, out);
    return ro_phase_from_u32(out[word]);
}

void ro_phase_fill(uint64_t seed, uint32_t channel, uint64_t frame, uint32_t W, float *phases) {
    uint32_t H = W / 2, Q = H / 2;
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    if (W < 4 || (W & 3)) { /* tiny / odd sizes: per-bin path */
        for (uint32_t b = 0; b < W; b++) phases[b] = ro_phase(seed, channel, frame, b, W);
        return;
    }
    for (uint32_t pair = 0; pair <= Q; pair++) {
        uint32_t ctr[4] = { pair, (uint32_t)frame, (uint32_t)(frame >> 32), channel }, o[4];
        ro_philox4x32_10(ctr, key, o);
        phases[pair] = ro_phase_from_u32(o[0]);
        if (pair == 0) { phases[H] = ro_phase_from_u32(o[2]); continue; }
        phases[W - pair] = ro_phase_from_u32(o[1]);
        if (pair == Q) continue;
        phases[H - pair] = ro_phase_from_u32(o[2]);
        phases[H + pair] = ro_phase_from_u32(o[3]);
    }
}

/* ---------------------------------------------------------------- numeric cores -- */
static float ro_cos_f32(float x) { return cosf(x); }
static float ro_sin_f32(float x) { return sinf(x); }
static float ro_hypot_f32(float a, float b) { return hypotf(a, b); }
static double ro_cos_f64(double x) { return cos(x); }
static double ro_sin_f64(double x) { return sin(x); }
static double ro_hypot_f64(double a, double b) { return hypot(a, b); }

#define REAL float
#define SUF(x) x##_f32
#include "oracle_core.inc"
#undef REAL
#undef SUF
#define REAL double
#define SUF(x) x##_f64
#include "oracle_core.inc"
#undef REAL
#undef SUF

void ro_set_fft_hook_f32(ro_fft_hook_f32 hook) { g_fft_hook_f32 = hook; }

void ro_fft_f32(float *re_im, size_t n, int inverse) {
    fft_plan_f32 *p = fft_plan_new_f32(n);
    fft_exec_f32(p, (cpx_f32 *)re_im, inverse);
    fft_plan_free_f32(p);
}
void ro_fft_f64(double *re_im, size_t n, int inverse) {
    fft_plan_f64 *p = fft_plan_new_f64(n);
    fft_exec_f64(p, (cpx_f64 *)re_im, inverse);
    fft_plan_free_f64(p);
}

int ro_resample(const float *in, size_t n, int factor, float *out, size_t *n_out) {
    return resample_f32(in, n, factor, out, n_out);
}

/* ---------------------------------------------------------------- public Stretcher */
struct ro_stretcher { int precision; stretcher_f32 *s32; stretcher_f64 *s64; };

#define DISPATCH(s, field) ((s)->precision == 64 ? (s)->s64->field : (s)->s32->field)

ro_stretcher *ro_stretcher_new(uint32_t channels, uint32_t sample_rate, float factor,
                               float amplitude, int pitch_multiple, const float *window,
                               size_t window_len, float buffer_dur_s, int precision) {
    ro_stretcher *s = (ro_stretcher *)calloc(1, sizeof(*s));
    s->precision = precision == 64 ? 64 : 32;
    if (s->precision == 64)
        s->s64 = stretcher_new_f64(channels, sample_rate, factor, amplitude, pitch_multiple,
                                   window, window_len, buffer_dur_s);
    else
        s->s32 = stretcher_new_f32(channels, sample_rate, factor, amplitude, pitch_multiple,
                                   window, window_len, buffer_dur_s);
    if (!s->s32 && !s->s64) { free(s); return NULL; }
    return s;
}
void ro_stretcher_free(ro_stretcher *s) {
    if (!s) return;
    stretcher_free_f32(s->s32);
    stretcher_free_f64(s->s64);
    free(s);
}
#define SET(s, stmt32, stmt64) do { if ((s)->precision == 64) { stretcher_f64 *x = (s)->s64; stmt64; } \
                                    else { stretcher_f32 *x = (s)->s32; stmt32; } } while (0)
void ro_stretcher_set_phase_philox(ro_stretcher *s, uint64_t seed, uint32_t channel) {
    SET(s, (x->phase_mode = RO_PHASE_PHILOX, x->seed = seed, x->channel = channel),
           (x->phase_mode = RO_PHASE_PHILOX, x->seed = seed, x->channel = channel));
}
void ro_stretcher_set_phase_table(ro_stretcher *s, const float *phases, size_t n_frames) {
    SET(s, (x->phase_mode = RO_PHASE_TABLE, x->phase_table = phases, x->phase_table_frames = n_frames),
           (x->phase_mode = RO_PHASE_TABLE, x->phase_table = phases, x->phase_table_frames = n_frames));
}
void ro_stretcher_set_phase_bypass(ro_stretcher *s) {
    SET(s, x->phase_mode = RO_PHASE_BYPASS, x->phase_mode = RO_PHASE_BYPASS);
}
void ro_stretcher_set_first_frame(ro_stretcher *s, uint64_t frame) {
    SET(s, (x->frame = frame, x->first_frame = frame), (x->frame = frame, x->first_frame = frame));
}
void ro_stretcher_push_plugin(ro_stretcher *s, ro_apply_fn fn) {
    SET(s, (x->has_kernel_recv = 1, x->n_kernels < 16 ? (void)(x->kernels[x->n_kernels++] = fn) : (void)0),
           (x->has_kernel_recv = 1, x->n_kernels < 16 ? (void)(x->kernels[x->n_kernels++] = fn) : (void)0));
}
void ro_stretcher_send(ro_stretcher *s, const float *chunk, size_t n) {
    if (s->precision == 64) stretcher_send_f64(s->s64, chunk, n); else stretcher_send_f32(s->s32, chunk, n);
}
void ro_stretcher_close(ro_stretcher *s) { SET(s, x->closed = 1, x->closed = 1); }
int ro_stretcher_is_done(const ro_stretcher *s) { return DISPATCH(s, done); }

/* stretcher.rs:82-85 */
size_t ro_stretcher_channel_bound(const ro_stretcher *s) {
    float wl = (float)DISPATCH(s, window_len), sr = (float)DISPATCH(s, sample_rate);
    return ro_f32_to_usize(ceilf((wl / sr) / DISPATCH(s, buffer_dur_s)));
}
int ro_stretcher_ensure_input_samples_available(ro_stretcher *s, size_t n) {
    return s->precision == 64 ? stretcher_ensure_input_f64(s->s64, n) : stretcher_ensure_input_f32(s->s32, n);
}
size_t ro_stretcher_input_buf_len(const ro_stretcher *s) { return DISPATCH(s, input_buf.len); }
void ro_stretcher_input_buf_copy(const ro_stretcher *s, float *out) {
    size_t n = ro_stretcher_input_buf_len(s);
    for (size_t i = 0; i < n; i++)
        out[i] = s->precision == 64 ? (float)s->s64->input_buf.data[s->s64->input_buf.head + i]
                                    : s->s32->input_buf.data[s->s32->input_buf.head + i];
}
size_t ro_stretcher_out_len(const ro_stretcher *s) {
    return ro_resample_len(DISPATCH(s, samples_needed_per_window), DISPATCH(s, pitch_multiple));
}
int ro_stretcher_next_window(ro_stretcher *s, float *out, size_t *n_out) {
    if (s->precision == 64) {
        size_t cap = ro_stretcher_out_len(s), n = 0;
        double *tmp = (double *)malloc(sizeof(double) * (cap + 1));
        int rc = stretcher_next_window_f64(s->s64, tmp, &n);
        for (size_t i = 0; i < n; i++) out[i] = (float)tmp[i];
        free(tmp); *n_out = n; return rc;
    }
    return stretcher_next_window_f32(s->s32, out, n_out);
}
int ro_stretcher_next_window_f64(ro_stretcher *s, double *out, size_t *n_out) {
    if (s->precision == 64) return stretcher_next_window_f64(s->s64, out, n_out);
    size_t cap = ro_stretcher_out_len(s), n = 0;
    float *tmp = (float *)malloc(sizeof(float) * (cap + 1));
    int rc = stretcher_next_window_f32(s->s32, tmp, &n);
    for (size_t i = 0; i < n; i++) out[i] = tmp[i];
    free(tmp); *n_out = n; return rc;
}
void ro_stretcher_constants(const ro_stretcher *s, ro_constants *c) {
    c->pitch_shifted_factor = DISPATCH(s, pitch_shifted_factor);
    c->corrected_amp_factor = DISPATCH(s, corrected_amp_factor);
    c->window_len = DISPATCH(s, window_len);
    c->half_window_len = DISPATCH(s, half_window_len);
    c->samples_needed_per_window = DISPATCH(s, samples_needed_per_window);
    c->sample_step_len = DISPATCH(s, sample_step_len);
    c->frames_per_call = c->half_window_len
        ? (c->samples_needed_per_window + c->half_window_len - 1) / c->half_window_len : 0;
    c->out_len_per_call = ro_stretcher_out_len(s);
}
uint64_t ro_stretcher_frames_done(const ro_stretcher *s) { return DISPATCH(s, frame); }

/* ---------------------------------------------------------------- driver --------- */
int ro_expected_counts(size_t n_in, float factor, int pitch_multiple, size_t W,
                       size_t *calls, size_t *frames, size_t *out_len) {
    float w1 = 1.0f;
    ro_stretcher *s = ro_stretcher_new(1, 44100, factor, 1.0f, pitch_multiple, &w1, 1, 1.0f, 32);
    if (!s || pitch_multiple == -1) { ro_stretcher_free(s); return -1; }
    ro_stretcher_free(s);
    /* constants depend on W only through arithmetic: redo them for the real W */
    int abs_p = pitch_multiple < 0 ? -pitch_multiple : pitch_multiple;
    float psf = pitch_multiple < 0 ? factor / (float)abs_p : factor * (float)abs_p;
    size_t S = pitch_multiple < 0 ? ro_f32_to_usize(ceilf((float)W / (float)abs_p)) : W * (size_t)abs_p;
    size_t H = W / 2, step = ro_f32_to_usize((float)W / (psf * 2.0f));
    if (step < 1 || step > W || H == 0) return -3;
    size_t k = (S + H - 1) / H;
    size_t j1 = n_in < W ? 0 : (n_in - W) / step + 1; /* first zero-padded frame */
    size_t c = j1 / k + 1;
    if (calls) *calls = c;
    if (frames) *frames = c * k;
    if (out_len) *out_len = c * ro_resample_len(S, pitch_multiple);
    return 0;
}

#define PROCESS_ALL_BODY(OUT_T, NEXT)                                                          \
    float *window = (float *)malloc(sizeof(float) * (window_len + 1));                         \
    ro_hanning(window, window_len);                        /* main.rs:131 */                   \
    ro_stretcher **st = (ro_stretcher **)calloc(channels, sizeof(*st));                        \
    long rc = 0, calls = 0;                                                                    \
    for (uint32_t c = 0; c < channels; c++) {              /* main.rs:133-153 */               \
        st[c] = ro_stretcher_new(channels, sample_rate, factor, amplitude, pitch_multiple,     \
                                 window, window_len, 1.0f, PREC);                              \
        if (!st[c]) { rc = -1; goto done; }                                                    \
        if (phase_mode == RO_PHASE_BYPASS) ro_stretcher_set_phase_bypass(st[c]);               \
        else ro_stretcher_set_phase_philox(st[c], seed, c);                                    \
        if (plugin) ro_stretcher_push_plugin(st[c], plugin);                                   \
        ro_stretcher_send(st[c], in[c], n_in);                                                 \
        ro_stretcher_close(st[c]);                                                             \
        n_out[c] = 0;                                                                          \
    }                                                                                          \
    for (;;) {                                             /* stretcher_processor.rs:56-71 */  \
        int stop = 0;                                                                          \
        for (uint32_t c = 0; c < channels; c++) {                                              \
            if (ro_stretcher_is_done(st[c])) { stop = 1; break; }                              \
            size_t n = 0, need = ro_stretcher_out_len(st[c]);                                  \
            if (n_out[c] + need > out_cap) { rc = -4; goto done; }                             \
            int e = NEXT(st[c], out[c] + n_out[c], &n);                                        \
            if (e) { rc = e; goto done; }                                                      \
            n_out[c] += n;                                 /* audio.rs:152-172 concat */       \
            if (c == 0) calls++;                                                               \
        }                                                                                      \
        if (stop) break;                                                                       \
    }                                                                                          \
done:                                                                                          \
    for (uint32_t c = 0; c < channels; c++) ro_stretcher_free(st[c]);                          \
    free(st); free(window);                                                                    \
    return rc ? rc : calls;

long ro_process_all(uint32_t channels, uint32_t sample_rate, const float *const *in, size_t n_in,
                    float factor, float amplitude, int pitch_multiple, size_t window_len,
                    uint64_t seed, int phase_mode, ro_apply_fn plugin, float *const *out,
                    size_t out_cap, size_t *n_out) {
#define PREC 32
    PROCESS_ALL_BODY(float, ro_stretcher_next_window)
#undef PREC
}

long ro_process_all_f64(uint32_t channels, uint32_t sample_rate, const float *const *in, size_t n_in,
                        float factor, float amplitude, int pitch_multiple, size_t window_len,
                        uint64_t seed, int phase_mode, ro_apply_fn plugin, double *const *out,
                        size_t out_cap, size_t *n_out) {
#define PREC 64
    PROCESS_ALL_BODY(double, ro_stretcher_next_window_f64)
#undef PREC
}

long ro_bench_channel(const float *in, size_t n_in, uint32_t sample_rate, float factor,
                      float amplitude, int pitch_multiple, size_t window_len, uint64_t seed,
                      uint32_t channel, size_t max_calls, float *sink, size_t *n_out_total) {
    float *window = (float *)malloc(sizeof(float) * (window_len + 1));
    ro_hanning(window, window_len);
    ro_stretcher *st = ro_stretcher_new(1, sample_rate, factor, amplitude, pitch_multiple, window,
                                        window_len, 1.0f, 32);
    if (!st) { free(window); return -1; }
    ro_stretcher_set_phase_philox(st, seed, channel);
    ro_stretcher_send(st, in, n_in);
    ro_stretcher_close(st);
    long calls = 0; size_t total = 0;
    while (!ro_stretcher_is_done(st) && (max_calls == 0 || (size_t)calls < max_calls)) {
        size_t n = 0;
        int e = ro_stretcher_next_window(st, sink, &n);
        if (e) { calls = e; break; }
        total += n; calls++;
    }
    if (n_out_total) *n_out_total = total;
    ro_stretcher_free(st); free(window);
    return calls;
}

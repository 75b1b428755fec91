// rcg_engine.cu -- host side of librocoder_b200.so: the C ABI declared in
// include/rocoder_b200.h.  Mirrors Stretcher / StretcherProcessor (reference
// stretcher.rs, stretcher_processor.rs) on top of the fused sm_100a kernels.  CUDA only:
// no CPU fallback exists; without a device every engine entry point fails loudly.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rocoder_b200.h"
#include "rcg_hotswap.h"
#include "rcg_kernels.h"
#include "rcg_tables.h"

using namespace rcg;

// ---------------------------------------------------------------------------- errors ------
static thread_local char g_err[512] = "";

static int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver           \
                            ? RCG_ERR_NO_DEVICE : RCG_ERR_CUDA,                                \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* rcg_last_error(void) { return g_err; }
extern "C" int rcg_abi_version(void) { return RCG_ABI_VERSION; }
extern "C" const char* rcg_status_string(int s) {
    switch (s) {
        case RCG_OK: return "ok";
        case RCG_ERR_INVALID_ARG: return "invalid argument";
        case RCG_ERR_PITCH: return "invalid pitch multiple";
        case RCG_ERR_WINDOW: return "unsupported window length";
        case RCG_ERR_STEP: return "sample step outside 1..=window";
        case RCG_ERR_CUDA: return "CUDA error";
        case RCG_ERR_NO_DEVICE: return "no CUDA device";
        case RCG_ERR_WOULD_BLOCK: return "input not available";
        case RCG_ERR_CAPACITY: return "output buffer too small";
        case RCG_ERR_PLUGIN: return "frequency kernel error";
        case RCG_ERR_STATE: return "invalid engine state";
    }
    return "unknown status";
}

// ---------------------------------------------------------------------------- plan --------
// Rust `as usize` on an f32: saturating, NaN -> 0
static uint64_t f32_to_usize(float v) {
    if (!(v == v) || v <= 0.0f) return 0;
    if (v >= 18446744073709551615.0f) return UINT64_MAX;
    return (uint64_t)v;
}

static uint64_t resample_len(uint64_t n, int p) {
    if (p == 1) return n;
    if (p > 1) return (n + (uint64_t)p - 1) / (uint64_t)p;
    return n ? (n - 1) * (uint64_t)(-p) : 0;
}

// longest window: the HBM-resident chirp-z path convolves at length 2^21 >= 2W - 1 (rcg_big.h)
static constexpr uint32_t kMaxWindow = 1u << (kBigMaxLog2M - 1);

// FFT_M of the chirp filter / M in the position order the forward passes of the HBM path leave (rcg_big.h)
static std::vector<cf> host_big_filter(int W, int lm) {
    std::vector<double> re, im;
    host_bluestein_filter_spectrum(W, lm, re, im);
    const size_t M = (size_t)1 << lm;
    std::vector<cf> f(M);
    for (size_t pos = 0; pos < M; pos++) {
        const size_t k = (size_t)big_pos_to_bin(lm, (long long)pos);
        f[pos] = make_cf((float)(re[k] / (double)M), (float)(im[k] / (double)M));
    }
    return f;
}

static int log2_exact(uint32_t w) {
    int l = 0;
    while ((1u << l) < w) l++;
    return (1u << l) == w ? l : -1;
}

extern "C" int rcg_plan(const rcg_config* cfg, rcg_plan_info* out) {
    if (!cfg || !out) return fail(RCG_ERR_INVALID_ARG, "null argument");
    const int p = cfg->pitch_multiple;
    if (p == 0) return fail(RCG_ERR_PITCH, "pitch_multiple must be non-zero (stretcher.rs:40)");
    if (p == -1) return fail(RCG_ERR_PITCH, "invalid resample factor -1 (resampler.rs:11)");
    if (p < -127 || p > 127) return fail(RCG_ERR_PITCH, "pitch_multiple outside the reference's i8 range");
    const uint32_t W = cfg->window_len;
    const int l2 = log2_exact(W);
    if (l2 < 0) {
        // not a power of two (README.md:100 allows any length): the chirp-z paths -- in one SM up to 8192
        // (rcg_frame_bluestein.h), through HBM above that (rcg_big.h)
        if (W & 1u)
            return fail(RCG_ERR_WINDOW, "window_len %u is odd: the reference appends W - W/2 samples per frame while "
                                        "advancing by W/2 (stretcher.rs:101-103); not supported", W);
        if (W < 64 || W > kMaxWindow)
            return fail(RCG_ERR_WINDOW, "window_len %u: even window lengths from 64 to %u are supported", W, kMaxWindow);
    } else if (l2 < 6 || W > kMaxWindow) {
        return fail(RCG_ERR_WINDOW, "window_len %u: need an even length in [64, %u]", W, kMaxWindow);
    }
    if (cfg->channels == 0) return fail(RCG_ERR_INVALID_ARG, "channels must be >= 1");
    const float abs_p = (float)(p < 0 ? -p : p);
    // stretcher.rs:42-46
    const float psf = p < 0 ? cfg->factor / abs_p : cfg->factor * abs_p;
    // stretcher.rs:47-51
    const uint64_t S = p < 0 ? f32_to_usize(ceilf((float)W / abs_p)) : (uint64_t)W * (uint64_t)(p < 0 ? -p : p);
    const uint64_t H = W / 2;
    // stretcher.rs:55
    const uint64_t step = f32_to_usize((float)W / (psf * 2.0f));
    if (step < 1 || step > W)
        return fail(RCG_ERR_STEP, "sample_step_len %llu outside 1..=%u (factor %g, pitch %d)",
                    (unsigned long long)step, W, (double)cfg->factor, p);
    out->pitch_shifted_factor = psf;
    out->corrected_amp_factor = fmaxf(4.0f, psf / 4.0f) * cfg->amplitude;  // stretcher.rs:53
    out->window_len = W;
    out->half_window_len = H;
    out->samples_needed_per_window = S;
    out->sample_step_len = step;
    out->frames_per_call = (S + H - 1) / H;
    out->out_len_per_call = resample_len(S, p);
    return RCG_OK;
}

static void counts_for(const rcg_plan_info& pl, uint64_t n_in, uint64_t* calls, uint64_t* frames, uint64_t* out_len) {
    // first frame j with j*step + W > n_in sets `done` (stretcher.rs:124-131); the call that
    // contains it is still completed and emitted; the processor stops before the next one.
    const uint64_t j1 = n_in < pl.window_len ? 0 : (n_in - pl.window_len) / pl.sample_step_len + 1;
    const uint64_t c = j1 / pl.frames_per_call + 1;
    if (calls) *calls = c;
    if (frames) *frames = c * pl.frames_per_call;
    if (out_len) *out_len = c * pl.out_len_per_call;
}

extern "C" int rcg_expected_counts(const rcg_config* cfg, uint64_t n_in, uint64_t* calls, uint64_t* frames,
                                   uint64_t* out_len) {
    rcg_plan_info pl;
    int rc = rcg_plan(cfg, &pl);
    if (rc) return rc;
    counts_for(pl, n_in, calls, frames, out_len);
    return RCG_OK;
}

extern "C" int rcg_hanning(float* out, size_t len) {
    if (!out) return fail(RCG_ERR_INVALID_ARG, "null argument");
    host_hanning(out, len);
    return RCG_OK;
}
extern "C" int rcg_hanning_crossfade_compensation(float* out, size_t len) {
    if (!out) return fail(RCG_ERR_INVALID_ARG, "null argument");
    host_crossfade_compensation(out, len);
    return RCG_OK;
}
extern "C" int rcg_channel_bound(const rcg_config* cfg, uint64_t* bound) {
    if (!cfg || !bound) return fail(RCG_ERR_INVALID_ARG, "null argument");
    // stretcher.rs:82-85
    *bound = f32_to_usize(ceilf(((float)cfg->window_len / (float)cfg->sample_rate) / cfg->buffer_dur_s));
    return RCG_OK;
}
extern "C" float rcg_phase(uint64_t seed, uint32_t channel, uint64_t frame, uint32_t bin, uint32_t W) {
    const uint32_t H = W / 2, Q = H / 2;
    uint32_t pair, word;
    if (bin <= Q) { pair = bin; word = 0; }
    else if (bin <= H) { pair = H - bin; word = 2; }
    else if (bin < H + Q) { pair = bin - H; word = 3; }
    else { pair = W - bin; word = 1; }
    uint32_t u[4];
    philox4x32_10(pair, (uint32_t)frame, (uint32_t)(frame >> 32), channel, make_philox_keys(seed), u);
    return phase_from_u32(u[word]);
}

// ---------------------------------------------------------------------------- engine ------
struct ChannelState {
    std::vector<float> input;      // samples [input_base, input_base + input.size())
    uint64_t input_base = 0;
    uint64_t received = 0;         // total samples sent
    bool closed = false;
    bool done = false;
    uint64_t next_call = 0;        // next call to compute
    std::deque<std::pair<std::vector<float>, bool>> ready;  // (window, used zero padding)
    float* d_tail[2] = {nullptr, nullptr};
    int tail_cur = 0;
};

// one entry of the frequency-kernel stack (fft.rs:22 `kernels: Vec<Library>`): a host `apply`
// or a device `apply` compiled with NVRTC
struct FreqKernel {
    rcg_apply_fn host = nullptr;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t dev = nullptr;     // gather form: runs on the spectra in HBM between analysis and synthesis
    cudaKernel_t fused = nullptr;   // bin-local form: the whole stretch kernel with the user's apply compiled in
    cudaLibrary_t twin_lib = nullptr;   // bin-local form again as a plain per-bin kernel over spectra in HBM (`dev` points
                                        // into it): takes over mid-batch when the kernel above it on the stack fails
};

struct rcg_engine {
    rcg_config cfg{};
    rcg_plan_info plan{};
    int log2w = 0;
    int device = 0;
    int sm_count = 0;
    LaunchGeometry geo{};
    int max_grid = 0;
    PhiloxKeys keys{};
    uint64_t seed = 0;
    float *d_window = nullptr, *d_env = nullptr, *d_scratch = nullptr, *d_window_even = nullptr, *d_env_even = nullptr;
    float* d_lerp_ratio = nullptr;
    uint2* d_pair_tab = nullptr;
    uint4* d_quad_tab = nullptr;
    cf *d_tw = nullptr, *d_spec_tw = nullptr, *d_spec_tw_hr = nullptr;
    // pass twiddles (and, W = 16384, the spectral step's table entries) as Tensor Memory rows (rcg_common.h): the image
    // of the full-rate path and, where it differs (spectral_in_tmem), the one of the half-rate path
    float *d_tw_image = nullptr, *d_tw_image_hr = nullptr;
    bool bluestein = false;   // window_len is not a power of two: chirp-z path (rcg_frame_bluestein.h)
    int bs_lm = 0;            // log2 of its convolution length
    BluesteinTables bt{};
    cf *d_bs_chirp = nullptr, *d_bs_filter = nullptr;
    bool big = false;         // window too long for an SM (pair): chirp-z convolutions through HBM (rcg_big.h)
    BigTables bgt{};
    cf* d_big_work = nullptr;
    float *d_big_ybuf = nullptr, *d_big_carry = nullptr;
    bool pair = false;        // W = 65536: a frame across a 2-CTA cluster
    PairTables pt{};
    cf *d_pair_cross = nullptr, *d_pair_even = nullptr, *d_pair_odd = nullptr;
    cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;
    uint64_t launches = 0;
    // process_all buffers (grow only)
    float *d_in = nullptr, *d_out = nullptr, *d_tail_pp[2] = {nullptr, nullptr};
    size_t d_in_cap = 0, d_out_cap = 0;
    std::vector<cudaEvent_t> events;
    // streaming
    std::vector<ChannelState> ch;
    float *d_sin = nullptr, *d_sout = nullptr, *h_sout = nullptr;
    size_t d_sin_cap = 0, d_sout_cap = 0, h_sout_cap = 0;
    // plugin path buffers: spectra of one batch of frames, device + pinned host
    cf *d_spec = nullptr, *h_spec = nullptr, *d_spec2 = nullptr;
    size_t spec_cap = 0, h_spec_cap = 0;  // in cf
    float* d_plug_tail[2] = {nullptr, nullptr};
    std::vector<float> plug_tmp;
    // plugins (fft.rs:22: kernels: Vec<Library>)
    std::vector<FreqKernel> kernels;
    // a kernel that needs the spectra in memory (host or gather form), or a watched file, was installed at some
    // point (even if all have since failed): the engine then stays on the full-rate path.  Fused bin-local
    // kernels do not set it -- they run inside the stretch kernel, half-rate synthesis included.
    bool had_kernels = false;
    std::vector<void*> libs;
    // hotswapper.rs:14-33: watched source file, builds queued for the next batch
    std::unique_ptr<rcg::KernelWatcher> watcher;
    uint32_t swaps = 0;
    // ---- several devices behind one engine (rcg_config.n_devices > 1) --------------------------------------
    // The object the caller holds only routes: `shards` are complete single-device engines, one per listed device
    // (all channels each, so any shard can take any channel range).  Exchange buffers of a shard that runs a
    // time range: raw first half of its first block, the neighbour's carried half window, and the event that says
    // its own last carried half window is final.
    std::vector<rcg_engine*> shards;
    float *d_head = nullptr, *d_prev_tail = nullptr;
    cudaEvent_t tail_ready = nullptr;
    const float* final_tail = nullptr;      // where the last launch of the current job left the carried half window
    // RCG_FLAG_FRAME_CLOCK: elapsed_ms of frame 0 (0 = not fixed yet: the wall clock at the first processing call)
    uint64_t clock_base = 0;
    bool clock_base_set = false;
    // how the last rcg_process_all was partitioned
    uint32_t part_devices = 1, part_exchanges = 0;
    const char* part_split = "none";
};
static int drain_watcher(rcg_engine* e, uint32_t* swapped);
static uint64_t total_launches(const rcg_engine* e);
static int install_everywhere(rcg_engine* e, const std::vector<char>& cubin, const std::vector<char>* twin);

static int pick_run_len(uint64_t frames_per_channel, uint32_t channels, int slots) {
    // minimise waves * (frames per item + 1 halo frame)
    if (frames_per_channel == 0) return 1;
    uint64_t best_cost = UINT64_MAX;
    int best = 1;
    const int rmax = (int)std::min<uint64_t>(frames_per_channel, 512);
    for (int r = 1; r <= rmax; r++) {
        uint64_t items = (uint64_t)channels * ((frames_per_channel + r - 1) / r);
        uint64_t waves = (items + slots - 1) / slots;
        uint64_t cost = waves * (uint64_t)(r + 1);
        if (cost < best_cost) { best_cost = cost; best = r; }
    }
    return best;
}

// rcg_config as ABI 1 laid it out (everything before n_devices): still accepted, means one device
static constexpr size_t kConfigSizeV1 = offsetof(rcg_config, n_devices);

static int create_group(const rcg_config& cfg, const rcg_plan_info& pl, rcg_engine** out);

extern "C" int rcg_create(const rcg_config* cfg_in, rcg_engine** out) {
    if (!cfg_in || !out) return fail(RCG_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (cfg_in->struct_size != sizeof(rcg_config) && cfg_in->struct_size != kConfigSizeV1)
        return fail(RCG_ERR_INVALID_ARG, "rcg_config.struct_size %u: expected %zu (or %zu, ABI 1)", cfg_in->struct_size,
                    sizeof(rcg_config), kConfigSizeV1);
    rcg_config cfg_full;
    memset(&cfg_full, 0, sizeof(cfg_full));
    memcpy(&cfg_full, cfg_in, cfg_in->struct_size);
    cfg_full.struct_size = sizeof(rcg_config);
    const rcg_config* cfg = &cfg_full;
    if (cfg->n_devices > RCG_MAX_DEVICES) return fail(RCG_ERR_INVALID_ARG, "n_devices %u > %d", cfg->n_devices, RCG_MAX_DEVICES);
    rcg_plan_info pl;
    int rc = rcg_plan(cfg, &pl);
    if (rc) return rc;
    if (cfg->n_devices > 1) return create_group(cfg_full, pl, out);
    if (cfg->n_devices == 1) cfg_full.device = cfg_full.devices[0];
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail(RCG_ERR_NO_DEVICE, "no CUDA device");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(RCG_ERR_INVALID_ARG, "device %d of %d", cfg->device, ndev);
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(RCG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", cfg->device,
                    prop.major, prop.minor);
    rcg_engine* e = new rcg_engine();
    e->cfg = *cfg;
    e->plan = pl;
    e->log2w = log2_exact(cfg->window_len);
    e->device = cfg->device;
    e->sm_count = prop.multiProcessorCount;
    e->seed = cfg->seed;
    if (cfg->flags & RCG_FLAG_RANDOM_SEED) {
        std::random_device rd;
        e->seed = ((uint64_t)rd() << 32) ^ rd();
    }
    e->keys = make_philox_keys(e->seed);
    auto bail = [&](int status) { rcg_destroy(e); return status; };
#define TRY_E(expr)                                                                                  \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return bail(fail(RCG_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)));        \
    } while (0)
    e->pair = (e->log2w == 16);
    e->big = (e->log2w < 0 && cfg->window_len > 8192) || e->log2w > 16;
    e->bluestein = (e->log2w < 0) || e->big;
    if (e->big) {
        e->bs_lm = bluestein_log2m((int)cfg->window_len);
        e->geo = LaunchGeometry{256, 1, 0, 1};
        e->max_grid = e->sm_count;
    } else if (e->bluestein) {
        e->bs_lm = bluestein_log2m((int)cfg->window_len);
        TRY_E(bluestein_geometry(e->bs_lm, (int)cfg->window_len, &e->geo));
        if (e->geo.max_ctas_per_sm < 1) return bail(fail(RCG_ERR_CUDA, "chirp-z kernel does not fit on an SM"));
        e->max_grid = e->sm_count * e->geo.max_ctas_per_sm;
    } else if (e->pair) {
        TRY_E(pair_geometry(&e->geo));
        e->max_grid = e->sm_count / 2;   // clusters
    } else {
        TRY_E(stretch_geometry(e->log2w, &e->geo));
        if (e->geo.max_ctas_per_sm < 1) return bail(fail(RCG_ERR_CUDA, "stretch kernel does not fit on an SM"));
        e->max_grid = e->sm_count * e->geo.max_ctas_per_sm;
    }
    const size_t W = cfg->window_len, H = W / 2;
    std::vector<float> window(W), env(H);
    host_hanning(window.data(), W);           // main.rs:131
    host_crossfade_compensation(env.data(), H);  // stretcher.rs:56
    // the in-CTA transform (the HBM path has none: its tables of this kind are dummies)
    const int plan_log2h = e->big ? 6 : e->bluestein ? e->bs_lm : e->pair ? kPairLog2HL : e->log2w - 1;
    std::vector<cf> tw = host_pass_twiddles(plan_log2h), stw = host_spectral_twiddles(plan_log2h);
    std::vector<cf> stw_hr = host_spectral_twiddles_hr(plan_log2h);
    if (e->pair) {
        std::vector<cf> cross = host_pair_cross_twiddles(e->log2w - 1);
        std::vector<cf> se = host_pair_spectral_twiddles(e->log2w - 1, 0), so = host_pair_spectral_twiddles(e->log2w - 1, 1);
        TRY_E(cudaMalloc(&e->d_pair_cross, cross.size() * sizeof(cf)));
        TRY_E(cudaMalloc(&e->d_pair_even, se.size() * sizeof(cf)));
        TRY_E(cudaMalloc(&e->d_pair_odd, so.size() * sizeof(cf)));
        TRY_E(cudaMemcpy(e->d_pair_cross, cross.data(), cross.size() * sizeof(cf), cudaMemcpyHostToDevice));
        TRY_E(cudaMemcpy(e->d_pair_even, se.data(), se.size() * sizeof(cf), cudaMemcpyHostToDevice));
        TRY_E(cudaMemcpy(e->d_pair_odd, so.data(), so.size() * sizeof(cf), cudaMemcpyHostToDevice));
    }
    if (e->big) {
        std::vector<cf> chirp = host_bluestein_chirp((int)W), filt = host_big_filter((int)W, e->bs_lm);
        TRY_E(cudaMalloc(&e->d_bs_chirp, chirp.size() * sizeof(cf)));
        TRY_E(cudaMalloc(&e->d_bs_filter, filt.size() * sizeof(cf)));
        TRY_E(cudaMemcpy(e->d_bs_chirp, chirp.data(), chirp.size() * sizeof(cf), cudaMemcpyHostToDevice));
        TRY_E(cudaMemcpy(e->d_bs_filter, filt.data(), filt.size() * sizeof(cf), cudaMemcpyHostToDevice));
        // frames per batch: 256 MB of convolution workspace
        const int batch = (int)std::max<long long>(1, (1LL << 25) >> e->bs_lm);
        TRY_E(cudaMalloc(&e->d_big_work, ((size_t)batch << e->bs_lm) * sizeof(cf)));
        TRY_E(cudaMalloc(&e->d_big_ybuf, (size_t)batch * W * sizeof(float)));
        TRY_E(cudaMalloc(&e->d_big_carry, (size_t)cfg->channels * H * sizeof(float)));
        e->bgt = BigTables{};
        e->bgt.chirp = e->d_bs_chirp; e->bgt.filter = e->d_bs_filter; e->bgt.work = e->d_big_work;
        e->bgt.ybuf = e->d_big_ybuf; e->bgt.carry = e->d_big_carry;
        e->bgt.w = (int)W; e->bgt.lm = e->bs_lm; e->bgt.batch = batch;
        e->bgt.n_top = big_top_plan(e->bs_lm, e->bgt.top_rlog2);
    } else if (e->bluestein) {
        std::vector<cf> chirp = host_bluestein_chirp((int)W), filt = host_bluestein_filter((int)W, e->bs_lm);
        TRY_E(cudaMalloc(&e->d_bs_chirp, chirp.size() * sizeof(cf)));
        TRY_E(cudaMalloc(&e->d_bs_filter, filt.size() * sizeof(cf)));
        TRY_E(cudaMemcpy(e->d_bs_chirp, chirp.data(), chirp.size() * sizeof(cf), cudaMemcpyHostToDevice));
        TRY_E(cudaMemcpy(e->d_bs_filter, filt.data(), filt.size() * sizeof(cf), cudaMemcpyHostToDevice));
        e->bt = BluesteinTables{e->d_bs_chirp, e->d_bs_filter, (int)W};
    }
    std::vector<float> we = host_decimate2(window), ee = host_decimate2(env);
    {   // frame-invariant slot-index tables of the spectral steps
        std::vector<uint2> ptab = host_pair_table(plan_log2h);
        std::vector<uint4> qtab = host_quad_table(plan_log2h);
        TRY_E(cudaMalloc(&e->d_pair_tab, ptab.size() * sizeof(uint2)));
        TRY_E(cudaMalloc(&e->d_quad_tab, qtab.size() * sizeof(uint4)));
        TRY_E(cudaMemcpy(e->d_pair_tab, ptab.data(), ptab.size() * sizeof(uint2), cudaMemcpyHostToDevice));
        TRY_E(cudaMemcpy(e->d_quad_tab, qtab.data(), qtab.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    }
    {   // j as f32 / factor as f32, resampler.rs:29
        const int q = cfg->pitch_multiple < 0 ? -cfg->pitch_multiple : 1;
        std::vector<float> ratio(128, 0.0f);
        for (int j = 0; j < q && j < 128; j++) ratio[j] = (float)j / (float)q;
        TRY_E(cudaMalloc(&e->d_lerp_ratio, ratio.size() * sizeof(float)));
        TRY_E(cudaMemcpy(e->d_lerp_ratio, ratio.data(), ratio.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (!e->bluestein && tw_in_tmem(plan_log2h)) {
        std::vector<float> img = host_tmem_image(plan_log2h, false, window.data(), env.data());
        TRY_E(cudaMalloc(&e->d_tw_image, img.size() * sizeof(float)));
        TRY_E(cudaMemcpy(e->d_tw_image, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
        if (spectral_in_tmem(plan_log2h)) {
            img = host_tmem_image(plan_log2h, true);
            TRY_E(cudaMalloc(&e->d_tw_image_hr, img.size() * sizeof(float)));
            TRY_E(cudaMemcpy(e->d_tw_image_hr, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    TRY_E(cudaMalloc(&e->d_window_even, we.size() * sizeof(float)));
    TRY_E(cudaMalloc(&e->d_env_even, ee.size() * sizeof(float)));
    TRY_E(cudaMalloc(&e->d_spec_tw_hr, stw_hr.size() * sizeof(cf)));
    TRY_E(cudaMemcpy(e->d_window_even, we.data(), we.size() * sizeof(float), cudaMemcpyHostToDevice));
    TRY_E(cudaMemcpy(e->d_env_even, ee.data(), ee.size() * sizeof(float), cudaMemcpyHostToDevice));
    TRY_E(cudaMemcpy(e->d_spec_tw_hr, stw_hr.data(), stw_hr.size() * sizeof(cf), cudaMemcpyHostToDevice));
    TRY_E(cudaMalloc(&e->d_window, W * sizeof(float)));
    TRY_E(cudaMalloc(&e->d_env, H * sizeof(float)));
    TRY_E(cudaMalloc(&e->d_tw, std::max<size_t>(tw.size(), 1) * sizeof(cf)));
    TRY_E(cudaMalloc(&e->d_spec_tw, stw.size() * sizeof(cf)));
    const size_t scratch_floats = e->pair ? (size_t)std::max<int>(e->max_grid, (int)cfg->channels) * 3 * H
                                          : (size_t)e->max_grid * e->geo.frames_per_cta * H;
    TRY_E(cudaMalloc(&e->d_scratch, scratch_floats * sizeof(float)));
    e->pt = PairTables{e->d_pair_cross, e->d_pair_even, e->d_pair_odd, e->d_scratch};
    TRY_E(cudaMemcpy(e->d_window, window.data(), W * sizeof(float), cudaMemcpyHostToDevice));
    TRY_E(cudaMemcpy(e->d_env, env.data(), H * sizeof(float), cudaMemcpyHostToDevice));
    TRY_E(cudaMemcpy(e->d_tw, tw.data(), tw.size() * sizeof(cf), cudaMemcpyHostToDevice));
    TRY_E(cudaMemcpy(e->d_spec_tw, stw.data(), stw.size() * sizeof(cf), cudaMemcpyHostToDevice));
    TRY_E(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    TRY_E(cudaStreamCreateWithFlags(&e->h2d, cudaStreamNonBlocking));
    TRY_E(cudaStreamCreateWithFlags(&e->d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) TRY_E(cudaMalloc(&e->d_tail_pp[i], (size_t)cfg->channels * H * sizeof(float)));
    TRY_E(cudaMalloc(&e->d_head, (size_t)cfg->channels * H * sizeof(float)));
    TRY_E(cudaMalloc(&e->d_prev_tail, (size_t)cfg->channels * H * sizeof(float)));
    TRY_E(cudaEventCreateWithFlags(&e->tail_ready, cudaEventDisableTiming));
    e->ch.resize(cfg->channels);
    for (auto& c : e->ch)
        for (int i = 0; i < 2; i++) {
            TRY_E(cudaMalloc(&c.d_tail[i], H * sizeof(float)));
            TRY_E(cudaMemset(c.d_tail[i], 0, H * sizeof(float)));
        }
#undef TRY_E
    *out = e;
    return RCG_OK;
}

// n_devices > 1: one complete engine per listed device; the returned object routes
static int create_group(const rcg_config& cfg, const rcg_plan_info& pl, rcg_engine** out) {
    rcg_engine* g = new rcg_engine();
    g->cfg = cfg;
    g->plan = pl;
    g->log2w = log2_exact(cfg.window_len);
    g->device = cfg.devices[0];
    g->seed = cfg.seed;
    if (cfg.flags & RCG_FLAG_RANDOM_SEED) {   // one seed for every shard, or the shards would disagree
        std::random_device rd;
        g->seed = ((uint64_t)rd() << 32) ^ rd();
    }
    for (uint32_t i = 0; i < cfg.n_devices; i++) {
        rcg_config c = cfg;
        c.n_devices = 0;
        c.device = cfg.devices[i];
        c.seed = g->seed;
        c.flags &= ~(uint32_t)RCG_FLAG_RANDOM_SEED;
        rcg_engine* sh = nullptr;
        const int rc = rcg_create(&c, &sh);
        if (rc) { rcg_destroy(g); return rc; }
        g->shards.push_back(sh);
    }
    // neighbours exchange the carried half window: direct peer access where the hardware offers it (NVLink on an
    // HGX board); cudaMemcpyPeerAsync stages through the host otherwise
    for (size_t i = 0; i + 1 < g->shards.size(); i++) {
        const int a = g->shards[i]->device, b = g->shards[i + 1]->device;
        if (a == b) continue;
        int ok = 0;
        if (cudaDeviceCanAccessPeer(&ok, b, a) == cudaSuccess && ok) {
            cudaSetDevice(b);
            if (cudaDeviceEnablePeerAccess(a, 0) != cudaSuccess) (void)cudaGetLastError();   // already enabled is fine
        }
        (void)cudaGetLastError();
    }
    *out = g;
    return RCG_OK;
}

extern "C" void rcg_destroy(rcg_engine* e) {
    if (!e) return;
    e->watcher.reset();   // joins the polling thread
    if (!e->shards.empty()) {
        for (rcg_engine* sh : e->shards) rcg_destroy(sh);
        for (void* l : e->libs) dlclose(l);
        delete e;
        return;
    }
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->d2h) cudaStreamSynchronize(e->d2h);
    if (e->h2d) cudaStreamSynchronize(e->h2d);
    cudaFree(e->d_window); cudaFree(e->d_env); cudaFree(e->d_tw); cudaFree(e->d_spec_tw); cudaFree(e->d_scratch);
    cudaFree(e->d_window_even); cudaFree(e->d_env_even); cudaFree(e->d_spec_tw_hr); cudaFree(e->d_lerp_ratio);
    cudaFree(e->d_pair_tab); cudaFree(e->d_quad_tab); cudaFree(e->d_tw_image); cudaFree(e->d_tw_image_hr);
    cudaFree(e->d_pair_cross); cudaFree(e->d_pair_even); cudaFree(e->d_pair_odd);
    cudaFree(e->d_bs_chirp); cudaFree(e->d_bs_filter);
    cudaFree(e->d_big_work); cudaFree(e->d_big_ybuf); cudaFree(e->d_big_carry);
    cudaFree(e->d_in); cudaFree(e->d_out); cudaFree(e->d_tail_pp[0]); cudaFree(e->d_tail_pp[1]);
    cudaFree(e->d_sin); cudaFree(e->d_sout); cudaFree(e->d_spec); cudaFree(e->d_spec2);
    for (auto& k : e->kernels) { if (k.lib) cudaLibraryUnload(k.lib); if (k.twin_lib) cudaLibraryUnload(k.twin_lib); }
    cudaFree(e->d_head); cudaFree(e->d_prev_tail);
    if (e->tail_ready) cudaEventDestroy(e->tail_ready);
    cudaFree(e->d_plug_tail[0]); cudaFree(e->d_plug_tail[1]);
    if (e->h_spec) cudaFreeHost(e->h_spec);
    if (e->h_sout) cudaFreeHost(e->h_sout);
    for (auto& c : e->ch) { cudaFree(c.d_tail[0]); cudaFree(c.d_tail[1]); }
    for (auto ev : e->events) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->h2d) cudaStreamDestroy(e->h2d);
    if (e->d2h) cudaStreamDestroy(e->d2h);
    for (void* l : e->libs) dlclose(l);
    delete e;
}

// Stretcher::new takes `window: Vec<f32>` (stretcher.rs:36); main.rs:131 always passes
// windows::hanning(len), which rcg_create installs.  This replaces it (e.g. windows::rectangular).
extern "C" int rcg_set_window(rcg_engine* e, const float* window, size_t len) {
    if (!e || !window) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (len != e->cfg.window_len) return fail(RCG_ERR_INVALID_ARG, "window has %zu samples, engine window_len is %u", len, e->cfg.window_len);
    if (!e->shards.empty()) {
        for (rcg_engine* sh : e->shards)
            if (int rc = rcg_set_window(sh, window, len)) return rc;
        return RCG_OK;
    }
    if (total_launches(e) > 0) return fail(RCG_ERR_STATE, "the window must be set before processing starts");
    CUDA_TRY(cudaSetDevice(e->device));
    std::vector<float> w(window, window + len), we = host_decimate2(w);
    CUDA_TRY(cudaMemcpy(e->d_window, w.data(), len * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(e->d_window_even, we.data(), we.size() * sizeof(float), cudaMemcpyHostToDevice));
    const int plan_log2h = e->log2w - 1;
    if (e->d_tw_image && !e->pair && !e->bluestein && !e->big && window_in_tmem(plan_log2h)) {
        // the full-rate launches read the window from the Tensor Memory image (rcg_common.h: window_in_tmem)
        std::vector<float> env(len / 2);
        host_crossfade_compensation(env.data(), env.size());
        const std::vector<float> img = host_tmem_image(plan_log2h, false, w.data(), env.data());
        CUDA_TRY(cudaMemcpy(e->d_tw_image, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return RCG_OK;
}

static uint64_t total_launches(const rcg_engine* e) {
    uint64_t n = e->launches;
    for (const rcg_engine* sh : e->shards) n += sh->launches;
    return n;
}

// the shard of a multi-device engine whose device owns the device pointer `p` (the engine itself when it has one device)
static rcg_engine* shard_for_pointer(rcg_engine* e, const void* p) {
    if (e->shards.empty()) return e;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
        (void)cudaGetLastError();
        return nullptr;
    }
    for (rcg_engine* sh : e->shards)
        if (sh->device == at.device) return sh;
    return nullptr;
}

extern "C" int rcg_get_plan(const rcg_engine* e, rcg_plan_info* out) {
    if (!e || !out) return fail(RCG_ERR_INVALID_ARG, "null argument");
    *out = e->plan;
    return RCG_OK;
}
extern "C" int rcg_launch_count(const rcg_engine* e, uint64_t* n) {
    if (!e || !n) return fail(RCG_ERR_INVALID_ARG, "null argument");
    *n = total_launches(e);
    return RCG_OK;
}
extern "C" int rcg_launch_geometry(const rcg_engine* e, uint32_t* threads, uint32_t* lanes_per_cta, uint64_t* smem_bytes,
                                   uint32_t* ctas_per_sm, uint32_t* max_grid) {
    if (!e) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (!e->shards.empty()) e = e->shards[0];
    if (threads) *threads = (uint32_t)e->geo.threads;
    if (lanes_per_cta) *lanes_per_cta = (uint32_t)e->geo.frames_per_cta;
    if (smem_bytes) *smem_bytes = (uint64_t)e->geo.smem_bytes;
    if (ctas_per_sm) *ctas_per_sm = (uint32_t)e->geo.max_ctas_per_sm;
    if (max_grid) *max_grid = (uint32_t)e->max_grid;
    return RCG_OK;
}
extern "C" int rcg_synchronize(rcg_engine* e) {
    if (!e) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (!e->shards.empty()) {
        for (rcg_engine* sh : e->shards)
            if (int rc = rcg_synchronize(sh)) return rc;
        return RCG_OK;
    }
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->h2d));
    CUDA_TRY(cudaStreamSynchronize(e->d2h));
    return RCG_OK;
}

// Even pitch multiples keep only even output samples (resampler.rs:15-18): the synthesis side
// then runs at half size.  Not once a frequency kernel that works on spectra in memory was loaded (that path
// keeps the full layout), so the carried half windows of fused and plugin launches are never mixed within one
// engine state.
static bool engine_half_rate(const rcg_engine* e) {
    return use_half_rate(e->cfg.pitch_multiple, 0u) && !e->had_kernels && !e->pair && !e->bluestein;
}

static void fill_params(const rcg_engine* e, StretchParams& p) {
    memset(&p, 0, sizeof(p));
    p.window = e->d_window; p.env = e->d_env; p.tw = e->d_tw; p.tw0 = e->d_tw; p.spec_tw = e->d_spec_tw;
    p.tw_image = e->d_tw_image;
    p.spec_tw_hr = e->d_spec_tw_hr; p.window_even = e->d_window_even; p.env_even = e->d_env_even;
    p.lerp_ratio = e->d_lerp_ratio;
    p.pair_tab = e->d_pair_tab; p.quad_tab = e->d_quad_tab;
    p.tail_scratch = e->d_scratch;
    p.channels = (int)e->cfg.channels;
    p.step = (int)e->plan.sample_step_len;
    p.pitch = e->cfg.pitch_multiple;
    p.samples_needed = (int)e->plan.samples_needed_per_window;
    p.amp = e->plan.corrected_amp_factor;
    p.inv_w = 1.0f / (float)e->cfg.window_len;
    p.flags = (e->cfg.flags & RCG_FLAG_BYPASS_RESYNTH) ? kFlagBypass : 0u;
    p.keys = e->keys;
}

static unsigned long long wall_clock_ms() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::milliseconds>(
               std::chrono::system_clock::now().time_since_epoch()).count();   // fft.rs:89-92
}
// the clock a device frequency kernel sees: {base, num, den} with elapsed_ms(frame) = base + frame * num / den.
// Default: the wall clock of this launch for every frame.  RCG_FLAG_FRAME_CLOCK: base = the engine's clock base,
// advancing by one frame's share of output time, 1000 * out_len_per_call / (frames_per_call * sample_rate) ms.
struct KernelClock { unsigned long long base, num, den; };
static KernelClock kernel_clock(rcg_engine* e) {
    if (!(e->cfg.flags & RCG_FLAG_FRAME_CLOCK)) return KernelClock{wall_clock_ms(), 0ull, 1ull};
    if (!e->clock_base_set) { e->clock_base = wall_clock_ms(); e->clock_base_set = true; }
    return KernelClock{e->clock_base, 1000ull * e->plan.out_len_per_call,
                       std::max<unsigned long long>(1ull, e->plan.frames_per_call * (unsigned long long)e->cfg.sample_rate)};
}

// launch calls [first_call, first_call + n_calls) for `channels` rows
struct SpectrumRef { cf* spectrum; long long frame0, frames; };

// Caller errors in a segment description, checked once before any kernel is tried: a bad argument must never be
// mistaken for a failing frequency kernel (which is dropped for good, fft.rs:101-106).
static int validate_segment(const rcg_engine* e, const rcg_segment& s, bool spectra_given) {
    auto misaligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7u) != 0; };
    // float2 stores: the carried / head buffers always; the output only where it is written
    // directly from the overlap-add phase (pitch 1, and pitch 2 on the half-rate path)
    const bool direct_out = !e->bluestein && (e->cfg.pitch_multiple == 1 || (e->cfg.pitch_multiple == 2 && engine_half_rate(e)));
    if (misaligned(s.d_tail_in) || misaligned(s.d_tail_out) || misaligned(s.d_head_out))
        return fail(RCG_ERR_INVALID_ARG, "tail / head buffers must be 8-byte aligned");
    if (direct_out && (misaligned(s.d_out) || (s.out_stride & 1)))
        return fail(RCG_ERR_INVALID_ARG, "d_out must be 8-byte aligned and out_stride even for pitch multiples 1 and 2");
    if ((s.flags & RCG_SEG_DEFER_FIRST) && !s.d_head_out) return fail(RCG_ERR_INVALID_ARG, "RCG_SEG_DEFER_FIRST needs d_head_out");
    const long long frame_begin = (long long)(s.first_call * e->plan.frames_per_call);
    const bool needs_halo = frame_begin > 0 && !s.d_tail_in && !(s.flags & RCG_SEG_DEFER_FIRST);
    if (needs_halo && !spectra_given && s.in_origin > (frame_begin - 1) * (long long)e->plan.sample_step_len)
        return fail(RCG_ERR_INVALID_ARG, "segment starts at frame %lld but its input starts past the halo frame", frame_begin);
    return RCG_OK;
}

static int launch_segment(rcg_engine* e, const rcg_segment& s, int channels, cudaStream_t stream,
                          int channel_offset = 0, const SpectrumRef* spec = nullptr, cudaKernel_t fused = nullptr) {
    if (s.n_calls == 0) return RCG_OK;
    if (int rc = validate_segment(e, s, spec != nullptr)) return rc;
    StretchParams p;
    fill_params(e, p);
    if (spec) {
        p.flags |= kFlagSpectrumIn;
        p.spectrum = spec->spectrum; p.spec_frame0 = spec->frame0; p.spec_frames = spec->frames;
    }
    if (engine_half_rate(e)) {
        p.flags |= kFlagHalfRate;
        if (e->d_tw_image_hr) p.tw_image = e->d_tw_image_hr;
    }
    const uint64_t k = e->plan.frames_per_call;
    p.in = s.d_in; p.in_stride = s.in_stride; p.in_origin = s.in_origin; p.in_valid = s.in_valid;
    p.out = s.d_out; p.out_stride = s.out_stride;
    p.out_origin = (long long)(s.first_call * e->plan.out_len_per_call);
    p.frame_begin = (long long)(s.first_call * k);
    p.frame_end = (long long)((s.first_call + s.n_calls) * k);
    p.channels = channels;
    p.channel_offset = (int)e->cfg.first_channel + channel_offset;
    p.tail_in = s.d_tail_in; p.tail_out = s.d_tail_out; p.head_out = s.d_head_out;
    if (s.flags & RCG_SEG_DEFER_FIRST) p.flags |= kFlagDeferFirst;
    const uint64_t nframes = (uint64_t)(p.frame_end - p.frame_begin);
    const int max_grid = e->max_grid;
    const int slots = max_grid * e->geo.frames_per_cta;
    p.run_len = pick_run_len(nframes, (uint32_t)channels, slots);
    p.n_runs = (int)((nframes + p.run_len - 1) / p.run_len);
    const long long items = (long long)p.n_runs * channels;
    // one item per CTA first (every SM busy), frame lanes double up only once the CTAs run out; W = 65536
    // counts clusters (frames_per_cta = 1)
    const int grid = (int)std::min<long long>(items, max_grid);
    cudaError_t ce;
    if (fused) {   // the stretch kernel NVRTC built around a bin-local frequency kernel (same geometry as the fused one)
        const KernelClock kc = kernel_clock(e);
        p.elapsed_ms = kc.base; p.clock_num = kc.num; p.clock_den = kc.den;
        int n_items = (int)items;
        void* args[] = {(void*)&p, (void*)&n_items};
        void* pair_args[] = {(void*)&p, (void*)&e->pt, (void*)&n_items};   // W = 65536: `grid` counts 2-CTA clusters
        ce = items <= 0 ? cudaSuccess
             : e->pair  ? cudaLaunchKernel((const void*)fused, dim3(2 * grid), dim3(e->geo.threads), pair_args, e->geo.smem_bytes, stream)
                        : cudaLaunchKernel((const void*)fused, dim3(grid), dim3(e->geo.threads), args, e->geo.smem_bytes, stream);
    } else {
        ce = e->big         ? launch_big(p, e->bgt, stream)
             : e->bluestein ? launch_bluestein(e->bs_lm, p, e->bt, grid, stream)
             : e->pair    ? launch_stretch_pair(p, e->pt, grid, stream)
                          : launch_stretch(e->log2w, p, grid, stream);
    }
    if (ce != cudaSuccess) return fail(RCG_ERR_CUDA, "stretch kernel launch failed: %s", cudaGetErrorString(ce));
    e->launches++;
    return RCG_OK;
}


// drop the newest frequency kernel for good ("kernel panicked, retrying with last or noop.", fft.rs:101-106)
static void pop_kernel(rcg_engine* e) {
    FreqKernel& k = e->kernels.back();
    if (k.lib) cudaLibraryUnload(k.lib);
    if (k.twin_lib) cudaLibraryUnload(k.twin_lib);
    e->kernels.pop_back();
}

// A kernel that failed was dropped for good by the shard that met it; every shard of a multi-device engine carries
// the same stack, pushed in the same order, so after a job all of them drop down to the shortest one left.
static void equalise_kernel_stacks(rcg_engine* g) {
    size_t depth = SIZE_MAX;
    for (rcg_engine* sh : g->shards) depth = std::min(depth, sh->kernels.size());
    for (rcg_engine* sh : g->shards) {
        if (sh->kernels.size() > depth) cudaSetDevice(sh->device);
        while (sh->kernels.size() > depth) pop_kernel(sh);
    }
}

// ---- frequency-kernel path (fft.rs:76-108) ---------------------------------------------------
// Host `apply` on one frame's spectrum, in place; mirrors the pop-and-retry of fft.rs:101-106: a failing kernel is
// dropped for good and the previous one is tried.  Returns false when the frame is still unprocessed because the
// kernel now on top is a device kernel (the caller continues on the device from this frame on); true otherwise
// (applied, or no kernel left: the spectrum passes through unchanged).
static bool apply_host_kernels(rcg_engine* e, float* spectrum, size_t n_bins) {
    while (!e->kernels.empty()) {
        if (!e->kernels.back().host) return false;
        const size_t ms = (size_t)std::chrono::duration_cast<std::chrono::milliseconds>(
                              std::chrono::system_clock::now().time_since_epoch()).count();  // fft.rs:89-92
        e->plug_tmp.resize(2 * n_bins);
        if (e->kernels.back().host(ms, spectrum, n_bins, e->plug_tmp.data()) == 0) {
            memcpy(spectrum, e->plug_tmp.data(), 2 * n_bins * sizeof(float));
            return true;
        }
        pop_kernel(e);  // "kernel panicked, retrying with last or noop."
    }
    return true;
}

// calls [first_call, first_call + n_calls) with frequency kernels that need the spectra in memory: analysis kernel
// -> `apply` per frame (host kernels: over PCIe, in the reference's call order; device kernels: in HBM) -> synthesis
// kernel.  The kernel on top of the stack can change in the middle of a batch (a host kernel fails and a device
// kernel lies beneath it, or the other way round): the frames not yet processed then continue with the new top,
// exactly where the reference's retry would (fft.rs:101-106).
static int run_segment_plugins(rcg_engine* e, const rcg_segment& s, int channels, cudaStream_t stream,
                               int channel_offset) {
    const uint64_t k = e->plan.frames_per_call, W = e->plan.window_len, H = W / 2;
    // spectra per batch: 256 MiB when they visit the host, 1 GiB when a device kernel keeps them in HBM
    const bool dev_top = !e->kernels.empty() && e->kernels.back().dev;
    const uint64_t budget_cf = (uint64_t)(dev_top ? 1024 : 256) << 17;
    uint64_t fpb = std::max<uint64_t>(k + 1, budget_cf / ((uint64_t)channels * W));  // frames/channel/batch
    uint64_t cpb = std::max<uint64_t>(1, (fpb - 1) / k);
    fpb = cpb * k + 1;
    const size_t need = (size_t)channels * fpb * W;
    if (e->spec_cap < need) {
        cudaFree(e->d_spec); e->d_spec = nullptr;
        e->spec_cap = 0;
        cudaFree(e->d_spec2); e->d_spec2 = nullptr;
        CUDA_TRY(cudaMalloc(&e->d_spec, need * sizeof(cf)));
        CUDA_TRY(cudaMalloc(&e->d_spec2, need * sizeof(cf)));
        e->spec_cap = need;
    }
    for (int i = 0; i < 2; i++)
        if (!e->d_plug_tail[i]) CUDA_TRY(cudaMalloc(&e->d_plug_tail[i], (size_t)e->cfg.channels * H * sizeof(float)));
    const uint64_t end = s.first_call + s.n_calls;
    int pp = 0;
    for (uint64_t c0 = s.first_call; c0 < end; c0 += cpb) {
        const uint64_t nb = std::min(cpb, end - c0);
        const bool first = (c0 == s.first_call), last = (c0 + nb == end);
        long long f0 = (long long)(c0 * k), f1 = (long long)((c0 + nb) * k);
        // the first batch recomputes the halo frame unless the carried half window is provided
        const bool halo = first && f0 > 0 && !s.d_tail_in && !(s.flags & RCG_SEG_DEFER_FIRST);
        const long long a0 = halo ? f0 - 1 : f0;
        if (halo && s.in_origin > a0 * (long long)e->plan.sample_step_len)
            return fail(RCG_ERR_INVALID_ARG, "segment input starts past its halo frame");
        StretchParams p;
        fill_params(e, p);
        p.in = s.d_in; p.in_stride = s.in_stride; p.in_origin = s.in_origin; p.in_valid = s.in_valid;
        p.channels = channels;
        p.channel_offset = (int)e->cfg.first_channel + channel_offset;
        p.frame_begin = a0; p.frame_end = f1;
        p.spectrum = e->d_spec; p.spec_frame0 = a0; p.spec_frames = f1 - a0;
        const int slots = e->max_grid * e->geo.frames_per_cta;
        const uint64_t nframes = (uint64_t)(f1 - a0);
        p.run_len = (int)std::max<uint64_t>(1, (nframes * channels + slots - 1) / slots);
        p.n_runs = (int)((nframes + p.run_len - 1) / p.run_len);
        const long long items = (long long)p.n_runs * channels;
        const int grid = (int)std::min<long long>(items, e->max_grid);
        cudaError_t ce = e->pair ? launch_analysis_pair(p, e->pt, grid, stream) : launch_analysis(e->log2w, p, grid, stream);
        if (ce != cudaSuccess) return fail(RCG_ERR_CUDA, "analysis kernel launch failed: %s", cudaGetErrorString(ce));
        e->launches++;
        const size_t bytes = (size_t)channels * nframes * W * sizeof(cf);

        // Frames of the batch in the reference's order (stretcher_processor.rs:62-70: call by call, channel by
        // channel, frame by frame; a recomputed halo frame first).  next[c] = first frame index (relative to a0) of
        // channel c that no kernel has processed yet; the spectra live in d_spec, and in h_spec while `on_host`.
        std::vector<uint64_t> next((size_t)channels, 0);
        bool on_host = false;
        cf* synth_src = e->d_spec;
        while (!e->kernels.empty()) {
            bool all_done = true;
            for (int c = 0; c < channels; c++) all_done = all_done && next[c] >= nframes;
            if (all_done) break;
            if (e->kernels.back().host) {
                if (!on_host) {   // pinned staging only when spectra really visit the host
                    if (e->h_spec_cap < need) {
                        if (e->h_spec) cudaFreeHost(e->h_spec);
                        e->h_spec = nullptr; e->h_spec_cap = 0;
                        CUDA_TRY(cudaMallocHost(&e->h_spec, need * sizeof(cf)));
                        e->h_spec_cap = need;
                    }
                    CUDA_TRY(cudaMemcpyAsync(e->h_spec, e->d_spec, bytes, cudaMemcpyDeviceToHost, stream));
                    CUDA_TRY(cudaStreamSynchronize(stream));
                    on_host = true;
                }
                // walk the reference order, skipping what is already done; stop when the top stops being a host kernel
                bool handed_over = false;
                const uint64_t lead = halo ? 1 : 0;   // frame 0 of the batch is the halo frame
                for (uint64_t call = 0; call <= nb && !handed_over; call++) {   // call 0 = the halo frames, if any
                    const uint64_t fb = call == 0 ? 0 : lead + (call - 1) * k, fe = call == 0 ? lead : lead + call * k;
                    for (int c = 0; c < channels && !handed_over; c++)
                        for (uint64_t fi = std::max(fb, next[c]); fi < fe; fi++) {
                            if (!apply_host_kernels(e, reinterpret_cast<float*>(e->h_spec + ((size_t)c * nframes + fi) * W), W)) {
                                handed_over = true;
                                break;
                            }
                            next[c] = fi + 1;
                        }
                }
                if (!handed_over)
                    for (int c = 0; c < channels; c++) next[c] = nframes;
                continue;
            }
            // device `apply` (the gather form, or the per-bin twin of a fused kernel): the spectra stay in HBM
            if (on_host) {
                CUDA_TRY(cudaMemcpyAsync(e->d_spec, e->h_spec, bytes, cudaMemcpyHostToDevice, stream));
                on_host = false;
            }
            const unsigned n_bins = (unsigned)W;
            const KernelClock kc = kernel_clock(e);
            const unsigned long long ms = kc.base, ck_num = kc.num, ck_den = kc.den, fpc = nframes;
            bool failed = false;
            bool untouched = true;
            for (int c = 0; c < channels; c++) untouched = untouched && next[c] == 0;
            if (untouched) {   // the usual case: one launch over the whole batch, synthesis reads its output buffer
                const unsigned long long n_fr = (unsigned long long)channels * nframes;
                const cf* in_p = e->d_spec;
                cf* out_p = e->d_spec2;
                const unsigned long long fr0 = (unsigned long long)a0;   // frame of spectrum 0 of every channel
                void* args[] = {(void*)&in_p, (void*)&out_p, (void*)&n_bins, (void*)&n_fr, (void*)&ms, (void*)&fr0, (void*)&fpc,
                                (void*)&ck_num, (void*)&ck_den};
                const unsigned long long total = n_fr * n_bins;
                const unsigned blocks = (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)e->sm_count * 16);
                if (cudaLaunchKernel((const void*)e->kernels.back().dev, dim3(blocks), dim3(256), args, 0, stream) == cudaSuccess) {
                    e->launches++;
                    synth_src = e->d_spec2;
                    break;
                }
                failed = true;
            }
            for (int c = 0; c < channels && !failed; c++) {
                if (next[c] >= nframes) continue;
                const unsigned long long n_fr = nframes - next[c];
                const cf* in_p = e->d_spec + ((size_t)c * nframes + next[c]) * W;
                cf* out_p = e->d_spec2 + ((size_t)c * nframes + next[c]) * W;
                const unsigned long long fr0 = (unsigned long long)a0 + next[c];   // one channel's remaining frames
                void* args[] = {(void*)&in_p, (void*)&out_p, (void*)&n_bins, (void*)&n_fr, (void*)&ms, (void*)&fr0, (void*)&n_fr,
                                (void*)&ck_num, (void*)&ck_den};
                const unsigned long long total = n_fr * n_bins;
                const unsigned blocks = (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)e->sm_count * 16);
                if (cudaLaunchKernel((const void*)e->kernels.back().dev, dim3(blocks), dim3(256), args, 0, stream) != cudaSuccess) {
                    failed = true;
                    break;
                }
                e->launches++;
                // (a gather kernel reads other bins of its frame: it must not write in place)
                CUDA_TRY(cudaMemcpyAsync(e->d_spec + ((size_t)c * nframes + next[c]) * W, out_p, (size_t)total * sizeof(cf),
                                         cudaMemcpyDeviceToDevice, stream));
                next[c] = nframes;
            }
            if (failed) {
                (void)cudaGetLastError();
                pop_kernel(e);   // unusable kernel: fall back to the previous one (fft.rs:101-106)
            }
        }
        if (on_host) CUDA_TRY(cudaMemcpyAsync(e->d_spec, e->h_spec, bytes, cudaMemcpyHostToDevice, stream));
        rcg_segment b = s;
        b.first_call = c0; b.n_calls = nb;
        b.d_out = s.d_out + (c0 - s.first_call) * e->plan.out_len_per_call;
        b.d_tail_in = first ? s.d_tail_in : e->d_plug_tail[pp];
        b.d_tail_out = (last && s.d_tail_out) ? s.d_tail_out : e->d_plug_tail[pp ^ 1];
        b.d_head_out = first ? s.d_head_out : nullptr;
        b.flags = first ? s.flags : 0u;
        pp ^= 1;
        SpectrumRef ref{synth_src, a0, f1 - a0};
        int rc = launch_segment(e, b, channels, stream, channel_offset, &ref);
        if (rc) return rc;
        // (the next batch reuses the spectrum buffers in stream order: no host synchronisation needed)
    }
    return RCG_OK;
}

static int run_segment(rcg_engine* e, const rcg_segment& s, int channels, cudaStream_t stream, int channel_offset) {
    if (e->watcher)   // fft.rs:78-81: take what the hot-swapper has built, without blocking
        if (int rc = drain_watcher(e, nullptr)) return rc;
    if (s.n_calls == 0) return RCG_OK;
    if (int rc = validate_segment(e, s, false)) return rc;
    while (!e->kernels.empty() && !(e->cfg.flags & RCG_FLAG_BYPASS_RESYNTH)) {
        if (!e->kernels.back().fused) return run_segment_plugins(e, s, channels, stream, channel_offset);
        // newest kernel is bin-local: one fused launch, no spectra in memory
        const int rc = launch_segment(e, s, channels, stream, channel_offset, nullptr, e->kernels.back().fused);
        if (rc == RCG_OK) return RCG_OK;
        if (rc != RCG_ERR_CUDA) return rc;   // not the kernel's fault: keep it
        (void)cudaGetLastError();
        pop_kernel(e);   // unusable kernel: fall back to the previous one (fft.rs:101-106)
    }
    return launch_segment(e, s, channels, stream, channel_offset);
}

extern "C" int rcg_process_segment(rcg_engine* e, const rcg_segment* seg) {
    if (!e || !seg || !seg->d_in || !seg->d_out) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (!e->shards.empty()) {
        if (e->watcher)
            if (int rc = drain_watcher(e, nullptr)) return rc;
        rcg_engine* sh = shard_for_pointer(e, seg->d_out);
        if (!sh) return fail(RCG_ERR_INVALID_ARG, "d_out is not memory of a device this engine lists");
        const int rc = rcg_process_segment(sh, seg);
        equalise_kernel_stacks(e);
        return rc;
    }
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = seg->stream ? (cudaStream_t)seg->stream : e->stream;
    return run_segment(e, *seg, (int)e->cfg.channels, st, 0);
}

extern "C" int rcg_finish_boundary(rcg_engine* e, float* d_out, int64_t out_stride, uint64_t first_call,
                                   const float* d_head, const float* d_tail_prev, void* stream) {
    if (!e || !d_out || !d_head || !d_tail_prev) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (!e->shards.empty()) {
        rcg_engine* sh = shard_for_pointer(e, d_out);
        if (!sh) return fail(RCG_ERR_INVALID_ARG, "d_out is not memory of a device this engine lists");
        return rcg_finish_boundary(sh, d_out, out_stride, first_call, d_head, d_tail_prev, stream);
    }
    CUDA_TRY(cudaSetDevice(e->device));
    StretchParams p;
    fill_params(e, p);
    p.out = d_out; p.out_stride = out_stride;
    p.out_origin = (long long)(first_call * e->plan.out_len_per_call);
    p.frame_begin = (long long)(first_call * e->plan.frames_per_call);
    if (engine_half_rate(e)) p.flags |= kFlagHalfRate;
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    cudaError_t ce = e->big         ? launch_boundary_big(p, e->bgt, d_head, d_tail_prev, st)
                     : e->bluestein ? launch_boundary_bluestein(e->bs_lm, p, e->bt, d_head, d_tail_prev, st)
                     : e->pair      ? launch_boundary_pair(p, e->pt, d_head, d_tail_prev, st)
                                    : launch_boundary(e->log2w, p, d_head, d_tail_prev, st);
    if (ce != cudaSuccess) return fail(RCG_ERR_CUDA, "boundary kernel launch failed: %s", cudaGetErrorString(ce));
    e->launches++;
    return RCG_OK;
}

static cudaEvent_t get_event(rcg_engine* e, size_t i) {
    while (e->events.size() <= i) {
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        e->events.push_back(ev);
    }
    return e->events[i];
}

// -- batched, host buffers -----------------------------------------------------------------------
// What one device does for its share of a job: channels [ch0, ch0 + nch) x calls [call0, call1).
// Time shards (call0 > 0) get the half window carried into their first block from the neighbour device
// (`prev`), or recompute the frame before it (RCG_FLAG_SHARD_HALO).
struct TailSignal {   // "the event after my last launch has been recorded" (or: I failed), shard g -> shard g + 1
    std::mutex m;
    std::condition_variable cv;
    int state = 0;    // 0 pending, 1 recorded, -1 failed
    void set(int v) { { std::lock_guard<std::mutex> l(m); state = v; } cv.notify_all(); }
    int wait() { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return state != 0; }); return state; }
};
struct RangeJob {
    uint32_t ch0 = 0, nch = 0;
    uint64_t call0 = 0, call1 = 0;
    rcg_engine* prev = nullptr;       // exchange: shard that owns the calls before call0
    TailSignal* prev_signal = nullptr;
    TailSignal* my_signal = nullptr;  // set when a later shard waits for this one's carried half window
};

static int process_range_impl(rcg_engine* e, const float* const* in, uint64_t n_in, float* const* out, const RangeJob& job) {
    CUDA_TRY(cudaSetDevice(e->device));
    const uint32_t C = job.nch;
    const uint64_t k = e->plan.frames_per_call, step = e->plan.sample_step_len, W = e->plan.window_len;
    const uint64_t olen = e->plan.out_len_per_call, H = W / 2;
    const uint64_t calls = job.call1 - job.call0;
    if (calls == 0 || C == 0) return RCG_OK;
    const bool exchange = job.call0 > 0 && job.prev != nullptr;
    const bool halo = job.call0 > 0 && !exchange;
    // device input = samples [s_lo, s_hi) of every channel of the range
    // (rounded down to a multiple of 4 samples: frames keep the 8-byte alignment they have in the whole input, and
    // with it the branch-free float2 load path)
    const uint64_t s_lo = std::min<uint64_t>(n_in, (job.call0 * k - (halo ? 1 : 0)) * step) & ~(uint64_t)3;
    const uint64_t s_hi = std::max(s_lo, std::min<uint64_t>(n_in, (job.call1 * k - 1) * step + W));
    const uint64_t n_span = s_hi - s_lo, out_len = calls * olen;
    const size_t in_stride = ((size_t)n_span + 3) & ~(size_t)3;
    const size_t out_stride = ((size_t)out_len + 3) & ~(size_t)3;
    if (e->d_in_cap < in_stride * C) {
        cudaFree(e->d_in); e->d_in = nullptr; e->d_in_cap = 0;
        CUDA_TRY(cudaMalloc(&e->d_in, std::max<size_t>(in_stride * C, 4) * sizeof(float)));
        e->d_in_cap = in_stride * C;
    }
    if (e->d_out_cap < out_stride * C) {
        cudaFree(e->d_out); e->d_out = nullptr; e->d_out_cap = 0;
        CUDA_TRY(cudaMalloc(&e->d_out, std::max<size_t>(out_stride * C, 4) * sizeof(float)));
        e->d_out_cap = out_stride * C;
    }
    // chunks: enough frames per chunk to keep every work lane busy for >= 16 frames,
    // at most 16 chunks, so that device->host copies of finished chunks overlap compute
    const uint64_t slots = (uint64_t)e->max_grid * e->geo.frames_per_cta;
    uint64_t min_calls = (slots * 16 + k * C - 1) / (k * C);
    uint64_t chunk_calls = std::max<uint64_t>(min_calls, (calls + 15) / 16);
    chunk_calls = std::max<uint64_t>(1, std::min(chunk_calls, calls));
    uint64_t uploaded = 0;  // samples of the span already on the device
    size_t ev_i = 0;
    int pp = 0;
    for (uint64_t c0 = job.call0; c0 < job.call1; c0 += chunk_calls) {
        const uint64_t nc = std::min(chunk_calls, job.call1 - c0);
        const bool first = (c0 == job.call0), last = (c0 + nc == job.call1);
        // input this chunk reads: up to the end of its last frame
        const uint64_t need = std::min<uint64_t>(n_in, ((c0 + nc) * k - 1) * step + W);
        const uint64_t need_rel = need > s_lo ? need - s_lo : 0;
        if (need_rel > uploaded) {
            for (uint32_t c = 0; c < C; c++)
                CUDA_TRY(cudaMemcpyAsync(e->d_in + c * in_stride + uploaded, in[job.ch0 + c] + s_lo + uploaded,
                                         (need_rel - uploaded) * sizeof(float), cudaMemcpyHostToDevice, e->h2d));
            uploaded = need_rel;
            cudaEvent_t up = get_event(e, ev_i++);
            CUDA_TRY(cudaEventRecord(up, e->h2d));
            CUDA_TRY(cudaStreamWaitEvent(e->stream, up, 0));
        }
        rcg_segment s{};
        s.d_in = e->d_in; s.in_stride = (int64_t)in_stride; s.in_origin = (int64_t)s_lo; s.in_valid = (int64_t)n_span;
        s.d_out = e->d_out + (c0 - job.call0) * olen; s.out_stride = (int64_t)out_stride;
        s.first_call = c0; s.n_calls = nc;
        s.d_tail_in = first ? nullptr : e->d_tail_pp[pp];
        s.d_tail_out = e->d_tail_pp[pp ^ 1];
        if (first && exchange) { s.flags |= RCG_SEG_DEFER_FIRST; s.d_head_out = e->d_head; }
        pp ^= 1;
        int rc = run_segment(e, s, (int)C, e->stream, (int)job.ch0);
        if (rc) return rc;
        if (last) {
            e->final_tail = s.d_tail_out;
            if (job.my_signal) {
                CUDA_TRY(cudaEventRecord(e->tail_ready, e->stream));
                job.my_signal->set(1);
            }
        }
        cudaEvent_t done = get_event(e, ev_i++);
        CUDA_TRY(cudaEventRecord(done, e->stream));
        CUDA_TRY(cudaStreamWaitEvent(e->d2h, done, 0));
        // the deferred first block leaves later, once the neighbour's half window has been added
        const uint64_t skip = (first && exchange) ? olen : 0;
        const uint64_t o0 = (c0 - job.call0) * olen + skip, on = nc * olen - skip;
        if (on)
            for (uint32_t c = 0; c < C; c++)
                CUDA_TRY(cudaMemcpyAsync(out[job.ch0 + c] + job.call0 * olen + o0, e->d_out + c * out_stride + o0,
                                         on * sizeof(float), cudaMemcpyDeviceToHost, e->d2h));
    }
    if (exchange) {
        // one exchange per boundary: the neighbour's last carried half window (C x W/2 floats), peer to peer
        if (job.prev_signal->wait() < 0) return fail(RCG_ERR_STATE, "the device before this time shard failed");
        CUDA_TRY(cudaStreamWaitEvent(e->stream, job.prev->tail_ready, 0));
        CUDA_TRY(cudaMemcpyPeerAsync(e->d_prev_tail, e->device, job.prev->final_tail, job.prev->device,
                                     (size_t)C * H * sizeof(float), e->stream));
        int rc = rcg_finish_boundary(e, e->d_out, (int64_t)out_stride, job.call0, e->d_head, e->d_prev_tail, e->stream);
        if (rc) return rc;
        cudaEvent_t done = get_event(e, ev_i++);
        CUDA_TRY(cudaEventRecord(done, e->stream));
        CUDA_TRY(cudaStreamWaitEvent(e->d2h, done, 0));
        for (uint32_t c = 0; c < C; c++)
            CUDA_TRY(cudaMemcpyAsync(out[job.ch0 + c] + job.call0 * olen, e->d_out + c * out_stride, olen * sizeof(float),
                                     cudaMemcpyDeviceToHost, e->d2h));
    }
    CUDA_TRY(cudaStreamSynchronize(e->d2h));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return RCG_OK;
}

static int process_range(rcg_engine* e, const float* const* in, uint64_t n_in, float* const* out, const RangeJob& job) {
    const int rc = process_range_impl(e, in, n_in, out, job);
    if (rc && job.my_signal) job.my_signal->set(-1);   // never leave the next shard waiting
    return rc;
}

// a host `apply` on top of the stack is called from one thread, in the reference's order: such jobs stay on one device
static bool top_kernel_is_host(const rcg_engine* e) { return !e->kernels.empty() && e->kernels.back().host != nullptr; }

static int process_all_group(rcg_engine* g, const float* const* in, uint64_t n_in, float* const* out, uint64_t calls,
                             uint64_t out_len) {
    if (g->watcher)   // fft.rs:78-81: take what the hot-swapper has built, for every shard, before the job starts
        if (int rc = drain_watcher(g, nullptr)) return rc;
    if ((g->cfg.flags & RCG_FLAG_FRAME_CLOCK) && !g->clock_base_set) rcg_set_clock_base(g, wall_clock_ms());   // one base for every shard
    const uint32_t C = g->cfg.channels;
    // a device must have something to amortise its launch, copies and thread on: >= 8 MB of samples in + out
    const uint64_t bytes = (uint64_t)C * (n_in + out_len) * sizeof(float);
    uint32_t G = (uint32_t)std::min<uint64_t>(g->shards.size(), std::max<uint64_t>(1, bytes >> 23));
    if (g->cfg.flags & RCG_FLAG_SHARD_ALWAYS) G = (uint32_t)g->shards.size();
    if (top_kernel_is_host(g->shards[0]) && !(g->cfg.flags & RCG_FLAG_BYPASS_RESYNTH)) G = 1;
    std::vector<RangeJob> jobs;
    std::vector<std::unique_ptr<TailSignal>> signals;
    const char* split = "none";
    if (G > 1 && C % G == 0) {            // by channel: no communication at all (main.rs:133-153)
        split = "channel";
        for (uint32_t i = 0; i < G; i++) {
            RangeJob j;
            j.ch0 = i * (C / G); j.nch = C / G; j.call0 = 0; j.call1 = calls;
            jobs.push_back(j);
        }
    } else if (G > 1 && calls >= 2) {     // by time: contiguous call ranges of every channel
        split = "time";
        G = (uint32_t)std::min<uint64_t>(G, calls);
        const bool exchange = !(g->cfg.flags & RCG_FLAG_SHARD_HALO);
        for (uint32_t i = 0; i < G; i++) signals.emplace_back(new TailSignal());
        for (uint32_t i = 0; i < G; i++) {
            RangeJob j;
            j.ch0 = 0; j.nch = C;
            j.call0 = calls * i / G; j.call1 = calls * (i + 1) / G;
            if (exchange && i > 0) { j.prev = g->shards[i - 1]; j.prev_signal = signals[i - 1].get(); }
            if (exchange && i + 1 < G) j.my_signal = signals[i].get();
            jobs.push_back(j);
        }
    } else {
        RangeJob j;
        j.ch0 = 0; j.nch = C; j.call0 = 0; j.call1 = calls;
        jobs.push_back(j);
    }
    g->part_devices = (uint32_t)jobs.size();
    g->part_split = jobs.size() > 1 ? split : "none";
    g->part_exchanges = 0;
    for (auto& j : jobs) if (j.prev) g->part_exchanges++;
    std::vector<int> rcs(jobs.size(), RCG_OK);
    std::vector<std::string> errs(jobs.size());
    auto work = [&](size_t i) {
        rcs[i] = process_range(g->shards[i], in, n_in, out, jobs[i]);
        if (rcs[i]) errs[i] = g_err;   // g_err is thread-local: carry the message back to the caller's thread
    };
    std::vector<std::thread> threads;
    for (size_t i = 1; i < jobs.size(); i++) threads.emplace_back(work, i);
    work(0);
    for (auto& t : threads) t.join();
    equalise_kernel_stacks(g);
    for (size_t i = 0; i < jobs.size(); i++)
        if (rcs[i]) return fail(rcs[i], "device %d: %s", g->shards[i]->device, errs[i].c_str());
    return RCG_OK;
}

extern "C" int rcg_process_all(rcg_engine* e, const float* const* in, uint64_t n_in, float* const* out,
                               uint64_t out_cap, uint64_t* n_out) {
    if (!e || !out || (!in && n_in)) return fail(RCG_ERR_INVALID_ARG, "null argument");
    const uint32_t C = e->cfg.channels;
    uint64_t calls, frames, out_len;
    counts_for(e->plan, n_in, &calls, &frames, &out_len);
    if (out_len > out_cap)
        return fail(RCG_ERR_CAPACITY, "output needs %llu samples per channel, capacity %llu",
                    (unsigned long long)out_len, (unsigned long long)out_cap);
    int rc;
    if (!e->shards.empty()) {
        rc = process_all_group(e, in, n_in, out, calls, out_len);
    } else {
        RangeJob j;
        j.ch0 = 0; j.nch = C; j.call0 = 0; j.call1 = calls;
        rc = process_range(e, in, n_in, out, j);
    }
    if (rc) return rc;
    if (n_out)
        for (uint32_t c = 0; c < C; c++) n_out[c] = out_len;
    return RCG_OK;
}

extern "C" int rcg_last_partition(const rcg_engine* e, uint32_t* devices_used, const char** split, uint32_t* exchanges) {
    if (!e) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (devices_used) *devices_used = e->part_devices;
    if (split) *split = e->part_split;
    if (exchanges) *exchanges = e->part_exchanges;
    return RCG_OK;
}

// -- streaming: Stretcher::next_window ---------------------------------------------------------
extern "C" int rcg_send(rcg_engine* e, uint32_t channel, const float* chunk, uint64_t n) {
    if (!e || channel >= e->cfg.channels || (!chunk && n)) return fail(RCG_ERR_INVALID_ARG, "bad channel or chunk");
    if (!e->shards.empty()) return rcg_send(e->shards[channel % e->shards.size()], channel, chunk, n);
    ChannelState& c = e->ch[channel];
    if (c.closed) return fail(RCG_ERR_STATE, "channel %u input is closed", channel);
    c.input.insert(c.input.end(), chunk, chunk + n);
    c.received += n;
    return RCG_OK;
}
extern "C" int rcg_close_input(rcg_engine* e, uint32_t channel) {
    if (!e || channel >= e->cfg.channels) return fail(RCG_ERR_INVALID_ARG, "bad channel");
    if (!e->shards.empty()) return rcg_close_input(e->shards[channel % e->shards.size()], channel);
    e->ch[channel].closed = true;
    return RCG_OK;
}
extern "C" int rcg_is_done(rcg_engine* e, uint32_t channel, int* done) {
    if (!e || channel >= e->cfg.channels || !done) return fail(RCG_ERR_INVALID_ARG, "bad channel");
    if (!e->shards.empty()) return rcg_is_done(e->shards[channel % e->shards.size()], channel, done);
    *done = e->ch[channel].done ? 1 : 0;
    return RCG_OK;
}

static int stream_compute(rcg_engine* e, uint32_t channel) {
    ChannelState& c = e->ch[channel];
    const uint64_t k = e->plan.frames_per_call, step = e->plan.sample_step_len, W = e->plan.window_len;
    const uint64_t H = W / 2, olen = e->plan.out_len_per_call;
    const uint64_t N = c.received;
    uint64_t lookahead = e->cfg.max_lookahead_calls ? e->cfg.max_lookahead_calls : 64;
    // calls whose every frame is fully covered by received input
    auto call_end_sample = [&](uint64_t call) { return ((call + 1) * k - 1) * step + W; };
    uint64_t n = 0;
    if (c.closed) {
        // recv() fails: frames are zero padded (stretcher.rs:129-131); run to the last call the
        // reference would make, or one call at a time beyond it
        uint64_t total;
        counts_for(e->plan, N, &total, nullptr, nullptr);
        n = c.next_call < total ? std::min(lookahead, total - c.next_call) : 1;
    } else {
        while (n < lookahead && call_end_sample(c.next_call + n) <= N) n++;
        if (n == 0) return fail(RCG_ERR_WOULD_BLOCK, "channel %u: call %llu needs %llu samples, %llu sent", channel,
                                (unsigned long long)c.next_call, (unsigned long long)call_end_sample(c.next_call),
                                (unsigned long long)N);
    }
    const uint64_t f_first = c.next_call * k;
    const uint64_t s_first = f_first * step;
    const uint64_t s_last = std::min<uint64_t>(N, call_end_sample(c.next_call + n - 1));
    const uint64_t n_samp = s_last > s_first ? s_last - s_first : 0;
    if (e->d_sin_cap < n_samp + 4) {
        cudaFree(e->d_sin); e->d_sin = nullptr;
        CUDA_TRY(cudaMalloc(&e->d_sin, (n_samp + 4) * 2 * sizeof(float)));
        e->d_sin_cap = (n_samp + 4) * 2;
    }
    if (e->d_sout_cap < n * olen) {
        cudaFree(e->d_sout); e->d_sout = nullptr;
        if (e->h_sout) cudaFreeHost(e->h_sout);
        e->h_sout = nullptr;
        CUDA_TRY(cudaMalloc(&e->d_sout, std::max<uint64_t>(n * olen, 4) * sizeof(float)));
        CUDA_TRY(cudaMallocHost(&e->h_sout, std::max<uint64_t>(n * olen, 4) * sizeof(float)));
        e->d_sout_cap = n * olen;
    }
    if (n_samp)
        CUDA_TRY(cudaMemcpyAsync(e->d_sin, c.input.data() + (s_first - c.input_base), n_samp * sizeof(float),
                                 cudaMemcpyHostToDevice, e->stream));
    rcg_segment s{};
    s.d_in = e->d_sin; s.in_stride = (int64_t)e->d_sin_cap; s.in_origin = (int64_t)s_first; s.in_valid = (int64_t)n_samp;
    s.d_out = e->d_sout; s.out_stride = (int64_t)e->d_sout_cap;
    s.first_call = c.next_call; s.n_calls = n;
    s.d_tail_in = c.next_call ? c.d_tail[c.tail_cur] : nullptr;
    s.d_tail_out = c.d_tail[c.tail_cur ^ 1];
    int rc = run_segment(e, s, 1, e->stream, (int)channel);
    if (rc) return rc;
    c.tail_cur ^= 1;
    CUDA_TRY(cudaMemcpyAsync(e->h_sout, e->d_sout, n * olen * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (uint64_t i = 0; i < n; i++) {
        const bool padded = c.closed && call_end_sample(c.next_call + i) > N;
        c.ready.emplace_back(std::vector<float>(e->h_sout + i * olen, e->h_sout + (i + 1) * olen), padded);
    }
    c.next_call += n;
    // drop consumed input (truncate_front at stretcher.rs:105-106)
    const uint64_t keep_from = std::min<uint64_t>(N, c.next_call * k * step);
    if (keep_from > c.input_base + (1u << 20) || (c.closed && keep_from >= N)) {
        c.input.erase(c.input.begin(), c.input.begin() + (keep_from - c.input_base));
        c.input_base = keep_from;
    }
    (void)H;
    return RCG_OK;
}

extern "C" int rcg_next_window(rcg_engine* e, uint32_t channel, float* out, uint64_t cap, uint64_t* n_out) {
    if (!e || channel >= e->cfg.channels || !out) return fail(RCG_ERR_INVALID_ARG, "bad channel or buffer");
    if (!e->shards.empty()) {   // every channel is an independent Stretcher (main.rs:133-153): device = channel mod n
        if (e->watcher)
            if (int rc = drain_watcher(e, nullptr)) return rc;
        const int rc = rcg_next_window(e->shards[channel % e->shards.size()], channel, out, cap, n_out);
        equalise_kernel_stacks(e);
        return rc;
    }
    if (cap < e->plan.out_len_per_call)
        return fail(RCG_ERR_CAPACITY, "window needs %llu samples", (unsigned long long)e->plan.out_len_per_call);
    CUDA_TRY(cudaSetDevice(e->device));
    ChannelState& c = e->ch[channel];
    if (c.ready.empty()) {
        int rc = stream_compute(e, channel);
        if (rc) return rc;
    }
    auto& w = c.ready.front();
    memcpy(out, w.first.data(), w.first.size() * sizeof(float));
    if (n_out) *n_out = w.first.size();
    if (w.second) c.done = true;  // stretcher.rs:131
    c.ready.pop_front();
    return RCG_OK;
}

// -- plugins -------------------------------------------------------------------------------------
// With an even pitch multiple the fused path carries even samples only between launches; the
// first kernel must therefore be pushed before any audio has been processed.
static int check_kernel_push(rcg_engine* e) {
    if (log2_exact(e->cfg.window_len) < 0)
        return fail(RCG_ERR_STATE, "frequency kernels need a power-of-two window (window_len %u runs the chirp-z path)", e->cfg.window_len);
    const bool had = e->shards.empty() ? e->had_kernels : e->shards[0]->had_kernels;
    if (total_launches(e) > 0 && !had && use_half_rate(e->cfg.pitch_multiple, 0u))
        return fail(RCG_ERR_STATE, "with an even pitch multiple the first frequency kernel must be pushed before processing starts");
    return RCG_OK;
}

extern "C" int rcg_push_frequency_kernel(rcg_engine* e, rcg_apply_fn fn) {
    if (!e || !fn) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (int rc = check_kernel_push(e)) return rc;
    if (!e->shards.empty()) {
        for (rcg_engine* sh : e->shards)
            if (int rc = rcg_push_frequency_kernel(sh, fn)) return rc;
        return RCG_OK;
    }
    FreqKernel k;
    k.host = fn;
    e->kernels.push_back(k);
    e->had_kernels = true;
    return RCG_OK;
}
extern "C" int rcg_load_frequency_kernel(rcg_engine* e, const char* so_path) {
    if (!e || !so_path) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (int rc = check_kernel_push(e)) return rc;
    void* lib = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) return fail(RCG_ERR_PLUGIN, "dlopen(%s): %s", so_path, dlerror());
    void* sym = dlsym(lib, "apply");
    if (!sym) {
        dlclose(lib);
        return fail(RCG_ERR_PLUGIN, "%s has no `apply` symbol", so_path);
    }
    e->libs.push_back(lib);   // (a multi-device engine keeps the handle itself and hands `apply` to every shard)
    if (!e->shards.empty()) {
        for (rcg_engine* sh : e->shards)
            if (int rc = rcg_push_frequency_kernel(sh, (rcg_apply_fn)sym)) return rc;
        return RCG_OK;
    }
    FreqKernel k;
    k.host = (rcg_apply_fn)sym;
    e->kernels.push_back(k);
    e->had_kernels = true;
    return RCG_OK;
}
// ---- device frequency kernels: CUDA source -> NVRTC -> cubin for sm_100a -----------------------------
// (the role of hotswapper.rs:58-86, which shells out to rustc; libnvrtc is loaded on first use so
// that the library itself has no NVRTC dependency)
namespace {
typedef void* nvrtcProgram_t;
struct Nvrtc {
    void* lib = nullptr;
    int (*create)(nvrtcProgram_t*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*compile)(nvrtcProgram_t, int, const char* const*) = nullptr;
    int (*log_size)(nvrtcProgram_t, size_t*) = nullptr;
    int (*get_log)(nvrtcProgram_t, char*) = nullptr;
    int (*cubin_size)(nvrtcProgram_t, size_t*) = nullptr;
    int (*get_cubin)(nvrtcProgram_t, char*) = nullptr;
    int (*destroy)(nvrtcProgram_t*) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
        for (const char* n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!lib) return false;
        create = (decltype(create))dlsym(lib, "nvrtcCreateProgram");
        compile = (decltype(compile))dlsym(lib, "nvrtcCompileProgram");
        log_size = (decltype(log_size))dlsym(lib, "nvrtcGetProgramLogSize");
        get_log = (decltype(get_log))dlsym(lib, "nvrtcGetProgramLog");
        cubin_size = (decltype(cubin_size))dlsym(lib, "nvrtcGetCUBINSize");
        get_cubin = (decltype(get_cubin))dlsym(lib, "nvrtcGetCUBIN");
        destroy = (decltype(destroy))dlsym(lib, "nvrtcDestroyProgram");
        return create && compile && log_size && get_log && cubin_size && get_cubin && destroy;
    }
} g_nvrtc;

const char* kDeviceWrapper = R"(
// spectra [channel][frame][bin]; spectrum f of the buffer is frame frame0 + f % frames_per_channel of its channel,
// and that frame's clock is elapsed_ms + frame * clock_num / clock_den (clock_num = 0: one clock for the launch)
extern "C" __global__ void rcg_apply_kernel(const float2* in, float2* out, unsigned n_bins,
                                            unsigned long long n_frames, unsigned long long elapsed_ms,
                                            unsigned long long frame0, unsigned long long frames_per_channel,
                                            unsigned long long clock_num, unsigned long long clock_den) {
    const unsigned long long total = n_frames * n_bins;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long f = i / n_bins;
        const unsigned b = (unsigned)(i - f * n_bins);
        const unsigned long long ms = elapsed_ms + (clock_num ? (frame0 + f % frames_per_channel) * clock_num / clock_den : 0ull);
        out[i] = apply((size_t)ms, b, n_bins, in + f * n_bins);
    }
}
)";
// the bin-local form as a plain kernel over spectra in memory (same entry point and arguments as the gather form)
const char* kLocalTwinWrapper = R"(
extern "C" __global__ void rcg_apply_kernel(const float2* in, float2* out, unsigned n_bins,
                                            unsigned long long n_frames, unsigned long long elapsed_ms,
                                            unsigned long long frame0, unsigned long long frames_per_channel,
                                            unsigned long long clock_num, unsigned long long clock_den) {
    const unsigned long long total = n_frames * n_bins;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long f = i / n_bins;
        const unsigned b = (unsigned)(i - f * n_bins);
        const unsigned long long ms = elapsed_ms + (clock_num ? (frame0 + f % frames_per_channel) * clock_num / clock_den : 0ull);
        out[i] = apply((size_t)ms, b, n_bins, in[i]);
    }
}
)";
}  // namespace

// user source -> cubin for sm_100a; no CUDA context needed (runs on the watcher thread too)
static bool compile_device_source(const std::string& user_src, std::vector<char>& cubin, std::string& log,
                                  const char* wrapper = kDeviceWrapper) {
    static std::mutex load_mutex;
    {
        std::lock_guard<std::mutex> g(load_mutex);
        if (!g_nvrtc.load()) { log = std::string("libnvrtc could not be loaded: ") + (dlerror() ? dlerror() : ""); return false; }
    }
    const std::string src = user_src + wrapper;
    nvrtcProgram_t prog = nullptr;
    if (g_nvrtc.create(&prog, src.c_str(), "frequency_kernel.cu", 0, nullptr, nullptr) != 0) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device"};
    const int cr = g_nvrtc.compile(prog, 3, opts);
    if (cr != 0) {
        size_t n = 0;
        g_nvrtc.log_size(prog, &n);
        log.assign(n + 1, '\0');
        if (n) g_nvrtc.get_log(prog, &log[0]);
        g_nvrtc.destroy(&prog);
        return false;
    }
    size_t cn = 0;
    g_nvrtc.cubin_size(prog, &cn);
    cubin.resize(cn);
    g_nvrtc.get_cubin(prog, cubin.data());
    g_nvrtc.destroy(&prog);
    return true;
}

static int install_cubin(rcg_engine* e, const std::vector<char>& cubin) {
    FreqKernel k;
    CUDA_TRY(cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaError_t ke = cudaLibraryGetKernel(&k.dev, k.lib, "rcg_apply_kernel");
    if (ke != cudaSuccess) {
        cudaLibraryUnload(k.lib);
        return fail(RCG_ERR_PLUGIN, "compiled frequency kernel has no entry point: %s", cudaGetErrorString(ke));
    }
    e->kernels.push_back(k);
    e->had_kernels = true;
    return RCG_OK;
}

extern "C" int rcg_push_device_frequency_kernel(rcg_engine* e, const char* cuda_source) {
    if (!e || !cuda_source) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (int rc = check_kernel_push(e)) return rc;
    std::vector<char> cubin;
    std::string log;
    if (!compile_device_source(cuda_source, cubin, log))   // compile failure keeps the previous kernel (hotswapper.rs:36-42)
        return fail(RCG_ERR_PLUGIN, "device frequency kernel failed to compile: %.400s", log.c_str());
    return install_everywhere(e, cubin, nullptr);
}

// ---- bin-local device kernels: NVRTC builds the whole stretch kernel around the user's apply ----------------
// The scalar configuration a fused kernel may assume at compile time.  Even positive pitch multiples run the half-rate
// path only until a kernel that needs spectra in memory is installed, so their mode can change: generic build.
static int fused_kernel_mode(const rcg_config& cfg) {
    if (cfg.flags & RCG_FLAG_BYPASS_RESYNTH) return 0;
    if (use_half_rate(cfg.pitch_multiple, 0u)) return 0;
    return stretch_mode_of(cfg.pitch_multiple, 0u);
}

#include "rcg_jit_sources.inc"

// mode: the launch configuration to fold in as constants (rcg_jit_entry.h), 0 = none
static bool compile_fused_source(const std::string& user_src, int log2h, int mode, std::vector<char>& cubin,
                                 std::vector<char>& twin, std::string& log) {
    // the per-bin twin first (small): it takes over when the kernel above this one fails in the middle of a batch
    if (!compile_device_source(user_src, twin, log, kLocalTwinWrapper)) return false;
    static std::mutex load_mutex;
    {
        std::lock_guard<std::mutex> g(load_mutex);
        if (!g_nvrtc.load()) { log = "libnvrtc could not be loaded"; return false; }
    }
    // the kernel NVRTC builds reads the same Tensor Memory image and tables as the library's own kernels: it must see
    // the build-time switches the library was compiled with (they decide what the image holds)
#define RCG_JIT_STR2(x) #x
#define RCG_JIT_STR(x) RCG_JIT_STR2(x)
    static const char kBuildSwitches[] =
        "\n#define RCG_TW_TMEM " RCG_JIT_STR(RCG_TW_TMEM) "\n#define RCG_TW_TMEM_512 " RCG_JIT_STR(RCG_TW_TMEM_512)
        "\n#define RCG_TW_TMEM_SMALL " RCG_JIT_STR(RCG_TW_TMEM_SMALL) "\n#define RCG_WIN_TMEM " RCG_JIT_STR(RCG_WIN_TMEM)
        "\n#define RCG_TMEM_CHUNK " RCG_JIT_STR(RCG_TMEM_CHUNK) "\n#define RCG_PAIR_SPLIT_LOAD " RCG_JIT_STR(RCG_PAIR_SPLIT_LOAD) "\n";
    const std::string src = user_src + kBuildSwitches + "#define RCG_JIT_LOG2H " + std::to_string(log2h) +
                            "\n#define RCG_JIT_MODE " + std::to_string(mode) + "\n#include \"rcg_jit_entry.h\"\n";
    const int n_headers = (int)(sizeof(kJitHeaderNames) / sizeof(kJitHeaderNames[0]));
    nvrtcProgram_t prog = nullptr;
    if (g_nvrtc.create(&prog, src.c_str(), "fused_frequency_kernel.cu", n_headers, kJitHeaderSources, kJitHeaderNames) != 0) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device"};
    const int cr = g_nvrtc.compile(prog, 3, opts);
    if (cr != 0) {
        size_t n = 0;
        g_nvrtc.log_size(prog, &n);
        log.assign(n + 1, '\0');
        if (n) g_nvrtc.get_log(prog, &log[0]);
        g_nvrtc.destroy(&prog);
        return false;
    }
    size_t cn = 0;
    g_nvrtc.cubin_size(prog, &cn);
    cubin.resize(cn);
    g_nvrtc.get_cubin(prog, cubin.data());
    g_nvrtc.destroy(&prog);
    return true;
}

static int install_fused_cubin(rcg_engine* e, const std::vector<char>& cubin, const std::vector<char>& twin) {
    FreqKernel k;
    CUDA_TRY(cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaError_t ke = cudaLibraryGetKernel(&k.fused, k.lib, "rcg_jit_stretch");
    if (ke == cudaSuccess)
        ke = cudaFuncSetAttribute((const void*)k.fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->geo.smem_bytes);
    if (ke == cudaSuccess) ke = cudaLibraryLoadData(&k.twin_lib, twin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ke == cudaSuccess) ke = cudaLibraryGetKernel(&k.dev, k.twin_lib, "rcg_apply_kernel");
    if (ke != cudaSuccess) {
        (void)cudaGetLastError();
        cudaLibraryUnload(k.lib);
        if (k.twin_lib) cudaLibraryUnload(k.twin_lib);
        return fail(RCG_ERR_PLUGIN, "fused frequency kernel could not be loaded: %s", cudaGetErrorString(ke));
    }
    e->kernels.push_back(k);
    return RCG_OK;
}

// load a compiled device kernel on the engine's device -- on every shard's device for a multi-device engine.
// twin != nullptr: the bin-local form (cubin = the fused stretch kernel, *twin = its per-bin twin).
static int install_everywhere(rcg_engine* e, const std::vector<char>& cubin, const std::vector<char>* twin) {
    if (!e->shards.empty()) {
        for (rcg_engine* sh : e->shards)
            if (int rc = install_everywhere(sh, cubin, twin)) return rc;
        return RCG_OK;
    }
    CUDA_TRY(cudaSetDevice(e->device));
    return twin ? install_fused_cubin(e, cubin, *twin) : install_cubin(e, cubin);
}

extern "C" int rcg_push_local_device_frequency_kernel(rcg_engine* e, const char* cuda_source) {
    if (!e || !cuda_source) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (log2_exact(e->cfg.window_len) < 0)
        return fail(RCG_ERR_STATE, "frequency kernels need a power-of-two window (window_len %u runs the chirp-z path)", e->cfg.window_len);
    std::vector<char> cubin, twin;
    std::string log;
    if (!compile_fused_source(cuda_source, e->log2w - 1, fused_kernel_mode(e->cfg), cubin, twin, log))   // compile failure keeps the previous kernel
        return fail(RCG_ERR_PLUGIN, "bin-local frequency kernel failed to compile: %.400s", log.c_str());
    return install_everywhere(e, cubin, &twin);
}

// ---- hot swap from a watched file (hotswapper.rs:14-33 + fft.rs:78-81) ------------------------------
static int drain_watcher(rcg_engine* e, uint32_t* swapped) {
    uint32_t n = 0;
    rcg::BuiltKernel b;
    int status = RCG_OK;
    while (e->watcher->try_recv(b)) {
        if (b.host) {
            e->libs.push_back(b.dl);       // every library stays loaded (fft.rs:22)
            if (e->shards.empty()) {
                FreqKernel k;
                k.host = (rcg_apply_fn)b.host;
                e->kernels.push_back(k);
                e->had_kernels = true;
            } else {
                for (rcg_engine* sh : e->shards) rcg_push_frequency_kernel(sh, (rcg_apply_fn)b.host);
            }
            n++;
        } else {
            const int rc = install_everywhere(e, b.cubin, b.local ? &b.twin : nullptr);
            if (rc) status = rc; else n++;
        }
        b = rcg::BuiltKernel();
    }
    e->swaps += n;
    if (swapped) *swapped = n;
    return status;
}

extern "C" int rcg_watch_frequency_kernel(rcg_engine* e, const char* path, uint32_t poll_ms) {
    if (!e || !path) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (e->watcher) return fail(RCG_ERR_STATE, "this engine already watches %s", e->watcher->path().c_str());
    if (int rc = check_kernel_push(e)) return rc;
    if (rcg::KernelWatcher::kind_of(path) == rcg::KernelWatcher::kUnknown)
        return fail(RCG_ERR_PLUGIN, "frequency kernel %s: want a .cu, .c, .cc, .cpp, .so or .rs file", path);
    // from here on this engine runs the full-rate path: a kernel may arrive between any two batches
    e->had_kernels = true;
    for (rcg_engine* sh : e->shards) sh->had_kernels = true;
    const int log2h = e->log2w - 1, mode = fused_kernel_mode(e->cfg);
    e->watcher.reset(new rcg::KernelWatcher(path, poll_ms,
        [log2h, mode](const std::string& src, bool local, std::vector<char>& cubin, std::vector<char>& twin, std::string& log) {
            return local ? compile_fused_source(src, log2h, mode, cubin, twin, log) : compile_device_source(src, cubin, log);
        }));
    e->watcher->start();   // first build happens here, on the caller's thread (hotswapper.rs:17)
    return drain_watcher(e, nullptr);
}

extern "C" int rcg_poll_frequency_kernels(rcg_engine* e, uint32_t* swapped) {
    if (!e) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (swapped) *swapped = 0;
    if (!e->watcher) return RCG_OK;
    return drain_watcher(e, swapped);
}

extern "C" int rcg_watch_stats(const rcg_engine* e, uint32_t* builds_ok, uint32_t* builds_failed, uint32_t* swaps) {
    if (!e) return fail(RCG_ERR_INVALID_ARG, "null argument");
    if (builds_ok) *builds_ok = e->watcher ? e->watcher->builds_ok() : 0;
    if (builds_failed) *builds_failed = e->watcher ? e->watcher->builds_failed() : 0;
    if (swaps) *swaps = e->swaps;
    return RCG_OK;
}

extern "C" int rcg_watch_log(const rcg_engine* e, char* buf, size_t cap) {
    if (!e || !buf || !cap) return fail(RCG_ERR_INVALID_ARG, "null argument");
    const std::string log = e->watcher ? e->watcher->last_log() : std::string();
    snprintf(buf, cap, "%s", log.c_str());
    return RCG_OK;
}

extern "C" int rcg_set_clock_base(rcg_engine* e, uint64_t base_ms) {
    if (!e) return fail(RCG_ERR_INVALID_ARG, "null argument");
    e->clock_base = base_ms;
    e->clock_base_set = true;
    for (rcg_engine* sh : e->shards) { sh->clock_base = base_ms; sh->clock_base_set = true; }
    return RCG_OK;
}

extern "C" int rcg_frequency_kernel_count(const rcg_engine* e, uint32_t* n) {
    if (!e || !n) return fail(RCG_ERR_INVALID_ARG, "null argument");
    *n = (uint32_t)(e->shards.empty() ? e->kernels.size() : e->shards[0]->kernels.size());
    return RCG_OK;
}

// -- helpers -------------------------------------------------------------------------------------
extern "C" int rcg_device_count(int* n) {
    if (!n) return fail(RCG_ERR_INVALID_ARG, "null argument");
    *n = 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { (void)cudaGetLastError(); return RCG_OK; }
    for (int d = 0; d < ndev; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) (*n)++;
    }
    return RCG_OK;
}
extern "C" int rcg_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(RCG_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 1));
    return RCG_OK;
}
extern "C" int rcg_host_free(void* ptr) {
    CUDA_TRY(cudaFreeHost(ptr));
    return RCG_OK;
}

// rocoder_gpu -- command-line front end with the reference's flags (src/main.rs:27-122) for the
// file-to-file path (main.rs:124-160, 190-204):  rocoder_gpu -i in.wav -o out.wav -f 8 -p 2 -w 16384
// Recording (no -i) and live playback (no -o) belong to the reference's audio-device layer and are
// not provided: both are reported as errors.  Extra flags: --seed N (phase seed; default: random,
// like rand::thread_rng), --device N or --device 0,1,... (the job is partitioned over every listed GPU: by channel
// when the channels divide evenly, else by time with one carried-half-window exchange per boundary), --bench
// (print timing as JSON).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../include/rocoder_b200.h"
#include "../rocoder_b200/host/rocoder_audio.hpp"

using namespace rocoder;

struct Opt {                       // main.rs:27-122
    uint32_t window_len = 16384;   // -w, --window
    std::string buffer_dur = "1";  // -b, --buffer
    float factor = 1.0f;           // -f, --factor
    int pitch_multiple = 1;        // -p, --pitch_multiple
    float amplitude = 1.0f;        // -a, --amplitude
    std::string input;             // -i, --input   ('-' = stdin)
    bool rotate_channels = false;  // --rotate-channels
    std::string freq_kernel;       // --freq-kernel
    std::string fade = "1";        // -x, --fade  (only used by live playback in the reference, main.rs:195-208)
    std::string start, duration;   // -s, -d
    std::string output;            // -o, --output
    bool has_seed = false;
    uint64_t seed = 0;
    std::vector<int> devices{0};   // --device 0,1,...
    bool bench = false;
};

static void usage() {
    std::fputs(
        "rocoder_gpu: B200 engine for rocoder's stretch path (file mode)\n"
        "  -w, --window <n>           Processing window size [16384]\n"
        "  -b, --buffer <dur>         Maximum amount of audio to process ahead of time [1]\n"
        "  -f, --factor <x>           Stretch factor; e.g. 5 to slow 5x and 0.2 to speed up 5x [1]\n"
        "  -p, --pitch_multiple <n>   A non-zero integer pitch multiplier [1]\n"
        "  -a, --amplitude <x>        Output amplitude [1]\n"
        "  -i, --input <path>         A .wav file ('-' for stdin)\n"
        "      --rotate-channels      Rotate the input audio channels\n"
        "      --freq-kernel <path>   Frequency kernel, rebuilt and swapped in whenever the file changes: a .cu\n"
        "                             (device `apply`), .c/.cpp (host `apply`), .so exporting the C-ABI `apply`, or a\n"
        "                             reference-style .rs kernel (needs rustc)\n"
        "  -x, --fade <dur>           (live playback only in the reference; accepted and ignored)\n"
        "  -s, --start <dur>          Start time in input audio (hh:mm:ss.ss)\n"
        "  -d, --duration <dur>       Duration to use from input audio (hh:mm:ss.ss)\n"
        "  -o, --output <path>        Output .wav file path. Uses 32-bit float.\n"
        "      --seed <n>  --device <n[,n...]>  --bench\n", stderr);
}

static bool parse(int argc, char** argv, Opt& o) {
    auto need = [&](int& i) -> const char* { if (i + 1 >= argc) throw AudioError(std::string("missing value for ") + argv[i]); return argv[++i]; };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], v;
        size_t eq = a.find('=');
        if (a.rfind("--", 0) == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); }
        auto val = [&]() -> std::string { return v.empty() ? std::string(need(i)) : v; };
        if (a == "-w" || a == "--window") o.window_len = (uint32_t)std::stoul(val());
        else if (a == "-b" || a == "--buffer") o.buffer_dur = val();
        else if (a == "-f" || a == "--factor") o.factor = std::stof(val());
        else if (a == "-p" || a == "--pitch_multiple") {
            long p = std::stol(val());
            if (p < -128 || p > 127) throw AudioError("pitch_multiple does not fit i8");
            o.pitch_multiple = (int)p;
        }
        else if (a == "-a" || a == "--amplitude") o.amplitude = std::stof(val());
        else if (a == "-i" || a == "--input") o.input = val();
        else if (a == "--rotate-channels") o.rotate_channels = true;
        else if (a == "--freq-kernel") o.freq_kernel = val();
        else if (a == "-x" || a == "--fade") o.fade = val();
        else if (a == "-s" || a == "--start") o.start = val();
        else if (a == "-d" || a == "--duration") o.duration = val();
        else if (a == "-o" || a == "--output") o.output = val();
        else if (a == "--seed") { o.seed = std::stoull(val()); o.has_seed = true; }
        else if (a == "--device") {
            o.devices.clear();
            std::string list = val();
            for (size_t pos = 0; pos <= list.size();) {
                const size_t comma = std::min(list.find(',', pos), list.size());
                o.devices.push_back(std::stoi(list.substr(pos, comma - pos)));
                pos = comma + 1;
            }
            if (o.devices.empty() || o.devices.size() > RCG_MAX_DEVICES) throw AudioError("--device: 1 to 16 ordinals");
        }
        else if (a == "--bench") o.bench = true;
        else if (a == "-h" || a == "--help") { usage(); return false; }
        else throw AudioError("unknown argument " + a);
    }
    return true;
}

int main(int argc, char** argv) {
    try {
        Opt opt;
        if (!parse(argc, argv, opt)) return 0;
        if (opt.input.empty()) throw AudioError("recording from an input device is not supported here: pass -i <file.wav>");
        if (opt.output.empty()) throw AudioError("live playback is not supported here: pass -o <file.wav>");
        parse_duration_ms(opt.fade);
        const uint64_t buffer_ms = parse_duration_ms(opt.buffer_dur);
        // load_audio, main.rs:162-188
        Audio audio = opt.input == "-" ? read_wav(std::cin) : read_wav_file(opt.input);
        if (!opt.start.empty() || !opt.duration.empty()) {
            uint64_t s = 0, d = 0;
            if (!opt.start.empty()) s = parse_duration_ms(opt.start);
            if (!opt.duration.empty()) d = parse_duration_ms(opt.duration);
            audio.clip_in_place(opt.start.empty() ? nullptr : &s, opt.duration.empty() ? nullptr : &d);
        }
        if (opt.rotate_channels) audio.rotate_channels();
        const size_t n_in = audio.data.empty() ? 0 : audio.data[0].size();

        rcg_config cfg{};
        cfg.struct_size = sizeof(cfg);
        cfg.channels = audio.spec.channels;
        cfg.sample_rate = audio.spec.sample_rate;
        cfg.window_len = opt.window_len;
        cfg.factor = opt.factor;
        cfg.amplitude = opt.amplitude;
        cfg.pitch_multiple = opt.pitch_multiple;
        cfg.buffer_dur_s = (float)buffer_ms / 1000.0f;
        cfg.seed = opt.seed;
        cfg.device = opt.devices[0];
        cfg.n_devices = (uint32_t)opt.devices.size();
        for (size_t i = 0; i < opt.devices.size(); i++) cfg.devices[i] = opt.devices[i];
        cfg.flags = opt.has_seed ? 0u : RCG_FLAG_RANDOM_SEED;
        rcg_engine* eng = nullptr;
        if (rcg_create(&cfg, &eng) != RCG_OK) throw AudioError(rcg_last_error());
        if (!opt.freq_kernel.empty()) {
            // ReFFT::new -> hotswapper::hotswap (fft.rs:31-35): build now, rebuild and swap whenever the file changes
            if (rcg_watch_frequency_kernel(eng, opt.freq_kernel.c_str(), 0) != RCG_OK) throw AudioError(rcg_last_error());
            uint32_t n = 0, bad = 0;
            rcg_frequency_kernel_count(eng, &n);
            rcg_watch_stats(eng, nullptr, &bad, nullptr);
            if (n == 0 && bad) {   // like the reference: warn and pass frames through until a build succeeds
                char log[2048];
                rcg_watch_log(eng, log, sizeof log);
                fprintf(stderr, "warning: failed to compile %s\n%s\n", opt.freq_kernel.c_str(), log);
            }
        }
        uint64_t calls = 0, frames = 0, out_len = 0;
        if (rcg_expected_counts(&cfg, n_in, &calls, &frames, &out_len) != RCG_OK) throw AudioError(rcg_last_error());
        std::vector<std::vector<float>> out(cfg.channels, std::vector<float>(out_len));
        std::vector<const float*> in_ptrs;
        std::vector<float*> out_ptrs;
        for (auto& c : audio.data) in_ptrs.push_back(c.data());
        for (auto& c : out) out_ptrs.push_back(c.data());
        std::vector<uint64_t> n_out(cfg.channels);
        const auto t0 = std::chrono::steady_clock::now();
        if (rcg_process_all(eng, in_ptrs.data(), n_in, out_ptrs.data(), out_len, n_out.data()) != RCG_OK)
            throw AudioError(rcg_last_error());
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint32_t used = 1, exchanges = 0;
        const char* split = "none";
        rcg_last_partition(eng, &used, &split, &exchanges);
        rcg_destroy(eng);
        write_wav_file(opt.output, out, audio.spec.sample_rate);   // handle_result, main.rs:195-204
        std::fprintf(stderr, "stretch process completed: %llu calls, %llu frames, %llu samples/channel in %.3f s\n",
                     (unsigned long long)calls, (unsigned long long)frames, (unsigned long long)out_len, secs);
        if (opt.bench)
            std::printf("{\"calls\": %llu, \"frames\": %llu, \"out_samples_per_channel\": %llu, \"channels\": %u, \"seconds\": %.6f, "
                        "\"devices_used\": %u, \"split\": \"%s\", \"exchanges\": %u}\n",
                        (unsigned long long)calls, (unsigned long long)frames, (unsigned long long)out_len, cfg.channels, secs,
                        used, split, exchanges);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rocoder_gpu: %s\n", e.what());
        return 1;
    }
}

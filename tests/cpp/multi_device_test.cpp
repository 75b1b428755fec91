// One engine, several devices (rcg_config.n_devices): rcg_process_all partitions the job by channel or by time and
// the output must be bit-identical to the single-device run for every device list -- the reference's shape is one
// call per file for all channels (main.rs:133-156).  Plain C ABI, no torch.  The same ordinal may repeat, so the
// partition, the carried-half-window exchange and the halo variant are exercised on a one-GPU box too; the real
// devices of the box are added when it has more than one.  Built and run by tests/test_multi_device.py.
//   usage: multi_device_test [max_real_devices]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rocoder_b200.h"

struct Case { const char* name; uint32_t channels, window; float factor; int pitch; uint64_t n_in; uint32_t flags; };

static std::vector<std::vector<float>> make_input(uint32_t channels, uint64_t n, uint64_t seed) {
    std::vector<std::vector<float>> x(channels, std::vector<float>(n));
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
    for (auto& ch : x)
        for (auto& v : ch) {   // SplitMix64 -> [-0.5, 0.5)
            s += 0x9E3779B97F4A7C15ull;
            uint64_t z = s;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            v = (float)((double)(z >> 11) / 9007199254740992.0 - 0.5);
        }
    return x;
}

static int run(const Case& c, const std::vector<int>& devices, uint32_t extra_flags, const std::vector<std::vector<float>>& x,
               std::vector<std::vector<float>>& out, uint32_t* used, std::string* split, uint32_t* exchanges) {
    rcg_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.channels = c.channels; cfg.sample_rate = 48000; cfg.window_len = c.window; cfg.factor = c.factor;
    cfg.amplitude = 1.0f; cfg.pitch_multiple = c.pitch; cfg.buffer_dur_s = 1.0f; cfg.seed = 77;
    cfg.flags = c.flags | extra_flags;
    cfg.device = devices[0];
    cfg.n_devices = (uint32_t)devices.size();
    for (size_t i = 0; i < devices.size(); i++) cfg.devices[i] = devices[i];
    rcg_engine* e = nullptr;
    if (rcg_create(&cfg, &e) != RCG_OK) { fprintf(stderr, "rcg_create: %s\n", rcg_last_error()); return 1; }
    uint64_t calls = 0, frames = 0, out_len = 0;
    rcg_expected_counts(&cfg, c.n_in, &calls, &frames, &out_len);
    out.assign(c.channels, std::vector<float>(out_len, -12345.0f));
    std::vector<const float*> ip; std::vector<float*> op;
    for (auto& v : x) ip.push_back(v.data());
    for (auto& v : out) op.push_back(v.data());
    std::vector<uint64_t> n_out(c.channels);
    int rc = rcg_process_all(e, ip.data(), c.n_in, op.data(), out_len, n_out.data());
    if (rc != RCG_OK) { fprintf(stderr, "rcg_process_all: %s\n", rcg_last_error()); rcg_destroy(e); return 2; }
    const char* sp = "";
    rcg_last_partition(e, used, &sp, exchanges);
    *split = sp;
    rcg_destroy(e);
    return 0;
}

int main(int argc, char** argv) {
    int real = 0;
    if (rcg_device_count(&real) != RCG_OK || real < 1) { fprintf(stderr, "no sm_100 device\n"); return 77; }
    if (argc > 1) real = std::min(real, atoi(argv[1]));
    const Case cases[] = {
        {"time split, pitch 1", 2, 1024, 2.0f, 1, 300000, 0},
        {"time split, half-rate pitch 2, W=16384", 2, 16384, 8.0f, 2, 700000, 0},
        {"time split, pitch -3 (drops H-S samples per call)", 2, 4096, 20.0f, -3, 400000, 0},
        {"time split, three channels, pitch 3", 3, 2048, 1.5f, 3, 250000, 0},
        {"channel split, 8 channels", 8, 4096, 2.0f, 1, 120000, 0},
        {"time split, W=65536 (CTA pair)", 2, 65536, 2.0f, 1, 1500000, 0},
        {"index path (bypass), pitch -2", 2, 512, 3.0f, -2, 90000, RCG_FLAG_BYPASS_RESYNTH},
    };
    std::vector<std::vector<int>> lists = {{0, 0}, {0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}};
    if (real >= 2) {
        std::vector<int> all;
        for (int d = 0; d < real; d++) all.push_back(d);
        lists.push_back({0, 1});
        if (real > 2) lists.push_back(all);
    }
    int failures = 0, checks = 0;
    for (const Case& c : cases) {
        const auto x = make_input(c.channels, c.n_in, c.window + c.channels);
        std::vector<std::vector<float>> ref, got;
        uint32_t used = 0, ex = 0;
        std::string split;
        if (run(c, {0}, 0, x, ref, &used, &split, &ex)) return 1;
        if (used != 1 || split != "none") { fprintf(stderr, "%s: single device reported %u devices (%s)\n", c.name, used, split.c_str()); failures++; }
        for (const auto& list : lists)
            for (uint32_t mode : {0u, (uint32_t)RCG_FLAG_SHARD_HALO}) {
                if (run(c, list, RCG_FLAG_SHARD_ALWAYS | mode, x, got, &used, &split, &ex)) return 1;
                checks++;
                bool same = got.size() == ref.size();
                for (size_t ch = 0; same && ch < ref.size(); ch++)
                    same = got[ch].size() == ref[ch].size() && memcmp(got[ch].data(), ref[ch].data(), ref[ch].size() * sizeof(float)) == 0;
                const bool by_channel = c.channels % list.size() == 0;
                const bool shape_ok = used == list.size() && split == (by_channel ? "channel" : "time") &&
                                      ex == ((by_channel || mode) ? 0u : (uint32_t)list.size() - 1);
                printf("%-52s %zu device(s)%s: %s split, %u exchange(s): %s\n", c.name, list.size(), mode ? " halo" : "",
                       split.c_str(), ex, same && shape_ok ? "bit-identical" : "MISMATCH");
                if (!same || !shape_ok) failures++;
            }
    }
    // a small job stays on one device unless forced
    {
        const Case c = {"small job", 2, 1024, 2.0f, 1, 50000, 0};
        const auto x = make_input(2, c.n_in, 3);
        std::vector<std::vector<float>> got;
        uint32_t used = 0, ex = 0;
        std::string split;
        if (run(c, {0, 0}, 0, x, got, &used, &split, &ex)) return 1;
        if (used != 1) { fprintf(stderr, "small job used %u devices\n", used); failures++; }
    }
    printf("%d checks, %d failure(s), %d real device(s)\n", checks, failures, real);
    return failures ? 1 : 0;
}

// Exercises include/rocoder_b200.hpp the way the reference wires the path (main.rs:133-160) and
// re-states the reference's two Stretcher tests (stretcher.rs:145-161).  Built and run by
// tests/test_cpp_mirror.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "rocoder_b200.hpp"

using namespace rocoder;

static std::pair<Stretcher, Sender<std::vector<float>>> basic_stretcher(size_t window_len) {
    // stretcher.rs:163-179
    auto ch = unbounded<std::vector<float>>();
    Stretcher st(AudioSpec{2, 44100}, std::move(ch.second), 1.0f, 1.0f, 1, window_len, std::chrono::duration<float>(1.0f));
    return {std::move(st), std::move(ch.first)};
}

static int test_ensure() {
    {   // ensure_input_samples_available_when_channel_closed_fills_with_zeros
        auto ch = unbounded<std::vector<float>>();
        Stretcher st(AudioSpec{2, 44100}, std::move(ch.second), 1.0f, 1.0f, 1, 1024, std::chrono::duration<float>(1.0f));
        { Sender<std::vector<float>> dropped(std::move(ch.first)); }
        st.ensure_input_samples_available(4);
        if (!st.is_done() || st.input_buf_len() != 4) return 1;
    }
    {   // ensure_input_samples_available_loading_multiple_chunks
        auto p = basic_stretcher(1024);
        p.second.send({1.0f, 2.0f, 3.0f});
        p.second.send({4.0f, 5.0f});
        p.first.ensure_input_samples_available(4);
        if (p.first.is_done() || p.first.input_buf_len() != 5) return 2;
    }
    {   // invalid pitch: assert!(pitch_multiple != 0), resampler panic on -1
        try { auto ch = unbounded<std::vector<float>>(); Stretcher st(AudioSpec{1, 44100}, std::move(ch.second), 1.0f, 1.0f, 0, 1024, std::chrono::duration<float>(1.0f)); return 3; }
        catch (const Error& e) { if (e.status != RCG_ERR_PITCH) return 4; }
        try { auto ch = unbounded<std::vector<float>>(); Stretcher st(AudioSpec{1, 44100}, std::move(ch.second), 1.0f, 1.0f, -1, 1024, std::chrono::duration<float>(1.0f)); return 5; }
        catch (const Error& e) { if (e.status != RCG_ERR_PITCH) return 6; }
    }
    std::puts("ensure tests ok");
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "ensure")) return test_ensure();
    if (argc < 10 || std::strcmp(argv[1], "run")) {
        std::fprintf(stderr, "usage: %s ensure | run in.f32 channels n factor pitch window seed out.f32\n", argv[0]);
        return 64;
    }
    const char* in_path = argv[2];
    const int channels = std::atoi(argv[3]);
    const size_t n = std::strtoull(argv[4], nullptr, 10);
    const float factor = (float)std::atof(argv[5]);
    const int pitch = std::atoi(argv[6]);
    const size_t window = std::strtoull(argv[7], nullptr, 10);
    const uint64_t seed = std::strtoull(argv[8], nullptr, 10);
    const char* out_path = argv[9];
    std::vector<std::vector<float>> audio(channels, std::vector<float>(n));
    std::ifstream in(in_path, std::ios::binary);
    for (auto& c : audio) in.read(reinterpret_cast<char*>(c.data()), (std::streamsize)(n * sizeof(float)));
    try {
        // main.rs:133-153: one Stretcher per channel, the whole channel sent as one chunk
        std::vector<Stretcher> stretchers;
        for (int c = 0; c < channels; c++) {
            auto ch = unbounded<std::vector<float>>();
            stretchers.emplace_back(AudioSpec{(uint16_t)channels, 48000}, std::move(ch.second), factor, 1.0f, (int8_t)pitch,
                                    window, std::chrono::duration<float>(1.0f), "", (uint32_t)c, seed);
            ch.first.send(audio[c]);
        }   // senders dropped here
        auto pb = StretcherProcessor::create(std::move(stretchers), (size_t)((float)n * factor));  // main.rs:154-155
        pb.first->start();
        auto out = pb.second.into_audio();                                                          // main.rs:200
        pb.first->join();
        if (!pb.first->error().empty()) { std::fprintf(stderr, "%s\n", pb.first->error().c_str()); return 2; }
        std::ofstream of(out_path, std::ios::binary);
        for (auto& c : out) of.write(reinterpret_cast<const char*>(c.data()), (std::streamsize)(c.size() * sizeof(float)));
        std::printf("%zu samples per channel\n", out.empty() ? 0 : out[0].size());
    } catch (const Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    return 0;
}

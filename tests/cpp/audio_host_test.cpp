// Re-states the reference's own unit tests for the host helpers the file-to-file path needs:
// duration_parser.rs:31-43 (8 cases), audio.rs clip_in_place / rotate_channels tests, the
// integer-sample conversions (audio.rs:16-29) and a WAV write/read round trip.
#include <cstdio>
#include <sstream>

#include "../../rocoder_b200/host/rocoder_audio.hpp"

using namespace rocoder;

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); fails++; } } while (0)

static bool dur(const char* s, uint64_t* ms) { try { *ms = parse_duration_ms(s); return true; } catch (const AudioError&) { return false; } }
static Audio generate_audio(float fill, size_t len, uint16_t channels, uint32_t rate) {   // test_utils.rs:70-82
    Audio a; a.spec.channels = channels; a.spec.sample_rate = rate;
    a.data.assign(channels, std::vector<float>(len, fill));
    return a;
}

int main(int argc, char** argv) {
    uint64_t ms = 0;
    // duration_parser.rs:31-43
    CHECK(!dur("adkjfn", &ms));
    CHECK(dur("1", &ms) && ms == 1000);
    CHECK(dur("1:1", &ms) && ms == 61000);
    CHECK(dur("1:1:1", &ms) && ms == 3661000);
    CHECK(dur("1:1:1.234", &ms) && ms == 3661234);
    CHECK(!dur("1:2:3:4", &ms));
    CHECK(!dur("1:2.9:4", &ms));
    CHECK(!dur("1.9:2:4", &ms));
    CHECK(dur("0.5", &ms) && ms == 500);
    // audio.rs tests: generate_audio(0.0, 5, 2, 2) = 5 samples, 2 channels, 2 Hz
    { Audio a = generate_audio(0, 5, 2, 2); a.clip_in_place(nullptr, nullptr); CHECK(a.data[0].size() == 5 && a.data[1].size() == 5); }
    { Audio a = generate_audio(0, 5, 2, 2); uint64_t s = 500; a.clip_in_place(&s, nullptr); CHECK(a.data[0].size() == 4 && a.data[1].size() == 4); }
    { Audio a = generate_audio(0, 5, 2, 2); uint64_t d = 500; a.clip_in_place(nullptr, &d); CHECK(a.data[0].size() == 1 && a.data[1].size() == 1); }
    { Audio a = generate_audio(0, 5, 2, 2); uint64_t s = 500, d = 1000; a.clip_in_place(&s, &d); CHECK(a.data[0].size() == 2 && a.data[1].size() == 2); }
    { Audio a = generate_audio(5, 2, 2, 2); a.data[1][0] = 6; a.rotate_channels(); CHECK(a.data[0][0] == 6 && a.data[1][0] == 5); }
    { Audio a = generate_audio(0, 5, 2, 2); uint64_t s = 5000; bool threw = false; try { a.clip_in_place(&s, nullptr); } catch (const AudioError&) { threw = true; } CHECK(threw); }
    // audio.rs:16-29
    CHECK(sample_from_i8(127) == 1.0f && sample_from_i16(32767) == 1.0f && sample_from_i16(-32768) < -1.0f);
    CHECK(sample_from_i24(8388607) < 1.0f && sample_from_i24(-8388608) == -1.0f && sample_from_i32(INT32_MAX) == 1.0f);
    // WAV: write 32-bit float, read it back bit for bit
    if (argc > 1) {
        std::vector<std::vector<float>> ch = {{0.0f, 0.25f, -0.5f, 1.0f}, {0.125f, -1.0f, 0.75f, 1e-8f}};
        write_wav_file(argv[1], ch, 48000);
        Audio a = read_wav_file(argv[1]);
        CHECK(a.spec.channels == 2 && a.spec.sample_rate == 48000 && a.data == ch);
    }
    std::printf(fails ? "%d failures\n" : "audio host tests ok\n", fails);
    return fails ? 1 : 0;
}

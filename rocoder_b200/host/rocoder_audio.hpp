// rocoder_audio.hpp -- host-side helpers around the stretch path, restating the reference's
// data formats so that file-to-file runs (`rocoder -i in.wav -o out.wav ...`) behave alike:
//   Audio / AudioSpec, clip_in_place, rotate_channels      reference src/audio.rs:32-75, 115-130
//   integer sample -> f32 conversions                      src/audio.rs:16-29
//   WavReader (8/16/24/32-bit int, 32-bit float)           src/audio_files.rs:115-189
//   WavWriter (32-bit float, interleaved)                  src/audio_files.rs:68-81, 199-227
//   parse_duration ("hh:mm:ss.ss")                         src/duration_parser.rs:5-25
// Plain C++17, no CUDA.  Errors are exceptions (the reference panics / returns Err).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <istream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace rocoder {

struct AudioError : std::runtime_error { using std::runtime_error::runtime_error; };

// src/duration_parser.rs:5-25: rsplit(':'), at most 3 fields; seconds parse as f32 and are
// truncated to milliseconds; minutes and hours must parse as u64.  Returns milliseconds.
inline uint64_t parse_duration_ms(const std::string& text) {
    std::vector<std::string> parts;
    size_t end = text.size();
    while (true) {
        size_t pos = text.rfind(':', end == 0 ? 0 : end - 1);
        if (end == 0 || pos == std::string::npos) { parts.push_back(text.substr(0, end)); break; }
        parts.push_back(text.substr(pos + 1, end - pos - 1));
        end = pos;
    }
    if (parts.size() > 3 || parts.empty()) throw AudioError("Invalid duration specification");
    auto parse_f32 = [](const std::string& s) {
        if (s.empty()) throw AudioError("invalid float literal");
        for (char c : s) if (!(std::isdigit((unsigned char)c) || c == '.' || c == '+' || c == '-' || c == 'e' || c == 'E'))
            throw AudioError("invalid float literal");
        char* e = nullptr;
        float v = std::strtof(s.c_str(), &e);
        if (e == s.c_str() || *e != '\0') throw AudioError("invalid float literal");
        return v;
    };
    auto parse_u64 = [](const std::string& s) {
        if (s.empty()) throw AudioError("invalid digit found in string");
        size_t i = s[0] == '+' ? 1 : 0;
        if (i == s.size()) throw AudioError("invalid digit found in string");
        uint64_t v = 0;
        for (; i < s.size(); i++) {
            if (!std::isdigit((unsigned char)s[i])) throw AudioError("invalid digit found in string");
            v = v * 10 + (uint64_t)(s[i] - '0');
        }
        return v;
    };
    const float seconds = parse_f32(parts[0]);
    const float ms_f = seconds * 1000.0f;
    uint64_t ms = (!(ms_f == ms_f) || ms_f <= 0.0f) ? 0 : (uint64_t)ms_f;   // `as u64`: saturating, NaN -> 0
    if (parts.size() > 1) ms += parse_u64(parts[1]) * 60 * 1000;
    if (parts.size() > 2) ms += parse_u64(parts[2]) * 60 * 60 * 1000;
    return ms;
}

// src/audio.rs:32-37, 39-43
struct AudioSpecH {
    uint16_t channels = 0;
    uint32_t sample_rate = 0;
};
struct Audio {
    std::vector<std::vector<float>> data;   // planar: one Vec<f32> per channel
    AudioSpecH spec;

    // src/audio.rs:57-63, 115-130: positions computed in f64 from the Duration
    void clip_in_place(const uint64_t* start_ms, const uint64_t* duration_ms) {
        const size_t len = data.empty() ? 0 : data[0].size();
        const size_t start = start_ms ? (size_t)(((double)*start_ms / 1000.0) * (double)spec.sample_rate) : 0;
        const size_t end = duration_ms ? start + (size_t)(((double)*duration_ms / 1000.0) * (double)spec.sample_rate) : len;
        if (start > end || end > len) throw AudioError("clip range out of bounds");   // slice index panic
        for (auto& c : data) c = std::vector<float>(c.begin() + (long)start, c.begin() + (long)end);
    }
    // src/audio.rs:73: data.rotate_right(1)
    void rotate_channels() {
        if (!data.empty()) std::rotate(data.rbegin(), data.rbegin() + 1, data.rend());
    }
};

// src/audio.rs:16-29
inline float sample_from_i8(int8_t n) { return (float)n / 127.0f; }
inline float sample_from_i16(int16_t n) { return (float)n / 32767.0f; }
inline float sample_from_i24(int32_t n) { return (float)n / 8388608.0f; }
inline float sample_from_i32(int32_t n) { return (float)n / (float)INT32_MAX; }

namespace detail {
inline uint32_t rd_u32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd_u16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
}  // namespace detail

// src/audio_files.rs:115-189 + read_all (:30-49): RIFF/WAVE, PCM (1), IEEE float (3) and
// WAVE_FORMAT_EXTENSIBLE (0xFFFE) with those sub-formats; samples are de-interleaved.
inline Audio read_wav(std::istream& in) {
    std::vector<unsigned char> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (buf.size() < 12 || std::memcmp(buf.data(), "RIFF", 4) || std::memcmp(buf.data() + 8, "WAVE", 4))
        throw AudioError("not a RIFF/WAVE stream");
    size_t pos = 12;
    uint16_t fmt_tag = 0, channels = 0, bits = 0;
    uint32_t rate = 0;
    const unsigned char* data = nullptr;
    size_t data_len = 0;
    while (pos + 8 <= buf.size()) {
        const uint32_t len = detail::rd_u32(buf.data() + pos + 4);
        const unsigned char* body = buf.data() + pos + 8;
        const size_t avail = std::min<size_t>(len, buf.size() - pos - 8);
        if (!std::memcmp(buf.data() + pos, "fmt ", 4) && avail >= 16) {
            fmt_tag = detail::rd_u16(body); channels = detail::rd_u16(body + 2); rate = detail::rd_u32(body + 4);
            bits = detail::rd_u16(body + 14);
            if (fmt_tag == 0xFFFE && avail >= 26) fmt_tag = detail::rd_u16(body + 24);  // sub-format GUID's first word
        } else if (!std::memcmp(buf.data() + pos, "data", 4)) {
            data = body; data_len = avail;
            break;
        }
        pos += 8 + (size_t)len + (len & 1);
    }
    if (!channels || !data) throw AudioError("missing fmt or data chunk");
    const size_t bytes = bits / 8;
    if (!((fmt_tag == 3 && bits == 32) || (fmt_tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32))))
        throw AudioError("Cannot read unsupported .wav format");   // audio_files.rs:186
    const size_t num_samples = data_len / bytes;
    if (num_samples % channels) throw AudioError("num_samples is not a multiple of channel count");   // :107-113
    Audio a;
    a.spec.channels = channels; a.spec.sample_rate = rate;
    a.data.assign(channels, std::vector<float>());
    for (auto& c : a.data) c.reserve(num_samples / channels);
    for (size_t i = 0; i < num_samples; i++) {
        const unsigned char* p = data + i * bytes;
        float v;
        if (fmt_tag == 3) { uint32_t u = detail::rd_u32(p); std::memcpy(&v, &u, 4); }
        else if (bits == 8) v = sample_from_i8((int8_t)((int)p[0] - 128));        // 8-bit WAV is unsigned on disk
        else if (bits == 16) v = sample_from_i16((int16_t)detail::rd_u16(p));
        else if (bits == 24) { int32_t s = (int32_t)(p[0] | (p[1] << 8) | (p[2] << 16)); if (s & 0x800000) s |= ~0xFFFFFF; v = sample_from_i24(s); }
        else v = sample_from_i32((int32_t)detail::rd_u32(p));
        a.data[i % channels].push_back(v);
    }
    return a;
}
inline Audio read_wav_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw AudioError("cannot open " + path);
    return read_wav(f);
}

// src/audio_files.rs:68-81, 203-219: 32-bit float, samples interleaved channel by channel
inline void write_wav_file(const std::string& path, const std::vector<std::vector<float>>& channels, uint32_t sample_rate) {
    if (channels.empty()) throw AudioError("no channels to write");
    const size_t n = channels[0].size();
    for (auto& c : channels) if (c.size() != n) throw AudioError("channels differ in length");
    const uint16_t ch = (uint16_t)channels.size();
    const uint64_t data_bytes = (uint64_t)n * ch * 4;
    if (data_bytes > 0xFFFFFFFFull - 44) throw AudioError("output exceeds the 4 GiB RIFF limit");
    std::ofstream f(path, std::ios::binary);
    if (!f) throw AudioError("cannot create " + path);
    auto w32 = [&](uint32_t v) { unsigned char b[4] = {(unsigned char)v, (unsigned char)(v >> 8), (unsigned char)(v >> 16), (unsigned char)(v >> 24)}; f.write((char*)b, 4); };
    auto w16 = [&](uint16_t v) { unsigned char b[2] = {(unsigned char)v, (unsigned char)(v >> 8)}; f.write((char*)b, 2); };
    f.write("RIFF", 4); w32((uint32_t)(36 + data_bytes)); f.write("WAVE", 4);
    f.write("fmt ", 4); w32(16); w16(3); w16(ch); w32(sample_rate); w32(sample_rate * ch * 4); w16((uint16_t)(ch * 4)); w16(32);
    f.write("data", 4); w32((uint32_t)data_bytes);
    std::vector<float> inter((size_t)std::min<uint64_t>(n, 1 << 16) * ch);
    for (size_t i0 = 0; i0 < n; i0 += (1 << 16)) {
        const size_t m = std::min<size_t>(1 << 16, n - i0);
        for (size_t i = 0; i < m; i++)
            for (uint16_t c = 0; c < ch; c++) inter[i * ch + c] = channels[c][i0 + i];
        f.write(reinterpret_cast<const char*>(inter.data()), (std::streamsize)(m * ch * 4));
    }
    if (!f) throw AudioError("write failed: " + path);
}

}  // namespace rocoder

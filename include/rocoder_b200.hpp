// rocoder_b200.hpp -- C++ host-side mirror of the reference's signal-flow API for the stretch
// path, on top of the C ABI (rocoder_b200.h).  Same names, argument meaning and error behaviour
// as reference src/stretcher.rs, src/stretcher_processor.rs and the parts of src/audio.rs they
// use; panics become rocoder::Error exceptions on the C++ side of the boundary.
//
//   rocoder::Channel<T> / Sender / Receiver   crossbeam_channel (main.rs:137, stretcher_processor.rs:34)
//   rocoder::AudioSpec                        audio.rs:32-37
//   rocoder::Stretcher                        stretcher.rs:12-136
//   rocoder::StretcherProcessor               stretcher_processor.rs:22-75
//   rocoder::AudioBus                         audio.rs:142-172
//
// Header-only; link with -lrocoder_b200.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "rocoder_b200.h"

namespace rocoder {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error("rocoder_b200: " + m), status(s) {}
};
inline void check(int status) {
    if (status != RCG_OK) throw Error(status, rcg_last_error());
}

// ---- crossbeam-like channel: unbounded (cap 0) or bounded; recv fails once empty and all
// senders are gone ---------------------------------------------------------------------------
template <class T>
class Channel {
  public:
    explicit Channel(size_t cap) : cap_(cap) {}
    void send(T v) {
        std::unique_lock<std::mutex> l(m_);
        cv_space_.wait(l, [&] { return cap_ == 0 || q_.size() < cap_ || receiver_gone_; });
        if (receiver_gone_) throw Error(RCG_ERR_STATE, "send on a disconnected channel");
        q_.push_back(std::move(v));
        cv_item_.notify_one();
    }
    // nullopt <=> RecvError (disconnected and empty)
    std::optional<T> recv() {
        std::unique_lock<std::mutex> l(m_);
        cv_item_.wait(l, [&] { return !q_.empty() || senders_ == 0; });
        if (q_.empty()) return std::nullopt;
        T v = std::move(q_.front());
        q_.pop_front();
        cv_space_.notify_one();
        return v;
    }
    // 0 = got an item, 1 = timed out, 2 = disconnected (recv_timeout, audio.rs:158-166)
    int recv_timeout(std::chrono::milliseconds d, T& out) {
        std::unique_lock<std::mutex> l(m_);
        if (!cv_item_.wait_for(l, d, [&] { return !q_.empty() || senders_ == 0; })) return 1;
        if (q_.empty()) return 2;
        out = std::move(q_.front());
        q_.pop_front();
        cv_space_.notify_one();
        return 0;
    }
    void add_sender() { std::lock_guard<std::mutex> l(m_); senders_++; }
    void drop_sender() { std::lock_guard<std::mutex> l(m_); if (senders_ > 0) senders_--; cv_item_.notify_all(); }
    void drop_receiver() { std::lock_guard<std::mutex> l(m_); receiver_gone_ = true; cv_space_.notify_all(); }

  private:
    std::mutex m_;
    std::condition_variable cv_item_, cv_space_;
    std::deque<T> q_;
    size_t cap_;
    int senders_ = 0;
    bool receiver_gone_ = false;
};

template <class T>
class Sender {
  public:
    explicit Sender(std::shared_ptr<Channel<T>> c) : c_(std::move(c)) { c_->add_sender(); }
    Sender(const Sender& o) : c_(o.c_) { c_->add_sender(); }
    Sender(Sender&& o) noexcept : c_(std::move(o.c_)) {}
    Sender& operator=(const Sender&) = delete;
    ~Sender() { if (c_) c_->drop_sender(); }
    void send(T v) { c_->send(std::move(v)); }

  private:
    std::shared_ptr<Channel<T>> c_;
};
template <class T>
class Receiver {
  public:
    explicit Receiver(std::shared_ptr<Channel<T>> c) : c_(std::move(c)) {}
    Receiver(Receiver&&) noexcept = default;
    Receiver(const Receiver&) = delete;
    ~Receiver() { if (c_) c_->drop_receiver(); }
    std::optional<T> recv() { return c_->recv(); }
    int recv_timeout(std::chrono::milliseconds d, T& out) { return c_->recv_timeout(d, out); }

  private:
    std::shared_ptr<Channel<T>> c_;
};
template <class T>
std::pair<Sender<T>, Receiver<T>> unbounded() {
    auto c = std::make_shared<Channel<T>>(0);
    return {Sender<T>(c), Receiver<T>(c)};
}
template <class T>
std::pair<Sender<T>, Receiver<T>> bounded(size_t cap) {
    auto c = std::make_shared<Channel<T>>(cap ? cap : 1);
    return {Sender<T>(c), Receiver<T>(c)};
}

// audio.rs:32-37
struct AudioSpec {
    uint16_t channels;
    uint32_t sample_rate;
};

// ---- Stretcher (stretcher.rs) -----------------------------------------------------------------
// concurrent vocoder for one channel of audio
class Stretcher {
  public:
    AudioSpec spec;

    // Stretcher::new (stretcher.rs:30-39).  `window_len` stands for `window: Vec<f32>` (the engine
    // builds windows::hanning(window_len), what main.rs:131 passes); frequency_kernel_src is the
    // path of a kernel file (.cu / .c / .cpp source or a .so exporting the C-ABI `apply`) that is built,
    // watched and hot-swapped like ReFFT::new does (fft.rs:31-35, hotswapper.rs:14-33), or empty.  channel / seed / device are
    // the additions: phases are drawn from a counter-based generator keyed by them.
    Stretcher(AudioSpec spec_, Receiver<std::vector<float>> input, float factor, float amplitude,
              int8_t pitch_multiple, size_t window_len, std::chrono::duration<float> buffer_dur,
              const std::string& frequency_kernel_src = "", uint32_t channel = 0, uint64_t seed = 0,
              int device = 0, uint32_t flags = 0)
        : spec(spec_), input_(std::move(input)), window_len_(window_len) {
        if (pitch_multiple == 0) throw Error(RCG_ERR_PITCH, "assertion failed: pitch_multiple != 0");  // stretcher.rs:40
        cfg_ = rcg_config{};
        cfg_.struct_size = sizeof(rcg_config);
        cfg_.channels = 1;
        cfg_.sample_rate = spec.sample_rate;
        cfg_.window_len = (uint32_t)window_len;
        cfg_.factor = factor;
        cfg_.amplitude = amplitude;
        cfg_.pitch_multiple = pitch_multiple;
        cfg_.buffer_dur_s = buffer_dur.count();
        cfg_.seed = seed;
        cfg_.device = device;
        cfg_.flags = flags;
        cfg_.first_channel = channel;
        check(rcg_create(&cfg_, &engine_));
        check(rcg_get_plan(engine_, &plan_));
        if (!frequency_kernel_src.empty()) check(rcg_watch_frequency_kernel(engine_, frequency_kernel_src.c_str(), 0));
    }
    Stretcher(Stretcher&& o) noexcept
        : spec(o.spec), input_(std::move(o.input_)), engine_(o.engine_), cfg_(o.cfg_), plan_(o.plan_),
          window_len_(o.window_len_), done_(o.done_), sent_(o.sent_), consumed_(o.consumed_), closed_(o.closed_) {
        o.engine_ = nullptr;
    }
    Stretcher(const Stretcher&) = delete;
    ~Stretcher() { rcg_destroy(engine_); }

    // stretcher.rs:78
    bool is_done() const { return done_; }

    // stretcher.rs:82-85
    size_t channel_bound() const {
        uint64_t b = 0;
        check(rcg_channel_bound(&cfg_, &b));
        return (size_t)b;
    }

    size_t input_buf_len() const { return (size_t)(sent_ - consumed_); }

    // stretcher.rs:123-135
    void ensure_input_samples_available(size_t n) {
        while (input_buf_len() < n) {
            auto chunk = input_.recv();
            if (chunk) {                                            // Ok(chunk) => input_buf.extend(chunk)
                check(rcg_send(engine_, 0, chunk->data(), chunk->size()));
                sent_ += chunk->size();
            } else {                                                // Err(_) => resize(n, 0.0); done = true
                if (!closed_) { check(rcg_close_input(engine_, 0)); closed_ = true; }
                sent_ = consumed_ + n;
                done_ = true;
            }
        }
    }

    // stretcher.rs:87-121
    std::vector<float> next_window() {
        for (uint64_t i = 0; i < plan_.frames_per_call; i++) {
            ensure_input_samples_available(window_len_);
            consumed_ += plan_.sample_step_len;                     // truncate_front, stretcher.rs:105-106
        }
        std::vector<float> out(plan_.out_len_per_call);
        uint64_t n = 0;
        check(rcg_next_window(engine_, 0, out.data(), out.size(), &n));
        out.resize(n);
        return out;
    }

    const rcg_plan_info& plan() const { return plan_; }

  private:
    Receiver<std::vector<float>> input_;
    rcg_engine* engine_ = nullptr;
    rcg_config cfg_{};
    rcg_plan_info plan_{};
    size_t window_len_;
    bool done_ = false;
    uint64_t sent_ = 0, consumed_ = 0;
    bool closed_ = false;
};

// ---- AudioBus (audio.rs:142-172) ----------------------------------------------------------------
struct AudioBus {
    AudioSpec spec;
    std::vector<Receiver<std::vector<float>>> channels;
    std::optional<size_t> expected_total_samples;

    // audio.rs:152-172: round-robin recv_timeout(5 ms) per channel until all are disconnected
    std::vector<std::vector<float>> into_audio() {
        std::vector<std::vector<float>> data(channels.size());
        std::vector<bool> live(channels.size(), true);
        size_t n_live = channels.size();
        while (n_live) {
            for (size_t i = 0; i < channels.size(); i++) {
                if (!live[i]) continue;
                std::vector<float> w;
                int rc = channels[i].recv_timeout(std::chrono::milliseconds(5), w);
                if (rc == 0) data[i].insert(data[i].end(), w.begin(), w.end());
                else if (rc == 2) { live[i] = false; n_live--; }
            }
        }
        return data;
    }
};

// ---- StretcherProcessor (stretcher_processor.rs:22-75) -----------------------------------------
class StretcherProcessor {
  public:
    // stretcher_processor.rs:26-46
    static std::pair<std::unique_ptr<StretcherProcessor>, AudioBus> create(std::vector<Stretcher> channel_stretchers,
                                                                          std::optional<size_t> expected_total_samples) {
        auto p = std::unique_ptr<StretcherProcessor>(new StretcherProcessor());
        AudioBus bus{channel_stretchers.at(0).spec, {}, expected_total_samples};
        for (auto& st : channel_stretchers) {
            auto ch = bounded<std::vector<float>>(st.channel_bound());
            p->channels_.emplace_back(std::move(ch.first), std::move(st));
            bus.channels.push_back(std::move(ch.second));
        }
        return {std::move(p), std::move(bus)};
    }

    // stretcher_processor.rs:50-75: one thread drives every channel; stops when the first
    // stretcher reports done ("assuming each stretcher finishes at the same time") or on Shutdown
    void start() {
        thread_ = std::thread([this] {
            try {
                for (bool stop = false; !stop && !shutdown_.load();) {
                    for (auto& [output, stretcher] : channels_) {
                        if (stretcher.is_done()) { stop = true; break; }
                        output.send(stretcher.next_window());
                    }
                }
            } catch (const std::exception& e) {
                error_ = e.what();
            }
            channels_.clear();  // drops the senders: receivers see the disconnect
            finished_.store(true);
        });
    }
    void shutdown() { shutdown_.store(true); }   // StretcherProcessorControlMessage::Shutdown
    void join() { if (thread_.joinable()) thread_.join(); }
    bool finished() const { return finished_.load(); }
    const std::string& error() const { return error_; }
    ~StretcherProcessor() { shutdown(); join(); }

  private:
    StretcherProcessor() = default;
    std::vector<std::pair<Sender<std::vector<float>>, Stretcher>> channels_;
    std::thread thread_;
    std::atomic<bool> shutdown_{false}, finished_{false};
    std::string error_;
};

}  // namespace rocoder

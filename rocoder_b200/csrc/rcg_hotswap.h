// rcg_hotswap.h -- compile-and-swap of frequency kernels from a watched source file.
//
// Mirrors reference src/hotswapper.rs:14-86: `hotswap(path)` builds the file once, then a background
// thread polls it every 100 ms (WATCHER_POLL_DUR, :12) and builds it again whenever it is modified;
// every successful build is sent down an unbounded channel, a failed one only logs (:36-42).  The DSP
// side takes what has arrived without blocking (fft.rs:78-81).  Here the channel is a mutex-guarded
// queue drained by the engine at the start of each batch of frames.
//
// Host-only C++ (no CUDA types): device sources are handed to a compile callback that returns a cubin.
// What a build produces, by suffix of the watched path:
//   .cu            -> cubin via the callback (NVRTC in the engine), a device `apply`; `*.local.cu` is the bin-local
//                     form that the engine fuses into the stretch kernel itself
//   .c .cc .cpp    -> `$RCG_CXX (default g++) -O3 -shared -fPIC` into a temp .so, dlopen, host `apply`
//   .so            -> dlopen of a private copy of the file, host `apply`
//   .rs            -> a kernel written for the reference itself (`#[no_mangle] pub fn apply(usize, Vec<(f32, f32)>)
//                     -> Vec<(f32, f32)>`, README.md:108-113): the source plus a C-ABI shim is built with
//                     `$RCG_RUSTC (default rustc) -C opt-level=3 --crate-type cdylib` (hotswapper.rs:60-72 uses
//                     `--crate-type dylib`), when a Rust compiler is installed; a panic maps to a non-zero return
#pragma once
#include <dlfcn.h>
#include <fcntl.h>
#include <spawn.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

extern char** environ;

namespace rcg {

typedef int (*host_apply_fn)(size_t, const float*, size_t, float*);

struct BuiltKernel {
    void* dl = nullptr;           // dlopen handle (host kernels); owned by whoever pops it
    host_apply_fn host = nullptr;
    std::vector<char> cubin;      // device kernels
    std::vector<char> twin;       // bin-local form only: the same `apply` as a plain per-bin kernel over spectra in memory
    bool local = false;           // device kernel of the bin-local form, fused into the stretch kernel
};

// source text -> cubin; returns false and fills `log` on a compile error
// local: the bin-local form `float2 apply(size_t, unsigned bin, unsigned n_bins, float2 x)` (file *.local.cu)
typedef std::function<bool(const std::string& src, bool local, std::vector<char>& cubin, std::vector<char>& twin, std::string& log)> DeviceCompileFn;

class KernelWatcher {
  public:
    KernelWatcher(std::string path, unsigned poll_ms, DeviceCompileFn dev)
        : path_(std::move(path)), poll_ms_(poll_ms ? poll_ms : 100), dev_(std::move(dev)) {}
    ~KernelWatcher() { stop(); }
    KernelWatcher(const KernelWatcher&) = delete;
    KernelWatcher& operator=(const KernelWatcher&) = delete;

    enum Kind { kUnknown, kDeviceSource, kHostSource, kSharedObject, kRustSource };
    static Kind kind_of(const std::string& p) {
        auto ends = [&](const char* s) {
            const size_t n = strlen(s);
            return p.size() >= n && p.compare(p.size() - n, n, s) == 0;
        };
        if (ends(".cu")) return kDeviceSource;
        if (ends(".c") || ends(".cc") || ends(".cpp") || ends(".cxx")) return kHostSource;
        if (ends(".so")) return kSharedObject;
        if (ends(".rs")) return kRustSource;
        return kUnknown;
    }

    static bool is_local_device_source(const std::string& p) {
        const char* suf = ".local.cu";
        const size_t n = strlen(suf);
        return p.size() >= n && p.compare(p.size() - n, n, suf) == 0;
    }

    // hotswapper.rs:17 + :20-30: first attempt on the calling thread, then the polling thread
    void start() {
        remember_stamp();
        attempt_update();
        thread_ = std::thread([this] { run(); });
    }
    void stop() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        if (thread_.joinable()) thread_.join();
        // builds nobody took
        for (auto& b : queue_) if (b.dl) dlclose(b.dl);
        queue_.clear();
    }
    // Receiver::try_recv (fft.rs:78): false when nothing has arrived
    bool try_recv(BuiltKernel& out) {
        std::lock_guard<std::mutex> g(m_);
        if (queue_.empty()) return false;
        out = std::move(queue_.front());
        queue_.pop_front();
        return true;
    }
    unsigned builds_ok() const { return ok_.load(); }
    unsigned builds_failed() const { return failed_.load(); }
    std::string last_log() const {
        std::lock_guard<std::mutex> g(m_);
        return log_;
    }
    const std::string& path() const { return path_; }

  private:
    struct Stamp {
        bool exists = false;
        long long sec = 0, nsec = 0, size = 0;
        unsigned long long ino = 0;
        bool operator==(const Stamp& o) const {
            return exists == o.exists && sec == o.sec && nsec == o.nsec && size == o.size && ino == o.ino;
        }
    };
    Stamp stamp_now() const {
        struct stat st;
        Stamp s;
        if (stat(path_.c_str(), &st) != 0) return s;
        s.exists = true;
        s.sec = (long long)st.st_mtim.tv_sec;
        s.nsec = (long long)st.st_mtim.tv_nsec;
        s.size = (long long)st.st_size;
        s.ino = (unsigned long long)st.st_ino;
        return s;
    }
    void remember_stamp() { stamp_ = stamp_now(); }

    void run() {
        std::unique_lock<std::mutex> lk(m_);
        while (!stop_) {
            cv_.wait_for(lk, std::chrono::milliseconds(poll_ms_));   // thread::sleep(WATCHER_POLL_DUR)
            if (stop_) break;
            lk.unlock();
            const Stamp now = stamp_now();
            // Transition::Modified only: a file that vanished or is mid-rename is ignored (:24-27)
            if (now.exists && !(now == stamp_)) {
                stamp_ = now;
                attempt_update();
            }
            lk.lock();
        }
    }

    // hotswapper.rs:35-50
    void attempt_update() {
        BuiltKernel b;
        std::string log;
        if (!compile(b, log)) {
            failed_++;
            std::lock_guard<std::mutex> g(m_);
            log_ = log;
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            queue_.push_back(std::move(b));
            log_.clear();
        }
        ok_++;
    }

    static bool read_file(const std::string& p, std::string& out) {
        std::ifstream f(p, std::ios::binary);
        if (!f) return false;
        std::ostringstream ss;
        ss << f.rdbuf();
        out = ss.str();
        return true;
    }
    static std::string temp_path(const char* suffix) {
        const char* dir = getenv("TMPDIR");
        std::string t = std::string(dir && *dir ? dir : "/tmp") + "/rcg_kernel_XXXXXX" + suffix;
        std::vector<char> buf(t.begin(), t.end());
        buf.push_back('\0');
        const int fd = mkstemps(buf.data(), (int)strlen(suffix));
        if (fd < 0) return std::string();
        close(fd);
        return std::string(buf.data());
    }
    // run argv, capture stderr+stdout into `log`; true on exit status 0
    static bool run_tool(const std::vector<std::string>& argv, std::string& log) {
        const std::string logfile = temp_path(".log");
        if (logfile.empty()) { log = "mkstemps failed"; return false; }
        posix_spawn_file_actions_t fa;
        posix_spawn_file_actions_init(&fa);
        posix_spawn_file_actions_addopen(&fa, 1, logfile.c_str(), O_WRONLY | O_TRUNC, 0600);
        posix_spawn_file_actions_adddup2(&fa, 1, 2);
        std::vector<char*> av;
        for (auto& a : argv) av.push_back(const_cast<char*>(a.c_str()));
        av.push_back(nullptr);
        pid_t pid = 0;
        const int rc = posix_spawnp(&pid, av[0], &fa, nullptr, av.data(), environ);
        posix_spawn_file_actions_destroy(&fa);
        bool ok = false;
        if (rc == 0) {
            int status = 0;
            while (waitpid(pid, &status, 0) < 0 && errno == EINTR) {}
            ok = WIFEXITED(status) && WEXITSTATUS(status) == 0;
        } else {
            log = std::string("cannot run ") + argv[0] + ": " + strerror(rc);
        }
        std::string text;
        if (read_file(logfile, text) && !text.empty()) log = text;
        unlink(logfile.c_str());
        return ok;
    }
    bool open_library(const std::string& so, BuiltKernel& b, std::string& log, const char* symbol = "apply") {
        void* lib = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!lib) { log = std::string("dlopen: ") + dlerror(); return false; }
        void* sym = dlsym(lib, symbol);   // fft.rs:93-94
        if (!sym) { log = so + " has no `" + symbol + "` symbol"; dlclose(lib); return false; }
        b.dl = lib;
        b.host = (host_apply_fn)sym;
        return true;
    }

  public:
    // C-ABI entry appended to a reference-style Rust kernel: same contract as rcg_apply_fn, built over the user's
    // `apply` (fft.rs:93: fn(usize, Vec<(f32, f32)>) -> Vec<(f32, f32)>); catch_unwind is what fft.rs:88 does
    static const char* rust_shim() {
        return "// ---- appended by rocoder_b200: C-ABI entry over `apply` ----\n"
               "#[no_mangle]\n"
               "pub unsafe extern \"C\" fn rcg_apply(elapsed_ms: usize, input: *const f32, n_bins: usize, output: *mut f32) -> i32 {\n"
               "    let inp: Vec<(f32, f32)> = (0..n_bins).map(|i| (*input.add(2 * i), *input.add(2 * i + 1))).collect();\n"
               "    match std::panic::catch_unwind(move || apply(elapsed_ms, inp)) {\n"
               "        Ok(v) => {\n"
               "            if v.len() != n_bins { return 2; }\n"
               "            for (i, c) in v.iter().enumerate() { *output.add(2 * i) = c.0; *output.add(2 * i + 1) = c.1; }\n"
               "            0\n"
               "        }\n"
               "        Err(_) => 1,\n"
               "    }\n"
               "}\n";
    }

  private:
    // hotswapper.rs:52-86
    bool compile(BuiltKernel& b, std::string& log) {
        switch (kind_of(path_)) {
            case kDeviceSource: {
                std::string src;
                if (!read_file(path_, src)) { log = "cannot read " + path_; return false; }
                if (!dev_) { log = "no device compiler"; return false; }
                b.local = is_local_device_source(path_);
                return dev_(src, b.local, b.cubin, b.twin, log);
            }
            case kHostSource: {
                const std::string so = temp_path(".so");   // tempfile::Builder::new().suffix(".so") (:58)
                if (so.empty()) { log = "mkstemps failed"; return false; }
                const char* cxx = getenv("RCG_CXX");
                const bool is_c = path_.size() > 2 && path_.compare(path_.size() - 2, 2, ".c") == 0;
                std::vector<std::string> argv = {cxx && *cxx ? cxx : (is_c ? "gcc" : "g++"), "-O3", "-shared", "-fPIC", "-w",
                                                 path_, "-o", so};
                bool ok = run_tool(argv, log) && open_library(so, b, log);
                unlink(so.c_str());   // the mapping stays valid; nothing is left behind in TMPDIR
                return ok;
            }
            case kSharedObject: {
                // dlopen caches by path: load a private copy so that a rebuilt file really is a new library
                const std::string so = temp_path(".so");
                std::string bytes;
                if (so.empty() || !read_file(path_, bytes)) { log = "cannot read " + path_; return false; }
                {
                    std::ofstream o(so, std::ios::binary | std::ios::trunc);
                    o.write(bytes.data(), (std::streamsize)bytes.size());
                }
                bool ok = open_library(so, b, log);
                unlink(so.c_str());
                return ok;
            }
            case kRustSource: {
                std::string src;
                if (!read_file(path_, src)) { log = "cannot read " + path_; return false; }
                const std::string rs = temp_path(".rs"), so = temp_path(".so");
                if (rs.empty() || so.empty()) { log = "mkstemps failed"; return false; }
                {
                    std::ofstream o(rs, std::ios::binary | std::ios::trunc);
                    o << src << "\n" << rust_shim();
                }
                const char* rustc = getenv("RCG_RUSTC");
                std::vector<std::string> argv = {rustc && *rustc ? rustc : "rustc", "-A", "warnings", "-C", "opt-level=3",
                                                 "--crate-type", "cdylib", "--crate-name", "rcg_kernel", rs, "-o", so};
                bool ok = run_tool(argv, log) && open_library(so, b, log, "rcg_apply");
                unlink(rs.c_str());
                unlink(so.c_str());
                return ok;
            }
            default:
                log = "unknown frequency-kernel file type (want .cu, .c, .cc, .cpp, .so or .rs): " + path_;
                return false;
        }
    }

    const std::string path_;
    const unsigned poll_ms_;
    DeviceCompileFn dev_;
    std::thread thread_;
    mutable std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
    std::deque<BuiltKernel> queue_;
    std::string log_;
    Stamp stamp_;
    std::atomic<unsigned> ok_{0}, failed_{0};
};

}  // namespace rcg

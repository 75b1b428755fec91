// CPU test of rocoder_b200/csrc/rcg_hotswap.h (the hot-swapper of reference src/hotswapper.rs):
// first build on the calling thread, rebuild on modification, a failed build changes nothing,
// non-blocking receive.  Device sources go through a fake compile callback.
#include "../../rocoder_b200/csrc/rcg_hotswap.h"

#include <cmath>
#include <cstdio>

static int g_fail = 0;
#define CHECK(c)                                                          \
    do {                                                                  \
        if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); g_fail++; } \
    } while (0)

static void write_file(const std::string& p, const std::string& text) {
    const std::string tmp = p + ".tmp";
    { std::ofstream o(tmp, std::ios::trunc); o << text; }
    rename(tmp.c_str(), p.c_str());   // what editors do: a new inode
}
static std::string kernel_src(float gain) {
    char b[512];
    snprintf(b, sizeof b,
             "#include <stddef.h>\nint apply(size_t ms, const float* in, size_t n, float* out) {\n"
             "  for (size_t i = 0; i < 2 * n; i++) out[i] = in[i] * %.1ff;\n  return 0;\n}\n", gain);
    return b;
}
static bool wait_for(const std::function<bool()>& f, int ms = 20000) {
    for (int t = 0; t < ms; t += 10) {
        if (f()) return true;
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    return f();
}
static float gain_of(const rcg::BuiltKernel& b) {
    float in[4] = {1, 2, 3, 4}, out[4] = {0, 0, 0, 0};
    if (!b.host || b.host(0, in, 2, out) != 0) return NAN;
    return out[3] / in[3];
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    {   // host source: build, rebuild on change, failed build, overwrite in place
        const std::string src = dir + "/k.c";
        write_file(src, kernel_src(2));
        rcg::KernelWatcher w(src, 20, nullptr);
        w.start();
        CHECK(w.builds_ok() == 1 && w.builds_failed() == 0);     // synchronous first attempt
        rcg::BuiltKernel a, b, c, d;
        CHECK(w.try_recv(a));
        CHECK(gain_of(a) == 2.0f);
        CHECK(!w.try_recv(b));                                   // try_recv does not block
        write_file(src, kernel_src(3));
        CHECK(wait_for([&] { return w.builds_ok() == 2; }));
        CHECK(w.try_recv(b));
        CHECK(gain_of(b) == 3.0f);
        CHECK(gain_of(a) == 2.0f);                               // earlier libraries stay usable
        write_file(src, "this is not C\n");
        CHECK(wait_for([&] { return w.builds_failed() == 1; }));
        CHECK(w.builds_ok() == 2);
        CHECK(!w.try_recv(c));
        CHECK(!w.last_log().empty());
        { std::ofstream o(src, std::ios::trunc); o << kernel_src(5); }   // in-place write, same inode
        CHECK(wait_for([&] { return w.builds_ok() == 3; }));
        CHECK(w.try_recv(d));
        CHECK(gain_of(d) == 5.0f);
        CHECK(w.last_log().empty());
        unlink(src.c_str());                                     // a vanished file is not a modification
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
        CHECK(w.builds_ok() == 3 && w.builds_failed() == 1);
        w.stop();
        for (auto* k : {&a, &b, &d}) dlclose(k->dl);
    }
    {   // a file that does not build at the start: no kernel, then picked up once fixed (hotswapper.rs:17, 36-42)
        const std::string src = dir + "/late.cpp";
        write_file(src, "int apply(;\n");
        rcg::KernelWatcher w(src, 20, nullptr);
        w.start();
        rcg::BuiltKernel a;
        CHECK(w.builds_ok() == 0 && w.builds_failed() == 1 && !w.try_recv(a));
        write_file(src, "#include <stddef.h>\nextern \"C\" int apply(size_t, const float* in, size_t n, float* out) {"
                        " for (size_t i = 0; i < 2 * n; i++) out[i] = in[i] * 7.0f; return 0; }\n");
        CHECK(wait_for([&] { return w.builds_ok() == 1; }));
        CHECK(w.try_recv(a) && gain_of(a) == 7.0f);
        // builds left in the queue are released by stop()
        write_file(src, "#include <stddef.h>\nextern \"C\" int apply(size_t, const float* in, size_t n, float* out) {"
                        " for (size_t i = 0; i < 2 * n; i++) out[i] = in[i] * 9.0f; return 0; }\n");
        CHECK(wait_for([&] { return w.builds_ok() == 2; }));
    }
    {   // a source without `apply`, and an unknown suffix
        const std::string src = dir + "/noapply.c";
        write_file(src, "int other(void) { return 0; }\n");
        rcg::KernelWatcher w(src, 20, nullptr);
        w.start();
        CHECK(w.builds_failed() == 1 && w.last_log().find("apply") != std::string::npos);
        CHECK(rcg::KernelWatcher::kind_of("k.txt") == rcg::KernelWatcher::kUnknown);
        CHECK(rcg::KernelWatcher::kind_of("a/b.cu") == rcg::KernelWatcher::kDeviceSource);
        CHECK(rcg::KernelWatcher::kind_of("x.so") == rcg::KernelWatcher::kSharedObject);
        CHECK(rcg::KernelWatcher::kind_of("gain.local.cu") == rcg::KernelWatcher::kDeviceSource);
        CHECK(rcg::KernelWatcher::is_local_device_source("a/gain.local.cu") && !rcg::KernelWatcher::is_local_device_source("gain.cu"));
    }
    {   // device source: text reaches the compile callback, cubin comes back
        const std::string src = dir + "/k.cu";
        write_file(src, "AAA");
        int calls = 0;
        rcg::KernelWatcher w(src, 20, [&](const std::string& s, bool local, std::vector<char>& cubin, std::vector<char>&, std::string& log) {
            CHECK(!local);
            calls++;
            if (s == "bad") { log = "syntax error"; return false; }
            cubin.assign(s.begin(), s.end());
            return true;
        });
        w.start();
        rcg::BuiltKernel a;
        CHECK(w.try_recv(a) && a.host == nullptr && std::string(a.cubin.begin(), a.cubin.end()) == "AAA");
        write_file(src, "bad");
        CHECK(wait_for([&] { return w.builds_failed() == 1; }));
        CHECK(w.last_log() == "syntax error");
        write_file(src, "BBBB");
        CHECK(wait_for([&] { return w.builds_ok() == 2; }));
        CHECK(w.try_recv(a) && a.cubin.size() == 4);
        CHECK(calls == 3);
    }
    {   // .so: a rebuilt file at the same path is a new library
        const std::string c1 = dir + "/s.c", so = dir + "/s.so";
        write_file(c1, kernel_src(4));
        CHECK(system(("gcc -shared -fPIC -o " + so + " " + c1).c_str()) == 0);
        rcg::KernelWatcher w(so, 20, nullptr);
        w.start();
        rcg::BuiltKernel a, b;
        CHECK(w.try_recv(a) && gain_of(a) == 4.0f);
        write_file(c1, kernel_src(6));
        CHECK(system(("gcc -shared -fPIC -o " + so + ".new " + c1 + " && mv " + so + ".new " + so).c_str()) == 0);
        CHECK(wait_for([&] { return w.builds_ok() >= 2; }));
        CHECK(w.try_recv(b) && gain_of(b) == 6.0f && gain_of(a) == 4.0f);
    }
    {   // .rs: the user's source plus the C-ABI shim goes to $RCG_RUSTC; `rcg_apply` of the result is the kernel.
        // No Rust compiler exists here: a stand-in script checks what it is handed and builds a C body instead.
        const std::string rs = dir + "/k.rs", fake = dir + "/fake_rustc.sh", seen = dir + "/seen.rs";
        write_file(rs, "#[no_mangle]\npub fn apply(_t: usize, input: Vec<(f32, f32)>) -> Vec<(f32, f32)> {\n"
                       "    input.iter().map(|c| (c.0 * 2.0, c.1 * 2.0)).collect()\n}\n");
        write_file(fake, "#!/bin/sh\n# args: -A warnings -C opt-level=3 --crate-type cdylib --crate-name rcg_kernel <src> -o <out>\n"
                         "[ \"$6\" = cdylib ] || exit 3\nsrc=\"$9\"; out=\"${11}\"\ncp \"$src\" " + seen + "\n"
                         "grep -q 'pub fn apply' \"$src\" || exit 4\ngrep -q 'extern \"C\" fn rcg_apply' \"$src\" || exit 5\n"
                         "printf '#include <stddef.h>\\nint rcg_apply(size_t t, const float* i, size_t n, float* o) { for (size_t k = 0; k < 2 * n; k++) o[k] = i[k] * 2.0f; return 0; }\\n' > \"$src.c\"\n"
                         "gcc -shared -fPIC -o \"$out\" \"$src.c\"; rc=$?; rm -f \"$src.c\"; exit $rc\n");
        chmod(fake.c_str(), 0755);
        setenv("RCG_RUSTC", fake.c_str(), 1);
        {
            rcg::KernelWatcher w(rs, 20, nullptr);
            w.start();
            rcg::BuiltKernel a;
            CHECK(w.builds_ok() == 1);
            if (w.builds_ok() != 1) printf("log: %s\n", w.last_log().c_str());
            CHECK(w.try_recv(a) && gain_of(a) == 2.0f);
            std::ifstream f(seen);
            std::stringstream ss; ss << f.rdbuf();
            CHECK(ss.str().find("catch_unwind") != std::string::npos && ss.str().find("input.iter().map") != std::string::npos);
        }
        setenv("RCG_RUSTC", "/nonexistent/rustc", 1);   // no compiler: a failed build, nothing else
        {
            rcg::KernelWatcher w(rs, 20, nullptr);
            w.start();
            rcg::BuiltKernel a;
            CHECK(w.builds_failed() == 1 && !w.try_recv(a) && !w.last_log().empty());
        }
        unsetenv("RCG_RUSTC");
        CHECK(rcg::KernelWatcher::kind_of("k.rs") == rcg::KernelWatcher::kRustSource);
    }
    if (g_fail) return 1;
    printf("hotswap_test: ok\n");
    return 0;
}

# Builds everything in-tree (same as `python -m rocoder_b200.build`).
#   make            product library (nvcc, sm_100a only) + CLI
#   make test-infra oracle + host emulator (test infrastructure, never linked into the product)
NVCC ?= nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC
CSRC := $(wildcard rocoder_b200/csrc/*.cu)
HDRS := $(wildcard rocoder_b200/csrc/*.h) include/rocoder_b200.h

all: rocoder_b200/librocoder_b200.so rocoder_b200/rocoder_gpu

# text of the kernel headers, embedded for the run-time (NVRTC) build of fused frequency kernels
rocoder_b200/csrc/rcg_jit_sources.inc: $(HDRS) rocoder_b200/build.py
	python -c "from rocoder_b200 import build; build.generate_jit_sources()"

rocoder_b200/librocoder_b200.so: $(CSRC) $(HDRS) rocoder_b200/csrc/rcg_jit_sources.inc
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC) -ldl

rocoder_b200/rocoder_gpu: tools/rocoder_gpu.cpp rocoder_b200/host/rocoder_audio.hpp rocoder_b200/librocoder_b200.so
	g++ -O2 -std=c++17 $< -o $@ -Lrocoder_b200 -lrocoder_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

test-infra:
	$(MAKE) -C oracle
	g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -o tests/emul/librcg_emul.so tests/emul/rcg_emul.cpp

clean:
	rm -f rocoder_b200/librocoder_b200.so rocoder_b200/rocoder_gpu tests/emul/librcg_emul.so rocoder_b200/csrc/rcg_jit_sources.inc
	$(MAKE) -C oracle clean
.PHONY: all test-infra clean

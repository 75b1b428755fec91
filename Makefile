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

# rcg_kernels.cu is one translation unit per window size (-DRCG_TU_LOG2H=L) plus the dispatch unit: `make -j`
OBJDIR := rocoder_b200/csrc/build
SIZES := 5 6 7 8 9 10 11 12 13 14
OBJS := $(patsubst rocoder_b200/csrc/%.cu,$(OBJDIR)/%.o,$(CSRC)) $(foreach l,$(SIZES),$(OBJDIR)/rcg_kernels_l$(l).o)

$(OBJDIR)/%.o: rocoder_b200/csrc/%.cu $(HDRS) rocoder_b200/csrc/rcg_jit_sources.inc
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $<
$(OBJDIR)/rcg_kernels_l%.o: rocoder_b200/csrc/rcg_kernels.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -DRCG_TU_LOG2H=$* -c -o $@ $<

rocoder_b200/librocoder_b200.so: $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -ldl

rocoder_b200/rocoder_gpu: tools/rocoder_gpu.cpp rocoder_b200/host/rocoder_audio.hpp rocoder_b200/librocoder_b200.so
	g++ -O2 -std=c++17 $< -o $@ -Lrocoder_b200 -lrocoder_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

test-infra:
	$(MAKE) -C oracle
	g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -o tests/emul/librcg_emul.so tests/emul/rcg_emul.cpp

clean:
	rm -rf rocoder_b200/csrc/build; rm -f rocoder_b200/librocoder_b200.so rocoder_b200/rocoder_gpu tests/emul/librcg_emul.so rocoder_b200/csrc/rcg_jit_sources.inc
	$(MAKE) -C oracle clean
.PHONY: all test-infra clean

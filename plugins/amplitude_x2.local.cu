// Bin-local device frequency kernel: the reference README's example (README.md:121-129, "multiply the input by a
// constant") in the form rcg_push_local_device_frequency_kernel / `--freq-kernel x.local.cu` take.  The output of a
// bin may depend on that bin's input, its index and the time only; NVRTC then builds the whole stretch kernel
// around the function and |apply(X[k])| is evaluated in registers -- no spectrum is ever written to memory.
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, float2 x) {
    return make_float2(x.x * 2.0f, x.y * 2.0f);
}

// Device frequency kernel: the same example for rcg_push_device_frequency_kernel (compiled with NVRTC
// for sm_100a; the spectra never leave HBM).  The function returns the output value of `bin` given
// the frame's whole W-bin input spectrum, so gather-style kernels (bin shifts, smoothing) work too.
__device__ float2 apply(size_t elapsed_ms, unsigned bin, unsigned n_bins, const float2* spectrum) {
    const float2 v = spectrum[bin];
    return make_float2(v.x * 2.0f, v.y * 2.0f);
}

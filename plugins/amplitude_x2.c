/* Host frequency kernel: the reference README's example (README.md:121-129, "multiply the input by a
 * constant") in the C-ABI form of `apply` (include/rocoder_b200.h; reference fft.rs:93).
 *   gcc -O3 -shared -fPIC -o amplitude_x2.so amplitude_x2.c
 *   rocoder_gpu -i in.wav -o out.wav --freq-kernel amplitude_x2.so      (or pass the .c: the CLI compiles it) */
#include <stddef.h>

int apply(size_t elapsed_ms, const float *in_re_im, size_t n_bins, float *out_re_im) {
    (void)elapsed_ms;
    for (size_t i = 0; i < 2 * n_bins; i++) out_re_im[i] = in_re_im[i] * 2.0f;
    return 0; /* non-zero = "panicked": the engine falls back to the previous kernel (fft.rs:101-106) */
}

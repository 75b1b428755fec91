"""Shared parity cases: (n_in, factor, pitch, window_len).  Sized so the oracle finishes in
seconds.  They cover every resample branch (p=1, p>1 incl. odd, p<-1 incl. the dropped
samples quirk at p<=-3), ragged / short / empty input, and every FFT plan shape."""

SMALL_CASES = [
    (5000, 1.0, 1, 256),
    (5000, 2.0, 2, 256),
    (7001, 3.0, -2, 512),
    (9000, 20.0, -3, 1024),     # BASELINE config 4's flags at a small window
    (3000, 1.5, 3, 128),
    (100, 1.0, 1, 64),          # input shorter than one window
    (0, 1.0, 1, 64),            # empty input: one all-zero call
    (4097, 0.5, 1, 2048),       # speed-up: step == W
    (30000, 4.0, 1, 2048),
    (40000, 8.0, 2, 4096),
    (50000, 2.0, -4, 8192),
    (6000, 2.5, 5, 512),
    (2048, 1.0, 1, 1024),       # exactly two windows
    (8000, 1.0, 7, 256),        # large odd harmonic: 14 frames per call
    (8000, 7.0, -7, 512),       # deep subharmonic: S = 74 of 256 finished samples kept per call
    (5000, 1.0, 12, 128),       # even, non power of two (half-rate path with a decimating gather)
]
LARGE_CASES = [
    (70000, 8.0, 2, 16384),     # BASELINE config 2's flags
    (100000, 1.0, 1, 16384),    # config 1
    (60000, 20.0, -3, 16384),   # config 4
    (90000, 3.0, 1, 32768),
    (50000, 2.0, 1, 16384),     # config 5
    (300000, 4.0, 1, 65536),    # config 3's largest window: a frame across a 2-CTA cluster
    (200000, 4.0, 2, 65536),
    (200000, 3.0, -3, 65536),
]

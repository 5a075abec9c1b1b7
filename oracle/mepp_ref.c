/*
 * oracle/mepp_ref.c -- plain C restatement of the reference's per-task scan, TEST INFRASTRUCTURE ONLY
 * (checker and CPU baseline; never linked into or called by the product).
 *
 * Follows npdeloss/mepp: mepp/motif_layers.py:80-165 (valid Conv1D over the normalised one-hot,
 * max over the two filters, ZeroPadding1D((m-1)//2, ...), ReLU) and mepp/core.py:59-145 (margin
 * blur), in the declared canonical float32 order of oracle/profile.py.  PARITY UNPINNED (the
 * reference has no golden vectors; see oracle/__init__.py).  Build: `make -C oracle`.
 * Compile WITHOUT -ffast-math: the float32 addition order is the specification.
 */
#include <stdint.h>
#include <stddef.h>

/* codes: int8 [N][L] in 0..4 (4 = N).  W0/W1: float32 [4][m] filters (W1 may be NULL).
 * X0: float32 [N][L] output.  Callers thread over row ranges (oracle/cref.py); this image has no libgomp. */
void mepp_ref_scan(const int8_t* codes, long N, int L, const float* W0, const float* W1, int m,
                   float bias, float* X0)
{
    const int nout = L - m + 1;
    const int padl = (m - 1) / 2;
    long n;

    for (n = 0; n < N; ++n) {
        const int8_t* s = codes + n * (long)L;
        float* x = X0 + n * (long)L;
        int i, j, f;
        for (i = 0; i < L; ++i) x[i] = 0.0f;
        for (i = 0; i < nout; ++i) {
            float best = 0.0f;
            for (f = 0; f < (W1 ? 2 : 1); ++f) {
                const float* W = f ? W1 : W0;
                volatile float acc = 0.0f; /* volatile: forbid reassociation / vector reductions */
                for (j = 0; j < m; ++j) {
                    int c = s[i + j];
                    if (c < 4) {
                        acc = acc + W[c * m + j];
                    } else {
                        acc = acc + 0.25f * W[0 * m + j];
                        acc = acc + 0.25f * W[1 * m + j];
                        acc = acc + 0.25f * W[2 * m + j];
                        acc = acc + 0.25f * W[3 * m + j];
                    }
                }
                {
                    float sc = acc + bias;
                    if (f == 0 || sc > best) best = sc;
                }
            }
            x[i + padl] = best > 0.0f ? best : 0.0f;
        }
    }
}

/* AveragePooling1D(k = 1 + 2*margin, stride 1, valid) + zero pad `margin` each side, float32. */
void mepp_ref_blur(const float* X0, long N, int L, int margin, float* X)
{
    const int k = 1 + 2 * margin;
    long n;

    for (n = 0; n < N; ++n) {
        const float* a = X0 + n * (long)L;
        float* o = X + n * (long)L;
        int p, d;
        for (p = 0; p < L; ++p) o[p] = 0.0f;
        if (margin == 0) { for (p = 0; p < L; ++p) o[p] = a[p]; continue; }
        for (p = margin; p <= L - 1 - margin; ++p) {
            float acc = a[p - margin];
            for (d = 1; d < k; ++d) acc = acc + a[p - margin + d];
            o[p] = acc / (float)k;
        }
    }
}

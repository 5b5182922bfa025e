/* Correctly rounded fused multiply-add over arrays, for the CPU oracle.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * numpy has no fma; the reference's third-party arithmetic (OpenBLAS dgemm
 * behind np.dot at /root/reference/code/utils.py:47, torch's linspace / bmm
 * behind F.affine_grid at /root/reference/code/models.py:378) uses hardware
 * FMA, so the restatement needs a single-rounding a*b+c.  Build:
 *   gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/liboracle_fma.so oracle/fma_helper.c -lm
 */
#include <math.h>
#include <stddef.h>

void fma64_vec(const double *a, const double *b, const double *c, double *out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = fma(a[i], b[i], c[i]);
}

void fma32_vec(const float *a, const float *b, const float *c, float *out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = fmaf(a[i], b[i], c[i]);
}

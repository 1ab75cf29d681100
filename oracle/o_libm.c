/* ORACLE — test infrastructure only (see o_common.h).
 * this machine's libm, element by element: what the restatement's expf / exp2f / log2f / powf calls evaluate to.
 * the strict kernels reproduce these bit for bit (vkdt_b200/csrc/kernels/libm_exact.h); tests/test_libm_exact_gpu.py compares. */
#include <math.h>
#include <stddef.h>
void o_libm_apply(int op, size_t n, const float *a, const float *b, float *out)
{
#pragma omp parallel for schedule(static)
  for(size_t i = 0; i < n; i++)
    out[i] = op == 0 ? expf(a[i]) : (op == 1 ? exp2f(a[i]) : (op == 2 ? log2f(a[i]) : powf(a[i], b[i])));
}

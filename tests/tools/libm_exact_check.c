// compares vkdt_b200/csrc/kernels/libm_exact.h, compiled for the host, with this machine's libm.
//   libm_exact_check <stride> <pow_pairs>     every stride-th float bit pattern for expf/exp2f/logf/log2f (1 = all 2^32),
//                                            pow_pairs random (x, y) pairs + structured pairs for powf
// prints the number of mismatching results per function (NaNs compare equal when both are NaN); exit code 1 on any.
// build: g++ -x c++ -std=c++17 -O2 -mfma -ffp-contract=off -fopenmp libm_exact_check.c -lm   (test infrastructure, not product)
#include "../../vkdt_b200/csrc/kernels/libm_exact.h"
#include <stdio.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int same(float a, float b)
{
  if(isnan(a) && isnan(b)) return 1;
  return lme_f2u(a) == lme_f2u(b);
}
static inline uint64_t rng(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

int main(int argc, char **argv)
{
  const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 257;
  const uint64_t pairs = argc > 2 ? strtoull(argv[2], 0, 10) : 1ull << 24;
  uint64_t bad[5] = {0, 0, 0, 0, 0};
  uint64_t b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0;
#pragma omp parallel for reduction(+:b0,b1,b2,b3) schedule(static)
  for(uint64_t u = 0; u < (1ull << 32); u += stride)
  {
    const float x = lme_u2f((uint32_t)u);
    if(!same(lme_expf(x), expf(x))) b0++;
    if(!same(lme_exp2f(x), exp2f(x))) b1++;
    if(!same(lme_logf(x), logf(x))) b2++;
    if(!same(lme_log2f(x), log2f(x))) b3++;
  }
  bad[0] = b0; bad[1] = b1; bad[2] = b2; bad[3] = b3;
  // powf: random bit patterns, random "ordinary" pairs (x in (0, 65504], |y| <= 8), and every float x in [0, 2) for the
  // exponents the path uses
#pragma omp parallel reduction(+:b4)
  {
    uint64_t s = 0x9E3779B97F4A7C15ull ^ (0x1234567ull * (uint64_t)(1 +
#ifdef _OPENMP
        omp_get_thread_num()
#else
        0
#endif
        ));
#pragma omp for schedule(static)
    for(uint64_t n = 0; n < pairs; n++)
    {
      const uint64_t r = rng(&s);
      float x = lme_u2f((uint32_t)r), y = lme_u2f((uint32_t)(r >> 32));
      if(!same(lme_powf(x, y), powf(x, y))) b4++;
      const uint64_t q = rng(&s);
      x = ldexpf((float)((q & 0xffffff) | 0x800000) / 16777216.0f, (int)((q >> 24) % 40) - 23);
      y = ((float)((q >> 32) & 0xffffff) / 16777216.0f - 0.5f) * 16.0f;
      if(!same(lme_powf(x, y), powf(x, y))) b4++;
    }
  }
  const float ys[] = { 0.8f, 1.0f / 2.2f, 2.2f, 2.4f, 1.0f / 2.4f, 1.0f / 3.0f, 0.631651345306265f, 0.6523997524738018f, 0.6007557017508491f,
                       0.8322850678616855f, 1.5831518565279648f, 1.2f, 0.2f, 2.6f, 0.5f, 3.0f, -1.0f, 2.0f, 0.45f };
  for(size_t j = 0; j < sizeof(ys) / sizeof(ys[0]); j++)
  {
    uint64_t b = 0;
#pragma omp parallel for reduction(+:b) schedule(static)
    for(uint64_t u = 0; u <= 0x40000000u; u += (stride > 16 ? 17 : 1))
    {
      const float x = lme_u2f((uint32_t)u);
      if(!same(lme_powf(x, ys[j]), powf(x, ys[j]))) b++;
      if(fabsf(ys[j]) <= 0.84f && !same(lme_powf_smally(x, ys[j]), powf(x, ys[j]))) b++;
      // +0, positive normal, (+inf, nan: below) and |y log2 x| < 120, positive y: the variant that answers the specials by selects
      if((u == 0 || u >= 0x00800000u) && ys[j] > 0.0f && (u == 0 || fabs((double)ys[j] * log2((double)x)) < 120.0) && !same(lme_powf_nonneg(x, ys[j]), powf(x, ys[j]))) b++;
      // +0 or positive normal x <= 1, positive y of any size: the variant that also answers underflow in line (denoise's test^16)
      if((u == 0 || u >= 0x00800000u) && u <= 0x3f800000u && ys[j] > 0.0f)
      {
        if(!same(lme_powf_nonneg_le1(x, ys[j]), powf(x, ys[j]))) b++;
        if(!same(lme_powf_nonneg_le1(x, 16.0f), powf(x, 16.0f))) b++;
        if(!same(lme_powf_nonneg_le1(x, 40.0f), powf(x, 40.0f))) b++;
      }
      // positive normal x, |y log2 x| < 120: the variant without any special case (the default tone curve's power)
      if(u >= 0x00800000u && fabs((double)ys[j] * log2((double)x)) < 120.0 && !same(lme_powf_safe(x, ys[j]), powf(x, ys[j]))) b++;
    }
    b4 += b;
  }
  { const float sp[3] = { INFINITY, NAN, lme_u2f(0x7fc12345u) };
    for(int k = 0; k < 3; k++) { if(!same(lme_powf_nonneg(sp[k], 0.8f), powf(sp[k], 0.8f))) b4++; if(!same(lme_powf_nonneg(sp[k], 4.0f), powf(sp[k], 4.0f))) b4++; } }
  bad[4] = b4;
  const char *name[5] = { "expf", "exp2f", "logf", "log2f", "powf" };
  int rc = 0;
  for(int k = 0; k < 5; k++) { printf("%s %llu\n", name[k], (unsigned long long)bad[k]); if(bad[k]) rc = 1; }
  return rc;
}

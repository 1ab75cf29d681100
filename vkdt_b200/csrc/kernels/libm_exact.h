// single precision exp / exp2 / log / log2 / pow whose results are those of glibc 2.39's libm BIT FOR BIT, on the device.
//
// why: every edge of the reference graph is an f16 image.  the SFU's ex2.approx / lg2.approx (what __expf / __powf compile
// to) differ from libm in the last fp32 bits, which flips an f16 rounding once in a few thousand values per edge; a flip in
// front of a discontinuous decision (demosaic's eigenvector snap, denoise's covariance pick) or stacked through llap's
// pyramid is the whole tail beyond BASELINE.json's max-abs 1e-3 against the oracle.  the oracle
// (oracle/*.c, and the reference's shaders compiled as C++ it is pinned to) calls libm, so the strict kernels evaluate
// the same algorithm with the same operation order and the same fused multiply-adds:
//   glibc 2.39 sysdeps/ieee754/flt-32/{e_expf,e_exp2f,e_logf,e_log2f,e_powf}.c (Szabolcs Nagy's ARM optimized routines):
//   table driven range reduction and a short polynomial, all in double, result rounded to float once.
//   the operation sequence below (which products are fused into fma) is the one of the x86-64 `_fma` ifunc variants the
//   dynamic loader selects on any AVX2+FMA host (read off the disassembly of this image's libm.so.6); the tables are
//   __exp2f_data / __logf_data / __log2f_data / __powf_log2_data.
// B200 runs fp64 at half the fp32 rate, one call is 6-14 double instructions: see DESIGN.md for what strict costs.
// tests/test_libm_exact_cpu.py compares this header (compiled for the host, -ffp-contract=off) with libm over all 2^32
// float arguments of the one-argument functions and 2^32 random pairs of powf; the GPU tests compare the device code.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define LME_FN __host__ __device__ __forceinline__
#define LME_MFN __host__ __device__ __forceinline__
#define LME_COLD static __host__ __device__ __noinline__
#else
#define LME_FN static inline
#define LME_MFN inline
#define LME_COLD static __attribute__((noinline))
#endif

#if defined(__CUDA_ARCH__)
#define LME_T(name) name##_dev
#define LME_FMA(a, b, c) __fma_rn((a), (b), (c))
#define LME_MUL(a, b) __dmul_rn((a), (b))
#define LME_ADD(a, b) __dadd_rn((a), (b))
#define LME_F2U(x) __float_as_uint(x)
#define LME_U2F(x) __uint_as_float(x)
#define LME_D2U(x) ((uint64_t)__double_as_longlong(x))
#define LME_U2D(x) __longlong_as_double((long long)(x))
#define LME_D2F(x) __double2float_rn(x)
#else
#define LME_T(name) name##_host
#define LME_FMA(a, b, c) fma((a), (b), (c))
#define LME_MUL(a, b) ((a) * (b))
#define LME_ADD(a, b) ((a) + (b))
static inline uint32_t lme_f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float lme_u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
static inline uint64_t lme_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double lme_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define LME_F2U(x) lme_f2u(x)
#define LME_U2F(x) lme_u2f(x)
#define LME_D2U(x) lme_d2u(x)
#define LME_U2D(x) lme_u2d(x)
#define LME_D2F(x) ((float)(x))
#endif

// the tables exist twice under nvcc: in device global memory (256 B each, L1 resident) and in host memory
#define LME_TABLES(SUFFIX, QUAL) \
QUAL uint64_t lme_exp2_tab##SUFFIX[32] = { 0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull }; \
QUAL double lme_logf_tab##SUFFIX[32] = { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, 0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3, 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0, 0x0.0p+0, 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4, 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3, 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2, 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 }; \
QUAL double lme_log2f_tab##SUFFIX[32] = { 0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2, 0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2, 0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3, 0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4, 0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1.0000000000000p+0, 0x0.0p+0, 0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3, 0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2, 0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2 };
#if defined(__CUDACC__)
LME_TABLES(_dev, static __device__ const)
#endif
LME_TABLES(_host, static const)

#define LME_SHIFT        0x1.8p+52                  /* __exp2f_data.shift */
#define LME_SHIFT_SCALED 0x1.8p+47                  /* shift / 32 */
#define LME_INVLN2N      0x1.71547652b82fep+5    /* 32 / ln 2 */
#define LME_EXP2_C0 0x1.c6af84b912394p-5
#define LME_EXP2_C1 0x1.ebfce50fac4f3p-3
#define LME_EXP2_C2 0x1.62e42ff0c52d6p-1
#define LME_EXP_C0  0x1.c6af84b912394p-20
#define LME_EXP_C1  0x1.ebfce50fac4f3p-13
#define LME_EXP_C2  0x1.62e42ff0c52d6p-6
#define LME_LN2      0x1.62e42fefa39efp-1
#define LME_LOG_A0  -0x1.00ea348b88334p-2
#define LME_LOG_A1  0x1.5575b0be00b6ap-2
#define LME_LOG_A2  -0x1.ffffef20a4123p-2
#define LME_LOG2_A0 -0x1.712b6f70a7e4dp-2
#define LME_LOG2_A1 0x1.ecabf496832e0p-2
#define LME_LOG2_A2 -0x1.715479ffae3dep-1
#define LME_LOG2_A3 0x1.715475f35c8b8p+0
#define LME_POW_A0  0x1.27616c9496e0bp-2
#define LME_POW_A1  -0x1.71969a075c67ap-2
#define LME_POW_A2  0x1.ec70a6ca7baddp-2
#define LME_POW_A3  -0x1.7154748bef6c8p-1
#define LME_POW_A4  0x1.71547652ab82bp+0

// 2^(k/32) * poly(r), the tail shared by expf, exp2f and powf (exp2_inline of e_powf.c with sign_bias = 0)
LME_FN double lme_exp2_tail(double kd_plus_shift, double r, double c0, double c1, double c2)
{
  const uint64_t ki = LME_D2U(kd_plus_shift);
  uint64_t t = LME_T(lme_exp2_tab)[ki & 31];
  t += ki << 47;
  const double s = LME_U2D(t);
  const double z = LME_FMA(c0, r, c1);
  const double r2 = LME_MUL(r, r);
  double y = LME_FMA(c2, r, 1.0);
  y = LME_FMA(z, r2, y);
  return LME_MUL(y, s);
}

LME_FN float lme_expf_gen(float x)
{
  const uint32_t abstop = (LME_F2U(x) >> 20) & 0x7ff;
  if(abstop > 0x42a)
  { // |x| >= 88 or not finite
    if(LME_F2U(x) == 0xff800000u) return 0.0f;
    if(abstop > 0x7f7) return x + x;
    if(x > 0x1.62e42ep6f) return INFINITY;
    if(x < -0x1.9fe368p6f) return 0.0f;
  }
  const double xd = (double)x;
  double kd = LME_FMA(LME_INVLN2N, xd, LME_SHIFT);
  const double ks = kd;
  kd = LME_ADD(kd, -LME_SHIFT);
  const double r = LME_FMA(LME_INVLN2N, xd, -kd);
  return LME_D2F(lme_exp2_tail(ks, r, LME_EXP_C0, LME_EXP_C1, LME_EXP_C2));
}

LME_FN float lme_exp2f(float x)
{
  const uint32_t abstop = (LME_F2U(x) >> 20) & 0x7ff;
  if(abstop > 0x42f)
  { // |x| >= 128 or not finite
    if(LME_F2U(x) == 0xff800000u) return 0.0f;
    if(abstop > 0x7f7) return x + x;
    if(x > 0.0f) return INFINITY;
    if(x <= -150.0f) return 0.0f;
  }
  const double xd = (double)x;
  double kd = LME_ADD(xd, LME_SHIFT_SCALED);
  const double ks = kd;
  kd = LME_ADD(kd, -LME_SHIFT_SCALED);
  const double r = LME_ADD(xd, -kd);
  return LME_D2F(lme_exp2_tail(ks, r, LME_EXP2_C0, LME_EXP2_C1, LME_EXP2_C2));
}

// common front end of logf / log2f / powf: x = 2^k * z, z in [0x1.66p-1, 0x1.66p0), table index from the top mantissa bits
#define LME_LOG_FRONT(ix, TAB)                                         \
  const uint32_t tmp = (ix) - 0x3f330000u;                             \
  const int i = (tmp >> 19) & 15;                                      \
  const uint32_t top = tmp & 0xff800000u;                              \
  const uint32_t iz = (ix) - top;                                      \
  const int k = (int32_t)tmp >> 23;                                    \
  const double invc = LME_T(TAB)[2 * i], logc = LME_T(TAB)[2 * i + 1]; \
  const double z = (double)LME_U2F(iz);                                \
  const double r = LME_FMA(z, invc, -1.0);

LME_FN float lme_logf(float x)
{
  uint32_t ix = LME_F2U(x);
  if(ix == 0x3f800000u) return 0.0f;
  if(ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
  {
    if(ix * 2 == 0) return -INFINITY;
    if(ix == 0x7f800000u) return x;
    if((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
    ix = LME_F2U(x * 0x1p23f);
    ix -= 23u << 23;
  }
  LME_LOG_FRONT(ix, lme_logf_tab)
  const double y0 = LME_FMA((double)k, LME_LN2, logc);
  const double r2 = LME_MUL(r, r);
  double y = LME_FMA(LME_LOG_A1, r, LME_LOG_A2);
  y = LME_FMA(LME_LOG_A0, r2, y);
  return LME_D2F(LME_FMA(y, r2, LME_ADD(y0, r)));
}

LME_FN float lme_log2f(float x)
{
  uint32_t ix = LME_F2U(x);
  if(ix == 0x3f800000u) return 0.0f;
  if(ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
  {
    if(ix * 2 == 0) return -INFINITY;
    if(ix == 0x7f800000u) return x;
    if((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
    ix = LME_F2U(x * 0x1p23f);
    ix -= 23u << 23;
  }
  LME_LOG_FRONT(ix, lme_log2f_tab)
  const double y0 = LME_ADD(logc, (double)k);
  const double r2 = LME_MUL(r, r);
  double y = LME_FMA(LME_LOG2_A1, r, LME_LOG2_A2);
  y = LME_FMA(LME_LOG2_A0, r2, y);
  const double p = LME_FMA(LME_LOG2_A3, r, y0);
  return LME_D2F(LME_FMA(y, r2, p));
}

// x^y for x >= 0 or integer y like libm; the shaders guard their arguments (pow of a negative base is undefined in GLSL)
LME_FN float lme_powf_gen(float x, float y)
{
  uint32_t ix = LME_F2U(x);
  const uint32_t iy = LME_F2U(y);
  uint64_t sign_bias = 0;
  if(ix - 0x00800000u >= 0x7f800000u - 0x00800000u || 2 * iy - 1 >= 2u * 0x7f800000u - 1)
  { // x is zero, subnormal, negative, inf or nan, or y is zero, inf or nan (e_powf.c)
    if(2 * iy - 1 >= 2u * 0x7f800000u - 1)
    {
      if(2 * iy == 0) return ((ix ^ 0x00400000u) & 0x7fffffffu) > 0x7fc00000u ? x + y : 1.0f;
      if(ix == 0x3f800000u) return ((iy ^ 0x00400000u) & 0x7fffffffu) > 0x7fc00000u ? x + y : 1.0f;
      if(2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return x + y;
      if(2 * ix == 2 * 0x3f800000u) return 1.0f;
      if((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;
      return y * y;
    }
    // integer test of y for negative bases: 0 not an integer, 1 odd, 2 even
    int yint = 0;
    {
      const int e = (iy >> 23) & 0xff;
      if(e >= 0x7f) { if(e > 0x7f + 23) yint = 2; else if(!(iy & ((1u << (0x7f + 23 - e)) - 1))) yint = (iy & (1u << (0x7f + 23 - e))) ? 1 : 2; }
    }
    if(2 * ix - 1 >= 2u * 0x7f800000u - 1)
    { // x is +-0, +-inf or nan
      float x2 = x * x;
      if((ix & 0x80000000u) && yint == 1) x2 = -x2;
      return (iy & 0x80000000u) ? 1.0f / x2 : x2;
    }
    if(ix & 0x80000000u)
    {
      if(yint == 0) return NAN;
      if(yint == 1) sign_bias = 1ull << 16; // SIGN_BIAS = 1 << (EXP2F_TABLE_BITS + 11)
      ix &= 0x7fffffffu;
    }
    if(ix < 0x00800000u)
    {
      ix = LME_F2U(LME_U2F(ix) * 0x1p23f);
      ix &= 0x7fffffffu;
      ix -= 23u << 23;
    }
  }
  LME_LOG_FRONT(ix, lme_log2f_tab)
  const double y0 = LME_ADD(logc, (double)k);
  const double r2 = LME_MUL(r, r);
  const double yy = LME_FMA(LME_POW_A0, r, LME_POW_A1);
  const double p = LME_FMA(LME_POW_A2, r, LME_POW_A3);
  const double r4 = LME_MUL(r2, r2);
  double q = LME_FMA(LME_POW_A4, r, y0);
  q = LME_FMA(p, r2, q);
  const double logx = LME_FMA(yy, r4, q);
  const double ylogx = LME_MUL((double)y, logx);
  if(((LME_D2U(ylogx) >> 47) & 0xffff) >= 0x80bf)
  { // |y log2 x| >= 126
    if(ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? -INFINITY : INFINITY;
    if(ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
  }
  double kd = LME_ADD(ylogx, LME_SHIFT_SCALED);
  const uint64_t ki = LME_D2U(kd);
  kd = LME_ADD(kd, -LME_SHIFT_SCALED);
  const double rr = LME_ADD(ylogx, -kd);
  uint64_t t = LME_T(lme_exp2_tab)[ki & 31];
  t += (ki + sign_bias) << 47;
  const double s = LME_U2D(t);
  const double zz = LME_FMA(LME_EXP2_C0, rr, LME_EXP2_C1);
  const double rr2 = LME_MUL(rr, rr);
  double res = LME_FMA(LME_EXP2_C2, rr, 1.0);
  res = LME_FMA(zz, rr2, res);
  return LME_D2F(LME_MUL(res, s));
}

// ---- what the kernels call: the same results, without a data dependent branch in the common case ----
// the generic functions above spend a third of their issued instructions on libm's special case ladders, and every double
// literal costs two moves in front of the instruction that uses it.  here the common case is one straight line, everything
// rare is one predicated call of the generic function kept out of line, the polynomial coefficients sit in the constant
// bank (an operand of DFMA / DMUL, no instruction), and the two tables can be staged in shared memory by the kernel
// (lme_smem_t: a 32 bit address instead of 64 bit arithmetic on a global one).
LME_COLD float lme_powf_cold(float x, float y) { return lme_powf_gen(x, y); }

#define LME_KI_INVLN2N 0
#define LME_KI_EXP_C0  1
#define LME_KI_EXP_C1  2
#define LME_KI_EXP_C2  3
#define LME_KI_EXP2_C0 4
#define LME_KI_EXP2_C1 5
#define LME_KI_EXP2_C2 6
#define LME_KI_POW_A0  7
#define LME_KI_POW_A1  8
#define LME_KI_POW_A2  9
#define LME_KI_POW_A3  10
#define LME_KI_POW_A4  11
#if defined(__CUDACC__)
static __constant__ double lme_kc[12] = { LME_INVLN2N, LME_EXP_C0, LME_EXP_C1, LME_EXP_C2, LME_EXP2_C0, LME_EXP2_C1, LME_EXP2_C2,
                                          LME_POW_A0, LME_POW_A1, LME_POW_A2, LME_POW_A3, LME_POW_A4 };
struct __align__(16) lme_smem_t { double log2tab[32]; uint64_t exp2tab[32]; };
// the first 32 threads of a CTA fill the tables; the caller synchronises before the first use
__device__ __forceinline__ void lme_smem_fill(lme_smem_t &s, int tid)
{
  if(tid < 32) { s.log2tab[tid] = lme_log2f_tab_dev[tid]; s.exp2tab[tid] = lme_exp2_tab_dev[tid]; }
}
#endif
#if defined(__CUDA_ARCH__)
#define LME_K(n) lme_kc[LME_KI_##n]
#else
#define LME_K(n) LME_##n
#endif

// expf: arguments below -104 give +0 like libm's underflow return (the formula rounds 2^-150.0x to zero as well), above 89
// +inf like its overflow return (the conversion overflows)
template <class TAB> LME_FN float lme_expf_t(float x, const TAB tab)
{
  const float xc = fminf(fmaxf(x, -104.0f), 89.0f);
  const double xd = (double)xc;
  double kd = LME_FMA(LME_K(INVLN2N), xd, LME_SHIFT);
  const uint64_t ki = LME_D2U(kd);
  kd = LME_ADD(kd, -LME_SHIFT);
  const double r = LME_FMA(LME_K(INVLN2N), xd, -kd);
  uint64_t t = tab.exp2(ki);
  t += ki << 47;
  const double sc = LME_U2D(t);
  const double z = LME_FMA(LME_K(EXP_C0), r, LME_K(EXP_C1));
  const double r2 = LME_MUL(r, r);
  double y = LME_FMA(LME_K(EXP_C2), r, 1.0);
  y = LME_FMA(z, r2, y);
  const float res = LME_D2F(LME_MUL(y, sc));
  return x != x ? x + x : res;
}

// powf: straight line for a positive normal x and a finite non zero y whose y log2 x stays inside +-126; x = +0 with a
// positive y is answered in line (black pixels are common); the rest (subnormal, negative, inf, nan, over/underflow) is rare
// SMALLY: the caller guarantees |y| <= 0.84, so that |y log2 x| < 126 for every positive finite x (|log2 x| < 150) and the
// over / underflow test of the general case is dead
// POSX: the caller guarantees a positive, normal, finite x and a finite non zero y: no special case is left in front of the
// logarithm; together with SMALLY (here: |y log2 x| < 126 checked by the caller for its range of x) nothing can call
// POSX == 2: x is +0, positive normal, +inf or nan (an image value that cannot be negative or subnormal, e.g. a non negative
// combination of f16 texels) and y is positive and finite: the three special values are answered by selects, nothing calls
template <bool SMALLY, int POSX, class TAB> LME_FN float lme_powf_ttt(float x, float y, const TAB tab)
{
  const uint32_t ix = LME_F2U(x), iy = LME_F2U(y);
  int rare = POSX == 1 ? 0 : ((ix - 0x00800000u >= 0x7f000000u) || (POSX == 0 && (2 * iy - 1 >= 2u * 0x7f800000u - 1)));
  const uint32_t ixs = rare ? 0x3f800000u : ix;
  const uint32_t tmp = ixs - 0x3f330000u;
  const int i = (tmp >> 19) & 15;
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ixs - top;
  const int k = (int32_t)tmp >> 23;
  double invc, logc; tab.log2(i, invc, logc);
  const double z = (double)LME_U2F(iz);
  const double r = LME_FMA(z, invc, -1.0);
  const double y0 = LME_ADD(logc, (double)k);
  const double r2 = LME_MUL(r, r);
  const double yy = LME_FMA(LME_K(POW_A0), r, LME_K(POW_A1));
  const double p = LME_FMA(LME_K(POW_A2), r, LME_K(POW_A3));
  const double r4 = LME_MUL(r2, r2);
  double q = LME_FMA(LME_K(POW_A4), r, y0);
  q = LME_FMA(p, r2, q);
  const double logx = LME_FMA(yy, r4, q);
  const double ylogx = LME_MUL((double)y, logx);
  double ylogx_c = ylogx;
  if(!SMALLY && POSX == 2)
  { // underflow answered in line: at y log2 x <= -150 libm returns +0, and so does the formula from -150 down (2^-151 rounds to
    // zero); the clamp keeps the exponent arithmetic in range.  (the overflow side is the caller's to exclude.)
    ylogx_c = ylogx < -151.0 ? -151.0 : ylogx;
  }
  else if(!SMALLY) rare = rare || (((uint32_t)(LME_D2U(ylogx) >> 47) & 0xffffu) >= 0x80bfu);
  double kd = LME_ADD(ylogx_c, LME_SHIFT_SCALED);
  const uint64_t ki = LME_D2U(kd);
  kd = LME_ADD(kd, -LME_SHIFT_SCALED);
  const double rr = LME_ADD(ylogx_c, -kd);
  uint64_t t = tab.exp2(ki);
  t += ki << 47;
  const double sc = LME_U2D(t);
  const double zz = LME_FMA(LME_K(EXP2_C0), rr, LME_K(EXP2_C1));
  const double rr2 = LME_MUL(rr, rr);
  double res = LME_FMA(LME_K(EXP2_C2), rr, 1.0);
  res = LME_FMA(zz, rr2, res);
  const float out = LME_D2F(LME_MUL(res, sc));
  if(POSX == 2) return (ix << 1) == 0 ? 0.0f : (ix >= 0x7f800000u ? x + x : out);   // +-0 -> +0 (y is no odd integer), +inf -> +inf, nan -> nan
  if(rare)
  {
    if(ix == 0 && iy - 1u < 0x7f7fffffu) return 0.0f;   // +0 ^ (positive finite y)
    return lme_powf_cold(x, y);
  }
  return out;
}
template <bool SMALLY, class TAB> LME_FN float lme_powf_tt(float x, float y, const TAB tab) { return lme_powf_ttt<SMALLY, 0>(x, y, tab); }
template <class TAB> LME_FN float lme_powf_t(float x, float y, const TAB tab) { return lme_powf_tt<false>(x, y, tab); }
// tables in global memory (L1 resident)
struct lme_gtab_t
{
  LME_MFN uint64_t exp2(uint64_t ki) const { return LME_T(lme_exp2_tab)[ki & 31]; }
  LME_MFN void log2(int i, double &invc, double &logc) const { invc = LME_T(lme_log2f_tab)[2 * i]; logc = LME_T(lme_log2f_tab)[2 * i + 1]; }
};
LME_FN float lme_expf(float x) { return lme_expf_t(x, lme_gtab_t()); }
LME_FN float lme_powf(float x, float y) { return lme_powf_t(x, y, lme_gtab_t()); }
LME_FN float lme_powf_smally(float x, float y) { return lme_powf_tt<true>(x, y, lme_gtab_t()); }
LME_FN float lme_powf_safe(float x, float y) { return lme_powf_ttt<true, 1>(x, y, lme_gtab_t()); }
LME_FN float lme_powf_nonneg(float x, float y) { return lme_powf_ttt<true, 2>(x, y, lme_gtab_t()); }
LME_FN float lme_powf_nonneg_le1(float x, float y) { return lme_powf_ttt<false, 2>(x, y, lme_gtab_t()); }   // x <= 1: may underflow, cannot overflow
#if defined(__CUDACC__)
// tables in shared memory: `base` is the 32 bit shared address of a filled lme_smem_t, kept in one register
struct lme_stab_t
{
  uint32_t base;
  __device__ __forceinline__ uint64_t exp2(uint64_t ki) const
  {
    uint64_t t; asm("ld.shared.u64 %0, [%1+256];" : "=l"(t) : "r"(base + (((uint32_t)ki & 31u) << 3))); return t;
  }
  __device__ __forceinline__ void log2(int i, double &invc, double &logc) const
  {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(logc) : "r"(base + ((uint32_t)i << 4)));
  }
};
__device__ __forceinline__ lme_stab_t lme_stab(const lme_smem_t &s)
{
  lme_stab_t t; t.base = (uint32_t)__cvta_generic_to_shared(&s);
  asm volatile("" : "+r"(t.base));   // one register for the kernel's lifetime instead of five uniform instructions per use
  return t;
}
#endif

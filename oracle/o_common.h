/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp32, OpenMP) of the GLSL kernels on vkdt's raw->display path.
 * Nothing under oracle/ is shipped or measured as product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it (as the checker).
 *
 * Parity status: the MLV bit-unpack restatement is pinned against the reference's own compiled
 * code (oracle/_ref/libmlvref.so, built from src/pipe/modules/i-mlv/video_mlv.c); the host side
 * (crop / colour parameter blocks, node graphs) against its own main.c files (libhostref.so); the
 * float kernels against the reference's own compute shaders compiled as C++ (oracle/glsl,
 * libshaderref.so, tests/test_shader_ref_cpu.py: bit exact, one f16 ulp on a few values for the
 * kernels that filter at fractional coordinates; demosaic/rcd_fill away from the reference's
 * tile seams, where its output depends on the tiling).  What stays unpinned: no
 * Vulkan driver's output is available: the reference's pipeline cannot be run in this container
 * (no vulkan headers/loader/ICD, no glslang) and its tree holds no golden images.
 *
 * Conventions that every restated kernel follows:
 *  - an image is w*h*c floats, c in {1,4}; a value stored through an f16 connector has been
 *    rounded to binary16 (RNE) by the producing kernel ("imageStore to f16 image").
 *  - texture(): bilinear, normalised coordinates, MIRRORED_REPEAT, exact float weights
 *    (src/qvk/qvk.c:596-611).
 *  - texelFetch(): exact texel, out-of-range coordinates clamp to the edge (Vulkan leaves this
 *    undefined; documented in DESIGN.md).
 *  - a 1-channel image read as vec4 gives (r,0,0,1).
 */
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oimg_t
{
  int w, h, c;   /* c = 1 or 4 */
  float *p;      /* w*h*c floats, row major, channel interleaved */
} oimg_t;

/* ---- binary16 round trip, round-to-nearest-even, overflow -> inf (imageStore to f16) ---- */
static inline uint16_t o_f32_to_f16_bits(float f)
{
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if(x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((x > 0x7f800000u) ? 0x200u : 0u)); /* inf/nan */
  if(x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);            /* rounds to >= 65520 -> inf */
  if(x <  0x33000001u) return (uint16_t)sign;                        /* < 2^-25 (or == 2^-25: ties to even 0) */
  if(x <  0x38800000u)
  { /* subnormal half */
    const int e = (int)(x >> 23);              /* biased exponent, 102..112 */
    uint32_t m = (x & 0x7fffffu) | 0x800000u;  /* 24 bit mantissa */
    const int shift = 126 - e;                 /* 14..24 */
    uint32_t r = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    if(rem > half || (rem == half && (r & 1u))) r++;
    return (uint16_t)(sign | r);
  }
  { /* normal */
    uint32_t r = (x - 0x38000000u) >> 13;
    const uint32_t rem = x & 0x1fffu;
    if(rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
    return (uint16_t)(sign | r);
  }
}
static inline float o_f16_bits_to_f32(uint16_t h)
{
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1f, m = h & 0x3ffu, x;
  if(e == 0)
  {
    if(m == 0) x = sign;
    else
    { /* subnormal */
      int s = 0;
      while(!(m & 0x400u)) { m <<= 1; s++; }
      m &= 0x3ffu;
      x = sign | ((uint32_t)(113 - s) << 23) | (m << 13);
    }
  }
  else if(e == 31) x = sign | 0x7f800000u | (m << 13);
  else x = sign | ((e + 112u) << 23) | (m << 13);
  float f; memcpy(&f, &x, 4);
  return f;
}
static inline float o_f16r(float f) { return o_f16_bits_to_f32(o_f32_to_f16_bits(f)); }

/* ---- image helpers ---- */
static inline oimg_t o_img_alloc(int w, int h, int c)
{
  oimg_t im = { w, h, c, 0 };
  im.p = (float *)calloc((size_t)w * h * c, sizeof(float));
  return im;
}
static inline void o_img_free(oimg_t *im) { free(im->p); im->p = 0; }

static inline int o_clampi(int v, int a, int b) { return v < a ? a : (v > b ? b : v); }
static inline int o_mirror(int i, int n)
{ /* VK_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT on texel indices */
  const int p = 2 * n;
  i %= p; if(i < 0) i += p;
  return i >= n ? p - 1 - i : i;
}
/* read texel as vec4 with vulkan's channel fill rules */
static inline void o_px4(const oimg_t *im, int x, int y, float *o)
{
  const float *s = im->p + ((size_t)y * im->w + x) * im->c;
  if(im->c == 4)      { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; }
  else if(im->c == 2) { o[0] = s[0]; o[1] = s[1]; o[2] = 0.0f; o[3] = 1.0f; }
  else                { o[0] = s[0]; o[1] = 0.0f; o[2] = 0.0f; o[3] = 1.0f; }
}
/* texelFetch(img, ivec2(x,y), 0): clamp out of range to the edge */
static inline void o_fetch4(const oimg_t *im, int x, int y, float *o)
{
  o_px4(im, o_clampi(x, 0, im->w - 1), o_clampi(y, 0, im->h - 1), o);
}
static inline float o_fetch1(const oimg_t *im, int x, int y)
{
  x = o_clampi(x, 0, im->w - 1); y = o_clampi(y, 0, im->h - 1);
  return im->p[((size_t)y * im->w + x) * im->c];
}
/* texture(img, vec2(u,v)): bilinear, mirrored repeat, normalised coordinates */
static inline void o_tex4(const oimg_t *im, double u, double v, float *o)
{
  /* the oracle is an IDEAL sampler: texture coordinates are carried in double so that the fp32 noise a shader's
   * (i+0.5)/size division adds (implementation specific, below any texture unit's weight precision) is absent */
  double x = u * (double)im->w - 0.5, y = v * (double)im->h - 0.5;
  /* taps the shaders aim at texel centres ((i+0.5)/size) come back from the float division a few ulps off.
   * every real texture unit resolves those to the exact texel (NVIDIA filters with 8 fractional bits, Vulkan
   * requires >= 4): snap coordinates closer than 1/4096 to a texel centre.  genuinely fractional taps
   * (flower +-1.2/0.4, soft +-1.5, semisoft +-0.5) keep exact float weights. */
  if(fabs(x - rint(x)) < 1.0 / 4096.0) x = rint(x);
  if(fabs(y - rint(y)) < 1.0 / 4096.0) y = rint(y);
  const double fx = floor(x), fy = floor(y);
  const float ax = (float)(x - fx), ay = (float)(y - fy);
  const int x0 = o_mirror((int)fx, im->w), x1 = o_mirror((int)fx + 1, im->w);
  const int y0 = o_mirror((int)fy, im->h), y1 = o_mirror((int)fy + 1, im->h);
  float t00[4], t10[4], t01[4], t11[4];
  o_px4(im, x0, y0, t00); o_px4(im, x1, y0, t10);
  o_px4(im, x0, y1, t01); o_px4(im, x1, y1, t11);
  for(int k = 0; k < 4; k++)
    o[k] = (t00[k] * (1.0f - ax) + t10[k] * ax) * (1.0f - ay)
         + (t01[k] * (1.0f - ax) + t11[k] * ax) * ay;
}
static inline float o_tex1(const oimg_t *im, double u, double v)
{
  float t[4]; o_tex4(im, u, v, t); return t[0];
}
/* textureGather(img, vec2(u,v), 0): x=(i0,j1) y=(i1,j1) z=(i1,j0) w=(i0,j0), red channel */
static inline void o_gather(const oimg_t *im, double u, double v, float *o)
{
  const double x = u * (double)im->w - 0.5, y = v * (double)im->h - 0.5;
  const int fx = (int)floor(x + 1e-6), fy = (int)floor(y + 1e-6);
  const int x0 = o_mirror(fx, im->w), x1 = o_mirror(fx + 1, im->w);
  const int y0 = o_mirror(fy, im->h), y1 = o_mirror(fy + 1, im->h);
  o[0] = im->p[((size_t)y1 * im->w + x0) * im->c];
  o[1] = im->p[((size_t)y1 * im->w + x1) * im->c];
  o[2] = im->p[((size_t)y0 * im->w + x1) * im->c];
  o[3] = im->p[((size_t)y0 * im->w + x0) * im->c];
}
/* imageStore helpers (f16 or f32 destination) */
static inline void o_store4(oimg_t *im, int x, int y, const float *v, int f16)
{
  if(x < 0 || y < 0 || x >= im->w || y >= im->h) return;
  float *d = im->p + ((size_t)y * im->w + x) * im->c;
  for(int k = 0; k < im->c; k++) d[k] = f16 ? o_f16r(v[k]) : v[k];
}
static inline void o_store1(oimg_t *im, int x, int y, float v, int f16)
{
  if(x < 0 || y < 0 || x >= im->w || y >= im->h) return;
  im->p[((size_t)y * im->w + x) * im->c] = f16 ? o_f16r(v) : v;
}

/* ---- glsl builtins ---- */
/* glsl leaves min/max with a NaN operand undefined; GPUs (and llvm's minnum/maxnum) return the non-NaN operand.
 * it matters: evd2x2's sqrt() of a discriminant that rounds below zero yields NaN in flat regions, and the
 * shaders' max(1e-9, exp(NaN)) / clamp(NaN, ..) then fall back to finite values (denoise/cov.glsl:116-133). */
static inline float o_min(float a, float b) { return fminf(a, b); }
static inline float o_max(float a, float b) { return fmaxf(a, b); }
static inline float o_clamp(float x, float a, float b) { return o_min(o_max(x, a), b); }
static inline float o_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float o_smoothstep(float e0, float e1, float x)
{
  const float t = o_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
static inline float o_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
/* shared.glsl:150-154 */
static inline float o_lum2020(const float *rgb)
{
  return 2.62700212e-01f * rgb[0] + 6.77998072e-01f * rgb[1] + 5.93017165e-02f * rgb[2];
}
/* y = M x for row-major 3x3 (the C flavour of matrices.h `makemat`) */
static inline void o_mat3mulv(const float *M, const float *x, float *y)
{
  float t[3];
  for(int j = 0; j < 3; j++) t[j] = M[3*j+0]*x[0] + M[3*j+1]*x[1] + M[3*j+2]*x[2];
  y[0] = t[0]; y[1] = t[1]; y[2] = t[2];
}

/* shared.glsl:244-293, eigen decomposition of the symmetric 2x2 (a b; b c) */
static inline void o_evd2x2(float a, float b, float c, float *eval, float *evec0, float *evec1)
{
  const float pHalf = -0.5f * (a + c);
  const float q = a*c - b*b;
  const float dr = sqrtf(pHalf * pHalf - q);
  eval[0] = -pHalf + dr;
  eval[1] = -pHalf - dr;
  const float a0 = a - eval[0], b0 = b, c0 = c - eval[0];
  const float sl0 = a0*a0 + b0*b0, sl1 = b0*b0 + c0*c0;
  float sl;
  if(sl0 > sl1) { evec1[0] = a0; evec1[1] = b0; sl = sl0; }
  else          { evec1[0] = b0; evec1[1] = c0; sl = sl1; }
  evec1[0] = (sl == 0.0f) ? 1.0f : evec1[0];
  sl = (sl == 0.0f) ? 1.0f : sl;
  const float il = 1.0f / sqrtf(sl);
  evec1[0] *= il; evec1[1] *= il;
  evec0[0] = evec1[1]; evec0[1] = -evec1[0];
}

#ifdef __cplusplus
}
#endif

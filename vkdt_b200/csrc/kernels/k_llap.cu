// local laplacian pyramids, restructured for HBM: the 11 x full-res level-0 stack of the reference
// (llap/curve.comp writes 22 B/px, reduce and assemble re-read it) is never materialised.
//  - (b200, llapr0)  curve + first reduce fused: rgba in -> 11 planes of level 1 (tile staged in shared memory)
//  - (llap, reduce)  11-layer [1 2 1]^2/16 reduce of the coarse levels (reduce.comp:16-35, sample_semisoft)
//  - (llap, assemble) coarse-level collapse (assemble.comp:52-88, sample_soft)
//  - (b200, llapfin) finest assemble + llap/colour.comp (+ grade) fused: level-0 laplacians are recomputed
//    from the input pixel (curve() rounded to f16 in registers = the value the reference would have stored)
// wiring: llap/main.c:31-105.  gamma/curve: llap/llap.glsl:3-22, llap/curve.comp:40-63.
#include <string.h>
#include "pointwise.cuh"

#define NUM_GAMMA 10
#define NL (NUM_GAMMA + 1)

struct llap_params_t { float sigma, shadows, hilights, clarity; };

VKB_DEV float gamma_from_i(int i) { return div_c((float)i, NUM_GAMMA - 1.0f); }
// for an index the compiler knows (unrolled loops): the plain quotient folds to a constant, the intrinsics of div_c do not
VKB_DEV float gamma_from_const_i(int i) { return (float)i / (NUM_GAMMA - 1.0f); }
VKB_DEV int gamma_hi_from_v(float v)
{ // llap.glsl:17-22 without a loop of divisions: 1 + #{ i in 1..8 : i/9 <= v }, the i/9 fold to constants
  int hi = 1;
#pragma unroll
  for(int i = 1; i < NUM_GAMMA - 1; i++) hi += ((float)i / (NUM_GAMMA - 1.0f) <= v) ? 1 : 0;
  return hi;
}
// curve() with its two divisions by per-launch constants turned into multiplications by their reciprocals
// (inv2s = 1/(2 sigma), invd = 1/(2 sigma^2/3)): <= 1 ulp of difference in t and in the exponent, the value is
// rounded to f16 right after.  the level-0 stack costs 10 of these per input pixel, the divisions were a third of it.
// CLARITY = false: p.clarity == 0 (the default).  the gaussian term is then +-0 with the sign of c, and adding
// 0 * c reproduces `val + 0 * c * exp(..)` bit for bit (signed zeros included) without the exponential.
VKB_DEV float llap_curve(float x, float g, const llap_params_t &p)
{ // llap/curve.comp:40-63
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? p.sigma : -p.sigma;
  const float shadhi = c > 0.0f ? p.shadows : p.hilights;
  if(fabsf(c) > 2 * p.sigma) val = g + ssigma + shadhi * (c - ssigma);
  else
  {
    const float t = clampf(c / (2.0f * ssigma), 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  val += p.clarity * c * m_exp(-c * c / (2.0f * p.sigma * p.sigma / 3.0f));
  return val;
}
// strict: the shader's two quotients (divisors 2 ssigma and 2 sigma^2 / 3 are launch constants: div_rd) and libm's exponential
struct llap_rd_t { double rd2s, rdk; };
static llap_rd_t llap_rd_host(const llap_params_t &p)
{ // the divisors as the shader's fp32 expressions form them (volatile: one rounding per operation on the host too)
  const volatile float two_s = 2.0f * p.sigma;
  const volatile float ss = two_s * p.sigma;
  const volatile float k = ss / 3.0f;
  llap_rd_t r = { 1.0 / (double)two_s, 1.0 / (double)k };
  return r;
}
VKB_DEV float llap_curve_x(float x, float g, const llap_params_t &p, const llap_rd_t &R, const lme_ctx_t &L)
{
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? p.sigma : -p.sigma;
  const float shadhi = c > 0.0f ? p.shadows : p.hilights;
  if(fabsf(c) > 2 * p.sigma) val = g + ssigma + shadhi * (c - ssigma);
  else
  {
    const float t = clampf(div_rd(c, c > 0.0f ? R.rd2s : -R.rd2s), 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  val += p.clarity * c * m_exp_s(div_rd(-c * c, R.rdk), L);
  return val;
}
template <bool CLARITY>
VKB_DEV float llap_curve_k(float x, float g, const llap_params_t &p, float inv2s, float invd, const llap_rd_t &R, const lme_ctx_t &L)
{
#if !VKB_FAST
  if(CLARITY) return llap_curve_x(x, g, p, R, L);
#endif
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? p.sigma : -p.sigma;
  const float shadhi = c > 0.0f ? p.shadows : p.hilights;
  if(fabsf(c) > 2 * p.sigma) val = g + ssigma + shadhi * (c - ssigma);
  else
  {
#if VKB_FAST
    const float t = clampf(fabsf(c) * inv2s, 0.0f, 1.0f);
#else
    const float t = clampf(c / (2.0f * ssigma), 0.0f, 1.0f);
#endif
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  if(CLARITY) val += p.clarity * c * exp_ftz(-c * c * invd);
  else        val += 0.0f * c;
  return val;
}
VKB_DEV float llap_grey(float4 px)
{ // curve.comp:72: clamp away nans and infs
  return lum2020(clampf(px.x, -1000.0f, 1000.0f), clampf(px.y, -1000.0f, 1000.0f), clampf(px.z, -1000.0f, 1000.0f));
}

// ---- curve + reduce to level 1 ----
// block = 32x8 outputs, input tile (2*32+1) x (2*8+1) texels around them, 11 f16 values per texel in smem.
#define R0_TW 65
#define R0_TH 17
template <bool CLARITY>
__global__ void __launch_bounds__(256, 6) k_llap_reduce0(const uint2 *__restrict__ in, int iw, int ih,
    __half *__restrict__ out, int ow, int oh, llap_params_t p, const llap_rd_t R, const band_t bd)
{
  __shared__ __align__(16) __half tile[NL][R0_TH][R0_TW + 1];
  const int tx0 = blockIdx.x * 64 - 1, ty0 = BAND_BY * 16 - 1;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const float inv2s = 1.0f / (2.0f * p.sigma), invd = 1.0f / (2.0f * p.sigma * p.sigma / 3.0f);
  LME_SMEM_STAGE(tid);
  __syncthreads();
  for(int t = tid; t < R0_TW * R0_TH; t += 256)
  {
    const int lx = t % R0_TW, ly = t / R0_TW;
    const bool big = iw >= 66 && ih >= 18; // the tile overhangs the image by < one tile: the cheap mirror is enough
    const int gx = big ? mirror1(tx0 + lx, iw) : mirrori(tx0 + lx, iw), gy = big ? mirror1(ty0 + ly, ih) : mirrori(ty0 + ly, ih);
    const float y = llap_grey(ld_rgba(in, iw, gx, gy));
#pragma unroll
    for(int g = 0; g < NUM_GAMMA; g++) tile[g][ly][lx] = __float2half_rn(llap_curve_k<CLARITY>(y, gamma_from_const_i(g), p, inv2s, invd, R, lme_ctx));
    tile[NUM_GAMMA][ly][lx] = __float2half_rn(y);
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  const int lx = 2 * threadIdx.x, ly = 2 * threadIdx.y; // tile coords of texel (2x-1, 2y-1)
  const size_t plane = (size_t)ow * oh;
#pragma unroll
  for(int g = 0; g < NL; g++)
  {
    float t[3][3];
#pragma unroll
    for(int j = 0; j < 3; j++)
#pragma unroll
      for(int i = 0; i < 3; i += 2)
      { // lx is even and the rows are 66 halves long: (lx, lx+1) is one aligned 32-bit load
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&tile[g][ly + j][lx + i]));
        t[j][i] = f.x;
        if(i == 0) t[j][1] = f.y;
      }
    // sample_semisoft: four bilinear taps with weights 1/2, summed, / 4
    const float b00 = (t[0][0] * 0.5f + t[0][1] * 0.5f) * 0.5f + (t[1][0] * 0.5f + t[1][1] * 0.5f) * 0.5f;
    const float b10 = (t[0][1] * 0.5f + t[0][2] * 0.5f) * 0.5f + (t[1][1] * 0.5f + t[1][2] * 0.5f) * 0.5f;
    const float b01 = (t[1][0] * 0.5f + t[1][1] * 0.5f) * 0.5f + (t[2][0] * 0.5f + t[2][1] * 0.5f) * 0.5f;
    const float b11 = (t[1][1] * 0.5f + t[1][2] * 0.5f) * 0.5f + (t[2][1] * 0.5f + t[2][2] * 0.5f) * 0.5f;
    out[g * plane + (size_t)y * ow + x] = __float2half_rn((((b00 + b10) + b01) + b11) / 4.0f);
  }
}

// ---- reduce of coarse levels: one thread per output pixel, all layers: the nine mirrored offsets are formed once and
// serve every plane (the per layer version spent four fifths of its instructions on them) ----
__global__ void __launch_bounds__(256, 4) k_llap_reduce(const __half *__restrict__ in, int iw, int ih,
    __half *__restrict__ out, int ow, int oh, int layers, const band_t bd)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  int xo[3]; size_t yo[3];
#pragma unroll
  for(int i = 0; i < 3; i++) { xo[i] = mirror1(2 * x - 1 + i, iw); yo[i] = (size_t)mirror1(2 * y - 1 + i, ih) * iw; }
  const size_t iplane = (size_t)iw * ih, oplane = (size_t)ow * oh;
  const __half *src = in;
  __half *dst = out + (size_t)y * ow + x;
#pragma unroll 2
  for(int g = 0; g < layers; g++, src += iplane, dst += oplane)
  {
    float t[3][3];
#pragma unroll
    for(int j = 0; j < 3; j++)
#pragma unroll
      for(int i = 0; i < 3; i++) t[j][i] = __half2float(__ldg(src + yo[j] + xo[i]));
    const float b00 = (t[0][0] * 0.5f + t[0][1] * 0.5f) * 0.5f + (t[1][0] * 0.5f + t[1][1] * 0.5f) * 0.5f;
    const float b10 = (t[0][1] * 0.5f + t[0][2] * 0.5f) * 0.5f + (t[1][1] * 0.5f + t[1][2] * 0.5f) * 0.5f;
    const float b01 = (t[1][0] * 0.5f + t[1][1] * 0.5f) * 0.5f + (t[2][0] * 0.5f + t[2][1] * 0.5f) * 0.5f;
    const float b11 = (t[1][1] * 0.5f + t[1][2] * 0.5f) * 0.5f + (t[2][1] * 0.5f + t[2][2] * 0.5f) * 0.5f;
    *dst = __float2half_rn((((b00 + b10) + b01) + b11) / 4.0f);
  }
}

// sample_soft(img, (opos*0.5+0.5)/size): 3x3 bilinear taps at -1.5, 0, +1.5 texels, / 9 (shared.glsl:99-127).
// per axis the taps land on: even o=2k: {k-2|k-1 (1/2), k, k+1|k+2 (1/2)}, odd o=2k+1: {k-1, k|k+1 (1/2), k+2}
struct soft_axis_t { int i0[3], i1[3]; float a[3]; };
VKB_DEV soft_axis_t soft_axis(int o, int n)
{ // texel indices are mirrored here, once per pixel and axis (the integer modulo of a general mirror dominated the kernel)
  soft_axis_t s;
  const int k = o >> 1;
  if(o & 1) { s.i0[0] = k - 1; s.a[0] = 0.0f; s.i0[1] = k; s.a[1] = 0.5f; s.i0[2] = k + 2; s.a[2] = 0.0f; }
  else      { s.i0[0] = k - 2; s.a[0] = 0.5f; s.i0[1] = k; s.a[1] = 0.0f; s.i0[2] = k + 1; s.a[2] = 0.5f; }
#pragma unroll
  for(int t = 0; t < 3; t++)
  {
    const int a = s.i0[t], b = a + 1;
    if(n >= 4) { s.i0[t] = mirror1(a, n); s.i1[t] = mirror1(b, n); }
    else       { s.i0[t] = mirrori(a, n); s.i1[t] = mirrori(b, n); }
  }
  return s;
}
VKB_DEV float gauss_expand(const __half *__restrict__ img, int w, int h, const soft_axis_t &sx, const soft_axis_t &sy)
{ // sx/sy hold mirrored texel indices (soft_axis_mirror), shared by every plane expanded at this pixel
  float r = 0.0f;
#pragma unroll
  for(int j = 0; j < 3; j++)
  {
    const __half *r0 = img + (size_t)sy.i0[j] * w, *r1 = img + (size_t)sy.i1[j] * w;
    const float ay = sy.a[j];
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
      const int x0 = sx.i0[i], x1 = sx.i1[i];
      const float ax = sx.a[i];
      float top = __half2float(__ldg(r0 + x0)) * (1.0f - ax);
      if(ax != 0.0f) top += __half2float(__ldg(r0 + x1)) * ax;
      float v = top * (1.0f - ay);
      if(ay != 0.0f)
      {
        float bot = __half2float(__ldg(r1 + x0)) * (1.0f - ax);
        if(ax != 0.0f) bot += __half2float(__ldg(r1 + x1)) * ax;
        v += bot * ay;
      }
      r += v;
    }
  }
  return div9(r);
}

// ---- assemble for coarse levels (both l0 and l1 stacks are in memory) ----
__global__ void __launch_bounds__(256) k_llap_assemble(const __half *__restrict__ coarse, const __half *__restrict__ l0,
    const __half *__restrict__ l1, int cw, int ch, __half *__restrict__ out, int ow, int oh, int first)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  const size_t p0 = (size_t)ow * oh, p1 = (size_t)cw * ch;
  const soft_axis_t sx = soft_axis(x, cw), sy = soft_axis(y, ch);
  const float res = first ? gauss_expand(l1 + NUM_GAMMA * p1, cw, ch, sx, sy) : gauss_expand(coarse, cw, ch, sx, sy);
  const float v = ld_h(l0 + NUM_GAMMA * p0, ow, x, y);
  const int hi = gamma_hi_from_v(v), lo = hi - 1;
  const float glo = gamma_from_i(lo), ghi = gamma_from_i(hi);
  const float a = clampf((v - glo) / (ghi - glo), 0.0f, 1.0f);
  const float lap0 = ld_h(l0 + lo * p0, ow, x, y) - gauss_expand(l1 + lo * p1, cw, ch, sx, sy);
  const float lap1 = ld_h(l0 + hi * p0, ow, x, y) - gauss_expand(l1 + hi * p1, cw, ch, sx, sy);
  // explicit _rn ops: this blend feeds the next pyramid level, keep it unfused whatever the compiler flags say
  out[(size_t)y * ow + x] = __float2half_rn(__fadd_rn(__fadd_rn(res, __fmul_rn(lap0, 1.0f - a)), __fmul_rn(lap1, a)));
}

// ---- shared-memory tiled variant of the coarse assemble: same arithmetic (shader tap order), but the 20x8 coarse
// window of all 12 planes is staged once per CTA of 32x8 outputs instead of ~60 mirrored global loads per pixel ----
#define AT_W 20
#define AT_H 8
struct soft_local_t { int i0[3]; float a[3]; };
VKB_DEV soft_local_t soft_local(int o, int origin)
{ // tile-local first column/row of each of the three taps
  soft_local_t s;
  const int k = (o >> 1) - origin;
  if(o & 1) { s.i0[0] = k - 1; s.a[0] = 0.0f; s.i0[1] = k; s.a[1] = 0.5f; s.i0[2] = k + 2; s.a[2] = 0.0f; }
  else      { s.i0[0] = k - 2; s.a[0] = 0.5f; s.i0[1] = k; s.a[1] = 0.0f; s.i0[2] = k + 1; s.a[2] = 0.5f; }
  return s;
}
VKB_DEV float gauss_expand_tile(const float (*T)[AT_W + 1], const soft_local_t &sx, const soft_local_t &sy)
{
  float r = 0.0f;
#pragma unroll
  for(int j = 0; j < 3; j++)
  {
    const float ay = sy.a[j];
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
      const float ax = sx.a[i];
      float top = T[sy.i0[j]][sx.i0[i]] * (1.0f - ax);
      if(ax != 0.0f) top += T[sy.i0[j]][sx.i0[i] + 1] * ax;
      float v = top * (1.0f - ay);
      if(ay != 0.0f)
      {
        float bot = T[sy.i0[j] + 1][sx.i0[i]] * (1.0f - ax);
        if(ax != 0.0f) bot += T[sy.i0[j] + 1][sx.i0[i] + 1] * ax;
        v += bot * ay;
      }
      r += v;
    }
  }
  return div9(r);
}
__global__ void __launch_bounds__(256) k_llap_assemble_tiled(const __half *__restrict__ coarse, const __half *__restrict__ l0,
    const __half *__restrict__ l1, int cw, int ch, __half *__restrict__ out, int ow, int oh, int first)
{
  __shared__ float tile[NL + 1][AT_H][AT_W + 1];
  __shared__ int s_pmin, s_pmax;
  const int cx0 = blockIdx.x * 16 - 2, cy0 = blockIdx.y * 4 - 2;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const size_t p0 = (size_t)ow * oh, p1 = (size_t)cw * ch;
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  const bool inside = x < ow && y < oh;
  // which gamma layers does this CTA need?  only those (plus the collapsed coarse level) are staged
  if(tid == 0) { s_pmin = NUM_GAMMA; s_pmax = 0; }
  float v = 0.0f; int hi = 1;
  if(inside) { v = ld_h(l0 + NUM_GAMMA * p0, ow, x, y); hi = gamma_hi_from_v(v); }
  __syncthreads();
  const int wlo = __reduce_min_sync(0xffffffffu, inside ? hi - 1 : NUM_GAMMA), whi = __reduce_max_sync(0xffffffffu, inside ? hi : 0);
  if(threadIdx.x == 0) { atomicMin(&s_pmin, wlo); atomicMax(&s_pmax, whi); }
  __syncthreads();
  const int pmin = s_pmin, pmax = s_pmax;
  const bool big = cw >= 24 && ch >= 12;
  if(tid < AT_H * AT_W)
  {
    const int r = tid / AT_W, c = tid - r * AT_W;
    const int gx = big ? mirror1(cx0 + c, cw) : mirrori(cx0 + c, cw), gy = big ? mirror1(cy0 + r, ch) : mirrori(cy0 + r, ch);
    const size_t off = (size_t)gy * cw + gx;
    tile[NL][r][c] = __half2float(__ldg((first ? l1 + NUM_GAMMA * p1 : coarse) + off));
    for(int pl = pmin; pl <= pmax; pl++) tile[pl][r][c] = __half2float(__ldg(l1 + pl * p1 + off));
  }
  __syncthreads();
  if(!inside) return;
  const soft_local_t sx = soft_local(x, cx0), sy = soft_local(y, cy0);
  const float res = gauss_expand_tile(tile[NL], sx, sy);
  const int lo = hi - 1;
  const float glo = gamma_from_i(lo), ghi = gamma_from_i(hi);
  const float a = clampf((v - glo) / (ghi - glo), 0.0f, 1.0f);
  const float lap0 = ld_h(l0 + lo * p0, ow, x, y) - gauss_expand_tile(tile[lo], sx, sy);
  const float lap1 = ld_h(l0 + hi * p0, ow, x, y) - gauss_expand_tile(tile[hi], sx, sy);
  // explicit _rn ops: this blend feeds the next pyramid level, keep it unfused whatever the compiler flags say
  out[(size_t)y * ow + x] = __float2half_rn(__fadd_rn(__fadd_rn(res, __fmul_rn(lap0, 1.0f - a)), __fmul_rn(lap1, a)));
}

// ---- finest assemble + recolouring (+ grade) ----
struct llapfin_t { llap_params_t p; int first; int have_grade; int out_f32; grade_params_t grade; };

__global__ void __launch_bounds__(256) k_llap_final(const uint2 *__restrict__ in, const __half *__restrict__ coarse,
    const __half *__restrict__ l1, int cw, int ch, void *__restrict__ outv, int ow, int oh, const __grid_constant__ llapfin_t P)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  const size_t p1 = (size_t)cw * ch;
  const float4 px = ld_rgba(in, ow, x, y);
  const float grey = llap_grey(px);
  const soft_axis_t sx = soft_axis(x, cw), sy = soft_axis(y, ch);
  const float res = P.first ? gauss_expand(l1 + NUM_GAMMA * p1, cw, ch, sx, sy) : gauss_expand(coarse, cw, ch, sx, sy);
  const float v = f16r(grey);
  const int hi = gamma_hi_from_v(v), lo = hi - 1;
  const float glo = gamma_from_i(lo), ghi = gamma_from_i(hi);
  const float a = clampf((v - glo) / (ghi - glo), 0.0f, 1.0f);
  const float lap0 = f16r(llap_curve(grey, glo, P.p)) - gauss_expand(l1 + lo * p1, cw, ch, sx, sy);
  const float lap1 = f16r(llap_curve(grey, ghi, P.p)) - gauss_expand(l1 + hi * p1, cw, ch, sx, sy);
  float l = f16r(res + lap0 * (1.0f - a) + lap1 * a);
  // llap/colour.comp:17-37
  const float yo = fmaxf(lum2020(px.x, px.y, px.z), 1e-8f);
  if(l < yo) l = yo * m_exp(1.0f * (l - yo));
  f3 c = { fmaxf(0.0f, px.x * l / yo), fmaxf(0.0f, px.y * l / yo), fmaxf(0.0f, px.z * l / yo) };
  if(P.have_grade)
  {
    c = { f16r(c.x), f16r(c.y), f16r(c.z) };
    c = grade_px(c, P.grade);
  }
  if(P.out_f32) reinterpret_cast<float4 *>(outv)[(size_t)y * ow + x] = make_float4(c.x, c.y, c.z, 1.0f);
  else st_rgba(reinterpret_cast<uint2 *>(outv), ow, x, y, make_float4(c.x, c.y, c.z, 1.0f));
}

static inline dim3 grid2d(unsigned w, unsigned h, unsigned z = 1) { return dim3(vkb_cdiv(w, 32), vkb_cdiv(h, 8), z); }
static const dim3 blk2d(32, 8);

int launch_llapr0_packed(const vkb_launch_t *l);
// conn: [0] input rgba f16 (level 0), [1] output y f16 x 11 layers (level 1).  params: llap params
static int launch_llapr0(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->params_size >= sizeof(llap_params_t));
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16 && out->chan == 1 && out->layers == NL && out->format == VKB_TOKEN_F16);
  VKB_REQUIRE(out->wd == (in->wd - 1) / 2 + 1 && out->ht == (in->ht - 1) / 2 + 1);
  const llap_params_t *lp = (const llap_params_t *)l->params;
  // fast: two layers per packed fp32 instruction (k_llap_r0.cu: fused multiply-adds, reciprocals, the SFU exponential);
  // strict: the scalar kernel below with the shader's curve operation for operation
  if(VKB_FAST && !getenv("VKB_LLAPR0_SCALAR")) return launch_llapr0_packed(l);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  if(lp->clarity == 0.0f)
    k_llap_reduce0<false><<<grid, blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht,
        (__half *)out->data, out->wd, out->ht, *lp, llap_rd_host(*lp), bd);
  else
    k_llap_reduce0<true><<<grid, blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht,
        (__half *)out->data, out->wd, out->ht, *lp, llap_rd_host(*lp), bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("b200", "llapr0", launch_llapr0);

// conn: [0] inhi y f16 x 11, [1] outlo y f16 x 11
static int launch_llap_reduce(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 1 && out->chan == 1 && in->layers == out->layers && in->format == VKB_TOKEN_F16 && out->format == VKB_TOKEN_F16);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_llap_reduce<<<grid, blk2d, 0, l->stream>>>((const __half *)in->data, in->wd, in->ht,
      (__half *)out->data, out->wd, out->ht, (int)out->layers, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("llap", "reduce", launch_llap_reduce);

// conn: [0] coarse y f16 (ignored when push.first), [1] currlo x11 (fine), [2] currhi x11 (coarse), [3] fine out y f16
// push: { u32 num_gamma; u32 first } (llap/main.c:66,92)
int launch_llap_assemble4(const vkb_launch_t *l, int first);
static int launch_llap_assemble(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 4 && l->push_size >= 8);
  const uint32_t *pc = (const uint32_t *)l->push;
  const vkb_image_t *coarse = l->conn, *l0 = l->conn + 1, *l1 = l->conn + 2, *out = l->conn + 3;
  VKB_REQUIRE(pc[0] == NUM_GAMMA && l0->layers == NL && l1->layers == NL && out->chan == 1);
  VKB_REQUIRE(l0->wd == out->wd && l0->ht == out->ht);
  VKB_REQUIRE(pc[1] || (coarse->wd == l1->wd && coarse->ht == l1->ht));
  VKB_REQUIRE(l1->wd == (out->wd - 1) / 2 + 1 && l1->ht == (out->ht - 1) / 2 + 1);
  if(out->wd >= 80 && out->ht >= 32) return launch_llap_assemble4(l, (int)pc[1]); // one thread per 2x2 pixels, k_llap_asm4.cu
  if(l->band_y0 >= 0 && !(l->band_y0 == 0 && l->band_y1 >= (int)out->ht)) return vkb_set_error(VKB_ERR_BAD_ARG, "llap assemble: levels this small run whole, not banded");
  if(out->wd >= 64 && out->ht >= 16)
    k_llap_assemble_tiled<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const __half *)coarse->data, (const __half *)l0->data,
        (const __half *)l1->data, l1->wd, l1->ht, (__half *)out->data, out->wd, out->ht, (int)pc[1]);
  else
  k_llap_assemble<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const __half *)coarse->data, (const __half *)l0->data,
      (const __half *)l1->data, l1->wd, l1->ht, (__half *)out->data, out->wd, out->ht, (int)pc[1]);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("llap", "assemble", launch_llap_assemble);

// conn: [0] input rgba f16, [1] coarse y f16 (level 1 assembled; ignored when first), [2] level-1 stack x11, [3] output rgba f16|f32
// push: { u32 first; u32 have_grade }.  params: llap params (16 B) followed by grade params (76 B) when have_grade
static int launch_llapfin(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 4 && l->push_size >= 8 && l->params_size >= sizeof(llap_params_t));
  const uint32_t *pc = (const uint32_t *)l->push;
  const vkb_image_t *in = l->conn, *coarse = l->conn + 1, *l1 = l->conn + 2, *out = l->conn + 3;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16 && l1->layers == NL && out->chan == 4);
  VKB_REQUIRE(in->wd == out->wd && in->ht == out->ht);
  VKB_REQUIRE(out->format == VKB_TOKEN_F16 || out->format == VKB_TOKEN_F32);
  llapfin_t P;
  memset(&P, 0, sizeof(P));
  memcpy(&P.p, l->params, sizeof(llap_params_t));
  P.first = pc[0]; P.have_grade = pc[1]; P.out_f32 = out->format == VKB_TOKEN_F32;
  if(P.have_grade)
  {
    VKB_REQUIRE(l->params_size >= sizeof(llap_params_t) + sizeof(grade_params_t));
    memcpy(&P.grade, (const uint8_t *)l->params + sizeof(llap_params_t), sizeof(grade_params_t));
  }
  k_llap_final<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const uint2 *)in->data, (const __half *)coarse->data,
      (const __half *)l1->data, l1->wd, l1->ht, out->data, out->wd, out->ht, P);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
// the shader-order (9 tap) variant stays available for A/B checks; the executor uses k_llap_fin.cu's (b200, llapfin)
VKB_REGISTER("b200", "llapfinx", launch_llapfin);

VKB_NS_END

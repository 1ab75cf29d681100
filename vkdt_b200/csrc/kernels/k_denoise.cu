// noise-profile driven wavelet denoise on the mosaic (denoise:strength > 0).
// replaces src/pipe/modules/denoise/{half,downcov,down,assemble,doub}.comp, cov.glsl, noise.glsl
// (wired by denoise/main.c:227-326): half-size rgb -> structure-tensor guided blur (level 0) -> three edge-aware
// "flower" levels written in the quadrant-swizzled layout that turns the same taps into an a-trous wavelet ->
// soft-shrunk reassembly -> per-CFA-colour residual shrink on the full-resolution mosaic.
// all intermediates live at 1/block^2 resolution; the full-res `doub` is the only HBM-heavy kernel (2+2 B/px + coarse reads).
// compiled with --fmad=false: downcov picks between two covariance estimates by comparing determinants and rejects
// hot pixels by a threshold — discontinuous decisions that must fall like in the fp32 restatement.
#include <string.h>
#include "common.cuh"

struct denoise_params_t { float strength, luma, detail, pad; float edges[4]; int gainmap; };      // denoise/params
struct dn_push_half_t { float wb[4], black[4], white[4]; int32_t crop[4]; uint32_t filters; };     // main.c:284-289
struct dn_push_down_t { float wb[4], black[4], white[4]; int32_t crop[4]; float noise_a, noise_b; int32_t level; uint32_t block; }; // :244-253
struct dn_push_asm_t  { float wb[4], black[4], white[4]; int32_t crop[4]; float noise_a, noise_b; uint32_t filters; };             // :266-274
struct dn_push_doub_t { float wb[4], black[4], white[4]; int32_t crop[4]; uint32_t filters; float noise_a, noise_b; int32_t gainmap; float map_os[4]; }; // :301-308

// noise.glsl:1-11
// `escale[k]` = exp2(12*edges[k] + edges[3]) is a function of the module params only: evaluated once per launch on the host
VKB_DEV void noise_sigma(float a, float b, float black, float white, const float *escale, float val, float *sig)
{
#if VKB_FAST   // call free forms (white > black, a, b >= 0: a noise profile)
  const float s = sqrt_f(a + fmaxf(0.0f, div_f(val - black, white - black)) * b);
#else
  const float s = sqrtf(a + fmaxf(0.0f, (val - black) / (white - black)) * b);
#endif
#pragma unroll
  for(int k = 0; k < 3; k++) sig[k] = clampf(escale[k] * s, 1e-3f, 1e3f);
}
static void host_escale(denoise_params_t *p)
{ // overwrite edges[0..2] in the kernel's copy of the params with the scale factors
  for(int k = 0; k < 3; k++) p->edges[k] = exp2f(12.0f * p->edges[k] + p->edges[3]);
}
VKB_DEV void swizzle(int x, int y, int w, int h, int &ox, int &oy)
{ // downcov.comp:52-53, down.comp:103-104
  ox = x / 2 + ((x & 1) * (w + 1)) / 2;
  oy = y / 2 + ((y & 1) * (h + 1)) / 2;
}

// ---- half: cfa block -> rgb (half.comp:24-74); input is the ui16 source sampled as UNORM ----
__global__ void __launch_bounds__(256) k_denoise_half(const uint16_t *__restrict__ in, int iw, int ih,
    uint2 *__restrict__ out, int ow, int oh, int cx, int cy, float white, int xtrans, const band_t bd)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  float4 rgba;
#define RAW(X, Y) div_c((float)__ldg(in + (size_t)clampi((Y), 0, ih - 1) * iw + clampi((X), 0, iw - 1)), 65535.0f)
  if(xtrans)
  {
    float c[9];
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
      for(int j = 0; j < 3; j++) c[3 * i + j] = RAW(cx + 3 * x + i, cy + 3 * y + j);
    const float col0 = (c[1] + c[7]) * 0.5f, col1 = (c[3] + c[5]) * 0.5f;
    if(((x + y + cx + cy) & 1) > 0) { rgba.x = col0; rgba.z = col1; }
    else                            { rgba.z = col0; rgba.x = col1; }
    rgba.y = div_c((c[0] + c[2] + c[4] + c[6] + c[8]) * 1.0f, 5.0f);
    rgba.w = 1.0f;
  }
  else
  { // textureGather at the block centre: x=(0,1) y=(1,1) z=(1,0) w=(0,0), mirrored repeat
    const int x0 = mirrori(cx + 2 * x, iw), x1 = mirrori(cx + 2 * x + 1, iw), y0 = mirrori(cy + 2 * y, ih), y1 = mirrori(cy + 2 * y + 1, ih);
    float gx = div_c((float)__ldg(in + (size_t)y1 * iw + x0), 65535.0f), gy = div_c((float)__ldg(in + (size_t)y1 * iw + x1), 65535.0f);
    float gz = div_c((float)__ldg(in + (size_t)y0 * iw + x1), 65535.0f), gw = div_c((float)__ldg(in + (size_t)y0 * iw + x0), 65535.0f);
    if(gx >= white) gx = gz;
    if(gz >= white) gz = gx;
    rgba = make_float4(gw, (gx + gz) / 2.0f, gy, 1.0f);
  }
#undef RAW
  st_rgba(out, ow, x, y, rgba);
}

// ---- downcov: level 0, structure tensor guided blur (downcov.comp:41-62, cov.glsl:21-138) ----
// a CTA of 32x8 outputs stages its 36x12 input window once in shared memory as (r, g, b, luminance) floats: the three
// 5x5 passes of response() then read LDS.128 instead of 75 mirrored 8-byte global loads + conversions + luminance
// dot products per pixel.  arithmetic per pixel is unchanged (same order, IEEE divisions) so the covariance choice and
// the hot pixel test fall exactly like in the restatement.
#define DC_W 36
#define DC_H 12
// the j loops stay rolled: 55 registers instead of 128, four CTAs per SM; the unrolled i loop gives the ILP
__global__ void __launch_bounds__(256, 5) k_denoise_downcov(const uint2 *__restrict__ in, int w, int h,
    uint2 *__restrict__ out, uint2 *__restrict__ covimg, const band_t bd)
{
  __shared__ float4 tile[DC_H][DC_W];  // r g b lum/25
  __shared__ float4 tinv[DC_H][DC_W];  // lum, 1/lum | lum*lum, 1/(lum*lum): every tap's divisions, done once per input texel
  const int tx0 = blockIdx.x * 32 - 2, ty0 = BAND_BY * 8 - 2;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  LME_SMEM_STAGE(tid);
  for(int t = tid; t < DC_W * DC_H; t += 256)
  {
    const int r = t / DC_W, c = t - r * DC_W;
    const float4 v = ld_rgba(in, w, mirrori(tx0 + c, w), mirrori(ty0 + r, h));
    const float l = lum2020(v.x, v.y, v.z), l2 = l * l;
    tile[r][c] = make_float4(v.x, v.y, v.z, div_c(l, 25.0f));
    tinv[r][c] = make_float4(l, div_g(1.0f, l), l2, div_g(1.0f, l2));
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= w || y >= h || BAND_SKIP(y)) return;
  const int lx = threadIdx.x, ly = threadIdx.y; // tile coords of tap (-2,-2)
  // the "white" (weights lum, lum^2) and "black" (weights 1/lum, 1/lum^2) estimates of cov.glsl run through identical
  // arithmetic: they travel as the two lanes of packed fp32 pairs (FMUL2 / FFMA2), one instruction for both.
  // lane lo = white, lane hi = black.
  f2 MX = pk2(0.0f, 0.0f), MY = MX, SM = MX;
#pragma unroll 1
  for(int j = 0; j < 5; j++)
  {
    const float fj = (float)(j - 2);
    const f2 FJ = pk2(fj, fj);
#pragma unroll
    for(int i = 0; i < 5; i++)
    {
      const float fi = (float)(i - 2);
      // (lum, 1/lum).  fi / px == fi * (1 / px) bit for bit when fi is 0, +-1 or +-2: scaling a correctly rounded quotient by a power of two is exact
      const f2 L = *reinterpret_cast<const f2 *>(&tinv[ly + j][lx + i].x);
      MX = add2(MX, mul2(pk2(fi, fi), L));
      MY = add2(MY, mul2(FJ, L));
      SM = add2(SM, L);
    }
  }
  const float smw = lo2(SM), smb = hi2(SM);
  const float mwx = div_g(lo2(MX), smw), mwy = div_g(lo2(MY), smw), mbx = div_g(hi2(MX), smb), mby = div_g(hi2(MY), smb);
  // the products run packed; the sums of products stay scalar: ptxas contracts a packed multiply into a following packed
  // add (FFMA2) whatever --fmad says, and a fused sum would not round like the restatement's
  f2 SS = pk2(0.0f, 0.0f);
  float Sw0 = 0, Sw1 = 0, Sw2 = 0, Sw3 = 0, Sb0 = 0, Sb1 = 0, Sb2 = 0, Sb3 = 0, mean_b = 0;
  f2 P0[5];
#pragma unroll
  for(int i = 0; i < 5; i++) P0[i] = pk2((float)(i - 2) - mwx, (float)(i - 2) - mbx);
#pragma unroll 1
  for(int j = 0; j < 5; j++)
  {
    const f2 P1 = pk2((float)(j - 2) - mwy, (float)(j - 2) - mby);
#pragma unroll
    for(int i = 0; i < 5; i++)
    {
      mean_b += tile[ly + j][lx + i].w;
      const f2 Q = *reinterpret_cast<const f2 *>(&tinv[ly + j][lx + i].z); // (lum^2, 1/lum^2)
      const f2 T0 = mul2(Q, P0[i]), T1 = mul2(Q, P1);
      const f2 A = mul2(T0, P0[i]), B = mul2(T0, P1), C = mul2(T1, P0[i]), D = mul2(T1, P1);
      Sw0 += lo2(A); Sw1 += lo2(B); Sw2 += lo2(C); Sw3 += lo2(D);
      Sb0 += hi2(A); Sb1 += hi2(B); Sb2 += hi2(C); Sb3 += hi2(D);
      SS = add2(SS, Q);
    }
  }
  const float sw = lo2(SS), sb = hi2(SS);
  Sw0 = div_g(Sw0, sw); Sw1 = div_g(Sw1, sw); Sw2 = div_g(Sw2, sw); Sw3 = div_g(Sw3, sw);
  Sb0 = div_g(Sb0, sb); Sb1 = div_g(Sb1, sb); Sb2 = div_g(Sb2, sb); Sb3 = div_g(Sb3, sb);
  const bool usew = (Sw0 * Sw3 - Sw1 * Sw2) < (Sb0 * Sb3 - Sb1 * Sb2);
  float e0, e1, v0x, v0y, v1x, v1y;
  evd2x2(usew ? Sw0 : Sb0, usew ? Sw2 : Sb2, usew ? Sw3 : Sb3, e0, e1, v0x, v0y, v1x, v1y);
  e1 *= 0.05f;
  e0 = clampf(e0, 0.01f, 25.0f); e1 = clampf(e1, 0.01f, 25.0f);
  st_rgba(covimg, w, x, y, make_float4(e0, e1, v0x, v0y));
  float r = 0, g = 0, b = 0, wt = 0;
#if VKB_FAST
  // weight(i,j) = exp(-(q0^2/e0 + q1^2/e1)/2) with q = V^t (i,j) is exp2 of a quadratic form in (i,j): the three
  // coefficients are set up once, a tap costs two adds and one ex2.  (continuous in the inputs: ~1e-6 relative on the
  // weights, the blurred colour is rounded to f16 right after.)
  const float ie0 = 1.0f / e0, ie1 = 1.0f / e1, k2 = -0.5f * 1.4426950408889634f;
  const float qa = k2 * (v0x * v0x * ie0 + v1x * v1x * ie1), qc = k2 * (v0y * v0y * ie0 + v1y * v1y * ie1);
  const float qb = 2.0f * k2 * (v0x * v0y * ie0 + v1x * v1y * ie1);
  float ei[5], eb[5];
#pragma unroll
  for(int i = 0; i < 5; i++) { ei[i] = qa * (float)((i - 2) * (i - 2)); eb[i] = qb * (float)(i - 2); }
#pragma unroll 1
  for(int j = 0; j < 5; j++)
  {
    const float fj = (float)(j - 2), ej = qc * fj * fj;
#pragma unroll
    for(int i = 0; i < 5; i++)
    {
      const float4 t = tile[ly + j][lx + i];
      const float wgt = t.x > 2.0f * mean_b ? 0.0f : fmaxf(1e-9f, ex2_ftz((ei[i] + ej) + eb[i] * fj)); // hot pixels get no weight
      r = __fmaf_rn(wgt, t.x, r); g = __fmaf_rn(wgt, t.y, g); b = __fmaf_rn(wgt, t.z, b);
      wt += wgt;
    }
  }
#else
  // cov.glsl:116-133: rotate the offset into the eigenbasis, libm's exponential, unfused accumulation in the shader's tap
  // order.  the weight of tap (i, j) is that of tap (-i, -j) bit for bit (every operation on the way is odd or even in the
  // offset, nan included), so 13 exponentials serve the 25 taps; the two divisors are the pixel's: div_rd
  const double rde0 = rcp_dn(e0), rde1 = rcp_dn(e1);   // both clamped to [0.01, 25]
  float wq[13];
#pragma unroll
  for(int k = 0; k < 13; k++)
  {
    const float fi = (float)(k % 5 - 2), fj = (float)(k / 5 - 2);
    const float x0 = fi * v0x + fj * v0y;
    const float x1 = fi * v1x + fj * v1y;
    wq[k] = fmaxf(1e-9f, m_exp_s(-0.5f * (div_rd(x0, rde0) * x0 + div_rd(x1, rde1) * x1), lme_ctx));
  }
#pragma unroll
  for(int k = 0; k < 25; k++)
  {
    const float4 t = tile[ly + k / 5][lx + k % 5];
    const float wgt = wq[k < 13 ? k : 24 - k];
    if(!(t.x > 2.0f * mean_b)) // hot pixels get no weight
    {
      r += wgt * t.x; g += wgt * t.y; b += wgt * t.z;
      wt += wgt;
    }
  }
#endif
  const float iw_ = fmaxf(wt, 1e-8f);
  float edge = clampf(75.0f * fmaxf(0.0f, e1 - 0.09f), 0.0f, 1.0f);
  edge = smoothstepf(0.4f, 0.75f, edge);
  edge = clampf(0.02f + edge, 0.0f, 1.0f);
  int ox, oy; swizzle(x, y, w, h, ox, oy);
#if VKB_FAST
  st_rgba(out, w, ox, oy, make_float4(r / iw_, g / iw_, b / iw_, edge));
#else
  const double riw = rcp_dn(iw_);   // >= 1e-8
  st_rgba(out, w, ox, oy, make_float4(div_rd(r, riw), div_rd(g, riw), div_rd(b, riw), edge));
#endif
}

#define DN_F02 0.20000004768371582f
#define DN_F04 0.3999999761581421f
// x^0.8 on the SFU (fast): ~1e-6 relative, i.e. <= 1e-4 in the [0,1] edge weight even at the 1e-3 noise floor of noise.glsl
VKB_DEV float gamma08(float f) { return f < 0.0f ? f : m_pow(f, 0.8f); }
// (the branch stays: below black every other noisy value is negative, and a select would pay for the power all the same)
// f >= 0 here is a non negative combination of f16 texels: +0 or far above the subnormal range (m_pow_nn: no call)
VKB_DEV float gamma08(float f, const lme_ctx_t &L) { return f < 0.0f ? f : m_pow_nn(f, 0.8f, L); }

// ---- down: levels 1..3, 5 tap flower with edge stopping (down.comp:59-107) ----
__global__ void __launch_bounds__(256) k_denoise_down(const uint2 *__restrict__ in, int w, int h, uint2 *__restrict__ out,
    denoise_params_t p, float black, float white, float noise_a, float noise_b, float lv, float blk)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= w || y >= h) return;
  const float t = 0.2f;
  const float4 c0 = ld_rgba(in, w, x, y);
  float sigma[3], sum[3], wgt[3], wc[3], g0[3];
  noise_sigma(noise_a, noise_b, black, white, p.edges, c0.x, sigma);
  const float cc[3] = { c0.x, c0.y, c0.z };
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    sum[k] = t * cc[k]; wgt[k] = t;
    sigma[k] = lv * sigma[k] / blk;
    wc[k] = 1.0f / sigma[k];
    g0[k] = gamma08(cc[k]);
  }
  // taps at (+1.2,+0.4) (-1.2,-0.4) (+0.4,-1.2) (-0.4,+1.2) from the texel centre: base texel and bilinear fraction
  const int   bx[4] = { x + 1, x - 2, x,     x - 1 }, by[4] = { y,    y - 1, y - 2, y + 1 };
  // the shader's offsets are the fp32 constants 0.5+-1.2 = 1.7f, -0.7f and 0.5+-0.4 = 0.9f, 0.1f (glslang folds in double, then
  // rounds): an ideal sampler sees the fractions (1.7f - 1.5) = 0.20000005 and (0.9f - 0.5) = 0.39999998, not 0.2f and 0.4f
  const float ax[4] = { DN_F02, 0.8f, DN_F04, 0.6f }, ay[4] = { DN_F04, 0.6f, 0.8f, DN_F02 };
#pragma unroll
  for(int o = 0; o < 4; o++)
  {
    const float4 col = bilin_rgba(in, w, h, bx[o], by[o], ax[o], ay[o]);
    const float c[3] = { col.x, col.y, col.z };
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      const float e = clampf(1.0f - 0.5f * (wc[k] * fabsf(gamma08(c[k]) - g0[k])), 0.0f, 1.0f);
      const float ww = e * (1.0f - t) / 4.0f;
      sum[k] += ww * c[k];
      wgt[k] += ww;
    }
  }
  int ox, oy; swizzle(x, y, w, h, ox, oy);
  st_rgba(out, w, ox, oy, make_float4(sum[0] / wgt[0], sum[1] / wgt[1], sum[2] / wgt[2], 1.0f));
}

// same arithmetic, taps served from a CTA wide window of the packed f16 texels in shared memory: the 17 texels a pixel
// reads sit at constant offsets from one base address, which removes the per tap mirroring and 64-bit address
// arithmetic (a fifth of the instructions of the kernel above).  used when the image is larger than one window.
#define DD_W 36
#define DD_H 12
VKB_DEV float4 unpack_rgba(uint2 v)
{
  const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__global__ void __launch_bounds__(256, 6) k_denoise_down_tiled(const uint2 *__restrict__ in, int w, int h, uint2 *__restrict__ out,
    denoise_params_t p, float black, float white, float noise_a, float noise_b, float lv, float blk, double rd_wb, double rd_blk, const band_t bd)
{
  __shared__ __align__(128) uint2 tile[DD_H][DD_W];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  LME_SMEM_STAGE(tid);
  const int tx0 = blockIdx.x * 32 - 2, ty0 = BAND_BY * 8 - 2;
  // a window inside the image: twelve bulk copies of one 288 byte row segment each (rows start on 16 bytes when w is even)
  const bool interior = tx0 >= 0 && ty0 >= 0 && tx0 + DD_W <= w && ty0 + DD_H <= h && !(w & 1);
  if(interior)
  {
    if(tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if(tid == 0) mbar_expect_tx(&bar, DD_H * DD_W * (uint32_t)sizeof(uint2));
    if(tid < DD_H) bulk_g2s(tile[tid], in + (size_t)(ty0 + tid) * w + tx0, DD_W * (uint32_t)sizeof(uint2), &bar);
    mbar_wait(&bar, 0);
  }
  else
  // thread (tx, ty) stages rows ty, ty + 8 and columns tx, tx + 32 of the window: no division, one mirror per row / column
  {
    const int c0 = threadIdx.x, c1 = threadIdx.x + 32, r0 = threadIdx.y, r1 = threadIdx.y + 8;
    const int gx0 = mirror1(tx0 + c0, w), gx1 = c1 < DD_W ? mirror1(tx0 + c1, w) : 0;
    const size_t gy0 = (size_t)mirror1(ty0 + r0, h) * w, gy1 = r1 < DD_H ? (size_t)mirror1(ty0 + r1, h) * w : 0;
    tile[r0][c0] = __ldg(in + gy0 + gx0);
    if(c1 < DD_W) tile[r0][c1] = __ldg(in + gy0 + gx1);
    if(r1 < DD_H)
    {
      tile[r1][c0] = __ldg(in + gy1 + gx0);
      if(c1 < DD_W) tile[r1][c1] = __ldg(in + gy1 + gx1);
    }
    __syncthreads();
  }
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= w || y >= h || BAND_SKIP(y)) return;
  const int lx = threadIdx.x + 2, ly = threadIdx.y + 2;
  const float t = 0.2f;
  const float4 c0 = unpack_rgba(tile[ly][lx]);
  float sigma[3], sum[3], wgt[3], wc[3], g0[3];
#if VKB_FAST
  noise_sigma(noise_a, noise_b, black, white, p.edges, c0.x, sigma);
#else
  { // noise_sigma() with its quotient by the launch constant (white - black) through div_rd
    const float s = sqrt_f(noise_a + fmaxf(0.0f, div_rd(c0.x - black, rd_wb)) * noise_b);   // noise_a, noise_b > 0
#pragma unroll
    for(int k = 0; k < 3; k++) sigma[k] = clampf(p.edges[k] * s, 1e-3f, 1e3f);
  }
#endif
  const float cc[3] = { c0.x, c0.y, c0.z };
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    sum[k] = t * cc[k]; wgt[k] = t;
#if VKB_FAST
    sigma[k] = lv * sigma[k] / blk;
#else
    sigma[k] = div_rd(lv * sigma[k], rd_blk);
#endif
    wc[k] = div_f(1.0f, sigma[k]);   // clamped to [1e-3, 1e3] * lv / blk
    g0[k] = gamma08(cc[k], lme_ctx);
  }
  constexpr int   bx[4] = { 1, -2, 0, -1 }, by[4] = { 0, -1, -2, 1 };
  constexpr float ax[4] = { DN_F02, 0.8f, DN_F04, 0.6f }, ay[4] = { DN_F04, 0.6f, 0.8f, DN_F02 };
#pragma unroll
  for(int o = 0; o < 4; o++)
  {
    const float4 t00 = unpack_rgba(tile[ly + by[o]][lx + bx[o]]),     t10 = unpack_rgba(tile[ly + by[o]][lx + bx[o] + 1]);
    const float4 t01 = unpack_rgba(tile[ly + by[o] + 1][lx + bx[o]]), t11 = unpack_rgba(tile[ly + by[o] + 1][lx + bx[o] + 1]);
    const float c[3] = {
      (t00.x * (1.0f - ax[o]) + t10.x * ax[o]) * (1.0f - ay[o]) + (t01.x * (1.0f - ax[o]) + t11.x * ax[o]) * ay[o],
      (t00.y * (1.0f - ax[o]) + t10.y * ax[o]) * (1.0f - ay[o]) + (t01.y * (1.0f - ax[o]) + t11.y * ax[o]) * ay[o],
      (t00.z * (1.0f - ax[o]) + t10.z * ax[o]) * (1.0f - ay[o]) + (t01.z * (1.0f - ax[o]) + t11.z * ax[o]) * ay[o] };
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      const float e = clampf(1.0f - 0.5f * (wc[k] * fabsf(gamma08(c[k], lme_ctx) - g0[k])), 0.0f, 1.0f);
      const float ww = e * (1.0f - t) / 4.0f;
      sum[k] += ww * c[k];
      wgt[k] += ww;
    }
  }
  int ox, oy; swizzle(x, y, w, h, ox, oy);
  st_rgba(out, w, ox, oy, make_float4(div_f(sum[0], wgt[0]), div_f(sum[1], wgt[1]), div_f(sum[2], wgt[2]), 1.0f));   // wgt >= 0.2
}

struct asm_consts_t { float rgb_to_yuv[9], yuv_to_rgb[9]; float wb[3], black[3], white[3]; float noise_a, noise_b, blk, thrs0, i2thrs0; float inorm[3], denorm[3];
                      float bb[4], ibb[4];   // 0.7^(l+1) / blk and its reciprocal: launch constants, evaluated on the host
                      double rd_wb[3], rd_wbal[3], rd_2t0, rd_2t1; }; // strict, for div_rd: 1 / (white - black), 1 / wb, 1 / (2 thrs0), 1 / (2 * 10000)

// ---- assemble: wavelet shrinkage over the 4 detail bands (assemble.comp:43-165) ----
__global__ void __launch_bounds__(256, 5) k_denoise_assemble(const uint2 *__restrict__ s0, const uint2 *__restrict__ s1, const uint2 *__restrict__ s2,
    const uint2 *__restrict__ s3, const uint2 *__restrict__ s4, uint2 *__restrict__ out, int w, int h,
    const __grid_constant__ denoise_params_t p, const __grid_constant__ asm_consts_t K, const band_t bd)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= w || y >= h || BAND_SKIP(y)) return;
  const int szx = w + 1, szy = h + 1;
  const float4 orig = ld_rgba(s0, w, x, y);
  float d[5][3] = { { orig.x, orig.y, orig.z } };
  int ex = x, ey = y;
  const uint2 *s[5] = { s0, s1, s2, s3, s4 };
#pragma unroll
  for(int l = 1; l <= 4; l++)
  {
    ex = ex / 2 + ((ex & 1) * szx) / 2;
    ey = ey / 2 + ((ey & 1) * szy) / 2;
    const float4 v = ld_rgba_clamp(s[l], w, h, ex, ey);
    d[l][0] = v.x; d[l][1] = v.y; d[l][2] = v.z;
  }
  float sigma[3];
#if VKB_FAST
  noise_sigma(K.noise_a, K.noise_b, K.black[1], K.white[1], p.edges, fmaxf(d[2][0], 0.0f), sigma);
#else
  { // noise_sigma() with its quotient by the launch constant (white - black) through div_rd
    const float sq = sqrt_f(K.noise_a + fmaxf(0.0f, div_rd(fmaxf(d[2][0], 0.0f) - K.black[1], K.rd_wb[1])) * K.noise_b);
#pragma unroll
    for(int k = 0; k < 3; k++) sigma[k] = clampf(p.edges[k] * sq, 1e-3f, 1e3f);
  }
#endif
  const float bb[4] = { K.bb[0], K.bb[1], K.bb[2], K.bb[3] };
#if VKB_FAST
  const float isig[3] = { __frcp_rn(sigma[0]), __frcp_rn(sigma[1]), __frcp_rn(sigma[2]) };
#endif
  float down4[3] = { d[4][0], d[4][1], d[4][2] }, len[4];
#pragma unroll
  for(int l = 0; l < 4; l++)
  {
#pragma unroll
#if VKB_FAST
    for(int k = 0; k < 3; k++) d[l][k] = (d[l][k] - d[l + 1][k]) * (isig[k] * K.ibb[l]); // 1/(sigma bb) as a product of reciprocals: 3 instead of 12
#else
    for(int k = 0; k < 3; k++) d[l][k] = div_f(d[l][k] - d[l + 1][k], sigma[k] * bb[l]);   // sigma in [1e-3, 1e3], bb > 0
#endif
    len[l] = sqrt_f(d[l][0] * d[l][0] + d[l][1] * d[l][1] + d[l][2] * d[l][2]);
  }
  const float slope = div_c(div_c(len[3] - len[0], 3.0f) + (len[2] - len[1]) / 1.0f + (len[1] - len[0]) / 1.0f
      + (len[3] - len[2]) / 1.0f + (len[2] - len[0]) / 2.0f + (len[3] - len[1]) / 2.0f, 6.0f);
  float test = fmaxf(0.0f, -slope);
  test = fmaxf(0.0f, 1.0f - test);
#if VKB_FAST
  test = test * test; test = test * test; test = test * test; test = test * test; // pow(test, 16)
#else
  test = m_pow_nn_le1(test, 16.0f);   // test in [0, 1], a multiple of 2^-24 (or nan): no out of line call
#endif
  test = clampf(1.5f * test, 0.0f, 1.0f);
#pragma unroll
  for(int l = 3; l >= 0; l--)
  {
    const bool big = fabsf(d[l][0]) > 10.0f;
    const float thrs = big ? 10000.0f : K.thrs0, i2t = big ? 0.5f / 10000.0f : K.i2thrs0;
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      const float a = fabsf(d[l][k]);
#if VKB_FAST
      const float tt = fminf(1.0f, a * i2t);
#else
      const float tt = fminf(1.0f, div_rd(a, big ? K.rd_2t1 : K.rd_2t0));   // a / (2.0f * thrs)
#endif
      down4[k] += sigma[k] * bb[l] * signf(d[l][k]) * mixf(fmaxf(a - thrs, 0.0f), a, tt);
    }
  }
  float v[3], vo[3], yuv[3], yuvo[3], rgb[3];
  const float og[3] = { orig.x, orig.y, orig.z };
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
#if VKB_FAST
    v[k]  = (down4[k] - K.black[k]) * K.inorm[k];
    vo[k] = (og[k]    - K.black[k]) * K.inorm[k];
#else
    v[k]  = div_rd(down4[k] - K.black[k], K.rd_wb[k]) * K.wb[k];   // (x - black) / (white - black) * wb
    vo[k] = div_rd(og[k]    - K.black[k], K.rd_wb[k]) * K.wb[k];
#endif
  }
#pragma unroll
  for(int j = 0; j < 3; j++)
  {
    yuv[j]  = K.rgb_to_yuv[3 * j] * v[0]  + K.rgb_to_yuv[3 * j + 1] * v[1]  + K.rgb_to_yuv[3 * j + 2] * v[2];
    yuvo[j] = K.rgb_to_yuv[3 * j] * vo[0] + K.rgb_to_yuv[3 * j + 1] * vo[1] + K.rgb_to_yuv[3 * j + 2] * vo[2];
  }
  yuv[0] = mixf(yuvo[0], yuv[0], p.luma);
#pragma unroll
  for(int j = 0; j < 3; j++) rgb[j] = K.yuv_to_rgb[3 * j] * yuv[0] + K.yuv_to_rgb[3 * j + 1] * yuv[1] + K.yuv_to_rgb[3 * j + 2] * yuv[2];
#if VKB_FAST
  st_rgba(out, w, x, y, make_float4(rgb[0] * K.denorm[0] + K.black[0], rgb[1] * K.denorm[1] + K.black[1], rgb[2] * K.denorm[2] + K.black[2], test));
#else
  st_rgba(out, w, x, y, make_float4(div_rd(rgb[0], K.rd_wbal[0]) * (K.white[0] - K.black[0]) + K.black[0],
        div_rd(rgb[1], K.rd_wbal[1]) * (K.white[1] - K.black[1]) + K.black[1],
        div_rd(rgb[2], K.rd_wbal[2]) * (K.white[2] - K.black[2]) + K.black[2], test));   // rgb / wb * (white - black) + black
#endif
}

// ---- doub: per-colour residual shrink on the full resolution mosaic (doub.comp:35-115) ----
// the per pixel part after the two coarse lookups: upsm_c / down_c are the pixel's own colour channel of crs0 / crs1, upw = crs0.w
// strict: 1 / (white - black) per colour in double, from the launcher (div_rd: the quotient by a launch constant)
struct doub_rd_t { double rd[3]; };
VKB_DEV float doub_shrink(float val, float upsm_c, float down_c, float upw, int col, bool xt,
    const denoise_params_t &p, const dn_push_doub_t &P, const doub_rd_t &R)
{
  float black = P.black[1], white = P.white[1];
  float T = 0.5f * p.strength * upw, blendw = p.luma;
  if(col != 1)
  {
    black = col == 0 ? P.black[0] : P.black[2]; white = col == 0 ? P.white[0] : P.white[2];
    blendw = 1.0f;
    if(xt) T = div_f(T, fmaxf(1e-4f, upw));
  }
  float sigma[3];
#if VKB_FAST
  noise_sigma(P.noise_a, P.noise_b, black, white, p.edges, upsm_c, sigma);
#else
  { // noise_sigma() without an out of line call: the quotient by (white - black) through div_rd, sqrt.rn's fast path (a, b >= 0)
    const float s = sqrt_f(P.noise_a + fmaxf(0.0f, div_rd(upsm_c - black, R.rd[col])) * P.noise_b);
#pragma unroll
    for(int k = 0; k < 3; k++) sigma[k] = clampf(p.edges[k] * s, 1e-3f, 1e3f);
  }
#endif
  blendw = 0.5f * (blendw + 1.0f);
  if(val < white)
  {
    const float wav = div_f(val - down_c, fmaxf(sigma[0] + sigma[2], 1e-8f));
    const float tt = fminf(1.0f, div_f(wav, fmaxf(2.0f * T, 1e-8f)));
#if VKB_FAST
    float uw = fminf(1.0f, 1.0f * upw); uw = uw * uw; uw = uw * uw; // pow(.., 4)
#else
    float uw = m_pow_nn(fminf(1.0f, 1.0f * upw), 4.0f);   // upw: a bilinear blend of f16 edge values in [0, 1]: +0 or >= 2^-24 * 1/16, 4 log2 < 126
#endif
    uw = 1.0f - (1.0f - uw) * p.detail;
    val = mixf(val, fmaxf(0.0f, upsm_c + sigma[1] * signf(wav) * mixf(fmaxf(0.0f, fabsf(wav) - T), fabsf(wav), tt)), uw * blendw);
  }
#if VKB_FAST
  return fmaxf(0.0f, div_f(val - black, white - black));
#else
  return fmaxf(0.0f, div_rd(val - black, R.rd[col]));
#endif
}

__global__ void __launch_bounds__(256) k_denoise_doub(const uint16_t *__restrict__ in, int iw, int ih,
    const uint2 *__restrict__ crs0, const uint2 *__restrict__ crs1, int cw, int ch, __half *__restrict__ out, int ow, int oh,
    const __grid_constant__ denoise_params_t p, const __grid_constant__ dn_push_doub_t P, const gainmap_t G, const doub_rd_t R)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  // texture(img_crs, (ipos+0.5)/imageSize(out)): ideal sampler, coordinates in double
  int bx, by; float ax, ay;
  if(ow == 2 * cw && oh == 2 * ch)
  { // (x+0.5)/2 - 0.5 = x/2 - 0.25: even x -> texel x/2-1 + 0.75, odd x -> texel (x-1)/2 + 0.25, exactly
    bx = (x >> 1) - ((x & 1) ? 0 : 1); ax = (x & 1) ? 0.25f : 0.75f;
    by = (y >> 1) - ((y & 1) ? 0 : 1); ay = (y & 1) ? 0.25f : 0.75f;
  }
  else if(ow == 3 * cw && oh == 3 * ch)
  { // x-trans: (x+0.5)/3 - 0.5 = (x-1)/3: texel floor((x-1)/3), fraction 0, 1/3 or 2/3.  the double arithmetic of the general
    // branch below lands within 1e-12 of those, which rounds to the floats 0x3eaaaaab / 0x3f2aaaab (their neighbours' midpoints
    // are 5e-9 away); a fraction of 0 +- 1e-13 selects the same texel either way.  no double division per pixel.
    const int qx = x + 2, qy = y + 2;                      // (x - 1) + 3: non negative
    const int rx = qx % 3, ry = qy % 3;
    bx = qx / 3 - 1; by = qy / 3 - 1;
    ax = rx == 0 ? 0.0f : (rx == 1 ? __uint_as_float(0x3eaaaaabu) : __uint_as_float(0x3f2aaaabu));
    ay = ry == 0 ? 0.0f : (ry == 1 ? __uint_as_float(0x3eaaaaabu) : __uint_as_float(0x3f2aaaabu));
  }
  else
  {
    const double ux = ((double)x + 0.5) / (double)ow * (double)cw - 0.5, uy = ((double)y + 0.5) / (double)oh * (double)ch - 0.5;
    const double fx = floor(ux), fy = floor(uy);
    bx = (int)fx; by = (int)fy; ax = (float)(ux - fx); ay = (float)(uy - fy);
  }
  const float4 upsm = bilin_rgba(crs0, cw, ch, bx, by, ax, ay);
  const float4 down = bilin_rgba(crs1, cw, ch, bx, by, ax, ay);
  const bool xt = P.filters == 9;
  const int col = xt ? xtrans_colour(x, y) : bayer_colour(x, y);
  const float val = div_c((float)__ldg(in + (size_t)mirrori(y + P.crop[1], ih) * iw + mirrori(x + P.crop[0], iw)), 65535.0f);
  const float uc = col == 1 ? upsm.y : (col == 0 ? upsm.x : upsm.z), dc = col == 1 ? down.y : (col == 0 ? down.x : down.z);
  float res = doub_shrink(val, uc, dc, upsm.w, col, xt, p, P, R);
  if(G.map) res *= gainmap_gain(G, x, y, P.crop[0], P.crop[1], ow, oh, 2);   // doub.comp:106-114
  out[(size_t)y * ow + x] = __float2half_rn(res);
}

// bayer, output exactly twice the coarse size: one thread per 2x2 block.  the four pixels' bilinear taps (fractions
// .75/.25 on texels X-1..X+1) share one 3x3 window of each coarse image and each pixel only needs its own colour channel,
// so a block costs 9+9 texel loads and 25 f16 conversions instead of 32 loads and 96 conversions.  per pixel the
// expressions are those of bilin_rgba() term by term.
__global__ void __launch_bounds__(256, 5) k_denoise_doub_bayer(const uint16_t *__restrict__ in, int iw, int ih,
    const uint2 *__restrict__ crs0, const uint2 *__restrict__ crs1, int cw, int ch, __half *__restrict__ out, int ow, int oh,
    const __grid_constant__ denoise_params_t p, const __grid_constant__ dn_push_doub_t P, const doub_rd_t R, const band_t bd)
{
  const int X = blockIdx.x * 32 + threadIdx.x, Y = BAND_BY * 8 + threadIdx.y;
  if(X >= cw || Y >= ch || BAND_SKIP(Y)) return;
  int xi[3], yi[3];
#pragma unroll
  for(int k = 0; k < 3; k++) { xi[k] = mirrori(X - 1 + k, cw); yi[k] = mirrori(Y - 1 + k, ch); }
  // channel c of the window, as floats: [j][i]
  float u[4][3][3], d[3][3][3];
#pragma unroll
  for(int j = 0; j < 3; j++)
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
      const float4 a = ld_rgba(crs0, cw, xi[i], yi[j]), b = ld_rgba(crs1, cw, xi[i], yi[j]);
      u[0][j][i] = a.x; u[1][j][i] = a.y; u[2][j][i] = a.z; u[3][j][i] = a.w;
      d[0][j][i] = b.x; d[1][j][i] = b.y; d[2][j][i] = b.z;
    }
  // pixel (dx, dy) of the block: base texel index = dx/dy (0 or 1) in the window, fraction .75 (even) or .25 (odd)
#define BIL(T, DX, DY) ((T[DY][DX] * (1.0f - AX(DX)) + T[DY][DX + 1] * AX(DX)) * (1.0f - AX(DY)) + (T[DY + 1][DX] * (1.0f - AX(DX)) + T[DY + 1][DX + 1] * AX(DX)) * AX(DY))
#define AX(D) ((D) ? 0.25f : 0.75f)
  const int ry0 = mirrori(2 * Y + P.crop[1], ih), ry1 = mirrori(2 * Y + 1 + P.crop[1], ih);
  const int rx0 = mirrori(2 * X + P.crop[0], iw), rx1 = mirrori(2 * X + 1 + P.crop[0], iw);
  const float v00 = div_c((float)__ldg(in + (size_t)ry0 * iw + rx0), 65535.0f), v10 = div_c((float)__ldg(in + (size_t)ry0 * iw + rx1), 65535.0f);
  const float v01 = div_c((float)__ldg(in + (size_t)ry1 * iw + rx0), 65535.0f), v11 = div_c((float)__ldg(in + (size_t)ry1 * iw + rx1), 65535.0f);
  const float o00 = doub_shrink(v00, BIL(u[0], 0, 0), BIL(d[0], 0, 0), BIL(u[3], 0, 0), 0, false, p, P, R); // r
  const float o10 = doub_shrink(v10, BIL(u[1], 1, 0), BIL(d[1], 1, 0), BIL(u[3], 1, 0), 1, false, p, P, R); // g
  const float o01 = doub_shrink(v01, BIL(u[1], 0, 1), BIL(d[1], 0, 1), BIL(u[3], 0, 1), 1, false, p, P, R); // g
  const float o11 = doub_shrink(v11, BIL(u[2], 1, 1), BIL(d[2], 1, 1), BIL(u[3], 1, 1), 2, false, p, P, R); // b
#undef BIL
#undef AX
  *reinterpret_cast<__half2 *>(out + (size_t)(2 * Y) * ow + 2 * X)     = __floats2half2_rn(o00, o10);
  *reinterpret_cast<__half2 *>(out + (size_t)(2 * Y + 1) * ow + 2 * X) = __floats2half2_rn(o01, o11);
}

static inline dim3 grid2d(unsigned w, unsigned h, unsigned by = 8) { return dim3(vkb_cdiv(w, 32), vkb_cdiv(h, by)); }
static const dim3 blk2d(32, 8);

// conn: [0] input ui16 raw, [1] output rgba f16 (1/block res)
static int launch_half(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= sizeof(dn_push_half_t));
  const dn_push_half_t *pc = (const dn_push_half_t *)l->push;
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->format == VKB_TOKEN_UI16 && in->chan == 1 && out->chan == 4 && out->format == VKB_TOKEN_F16);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_denoise_half<<<grid, blk2d, 0, l->stream>>>((const uint16_t *)in->data, in->wd, in->ht, (uint2 *)out->data,
      out->wd, out->ht, pc->crop[0], pc->crop[1], pc->white[1], pc->filters == 9, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("denoise", "half", launch_half);

// conn: [0] input rgba f16, [1] output rgba f16 (swizzled), [2] cov rgba f16
static int launch_downcov(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 3);
  const vkb_image_t *in = l->conn, *out = l->conn + 1, *cov = l->conn + 2;
  VKB_REQUIRE(in->chan == 4 && out->chan == 4 && cov->chan == 4 && in->wd == out->wd && in->ht == out->ht && cov->wd == in->wd);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_denoise_downcov<<<grid, blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, (uint2 *)out->data, (uint2 *)cov->data, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("denoise", "downcov", launch_downcov);

// conn: [0] input rgba f16 (swizzled previous level), [1] output rgba f16
static int launch_down(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= sizeof(dn_push_down_t) && l->params_size >= sizeof(denoise_params_t) - 4);
  const dn_push_down_t *pc = (const dn_push_down_t *)l->push;
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && out->chan == 4 && in->wd == out->wd && in->ht == out->ht);
  VKB_REQUIRE(pc->level >= 0); // the level < 0 branch (response()) is dead in the reference wiring
  denoise_params_t p; memset(&p, 0, sizeof(p)); memcpy(&p, l->params, l->params_size < sizeof(p) ? l->params_size : sizeof(p));
  const float blk = pc->block == 3 ? 2.23607f : (pc->block == 2 ? 1.414213f : 1.0f);
  host_escale(&p);
  const volatile float wmb = pc->white[1] - pc->black[1];
  if(in->wd >= DD_W && in->ht >= DD_H) // a window overhangs the image by less than its size: one reflection is enough
  {
    dim3 grid = grid2d(out->wd, out->ht);
    const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
    if(!grid.y) return VKB_OK;
    k_denoise_down_tiled<<<grid, blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, (uint2 *)out->data,
        p, pc->black[1], pc->white[1], pc->noise_a, pc->noise_b, powf(0.7f, (float)pc->level), blk, 1.0 / (double)wmb, 1.0 / (double)blk, bd);
  }
  else
  k_denoise_down<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, (uint2 *)out->data,
      p, pc->black[1], pc->white[1], pc->noise_a, pc->noise_b, powf(0.7f, (float)pc->level), blk);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("denoise", "down", launch_down);

static void mat3mul(const float *A, const float *B, float *C)
{
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) C[3 * j + i] = A[3 * j + 0] * B[0 + i] + A[3 * j + 1] * B[3 + i] + A[3 * j + 2] * B[6 + i];
}
// conn: [0..4] s0..s4 rgba f16, [5] output rgba f16
static int launch_assemble(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 6 && l->push_size >= sizeof(dn_push_asm_t) && l->params_size >= sizeof(denoise_params_t) - 4);
  const dn_push_asm_t *pc = (const dn_push_asm_t *)l->push;
  const vkb_image_t *c = l->conn, *out = l->conn + 5;
  for(int k = 0; k < 5; k++) VKB_REQUIRE(c[k].chan == 4 && c[k].wd == out->wd && c[k].ht == out->ht);
  denoise_params_t p; memset(&p, 0, sizeof(p)); memcpy(&p, l->params, l->params_size < sizeof(p) ? l->params_size : sizeof(p));
  static const float rec709_to_yuv[9] = {0.299f, 0.587f, 0.114f, -0.147f, -0.289f, 0.436f, 0.615f, -0.515f, -0.100f};
  static const float yuv_to_rec709[9] = {1.0f, -3.94570707e-05f, 1.13982797f, 1.0f, -3.94610164e-01f, -5.80500316e-01f, 1.0f, 2.03199968f, -4.81376263e-04f};
  static const float rec2020_to_rec709[9] = {1.66022677f, -0.58754761f, -0.07283825f, -0.12455334f, 1.13292605f, -0.00834963f, -0.01815514f, -0.10060303f, 1.11899817f};
  static const float rec709_to_rec2020[9] = {0.62750375f, 0.32927542f, 0.04330266f, 0.06910828f, 0.91951916f, 0.0113596f, 0.01639406f, 0.08801125f, 0.89538035f};
  asm_consts_t K;
  mat3mul(rec709_to_yuv, rec2020_to_rec709, K.rgb_to_yuv);
  mat3mul(rec709_to_rec2020, yuv_to_rec709, K.yuv_to_rgb);
  for(int k = 0; k < 3; k++) { K.wb[k] = pc->wb[k]; K.black[k] = pc->black[k]; K.white[k] = pc->white[k]; }
  K.noise_a = pc->noise_a; K.noise_b = pc->noise_b;
  K.blk = pc->filters == 0u ? 1.0f : (pc->filters == 9u ? 2.23607f : 1.414213f);
  K.thrs0 = powf(p.strength, 4.0f);
  { const float pw[4] = { 0.7000f, 0.4900f, 0.3430f, 0.2401f }; for(int l = 0; l < 4; l++) { K.bb[l] = pw[l] / K.blk; K.ibb[l] = 1.0f / K.bb[l]; } }
  K.i2thrs0 = 1.0f / (2.0f * K.thrs0);
  for(int k = 0; k < 3; k++)
  { // per-launch constants: wb/(white-black) and its inverse, computed in double
    K.inorm[k]  = (float)((double)K.wb[k] / ((double)K.white[k] - (double)K.black[k]));
    K.denorm[k] = (float)(((double)K.white[k] - (double)K.black[k]) / (double)K.wb[k]);
  }
  for(int k = 0; k < 3; k++)
  {
    const volatile float wmb = K.white[k] - K.black[k];
    K.rd_wb[k] = 1.0 / (double)wmb; K.rd_wbal[k] = 1.0 / (double)K.wb[k];
  }
  { const volatile float t0 = 2.0f * K.thrs0; K.rd_2t0 = 1.0 / (double)t0; K.rd_2t1 = 1.0 / 20000.0; }
  host_escale(&p);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_denoise_assemble<<<grid, blk2d, 0, l->stream>>>((const uint2 *)c[0].data, (const uint2 *)c[1].data, (const uint2 *)c[2].data,
      (const uint2 *)c[3].data, (const uint2 *)c[4].data, (uint2 *)out->data, out->wd, out->ht, p, K, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("denoise", "assemble", launch_assemble);

// conn: [0] orig ui16 raw, [1] crs0 (assembled) rgba f16, [2] crs1 (half) rgba f16, [3] output mosaic f16, [4] gainmap (ignored)
static int launch_doub(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 4 && l->push_size >= sizeof(dn_push_doub_t) && l->params_size >= sizeof(denoise_params_t) - 4);
  const vkb_image_t *in = l->conn, *c0 = l->conn + 1, *c1 = l->conn + 2, *out = l->conn + 3;
  VKB_REQUIRE(in->format == VKB_TOKEN_UI16 && c0->chan == 4 && c1->chan == 4 && c0->wd == c1->wd && out->chan == 1 && out->format == VKB_TOKEN_F16);
  denoise_params_t p; memset(&p, 0, sizeof(p)); memcpy(&p, l->params, l->params_size < sizeof(p) ? l->params_size : sizeof(p));
  dn_push_doub_t P; memcpy(&P, l->push, sizeof(P));
  host_escale(&p);
  // [4] gain map rgba f32 (a dummy binding unless push.gainmap, denoise/main.c:315-319); bayer only (doub.comp:106)
  gainmap_t G = { 0, 0, 0, { 0, 0, 0, 0 } };
  if(P.filters != 9u && P.gainmap == 1 && p.gainmap == 1)
  {
    VKB_REQUIRE(l->num_conn >= 5 && l->conn[4].format == VKB_TOKEN_F32 && l->conn[4].chan == 4 && l->conn[4].data && l->band_y0 < 0);
    G.map = (const float4 *)l->conn[4].data; G.w = (int)l->conn[4].wd; G.h = (int)l->conn[4].ht;
    for(int k = 0; k < 4; k++) G.os[k] = P.map_os[k];
  }
  doub_rd_t R;
  for(int k = 0; k < 3; k++) { const volatile float wmb = P.white[k] - P.black[k]; R.rd[k] = 1.0 / (double)wmb; }
  if(!G.map && P.filters != 9u && P.filters != 0u && out->wd == 2 * c0->wd && out->ht == 2 * c0->ht)
  {
    dim3 grid = grid2d(c0->wd, c0->ht);
    const band_t bd = band_of(l, 2, 8, c0->ht, &grid.y); // band image: the output mosaic, two rows per thread row
    if(!grid.y) return VKB_OK;
    k_denoise_doub_bayer<<<grid, blk2d, 0, l->stream>>>((const uint16_t *)in->data, in->wd, in->ht, (const uint2 *)c0->data,
        (const uint2 *)c1->data, c0->wd, c0->ht, (__half *)out->data, out->wd, out->ht, p, P, R, bd);
  }
  else
  k_denoise_doub<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const uint16_t *)in->data, in->wd, in->ht, (const uint2 *)c0->data,
      (const uint2 *)c1->data, c0->wd, c0->ht, (__half *)out->data, out->wd, out->ht, p, P, G, R);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("denoise", "doub", launch_doub);

VKB_NS_END

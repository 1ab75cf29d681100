/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/denoise/{noop,half,downcov,down,assemble,doub}.comp,
 * cov.glsl and noise.glsl.  Gain maps (DNG opcode lists) are outside the hot-path scope. */
#include "o_common.h"
#include "vkdt_oracle.h"

/* denoise/noop.comp:36-57.  `in` is the ui16 source sampled as UNORM (x/65535).
 * the reference stores (v,0,0,1) into an rgba image; consumers read .r only, we keep .r. */
void o_denoise_noop(const oimg_t *in, oimg_t *out, const int *crop, const float *black, const float *white)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float col = o_fetch1(in, x + crop[0], y + crop[1]);
    col = o_max(0.0f, (col - black[0]) / (white[0] - black[0]));
    o_store1(out, x, y, col, 1);
  }
}

/* the DNG GainMap branch of noop.comp:48-57 and doub.comp:106-114: a low resolution rgba f32 texture (one gain per cfa site of
 * the 2x2 block) sampled at the pixel's position in the uncropped image; map_os = { origin x, origin y, 1 / extent x, 1 / extent y }
 * (denoise/main.c:188-195).  block = 1: noop (full resolution position), 2: doub (position of the 2x2 block) */
static float o_gainmap_gain(const oimg_t *gm, const float *map_os, int x, int y, int cx, int cy, int sw, int sh, int block)
{
  float px, py;
  if(block == 1) { px = (0.5f + (float)(x + cx)) / (float)sw; py = (0.5f + (float)(y + cy)) / (float)sh; }
  else           { px = (0.5f + (float)((x + cx) / 2)) / (float)(sw / 2); py = (0.5f + (float)((y + cy) / 2)) / (float)(sh / 2); }
  px = o_clamp(px * map_os[2] - map_os[0], 0.0f, 1.0f);
  py = o_clamp(py * map_os[3] - map_os[1], 0.0f, 1.0f);
  float g[4];
  o_tex4(gm, (double)px, (double)py, g);
  return g[(x & 1) + (y & 1) * 2];
}
/* noop.comp with a gain map (bayer only: filters != 0 && filters != 9, push.gainmap == 1 && params.gainmap == 1) */
void o_denoise_noop_gm(const oimg_t *in, oimg_t *out, const int *crop, const float *black, const float *white, const oimg_t *gm, const float *map_os)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float col = o_fetch1(in, x + crop[0], y + crop[1]);
    col = o_max(0.0f, (col - black[0]) / (white[0] - black[0]));
    col *= o_gainmap_gain(gm, map_os, x, y, crop[0], crop[1], in->w, in->h, 1);
    o_store1(out, x, y, col, 1);
  }
}
/* X-Trans colour at absolute position: demosaic/splat.comp:52-68, denoise/doub.comp:52-66.
 * returns 0 red, 1 green, 2 blue */
int o_xtrans_colour(int x, int y)
{
  const int blue_top = ((x / 3 + y / 3) & 1) > 0;
  const int qx = x - (x / 3) * 3, qy = y - (y / 3) * 3;
  if(((qx + qy) & 1) == 0) return 1;
  if(blue_top ^ (qy == 1)) return 2;
  return 0;
}

/* denoise/half.comp:24-74 */
void o_denoise_half(const oimg_t *in, oimg_t *out, const int *crop, const float *white4, uint32_t filters)
{
  const float white = white4[1];
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgba[4];
    if(filters == 9)
    {
      float c[9]; /* c[3*i+j] = texel (i,j) : index runs column first like the shader's c0..c8 */
      for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++)
        c[3*i+j] = o_fetch1(in, crop[0] + 3*x + i, crop[1] + 3*y + j);
      const float col0 = (c[1] + c[7]) * 0.5f, col1 = (c[3] + c[5]) * 0.5f;
      if(((x + y + crop[0] + crop[1]) & 1) > 0) { rgba[0] = col0; rgba[2] = col1; }
      else                                      { rgba[2] = col0; rgba[0] = col1; }
      rgba[1] = (c[0] + c[2] + c[4] + c[6] + c[8]) * 1.0f / 5.0f;
      rgba[3] = 1.0f;
    }
    else
    {
      float c[4];
      o_gather(in, (crop[0] + 2.0 * (x + .5)) / (double)in->w, (crop[1] + 2.0 * (y + .5)) / (double)in->h, c);
      if(c[0] >= white) c[0] = c[2];
      if(c[2] >= white) c[2] = c[0];
      rgba[0] = c[3]; rgba[1] = (c[0] + c[2]) / 2.0f; rgba[2] = c[1]; rgba[3] = 1.0f;
    }
    o_store4(out, x, y, rgba, 1);
  }
}

/* denoise/noise.glsl:1-11 */
static inline void noise_sigma(float a, float b, float black, float white, const float *edges, float val, float *sig)
{
  const float s = sqrtf(a + o_max(0.0f, (val - black) / (white - black)) * b);
  for(int k = 0; k < 3; k++)
    sig[k] = o_clamp(exp2f(12.0f * edges[k] + edges[3]) * s, 1e-3f, 1e3f);
}

/* denoise/cov.glsl:21-138 `response` */
static void response(const oimg_t *img, int px_, int py_, float *cov, float *res)
{
  const double iszx = 1.0 / (double)img->w, iszy = 1.0 / (double)img->h;
  float Sw[4] = {0}, Sb[4] = {0}; /* [0][0] [0][1] [1][0] [1][1] */
  float sw = 0.0f, sb = 0.0f;
  float mw[2] = {0}, mb[2] = {0};
  float smw = 0.0f, smb = 0.0f, mean_b = 0.0f;
  float t[4];
  for(int j = -2; j <= 2; j++) for(int i = -2; i <= 2; i++)
  {
    o_tex4(img, (px_ + 0.5 + i) * iszx, (py_ + 0.5 + j) * iszy, t);
    const float px = o_lum2020(t);
    const float w = 1.0f;
    mw[0] += (float)i * px * w; mw[1] += (float)j * px * w;
    smw += px * w;
    mb[0] += (float)i / px * w; mb[1] += (float)j / px * w;
    smb += 1.0f / px * w;
  }
  mw[0] /= smw; mw[1] /= smw;
  mb[0] /= smb; mb[1] /= smb;
  for(int j = -2; j <= 2; j++) for(int i = -2; i <= 2; i++)
  {
    o_tex4(img, (px_ + 0.5 + i) * iszx, (py_ + 0.5 + j) * iszy, t);
    const float px = o_lum2020(t);
    const float w = 1.0f;
    mean_b += px / 25.0f;
    float p2 = px * px * w;
    float p0 = (float)i - mw[0], p1 = (float)j - mw[1];
    Sw[0] += p2 * p0 * p0; Sw[1] += p2 * p0 * p1;
    Sw[2] += p2 * p1 * p0; Sw[3] += p2 * p1 * p1;
    sw += p2;
    p0 = (float)i - mb[0]; p1 = (float)j - mb[1];
    p2 = 1.0f / (px * px) * w;
    Sb[0] += p2 * p0 * p0; Sb[1] += p2 * p0 * p1;
    Sb[2] += p2 * p1 * p0; Sb[3] += p2 * p1 * p1;
    sb += p2;
  }
  for(int k = 0; k < 4; k++) { Sw[k] /= sw; Sb[k] /= sb; }
  const float detw = Sw[0] * Sw[3] - Sw[1] * Sw[2];
  const float detb = Sb[0] * Sb[3] - Sb[1] * Sb[2];
  const float *S = detw < detb ? Sw : Sb;
  float eval[2], evec0[2], evec1[2];
  o_evd2x2(S[0], S[2], S[3], eval, evec0, evec1);
  eval[1] *= 0.05f;
  eval[0] = o_clamp(eval[0], 0.01f, 25.0f);
  eval[1] = o_clamp(eval[1], 0.01f, 25.0f);
  cov[0] = eval[0]; cov[1] = eval[1]; cov[2] = evec0[0]; cov[3] = evec0[1];

  float acc[3] = {0}, wt = 0.0f;
  for(int j = -2; j <= 2; j++) for(int i = -2; i <= 2; i++)
  {
    o_tex4(img, (px_ + 0.5 + i) * iszx, (py_ + 0.5 + j) * iszy, t);
    const float ht = 2.0f;
    if(t[0] > ht * mean_b) continue; /* hot pixels */
    const float x0 = (float)i * evec0[0] + (float)j * evec0[1];
    const float x1 = (float)i * evec1[0] + (float)j * evec1[1];
    const float w = o_max(1e-9f, expf(-0.5f * (x0 / eval[0] * x0 + x1 / eval[1] * x1)));
    for(int k = 0; k < 3; k++) acc[k] += w * t[k];
    wt += w;
  }
  for(int k = 0; k < 3; k++) res[k] = acc[k] / o_max(wt, 1e-8f);
}

/* swizzled write position shared by downcov.comp:52-53 and down.comp:103-104 */
static inline void swizzle(int x, int y, int w, int h, int *ox, int *oy)
{
  *ox = x / 2 + ((x & 1) * (w + 1)) / 2;
  *oy = y / 2 + ((y & 1) * (h + 1)) / 2;
}

/* denoise/downcov.comp:41-62 (crop is zero on mosaic input, denoise/main.c:244-253) */
void o_denoise_downcov(const oimg_t *in, oimg_t *out, oimg_t *covimg)
{
#pragma omp parallel for schedule(dynamic, 4)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float cov[4], rgb[4];
    response(in, x, y, cov, rgb);
    o_store4(covimg, x, y, cov, 1);
    int ox, oy; swizzle(x, y, out->w, out->h, &ox, &oy);
    float edge = o_clamp(75.0f * o_max(0.0f, cov[1] - 0.09f), 0.0f, 1.0f);
    edge = o_smoothstep(0.4f, 0.75f, edge);
    edge = o_clamp(0.02f + edge, 0.0f, 1.0f);
    rgb[3] = edge;
    o_store4(out, ox, oy, rgb, 1);
  }
}

static inline float gamma08(float f) { return f < 0.0f ? f : powf(f, 0.8f); }

/* denoise/down.comp:59-107, level >= 0 branch */
void o_denoise_down(const oimg_t *in, oimg_t *out, const o_denoise_params_t *p,
    const float *black4, const float *white4, float noise_a, float noise_b, int level, uint32_t block)
{
  const float t = 0.2f;
  const float blk = (block == 3) ? 2.23607f : (block == 2 ? 1.414213f : 1.0f);
  const double szx = (double)in->w, szy = (double)in->h;
  static const float off[4][2] = {
    { (float)(0.5 + 1.2), (float)(0.5 + 0.4) }, { (float)(0.5 - 1.2), (float)(0.5 - 0.4) },
    { (float)(0.5 + 0.4), (float)(0.5 - 1.2) }, { (float)(0.5 - 0.4), (float)(0.5 + 1.2) } }; /* glslang folds constants in double */
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float c0[4];
    o_tex4(in, (x + 0.5) / szx, (y + 0.5) / szy, c0);
    float sum[3], wgt[3], sigma[3], wc[3];
    noise_sigma(noise_a, noise_b, black4[1], white4[1], p->edges, c0[0], sigma);
    const float lv = powf(0.7f, (float)level);
    for(int k = 0; k < 3; k++)
    {
      sum[k] = t * c0[k]; wgt[k] = t;
      sigma[k] = lv * sigma[k] / blk;
      wc[k] = 1.0f / sigma[k];
    }
    for(int o = 0; o < 4; o++)
    {
      float col[4];
      o_tex4(in, ((double)x + off[o][0]) / szx, ((double)y + off[o][1]) / szy, col);
      for(int k = 0; k < 3; k++)
      {
        const float e = o_clamp(1.0f - 0.5f * (wc[k] * fabsf(gamma08(col[k]) - gamma08(c0[k]))), 0.0f, 1.0f);
        const float w = e * (1.0f - t) / 4.0f;
        sum[k] += w * col[k];
        wgt[k] += w;
      }
    }
    float rgba[4] = { sum[0] / wgt[0], sum[1] / wgt[1], sum[2] / wgt[2], 1.0f };
    int ox, oy; swizzle(x, y, out->w, out->h, &ox, &oy);
    o_store4(out, ox, oy, rgba, 1);
  }
}

static void mat3mul(const float *A, const float *B, float *C)
{
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
    C[3*j+i] = A[3*j+0]*B[0+i] + A[3*j+1]*B[3+i] + A[3*j+2]*B[6+i];
}

/* denoise/assemble.comp:43-165 */
void o_denoise_assemble(const oimg_t *s0, const oimg_t *s1, const oimg_t *s2, const oimg_t *s3, const oimg_t *s4,
    oimg_t *out, const o_denoise_params_t *p, const float *wb, const float *black, const float *white,
    float noise_a, float noise_b, uint32_t filters)
{
  static const float rec709_to_yuv[9] = {0.299f, 0.587f, 0.114f, -0.147f, -0.289f, 0.436f, 0.615f, -0.515f, -0.100f};
  static const float yuv_to_rec709[9] = {1.0f, -3.94570707e-05f, 1.13982797f, 1.0f, -3.94610164e-01f, -5.80500316e-01f, 1.0f, 2.03199968f, -4.81376263e-04f};
  static const float rec2020_to_rec709[9] = {1.66022677f, -0.58754761f, -0.07283825f, -0.12455334f, 1.13292605f, -0.00834963f, -0.01815514f, -0.10060303f, 1.11899817f};
  static const float rec709_to_rec2020[9] = {0.62750375f, 0.32927542f, 0.04330266f, 0.06910828f, 0.91951916f, 0.0113596f, 0.01639406f, 0.08801125f, 0.89538035f};
  float rgb_to_yuv[9], yuv_to_rgb[9];
  mat3mul(rec709_to_yuv, rec2020_to_rec709, rgb_to_yuv);
  mat3mul(rec709_to_rec2020, yuv_to_rec709, yuv_to_rgb);
  const int szx = s1->w + 1, szy = s1->h + 1;
  const float blk = filters == 0u ? 1.0f : (filters == 9u ? 2.23607f : 1.414213f);
  const float bb[4] = { 0.7000f / blk, 0.4900f / blk, 0.3430f / blk, 0.2401f / blk };
  const float thrs0 = powf(p->strength, 4.0f);
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float orig[4], d[5][4];
    o_fetch4(s0, x, y, orig);
    int ex = x, ey = y;
    for(int k = 0; k < 3; k++) d[0][k] = orig[k];
    const oimg_t *s[5] = { s0, s1, s2, s3, s4 };
    for(int l = 1; l <= 4; l++)
    {
      ex = ex / 2 + ((ex & 1) * szx) / 2;
      ey = ey / 2 + ((ey & 1) * szy) / 2;
      o_fetch4(s[l], ex, ey, d[l]);
    }
    float sigma[3];
    noise_sigma(noise_a, noise_b, black[1], white[1], p->edges, o_max(d[2][0], 0.0f), sigma);
    float down4[3] = { d[4][0], d[4][1], d[4][2] };
    float len[4];
    for(int l = 0; l < 4; l++)
    {
      for(int k = 0; k < 3; k++) d[l][k] = (d[l][k] - d[l+1][k]) / (sigma[k] * bb[l]);
      len[l] = sqrtf(d[l][0]*d[l][0] + d[l][1]*d[l][1] + d[l][2]*d[l][2]);
    }
    const float slope = ((len[3] - len[0]) / 3.0f + (len[2] - len[1]) / 1.0f + (len[1] - len[0]) / 1.0f
        + (len[3] - len[2]) / 1.0f + (len[2] - len[0]) / 2.0f + (len[3] - len[1]) / 2.0f) / 6.0f;
    float test = o_max(0.0f, -slope);
    test = o_max(0.0f, 1.0f - test);
    test = powf(test, 16.0f);
    test = o_clamp(1.5f * test, 0.0f, 1.0f);
    for(int l = 3; l >= 0; l--)
    {
      const float thrs = fabsf(d[l][0]) > 10.0f ? 10000.0f : thrs0;
      for(int k = 0; k < 3; k++)
      {
        const float a = fabsf(d[l][k]);
        const float tt = o_min(1.0f, a / (2.0f * thrs));
        down4[k] += sigma[k] * bb[l] * o_sign(d[l][k]) * o_mix(o_max(a - thrs, 0.0f), a, tt);
      }
    }
    float v[3], vo[3], yuv[3], yuvo[3], rgb[3];
    for(int k = 0; k < 3; k++)
    {
      v[k]  = (down4[k] - black[k]) / (white[k] - black[k]) * wb[k];
      vo[k] = (orig[k]  - black[k]) / (white[k] - black[k]) * wb[k];
    }
    o_mat3mulv(rgb_to_yuv, v, yuv);
    o_mat3mulv(rgb_to_yuv, vo, yuvo);
    yuv[0] = o_mix(yuvo[0], yuv[0], p->luma);
    o_mat3mulv(yuv_to_rgb, yuv, rgb);
    float rgba[4];
    for(int k = 0; k < 3; k++) rgba[k] = rgb[k] / wb[k] * (white[k] - black[k]) + black[k];
    rgba[3] = test;
    o_store4(out, x, y, rgba, 1);
  }
}

/* denoise/doub.comp:35-115 */
void o_denoise_doub(const oimg_t *in, const oimg_t *crs0, const oimg_t *crs1, oimg_t *out,
    const o_denoise_params_t *p, const int *crop, const float *black4, const float *white4,
    float noise_a, float noise_b, uint32_t filters)
{
  o_denoise_doub_gm(in, crs0, crs1, out, p, crop, black4, white4, noise_a, noise_b, filters, 0, 0);
}
/* gm != 0: with the gain map branch (doub.comp:106-114; filters != 9 && push.gainmap == 1 && params.gainmap == 1) */
void o_denoise_doub_gm(const oimg_t *in, const oimg_t *crs0, const oimg_t *crs1, oimg_t *out,
    const o_denoise_params_t *p, const int *crop, const float *black4, const float *white4,
    float noise_a, float noise_b, uint32_t filters, const oimg_t *gm, const float *map_os)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float upsm[4], down[4];
    o_tex4(crs0, (x + 0.5) / (double)out->w, (y + 0.5) / (double)out->h, upsm);
    o_tex4(crs1, (x + 0.5) / (double)out->w, (y + 0.5) / (double)out->h, down);
    float black = black4[1], white = white4[1], crs = upsm[1], crs1v = down[1];
    float T = 0.5f * p->strength * upsm[3], blendw = p->luma;
    int col; /* 0 r 1 g 2 b */
    if(filters == 9) col = o_xtrans_colour(x, y);
    else col = ((x & 1) == (y & 1)) ? ((x & 1) ? 2 : 0) : 1;
    if(col != 1)
    {
      black = black4[col]; white = white4[col];
      crs = upsm[col]; crs1v = down[col]; blendw = 1.0f;
      if(filters == 9) T /= o_max(1e-4f, upsm[3]);
    }
    float sigma[3];
    noise_sigma(noise_a, noise_b, black, white, p->edges, crs, sigma);
    float val = o_tex1(in, (x + crop[0] + .5) / (double)in->w, (y + crop[1] + .5) / (double)in->h);
    blendw = 0.5f * (blendw + 1.0f);
    if(val < white)
    {
      const float wav = (val - crs1v) / o_max(sigma[0] + sigma[2], 1e-8f);
      const float tt = o_min(1.0f, wav / o_max(2.0f * T, 1e-8f));
      float uw = powf(o_min(1.0f, 1.0f * upsm[3]), 4.0f);
      uw = 1.0f - (1.0f - uw) * p->detail;
      val = o_mix(val, o_max(0.0f, crs + sigma[1] * o_sign(wav) * o_mix(o_max(0.0f, fabsf(wav) - T), fabsf(wav), tt)), uw * blendw);
    }
    val = o_max(0.0f, (val - black) / (white - black));
    if(gm && filters != 9) val *= o_gainmap_gain(gm, map_os, x, y, crop[0], crop[1], out->w, out->h, 2);
    o_store1(out, x, y, val, 1);
  }
}

/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/llap/{curve,reduce,assemble,colour}.comp, llap.glsl,
 * shared.glsl:98-148 (sample_soft, sample_semisoft), grade/main.comp:21-62 and
 * the pyramid construction of llap/main.c:31-105. */
#include "o_common.h"
#include "vkdt_oracle.h"

#define NUM_GAMMA 10

/* llap/llap.glsl:3-22 */
static inline float gamma_from_i(int i) { return (float)i / (NUM_GAMMA - 1.0f); }
static inline int gamma_hi_from_v(float v)
{
  int hi = 1;
  for(; hi < NUM_GAMMA - 1 && gamma_from_i(hi) <= v; hi++);
  return hi;
}

/* llap/curve.comp:40-63 */
static float curve(float x, float g, float sigma, float shadows, float highlights, float clarity)
{
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? sigma : -sigma;
  const float shadhi = c > 0.0f ? shadows : highlights;
  if(fabsf(c) > 2 * sigma) val = g + ssigma + shadhi * (c - ssigma);
  else
  {
    const float t = o_clamp(c / (2.0f * ssigma), 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  val += clarity * c * expf(-c * c / (2.0f * sigma * sigma / 3.0f));
  return val;
}

/* llap/curve.comp:65-80: out[0..9] remapped, out[10] grey */
void o_llap_curve(const oimg_t *in, oimg_t *out, const o_llap_params_t *p)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < in->h; y++) for(int x = 0; x < in->w; x++)
  {
    float rgb[4];
    o_fetch4(in, x, y, rgb);
    for(int k = 0; k < 3; k++) rgb[k] = o_clamp(rgb[k], -1000.0f, 1000.0f);
    const float l = o_lum2020(rgb);
    for(int i = 0; i < NUM_GAMMA; i++)
      o_store1(out + i, x, y, curve(l, gamma_from_i(i), p->sigma, p->shadows, p->hilights, p->clarity), 1);
    o_store1(out + NUM_GAMMA, x, y, l, 1);
  }
}

/* shared.glsl:130-148 */
static float sample_semisoft(const oimg_t *tex, double u, double v)
{
  const double sx = (double)tex->w, sy = (double)tex->h;
  const double cx = u * sx, cy = v * sy;
  const double x0 = (cx - .5) / sx, x1 = (cx + .5) / sx;
  const double y0 = (cy - .5) / sy, y1 = (cy + .5) / sy;
  float r = 0.0f;
  r += o_tex1(tex, x0, y0);
  r += o_tex1(tex, x1, y0);
  r += o_tex1(tex, x0, y1);
  r += o_tex1(tex, x1, y1);
  return r / 4.0f;
}
/* shared.glsl:99-127 */
static float sample_soft(const oimg_t *tex, double u, double v)
{
  const double sx = (double)tex->w, sy = (double)tex->h;
  const double cx = u * sx, cy = v * sy;
  const double px[3] = { (cx - 1.5) / sx, cx / sx, (cx + 1.5) / sx };
  const double py[3] = { (cy - 1.5) / sy, cy / sy, (cy + 1.5) / sy };
  float r = 0.0f;
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) r += o_tex1(tex, px[i], py[j]);
  return r / 9.0f;
}

/* llap/reduce.comp:16-35, one layer */
void o_llap_reduce(const oimg_t *in, oimg_t *out)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
    o_store1(out, x, y, sample_semisoft(in, (2 * x + 0.5) / (double)in->w, (2 * y + 0.5) / (double)in->h), 1);
}

static inline float gauss_expand(const oimg_t *im, int ox, int oy)
{ /* llap/assemble.comp:19-24 */
  return sample_soft(im, (ox * 0.5 + 0.5) / (double)im->w, (oy * 0.5 + 0.5) / (double)im->h);
}

/* llap/assemble.comp:52-88.  l0/l1: arrays of NUM_GAMMA+1 layers of the fine / coarse level.
 * first != 0: the "coarse" to expand is the grey layer l1[NUM_GAMMA] */
void o_llap_assemble(const oimg_t *coarse, const oimg_t *l0, const oimg_t *l1, oimg_t *out, int first)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    const float res = first ? gauss_expand(l1 + NUM_GAMMA, x, y) : gauss_expand(coarse, x, y);
    const float v = o_fetch1(l0 + NUM_GAMMA, x, y);
    const int hi = gamma_hi_from_v(v);
    const int lo = hi - 1;
    const float glo = gamma_from_i(lo), ghi = gamma_from_i(hi);
    const float a = o_clamp((v - glo) / (ghi - glo), 0.0f, 1.0f);
    const float lap0 = o_fetch1(l0 + lo, x, y) - gauss_expand(l1 + lo, x, y);
    const float lap1 = o_fetch1(l0 + hi, x, y) - gauss_expand(l1 + hi, x, y);
    o_store1(out, x, y, res + lap0 * (1.0f - a) + lap1 * a, 1);
  }
}

/* llap/colour.comp:17-37 */
void o_llap_colour(const oimg_t *lum, const oimg_t *org, oimg_t *out, int out_f16)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float l = o_fetch1(lum, x, y);
    float rgb[4];
    o_fetch4(org, x, y, rgb);
    const float yo = o_max(o_lum2020(rgb), 1e-8f);
    if(l < yo) l = yo * expf(1.0f * (l - yo));
    float o[4];
    for(int k = 0; k < 3; k++) o[k] = o_max(0.0f, rgb[k] * l / yo);
    o[3] = 1.0f;
    o_store4(out, x, y, o, out_f16);
  }
}

/* the whole module as wired by llap/main.c:31-105 */
void o_llap_module(const oimg_t *in, oimg_t *out, const o_llap_params_t *p, int out_f16)
{
  enum { MAXL = 12 };
  oimg_t *lev[MAXL];      /* lev[l] = 11 layers at level l */
  oimg_t asm_out[MAXL];   /* asm_out[l] = output of assemble node l (dims of level l-1) */
  int lw[MAXL], lh[MAXL];
  int nl = MAXL;
  lw[0] = in->w; lh[0] = in->h;
  for(int l = 1; l < MAXL; l++)
  {
    lw[l] = (lw[l-1] - 1) / 2 + 1; lh[l] = (lh[l-1] - 1) / 2 + 1;
    /* main.c:79-88: stop once the *next* level would be <= 1 px */
    const int nw = (lw[l] - 1) / 2 + 1, nh = (lh[l] - 1) / 2 + 1;
    if(nw <= 1 || nh <= 1) { nl = l + 1; break; }
  }
  for(int l = 0; l < nl; l++)
  {
    lev[l] = (oimg_t *)malloc(sizeof(oimg_t) * (NUM_GAMMA + 1));
    for(int i = 0; i <= NUM_GAMMA; i++) lev[l][i] = o_img_alloc(lw[l], lh[l], 1);
  }
  o_llap_curve(in, lev[0], p);
  for(int l = 1; l < nl; l++)
    for(int i = 0; i <= NUM_GAMMA; i++) o_llap_reduce(lev[l-1] + i, lev[l] + i);
  for(int l = nl - 1; l >= 1; l--)
  {
    asm_out[l] = o_img_alloc(lw[l-1], lh[l-1], 1);
    o_llap_assemble(l == nl - 1 ? 0 : &asm_out[l+1], lev[l-1], lev[l], &asm_out[l], l == nl - 1);
  }
  o_llap_colour(&asm_out[1], in, out, out_f16);
  for(int l = 1; l < nl; l++) o_img_free(&asm_out[l]);
  for(int l = 0; l < nl; l++)
  {
    for(int i = 0; i <= NUM_GAMMA; i++) o_img_free(lev[l] + i);
    free(lev[l]);
  }
}

/* grade/main.comp:21-62; q = {lift[4], gamma[4], gain[4], offset[4], mode, sh_pivot, hi_pivot} */
void o_grade_main(const oimg_t *in, oimg_t *out, const o_grade_params_t *q, int out_f16)
{
  float lift[3], gam[3], gain[3], off[3];
  for(int k = 0; k < 3; k++)
  {
    lift[k] = q->lift[k] + q->lift[3];
    gam[k]  = o_max(q->gamma[k] + q->gamma[3], 1e-6f);
    gain[k] = o_max(q->gain[k] + q->gain[3], 0.0f);
    off[k]  = q->offset[k] + q->offset[3];
  }
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgb[4];
    o_fetch4(in, x, y, rgb);
    if(q->mode == 0)
    {
      for(int k = 0; k < 3; k++)
      {
        float v = gain[k] * rgb[k];
        v = v * (1.0f - lift[k]) + lift[k];
        v = powf(o_max(v, 0.0f), 1.0f / gam[k]);
        rgb[k] = v + off[k];
      }
    }
    else
    {
      float L = o_max(rgb[0], 0.0f) * 0.2126f + o_max(rgb[1], 0.0f) * 0.7152f + o_max(rgb[2], 0.0f) * 0.0722f;
      L = o_clamp(0.67f + log2f(o_max(L, 1e-6f)) * 0.11f, 0.0f, 1.0f);
      const float sp = o_clamp(q->sh_pivot, 1e-3f, 1.0f - 1e-3f);
      const float hp = o_clamp(q->hi_pivot, sp + 1e-3f, 1.0f);
      const float w_s = 1.0f - o_smoothstep(0.0f, sp, L);
      const float w_h = o_smoothstep(hp, 1.0f, L);
      const float w_m = 1.0f - w_s - w_h;
      for(int k = 0; k < 3; k++)
      {
        const float ge = o_mix(1.0f, gain[k], w_h);
        const float le = lift[k] * w_s;
        const float me = o_mix(1.0f, gam[k], w_m);
        float v = ge * rgb[k];
        v = v * (1.0f - le) + le;
        v = powf(o_max(v, 0.0f), 1.0f / me);
        rgb[k] = v + off[k];
      }
    }
    rgb[3] = 1.0f;
    o_store4(out, x, y, rgb, out_f16);
  }
}

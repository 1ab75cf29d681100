/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/hilite/{half,reduce,assemble,doub}.comp */
#include "o_common.h"
#include "vkdt_oracle.h"

/* hilite/half.comp:24-72 */
void o_hilite_half(const oimg_t *in, oimg_t *out, const o_hilite_params_t *p, uint32_t filters)
{
  const float white = p->white;
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgba[4];
    if(filters == 9)
    {
      float c[9];
      for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++)
        c[3*i+j] = o_fetch1(in, 3*x + i, 3*y + j);
      if(c[1] >= white) c[1] = c[7];
      if(c[7] >= white) c[7] = c[1];
      if(c[3] >= white) c[3] = c[5];
      if(c[5] >= white) c[5] = c[3];
      const float maxg = o_max(o_max(o_max(c[0], c[2]), c[4]), o_max(c[6], c[8]));
      if(maxg >= white) c[0] = c[2] = c[4] = c[6] = c[8] = 1.0f;
      const float col0 = (c[1] + c[7]) * 0.5f, col1 = (c[3] + c[5]) * .5f;
      if(((x + y) & 1) > 0) { rgba[0] = col0; rgba[2] = col1; }
      else                  { rgba[2] = col0; rgba[0] = col1; }
      rgba[1] = (c[0] + c[2] + c[4] + c[6] + c[8]) / 5.0f;
      rgba[3] = 1.0f;
    }
    else
    {
      float c[4];
      o_gather(in, 2.0 * (x + .5) / (double)in->w, 2.0 * (y + .5) / (double)in->h, c);
      if(c[0] >= white) c[0] = c[2];
      if(c[2] >= white) c[2] = c[0];
      rgba[0] = c[3]; rgba[1] = (c[0] + c[2]) / 2.0f; rgba[2] = c[1]; rgba[3] = 1.0f;
    }
    o_store4(out, x, y, rgba, 1);
  }
}

/* hilite/reduce.comp:21-60 */
void o_hilite_reduce(const oimg_t *in, oimg_t *out, const o_hilite_params_t *p, const float *wb4)
{
  float white = p->white;
  if(!(white > 0.0f)) white = 1.0f;
  static const float w[5] = {1.0f/16.0f, 4.0f/16.0f, 6.0f/16.0f, 4.0f/16.0f, 1.0f/16.0f};
  static const float sw[5] = {1.0f, 2.0f, 0.0f, -2.0f, -1.0f};
  float wb[3] = { wb4[0], wb4[1], wb4[2] };
  if(!(wb[0]*wb[0] + wb[1]*wb[1] + wb[2]*wb[2] > 1e-3f)) wb[0] = wb[1] = wb[2] = 1.0f;
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float edge[2] = {0, 0}, col[3] = {0, 0, 0}, wgt = 0.0f;
    for(int jj = -2; jj <= 2; jj++) for(int ii = -2; ii <= 2; ii++)
    {
      float rgb[4];
      o_tex4(in, (double)(2*x + ii + 0.5) / (double)in->w, (double)(2*y + jj + 0.5) / (double)in->h, rgb);
      const float l = o_lum2020(rgb);
      edge[0] += w[jj+2] * sw[ii+2] * l;
      edge[1] += w[ii+2] * sw[jj+2] * l;
      const float u = w[ii+2] * w[jj+2];
      const float rw = rgb[0] * wb[0], gw = rgb[1] * wb[1], bw = rgb[2] * wb[2];
      const float cmax = o_max(rw, o_max(gw, bw));
      const float cmin = o_min(rw, o_min(gw, bw));
      const float sat = (cmax - cmin) / o_max(1e-3f, cmax);
      if(rgb[0] < white && rgb[1] < white && rgb[2] < white)
      {
        const float s = o_smoothstep(0.2f, 1.0f, o_max(rgb[0], o_max(rgb[1], rgb[2])) / white);
        float t = o_smoothstep(0.15f, 0.9f, sat);
        const float ds = p->desat * p->desat;
        t = o_clamp(5.0f * ds * ds * s * t, 0.0f, 1.0f);
        for(int k = 0; k < 3; k++)
        {
          const float c = o_mix(rgb[k], (cmax + cmin) / wb[k] * .5f, t);
          col[k] += c * u;
        }
        wgt += u;
      }
    }
    float o[4];
    if(wgt == 0.0f) o[0] = o[1] = o[2] = 1.0f;
    else for(int k = 0; k < 3; k++) o[k] = col[k] / wgt;
    o[3] = sqrtf(edge[0]*edge[0] + edge[1]*edge[1]);
    o_store4(out, x, y, o, 1);
  }
}

/* hilite/assemble.comp:23-86 `gauss_expand` */
static void gauss_expand(const oimg_t *im, int ox, int oy, float *c)
{
  static const float w[5] = {1.0f/16.0f, 4.0f/16.0f, 6.0f/16.0f, 4.0f/16.0f, 1.0f/16.0f};
  const int ix = ox / 2, iy = oy / 2;
  const int dx = ox & 1, dy = oy & 1;
  /* even: taps -1..1 with w[2*t+2]; odd: taps 0..1 with w[2*t+1] */
  const int i0 = dx ? 0 : -1, j0 = dy ? 0 : -1;
  float wgt = 0.0f;
  c[0] = c[1] = c[2] = 0.0f;
  for(int ii = i0; ii <= 1; ii++) for(int jj = j0; jj <= 1; jj++)
  {
    float rgb[4];
    o_tex4(im, (double)(ix + ii + 0.5) / (double)im->w, (double)(iy + jj + 0.5) / (double)im->h, rgb);
    const float wy = dy ? w[2*jj+1] : w[2*jj+2];
    const float wx = dx ? w[2*ii+1] : w[2*ii+2];
    for(int k = 0; k < 3; k++) c[k] += rgb[k] * wy * wx;
    wgt += wy * wx;
  }
  if(wgt == 0.0f) { c[0] = 0.0f; c[1] = 1.0f; c[2] = 1.0f; return; }
  for(int k = 0; k < 3; k++) c[k] /= wgt;
}

/* hilite/assemble.comp:88-118 */
void o_hilite_assemble(const oimg_t *fine_img, const oimg_t *coarse, oimg_t *out, const o_hilite_params_t *p)
{
  const float white = p->white;
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float upsm[3], fine[4];
    gauss_expand(coarse, x, y, upsm);
    o_fetch4(fine_img, x, y, fine);
    const float sr = fine[0] / o_max(0.001f, upsm[0]);
    const float sg = fine[1] / o_max(0.001f, upsm[1]);
    const float sb = fine[2] / o_max(0.001f, upsm[2]);
    const float wr = expf(upsm[0] - o_max(upsm[1], upsm[2]));
    const float wg = expf(upsm[1] - o_max(upsm[0], upsm[2]));
    const float wb = expf(upsm[2] - o_max(upsm[0], upsm[1]));
    const float scale = (sr * wr + sg * wg + sb * wb) / (wr + wg + wb);
    float t = p->soft;
    if(fine[0] >= white || fine[1] >= white || fine[2] >= white) t = 1.0f;
    if(isnan(fine[3])) fine[3] = 0.0f;
    t = o_clamp(o_mix(t, 1.0f, sqrtf(o_max(0.0f, fine[3]))), 0.0f, 1.0f);
    float o[4];
    for(int k = 0; k < 3; k++)
    {
      const float rec = o_clamp(upsm[k] * scale, -65535.0f, 65535.0f);
      o[k] = o_mix(fine[k], rec, t);
    }
    o[3] = 1.0f;
    o_store4(out, x, y, o, 1);
  }
}

/* hilite/doub.comp:22-121.  runs on the coarse (block) grid, writes block x block mosaic texels */
void o_hilite_doub(const oimg_t *in, const oimg_t *coarse, oimg_t *out, const o_hilite_params_t *p, uint32_t filters)
{
  const float white = p->white;
  const int bw = filters == 9 ? out->w / 3 : out->w / 2, bh = filters == 9 ? out->h / 3 : out->h / 2;
#pragma omp parallel for schedule(static)
  for(int y = 0; y < bh; y++) for(int x = 0; x < bw; x++)
  {
    float upsm[4];
    o_fetch4(coarse, x, y, upsm);
    if(filters == 9)
    {
      if(((x + y) & 1) == 0) { const float t = upsm[0]; upsm[0] = upsm[2]; upsm[2] = t; }
      float c[9];
      for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++)
        c[3*i+j] = o_fetch1(in, 3*x + i, 3*y + j);
      const float minr = o_min(c[1], c[7]);
      const float minb = o_min(c[3], c[5]);
      const float ming = o_min(o_min(o_min(c[0], c[2]), c[4]), o_min(c[6], c[8]));
      const float sr = minr / o_max(0.001f, upsm[0]);
      const float sg = ming / o_max(0.001f, upsm[1]);
      const float sb = minb / o_max(0.001f, upsm[2]);
      const float wr = expf(upsm[0] - o_max(upsm[1], upsm[2]));
      const float wg = expf(upsm[1] - o_max(upsm[0], upsm[2]));
      const float wb = expf(upsm[2] - o_max(upsm[0], upsm[1]));
      const float scale = (sr * wr + sg * wg + sb * wb) / (wr + wg + wb);
      const float maxr = o_max(c[1], c[7]);
      const float maxb = o_max(c[3], c[5]);
      const float maxg = o_max(o_max(o_max(c[0], c[2]), c[4]), o_max(c[6], c[8]));
      const float softw = 0.97f * white;
      const float maxrgb = o_max(maxr, o_max(maxg, maxb));
      if(maxrgb > softw)
      {
        float t = o_smoothstep(softw, white, maxrgb);
        c[1] = o_mix(c[1], upsm[0] * scale, t);
        c[7] = o_mix(c[7], upsm[0] * scale, t);
        c[3] = o_mix(c[3], upsm[2] * scale, t);
        c[5] = o_mix(c[5], upsm[2] * scale, t);
        t = o_smoothstep(softw, white, ming);
        c[0] = o_mix(c[0], upsm[1] * scale, t);
        c[2] = o_mix(c[2], upsm[1] * scale, t);
        c[4] = o_mix(c[4], upsm[1] * scale, t);
        c[6] = o_mix(c[6], upsm[1] * scale, t);
        c[8] = o_mix(c[8], upsm[1] * scale, t);
      }
      for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++)
        o_store1(out, 3*x + i, 3*y + j, c[3*i+j], 1);
    }
    else
    {
      float c[4];
      o_gather(in, 2.0 * (x + .5) / (double)in->w, 2.0 * (y + .5) / (double)in->h, c);
      const float ming = o_min(c[0], c[2]);
      const float sr = c[3] / o_max(0.001f, upsm[0]);
      const float sg = ming / o_max(0.001f, upsm[1]);
      const float sb = c[1] / o_max(0.001f, upsm[2]);
      const float wr = expf(upsm[0] - o_max(upsm[1], upsm[2]));
      const float wg = expf(upsm[1] - o_max(upsm[0], upsm[2]));
      const float wb = expf(upsm[2] - o_max(upsm[0], upsm[1]));
      const float scale = (sr * wr + sg * wg + sb * wb) / (wr + wg + wb);
      const float softw = 0.97f * white;
      const float maxrgb = o_max(o_max(c[0], c[1]), o_max(c[2], c[3]));
      if(maxrgb > softw)
      {
        const float t = o_smoothstep(softw, white, maxrgb);
        const float u[4] = { upsm[1], upsm[2], upsm[1], upsm[0] }; /* .gbgr */
        for(int k = 0; k < 4; k++) c[k] = o_mix(c[k], scale * u[k], t);
      }
      o_store1(out, 2*x,     2*y,     c[3], 1);
      o_store1(out, 2*x + 1, 2*y,     c[2], 1);
      o_store1(out, 2*x,     2*y + 1, c[0], 1);
      o_store1(out, 2*x + 1, 2*y + 1, c[1], 1);
    }
  }
}

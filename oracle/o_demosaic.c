/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/demosaic/{down,gauss,splat,fix,halfsize}.comp.
 * `down` is kept for completeness; `gauss` never samples its output (gauss.comp:12, only img_orig is read). */
#include "o_common.h"
#include "vkdt_oracle.h"

/* demosaic/down.comp:17-47 */
void o_demosaic_down(const oimg_t *in, oimg_t *out, uint32_t filters)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float lum;
    if(filters == 9)
    {
      float s = 0.0f; /* c0..c8 summed in the shader's order: (0,0)(0,1)(0,2)(1,0).. */
      for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) s += o_fetch1(in, 3*x + i, 3*y + j);
      lum = s / 9.0f;
    }
    else
    {
      const float c0 = o_fetch1(in, 2*x, 2*y), c1 = o_fetch1(in, 2*x, 2*y + 1);
      const float c2 = o_fetch1(in, 2*x + 1, 2*y), c3 = o_fetch1(in, 2*x + 1, 2*y + 1);
      lum = (c0 + c1 + c2 + c3) / 4.0f;
    }
    o_store1(out, x, y, lum, 1);
  }
}

/* demosaic/halfsize.comp:15-52 */
void o_demosaic_halfsize(const oimg_t *in, oimg_t *out, uint32_t filters)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgba[4];
    if(filters == 9)
    {
      float c[9];
      for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) c[3*i+j] = o_fetch1(in, 3*x + i, 3*y + j);
      rgba[1] = (c[0] + c[2] + c[4] + c[6] + c[8]) * 1.0f / 5.0f;
      const float col0 = (c[1] + c[7]) * 0.5f, col1 = (c[3] + c[5]) * .5f;
      if(((x + y) & 1) > 0) { rgba[0] = col0; rgba[2] = col1; }
      else                  { rgba[2] = col0; rgba[0] = col1; }
      rgba[3] = 1.0f; /* left undefined by the shader */
    }
    else
    {
      float c[4];
      o_gather(in, 2.0 * (x + .5) / (double)in->w, 2.0 * (y + .5) / (double)in->h, c);
      rgba[0] = c[3]; rgba[1] = (c[0] + c[2]) / 2.0f; rgba[2] = c[1]; rgba[3] = 1.0f;
    }
    o_store4(out, x, y, rgba, 1);
  }
}

/* shared/resample.comp:26-40 (returns after the catmull rom lookup) */
void o_resample(const oimg_t *in, oimg_t *out)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float res[4];
    o_sample_catmull_rom(in, ((float)x + 0.5f) / (float)out->w, ((float)y + 0.5f) / (float)out->h, res);
    o_store4(out, x, y, res, 1);
  }
}

/* demosaic/gauss.comp:17-126 */
void o_demosaic_gauss(const oimg_t *orig, oimg_t *out, uint32_t filters)
{
  const int xt = filters == 9;
  const int blk = xt ? 3 : 2;
  const int lo = xt ? 0 : -1, hi = 3;
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float Sw[4] = {0}, Sb[4] = {0}, sw = 0.0f, sb = 0.0f;
    float mw[2] = {0}, mb[2] = {0}, smw = 0.0f, smb = 0.0f;
    for(int j = lo; j < hi; j++) for(int i = lo; i < hi; i++)
    {
      if(xt ? (((j + i) & 1) == 1) : (((j + i) & 1) != 1)) continue;
      const float px = o_tex1(orig, (blk * x + 0.5 + (double)i) / (double)orig->w, (blk * y + 0.5 + (double)j) / (double)orig->h);
      mw[0] += (float)i * px; mw[1] += (float)j * px;
      smw += px;
      mb[0] += (float)i / px; mb[1] += (float)j / px;
      smb += 1.0f / px;
    }
    mw[0] /= smw; mw[1] /= smw;
    mb[0] /= smb; mb[1] /= smb;
    for(int j = lo; j < hi; j++) for(int i = lo; i < hi; i++)
    {
      if(xt ? (((j + i) & 1) == 1) : (((j + i) & 1) != 1)) continue;
      const float px = o_tex1(orig, (blk * x + 0.5 + (double)i) / (double)orig->w, (blk * y + 0.5 + (double)j) / (double)orig->h);
      float p2 = px * px;
      float p0 = (float)i - mw[0], p1 = (float)j - mw[1];
      Sw[0] += p2 * p0 * p0; Sw[1] += p2 * p0 * p1;
      Sw[2] += p2 * p1 * p0; Sw[3] += p2 * p1 * p1;
      sw += p2;
      p0 = (float)i - mb[0]; p1 = (float)j - mb[1];
      p2 = 1.0f / p2;
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
    if(!xt)
    {
      eval[0] *= 0.2f; eval[1] *= 0.2f;
      if(fabsf(evec0[0]) > fabsf(evec0[1])) { evec0[0] = 1; evec0[1] = 0; }
      else                                  { evec0[0] = 0; evec0[1] = 1; }
    }
    else
    {
      eval[0] *= 0.4f; eval[1] *= 0.4f;
      if     (fabsf(evec0[0]) > 2.f * fabsf(evec0[1])) { evec0[0] = 1; evec0[1] = 0; }
      else if(fabsf(evec0[1]) > 2.f * fabsf(evec0[0])) { evec0[0] = 0; evec0[1] = 1; }
      else eval[0] = eval[1] = .1f;
    }
    const float o[4] = { eval[0], eval[1], evec0[0], evec0[1] };
    o_store4(out, x, y, o, 1);
  }
}

/* demosaic/splat.comp:16-140 */
void o_demosaic_splat(const oimg_t *in, const oimg_t *gauss, oimg_t *out, uint32_t filters)
{
  const int xt = filters == 9;
  const int r = xt ? 2 : 1;
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float cov[4];
    if(xt) o_fetch4(gauss, x / 3, y / 3, cov);
    else   o_fetch4(gauss, (x + 1) / 2, (y + 1) / 2, cov);
    float rgb[3] = {0}, w[3] = {0};
    const float e0 = o_clamp(cov[0], 0.01f, 25.0f), e1 = o_clamp(cov[1], 0.01f, 25.0f);
    for(int j = -r; j <= r; j++) for(int i = -r; i <= r; i++)
    {
      int px = x + i, py = y + j;
      if(px < 0) px += 6;
      if(py < 0) py += 6;
      if(px >= in->w) px -= 6;
      if(py >= in->h) py -= 6;
      float col = o_fetch1(in, px, py);
      /* E = mat2(cov.z, -cov.w, cov.w, cov.z) given as columns: of = E * o */
      const float of0 = cov[2] * (float)i + cov[3] * (float)j;
      const float of1 = -cov[3] * (float)i + cov[2] * (float)j;
      float weight = o_clamp(expf(-0.5f * (of0 / e0 * of0 + of1 / e1 * of1)), 1e-4f, 1.0f);
      if(i == 0 && j == 0) weight = 666.0f;
      int c;
      if(xt) c = o_xtrans_colour(x + i + 6, y + j + 6);
      else
      {
        const int qx = x + i, qy = y + j; /* may be -1: & 1 on two's complement is what glsl does */
        c = ((qx & 1) == (qy & 1)) ? ((qx & 1) ? 2 : 0) : 1;
      }
      if(xt) { rgb[c] += col * weight; w[c] += weight; }
      else   { col *= weight; rgb[c] += col; w[c] += weight; }
    }
    o_store1(out, x, y, rgb[1] / o_max(1e-8f, w[1]), 1);
  }
}

static inline float fix_gauss(float e0, float e1, float cz, float cw, int i, int j)
{
  const float of0 = cz * (float)i + cw * (float)j;
  const float of1 = -cw * (float)i + cz * (float)j;
  return o_clamp(expf(-0.5f * (of0 / e0 * of0 + of1 / e1 * of1)), 1e-3f, 1.0f);
}

/* demosaic/fix.comp:25-135.  `fixup` is param 0 of the module (`colour`), fix.comp:7-10 */
void o_demosaic_fix(const oimg_t *in, const oimg_t *green, const oimg_t *covimg, oimg_t *out, uint32_t filters, int fixup)
{
  const int xt = filters == 9;
  const int r = xt ? o_clampi(fixup + 2, 2, 3) : o_clampi(fixup + 1, 1, 2);
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgb[3] = {0}, w[3] = {0};
    const float gc = o_tex1(green, (x + 0.5) / (double)green->w, (y + 0.5) / (double)green->h);
    float cov[4];
    if(xt) { o_fetch4(covimg, (x + 1) / 3, (y + 1) / 3, cov); cov[0] = o_clamp(cov[0], 1.f, 10.f); cov[1] = o_clamp(cov[1], 1.f, 10.f); }
    else   { o_fetch4(covimg, (x + 1) / 2, (y + 1) / 2, cov); cov[0] = o_clamp(cov[0], 1.0f, 49.f); cov[1] = o_clamp(cov[1], 1.0f, 49.f); }
    const float ks = xt ? 3.0f : 2.0f;
    for(int j = -r; j <= r; j++) for(int i = -r; i <= r; i++)
    {
      const float gh  = o_tex1(green, (x + i + 0.5) / (double)green->w, (y + j + 0.5) / (double)green->h);
      const float col = o_tex1(in,    (x + i + 0.5) / (double)in->w,    (y + j + 0.5) / (double)in->h);
      const int px = x + i, py = y + j;
      int c;
      if(xt) c = o_xtrans_colour(px, py); /* no +6 margin in fix.comp:38-40; int division truncates toward zero for negatives */
      else   c = ((px & 1) == (py & 1)) ? ((px & 1) ? 2 : 0) : 1;
      if(c == 1) { rgb[1] = gc; w[1] = 1.0f; }
      else
      {
        const float weight = fix_gauss(ks * cov[0], ks * cov[1], cov[2], cov[3], i, j);
        if(xt) { const float corr = (1e-4f + gc) / (1e-4f + gh); rgb[c] += col * corr * weight; }
        else   rgb[c] += col * (1e-4f + gc) / (1e-4f + gh) * weight;
        w[c] += weight;
      }
    }
    float o[4];
    for(int k = 0; k < 3; k++) o[k] = rgb[k] / o_max(1e-8f, w[k]);
    o[3] = 1.0f;
    o_store4(out, x, y, o, 1);
  }
}

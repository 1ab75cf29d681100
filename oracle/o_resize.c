/* ORACLE — test infrastructure, not product.  CPU restatement of the export side's resize module:
 *   resize/main.comp (:22-35: nearest slice / sample_flower / sample_catmull_rom by push.scale),
 *   shared/blurh.comp, shared/blurv.comp (:7-66: separable gaussian, weights by the w_i = w_{i-1} v_i recurrence),
 *   shared.glsl:199-221 (sample_flower), shared.glsl:47-96 (sample_catmull_rom, o_crop.c).
 * sampling is the ideal sampler of o_common.h (coordinates in double). */
#include "o_common.h"
#include "vkdt_oracle.h"

void o_sample_catmull_rom(const oimg_t *tex, float u, float v, float *res);

/* shared.glsl:199-221: five taps, rgb weighted plainly, the fourth channel carries the weighted max(r,g,b)^2 */
static void sample_flower(const oimg_t *tex, double tcx, double tcy, float *res)
{
  const double sx = (double)tex->w, sy = (double)tex->h;
  const float t = 36.0f / 256.0f;
  const float wq = (1.0f - t) / 4.0f;
  const double ox[5] = { 0.0, (double)1.2f, -(double)1.2f, -(double)0.4f, (double)0.4f };
  const double oy[5] = { 0.0, (double)0.4f, -(double)0.4f, (double)1.2f, -(double)1.2f };
  res[0] = res[1] = res[2] = res[3] = 0.0f;
  for(int k = 0; k < 5; k++)
  {
    float v[4];
    o_tex4(tex, (tcx + ox[k]) / sx, (tcy + oy[k]) / sy, v);
    const float W = k ? wq : t;
    const float l = o_max(v[0], o_max(v[1], v[2]));
    res[0] += W * v[0]; res[1] += W * v[1]; res[2] += W * v[2]; res[3] += W * l * l;
  }
}

/* resize/main.comp:22-35.  mode: 0 magnify, 1 slice, 2 minify */
void o_resize_main(const oimg_t *in, oimg_t *out, int mode, int out_f16)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgb[4];
    if(mode == 1)
    { /* texelFetch(img_in, ivec2(textureSize * (ipos + 0.5) / vec2(imageSize(img_out))), 0) */
      const int fx = (int)((float)in->w * ((float)x + 0.5f) / (float)out->w);
      const int fy = (int)((float)in->h * ((float)y + 0.5f) / (float)out->h);
      o_fetch4(in, fx, fy, rgb);
    }
    else if(mode < 1) sample_flower(in, ((double)x + 0.5) / (double)out->w * (double)in->w, ((double)y + 0.5) / (double)out->h * (double)in->h, rgb);
    else o_sample_catmull_rom(in, ((float)x + 0.5f) / (float)out->w, ((float)y + 0.5f) / (float)out->h, rgb);
    o_store4(out, x, y, rgb, out_f16);
  }
}

/* shared/blurh.comp / blurv.comp:7-66 */
void o_blur_sep(const oimg_t *in, oimg_t *out, float radius, int vertical, int out_f16)
{
  const int sp = (int)floorf(radius);
  const float sigma = radius / 3.0f;
  const float a = 0.5f / (sigma * sigma);
  const float c = expf(-2.0f * a), w1 = expf(-a), v2 = expf(-3.0f * a);
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    /* the taps sit on texel centres (uv +- i / size): the ideal sampler fetches the mirrored texel */
    float color[4], wgt = 1.0f;
    const double u = ((double)x + 0.5) / (double)in->w, vv = ((double)y + 0.5) / (double)in->h;
    o_tex4(in, u, vv, color);
    if(sp > 0)
    {
      float w = w1, v = v2;
      int i = 1;
      for(; i <= sp - 1; i += 2)
      {
        float p1[4], m1[4], p2[4], m2[4];
        const double d1 = (double)i, d2 = (double)(i + 1);
        if(vertical)
        {
          o_tex4(in, u, vv + d1 / (double)in->h, p1); o_tex4(in, u, vv - d1 / (double)in->h, m1);
          o_tex4(in, u, vv + d2 / (double)in->h, p2); o_tex4(in, u, vv - d2 / (double)in->h, m2);
        }
        else
        {
          o_tex4(in, u + d1 / (double)in->w, vv, p1); o_tex4(in, u - d1 / (double)in->w, vv, m1);
          o_tex4(in, u + d2 / (double)in->w, vv, p2); o_tex4(in, u - d2 / (double)in->w, vv, m2);
        }
        const float w2 = w * v;
        const float vn = v * c;
        wgt += 2.0f * (w + w2);
        for(int k = 0; k < 4; k++) color[k] += w * (p1[k] + m1[k]) + w2 * (p2[k] + m2[k]);
        w = w2 * vn;
        v = vn * c;
      }
      if(i == sp)
      {
        float p1[4], m1[4];
        const double d1 = (double)i;
        if(vertical) { o_tex4(in, u, vv + d1 / (double)in->h, p1); o_tex4(in, u, vv - d1 / (double)in->h, m1); }
        else         { o_tex4(in, u + d1 / (double)in->w, vv, p1); o_tex4(in, u - d1 / (double)in->w, vv, m1); }
        wgt += 2.0f * w;
        for(int k = 0; k < 4; k++) color[k] += w * (p1[k] + m1[k]);
      }
    }
    const float iw = 1.0f / wgt;
    for(int k = 0; k < 4; k++) color[k] *= iw;
    o_store4(out, x, y, color, out_f16);
  }
}

/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/filmcurv/main.comp:18-166, params.glsl:15-34,
 * shared.glsl:371-387 (adjust_colour_dng) and colourspaces.glsl:20-78 (oklab, "hsv").
 * colour mode 2: shared/munsell.glsl:7-137 (hue constancy along munsell hue lines) + colourspaces.glsl:2-19; its table of
 * chromaticities is data shared with the product (munsell_table.h, made by scripts/make_munsell_table.py). */
#include "o_common.h"
#include "vkdt_oracle.h"
#include "../vkdt_b200/csrc/kernels/munsell_table.h"

static const float M_2020_to_xyz[9] = {0.636958048301290991f, 0.144616903586208406f, 0.168880975164172054f, 0.26270021201126692f, 0.677998071518871148f, 0.0593017164698619384f, 4.9999999999999999e-17f, 0.0280726930490874452f, 1.06098505771079066f};
static const float M_xyz_to_2020[9] = {1.71665119f, -0.35567078f, -0.25336628f, -0.66668435f, 1.61648124f, 0.01576855f, 0.01763986f, -0.04277061f, 0.94210312f};

static inline float weibull_cdf(float x, float il, float k) { return 1.0f - expf(-powf(o_max(x, 1e-7f) * il, k)); }
static inline float weibull_pdf(float x, float il, float k)
{
  x = o_max(x, 1e-7f);
  return k * il * powf(x * il, k - 1.0f) * expf(-powf(x * il, k));
}
static inline float glsl_mod(float x, float y) { return x - y * floorf(x / y); }
static inline float glsl_fract(float x) { return x - floorf(x); }

/* shared.glsl:371-387 */
void o_adjust_colour_dng(const float *col0_, float *col1)
{
  float col0[3] = { col0_[0], col0_[1], col0_[2] };
  int fx = 0, fy = 0, fz = 0; float t;
#define SWAP(a, b) do { t = a; a = b; b = t; } while(0)
  if(col0[2] > col0[1]) { SWAP(col0[2], col0[1]); SWAP(col1[2], col1[1]); fx = 1; }
  if(col0[1] > col0[0]) { SWAP(col0[0], col0[1]); SWAP(col1[0], col1[1]); fy = 1; }
  if(col0[2] > col0[1]) { SWAP(col0[2], col0[1]); SWAP(col1[2], col1[1]); fz = 1; }
  col1[1] = o_mix(col1[2], col1[0], (col0[1] - col0[2] + 1e-6f) / (col0[0] - col0[2] + 1e-6f));
  if(fz) SWAP(col1[2], col1[1]);
  if(fy) SWAP(col1[0], col1[1]);
  if(fx) SWAP(col1[2], col1[1]);
#undef SWAP
}

/* colourspaces.glsl:20-49.  glsl mat3(a,b,c, d,e,f, g,h,i) lists COLUMNS */
static void rec2020_to_oklab(const float *rgb, float *lab)
{
  float lms[3];
  lms[0] = 0.61668844f * rgb[0] + 0.36015907f * rgb[1] + 0.02304329f * rgb[2];
  lms[1] = 0.2651402f  * rgb[0] + 0.63585648f * rgb[1] + 0.09903023f * rgb[2];
  lms[2] = 0.10015065f * rgb[0] + 0.20400432f * rgb[1] + 0.69632468f * rgb[2];
  for(int k = 0; k < 3; k++) lms[k] = powf(o_max(0.0f, lms[k]), 1.0f / 3.0f);
  lab[0] = 0.21045426f * lms[0] + 0.79361779f * lms[1] - 0.00407205f * lms[2];
  lab[1] = 1.9779985f  * lms[0] - 2.42859221f * lms[1] + 0.45059371f * lms[2];
  lab[2] = 0.02590404f * lms[0] + 0.78277177f * lms[1] - 0.80867577f * lms[2];
}
static void oklab_to_rec2020(const float *lab, float *rgb)
{
  float lms[3];
  lms[0] = 1.0f        * lab[0] + 0.39633779f * lab[1] + 0.21580376f * lab[2];
  lms[1] = 1.00000001f * lab[0] - 0.10556134f * lab[1] - 0.06385417f * lab[2];
  lms[2] = 1.00000005f * lab[0] - 0.08948418f * lab[1] - 1.29148554f * lab[2];
  for(int k = 0; k < 3; k++) lms[k] = lms[k] * lms[k] * lms[k];
  rgb[0] =  2.14014041f * lms[0] - 1.24635595f * lms[1] + 0.10643173f * lms[2];
  rgb[1] = -0.88483245f * lms[0] + 2.16317272f * lms[1] - 0.27836159f * lms[2];
  rgb[2] = -0.04857906f * lms[0] - 0.45449091f * lms[1] + 1.50235629f * lms[2];
}
static void rgb2hsv(const float *c, float *hsv)
{
  float lab[3];
  rec2020_to_oklab(c, lab);
  hsv[0] = glsl_fract(1.0f + atan2f(lab[2], lab[1]) / (2.0f * (float)M_PI));
  hsv[1] = sqrtf(lab[1] * lab[1] + lab[2] * lab[2]);
  hsv[2] = lab[0];
}
static void hsv2rgb(const float *hCL, float *rgb)
{
  const float lab[3] = { hCL[2], hCL[1] * cosf(2.0f * (float)M_PI * hCL[0]), hCL[1] * sinf(2.0f * (float)M_PI * hCL[0]) };
  if(lab[0] <= 0.0f) { rgb[0] = rgb[1] = rgb[2] = 0.0f; return; }
  oklab_to_rec2020(lab, rgb);
}
static float lerp_chromaticity_angle(float h1, float h2, float t)
{
  const float delta = h2 - h1;
  if(delta > 0.5f) h2 -= 1.0f;
  else if(delta < -0.5f) h2 += 1.0f;
  const float lerped = h1 + t * (h2 - h1);
  return glsl_mod(lerped, 1.0f);
}
static float hue_bump(float h, float h0, float w)
{
  const float pi = (float)M_PI;
  const float d = fabsf(glsl_mod(h - h0 + pi, 2.0f * pi) - pi);
  return d < w ? 0.5f + 0.5f * cosf(pi * d / w) : 0.0f;
}

/* ---- shared/munsell.glsl ---- */
static const uint32_t munsell_xy[VKB_MUNSELL_HDIM * VKB_MUNSELL_CDIM] = { VKB_MUNSELL_WORDS };
#define MH VKB_MUNSELL_HDIM
#define MC VKB_MUNSELL_CDIM
static const float mun_ill[2] = { 0.31271f, 0.32902f };
/* :7-16 */
static float xy_to_monotone_hue_angle(const float *xy)
{
  const float pi = (float)M_PI;
  return glsl_mod(2.0f * pi - 2.52f - atan2f(xy[1] - mun_ill[1], xy[0] - mun_ill[0]), 2.0f * pi);
}
/* :19-27 */
static void munsell_lookup(int hue_idx, int chroma_idx, float *xy)
{
  hue_idx = (hue_idx % MH + MH) % MH;
  chroma_idx = chroma_idx < 0 ? 0 : chroma_idx > MC - 1 ? MC - 1 : chroma_idx;
  const uint32_t w = munsell_xy[MC * hue_idx + chroma_idx];
  xy[0] = o_f16_bits_to_f32((uint16_t)(w & 0xffffu));
  xy[1] = o_f16_bits_to_f32((uint16_t)(w >> 16));
}
/* :30-38 */
static float munsell_side(const float *v0, const float *v1, const float *p)
{
  const float ax = v1[0] - v0[0], ay = v1[1] - v0[1], bx = p[0] - v0[0], by = p[1] - v0[1];
  return ax * by - ay * bx;
}
/* :41-62 */
static void munsell_to_xy(const float *mhc, float *xy)
{
  const float hm = mhc[0] * MH, cm = o_max(mhc[1], 0.0f) * MC;
  const int hidxm = (int)hm, cidxm = (int)cm;
  const float hu = hm - hidxm, cu = cm - cidxm;
  float r0[2], r1[2], r2[2], r3[2];
  munsell_lookup(hidxm, cidxm + 1, r3); munsell_lookup(hidxm + 1, cidxm + 1, r2);
  munsell_lookup(hidxm, cidxm, r0);     munsell_lookup(hidxm + 1, cidxm, r1);
  for(int c = 0; c < 2; c++)
    xy[c] = hu >= cu ? (1 - hu) * r0[c] + (hu - cu) * r1[c] + cu * r2[c]
                     : hu * r2[c] + (cu - hu) * r3[c] + (1 - cu) * r0[c];
}
/* :67-137 */
static void munsell_from_xy(const float *xy, float *mhc)
{
  int hidxm = 0, hidxM = MH, cidxm = 0, cidxM = MC - 1;
  const float theta = xy_to_monotone_hue_angle(xy);
  const float dx = xy[0] - mun_ill[0], dy = xy[1] - mun_ill[1];
  const float rad2 = dx * dx + dy * dy;
  for(int i = 0; i < 10; i++)
  {
    const int hidx = (hidxm + hidxM) / 2, cidx = (cidxm + cidxM) / 2;
    float res[2];
    munsell_lookup(hidx, cidx, res);
    const float th = xy_to_monotone_hue_angle(res);
    const float ex = res[0] - mun_ill[0], ey = res[1] - mun_ill[1];
    const float r2 = ex * ex + ey * ey;
    if(th <= theta) hidxm = hidx; else hidxM = hidx;
    if(r2 <= rad2)  cidxm = cidx; else cidxM = cidx;
    if(hidxM <= hidxm + 1 && cidxM <= cidxm + 1) break;
  }
  for(int i = 0; i < 10; i++)
  {
    float r0[2], r1[2], r2[2], r3[2];
    munsell_lookup(hidxm, cidxm + 1, r3); munsell_lookup(hidxm + 1, cidxm + 1, r2);
    munsell_lookup(hidxm, cidxm, r0);     munsell_lookup(hidxm + 1, cidxm, r1);
    const float s0 = munsell_side(r0, r1, xy), s1 = munsell_side(r1, r2, xy);
    const float s2 = munsell_side(r2, r3, xy), s3 = munsell_side(r3, r0, xy);
    /* the indices step before the containment test, and the result below uses the stepped ones, as the shader does */
    if(s0 < 0 && cidxm > 0) cidxm--;
    else if(s0 < 0 && cidxm == 0) hidxm = ((hidxm + MH / 2) % MH + MH) % MH;
    else if(s2 < 0 && cidxm < MC - 2) cidxm++;
    if(s1 < 0) hidxm++;
    else if(s3 < 0) hidxm--;
    if(s0 >= 0 && s1 >= 0 && s3 >= 0 && (s2 >= 0 || cidxm >= MC - 2))
    {
      const float t0 = munsell_side(r0, r1, r2), t1 = munsell_side(r2, r3, r0);
      float u0, u1, u2, u3;
      if(cidxm > 0 && s0 + s1 <= t0) { u2 = s0 / t0; u0 = s1 / t0; u1 = 1.0f - u0 - u2; u3 = 0.0f; }
      else                           { u2 = s3 / t1; u0 = s2 / t1; u3 = 1.0f - u0 - u2; u1 = 0.0f; }
      const float hi = u0 * hidxm + u1 * (hidxm + 1.0f) + u2 * (hidxm + 1.0f) + u3 * hidxm;
      const float ci = u0 * cidxm + u1 * cidxm + u2 * (cidxm + 1.0f) + u3 * (cidxm + 1.0f);
      mhc[0] = hi / MH; mhc[1] = o_max(0.0f, ci / MC);
      return;
    }
  }
  mhc[0] = mhc[1] = 1.0f;
}
/* colourspaces.glsl:2-19 */
static void fc_rec2020_to_xyY(const float *rgb, float *xyY)
{
  float xyz[3];
  o_mat3mulv(M_2020_to_xyz, rgb, xyz);
  const float s = 1.0f * xyz[0] + 1.0f * xyz[1] + 1.0f * xyz[2];
  xyY[0] = xyz[0] / s; xyY[1] = xyz[1] / s; xyY[2] = xyz[1];
}
static void fc_xyY_to_rec2020(const float *xyY, float *rgb)
{
  const float xyz[3] = { xyY[0] * xyY[2] / xyY[1], xyY[1] * xyY[2] / xyY[1], (1.0f - xyY[0] - xyY[1]) * xyY[2] / xyY[1] };
  o_mat3mulv(M_xyz_to_2020, xyz, rgb);
}

/* one pixel of filmcurv/main.comp:68-166 */
void o_filmcurv_px(const float *col_in, float *col1, const o_filmcurv_params_t *p)
{
  const float il = o_max(5e-3f, p->light);
  const float k  = o_max(1e-4f, p->contrast);
  float col0[3];
  for(int c = 0; c < 3; c++) col0[c] = col_in[c] + p->bias;
  if(p->colour == 0)
  {
    for(int c = 0; c < 3; c++) col1[c] = weibull_cdf(col0[c], il, k);
    float xyz0[3], xyz1[3];
    o_mat3mulv(M_2020_to_xyz, col0, xyz0);
    o_mat3mulv(M_2020_to_xyz, col1, xyz1);
    const float s0 = o_max(1e-4f, xyz0[0] + xyz0[1] + xyz0[2]);
    const float s1 = o_max(1e-4f, xyz1[0] + xyz1[1] + xyz1[2]);
    const float xyY0[3] = { xyz0[0] / s0, xyz0[1] / s0, xyz0[1] };
    float xyY1[3] = { xyz1[0] / s1, xyz1[1] / s1, xyz0[1] };
    float jch0[3], jch1[3];
    o_xyY_to_dt_UCS_JCH(xyY0, 1.0f, jch0);
    o_xyY_to_dt_UCS_JCH(xyY1, 1.0f, jch1);
    jch1[2] = jch0[2];
    o_dt_UCS_JCH_to_xyY(jch1, 1.0f, xyY1);
    /* vec3(..) * xyz1.y / max(..) evaluates as (v * y) / m */
    const float m = o_max(1e-4f, xyY1[1]);
    const float w[3] = { xyY1[0] * xyz1[1] / m, xyY1[1] * xyz1[1] / m, (1.0f - xyY1[0] - xyY1[1]) * xyz1[1] / m };
    o_mat3mulv(M_xyz_to_2020, w, col1);
  }
  else if(p->colour == 3)
  {
    for(int c = 0; c < 3; c++) col1[c] = weibull_cdf(col0[c], il, k);
    o_adjust_colour_dng(col0, col1);
  }
  else if(p->colour == 1)
  {
    for(int c = 0; c < 3; c++) col1[c] = weibull_cdf(col0[c], il, k);
  }
  else if(p->colour == 4)
  { /* agx, mat3 given as column vec3s */
    float c[3], hsv0[3], hsv1[3];
    c[0] = 0.856627153315983f * col0[0] + 0.0951212405381588f * col0[1] + 0.0482516061458583f * col0[2];
    c[1] = 0.137318972929847f * col0[0] + 0.761241990602591f  * col0[1] + 0.101439036467562f  * col0[2];
    c[2] = 0.11189821299995f  * col0[0] + 0.0767994186031903f * col0[1] + 0.811302368396859f  * col0[2];
    rgb2hsv(c, hsv0);
    for(int q = 0; q < 3; q++) c[q] = weibull_cdf(c[q], il, k);
    rgb2hsv(c, hsv1);
    hsv1[0] = lerp_chromaticity_angle(hsv0[0], hsv1[0], 0.4f);
    hsv2rgb(hsv1, c);
    col1[0] =  1.1271005818144368f * c[0] - 0.11060664309660323f * c[1] - 0.016493938717834573f * c[2];
    col1[1] = -0.1413297634984383f * c[0] + 1.157823702216272f   * c[1] - 0.016493938717834257f * c[2];
    col1[2] = -0.14132976349843826f * c[0] - 0.11060664309660294f * c[1] + 1.2519364065950405f * c[2];
  }
  else if(p->colour == 5)
  {
    const float pi = (float)M_PI;
    float lab0[3], lab_pc[3], pc[3];
    rec2020_to_oklab(col0, lab0);
    const float L0 = o_max(lab0[0], 1e-7f);
    const float lum0 = o_max(col0[0] * 0.2627f + col0[1] * 0.6780f + col0[2] * 0.0593f, 1e-7f);
    float lum1 = weibull_cdf(lum0, il, k);
    lum1 = lum1 + p->rolloff * lum1 * lum1 * (1.0f - lum1);
    if(p->shadows != 0.0f)
    {
      const float toe_gamma = 1.0f - 0.5f * p->shadows;
      const float lum_toe = powf(o_max(lum1, 1e-7f), toe_gamma);
      lum1 = o_mix(lum1, lum_toe, o_smoothstep(0.3f, 0.0f, lum1));
    }
    const float L1 = L0 * powf(lum1 / lum0, 1.0f / 3.0f);
    for(int c = 0; c < 3; c++) pc[c] = weibull_cdf(col0[c], il, k);
    rec2020_to_oklab(pc, lab_pc);
    float C1 = sqrtf(lab_pc[1] * lab_pc[1] + lab_pc[2] * lab_pc[2]);
    const float tame = 1.0f - 0.6f * p->rolloff * o_smoothstep(0.15f, 0.5f, lum1);
    C1 *= o_mix(tame, 1.0f, o_clamp(p->chroma - 1.0f, 0.0f, 1.0f));
    float h = atan2f(lab_pc[2], lab_pc[1]);
    const float h_target = 0.96f;
    const float h_dist = fabsf(glsl_mod(h - h_target + pi, 2.0f * pi) - pi);
    if(h_dist < 0.7f)
    {
      const float away = fabsf(lum1 - 0.35f);
      const float converge = 0.3f * o_smoothstep(0.0f, 0.3f, away);
      h = lerp_chromaticity_angle(h, h_target, converge);
    }
    const float c = p->chroma - 1.0f;
    const float deriv = weibull_pdf(lum0, il, k);
    const float hi = powf(o_max(1.0f, 1.0f / o_max(deriv, 0.15f)), c * 0.1f);
    const float sw = o_smoothstep(0.3f, 0.0f, lum1);
    const float lo = 1.0f + c * 0.15f * sw;
    float cr = p->chroma * hi * lo;
    cr *= 1.0f + p->red * hue_bump(h, 0.7f, 1.0f) + p->yellow * hue_bump(h, 1.76f, 1.0f) + p->blue * hue_bump(h, -1.76f, 1.0f);
    const float lab1[3] = { L1, C1 * cr * cosf(h), C1 * cr * sinf(h) };
    oklab_to_rec2020(lab1, col1);
  }
  else if(p->colour == 2)
  { /* :96-104 */
    for(int c = 0; c < 3; c++) col1[c] = weibull_cdf(col0[c], il, k);
    float xyY0[3], xyY1[3], m0[2], m1[2];
    fc_rec2020_to_xyY(col0, xyY0);
    munsell_from_xy(xyY0, m0);
    fc_rec2020_to_xyY(col1, xyY1);
    munsell_from_xy(xyY1, m1);
    const float mhc[2] = { m0[0], m1[1] };
    munsell_to_xy(mhc, xyY1);
    fc_xyY_to_rec2020(xyY1, col1);
  }
  else { col1[0] = col1[1] = col1[2] = 0.0f; } /* no such mode; glsl would store an undefined col1 */
}

void o_filmcurv_main(const oimg_t *in, oimg_t *out, const o_filmcurv_params_t *p, int out_f16)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float c0[4], c1[4];
    o_fetch4(in, x, y, c0);
    o_filmcurv_px(c0, c1, p);
    c1[3] = 1.0f;
    o_store4(out, x, y, c1, out_f16);
  }
}

/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/colour/main.c:88-186,219-365 (commit_params) and
 * colour/main-impl.glsl:49-67,118-341 for the paths that need no external LUT input
 * (clut / abney / spectra / picked connectors unconnected: have_clut = have_pick = have_abney = 0),
 * plus shared/dtucs.glsl:11-83 and colourspaces.glsl:1-18.
 * Out of scope (documented in DESIGN.md): camera-log TRCs 7..15 and camera gamuts 8,9,11..16. */
#include "o_common.h"
#include "vkdt_oracle.h"

static const float M_cat16_Mi[9] = {1.86206786f, -1.01125463f, 0.14918677f, 0.38752654f, 0.62144744f, -0.00897398f, -0.01584150f, -0.03412294f, 1.04996444f};
static const float M_cat16_M[9]  = {0.401288f, 0.650173f, -0.051461f, -0.250268f, 1.204414f, 0.045854f, -0.002079f, 0.048952f, 0.953127f};
static const float M_2020_to_xyz[9] = {0.636958048301290991f, 0.144616903586208406f, 0.168880975164172054f, 0.26270021201126692f, 0.677998071518871148f, 0.0593017164698619384f, 4.9999999999999999e-17f, 0.0280726930490874452f, 1.06098505771079066f};
static const float M_xyz_to_2020[9] = {1.71665119f, -0.35567078f, -0.25336628f, -0.66668435f, 1.61648124f, 0.01576855f, 0.01763986f, -0.04277061f, 0.94210312f};
static const float M_709_to_2020[9] = {0.62750375f, 0.32927542f, 0.04330266f, 0.06910828f, 0.91951916f, 0.0113596f, 0.01639406f, 0.08801125f, 0.89538035f};
static const float M_adobe_to_2020[9] = {0.87736306f, 0.07751751f, 0.04516292f, 0.0966218f, 0.89152263f, 0.01186405f, 0.02291617f, 0.04301452f, 0.93367996f};
static const float M_p3d65_to_2020[9] = {0.75386031f, 0.19861268f, 0.04757049f, 0.04575344f, 0.94178472f, 0.01247032f, -0.00121501f, 0.01760596f, 0.98321971f};
static const float M_ap0_to_2020[9] = {1.51286139f, -0.2589874f, -0.22978603f, -0.07903646f, 1.17706683f, -0.10075565f, 0.00209124f, -0.03114411f, 0.95350416f};
static const float M_ap1_to_2020[9] = {1.03866457f, -1.14744180e-02f, -2.72327263e-02f, -4.33683734e-04f, 1.00062477f, 1.01851049e-04f, -5.64306018e-03f, -2.23568741e-02f, 1.02483276f};
static const float M_redwg_to_2020[9] = {1.180431f, -0.094040f, -0.086391f, -0.028017f, 1.311442f, -0.283425f, -0.074360f, -0.362078f, 1.436437f};

/* colour/main.c:76-186: thin plate (linear kernel) rbf coefficients; coef has 4*(3+N) floats, zero-inited */
static void compute_coefficients(int N, const float *source, const float *target, float *coef)
{
  const int N2 = N + 3;
  if(N == 0) { for(int co = 0; co < 3; co++) coef[co*4+co] = 1.0f; return; }
  if(N == 1) { for(int co = 0; co < 3; co++) coef[co*4+co] = target[co] / source[co]; return; }
  double *A = (double *)malloc(sizeof(double) * N2 * N2);
  double *b = (double *)malloc(sizeof(double) * N2);
  double *A0 = (double *)malloc(sizeof(double) * N2 * N2);
  for(int j = 0; j < N; j++) for(int i = j; i < N; i++)
  {
    const float *x = source + 3*i, *y = source + 3*j;
    const double r2 = (x[0]-y[0])*(x[0]-y[0]) + (x[1]-y[1])*(x[1]-y[1]) + (x[2]-y[2])*(x[2]-y[2]);
    A[j*N2+i] = A[i*N2+j] = sqrt(r2);
  }
  for(int k = 0; k < 3; k++) for(int i = 0; i < N; i++) A[i*N2+N+k] = A[(N+k)*N2+i] = source[3*i+k];
  for(int j = N; j < N2; j++) for(int i = N; i < N2; i++) A[j*N2+i] = 0;
  /* the reference triangularises once and back-substitutes three times; solving the same
   * factorisation three times is arithmetically identical to re-running o_gauss_solve on a copy. */
  memcpy(A0, A, sizeof(double) * N2 * N2);
  for(int ch = 0; ch < 3; ch++)
  {
    memcpy(A, A0, sizeof(double) * N2 * N2);
    for(int i = 0; i < N; i++) b[i] = target[3*i+ch];
    for(int i = N; i < N2; i++) b[i] = 0;
    if(!o_gauss_solve(A, b, N2)) break;
    for(int i = 0; i < N; i++) coef[12 + 4*i + ch] = b[i];
    for(int i = 0; i < 3; i++) coef[4*i + ch] = b[N+i];
  }
  free(A); free(A0); free(b);
}

/* colour/main.c:219-365.  f has O_COLOUR_COMMITTED_FLOATS entries.  p_wb is written back (side effect of the reference). */
void o_colour_commit(const o_colour_params_t *p, float *p_wb, const float *img_wb, const float *img_cam_to_rec2020,
    int img_primaries, int img_trc, float *f)
{
  uint32_t *ii = (uint32_t *)f;
  memset(f, 0, sizeof(float) * O_COLOUR_COMMITTED_FLOATS);
  if(p_wb[0] == 0.0f && p_wb[1] == 0.0f && p_wb[2] == 0.0f)
  {
    float w0[3] = {0}, w[3] = { img_wb[0], img_wb[1], img_wb[2] };
    for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) w0[j] += img_cam_to_rec2020[3*j+i] / w[i];
    w0[0] /= w0[1]; w0[2] /= w0[1]; w0[1] = 1.0f;
    p_wb[0] = 1 / w0[0]; p_wb[1] = 1; p_wb[2] = 1 / w0[2];
  }
  if(!(p_wb[0] == p_wb[0]) || p_wb[0] == 0.0f || p_wb[1] == 0.0f || p_wb[2] == 0.0f)
    p_wb[0] = p_wb[1] = p_wb[2] = 1.0f;
  f[0] = p_wb[0] / p_wb[1];
  f[1] = 1.0f;
  f[2] = p_wb[2] / p_wb[1];
  f[3] = powf(2.0f, p->exposure);
  const int off = 4+12+4+12+4*24+4*24;
  /* clut unconnected: nbands = 3, legacy mapping (colour/main.c:272-292) */
  if(p->temp <= 0.0f) f[off+0] = -1.0f;
  else
  {
    float v = tanf(asinhf(46.3407f + p->temp)) + (-0.0287128f * cosf(0.000798585f * (714.855f - p->temp))) + 0.942275f;
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    f[off+0] = 1.0f - v;
  }
  ii[off+1] = p->matrix == 4 ? 1 : 0;
  f[off+2] = p->sat;
  ii[off+3] = p->picked;
  ii[off+4] = p->gamut;
  ii[off+5] = img_primaries;
  ii[off+6] = img_trc;
  f[off+7] = p->clip ? p->clipmax : 0.0;
  float awb[3] = { img_wb[0], img_wb[1], img_wb[2] };
  if(!(awb[0] > 0.0f) || !(awb[1] > 0.0f) || !(awb[2] > 0.0f)) awb[0] = awb[1] = awb[2] = 1.0f;
  f[off+8] = awb[0] / awb[1]; f[off+9] = 1.0f; f[off+10] = awb[2] / awb[1]; f[off+11] = 1.0f;
  if(p->matrix == 1)
  { for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) f[4+4*i+j] = img_cam_to_rec2020[3*j+i]; }
  else if(p->matrix == 2) { ii[off+5] = 5; ii[off+6] = 0; }  /* XYZ, linear */
  else if(p->matrix == 3) { ii[off+5] = 1; ii[off+6] = 0; }  /* rec709, linear */
  else if(p->matrix == 5)
  {
    ii[off+5] = 0; ii[off+6] = 0;
    for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) f[4+4*i+j] = p->mat[3*j+i];
  }
  else
  {
    ii[off+5] = 2; ii[off+6] = 0;
    for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) f[4+4*j+i] = i == j ? 1.0f : 0.0f;
  }
  if(p->mode == 1)
  {
    const int N = p->cnt < 0 ? 0 : (p->cnt > 24 ? 24 : p->cnt);
    ii[16] = N; ii[17] = ii[18] = ii[19] = 0;
    float src[72], tgt[72];
    for(int i = 0; i < N; i++) for(int k = 0; k < 3; k++)
    {
      src[3*i+k] = p->rbmap[6*i+k];
      tgt[3*i+k] = p->rbmap[6*i+3+k];
    }
    memset(f + 20, 0, sizeof(float) * (12 + 24*4 + 24*4));
    for(int k = 0; k < N; k++)
    {
      f[128 + 4*k + 0] = src[3*k+0];
      f[128 + 4*k + 1] = src[3*k+1];
      f[128 + 4*k + 2] = src[3*k+2];
      f[128 + 4*k + 3] = 0.0f;
    }
    compute_coefficients(N, src, tgt, f + 20);
  }
  else ii[16] = ii[17] = ii[18] = ii[19] = 0;
}

/* camera gamuts -> xyz (matrices.h:89-112), the fp32 value of every entry; index = primaries 8, 9, 11..16 */
static const float M_camgamut_to_xyz[8][9] = {
  /*  8 arriwg3         */ { 0.638007641f, 0.214703858f, 0.09774445f, 0.291953772f, 0.823841035f, -0.115794823f, 0.00279827905f, -0.0670342371f, 1.15329373f },
  /*  9 arriwg4         */ { 0.704858303f, 0.129760295f, 0.115837313f, 0.254524171f, 0.781477749f, -0.0360019095f, 0.0f, 0.0f, 1.0890578f },
  /* 11 sonysgamut3     */ { 0.706482708f, 0.128801048f, 0.115172163f, 0.270979673f, 0.786606431f, -0.0575860813f, -0.00967784505f, 0.00460003735f, 1.09413552f },
  /* 12 sonysgamut3cine */ { 0.5990839f, 0.248925522f, 0.102446489f, 0.215075821f, 0.885068476f, -0.100144319f, -0.0320658498f, -0.0276583899f, 1.14878201f },
  /* 13 vgamut          */ { 0.679644465f, 0.152211413f, 0.118600048f, 0.260685563f, 0.774894476f, -0.0355800129f, -0.00931019802f, -0.00461246725f, 1.10298038f },
  /* 14 egamut          */ { 0.705396831f, 0.164041325f, 0.0810177475f, 0.280130714f, 0.820206642f, -0.100337364f, -0.103781514f, -0.0729072541f, 1.26574647f },
  /* 15 egamut2         */ { 0.736477673f, 0.130739644f, 0.0832385793f, 0.275069982f, 0.828017771f, -0.103087775f, -0.124225155f, -0.0871597677f, 1.3004427f },
  /* 16 davinciwg       */ { 0.70062238f, 0.148774818f, 0.101058722f, 0.274118513f, 0.873631895f, -0.147750407f, -0.0989629105f, -0.137895331f, 1.32591593f },
};
static int camgamut_index(uint32_t prim) { return prim == 8 ? 0 : prim == 9 ? 1 : prim >= 11 && prim <= 16 ? (int)prim - 9 : -1; }

/* camera log curves to scene linear (shared/oetf.glsl:2-38), every operation in fp32 as written there; expressions of literals
 * alone are constants that glslang folds in double and rounds once ((float)(0.18 + 0.01), not 0.18f + 0.01f).
 * glsl's mix(a, b, cond) with a bvec selects: both sides are evaluated, b is taken where cond holds */
static float decode_log(float x, uint32_t trc)
{
  switch(trc)
  {
    case 7:  return x > 0.02740668f ? exp2f(x / 0.07329248f - 7.0f) - 0.0075f : x / 10.44426855f;                      /* davinci intermediate */
    case 8:  return x < 0.075f ? (x - 0.075f) / 16.184376489665897f : expf((x - 0.5520126568606655f) / 0.09232902596577353f) - 0.0057048244042473785f; /* filmlight t-log */
    case 9:  return x <= 0.155251141552511f ? (x - 0.0729055341958355f) / 10.5402377416545f : exp2f(x * 17.52f - 9.72f);   /* aces cct */
    case 10: return x < (float)(5.367655 * 0.010591 + 0.092809) ? (x - 0.092809f) / 5.367655f : (powf(10.0f, (x - 0.385537f) / 0.247190f) - 0.052272f) / 5.555556f; /* arri logC3 */
    case 11: return x < -0.7774983977293537f ? x * 0.3033266726886969f - 0.7774983977293537f
                  : (exp2f(14.0f * (x - 0.09286412512218964f) / 0.9071358748778103f + 6.0f) - 64.0f) / 2231.8263090676883f;  /* arri logC4 */
    case 12: return x < 0.0f ? (x / 15.1927f) - 0.01f : (powf(10.0f, x / 0.224282f) - 1.0f) / 155.975327f - 0.01f;       /* red log3G10 */
    case 13: return x < 0.181f ? (x - 0.125f) / 5.6f : powf(10.0f, (x - 0.598206f) / 0.241514f) - 0.00873f;              /* panasonic v-log */
    case 14: return x < (float)(171.2102946929 / 1023.0) ? (x * 1023.0f - 95.0f) * 0.01125f / (float)(171.2102946929 - 95.0)
                  : powf(10.0f, (x * 1023.0f - 420.0f) / 261.5f) * (float)(0.18 + 0.01) - 0.01f;                                /* sony s-log3 */
    case 15: return x < 0.100686685370811f ? (x - 0.092864f) / 8.799461f
                  : powf(10.0f, (x - 0.384316f) / 0.245281f) / 5.555556f - (float)(0.064829 / 5.555556);                          /* fuji f-log2 */
    default: return x;
  }
}

/* colour/main-impl.glsl:118-198 */
static void decode_colour(const float *f, float *rgb)
{
  const uint32_t *ii = (const uint32_t *)f;
  const int off = 224;
  const uint32_t trc = ii[off+6], prim = ii[off+5];
  if(trc == 1)
  {
    const float a = 1.09929682680944f;
    for(int k = 0; k < 3; k++) rgb[k] = rgb[k] > (float)(0.018053968510807 * 4.5) ? powf((rgb[k] + (float)(1.09929682680944 - 1.0)) / a, 2.2f) : rgb[k] / 4.5f; /* constants of constants: folded in double */
  }
  else if(trc == 2)
  { for(int k = 0; k < 3; k++) rgb[k] = rgb[k] > 0.04045f ? powf((rgb[k] + 0.055f) / 1.055f, 2.4f) : rgb[k] / 12.92f; }
  else if(trc == 3)
  {
    const float m1 = 1305.0f/8192.0f, m2 = 2523.0f/32.0f, c1 = 107.0f/128.0f, c2 = 2413.0f/128.0f, c3 = 2392.0f/128.0f;
    for(int k = 0; k < 3; k++)
    {
      const float xpow = powf(o_max(0.0f, rgb[k]), 1.0f / m2);
      const float num = o_max(xpow - c1, 0.0f);
      const float den = o_max(c2 - c3 * xpow, 1e-10f);
      rgb[k] = powf(num / den, 1.0f / m1);
    }
  }
  else if(trc == 4) { for(int k = 0; k < 3; k++) rgb[k] = powf(rgb[k], 2.6f); }
  else if(trc == 5)
  {
    const float a = 0.17883277f, b = 0.28466892f, c = 0.55991073f;
    for(int k = 0; k < 3; k++) rgb[k] = rgb[k] <= 0.5f ? rgb[k] * rgb[k] / 3.0f : (expf((rgb[k] - c) / a) + b) / 12.0f;
  }
  else if(trc == 6) { for(int k = 0; k < 3; k++) rgb[k] = powf(o_max(rgb[k], 0.0f), 2.2f); }
  else if(trc >= 7 && trc <= 15) { for(int k = 0; k < 3; k++) rgb[k] = decode_log(rgb[k], trc); }
  if(prim == 0)
  { /* custom matrix, uploaded column major in f[4..15] */
    const float r = f[4] * rgb[0] + f[8]  * rgb[1] + f[12] * rgb[2];
    const float g = f[5] * rgb[0] + f[9]  * rgb[1] + f[13] * rgb[2];
    const float b = f[6] * rgb[0] + f[10] * rgb[1] + f[14] * rgb[2];
    rgb[0] = r; rgb[1] = g; rgb[2] = b;
  }
  else if(prim == 1) o_mat3mulv(M_709_to_2020, rgb, rgb);
  else if(prim == 3) o_mat3mulv(M_adobe_to_2020, rgb, rgb);
  else if(prim == 4) o_mat3mulv(M_p3d65_to_2020, rgb, rgb);
  else if(prim == 5) o_mat3mulv(M_xyz_to_2020, rgb, rgb);
  else if(prim == 6) o_mat3mulv(M_ap0_to_2020, rgb, rgb);
  else if(prim == 7) o_mat3mulv(M_ap1_to_2020, rgb, rgb);
  else if(prim == 10) o_mat3mulv(M_redwg_to_2020, rgb, rgb);
  else if(camgamut_index(prim) >= 0)
  { /* main-impl.glsl:179-196: M1 * M0 * rgb, evaluated left to right: the fp32 matrix product first */
    const float *M0 = M_camgamut_to_xyz[camgamut_index(prim)];
    float M[9];
    for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
      M[3*j+i] = M_xyz_to_2020[3*j+0] * M0[i] + M_xyz_to_2020[3*j+1] * M0[3+i] + M_xyz_to_2020[3*j+2] * M0[6+i];
    o_mat3mulv(M, rgb, rgb);
  }
}

/* colour/main-impl.glsl:49-67.  glsl evaluates M16 * rec2020_to_xyz * v left to right: (M16*R)*v */
static void cat16(float *rgb, const float *src, const float *dst)
{
  float MR[9];
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
    MR[3*j+i] = M_cat16_M[3*j+0] * M_2020_to_xyz[i] + M_cat16_M[3*j+1] * M_2020_to_xyz[3+i] + M_cat16_M[3*j+2] * M_2020_to_xyz[6+i];
  float XM[9];
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
    XM[3*j+i] = M_xyz_to_2020[3*j+0] * M_cat16_Mi[i] + M_xyz_to_2020[3*j+1] * M_cat16_Mi[3+i] + M_xyz_to_2020[3*j+2] * M_cat16_Mi[6+i];
  float cs[3], cd[3], cl[3];
  o_mat3mulv(MR, src, cs);
  o_mat3mulv(MR, dst, cd);
  o_mat3mulv(MR, rgb, cl);
  for(int k = 0; k < 3; k++) cl[k] *= cd[k] / cs[k];
  o_mat3mulv(XM, cl, rgb);
}

/* colourspaces.glsl:2-18 */
static void rec2020_to_xyY(const float *rgb, float *xyY)
{
  float xyz[3];
  o_mat3mulv(M_2020_to_xyz, rgb, xyz);
  const float s = xyz[0] + xyz[1] + xyz[2];
  xyY[0] = xyz[0] / s; xyY[1] = xyz[1] / s; xyY[2] = xyz[1];
}
static void xyY_to_rec2020(const float *xyY, float *rgb)
{
  const float xyz[3] = { xyY[0] * xyY[2] / xyY[1], xyY[1] * xyY[2] / xyY[1], (1.0f - xyY[0] - xyY[1]) * xyY[2] / xyY[1] };
  o_mat3mulv(M_xyz_to_2020, xyz, rgb);
}

/* shared/dtucs.glsl:11-83 */
void o_xyY_to_dt_UCS_JCH(const float *xyY, float L_white, float *JCH)
{
  /* M1 given as column vectors */
  const float ux = -0.783941002840055f * xyY[0] + 0.277512987809202f * xyY[1] + 0.153836578598858f;
  const float uy =  0.745273540913283f * xyY[0] - 0.205375866083878f * xyY[1] - 0.165478376301988f;
  const float ud =  0.318707282433486f * xyY[0] + 2.16743692732158f  * xyY[1] + 0.291320554395942f;
  const float u = ux / ud, v = uy / ud;
  const float us = 1.39656225667f * u / (fabsf(u) + 1.49217352929f);
  const float vs = 1.4513954287f  * v / (fabsf(v) + 1.52488637914f);
  /* M2 = mat2(-1.124983854323892, 1.86323315098672, -0.980483721769325, 1.971853092390862) columns */
  const float Up = -1.124983854323892f * us - 0.980483721769325f * vs;
  const float Vp =  1.86323315098672f  * us + 1.971853092390862f * vs;
  const float Y_hat = powf(xyY[2], 0.631651345306265f);
  const float L_star = 2.098883786377f * Y_hat / (Y_hat + 1.12426773749357f);
  const float M2 = Up * Up + Vp * Vp;
  JCH[0] = L_star / L_white;
  JCH[1] = 15.932993652962535f * powf(L_star, 0.6523997524738018f) * powf(M2, 0.6007557017508491f) / L_white;
  JCH[2] = atan2f(Vp, Up);
}
void o_dt_UCS_JCH_to_xyY(const float *JCH, float L_white, float *xyY)
{
  const float L_star = JCH[0] * L_white;
  float M = powf(JCH[1] * L_white / (15.932993652962535f * powf(L_star, 0.6523997524738018f)), 0.8322850678616855f);
  M = o_clamp(M, 0.0f, 0.05f);
  const float a = M * cosf(JCH[2]), b = M * sinf(JCH[2]);
  /* M1 = mat2(-5.037522385190711, 4.760029407436461, -2.504856328185843, 2.874012963239247) columns */
  const float us = -5.037522385190711f * a - 2.504856328185843f * b;
  const float vs =  4.760029407436461f * a + 2.874012963239247f * b;
  const float U = -1.49217352929f * us / (fabsf(us) - 1.39656225667f);
  const float V = -1.52488637914f * vs / (fabsf(vs) - 1.4513954287f);
  const float x = 0.167171472114775f * U + 0.141299802443708f * V - 0.00801531300850582f;
  const float y = -0.150959086409163f * U - 0.155185060382272f * V - 0.00843312433578007f;
  const float d = 0.940254742367256f * U + 1.000000000000000f * V - 0.0256325967652889f;
  xyY[0] = x / d; xyY[1] = y / d;
  xyY[2] = powf((1.12426773749357f * L_star / (2.098883786377f - L_star)), 1.5831518565279648f);
}

/* ---- lut inputs: main-impl.glsl:76-102 (process_clut), clut.glsl:1-20 ---- */
static void tri2quad(float *tc)
{
  tc[1] = tc[1] / (1.0f - tc[0]);
  tc[0] = (1.0f - tc[0]) * (1.0f - tc[0]);
}
static void clut_chroma(const oimg_t *clut, const float *tc, int idx, int nbands, float *rb)
{
  const int band = (nbands == 3) ? 2 * idx : idx;
  float t[4];
  o_tex4(clut, (tc[0] + (float)band) / (float)nbands, tc[1], t);
  rb[0] = t[0]; rb[1] = t[1];
}
static float clut_luminance(const oimg_t *clut, const float *tc, int idx, int n, int nbands)
{
  float t[4];
  if(nbands == 3) { o_tex4(clut, (tc[0] + 1.0f) / 3.0f, tc[1], t); return t[idx]; }
  o_tex4(clut, (tc[0] + (float)(n + idx / 2)) / (float)nbands, tc[1], t);
  return (idx % 2 == 0) ? t[0] : t[1];
}
static void process_clut(const oimg_t *clut, const float *f, float auto_temp, float *rgb)
{
  const float b = rgb[0] + rgb[1] + rgb[2];
  float tc[2] = { rgb[0] / b, rgb[2] / b };
  tri2quad(tc);
  const int nbands = clut->w / clut->h;
  const int n = (nbands * 2) / 3;
  const float temp = f[224] < 0.0f ? auto_temp : f[224];
  const float bp = o_clamp(temp, 0.0f, 1.0f) * (float)(n - 1);
  const int k0 = (int)bp;
  const int k1 = k0 + 1 < n - 1 ? k0 + 1 : n - 1;
  const float frac = bp - (float)k0;
  float rb0[2], rb1[2];
  clut_chroma(clut, tc, k0, nbands, rb0);
  clut_chroma(clut, tc, k1, nbands, rb1);
  const float rbx = o_mix(rb0[0], rb1[0], frac), rby = o_mix(rb0[1], rb1[1], frac);
  const float L = o_mix(clut_luminance(clut, tc, k0, n, nbands), clut_luminance(clut, tc, k1, n, nbands), frac);
  rgb[0] = rbx * L * b; rgb[1] = (1.0f - rbx - rby) * L * b; rgb[2] = rby * L * b;
}

/* colour/atemp-impl.glsl:57-87 (autotemp.comp, one invocation): where between the clut's temperature anchors the as-shot white
 * balance comes out neutral; -1 if the committed temperature is explicit.  have_pick = 0 */
float o_colour_autotemp(const oimg_t *clut, const float *f)
{
  if(f[224] >= 0.0f) return -1.0f;
  const int nbands = clut->w / clut->h;
  const int n = (nbands * 2) / 3;
  const float neutral[3] = { 1.0f / o_max(f[232], 1e-6f), 1.0f, 1.0f / o_max(f[234], 1e-6f) };
  const float nb = o_max(neutral[0] + neutral[1] + neutral[2], 1e-6f);
  float ntc[2] = { neutral[0] / nb, neutral[2] / nb };
  tri2quad(ntc);
  const float target = 1.0f / 3.0f;
  float prev[2], next[2];
  clut_chroma(clut, ntc, 0, nbands, prev);
  float best_bp = 0.0f, best_res = 1e30f;
  for(int k = 0; k < n - 1; k++)
  {
    clut_chroma(clut, ntc, k + 1, nbands, next);
    const float dv[2] = { next[0] - prev[0], next[1] - prev[1] };
    const float denom = dv[0] * dv[0] + dv[1] * dv[1];
    const float m = denom > 1e-12f ? ((target - prev[0]) * dv[0] + (target - prev[1]) * dv[1]) / denom : 0.0f;
    const float mc = o_clamp(m, 0.0f, 1.0f);
    const float ex = target - (prev[0] + mc * dv[0]), ey = target - (prev[1] + mc * dv[1]);
    const float res = sqrtf(ex * ex + ey * ey);
    if(res < best_res) { best_res = res; best_bp = (float)k + mc; }
    prev[0] = next[0]; prev[1] = next[1];
  }
  return best_bp / (float)(n - 1 > 1 ? n - 1 : 1);
}

/* colour/main-impl.glsl:200-341.  clut / (abney and spectra) may be null: have_clut = 0 / have_abney = 0; have_pick = 0 always
 * (the colour picker is not part of the path).  auto_temp: what the autotemp node would deliver, read if the committed
 * temperature is negative */
void o_colour_main_lut(const oimg_t *in, oimg_t *out, const float *f, int out_f16, const oimg_t *clut, const oimg_t *abney, const oimg_t *spectra, float auto_temp)
{
  const uint32_t *ii = (const uint32_t *)f;
  const int off = 224;
  const float sat = f[off+2], clip_hl = f[off+7];
  const uint32_t N = ii[16] > 24 ? 24 : ii[16];
  const uint32_t gamut_mode = ii[off+4];
  const int use_clut = clut && ii[off+1] != 0;
  const int have_abney = abney && spectra;
  const float one[3] = {1.0f, 1.0f, 1.0f};
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgb[4];
    o_tex4(in, (x + 0.5) / (double)out->w, (y + 0.5) / (double)out->h, rgb);
    if(!use_clut) decode_colour(f, rgb);
    else process_clut(clut, f, auto_temp, rgb);
    cat16(rgb, one, f);
    if(clip_hl > 0.0f)
    {
      float clip[3] = { clip_hl, clip_hl, clip_hl };
      if(!use_clut) decode_colour(f, clip);
      else { rgb[0] = rgb[1] = rgb[2] = clip_hl; process_clut(clut, f, auto_temp, rgb); } /* :245 assigns the PIXEL, not the clip colour */
      cat16(clip, one, f);
      const float t = o_min(clip[0], o_min(clip[1], clip[2]));
      for(int k = 0; k < 3; k++) rgb[k] = o_min(rgb[k], t);
    }
    for(int k = 0; k < 3; k++) rgb[k] *= f[3];
    if(ii[16] > 0)
    { /* rbf_P at f[20..31] as 3 column vec4s, rbf_c at f[32..], rbf_p at f[128..] */
      float co[3];
      for(int k = 0; k < 3; k++) co[k] = f[20+k] * rgb[0] + f[24+k] * rgb[1] + f[28+k] * rgb[2];
      for(uint32_t i = 0; i < N; i++)
      {
        const float d0 = rgb[0] - f[128+4*i], d1 = rgb[1] - f[129+4*i], d2 = rgb[2] - f[130+4*i];
        const float r = sqrtf(d0*d0 + d1*d1 + d2*d2);
        for(int k = 0; k < 3; k++) co[k] += f[32+4*i+k] * r;
      }
      for(int k = 0; k < 3; k++) rgb[k] = co[k];
    }
    if(!have_abney && sat != 1.0f)
    {
      for(int k = 0; k < 3; k++) rgb[k] = o_max(rgb[k], 0.0f);
      float xyY[3], JCH[3];
      rec2020_to_xyY(rgb, xyY);
      o_xyY_to_dt_UCS_JCH(xyY, 1.0f, JCH);
      JCH[1] = o_clamp(JCH[1] * sat, 0.0f, 1.0f);
      o_dt_UCS_JCH_to_xyY(JCH, 1.0f, xyY);
      xyY_to_rec2020(xyY, rgb);
    }
    else if(have_abney && (sat != 1.0f || gamut_mode > 0))
    { /* :287-335 saturation along lines of constant dominant wavelength, gamut compression */
      float xyY[3], lut[4], t4[4];
      rec2020_to_xyY(rgb, xyY);
      tri2quad(xyY);
      o_tex4(spectra, xyY[0], xyY[1], lut);
      float slx = lut[3], sly = -lut[1] / (2.0f * lut[0]);
      const float norm = (sly - 400.0f) / (700.0f - 400.0f) - 0.5f;
      sly = 0.5f * (0.5f + 0.5f * norm / sqrtf(norm * norm + 0.25f));
      if(lut[0] > 0.0f) sly += 0.5f;
      float m = sat * slx;
      const int sw = abney->w, sh = abney->h;
      if(gamut_mode > 0)
      {
        float bound = 1.0f;
        if(gamut_mode == 1) { o_fetch4(abney, sw - 1, (int)(sly * sh), t4); bound = t4[1]; }
        else if(gamut_mode == 2)
        {
          o_fetch4(abney, sw - 1, (int)(sly * sh), t4);
          bound = t4[0];
          slx *= t4[0] / t4[1];
          m = sat * slx;
        }
        else if(gamut_mode == 3)
        {
          o_fetch4(abney, sw - 2, (int)(sly * sh), t4);
          bound = t4[0];
          slx *= t4[0] / t4[1];
          m = sat * slx;
        }
        if(sat > 1.0f) slx = o_mix(slx, bound, (m - slx) / (m - slx + 1.0f));
        else slx = m;
        if(slx > bound) slx = bound;
      }
      slx = o_clamp(slx, 0.0f, (sw - 3.0f) / sw);
      o_tex4(abney, slx, sly, t4);
      xyY[0] = t4[0]; xyY[1] = t4[1];
      xyY_to_rec2020(xyY, rgb);
    }
    for(int k = 0; k < 3; k++) rgb[k] = o_clamp(rgb[k], -65535.0f, 65535.0f);
    rgb[3] = 1.0f;
    o_store4(out, x, y, rgb, out_f16);
  }
}

/* the same without lut inputs */
void o_colour_main(const oimg_t *in, oimg_t *out, const float *f, int out_f16)
{
  o_colour_main_lut(in, out, f, out_f16, 0, 0, 0, 0.0f);
}

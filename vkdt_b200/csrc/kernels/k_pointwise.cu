// fused pointwise chain: any sequence of crop -> colour -> filmcurv -> grade runs as ONE kernel,
// one 8-byte read and one 8/16-byte write per pixel, f16 rounding applied in registers at exactly the
// edges where the reference stores an f16 image (so results match the unfused graph).
//  - (crop, main) (colour, main) (filmcurv, main) (grade, main): single-op chains, drop-in per node
//  - (b200, pointw): the fused chain.  push: { u32 n_ops; u32 op[7] }, params: the ops' parameter blobs
//    back to back (crop: 20 floats committed, colour: 242 floats committed, filmcurv: 10 x 4 B, grade: 19 x 4 B)
// replaces crop/main.comp, colour/main.comp, filmcurv/main.comp, grade/main.comp and the three HBM
// round trips between them.
#include <string.h>
#include <math.h>
#include "pointwise.cuh"

enum { PW_CROP = 1, PW_COLOUR = 2, PW_FILMCURV = 3, PW_GRADE = 4, PW_COLENC = 5 };

struct pw_chain_t
{
  int n_ops;
  int op[8];
  int shift, sx, sy;   // crop resolved to an integer translation on the host (the default 3 px micro-crop)
  int out_f32;         // 0 rgba f16, 1 rgba f32, 2 packed rgb f32 (PFM payload), 3 rgba ui8 (o-jpg's sink image), 4 packed rgb ui8
  colenc_params_t colenc;
  crop_committed_t crop;
  filmcurv_params_t film;
  grade_params_t grade;
  colour_digest_t colour;
};

VKB_DEV f3 round3(f3 c) { return { f16r(c.x), f16r(c.y), f16r(c.z) }; }
// 8 bit sinks: rgba ui8 as one 32 bit word per pixel (alpha 255), or packed rgb bytes
VKB_DEV void st_sink_ui8(void *__restrict__ outv, int ow, int x, int y, f3 c, int mode)
{
  const uint32_t r = unorm8(c.x), g = unorm8(c.y), b = unorm8(c.z);
  if(mode == 3) reinterpret_cast<uint32_t *>(outv)[(size_t)y * ow + x] = r | (g << 8) | (b << 16) | 0xff000000u;
  else { uint8_t *o = reinterpret_cast<uint8_t *>(outv) + ((size_t)y * ow + x) * 3; o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b; }
}

// compile-time specialisations of the common chains: the op sequence is a template parameter pack, so the loop
// below unrolls into straight-line code (the generic runtime-switch kernel needed 104 registers).
template <int O0, int O1, int O2, int O3, bool F32, bool ROT>
__global__ void __launch_bounds__(256) k_pointwise_t(const uint2 *__restrict__ in, int iw, int ih,
    void *__restrict__ outv, int ow, int oh, const __grid_constant__ pw_chain_t P)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  constexpr int ops[4] = { O0, O1, O2, O3 };
  constexpr int n = (O0 != 0) + (O1 != 0) + (O2 != 0) + (O3 != 0);
  float4 px;
  if(O0 == PW_CROP)
  {
    if(ROT) { px = ld_rgba_clamp(in, iw, ih, x + P.sx, y + P.sy); px.w = 1.0f; } // host proved the gather is an integer shift
    else px = crop_fetch<false>(in, iw, ih, x, y, P.crop);
  }
  else px = ld_rgba(in, iw, x, y);
  f3 c = { px.x, px.y, px.z };
  if(O0 == PW_CROP && n > 1) c = round3(c);
#pragma unroll
  for(int o = (O0 == PW_CROP ? 1 : 0); o < n; o++)
  {
    if(ops[o] == PW_COLOUR)        c = colour_px(c, P.colour);
    else if(ops[o] == PW_FILMCURV) c = filmcurv_px(c, P.film);
    else if(ops[o] == PW_GRADE)    c = grade_px(c, P.grade);
    else if(ops[o] == PW_COLENC)   c = colenc_px(c, P.colenc);
    if(o < n - 1) c = round3(c);
  }
  if(F32 && P.out_f32 >= 3) st_sink_ui8(outv, ow, x, y, c, P.out_f32);
  else if(F32 && P.out_f32 == 2)
  { // packed rgb sink: the lanes with x < ow are still here, contiguous from lane 0
    __shared__ __align__(16) float stage[8][100];
    const int x0 = blockIdx.x * 32, nl = min(32, ow - x0);
    const float v[3] = { c.x, c.y, c.z };
    st_rgb_coop<3>(stage[threadIdx.y], reinterpret_cast<float *>(outv) + ((size_t)y * ow + x0) * 3, threadIdx.x, nl, v, 3 * nl);
  }
  else if(F32) st_sink_f32(outv, ow, x, y, c.x, c.y, c.z, P.out_f32);
  else st_rgba(reinterpret_cast<uint2 *>(outv), ow, x, y, make_float4(c.x, c.y, c.z, 1.0f));
}

__global__ void __launch_bounds__(256) k_pointwise(const uint2 *__restrict__ in, int iw, int ih,
    void *__restrict__ outv, int ow, int oh, const __grid_constant__ pw_chain_t P)
{
  const int y = blockIdx.y * 8 + threadIdx.y;
  if(y >= oh) return;
#pragma unroll
  for(int rep = 0; rep < 2; rep++)
  {
    const int x = blockIdx.x * 64 + rep * 32 + threadIdx.x;
    if(x >= ow) continue;
    float4 px;
    int first = 0;
    if(P.op[0] == PW_CROP) { px = P.crop.r[0] != 1.0f ? crop_fetch<true>(in, iw, ih, x, y, P.crop) : crop_fetch<false>(in, iw, ih, x, y, P.crop); first = 1; }
    else px = ld_rgba(in, iw, min(x, iw - 1), min(y, ih - 1));
    f3 c = { px.x, px.y, px.z };
    // every edge between two modules is an f16 image in the reference: round in registers where it would store
    if(first && P.n_ops > 1) c = round3(c);
    for(int o = first; o < P.n_ops; o++)
    {
      if(P.op[o] == PW_COLOUR)        c = colour_px(c, P.colour);
      else if(P.op[o] == PW_FILMCURV) c = filmcurv_px(c, P.film);
      else if(P.op[o] == PW_GRADE)    c = grade_px(c, P.grade);
      else if(P.op[o] == PW_COLENC)   c = colenc_px(c, P.colenc);
      if(o < P.n_ops - 1) c = round3(c);
    }
    if(P.out_f32 >= 3) st_sink_ui8(outv, ow, x, y, c, P.out_f32);
    else if(P.out_f32) st_sink_f32(outv, ow, x, y, c.x, c.y, c.z, P.out_f32);
    else st_rgba(reinterpret_cast<uint2 *>(outv), ow, x, y, make_float4(c.x, c.y, c.z, 1.0f));
  }
}

// ---- host: digest of colour's committed block (layout colour/main.c:260-364) ----
static void mat3mul_d(const double *A, const double *B, double *C)
{
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
    C[3 * j + i] = A[3 * j + 0] * B[i] + A[3 * j + 1] * B[3 + i] + A[3 * j + 2] * B[6 + i];
}
static const double M_cat16_Mi[9] = {1.86206786, -1.01125463, 0.14918677, 0.38752654, 0.62144744, -0.00897398, -0.01584150, -0.03412294, 1.04996444};
static const double M_cat16_M[9]  = {0.401288, 0.650173, -0.051461, -0.250268, 1.204414, 0.045854, -0.002079, 0.048952, 0.953127};
static const double M_2020_to_xyz[9] = {0.636958048301290991, 0.144616903586208406, 0.168880975164172054, 0.26270021201126692, 0.677998071518871148, 0.0593017164698619384, 4.9999999999999999e-17, 0.0280726930490874452, 1.06098505771079066};
static const double M_xyz_to_2020[9] = {1.71665119, -0.35567078, -0.25336628, -0.66668435, 1.61648124, 0.01576855, 0.01763986, -0.04277061, 0.94210312};
static const double M_709_to_2020[9] = {0.62750375, 0.32927542, 0.04330266, 0.06910828, 0.91951916, 0.0113596, 0.01639406, 0.08801125, 0.89538035};
static const double M_adobe_to_2020[9] = {0.87736306, 0.07751751, 0.04516292, 0.0966218, 0.89152263, 0.01186405, 0.02291617, 0.04301452, 0.93367996};
static const double M_p3d65_to_2020[9] = {0.75386031, 0.19861268, 0.04757049, 0.04575344, 0.94178472, 0.01247032, -0.00121501, 0.01760596, 0.98321971};
static const double M_ap0_to_2020[9] = {1.51286139, -0.2589874, -0.22978603, -0.07903646, 1.17706683, -0.10075565, 0.00209124, -0.03114411, 0.95350416};
static const double M_ap1_to_2020[9] = {1.03866457, -1.14744180e-02, -2.72327263e-02, -4.33683734e-04, 1.00062477, 1.01851049e-04, -5.64306018e-03, -2.23568741e-02, 1.02483276};
static const double M_redwg_to_2020[9] = {1.180431, -0.094040, -0.086391, -0.028017, 1.311442, -0.283425, -0.074360, -0.362078, 1.436437};
static const double M_ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
// camera gamuts -> xyz (matrices.h:89-112), the fp32 value of every entry; rows: primaries 8, 9, 11..16
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

static double host_decode_trc(double v, uint32_t trc)
{
  switch(trc)
  {
    case 1: { const double a = 1.09929682680944, b = 0.018053968510807; return v > b * 4.5 ? pow((v + (a - 1)) / a, 2.2) : v / 4.5; }
    case 2: return v > 0.04045 ? pow((v + 0.055) / 1.055, 2.4) : v / 12.92;
    case 3: { const double m1 = 1305.0 / 8192.0, m2 = 2523.0 / 32.0, c1 = 107.0 / 128.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0;
              const double xp = pow(fmax(0.0, v), 1.0 / m2); return pow(fmax(xp - c1, 0.0) / fmax(c2 - c3 * xp, 1e-10), 1.0 / m1); }
    case 4: return pow(v, 2.6);
    case 5: { const double a = 0.17883277, b = 0.28466892, c = 0.55991073; return v <= 0.5 ? v * v / 3.0 : (exp((v - c) / a) + b) / 12.0; }
    case 6: return pow(fmax(v, 0.0), 2.2);
    case 7:  return v > 0.02740668 ? exp2(v / 0.07329248 - 7.0) - 0.0075 : v / 10.44426855;
    case 8:  return v < 0.075 ? (v - 0.075) / 16.184376489665897 : exp((v - 0.5520126568606655) / 0.09232902596577353) - 0.0057048244042473785;
    case 9:  return v <= 0.155251141552511 ? (v - 0.0729055341958355) / 10.5402377416545 : exp2(v * 17.52 - 9.72);
    case 10: return v < 5.367655 * 0.010591 + 0.092809 ? (v - 0.092809) / 5.367655 : (pow(10.0, (v - 0.385537) / 0.247190) - 0.052272) / 5.555556;
    case 11: return v < -0.7774983977293537 ? v * 0.3033266726886969 - 0.7774983977293537 : (exp2(14.0 * (v - 0.09286412512218964) / 0.9071358748778103 + 6.0) - 64.0) / 2231.8263090676883;
    case 12: return v < 0.0 ? (v / 15.1927) - 0.01 : (pow(10.0, v / 0.224282) - 1.0) / 155.975327 - 0.01;
    case 13: return v < 0.181 ? (v - 0.125) / 5.6 : pow(10.0, (v - 0.598206) / 0.241514) - 0.00873;
    case 14: return v < 171.2102946929 / 1023.0 ? (v * 1023.0 - 95.0) * 0.01125 / (171.2102946929 - 95.0) : pow(10.0, (v * 1023.0 - 420.0) / 261.5) * (0.18 + 0.01) - 0.01;
    case 15: return v < 0.100686685370811 ? (v - 0.092864) / 8.799461 : pow(10.0, (v - 0.384316) / 0.245281) / 5.555556 - 0.064829 / 5.555556;
    default: return v;
  }
}

// shared/oetf.glsl:2-38 in fp32, one rounding per operation (volatile keeps the host compiler from folding or widening)
static float host_decode_log_f(float x, uint32_t trc)
{
  volatile float t, u;
  switch(trc)
  {
    case 7:  if(x > 0.02740668f) { t = x / 0.07329248f; t = t - 7.0f; t = exp2f(t); t = t - 0.0075f; return t; } t = x / 10.44426855f; return t;
    case 8:  if(x < 0.075f) { t = x - 0.075f; t = t / 16.184376489665897f; return t; } t = x - 0.5520126568606655f; t = t / 0.09232902596577353f; t = expf(t); t = t - 0.0057048244042473785f; return t;
    case 9:  if(x <= 0.155251141552511f) { t = x - 0.0729055341958355f; t = t / 10.5402377416545f; return t; } t = x * 17.52f; t = t - 9.72f; t = exp2f(t); return t;
    case 10: u = (float)(5.367655 * 0.010591 + 0.092809);   // constants of literals alone: folded in double, rounded once (glslang)
             if(x < u) { t = x - 0.092809f; t = t / 5.367655f; return t; } t = x - 0.385537f; t = t / 0.247190f; t = powf(10.0f, t); t = t - 0.052272f; t = t / 5.555556f; return t;
    case 11: if(x < -0.7774983977293537f) { t = x * 0.3033266726886969f; t = t - 0.7774983977293537f; return t; }
             t = x - 0.09286412512218964f; t = 14.0f * t; t = t / 0.9071358748778103f; t = t + 6.0f; t = exp2f(t); t = t - 64.0f; t = t / 2231.8263090676883f; return t;
    case 12: if(x < 0.0f) { t = x / 15.1927f; t = t - 0.01f; return t; } t = x / 0.224282f; t = powf(10.0f, t); t = t - 1.0f; t = t / 155.975327f; t = t - 0.01f; return t;
    case 13: if(x < 0.181f) { t = x - 0.125f; t = t / 5.6f; return t; } t = x - 0.598206f; t = t / 0.241514f; t = powf(10.0f, t); t = t - 0.00873f; return t;
    case 14: u = (float)(171.2102946929 / 1023.0);
             if(x < u) { t = x * 1023.0f; t = t - 95.0f; t = t * 0.01125f; u = (float)(171.2102946929 - 95.0); t = t / u; return t; }
             t = x * 1023.0f; t = t - 420.0f; t = t / 261.5f; t = powf(10.0f, t); u = (float)(0.18 + 0.01); t = t * u; t = t - 0.01f; return t;
    case 15: if(x < 0.100686685370811f) { t = x - 0.092864f; t = t / 8.799461f; return t; }
             t = x - 0.384316f; t = t / 0.245281f; t = powf(10.0f, t); t = t / 5.555556f; u = (float)(0.064829 / 5.555556); t = t - u; return t;
    default: return x;
  }
}

// fp32 flavours for the strict build's launch constants (the host compiler neither contracts nor reassociates)
static float host_decode_trc_f(float v, uint32_t trc)
{
  switch(trc)
  {
    case 1: { const float a = 1.09929682680944f; return v > (float)(0.018053968510807 * 4.5) ? powf((v + (float)(1.09929682680944 - 1.0)) / a, 2.2f) : v / 4.5f; }
    case 2: return v > 0.04045f ? powf((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f;
    case 3: { const float m1 = 1305.0f / 8192.0f, m2 = 2523.0f / 32.0f, c1 = 107.0f / 128.0f, c2 = 2413.0f / 128.0f, c3 = 2392.0f / 128.0f;
              const float xp = powf(fmaxf(0.0f, v), 1.0f / m2); return powf(fmaxf(xp - c1, 0.0f) / fmaxf(c2 - c3 * xp, 1e-10f), 1.0f / m1); }
    case 4: return powf(v, 2.6f);
    case 5: { const float a = 0.17883277f, b = 0.28466892f, c = 0.55991073f; return v <= 0.5f ? v * v / 3.0f : (expf((v - c) / a) + b) / 12.0f; }
    case 6: return powf(fmaxf(v, 0.0f), 2.2f);
    case 7: case 8: case 9: case 10: case 11: case 12: case 13: case 14: case 15: return host_decode_log_f(v, trc);
    default: return v;
  }
}
static void host_mat3v_f(const float *M, const float *x, float *y)
{
  for(int j = 0; j < 3; j++) y[j] = M[3 * j + 0] * x[0] + M[3 * j + 1] * x[1] + M[3 * j + 2] * x[2];
}

static int colour_digest(const float *f, uint32_t size, colour_digest_t *d)
{
  if(size < 236 * 4) return vkb_set_error(VKB_ERR_BAD_ARG, "colour: committed params too small (%u bytes)", size);
  const uint32_t *ii = (const uint32_t *)f;
  const int off = 224;
  memset(d, 0, sizeof(*d));
  const uint32_t prim = ii[off + 5];
  d->trc = ii[off + 6];
  if(d->trc > 15) return vkb_set_error(VKB_ERR_BAD_ARG, "colour: unknown transfer curve %u", d->trc);
  double P[9];
  float Pf[9]; bool have_Pf = false; // a primaries matrix the shader itself forms in fp32 (camera gamuts)
  switch(prim)
  {
    case 0: for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) P[3 * j + i] = f[4 + 4 * i + j]; break; // column major upload
    case 1: memcpy(P, M_709_to_2020, sizeof(P)); break;
    case 2: memcpy(P, M_ident, sizeof(P)); break;
    case 3: memcpy(P, M_adobe_to_2020, sizeof(P)); break;
    case 4: memcpy(P, M_p3d65_to_2020, sizeof(P)); break;
    case 5: memcpy(P, M_xyz_to_2020, sizeof(P)); break;
    case 6: memcpy(P, M_ap0_to_2020, sizeof(P)); break;
    case 7: memcpy(P, M_ap1_to_2020, sizeof(P)); break;
    case 10: memcpy(P, M_redwg_to_2020, sizeof(P)); break;
    case 8: case 9: case 11: case 12: case 13: case 14: case 15: case 16:
    { // main-impl.glsl:179-196: xyz_to_rec2020 * gamut_to_xyz, the product formed in fp32 (left to right) before it meets the pixel
      const float *M0 = M_camgamut_to_xyz[prim < 10 ? prim - 8 : prim - 9];
      static const float fX[9] = {1.71665119f, -0.35567078f, -0.25336628f, -0.66668435f, 1.61648124f, 0.01576855f, 0.01763986f, -0.04277061f, 0.94210312f};
      volatile float t;
      for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
      {
        t = fX[3 * j + 0] * M0[i]; float a = t; t = fX[3 * j + 1] * M0[3 + i]; a = a + t; t = fX[3 * j + 2] * M0[6 + i]; a = a + t;
        Pf[3 * j + i] = a; P[3 * j + i] = a;
      }
      have_Pf = true;
      break;
    }
    default: return vkb_set_error(VKB_ERR_BAD_ARG, "colour: unknown primaries %u", prim);
  }
  // cat16(rgb, src = 1, dst = mul.rgb): xyz_to_rec2020 * M16i * diag(cl_dst / cl_src) * M16 * rec2020_to_xyz (main-impl.glsl:49-67)
  double MR[9], XM[9], D[9] = {0}, T0[9], T1[9], A[9];
  mat3mul_d(M_cat16_M, M_2020_to_xyz, MR);
  mat3mul_d(M_xyz_to_2020, M_cat16_Mi, XM);
  for(int j = 0; j < 3; j++)
  {
    const double cs = MR[3 * j] + MR[3 * j + 1] + MR[3 * j + 2];
    const double cd = MR[3 * j] * f[0] + MR[3 * j + 1] * f[1] + MR[3 * j + 2] * f[2];
    D[4 * j] = cd / cs;
  }
  mat3mul_d(D, MR, T0);
  mat3mul_d(XM, T0, T1);
  mat3mul_d(T1, P, A);
  for(int k = 0; k < 9; k++) d->A[k] = (float)A[k];
  { // the strict build's matrices: fp32, formed like the shader forms them (main-impl.glsl:49-67; oracle/o_colour.c cat16())
    static const float fM16[9]  = {0.401288f, 0.650173f, -0.051461f, -0.250268f, 1.204414f, 0.045854f, -0.002079f, 0.048952f, 0.953127f};
    static const float fM16i[9] = {1.86206786f, -1.01125463f, 0.14918677f, 0.38752654f, 0.62144744f, -0.00897398f, -0.01584150f, -0.03412294f, 1.04996444f};
    static const float fR[9] = {0.636958048301290991f, 0.144616903586208406f, 0.168880975164172054f, 0.26270021201126692f, 0.677998071518871148f, 0.0593017164698619384f, 4.9999999999999999e-17f, 0.0280726930490874452f, 1.06098505771079066f};
    static const float fX[9] = {1.71665119f, -0.35567078f, -0.25336628f, -0.66668435f, 1.61648124f, 0.01576855f, 0.01763986f, -0.04277061f, 0.94210312f};
    volatile float t; // every product and sum rounded to fp32, in the restatement's order
    for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
    {
      t = fM16[3 * j + 0] * fR[i]; float a = t; t = fM16[3 * j + 1] * fR[3 + i]; a = a + t; t = fM16[3 * j + 2] * fR[6 + i]; a = a + t; d->MR[3 * j + i] = a;
      t = fX[3 * j + 0] * fM16i[i]; a = t; t = fX[3 * j + 1] * fM16i[3 + i]; a = a + t; t = fX[3 * j + 2] * fM16i[6 + i]; a = a + t; d->XM[3 * j + i] = a;
    }
    for(int j = 0; j < 3; j++)
    {
      t = d->MR[3 * j] * 1.0f; float cs = t; t = d->MR[3 * j + 1] * 1.0f; cs = cs + t; t = d->MR[3 * j + 2] * 1.0f; cs = cs + t;
      t = d->MR[3 * j] * f[0]; float cd = t; t = d->MR[3 * j + 1] * f[1]; cd = cd + t; t = d->MR[3 * j + 2] * f[2]; cd = cd + t;
      t = cd / cs; d->ratio[j] = t;
    }
    d->has_P = prim != 2;
    for(int k = 0; k < 9; k++) d->P[k] = have_Pf ? Pf[k] : (float)P[k]; // the double constants above are the fp32 literals' decimal strings: exact round trip
  }
  d->exposure = f[3];
  const float clip = f[off + 7];
  d->clip_t = 0.0f;
  if(clip > 0.0f)
  {
#if VKB_FAST
    const double c = host_decode_trc(clip, d->trc);
    double t = 1e30;
    for(int j = 0; j < 3; j++) t = fmin(t, (A[3 * j] + A[3 * j + 1] + A[3 * j + 2]) * c);
    d->clip_t = (float)t;
#else
    // main-impl.glsl:241-247: the clip level runs through decode_colour and cat16 like a pixel, in fp32
    float c[3], o[3];
    c[0] = c[1] = c[2] = host_decode_trc_f(clip, d->trc);
    if(d->has_P) { host_mat3v_f(d->P, c, o); memcpy(c, o, sizeof(c)); }
    host_mat3v_f(d->MR, c, o);
    for(int k = 0; k < 3; k++) o[k] *= d->ratio[k];
    host_mat3v_f(d->XM, o, c);
    d->clip_t = fminf(c[0], fminf(c[1], c[2]));
#endif
    if(!(d->clip_t > 0.0f)) return vkb_set_error(VKB_ERR_BAD_ARG, "colour: non-positive highlight clip level");
  }
  d->N = ii[16] > 24 ? 24 : ii[16];
  d->sat = f[off + 2];
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) d->rbf_P[3 * j + i] = f[20 + 4 * i + j];
  memcpy(d->rbf_c, f + 32, sizeof(float) * 96);
  memcpy(d->rbf_p, f + 128, sizeof(float) * 96);
  return VKB_OK;
}

// the default darkroom chain with default-shaped parameters, as straight-line code: crop as an integer shift,
// colour = matrix * exposure (no trc decode, no clipping, no rbf, saturation 1), filmcurv colour mode 3 (per channel
// weibull curve + dng hue preservation).  the launcher checks those conditions; every other combination keeps the
// kernels above.  same device functions, same f16 roundings: values are identical to k_pointwise_t, the difference is
// ~100 instead of ~190 issued instructions per pixel (the uniform parameter tests and the dead paths' register
// pressure), which is what decides the speed of this kernel.
// weibull_cdf() of pointwise.cuh on the bare SFU instructions.  __powf / __expf expand to the same ex2.approx / lg2.approx
// but wrap each in a denormal guard (compare, scale, unscale: 4 instructions instead of 1); arguments here are >= 1e-7 * il
// and a result below 2^-126 only ever enters 1 - x, so flushing it changes nothing.
// SAFE (strict): the launcher has checked that il and k are finite and that k log2(x il) stays inside +-126 for every x in
// [1e-7, 65541] (the clamp in front of the curve): powf then has no special case left and the kernel no out of line call
template <bool SAFE>
VKB_DEV float weibull_cdf_ftz(float x, float il, float k)
{
#if VKB_FAST
  const float p = ex2_ftz(k * lg2_ftz(fmaxf(x, 1e-7f) * il));   // __powf(x * il, k)
  return 1.0f - ex2_ftz(-p * 1.4426950408889634f);               // __expf(-p)
#else
  if(SAFE) return 1.0f - m_exp(-lme_powf_safe(fmaxf(x, 1e-7f) * il, k));
  return weibull_cdf(x, il, k);                                  // libm's powf and expf bit for bit
#endif
}
#define PW_NPX 4
template <bool F32, bool SAFE>
__global__ void __launch_bounds__(256) k_pointwise_dflt(const uint2 *__restrict__ in, int iw, int ih,
    void *__restrict__ outv, int ow, int oh, const __grid_constant__ pw_chain_t P, const band_t bd)
{
  // PW_NPX pixels per thread, 32 apart: all loads are issued before the first dependent instruction, which keeps
  // enough bytes in flight per SM for HBM latency (one 8 byte load per thread does not: 2.7 TB/s)
  const int y = BAND_BY * 8 + threadIdx.y;
  if(y >= oh || BAND_SKIP(y)) return;
  uint2 raw[PW_NPX];
#pragma unroll
  for(int q = 0; q < PW_NPX; q++)
  {
    const int x = (blockIdx.x * PW_NPX + q) * 32 + threadIdx.x;
    if(x < ow) raw[q] = __ldg(in + (size_t)(y + P.sy) * iw + x + P.sx); // inside the input for every output pixel: proven on the host
  }
#pragma unroll
  for(int q = 0; q < PW_NPX; q++)
  {
  const int x = (blockIdx.x * PW_NPX + q) * 32 + threadIdx.x;
  if(x >= ow) continue;
  const float4 px = h4_to_f4(raw[q]);
  // crop's output edge is f16: the fetched texel already is
  const colour_digest_t &C = P.colour;
#if VKB_FAST
  f3 o = { C.A[0] * px.x + C.A[1] * px.y + C.A[2] * px.z,
           C.A[3] * px.x + C.A[4] * px.y + C.A[5] * px.z,
           C.A[6] * px.x + C.A[7] * px.y + C.A[8] * px.z };
#else
  f3 o = colour_matrices({ px.x, px.y, px.z }, C);
#endif
  o.x *= C.exposure; o.y *= C.exposure; o.z *= C.exposure;
  o.x = clampf(o.x, -65535.0f, 65535.0f); o.y = clampf(o.y, -65535.0f, 65535.0f); o.z = clampf(o.z, -65535.0f, 65535.0f);
  o = round3(o);
  const float il = fmaxf(5e-3f, P.film.light), k = fmaxf(1e-4f, P.film.contrast);
  const f3 col0 = { o.x + P.film.bias, o.y + P.film.bias, o.z + P.film.bias };
  // adjust_colour_dng(col0, curve(col0)) (shared.glsl:371-387) sorts the channels by col0, keeps the curve values of the
  // largest and smallest and replaces the middle one by a blend of those two: the middle channel's own curve value
  // is never used.  sorting col0 first (same comparisons, same tie breaking) and evaluating the curve on the sorted
  // maximum and minimum gives bit-identical results with two curve evaluations instead of three and half the selects.
  float s0 = col0.x, s1 = col0.y, s2 = col0.z, t;
  const bool fx = s2 > s1; if(fx) { t = s2; s2 = s1; s1 = t; }
  const bool fy = s1 > s0; if(fy) { t = s0; s0 = s1; s1 = t; }
  const bool fz = s2 > s1; if(fz) { t = s2; s2 = s1; s1 = t; }
  float r0 = weibull_cdf_ftz<SAFE>(s0, il, k), r2 = weibull_cdf_ftz<SAFE>(s2, il, k);
#if VKB_FAST
  float r1 = mixf(r2, r0, m_div(s1 - s2 + 1e-6f, s0 - s2 + 1e-6f));
#else  // sorted and clamped to +-65535: the divisor lies in [1e-6, 131071], the call free exact quotient applies
  float r1 = mixf(r2, r0, div_f(s1 - s2 + 1e-6f, s0 - s2 + 1e-6f));
#endif
  if(fz) { t = r2; r2 = r1; r1 = t; }
  if(fy) { t = r0; r0 = r1; r1 = t; }
  if(fx) { t = r2; r2 = r1; r1 = t; }
  const f3 c = { r0, r1, r2 };
  if(F32 && P.out_f32 == 2)
  {
    __shared__ float stage[8][96];
    const int x0 = (blockIdx.x * PW_NPX + q) * 32, nl = min(32, ow - x0);
    const float v[3] = { c.x, c.y, c.z };
    st_rgb_coop<3>(stage[threadIdx.y], reinterpret_cast<float *>(outv) + ((size_t)y * ow + x0) * 3, threadIdx.x, nl, v, 3 * nl);
  }
  else if(F32) st_sink_f32(outv, ow, x, y, c.x, c.y, c.z, P.out_f32);
  else st_rgba(reinterpret_cast<uint2 *>(outv), ow, x, y, make_float4(c.x, c.y, c.z, 1.0f));
  }
}

static int launch_chain(const vkb_launch_t *l, int n_ops, const uint32_t *ops)
{
  VKB_REQUIRE(l->num_conn >= 2 && n_ops >= 1 && n_ops <= 8);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16);
  VKB_REQUIRE(out->format == VKB_TOKEN_F16 || out->format == VKB_TOKEN_F32 || out->format == VKB_TOKEN_UI8);
  VKB_REQUIRE(out->chan == 4 || (out->chan == 3 && out->format != VKB_TOKEN_F16)); // 3: packed rgb sinks (VKB_SINK_RGB_F32 / VKB_SINK_RGB_UI8)
  pw_chain_t P;
  memset(&P, 0, sizeof(P));
  P.n_ops = n_ops;
  P.out_f32 = out->format == VKB_TOKEN_F32 ? (out->chan == 3 ? 2 : 1) : (out->format == VKB_TOKEN_UI8 ? (out->chan == 3 ? 4 : 3) : 0);
  const uint8_t *pp = (const uint8_t *)l->params;
  uint32_t left = l->params_size;
  for(int o = 0; o < n_ops; o++)
  {
    P.op[o] = ops[o];
    uint32_t need = 0;
    switch(ops[o])
    {
      case PW_CROP:
        need = 20 * 4; VKB_REQUIRE(o == 0 && left >= need);
        memcpy(&P.crop, pp, need); break;
      case PW_COLOUR:
      {
        need = 242 * 4; if(left < need) need = left; // callers may pass the 236 floats the shader reads
        const int r = colour_digest((const float *)pp, need, &P.colour);
        if(r) return r;
        break;
      }
      case PW_FILMCURV:
        need = sizeof(filmcurv_params_t); VKB_REQUIRE(left >= need);
        memcpy(&P.film, pp, need);
        if(P.film.colour < 0 || P.film.colour > 5) return vkb_set_error(VKB_ERR_BAD_ARG, "filmcurv: no colour mode %d", P.film.colour);
        break;
      case PW_GRADE:
        need = sizeof(grade_params_t); VKB_REQUIRE(left >= need);
        memcpy(&P.grade, pp, need); break;
      case PW_COLENC:
        need = sizeof(colenc_params_t); VKB_REQUIRE(left >= need);
        memcpy(&P.colenc, pp, need); break;
      default: return vkb_set_error(VKB_ERR_BAD_ARG, "pointwise chain: unknown op %u", ops[o]);
    }
    pp += need; left -= need;
  }
  if(P.op[0] != PW_CROP) VKB_REQUIRE(in->wd == out->wd && in->ht == out->ht);
  else if(P.crop.r[0] == 1.0f)
  { // is crop/main.comp's gather (crop offset, rotation, homography, texelFetch(ivec2(rd*size))) a pure integer translation?
    // evaluate the transform in double on a 3x3 grid of output pixels; every sample must land within 1/4 texel of the
    // centre of texel (x + sx, y + sy) (the shader's fp32 error is ~1e-3 texels) and inside the input.
    const crop_committed_t &c = P.crop;
    const double tsx = in->wd, tsy = in->ht;
    bool ok = true; long sx = 0, sy = 0;
    for(int j = 0; j < 3 && ok; j++) for(int i = 0; i < 3 && ok; i++)
    {
      const double x = i * 0.5 * (out->wd - 1), y = j * 0.5 * (out->ht - 1);
      double xx = floor(x) + 0.5 + (double)c.crop[0] * tsx, yy = floor(y) + 0.5 + (double)c.crop[2] * tsy;
      const double dx = xx - tsx * .5, dy = yy - tsy * .5;
      xx = c.r[0] * dx + c.r[2] * dy + tsx * .5; yy = c.r[1] * dx + c.r[3] * dy + tsy * .5;
      const double hx = c.H[0] * xx + c.H[4] * yy + c.H[8], hy = c.H[1] * xx + c.H[5] * yy + c.H[9], hz = c.H[2] * xx + c.H[6] * yy + c.H[10];
      const double u = hx / hz, v = hy / hz;
      const long tx = (long)floor(u), ty = (long)floor(v);
      if(fabs(u - (tx + 0.5)) > 0.25 || fabs(v - (ty + 0.5)) > 0.25 || tx < 0 || ty < 0 || tx >= (long)in->wd || ty >= (long)in->ht) ok = false;
      const long ex = tx - (long)floor(x), ey = ty - (long)floor(y);
      if(i == 0 && j == 0) { sx = ex; sy = ey; }
      else if(ex != sx || ey != sy) ok = false;
    }
    if(ok) { P.shift = 1; P.sx = (int)sx; P.sy = (int)sy; }
  }
  dim3 block(32, 8), grid(vkb_cdiv(out->wd, 64), vkb_cdiv(out->ht, 8)), grid1(vkb_cdiv(out->wd, 32), vkb_cdiv(out->ht, 8));
  const int sig = P.op[0] | (P.op[1] << 4) | (P.op[2] << 8) | (P.op[3] << 12) | (n_ops > 4 ? 1 << 20 : 0);
  if(l->band_y0 >= 0 && !(sig == (PW_CROP | (PW_COLOUR << 4) | (PW_FILMCURV << 8)) && P.shift && P.colour.trc == 0 && !(P.colour.clip_t > 0.0f) &&
     P.colour.N == 0 && P.colour.sat == 1.0f && P.film.colour == 3))
    return vkb_set_error(VKB_ERR_BAD_ARG, "pointwise chain: only the default parameter kernel runs banded");
  if(sig == (PW_CROP | (PW_COLOUR << 4) | (PW_FILMCURV << 8)) && P.shift && P.colour.trc == 0 && !(P.colour.clip_t > 0.0f) &&
     P.colour.N == 0 && P.colour.sat == 1.0f && P.film.colour == 3)
  { // the default darkroom parameters: straight-line kernel
    dim3 gridn(vkb_cdiv(out->wd, 32 * PW_NPX), vkb_cdiv(out->ht, 8));
    const band_t bd = band_of(l, 1, 8, out->ht, &gridn.y);
    if(!gridn.y) return VKB_OK;
    // the curve's power: x il in [1e-7 il, 65541 il] (col0 = clamp(+-65535) + bias; the shader's max(x, 1e-7)), exponent k
    const float il = fmaxf(5e-3f, P.film.light), kk = fmaxf(1e-4f, P.film.contrast);
    const double lo = log2(1e-7 * (double)il), hi = log2((65536.0 + fabs((double)P.film.bias)) * (double)il);
    const bool safe = std::isfinite(il) && std::isfinite(kk) && std::isfinite(P.film.bias) && il < 1e30f &&
                      (double)kk * fmax(fabs(lo), fabs(hi)) < 120.0;
#define GO(F, S) k_pointwise_dflt<F, S><<<gridn, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->wd, out->ht, P, bd)
    if(P.out_f32) { if(safe) GO(true, true); else GO(true, false); }
    else          { if(safe) GO(false, true); else GO(false, false); }
#undef GO
    VKB_CHECK_LAUNCH();
    return VKB_OK;
  }
#define PW_CASE(A, B, C, D) \
  case ((A) | ((B) << 4) | ((C) << 8) | ((D) << 12)): \
    if((A) == PW_CROP && P.crop.r[0] != 1.0f) goto generic; /* rotation / perspective: catmull-rom gather, generic kernel */ \
    if((A) == PW_CROP && P.shift) { \
      if(P.out_f32) k_pointwise_t<A, B, C, D, true, true><<<grid1, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->wd, out->ht, P); \
      else k_pointwise_t<A, B, C, D, false, true><<<grid1, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->wd, out->ht, P); \
      break; } \
    if(P.out_f32) k_pointwise_t<A, B, C, D, true, false><<<grid1, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->wd, out->ht, P); \
    else k_pointwise_t<A, B, C, D, false, false><<<grid1, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->wd, out->ht, P); \
    break;
  switch(sig)
  {
    PW_CASE(PW_CROP, PW_COLOUR, PW_FILMCURV, 0)
    PW_CASE(PW_CROP, PW_COLOUR, PW_FILMCURV, PW_GRADE)
    PW_CASE(PW_COLOUR, PW_FILMCURV, 0, 0)
    PW_CASE(PW_CROP, PW_COLOUR, 0, 0)
    PW_CASE(PW_CROP, 0, 0, 0)
    PW_CASE(PW_COLOUR, 0, 0, 0)
    PW_CASE(PW_FILMCURV, 0, 0, 0)
    PW_CASE(PW_GRADE, 0, 0, 0)
    PW_CASE(PW_COLENC, 0, 0, 0)
    PW_CASE(PW_GRADE, PW_COLENC, 0, 0)
    PW_CASE(PW_CROP, PW_COLOUR, PW_FILMCURV, PW_COLENC)
    default:
    generic:
  k_pointwise<<<grid, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->wd, out->ht, P);
  }
#undef PW_CASE
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

// ---- colour with lut inputs (colour/main-impl.glsl:76-102 process_clut + clut.glsl, :287-335 abney / spectra) ----
// the luts are small images of the reference's offline tools: clut (rg f16, nbands x 1 squares side by side), abney (rg f16,
// the last two columns hold gamut bounds), spectra (rgba).  any channel count / f16 or f32 storage is read.
struct lut_t { const void *p; int w, h, chan, f32; };
VKB_DEV float4 lut_px(const lut_t &t, int x, int y)
{ // vulkan's fill rules for absent channels: (0, 0, 0, 1)
  float v[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
  const size_t o = ((size_t)y * t.w + x) * t.chan;
  for(int c = 0; c < t.chan && c < 4; c++)
    v[c] = t.f32 ? __ldg((const float *)t.p + o + c) : __half2float(__ushort_as_half(__ldg((const unsigned short *)t.p + o + c)));
  return make_float4(v[0], v[1], v[2], v[3]);
}
VKB_DEV float4 lut_fetch(const lut_t &t, int x, int y) { return lut_px(t, min(max(x, 0), t.w - 1), min(max(y, 0), t.h - 1)); }
VKB_DEV float4 lut_tex(const lut_t &t, float u, float v)
{ // texture(): linear, mirrored repeat; the restatement's ideal sampler (oracle/o_common.h o_tex4) in both builds
  double x = (double)u * (double)t.w - 0.5, y = (double)v * (double)t.h - 0.5;
  if(fabs(x - rint(x)) < 1.0 / 4096.0) x = rint(x);
  if(fabs(y - rint(y)) < 1.0 / 4096.0) y = rint(y);
  const double fx = floor(x), fy = floor(y);
  const float ax = (float)(x - fx), ay = (float)(y - fy);
  const int x0 = mirrori((int)fx, t.w), x1 = mirrori((int)fx + 1, t.w), y0 = mirrori((int)fy, t.h), y1 = mirrori((int)fy + 1, t.h);
  const float4 a = lut_px(t, x0, y0), b = lut_px(t, x1, y0), c = lut_px(t, x0, y1), d = lut_px(t, x1, y1);
  return make_float4((a.x * (1.0f - ax) + b.x * ax) * (1.0f - ay) + (c.x * (1.0f - ax) + d.x * ax) * ay,
                     (a.y * (1.0f - ax) + b.y * ax) * (1.0f - ay) + (c.y * (1.0f - ax) + d.y * ax) * ay,
                     (a.z * (1.0f - ax) + b.z * ax) * (1.0f - ay) + (c.z * (1.0f - ax) + d.z * ax) * ay,
                     (a.w * (1.0f - ax) + b.w * ax) * (1.0f - ay) + (c.w * (1.0f - ax) + d.w * ax) * ay);
}
VKB_DEV void tri2quad(float &x, float &y) { y = y / (1.0f - x); x = (1.0f - x) * (1.0f - x); }
VKB_DEV float2 clut_chroma(const lut_t &clut, float tx, float ty, int idx, int nbands)
{
  const int band = nbands == 3 ? 2 * idx : idx;
  const float4 t = lut_tex(clut, (tx + (float)band) / (float)nbands, ty);
  return make_float2(t.x, t.y);
}
VKB_DEV float clut_luminance(const lut_t &clut, float tx, float ty, int idx, int n, int nbands)
{
  if(nbands == 3) { const float4 t = lut_tex(clut, (tx + 1.0f) / 3.0f, ty); return idx == 0 ? t.x : idx == 1 ? t.y : idx == 2 ? t.z : t.w; }
  const float4 t = lut_tex(clut, (tx + (float)(n + idx / 2)) / (float)nbands, ty);
  return (idx % 2 == 0) ? t.x : t.y;
}
VKB_DEV f3 process_clut(const lut_t &clut, float temp, f3 rgb)
{ // camera rgb -> rec2020 through the two nearest temperature anchors of the lut
  const float b = rgb.x + rgb.y + rgb.z;
  float tx = rgb.x / b, ty = rgb.z / b;
  tri2quad(tx, ty);
  const int nbands = clut.w / clut.h;
  const int n = (nbands * 2) / 3;
  const float bp = clampf(temp, 0.0f, 1.0f) * (float)(n - 1);
  const int k0 = (int)bp, k1 = min(k0 + 1, n - 1);
  const float frac = bp - (float)k0;
  const float2 rb0 = clut_chroma(clut, tx, ty, k0, nbands), rb1 = clut_chroma(clut, tx, ty, k1, nbands);
  const float rbx = mixf(rb0.x, rb1.x, frac), rby = mixf(rb0.y, rb1.y, frac);
  const float L = mixf(clut_luminance(clut, tx, ty, k0, n, nbands), clut_luminance(clut, tx, ty, k1, n, nbands), frac);
  return { rbx * L * b, (1.0f - rbx - rby) * L * b, rby * L * b };
}
struct colour_lut_t
{
  lut_t clut, abney, spectra;
  int use_clut, have_abney;
  float temp, clip_hl;
  uint32_t gamut_mode;
  const float *auto_temp;   // the autotemp node's 1x1 answer, read when the committed temperature is negative
};
__global__ void __launch_bounds__(256) k_colour_lut(const uint2 *__restrict__ in, int w, int h, void *__restrict__ outv, int out_f32, const colour_digest_t P, const colour_lut_t Q)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= w || y >= h) return;
  const float4 px = ld_rgba(in, w, x, y);
  const float temp = Q.temp < 0.0f && Q.auto_temp ? __ldg(Q.auto_temp) : Q.temp;
  colour_digest_t noP = P;     // cat16 on its own: the lut delivers rec2020
  noP.has_P = 0;
  f3 c = { px.x, px.y, px.z }, o;
  if(!Q.use_clut)
  {
    if(P.trc) { c.x = decode_trc(c.x, P.trc); c.y = decode_trc(c.y, P.trc); c.z = decode_trc(c.z, P.trc); }
    o = colour_matrices(c, P);
    if(P.clip_t > 0.0f) { o.x = fminf(o.x, P.clip_t); o.y = fminf(o.y, P.clip_t); o.z = fminf(o.z, P.clip_t); }
  }
  else
  {
    o = colour_matrices(process_clut(Q.clut, temp, c), noP);
    if(Q.clip_hl > 0.0f)
    { // main-impl.glsl:245 assigns the converted clip colour to the PIXEL and leaves the clip colour as it was: followed to the letter
      o = process_clut(Q.clut, temp, { Q.clip_hl, Q.clip_hl, Q.clip_hl });
      const f3 cl = colour_matrices({ Q.clip_hl, Q.clip_hl, Q.clip_hl }, noP);
      const float t = fminf(cl.x, fminf(cl.y, cl.z));
      o.x = fminf(o.x, t); o.y = fminf(o.y, t); o.z = fminf(o.z, t);
    }
  }
  o.x *= P.exposure; o.y *= P.exposure; o.z *= P.exposure;
  if(P.N > 0)
  {
    f3 co = { P.rbf_P[0] * o.x + P.rbf_P[1] * o.y + P.rbf_P[2] * o.z,
              P.rbf_P[3] * o.x + P.rbf_P[4] * o.y + P.rbf_P[5] * o.z,
              P.rbf_P[6] * o.x + P.rbf_P[7] * o.y + P.rbf_P[8] * o.z };
    for(uint32_t i = 0; i < P.N; i++)
    {
      const float d0 = o.x - P.rbf_p[i][0], d1 = o.y - P.rbf_p[i][1], d2 = o.z - P.rbf_p[i][2];
      const float r = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
      co.x += P.rbf_c[i][0] * r; co.y += P.rbf_c[i][1] * r; co.z += P.rbf_c[i][2] * r;
    }
    o = co;
  }
  if(!Q.have_abney && P.sat != 1.0f)
  {
    o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f);
    const f3 xyz = rec2020_to_xyz(o);
    const float s = xyz.x + xyz.y + xyz.z;
    float J, C, H, xx, yy, Y;
    xyY_to_dt_UCS_JCH(xyz.x / s, xyz.y / s, xyz.y, 1.0f, J, C, H);
    C = clampf(C * P.sat, 0.0f, 1.0f);
    dt_UCS_JCH_to_xyY(J, C, H, 1.0f, xx, yy, Y);
    o = xyz_to_rec2020({ xx * Y / yy, yy * Y / yy, (1.0f - xx - yy) * Y / yy });
  }
  else if(Q.have_abney && (P.sat != 1.0f || Q.gamut_mode > 0))
  { // saturation along lines of constant dominant wavelength, compressed into the chosen gamut
    const f3 xyz = rec2020_to_xyz(o);
    const float s = xyz.x + xyz.y + xyz.z;
    float qx = xyz.x / s, qy = xyz.y / s;
    const float Y = xyz.y;
    tri2quad(qx, qy);
    const float4 lut = lut_tex(Q.spectra, qx, qy);
    float slx = lut.w, sly = -lut.y / (2.0f * lut.x);
    const float norm = (sly - 400.0f) / (700.0f - 400.0f) - 0.5f;
    sly = 0.5f * (0.5f + 0.5f * norm / sqrtf(norm * norm + 0.25f));
    if(lut.x > 0.0f) sly += 0.5f;
    float m = P.sat * slx;
    const int sw = Q.abney.w, sh = Q.abney.h;
    if(Q.gamut_mode > 0)
    {
      float bound = 1.0f;
      if(Q.gamut_mode == 1) bound = lut_fetch(Q.abney, sw - 1, (int)(sly * (float)sh)).y;
      else if(Q.gamut_mode == 2 || Q.gamut_mode == 3)
      { // rec2020 / rec709: the lower bound moves with the spectral locus scaled into the triangle
        const float4 ms = lut_fetch(Q.abney, Q.gamut_mode == 2 ? sw - 1 : sw - 2, (int)(sly * (float)sh));
        bound = ms.x;
        slx *= ms.x / ms.y;
        m = P.sat * slx;
      }
      if(P.sat > 1.0f) slx = mixf(slx, bound, (m - slx) / (m - slx + 1.0f));
      else slx = m;
      if(slx > bound) slx = bound;
    }
    slx = clampf(slx, 0.0f, ((float)sw - 3.0f) / (float)sw);
    const float4 xy = lut_tex(Q.abney, slx, sly);
    o = xyz_to_rec2020({ xy.x * Y / xy.y, xy.y * Y / xy.y, (1.0f - xy.x - xy.y) * Y / xy.y });
  }
  o.x = clampf(o.x, -65535.0f, 65535.0f); o.y = clampf(o.y, -65535.0f, 65535.0f); o.z = clampf(o.z, -65535.0f, 65535.0f);
  if(out_f32) reinterpret_cast<float4 *>(outv)[(size_t)y * w + x] = make_float4(o.x, o.y, o.z, 1.0f);
  else st_rgba(reinterpret_cast<uint2 *>(outv), w, x, y, make_float4(o.x, o.y, o.z, 1.0f));
}
static int lut_of(const vkb_image_t *im, lut_t *t)
{
  VKB_REQUIRE(im->data && im->wd > 0 && im->ht > 0 && im->chan >= 1 && im->chan <= 4);
  VKB_REQUIRE(im->format == VKB_TOKEN_F16 || im->format == VKB_TOKEN_F32);
  *t = { im->data, (int)im->wd, (int)im->ht, (int)im->chan, im->format == VKB_TOKEN_F32 ? 1 : 0 };
  return VKB_OK;
}
// (colour, main) with the node's seven connectors (colour/main.c:444-465: input output clut picked abney spectra autotemp) and
// the push constants { have_clut, have_pick, have_abney }
static int launch_colour_lut(const vkb_launch_t *l)
{
  const int32_t *pc = (const int32_t *)l->push;
  VKB_REQUIRE(l->num_conn >= 6);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16 && out->chan == 4 && (out->format == VKB_TOKEN_F16 || out->format == VKB_TOKEN_F32));
  VKB_REQUIRE(in->wd == out->wd && in->ht == out->ht);
  if(pc[1]) return vkb_set_error(VKB_ERR_BAD_ARG, "colour: the colour picker input is outside the hot-path scope");
  colour_digest_t P;
  colour_lut_t Q;
  memset(&Q, 0, sizeof(Q));
  const float *f = (const float *)l->params;
  const uint32_t *fi = (const uint32_t *)l->params;
  const int r = colour_digest(f, l->params_size < 242 * 4 ? l->params_size : 242 * 4, &P);
  if(r) return r;
  Q.use_clut = pc[0] && fi[225] != 0;
  Q.have_abney = pc[2] != 0;
  Q.temp = f[224]; Q.clip_hl = f[231]; Q.gamut_mode = fi[228];
  if(Q.use_clut)
  {
    if(Q.temp < 0.0f)
    { // as shot: the (colour, autotemp) node's answer on connector 6
      const vkb_image_t *at = l->num_conn >= 7 ? l->conn + 6 : 0;
      if(!at || !at->data || at->format != VKB_TOKEN_F32 || at->wd != 1 || at->ht != 1 || at->chan != 1)
        return vkb_set_error(VKB_ERR_BAD_ARG, "colour: an as-shot clut temperature (temp <= 0) needs the autotemp node's 1x1 f32 image on connector 6");
      Q.auto_temp = (const float *)at->data;
    }
    if(lut_of(l->conn + 2, &Q.clut)) return VKB_ERR_BAD_ARG;
    VKB_REQUIRE(Q.clut.w / Q.clut.h >= 3);
  }
  if(Q.have_abney) { if(lut_of(l->conn + 4, &Q.abney) || lut_of(l->conn + 5, &Q.spectra)) return VKB_ERR_BAD_ARG; }
  dim3 block(32, 8), grid(vkb_cdiv(out->wd, 32), vkb_cdiv(out->ht, 8));
  k_colour_lut<<<grid, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->format == VKB_TOKEN_F32, P, Q);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

// (colour, autotemp) (colour/atemp-impl.glsl:57-87, one invocation): the position between the clut's temperature anchors at which
// the as-shot white balance comes out neutral; -1 when the committed temperature is explicit.  connectors: clut, temp (1x1 f32), picked
__global__ void k_colour_autotemp(const lut_t clut, float *__restrict__ out, float temp, float awb_r, float awb_b)
{
  if(threadIdx.x || blockIdx.x) return;
  if(temp >= 0.0f) { out[0] = -1.0f; return; }
  const int nbands = clut.w / clut.h;
  const int n = (nbands * 2) / 3;
  const float nr = 1.0f / fmaxf(awb_r, 1e-6f), nbl = 1.0f / fmaxf(awb_b, 1e-6f);
  const float nb = fmaxf(nr + 1.0f + nbl, 1e-6f);
  float tx = nr / nb, ty = nbl / nb;
  tri2quad(tx, ty);
  const float target = 1.0f / 3.0f;
  float2 prev = clut_chroma(clut, tx, ty, 0, nbands);
  float best_bp = 0.0f, best_res = 1e30f;
  for(int k = 0; k < n - 1; k++)
  {
    const float2 next = clut_chroma(clut, tx, ty, k + 1, nbands);
    const float dx = next.x - prev.x, dy = next.y - prev.y;
    const float denom = dx * dx + dy * dy;
    const float m = denom > 1e-12f ? ((target - prev.x) * dx + (target - prev.y) * dy) / denom : 0.0f;
    const float mc = clampf(m, 0.0f, 1.0f);
    const float ex = target - (prev.x + mc * dx), ey = target - (prev.y + mc * dy);
    const float res = sqrtf(ex * ex + ey * ey);
    if(res < best_res) { best_res = res; best_bp = (float)k + mc; }
    prev = next;
  }
  out[0] = best_bp / (float)max(n - 1, 1);
}
static int launch_colour_autotemp(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->params_size >= 236 * 4);
  const int32_t *pc = (const int32_t *)l->push;
  if(l->push_size >= 4 && pc[0]) return vkb_set_error(VKB_ERR_BAD_ARG, "colour: the colour picker input is outside the hot-path scope");
  const vkb_image_t *out = l->conn + 1;
  VKB_REQUIRE(out->data && out->format == VKB_TOKEN_F32 && out->wd == 1 && out->ht == 1 && out->chan == 1);
  lut_t clut;
  if(lut_of(l->conn, &clut)) return VKB_ERR_BAD_ARG;
  VKB_REQUIRE(clut.w / clut.h >= 3);
  const float *f = (const float *)l->params;
  k_colour_autotemp<<<1, 32, 0, l->stream>>>(clut, (float *)out->data, f[224], f[232], f[234]);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

static int launch_crop(const vkb_launch_t *l)     { const uint32_t op = PW_CROP;     return launch_chain(l, 1, &op); }
static int launch_colour(const vkb_launch_t *l)
{
  const int32_t *pc = (const int32_t *)l->push;
  if(l->push_size >= 12 && l->num_conn >= 6 && (pc[0] || pc[1] || pc[2])) return launch_colour_lut(l);
  const uint32_t op = PW_COLOUR; return launch_chain(l, 1, &op);
}
static int launch_filmcurv(const vkb_launch_t *l) { const uint32_t op = PW_FILMCURV; return launch_chain(l, 1, &op); }
static int launch_grade(const vkb_launch_t *l)    { const uint32_t op = PW_GRADE;    return launch_chain(l, 1, &op); }
static int launch_colenc(const vkb_launch_t *l)   { const uint32_t op = PW_COLENC;   return launch_chain(l, 1, &op); }
static int launch_pointw(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->push_size >= 8);
  const uint32_t *pc = (const uint32_t *)l->push;
  VKB_REQUIRE(pc[0] >= 1 && pc[0] <= 7 && l->push_size >= 4 * (1 + pc[0]));
  return launch_chain(l, pc[0], pc + 1);
}
VKB_REGISTER("crop", "main", launch_crop);
VKB_REGISTER("colour", "main", launch_colour);
VKB_REGISTER("colour", "autotemp", launch_colour_autotemp);
VKB_REGISTER("filmcurv", "main", launch_filmcurv);
VKB_REGISTER("grade", "main", launch_grade);
VKB_REGISTER("colenc", "main", launch_colenc);
VKB_REGISTER("b200", "pointw", launch_pointw);

VKB_NS_END

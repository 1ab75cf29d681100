// per-pixel device functions of the pointwise modules: crop (gather front end), colour, filmcurv, grade.
// restated from crop/main.comp:20-57, colour/main-impl.glsl:118-341, filmcurv/main.comp:68-166 +
// params.glsl:15-34, grade/main.comp:21-62, shared.glsl:47-96,371-387, shared/dtucs.glsl:11-83,
// colourspaces.glsl:2-78.  transcendentals: m_pow / m_exp of common.cuh, i.e. libm's results bit for bit in the strict
// build and the SFU intrinsics (what GLSL pow/exp compile to on a GPU) in the fast one.
#pragma once
#include "common.cuh"
#include "munsell_table.h"

#define PW_POW(x, y) m_pow((x), (y))
#define PW_EXP(x)    m_exp((x))

// ---- parameter blocks, same byte layout as the reference's uniform blocks ----
struct crop_committed_t { float H[12]; float r[4]; float crop[4]; };                     // crop/main.c:311-335
struct filmcurv_params_t { float light, contrast, bias; int colour; float chroma, rolloff, red, yellow, blue, shadows; };
struct grade_params_t { float lift[4], gamma[4], gain[4], off[4]; int mode; float sh_pivot, hi_pivot; };
struct colenc_params_t { int prim, trc; };                                               // colenc/params
// colour: host side digest of the 242-float committed block (colour/main.c:260-364)
struct colour_digest_t
{
  // strict build: the shader's own sequence per pixel (main-impl.glsl:152-198 decode_colour's primaries matrix, then
  // cat16() :49-67 = XM * (diag(cd/cs) * (MR * rgb))), every matrix in fp32 exactly as the shader / restatement forms it
  float P[9]; uint32_t has_P;   // primaries -> rec2020, row major (has_P == 0: rec2020 input, no multiplication)
  float MR[9], XM[9], ratio[3]; // M16 * rec2020_to_xyz, xyz_to_rec2020 * M16i, cl_dst / cl_src
  // fast build: all of it premultiplied on the host in double
  float A[9];        // row major: xyz_to_rec2020 * M16i * diag(cl_dst/cl_src) * M16 * rec2020_to_xyz * primaries
  float exposure;    // mul.w
  float clip_t;      // min channel of the processed clip colour, <= 0: no clipping
  uint32_t trc;
  uint32_t N;        // rbf points
  float sat;
  float rbf_P[9];    // row major
  float rbf_c[24][4];
  float rbf_p[24][4];
};

// ---- crop ----
VKB_DEV float4 catmull_rom_rgba(const uint2 *__restrict__ tex, int w, int h, float u, float v)
{ // shared.glsl:47-96
  const float sx = (float)w, sy = (float)h;
  const float spx = u * sx, spy = v * sy;
  const float t1x = floorf(spx - 0.5f) + 0.5f, t1y = floorf(spy - 0.5f) + 0.5f;
  const float fx = spx - t1x, fy = spy - t1y;
  const float w0x = fx * (-0.5f + fx * (1.0f - 0.5f * fx)), w0y = fy * (-0.5f + fy * (1.0f - 0.5f * fy));
  const float w1x = 1.0f + fx * fx * (-2.5f + 1.5f * fx),   w1y = 1.0f + fy * fy * (-2.5f + 1.5f * fy);
  const float w2x = fx * (0.5f + fx * (2.0f - 1.5f * fx)),  w2y = fy * (0.5f + fy * (2.0f - 1.5f * fy));
  const float w3x = fx * fx * (-0.5f + 0.5f * fx),          w3y = fy * fy * (-0.5f + 0.5f * fy);
  const float px[3] = { (t1x - 1.0f) / sx, (t1x + w2x / (w1x + w2x)) / sx, (t1x + 2.0f) / sx };
  const float py[3] = { (t1y - 1.0f) / sy, (t1y + w2y / (w1y + w2y)) / sy, (t1y + 2.0f) / sy };
  const float wx[3] = { w0x, w1x + w2x, w3x }, wy[3] = { w0y, w1y + w2y, w3y };
  float4 res = make_float4(0, 0, 0, 0);
#pragma unroll
  for(int j = 0; j < 3; j++)
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
      const float4 t = tex_rgba(tex, w, h, px[i], py[j]);
      res.x += t.x * wx[i] * wy[j]; res.y += t.y * wx[i] * wy[j];
      res.z += t.z * wx[i] * wy[j]; res.w += t.w * wx[i] * wy[j];
    }
  return res;
}

// crop/main.comp:20-57: output pixel (x,y) -> input texel.  ROT is the shader's `params.r0 != 1.0` branch, hoisted
// to a template parameter by the launcher (it is uniform): the catmull-rom path costs ~200 registers when inlined.
template <bool ROT>
VKB_DEV float4 crop_fetch(const uint2 *__restrict__ in, int iw, int ih, int x, int y, const crop_committed_t &c)
{
  const float tsx = (float)iw, tsy = (float)ih;
  float xx = (float)x + 0.5f, yy = (float)y + 0.5f;
  xx += c.crop[0] * tsx; yy += c.crop[2] * tsy;
  const float dx = xx - tsx * .5f, dy = yy - tsy * .5f;
  xx = c.r[0] * dx + c.r[2] * dy + tsx * .5f;
  yy = c.r[1] * dx + c.r[3] * dy + tsy * .5f;
  const float hx = c.H[0] * xx + c.H[4] * yy + c.H[8];
  const float hy = c.H[1] * xx + c.H[5] * yy + c.H[9];
  const float hz = c.H[2] * xx + c.H[6] * yy + c.H[10];
  float rdx = hx / hz, rdy = hy / hz;
  rdx /= tsx; rdy /= tsy;
  float4 rgba;
  if(rdx < 0.f || rdy < 0.f || rdx >= 1.f || rdy >= 1.f) rgba = make_float4(0, 0, 0, 0);
  else if(ROT) rgba = catmull_rom_rgba(in, iw, ih, rdx, rdy);
  else rgba = ld_rgba_clamp(in, iw, ih, (int)(rdx * tsx), (int)(rdy * tsy));
  rgba.w = 1.0f;
  return rgba;
}

// ---- dt ucs (shared/dtucs.glsl:11-83) ----
VKB_DEV void xyY_to_dt_UCS_JCH(float x, float y, float Y, float L_white, float &J, float &C, float &H)
{
  const float ux = -0.783941002840055f * x + 0.277512987809202f * y + 0.153836578598858f;
  const float uy =  0.745273540913283f * x - 0.205375866083878f * y - 0.165478376301988f;
  const float ud =  0.318707282433486f * x + 2.16743692732158f  * y + 0.291320554395942f;
  const float u = ux / ud, v = uy / ud;
  const float us = 1.39656225667f * u / (fabsf(u) + 1.49217352929f);
  const float vs = 1.4513954287f  * v / (fabsf(v) + 1.52488637914f);
  const float Up = -1.124983854323892f * us - 0.980483721769325f * vs;
  const float Vp =  1.86323315098672f  * us + 1.971853092390862f * vs;
  const float Y_hat = PW_POW(Y, 0.631651345306265f);
  const float L_star = 2.098883786377f * Y_hat / (Y_hat + 1.12426773749357f);
  const float M2 = Up * Up + Vp * Vp;
  J = L_star / L_white;
  C = 15.932993652962535f * PW_POW(L_star, 0.6523997524738018f) * PW_POW(M2, 0.6007557017508491f) / L_white;
  H = atan2f(Vp, Up);
}
VKB_DEV void dt_UCS_JCH_to_xyY(float J, float C, float H, float L_white, float &x, float &y, float &Y)
{
  const float L_star = J * L_white;
  float M = PW_POW(C * L_white / (15.932993652962535f * PW_POW(L_star, 0.6523997524738018f)), 0.8322850678616855f);
  M = clampf(M, 0.0f, 0.05f);
  float sh, ch;
  sincosf(H, &sh, &ch);
  const float a = M * ch, b = M * sh;
  const float us = -5.037522385190711f * a - 2.504856328185843f * b;
  const float vs =  4.760029407436461f * a + 2.874012963239247f * b;
  const float U = -1.49217352929f * us / (fabsf(us) - 1.39656225667f);
  const float V = -1.52488637914f * vs / (fabsf(vs) - 1.4513954287f);
  const float xx = 0.167171472114775f * U + 0.141299802443708f * V - 0.00801531300850582f;
  const float yy = -0.150959086409163f * U - 0.155185060382272f * V - 0.00843312433578007f;
  const float d = 0.940254742367256f * U + 1.000000000000000f * V - 0.0256325967652889f;
  x = xx / d; y = yy / d;
  Y = PW_POW((1.12426773749357f * L_star / (2.098883786377f - L_star)), 1.5831518565279648f);
}

#define M2020_XYZ_00 0.636958048301290991f
#define M2020_XYZ_01 0.144616903586208406f
#define M2020_XYZ_02 0.168880975164172054f
#define M2020_XYZ_10 0.26270021201126692f
#define M2020_XYZ_11 0.677998071518871148f
#define M2020_XYZ_12 0.0593017164698619384f
#define M2020_XYZ_20 4.9999999999999999e-17f
#define M2020_XYZ_21 0.0280726930490874452f
#define M2020_XYZ_22 1.06098505771079066f
VKB_DEV f3 rec2020_to_xyz(f3 c)
{
  return { M2020_XYZ_00 * c.x + M2020_XYZ_01 * c.y + M2020_XYZ_02 * c.z,
           M2020_XYZ_10 * c.x + M2020_XYZ_11 * c.y + M2020_XYZ_12 * c.z,
           M2020_XYZ_20 * c.x + M2020_XYZ_21 * c.y + M2020_XYZ_22 * c.z };
}
VKB_DEV f3 xyz_to_rec2020(f3 c)
{
  return { 1.71665119f * c.x - 0.35567078f * c.y - 0.25336628f * c.z,
          -0.66668435f * c.x + 1.61648124f * c.y + 0.01576855f * c.z,
           0.01763986f * c.x - 0.04277061f * c.y + 0.94210312f * c.z };
}

// ---- colour (main-impl.glsl:200-341, no lut inputs) ----
VKB_DEV float decode_trc(float v, uint32_t trc)
{ // main-impl.glsl:118-150
  switch(trc)
  {
    case 1: { const float a = 1.09929682680944f;
              return v > (float)(0.018053968510807 * 4.5) ? PW_POW((v + (float)(1.09929682680944 - 1.0)) / a, 2.2f) : v / 4.5f; }   // b * 4.5, a - 1: constants, folded in double
    case 2: return v > 0.04045f ? PW_POW((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f;
    case 3: { const float m1 = 1305.0f / 8192.0f, m2 = 2523.0f / 32.0f, c1 = 107.0f / 128.0f, c2 = 2413.0f / 128.0f, c3 = 2392.0f / 128.0f;
              const float xp = PW_POW(fmaxf(0.0f, v), 1.0f / m2);
              return PW_POW(fmaxf(xp - c1, 0.0f) / fmaxf(c2 - c3 * xp, 1e-10f), 1.0f / m1); }
    case 4: return PW_POW(v, 2.6f);
    case 5: { const float a = 0.17883277f, b = 0.28466892f, c = 0.55991073f;
              return v <= 0.5f ? v * v / 3.0f : (PW_EXP((v - c) / a) + b) / 12.0f; }
    case 6: return PW_POW(fmaxf(v, 0.0f), 2.2f);
    // camera log curves to scene linear (shared/oetf.glsl:2-38; mix() with a bvec selects; expressions of literals alone are
    // folded in double and rounded once, like glslang does)
    case 7:  return v > 0.02740668f ? m_exp2(v / 0.07329248f - 7.0f) - 0.0075f : v / 10.44426855f;
    case 8:  return v < 0.075f ? (v - 0.075f) / 16.184376489665897f : PW_EXP((v - 0.5520126568606655f) / 0.09232902596577353f) - 0.0057048244042473785f;
    case 9:  return v <= 0.155251141552511f ? (v - 0.0729055341958355f) / 10.5402377416545f : m_exp2(v * 17.52f - 9.72f);
    case 10: return v < (float)(5.367655 * 0.010591 + 0.092809) ? (v - 0.092809f) / 5.367655f : (PW_POW(10.0f, (v - 0.385537f) / 0.247190f) - 0.052272f) / 5.555556f;
    case 11: return v < -0.7774983977293537f ? v * 0.3033266726886969f - 0.7774983977293537f
                  : (m_exp2(14.0f * (v - 0.09286412512218964f) / 0.9071358748778103f + 6.0f) - 64.0f) / 2231.8263090676883f;
    case 12: return v < 0.0f ? (v / 15.1927f) - 0.01f : (PW_POW(10.0f, v / 0.224282f) - 1.0f) / 155.975327f - 0.01f;
    case 13: return v < 0.181f ? (v - 0.125f) / 5.6f : PW_POW(10.0f, (v - 0.598206f) / 0.241514f) - 0.00873f;
    case 14: return v < (float)(171.2102946929 / 1023.0) ? (v * 1023.0f - 95.0f) * 0.01125f / (float)(171.2102946929 - 95.0)
                  : PW_POW(10.0f, (v * 1023.0f - 420.0f) / 261.5f) * (float)(0.18 + 0.01) - 0.01f;
    case 15: return v < 0.100686685370811f ? (v - 0.092864f) / 8.799461f
                  : PW_POW(10.0f, (v - 0.384316f) / 0.245281f) / 5.555556f - (float)(0.064829 / 5.555556);
    default: return v;
  }
}
VKB_DEV f3 mat3v(const float *M, f3 v)
{ // y = M x, row major, each row summed left to right (matrices.h / o_mat3mulv)
  return { M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z, M[6] * v.x + M[7] * v.y + M[8] * v.z };
}
VKB_DEV f3 colour_matrices(f3 c, const colour_digest_t &p)
{ // primaries -> rec2020, then cat16 (main-impl.glsl:49-67), operation for operation
  if(p.has_P) c = mat3v(p.P, c);
  f3 cl = mat3v(p.MR, c);
  cl.x *= p.ratio[0]; cl.y *= p.ratio[1]; cl.z *= p.ratio[2];
  return mat3v(p.XM, cl);
}
VKB_DEV f3 colour_px(f3 c, const colour_digest_t &p)
{
  if(p.trc) { c.x = decode_trc(c.x, p.trc); c.y = decode_trc(c.y, p.trc); c.z = decode_trc(c.z, p.trc); }
#if VKB_FAST
  f3 o = { p.A[0] * c.x + p.A[1] * c.y + p.A[2] * c.z,
           p.A[3] * c.x + p.A[4] * c.y + p.A[5] * c.z,
           p.A[6] * c.x + p.A[7] * c.y + p.A[8] * c.z };
#else
  f3 o = colour_matrices(c, p);
#endif
  if(p.clip_t > 0.0f) { o.x = fminf(o.x, p.clip_t); o.y = fminf(o.y, p.clip_t); o.z = fminf(o.z, p.clip_t); }
  o.x *= p.exposure; o.y *= p.exposure; o.z *= p.exposure;
  if(p.N > 0)
  {
    f3 co = { p.rbf_P[0] * o.x + p.rbf_P[1] * o.y + p.rbf_P[2] * o.z,
              p.rbf_P[3] * o.x + p.rbf_P[4] * o.y + p.rbf_P[5] * o.z,
              p.rbf_P[6] * o.x + p.rbf_P[7] * o.y + p.rbf_P[8] * o.z };
    for(uint32_t i = 0; i < p.N; i++)
    {
      const float d0 = o.x - p.rbf_p[i][0], d1 = o.y - p.rbf_p[i][1], d2 = o.z - p.rbf_p[i][2];
      const float r = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
      co.x += p.rbf_c[i][0] * r; co.y += p.rbf_c[i][1] * r; co.z += p.rbf_c[i][2] * r;
    }
    o = co;
  }
  if(p.sat != 1.0f)
  {
    o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f);
    const f3 xyz = rec2020_to_xyz(o);
    const float s = xyz.x + xyz.y + xyz.z;
    float J, C, H, x, y, Y;
    xyY_to_dt_UCS_JCH(xyz.x / s, xyz.y / s, xyz.y, 1.0f, J, C, H);
    C = clampf(C * p.sat, 0.0f, 1.0f);
    dt_UCS_JCH_to_xyY(J, C, H, 1.0f, x, y, Y);
    const f3 q = { x * Y / y, y * Y / y, (1.0f - x - y) * Y / y };
    o = xyz_to_rec2020(q);
  }
  o.x = clampf(o.x, -65535.0f, 65535.0f); o.y = clampf(o.y, -65535.0f, 65535.0f); o.z = clampf(o.z, -65535.0f, 65535.0f);
  return o;
}

// ---- filmcurv ----
VKB_DEV float weibull_cdf(float x, float il, float k) { return 1.0f - PW_EXP(-PW_POW(fmaxf(x, 1e-7f) * il, k)); }
VKB_DEV float weibull_pdf(float x, float il, float k)
{
  x = fmaxf(x, 1e-7f);
  return k * il * PW_POW(x * il, k - 1.0f) * PW_EXP(-PW_POW(x * il, k));
}
VKB_DEV float glsl_mod(float x, float y) { return x - y * floorf(x / y); }

VKB_DEV f3 adjust_colour_dng(f3 col0, f3 col1)
{ // shared.glsl:371-387
  bool fx = false, fy = false, fz = false; float t;
#define SWP(a, b) { t = a; a = b; b = t; }
  if(col0.z > col0.y) { SWP(col0.z, col0.y) SWP(col1.z, col1.y) fx = true; }
  if(col0.y > col0.x) { SWP(col0.x, col0.y) SWP(col1.x, col1.y) fy = true; }
  if(col0.z > col0.y) { SWP(col0.z, col0.y) SWP(col1.z, col1.y) fz = true; }
  col1.y = mixf(col1.z, col1.x, m_div(col0.y - col0.z + 1e-6f, col0.x - col0.z + 1e-6f)); // blend factor (fast: 2 ulp is plenty)
  if(fz) SWP(col1.z, col1.y)
  if(fy) SWP(col1.x, col1.y)
  if(fx) SWP(col1.z, col1.y)
#undef SWP
  return col1;
}
VKB_DEV f3 rec2020_to_oklab(f3 c)
{ // colourspaces.glsl:20-33
  float l = 0.61668844f * c.x + 0.36015907f * c.y + 0.02304329f * c.z;
  float m = 0.2651402f  * c.x + 0.63585648f * c.y + 0.09903023f * c.z;
  float s = 0.10015065f * c.x + 0.20400432f * c.y + 0.69632468f * c.z;
  l = PW_POW(fmaxf(0.0f, l), 1.0f / 3.0f); m = PW_POW(fmaxf(0.0f, m), 1.0f / 3.0f); s = PW_POW(fmaxf(0.0f, s), 1.0f / 3.0f);
  return { 0.21045426f * l + 0.79361779f * m - 0.00407205f * s,
           1.9779985f  * l - 2.42859221f * m + 0.45059371f * s,
           0.02590404f * l + 0.78277177f * m - 0.80867577f * s };
}
VKB_DEV f3 oklab_to_rec2020(f3 lab)
{ // colourspaces.glsl:35-49
  float l = 1.0f        * lab.x + 0.39633779f * lab.y + 0.21580376f * lab.z;
  float m = 1.00000001f * lab.x - 0.10556134f * lab.y - 0.06385417f * lab.z;
  float s = 1.00000005f * lab.x - 0.08948418f * lab.y - 1.29148554f * lab.z;
  l = l * l * l; m = m * m * m; s = s * s * s;
  return {  2.14014041f * l - 1.24635595f * m + 0.10643173f * s,
           -0.88483245f * l + 2.16317272f * m - 0.27836159f * s,
           -0.04857906f * l - 0.45449091f * m + 1.50235629f * s };
}
VKB_DEV float lerp_chromaticity_angle(float h1, float h2, float t)
{
  const float delta = h2 - h1;
  if(delta > 0.5f) h2 -= 1.0f;
  else if(delta < -0.5f) h2 += 1.0f;
  return glsl_mod(h1 + t * (h2 - h1), 1.0f);
}
VKB_DEV float hue_bump(float h, float h0, float w)
{
  const float pi = 3.14159265358979323846f;
  const float d = fabsf(glsl_mod(h - h0 + pi, 2.0f * pi) - pi);
  return d < w ? 0.5f + 0.5f * cosf(pi * d / w) : 0.0f;
}

// ---- munsell hue lines (shared/munsell.glsl:7-137): chromaticity <-> (hue, chroma) on the renotation grid ----
static __device__ const uint32_t munsell_xy[VKB_MUNSELL_HDIM * VKB_MUNSELL_CDIM] = { VKB_MUNSELL_WORDS };
#define MUN_ILLX 0.31271f
#define MUN_ILLY 0.32902f
VKB_DEV float2 munsell_lookup(int hue_idx, int chroma_idx)
{
  hue_idx = (hue_idx % VKB_MUNSELL_HDIM + VKB_MUNSELL_HDIM) % VKB_MUNSELL_HDIM;
  chroma_idx = min(max(chroma_idx, 0), VKB_MUNSELL_CDIM - 1);
  const uint32_t w = __ldg(munsell_xy + VKB_MUNSELL_CDIM * hue_idx + chroma_idx);
  return make_float2(__half2float(__ushort_as_half((unsigned short)(w & 0xffffu))), __half2float(__ushort_as_half((unsigned short)(w >> 16))));
}
VKB_DEV float munsell_hue_angle(float2 xy)
{ // grows monotonically with the hue index, from zero at hue 0: what the search below needs
  const float pi = 3.14159265358979323846f;
  return glsl_mod(2.0f * pi - 2.52f - atan2f(xy.y - MUN_ILLY, xy.x - MUN_ILLX), 2.0f * pi);
}
VKB_DEV float munsell_side(float2 v0, float2 v1, float2 p)
{ // which side of the line v0--v1 p lies on
  const float ax = v1.x - v0.x, ay = v1.y - v0.y, bx = p.x - v0.x, by = p.y - v0.y;
  return ax * by - ay * bx;
}
VKB_DEV float2 munsell_to_xy(float2 mhc)
{
  const float hm = mhc.x * VKB_MUNSELL_HDIM, cm = fmaxf(mhc.y, 0.0f) * VKB_MUNSELL_CDIM;
  const int hi = (int)hm, ci = (int)cm;
  const float hu = hm - (float)hi, cu = cm - (float)ci;
  const float2 r3 = munsell_lookup(hi, ci + 1), r2 = munsell_lookup(hi + 1, ci + 1);
  const float2 r0 = munsell_lookup(hi, ci),     r1 = munsell_lookup(hi + 1, ci);
  if(hu >= cu) return make_float2((1.0f - hu) * r0.x + (hu - cu) * r1.x + cu * r2.x, (1.0f - hu) * r0.y + (hu - cu) * r1.y + cu * r2.y);
  return make_float2(hu * r2.x + (cu - hu) * r3.x + (1.0f - cu) * r0.x, hu * r2.y + (cu - hu) * r3.y + (1.0f - cu) * r0.y);
}
VKB_DEV float2 munsell_from_xy(float2 xy)
{
  int hm = 0, hM = VKB_MUNSELL_HDIM, cm = 0, cM = VKB_MUNSELL_CDIM - 1;
  const float theta = munsell_hue_angle(xy);
  const float dx = xy.x - MUN_ILLX, dy = xy.y - MUN_ILLY;
  const float rad2 = dx * dx + dy * dy;
  for(int i = 0; i < 10; i++)
  { // bisection on hue angle and on distance from the white point
    const int h = (hm + hM) / 2, c = (cm + cM) / 2;
    const float2 res = munsell_lookup(h, c);
    const float th = munsell_hue_angle(res);
    const float ex = res.x - MUN_ILLX, ey = res.y - MUN_ILLY;
    const float r2 = ex * ex + ey * ey;
    if(th <= theta) hm = h; else hM = h;
    if(r2 <= rad2)  cm = c; else cM = c;
    if(hM <= hm + 1 && cM <= cm + 1) break;
  }
  for(int i = 0; i < 10; i++)
  { // the grid is not polar: walk to the cell that contains xy
    const float2 r3 = munsell_lookup(hm, cm + 1), r2 = munsell_lookup(hm + 1, cm + 1);
    const float2 r0 = munsell_lookup(hm, cm),     r1 = munsell_lookup(hm + 1, cm);
    const float s0 = munsell_side(r0, r1, xy), s1 = munsell_side(r1, r2, xy);
    const float s2 = munsell_side(r2, r3, xy), s3 = munsell_side(r3, r0, xy);
    if(s0 < 0.0f && cm > 0) cm--;
    else if(s0 < 0.0f && cm == 0) hm = ((hm + VKB_MUNSELL_HDIM / 2) % VKB_MUNSELL_HDIM + VKB_MUNSELL_HDIM) % VKB_MUNSELL_HDIM;
    else if(s2 < 0.0f && cm < VKB_MUNSELL_CDIM - 2) cm++;
    if(s1 < 0.0f) hm++;
    else if(s3 < 0.0f) hm--;
    if(s0 >= 0.0f && s1 >= 0.0f && s3 >= 0.0f && (s2 >= 0.0f || cm >= VKB_MUNSELL_CDIM - 2))
    { // inside: barycentric coordinates in the triangle, interpolating the (stepped, like the shader's) corner indices
      const float t0 = munsell_side(r0, r1, r2), t1 = munsell_side(r2, r3, r0);
      float u0, u1, u2, u3;
      if(cm > 0 && s0 + s1 <= t0) { u2 = s0 / t0; u0 = s1 / t0; u1 = 1.0f - u0 - u2; u3 = 0.0f; }
      else                        { u2 = s3 / t1; u0 = s2 / t1; u3 = 1.0f - u0 - u2; u1 = 0.0f; }
      const float fh = (float)hm, fc = (float)cm;
      const float hi = u0 * fh + u1 * (fh + 1.0f) + u2 * (fh + 1.0f) + u3 * fh;
      const float ci = u0 * fc + u1 * fc + u2 * (fc + 1.0f) + u3 * (fc + 1.0f);
      return make_float2(hi / VKB_MUNSELL_HDIM, fmaxf(0.0f, ci / VKB_MUNSELL_CDIM));
    }
  }
  return make_float2(1.0f, 1.0f);
}

VKB_DEV f3 filmcurv_px(f3 in, const filmcurv_params_t &p)
{ // filmcurv/main.comp:68-166
  const float il = fmaxf(5e-3f, p.light);
  const float k  = fmaxf(1e-4f, p.contrast);
  const f3 col0 = { in.x + p.bias, in.y + p.bias, in.z + p.bias };
  f3 col1 = { weibull_cdf(col0.x, il, k), weibull_cdf(col0.y, il, k), weibull_cdf(col0.z, il, k) };
  if(p.colour == 3) return adjust_colour_dng(col0, col1);
  if(p.colour == 1) return col1;
  if(p.colour == 0)
  {
    const f3 xyz0 = rec2020_to_xyz(col0), xyz1 = rec2020_to_xyz(col1);
    const float s0 = fmaxf(1e-4f, xyz0.x + xyz0.y + xyz0.z), s1 = fmaxf(1e-4f, xyz1.x + xyz1.y + xyz1.z);
    float J0, C0, H0, J1, C1, H1, x, y, Y;
    xyY_to_dt_UCS_JCH(xyz0.x / s0, xyz0.y / s0, xyz0.y, 1.0f, J0, C0, H0);
    xyY_to_dt_UCS_JCH(xyz1.x / s1, xyz1.y / s1, xyz0.y, 1.0f, J1, C1, H1);
    dt_UCS_JCH_to_xyY(J1, C1, H0, 1.0f, x, y, Y);
    const float m = fmaxf(1e-4f, y);
    const f3 q = { x * xyz1.y / m, y * xyz1.y / m, (1.0f - x - y) * xyz1.y / m };
    return xyz_to_rec2020(q);
  }
  if(p.colour == 2)
  { // hue of the input, chroma of the per channel curve, along munsell's hue lines (main.comp:96-104, colourspaces.glsl:2-19)
    const f3 xyz0 = rec2020_to_xyz(col0), xyz1 = rec2020_to_xyz(col1);
    const float s0 = 1.0f * xyz0.x + 1.0f * xyz0.y + 1.0f * xyz0.z, s1 = 1.0f * xyz1.x + 1.0f * xyz1.y + 1.0f * xyz1.z;
    const float2 m0 = munsell_from_xy(make_float2(xyz0.x / s0, xyz0.y / s0));
    const float2 m1 = munsell_from_xy(make_float2(xyz1.x / s1, xyz1.y / s1));
    const float2 xy = munsell_to_xy(make_float2(m0.x, m1.y));
    const float Y = xyz1.y;
    return xyz_to_rec2020({ xy.x * Y / xy.y, xy.y * Y / xy.y, (1.0f - xy.x - xy.y) * Y / xy.y });
  }
  if(p.colour == 4)
  { // agx
    const float twopi = 2.0f * 3.14159265358979323846f;
    f3 c = { 0.856627153315983f * col0.x + 0.0951212405381588f * col0.y + 0.0482516061458583f * col0.z,
             0.137318972929847f * col0.x + 0.761241990602591f  * col0.y + 0.101439036467562f  * col0.z,
             0.11189821299995f  * col0.x + 0.0767994186031903f * col0.y + 0.811302368396859f  * col0.z };
    const f3 lab0 = rec2020_to_oklab(c);
    const float t0 = 1.0f + atan2f(lab0.z, lab0.y) / twopi;
    const float h0 = t0 - floorf(t0);
    c = { weibull_cdf(c.x, il, k), weibull_cdf(c.y, il, k), weibull_cdf(c.z, il, k) };
    const f3 lab1 = rec2020_to_oklab(c);
    const float t1 = 1.0f + atan2f(lab1.z, lab1.y) / twopi;
    float h1 = t1 - floorf(t1);
    const float C = sqrtf(lab1.y * lab1.y + lab1.z * lab1.z);
    h1 = lerp_chromaticity_angle(h0, h1, 0.4f);
    float sh, ch;
    sincosf(twopi * h1, &sh, &ch);
    f3 r = { 0, 0, 0 };
    if(!(lab1.x <= 0.0f)) r = oklab_to_rec2020({ lab1.x, C * ch, C * sh });
    return {  1.1271005818144368f * r.x - 0.11060664309660323f * r.y - 0.016493938717834573f * r.z,
             -0.1413297634984383f * r.x + 1.157823702216272f   * r.y - 0.016493938717834257f * r.z,
             -0.14132976349843826f * r.x - 0.11060664309660294f * r.y + 1.2519364065950405f * r.z };
  }
  if(p.colour == 5)
  {
    const float pi = 3.14159265358979323846f;
    const f3 lab0 = rec2020_to_oklab(col0);
    const float L0 = fmaxf(lab0.x, 1e-7f);
    const float lum0 = fmaxf(col0.x * 0.2627f + col0.y * 0.6780f + col0.z * 0.0593f, 1e-7f);
    float lum1 = weibull_cdf(lum0, il, k);
    lum1 = lum1 + p.rolloff * lum1 * lum1 * (1.0f - lum1);
    if(p.shadows != 0.0f)
    {
      const float toe_gamma = 1.0f - 0.5f * p.shadows;
      const float lum_toe = PW_POW(fmaxf(lum1, 1e-7f), toe_gamma);
      lum1 = mixf(lum1, lum_toe, smoothstepf(0.3f, 0.0f, lum1));
    }
    const float L1 = L0 * PW_POW(lum1 / lum0, 1.0f / 3.0f);
    const f3 lab_pc = rec2020_to_oklab(col1);
    float C1 = sqrtf(lab_pc.y * lab_pc.y + lab_pc.z * lab_pc.z);
    const float tame = 1.0f - 0.6f * p.rolloff * smoothstepf(0.15f, 0.5f, lum1);
    C1 *= mixf(tame, 1.0f, clampf(p.chroma - 1.0f, 0.0f, 1.0f));
    float h = atan2f(lab_pc.z, lab_pc.y);
    const float h_target = 0.96f;
    const float h_dist = fabsf(glsl_mod(h - h_target + pi, 2.0f * pi) - pi);
    if(h_dist < 0.7f)
    {
      const float away = fabsf(lum1 - 0.35f);
      h = lerp_chromaticity_angle(h, h_target, 0.3f * smoothstepf(0.0f, 0.3f, away));
    }
    const float c = p.chroma - 1.0f;
    const float deriv = weibull_pdf(lum0, il, k);
    const float hi = PW_POW(fmaxf(1.0f, 1.0f / fmaxf(deriv, 0.15f)), c * 0.1f);
    const float lo = 1.0f + c * 0.15f * smoothstepf(0.3f, 0.0f, lum1);
    float cr = p.chroma * hi * lo;
    cr *= 1.0f + p.red * hue_bump(h, 0.7f, 1.0f) + p.yellow * hue_bump(h, 1.76f, 1.0f) + p.blue * hue_bump(h, -1.76f, 1.0f);
    float sh, ch;
    sincosf(h, &sh, &ch);
    return oklab_to_rec2020({ L1, C1 * cr * ch, C1 * cr * sh });
  }
  return { 0, 0, 0 }; // colour == 2 (munsell lut) is out of scope
}

// ---- colenc (colenc/main.comp:17-81): rec2020 -> output primaries, output transfer curve ----
// constants: glslang folds constant expressions in double and rounds once
VKB_DEV f3 colenc_px(f3 c, const colenc_params_t &p)
{
  if(p.prim == 1)       c = f3{ 1.66022677f * c.x - 0.58754761f * c.y - 0.07283825f * c.z, -0.12455334f * c.x + 1.13292605f * c.y - 0.00834963f * c.z, -0.01815514f * c.x - 0.10060303f * c.y + 1.11899817f * c.z };
  else if(p.prim == 3)  c = f3{ 1.15194302f * c.x - 0.09753232f * c.y - 0.05448118f * c.z, -0.12454585f * c.x + 1.13290963f * c.y - 0.00837122f * c.z, -0.02253539f * c.x - 0.04979918f * c.y + 1.07275365f * c.z };
  else if(p.prim == 4)  c = f3{ 1.34353337f * c.x - 0.28218904f * c.y - 0.06142427f * c.z, -0.06530851f * c.x + 1.07578268f * c.y - 0.01048453f * c.z, 0.00282971f * c.x - 0.01961215f * c.y + 1.01717851f * c.z };
  else if(p.prim == 5)  c = rec2020_to_xyz(c);
  else if(p.prim == 6)  c = f3{ 6.68685575e-01f * c.x + 1.51817679e-01f * c.y + 1.77189677e-01f * c.z, 4.49002044e-02f * c.x + 8.62145497e-01f * c.y + 1.01922441e-01f * c.z, -2.66851927e-09f * c.x + 2.78271109e-02f * c.y + 1.05170358f * c.z };
  else if(p.prim == 7)  c = f3{ 9.62918591e-01f * c.x + 1.16137050e-02f * c.y + 2.55863361e-02f * c.z, 4.16800770e-04f * c.x + 9.99378426e-01f * c.y - 8.82457347e-05f * c.z, 5.31123331e-03f * c.x + 2.18655328e-02f * c.y + 9.75907920e-01f * c.z };
  else if(p.prim == 10) c = f3{ 0.853263f * c.x + 0.079695f * c.y + 0.067042f * c.z, 0.029375f * c.x + 0.809195f * c.y + 0.161430f * c.z, 0.051575f * c.x + 0.208097f * c.y + 0.740329f * c.z };
  float v[3] = { c.x, c.y, c.z };
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    float t = v[k];
    if(p.trc == 1)
    {
      const float a = 1.09929682680944f, b = 0.018053968510807f;
      t = t > b ? PW_POW(t, (float)(1.0 / 2.2)) * a - (float)(1.09929682680944 - 1.0) : t * 4.5f;   // a - 1 is a constant: folded in double, rounded once (glslang)
    }
    else if(p.trc == 2) t = t > 0.0031308f ? PW_POW(t, (float)(1.0 / 2.4)) * 1.055f - 0.055f : t * 12.92f;
    else if(p.trc == 3)
    {
      const float c3 = (float)(2392.0 / 128.0), c2 = (float)(2413.0 / 128.0), c1 = c3 - c2 + 1.0f;
      const float m1 = (float)(1305.0 / 8192.0), m2 = (float)(2523.0 / 32.0);
      t = fmaxf(0.0f, t);
      t = PW_POW(t, m1);
      const float num = (c1 - 1.0f) + (c2 - c3) * t, den = 1.0f + c3 * t;
      t = PW_POW(1.0f + num / den, m2);
    }
    else if(p.trc == 4) t = PW_POW(t, (float)(1.0 / 2.6));
    else if(p.trc == 5)
    {
      const float a = 0.17883277f, b = 1.0f - 4.0f * a, cc = 0.5f - a * -0.33500978350639343f; // c = 0.5 - a log(4a): logf(4 * 0.17883277f) = -0.33500978 (libm)
      t = t > (float)(1.0 / 12.0) ? a * m_log(12.0f * t - b) + cc : sqrtf(3.0f * t);
    }
    else if(p.trc == 6) t = PW_POW(t, (float)(1.0 / 2.2));
    v[k] = t;
  }
  return { v[0], v[1], v[2] };
}

// ---- grade (grade/main.comp:21-62) ----
VKB_DEV f3 grade_px(f3 c, const grade_params_t &q)
{
  const float lift[3] = { q.lift[0] + q.lift[3], q.lift[1] + q.lift[3], q.lift[2] + q.lift[3] };
  const float gam[3]  = { fmaxf(q.gamma[0] + q.gamma[3], 1e-6f), fmaxf(q.gamma[1] + q.gamma[3], 1e-6f), fmaxf(q.gamma[2] + q.gamma[3], 1e-6f) };
  const float gain[3] = { fmaxf(q.gain[0] + q.gain[3], 0.0f), fmaxf(q.gain[1] + q.gain[3], 0.0f), fmaxf(q.gain[2] + q.gain[3], 0.0f) };
  const float off[3]  = { q.off[0] + q.off[3], q.off[1] + q.off[3], q.off[2] + q.off[3] };
  float v[3] = { c.x, c.y, c.z };
  if(q.mode == 0)
  {
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      float t = gain[k] * v[k];
      t = t * (1.0f - lift[k]) + lift[k];
      const float ig = 1.0f / gam[k];
      t = fmaxf(t, 0.0f);
      t = (ig == 1.0f) ? t : PW_POW(t, ig); // pow(x, 1) == x exactly
      v[k] = t + off[k];
    }
  }
  else
  {
    float L = fmaxf(v[0], 0.0f) * 0.2126f + fmaxf(v[1], 0.0f) * 0.7152f + fmaxf(v[2], 0.0f) * 0.0722f;
    L = clampf(0.67f + m_log2(fmaxf(L, 1e-6f)) * 0.11f, 0.0f, 1.0f);
    const float sp = clampf(q.sh_pivot, 1e-3f, 1.0f - 1e-3f);
    const float hp = clampf(q.hi_pivot, sp + 1e-3f, 1.0f);
    const float w_s = 1.0f - smoothstepf(0.0f, sp, L);
    const float w_h = smoothstepf(hp, 1.0f, L);
    const float w_m = 1.0f - w_s - w_h;
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      const float ge = mixf(1.0f, gain[k], w_h), le = lift[k] * w_s, me = mixf(1.0f, gam[k], w_m);
      float t = ge * v[k];
      t = t * (1.0f - le) + le;
      t = PW_POW(fmaxf(t, 0.0f), 1.0f / me);
      v[k] = t + off[k];
    }
  }
  return { v[0], v[1], v[2] };
}

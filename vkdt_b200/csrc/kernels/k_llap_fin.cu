// finest level of the local laplacian pyramid, fused with llap/colour.comp and (optionally) grade/main.comp:
//   out = grade( recolour( expand(coarse) + blend of the two laplacians bracketing the local grey value ) )
// (llap/assemble.comp:52-88, llap/colour.comp:17-37, grade/main.comp:21-62).
// B200 shape: the three gauss_expand()s a pixel needs read 5x5 coarse texels each from planes that depend on the
// pixel's grey value.  a CTA of 32x8 output pixels stages the 20x8 coarse window of ALL 12 planes (11 gamma layers +
// the collapsed coarse level) in shared memory once (7.5 global loads per pixel instead of ~60 with per-tap address
// mirroring), and evaluates sample_soft's 3x3 bilinear taps in their separable form [1 1 2 1 1]/6 | [2 1 1 2]/6.
// the separable sum differs from the shader's 9-tap order by fp32 rounding only (~1e-7), far below the f16 store
// that follows; nothing downstream of this kernel feeds a pyramid level, so the difference cannot compound.
// level-0 gamma layers are never read: curve() is recomputed from the input pixel and rounded to f16 in registers.
#include <string.h>
#include "pointwise.cuh"

#define NUM_GAMMA 10
#define NL (NUM_GAMMA + 1)

struct llap_params_t { float sigma, shadows, hilights, clarity; };
// everything below `grade` is a function of the launch's parameters only and is evaluated once on the host with the same
// fp32 operations the kernel used to repeat per pixel (1/(2 sigma), 1/(2 sigma^2/3); grade: lift, 1 - lift, gain, offset, 1/gamma)
struct llap_rd_t { double rd2s, rdk; };   // 1 / (2 sigma), 1 / (2 sigma^2 / 3) for div_rd (strict)
struct llapfin_t { llap_rd_t rd; double rdg[NUM_GAMMA];   // rdg[i] = 1 / (gamma[i] - gamma[i - 1]): the blend weight's quotient (strict)
                   llap_params_t p; int first; int have_grade; int out_f32; grade_params_t grade;
                   float inv2s, invd; float g_lift[3], g_oml[3], g_gain[3], g_off[3], g_ig[3]; };

VKB_DEV float gamma_from_i(int i) { return div_c((float)i, NUM_GAMMA - 1.0f); }
// llap.glsl:17-22 gamma_hi_from_v without the loop of divisions: 1 + #{ i in 1..8 : i/9 <= v } (the i/9 are compile time constants)
VKB_DEV int gamma_hi(float v)
{
  int hi = 1;
#pragma unroll
  for(int i = 1; i < NUM_GAMMA - 1; i++) hi += ((float)i / (NUM_GAMMA - 1.0f) <= v) ? 1 : 0;
  return hi;
}
// curve() with the two divisions by per-launch constants turned into multiplications (1 ulp of difference in t and
// in the exponent; the result is rounded to f16 right after).  k.x = 1/(2 sigma), k.y = 1/(2 sigma^2 / 3)
VKB_DEV float llap_curve_k(float x, float g, const llap_params_t &p, float inv2s, float invd)
{
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? p.sigma : -p.sigma;
  const float shadhi = c > 0.0f ? p.shadows : p.hilights;
  if(fabsf(c) > 2 * p.sigma) val = g + ssigma + shadhi * (c - ssigma);
  else
  {
    const float t = clampf(fabsf(c) * inv2s, 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  val += p.clarity * c * exp_ftz(-c * c * invd);
  return val;
}
VKB_DEV float llap_curve(float x, float g, const llap_params_t &p)
{ // llap/curve.comp:40-63
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? p.sigma : -p.sigma;
  const float shadhi = c > 0.0f ? p.shadows : p.hilights;
  if(fabsf(c) > 2 * p.sigma) val = g + ssigma + shadhi * (c - ssigma);
  else
  {
    const float t = clampf(c / (2.0f * ssigma), 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  // the gaussian term is < 3% of val: __expf's 1e-6 relative error on it is below an fp32 ulp of val
  val += p.clarity * c * m_exp(-c * c / (2.0f * p.sigma * p.sigma / 3.0f));
  return val;
}

// strict: the two quotients by launch constants through div_rd, libm's exponential with its tables in shared memory
VKB_DEV float llap_curve_x(float x, float g, const llap_params_t &p, const llap_rd_t &R, const lme_ctx_t &L)
{
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? p.sigma : -p.sigma;
  const float shadhi = c > 0.0f ? p.shadows : p.hilights;
  if(fabsf(c) > 2 * p.sigma) val = g + ssigma + shadhi * (c - ssigma);
  else
  {
    const float t = clampf(div_rd(c, c > 0.0f ? R.rd2s : -R.rd2s), 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  val += p.clarity * c * m_exp_s(div_rd(-c * c, R.rdk), L);
  return val;
}

// ---- 2x2 output pixels per thread ----
// the four pixels (2k, 2k+1) x (2m, 2m+1) expand the same 5x5 coarse texels around (k, m); with the usual case of all
// four picking the same pair of gamma layers one 5x5 window per plane serves four outputs, and the separable row sums
// (even weights [1 1 2 1 1]/2, odd weights [0 2 1 1 2]/2) are shared between the two columns / rows.
// only the gamma layers the CTA's pixels actually select (+ the collapsed coarse level) are staged in shared memory.
#define F3_W 36
#define F3_H 12
VKB_DEV void expand4(const float (*T)[F3_W + 1], int lx, int ly, float &ee, float &oe, float &eo, float &oo)
{ // ee: x even y even, oe: x odd y even, eo: x even y odd, oo: both odd
  float he[5], ho[5];
#pragma unroll
  for(int r = 0; r < 5; r++)
  {
    const float *row = T[ly - 2 + r] + lx - 2;
    const float t0 = row[0], t1 = row[1], t2 = row[2], t3 = row[3], t4 = row[4];
    he[r] = 0.5f * (t0 + t1) + t2 + 0.5f * (t3 + t4);
    ho[r] = t1 + 0.5f * (t2 + t3) + t4;
  }
  ee = div9(0.5f * (he[0] + he[1]) + he[2] + 0.5f * (he[3] + he[4]));
  oe = div9(0.5f * (ho[0] + ho[1]) + ho[2] + 0.5f * (ho[3] + ho[4]));
  eo = div9(he[1] + 0.5f * (he[2] + he[3]) + he[4]);
  oo = div9(ho[1] + 0.5f * (ho[2] + ho[3]) + ho[4]);
}

VKB_DEV float expand1(const float (*T)[F3_W + 1], int lx, int ly, int q)
{ // one output of the 2x2 (q&1: x odd, q>>1: y odd): used when the four pixels bracket different gamma layers
  const bool ox = q & 1, oy = q >> 1;
  const float wx[5] = { ox ? 0.0f : 0.5f, ox ? 1.0f : 0.5f, ox ? 0.5f : 1.0f, 0.5f, ox ? 1.0f : 0.5f };
  const float wy[5] = { oy ? 0.0f : 0.5f, oy ? 1.0f : 0.5f, oy ? 0.5f : 1.0f, 0.5f, oy ? 1.0f : 0.5f };
  float acc = 0.0f;
#pragma unroll
  for(int r = 0; r < 5; r++)
  {
    const float *row = T[ly - 2 + r] + lx - 2;
    acc += (row[0] * wx[0] + row[1] * wx[1] + row[2] * wx[2] + row[3] * wx[3] + row[4] * wx[4]) * wy[r];
  }
  return div9(acc);
}

// strict build: sample_soft's nine bilinear taps in the shader's order (as k_llap_asm4.cu's expand_q), from the 5x5 window
VKB_DEV constexpr int  tap_i0(int d, int t)   { return d ? (t == 0 ? 1 : (t == 1 ? 2 : 4)) : (t == 0 ? 0 : (t == 1 ? 2 : 3)); }
VKB_DEV constexpr bool tap_half(int d, int t) { return d ? t == 1 : t != 1; }
template <int DX, int DY>
VKB_DEV float expand_q(const float (&W)[5][5])
{
  float r = 0.0f;
#pragma unroll
  for(int j = 0; j < 3; j++)
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
      const int x0 = tap_i0(DX, i), y0 = tap_i0(DY, j);
      float top = tap_half(DX, i) ? W[y0][x0] * 0.5f + W[y0][x0 + 1] * 0.5f : W[y0][x0];
      float v;
      if(tap_half(DY, j))
      {
        const float bot = tap_half(DX, i) ? W[y0 + 1][x0] * 0.5f + W[y0 + 1][x0 + 1] * 0.5f : W[y0 + 1][x0];
        v = top * 0.5f + bot * 0.5f;
      }
      else v = top;
      r += v;
    }
  return div9(r);
}
VKB_DEV void expand4_exact(const float (*T)[F3_W + 1], int lx, int ly, float *t)
{
  float W[5][5];
#pragma unroll
  for(int r = 0; r < 5; r++)
#pragma unroll
    for(int c = 0; c < 5; c++) W[r][c] = T[ly - 2 + r][lx - 2 + c];
  t[0] = expand_q<0, 0>(W); t[1] = expand_q<1, 0>(W); t[2] = expand_q<0, 1>(W); t[3] = expand_q<1, 1>(W);
}

// grade/main.comp:21-40 (mode 0) on the host-evaluated constants of llapfin_t: same operations, same order as grade_px()
// (the other grading modes stay out of line: inlined, their quotients, logarithms and powers cost the default path registers)
static __device__ __noinline__ f3 grade_px_other(f3 c, const grade_params_t &g) { return grade_px(c, g); }
// PLAIN: mode 0 with gamma 1 in every channel (the default parameters, checked by the launcher): no power, no other mode, and
// therefore no out of line call in the kernel
template <bool PLAIN>
VKB_DEV f3 grade_px_digest(f3 c, const llapfin_t &P)
{
  if(!PLAIN && P.grade.mode != 0) return grade_px_other(c, P.grade);
  float v[3] = { c.x, c.y, c.z };
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    float t = P.g_gain[k] * v[k];
    t = t * P.g_oml[k] + P.g_lift[k];
    t = fmaxf(t, 0.0f);
    if(!PLAIN) t = (P.g_ig[k] == 1.0f) ? t : PW_POW(t, P.g_ig[k]);
    v[k] = t + P.g_off[k];
  }
  return { v[0], v[1], v[2] };
}

// GRADE: 0 no grade module behind llap, 1 grade with its default shape (mode 0, gamma 1: grade_px_digest<true>), 2 any grade
template <bool F32, int GRADE>
__global__ void __launch_bounds__(256, 4) k_llap_final4(const uint2 *__restrict__ in, const __half *__restrict__ coarse,
    const __half *__restrict__ l1, int cw, int ch, void *__restrict__ outv, int ow, int oh, const __grid_constant__ llapfin_t P, const band_t bd)
{
  __shared__ float tile[NL + 1][F3_H][F3_W + 1];
  __shared__ int s_pmin, s_pmax;
  __shared__ float s_gamma[NUM_GAMMA]; // i / 9: the bracket values by index, instead of a division per pixel and bracket
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if(tid == 0) { s_pmin = NUM_GAMMA; s_pmax = 0; }
  if(tid < NUM_GAMMA) s_gamma[tid] = gamma_from_i(tid);
  LME_SMEM_STAGE(tid);
  const int kx = blockIdx.x * 32 + threadIdx.x, ky = BAND_BY * 8 + threadIdx.y;
  const int cx0 = blockIdx.x * 32 - 2, cy0 = BAND_BY * 8 - 2;
  const size_t p1 = (size_t)cw * ch;
  // 1) the four input pixels, their grey value and gamma bracket
  float4 px[4]; float grey[4], v[4]; int hi[4];
  int mylo = NUM_GAMMA, myhi = 0;
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    const int x = 2 * kx + (q & 1), y = 2 * ky + (q >> 1);
    hi[q] = -1;
    if(x < ow && y < oh && !BAND_SKIP(ky))
    {
      px[q] = ld_rgba(in, ow, x, y);
      grey[q] = lum2020(clampf(px[q].x, -1000.0f, 1000.0f), clampf(px[q].y, -1000.0f, 1000.0f), clampf(px[q].z, -1000.0f, 1000.0f));
      v[q] = f16r(grey[q]);
      const int h = gamma_hi(v[q]);
      hi[q] = h;
      mylo = min(mylo, h - 1); myhi = max(myhi, h);
    }
  }
  __syncthreads();
  mylo = __reduce_min_sync(0xffffffffu, mylo); myhi = __reduce_max_sync(0xffffffffu, myhi);
  if(threadIdx.x == 0) { atomicMin(&s_pmin, mylo); atomicMax(&s_pmax, myhi); }
  __syncthreads();
  const int pmin = s_pmin, pmax = s_pmax;
  // 2) stage the needed planes: gamma layers pmin..pmax and the coarse level (slot NL)
  const bool big = cw >= 40 && ch >= 16;
  const int nplanes = pmax >= pmin ? pmax - pmin + 2 : 0;
  // each thread owns <= 2 texel positions of the 36x12 window; the mirrored offset is computed once and reused per plane
#pragma unroll
  for(int e = 0; e < 2; e++)
  {
    const int t = tid + e * 256;
    if(t >= F3_H * F3_W) break;
    const int r = t / F3_W, c = t - r * F3_W;
    const int gx = big ? mirror1(cx0 + c, cw) : mirrori(cx0 + c, cw), gy = big ? mirror1(cy0 + r, ch) : mirrori(cy0 + r, ch);
    const size_t off = (size_t)gy * cw + gx;
    if(nplanes) tile[NL][r][c] = __half2float(__ldg((P.first ? l1 + NUM_GAMMA * p1 : coarse) + off));
    for(int pl = pmin; pl <= pmax; pl++) tile[pl][r][c] = __half2float(__ldg(l1 + pl * p1 + off));
  }
  __syncthreads();
  if(hi[0] < 0) return; // whole 2x2 outside the image
  const float inv2s = P.inv2s, invd = P.invd;
  const int lx = kx - cx0, ly = ky - cy0;
  float res[4], e0[4], e1[4];
  float oc[6]; // packed rgb sink: the thread's two pixels of a row, stored by the warp together
  __shared__ __align__(16) float stage[8][196];
#if VKB_FAST
  expand4(tile[NL], lx, ly, res[0], res[1], res[2], res[3]);
  const bool same = (hi[1] < 0 || hi[1] == hi[0]) && (hi[2] < 0 || hi[2] == hi[0]) && (hi[3] < 0 || hi[3] == hi[0]);
  if(same)
  {
    expand4(tile[hi[0] - 1], lx, ly, e0[0], e0[1], e0[2], e0[3]);
    expand4(tile[hi[0]],     lx, ly, e1[0], e1[1], e1[2], e1[3]);
  }
  else
  {
#pragma unroll
    for(int q = 0; q < 4; q++) if(hi[q] >= 0)
    {
      e0[q] = expand1(tile[hi[q] - 1], lx, ly, q);
      e1[q] = expand1(tile[hi[q]],     lx, ly, q);
    }
  }
#else
  { // the collapsed coarse level, then every gamma layer one of the four pixels brackets, each in the shader's tap order
    int hmin = NUM_GAMMA, hmax = 0;
#pragma unroll
    for(int q = 0; q < 4; q++) if(hi[q] >= 0) { hmin = min(hmin, hi[q]); hmax = max(hmax, hi[q]); }
    for(int pl = hmin - 2; pl <= hmax; pl++)
    {
      float t[4];
      expand4_exact(tile[pl == hmin - 2 ? NL : pl], lx, ly, t);
#pragma unroll
      for(int q = 0; q < 4; q++)
      {
        if(pl == hmin - 2) res[q] = t[q];
        else
        {
          if(pl == hi[q] - 1) e0[q] = t[q];
          if(pl == hi[q])     e1[q] = t[q];
        }
      }
    }
  }
#endif
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    if(hi[q] >= 0)
    {
    const int x = 2 * kx + (q & 1), y = 2 * ky + (q >> 1);
    const float glo = s_gamma[hi[q] - 1], ghi = s_gamma[hi[q]];
#if VKB_FAST
    const float a = clampf(__fdividef(v[q] - glo, ghi - glo), 0.0f, 1.0f);
    const float lap0 = f16r(llap_curve_k(grey[q], glo, P.p, inv2s, invd)) - e0[q];
    const float lap1 = f16r(llap_curve_k(grey[q], ghi, P.p, inv2s, invd)) - e1[q];
    float l = f16r(res[q] + lap0 * (1.0f - a) + lap1 * a);
    const float yo = fmaxf(lum2020(px[q].x, px[q].y, px[q].z), 1e-8f);
    if(l < yo) l = yo * exp_ftz(l - yo);
    const float ratio = __fdividef(l, yo); // nothing downstream but one f16/f32 store: 2 ulp is plenty
    f3 c = { fmaxf(0.0f, px[q].x * ratio), fmaxf(0.0f, px[q].y * ratio), fmaxf(0.0f, px[q].z * ratio) };
#else
    // assemble.comp:66-87 and colour.comp:23-35 operation for operation
    const float a = clampf(div_rd(v[q] - glo, P.rdg[hi[q]]), 0.0f, 1.0f);
    const float lap0 = f16r(llap_curve_x(grey[q], glo, P.p, P.rd, lme_ctx)) - e0[q];
    const float lap1 = f16r(llap_curve_x(grey[q], ghi, P.p, P.rd, lme_ctx)) - e1[q];
    float l = f16r(res[q] + lap0 * (1.0f - a) + lap1 * a);
    const float yo = fmaxf(lum2020(px[q].x, px[q].y, px[q].z), 1e-8f);
    if(l < yo) l = yo * m_exp_s(1.0f * (l - yo), lme_ctx);
    const double ryo = rcp_dn(yo);   // yo >= 1e-8: normal
    f3 c = { fmaxf(0.0f, div_rd(px[q].x * l, ryo)), fmaxf(0.0f, div_rd(px[q].y * l, ryo)), fmaxf(0.0f, div_rd(px[q].z * l, ryo)) };
#endif
    if(GRADE)
    {
      c = { f16r(c.x), f16r(c.y), f16r(c.z) };
      c = GRADE == 1 ? grade_px_digest<true>(c, P) : grade_px_digest<false>(c, P);
    }
    if(F32 && P.out_f32 == 2) { oc[3 * (q & 1)] = c.x; oc[3 * (q & 1) + 1] = c.y; oc[3 * (q & 1) + 2] = c.z; }
    else if(F32) st_sink_f32(outv, ow, x, y, c.x, c.y, c.z, P.out_f32);
    else st_rgba(reinterpret_cast<uint2 *>(outv), ow, x, y, make_float4(c.x, c.y, c.z, 1.0f));
    }
    if(F32 && (q & 1) && P.out_f32 == 2)
    { // end of a row: the lanes left of the image border are the ones still here (hi[0] >= 0), contiguous from lane 0
      const int x0 = blockIdx.x * 64, left = ow - x0, y = 2 * ky + (q >> 1);
      if(y < oh) st_rgb_coop<6>(stage[threadIdx.y], reinterpret_cast<float *>(outv) + ((size_t)y * ow + x0) * 3, threadIdx.x,
          min(32, (left + 1) >> 1), oc, 3 * min(64, left));
    }
  }
}

// conn: [0] input rgba f16, [1] coarse y f16 (level 1 assembled; ignored when first), [2] level-1 stack x11, [3] output rgba f16|f32
// push: { u32 first; u32 have_grade }.  params: llap params (16 B) followed by grade params (76 B) when have_grade
static int launch_llapfin2(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 4 && l->push_size >= 8 && l->params_size >= sizeof(llap_params_t));
  const uint32_t *pc = (const uint32_t *)l->push;
  const vkb_image_t *in = l->conn, *coarse = l->conn + 1, *l1 = l->conn + 2, *out = l->conn + 3;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16 && l1->layers == NL);
  VKB_REQUIRE(in->wd == out->wd && in->ht == out->ht);
  VKB_REQUIRE(l1->wd == (in->wd - 1) / 2 + 1 && l1->ht == (in->ht - 1) / 2 + 1);
  VKB_REQUIRE(out->format == VKB_TOKEN_F16 || out->format == VKB_TOKEN_F32);
  VKB_REQUIRE(out->chan == 4 || (out->chan == 3 && out->format == VKB_TOKEN_F32)); // 3: packed rgb f32 sink (VKB_SINK_RGB_F32)
  llapfin_t P;
  memset(&P, 0, sizeof(P));
  memcpy(&P.p, l->params, sizeof(llap_params_t));
  P.first = pc[0]; P.have_grade = pc[1]; P.out_f32 = out->format == VKB_TOKEN_F32 ? (out->chan == 3 ? 2 : 1) : 0;
  P.inv2s = 1.0f / (2.0f * P.p.sigma); P.invd = 1.0f / (2.0f * P.p.sigma * P.p.sigma / 3.0f);
  { // the divisors as the kernel's fp32 expressions form them (volatile: one rounding per operation on the host too)
    const volatile float two_s = 2.0f * P.p.sigma;
    const volatile float ss = two_s * P.p.sigma;
    const volatile float k = ss / 3.0f;
    P.rd.rd2s = 1.0 / (double)two_s; P.rd.rdk = 1.0 / (double)k;
    P.rdg[0] = 0.0;
    for(int i = 1; i < NUM_GAMMA; i++)
    { // gamma_from_i() in fp32: i / 9
      const volatile float g1 = (float)i / (NUM_GAMMA - 1.0f), g0 = (float)(i - 1) / (NUM_GAMMA - 1.0f);
      const volatile float d = g1 - g0;
      P.rdg[i] = 1.0 / (double)d;
    }
  }
  if(P.have_grade)
  {
    VKB_REQUIRE(l->params_size >= sizeof(llap_params_t) + sizeof(grade_params_t));
    memcpy(&P.grade, (const uint8_t *)l->params + sizeof(llap_params_t), sizeof(grade_params_t));
    const grade_params_t &q = P.grade;
    for(int k = 0; k < 3; k++)
    { // grade_px(): lift, gam, gain, off and the two expressions that only depend on them
      const volatile float lift = q.lift[k] + q.lift[3], gam = fmaxf(q.gamma[k] + q.gamma[3], 1e-6f);
      P.g_lift[k] = lift; P.g_oml[k] = 1.0f - lift;
      P.g_gain[k] = fmaxf(q.gain[k] + q.gain[3], 0.0f);
      P.g_off[k]  = q.off[k] + q.off[3];
      P.g_ig[k]   = 1.0f / gam;
    }
  }
  dim3 grid(vkb_cdiv(out->wd, 64), vkb_cdiv(out->ht, 16)), block(32, 8);
  const band_t bd = band_of(l, 2, 8, (out->ht + 1) / 2, &grid.y); // band image: the output, two rows per thread row
  if(!grid.y) return VKB_OK;
#define GO(F, G) k_llap_final4<F, G><<<grid, block, 0, l->stream>>>((const uint2 *)in->data, (const __half *)coarse->data, \
      (const __half *)l1->data, l1->wd, l1->ht, out->data, out->wd, out->ht, P, bd)
  const bool plain = P.have_grade && P.grade.mode == 0 && P.g_ig[0] == 1.0f && P.g_ig[1] == 1.0f && P.g_ig[2] == 1.0f;
  if(P.out_f32) { if(plain) GO(true, 1); else if(P.have_grade) GO(true, 2); else GO(true, 0); }
  else          { if(plain) GO(false, 1); else if(P.have_grade) GO(false, 2); else GO(false, 0); }
#undef GO
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("b200", "llapfin", launch_llapfin2);

VKB_NS_END

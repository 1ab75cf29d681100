// highlight reconstruction on the mosaic via a clip-aware gaussian pyramid.
// replaces src/pipe/modules/hilite/{half,reduce,assemble,doub}.comp (wired by hilite/main.c:5-90).
// all taps of the reference sit on exact texel centres, so the sampler reduces to mirrored texel fetches.
#include "common.cuh"

struct hilite_params_t { float white, desat, soft; };          // hilite/params
struct hilite_push_t   { float wb[4]; uint32_t filters; };     // hilite/main.c:26

// ---- half: mosaic block -> rgb, clipped greens replaced (half.comp:24-72) ----
__global__ void __launch_bounds__(256) k_hilite_half(const __half *__restrict__ in, int iw, int ih,
    uint2 *__restrict__ out, int ow, int oh, float white, int xtrans, const band_t bd)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  float4 rgba;
  if(xtrans)
  {
    float c[9];
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
      for(int j = 0; j < 3; j++) c[3 * i + j] = ld_h_clamp(in, iw, ih, 3 * x + i, 3 * y + j);
    if(c[1] >= white) c[1] = c[7];
    if(c[7] >= white) c[7] = c[1];
    if(c[3] >= white) c[3] = c[5];
    if(c[5] >= white) c[5] = c[3];
    const float maxg = fmaxf(fmaxf(fmaxf(c[0], c[2]), c[4]), fmaxf(c[6], c[8]));
    if(maxg >= white) c[0] = c[2] = c[4] = c[6] = c[8] = 1.0f;
    const float col0 = (c[1] + c[7]) * 0.5f, col1 = (c[3] + c[5]) * .5f;
    if(((x + y) & 1) > 0) { rgba.x = col0; rgba.z = col1; }
    else                  { rgba.z = col0; rgba.x = col1; }
    rgba.y = div_c(c[0] + c[2] + c[4] + c[6] + c[8], 5.0f);
    rgba.w = 1.0f;
  }
  else
  { // textureGather at the block centre: x=(0,1) y=(1,1) z=(1,0) w=(0,0)
    const int x0 = mirror1(2 * x, iw), x1 = mirror1(2 * x + 1, iw), y0 = mirror1(2 * y, ih), y1 = mirror1(2 * y + 1, ih);
    float cx = ld_h(in, iw, x0, y1), cy = ld_h(in, iw, x1, y1), cz = ld_h(in, iw, x1, y0), cw = ld_h(in, iw, x0, y0);
    if(cx >= white) cx = cz;
    if(cz >= white) cz = cx;
    rgba = make_float4(cw, (cx + cz) / 2.0f, cy, 1.0f);
  }
  st_rgba(out, ow, x, y, rgba);
}

// ---- reduce: 5x5 binomial over unclipped pixels, desaturating near white (reduce.comp:21-60) ----
#define HR_TW 67
#define HR_TH 19
#define HR_COL(c) ((((c) & 1) * 34) + ((c) >> 1))
__global__ void __launch_bounds__(256) k_hilite_reduce(const uint2 *__restrict__ in, int iw, int ih,
    uint2 *__restrict__ out, int ow, int oh, hilite_params_t p, float wbr, float wbg, float wbb, const band_t bd)
{
  // everything the shader evaluates per tap except the binomial weight depends on the input texel alone (luminance,
  // clip test, desaturated colour: seven divisions and two smoothsteps).  a CTA of 32x8 outputs evaluates it once per
  // texel of its 67x19 input window into shared memory; the 25 tap loop below only accumulates, in the shader's order.
  // columns are stored even ones first, odd ones behind (HR_COL): a warp's taps 2*lane+ii then hit consecutive words
  __shared__ float4 tile[HR_TH][HR_TW + 1]; // desaturated r g b, luminance
  __shared__ float  okay[HR_TH][HR_TW + 1]; // 1 if no channel is clipped
  float white = p.white;
  if(!(white > 0.0f)) white = 1.0f;
  const float ds = p.desat * p.desat;
  const int tx0 = blockIdx.x * 64 - 2, ty0 = BAND_BY * 16 - 2;
  // only the part of the window the CTA's valid outputs read: the last levels of the pyramid are a few pixels large
  const int need_w = min(HR_TW, 2 * (ow - (int)blockIdx.x * 32) + 3), need_h = min(HR_TH, 2 * (oh - BAND_BY * 8) + 3);
  for(int t = threadIdx.y * 32 + threadIdx.x; t < need_w * need_h; t += 256)
  {
    const int r = t / need_w, c = t - r * need_w;
    const float4 rgb = ld_rgba(in, iw, mirrori(tx0 + c, iw), mirrori(ty0 + r, ih));
    float4 m = make_float4(0.0f, 0.0f, 0.0f, lum2020(rgb.x, rgb.y, rgb.z));
    const bool ok = rgb.x < white && rgb.y < white && rgb.z < white;
    if(ok)
    {
      const float rw = rgb.x * wbr, gw = rgb.y * wbg, bw = rgb.z * wbb;
      const float cmax = fmaxf(rw, fmaxf(gw, bw));
      const float cmin = fminf(rw, fminf(gw, bw));
      const float sat = div_g(cmax - cmin, fmaxf(1e-3f, cmax));
      const float s = smoothstepf(0.2f, 1.0f, div_g(fmaxf(rgb.x, fmaxf(rgb.y, rgb.z)), white));
      float tt = smoothstepf(0.15f, 0.9f, sat);
      tt = clampf(5.0f * ds * ds * s * tt, 0.0f, 1.0f);
      m.x = mixf(rgb.x, div_g(cmax + cmin, wbr) * .5f, tt);
      m.y = mixf(rgb.y, div_g(cmax + cmin, wbg) * .5f, tt);
      m.z = mixf(rgb.z, div_g(cmax + cmin, wbb) * .5f, tt);
    }
    tile[r][HR_COL(c)] = m;
    okay[r][HR_COL(c)] = ok ? 1.0f : 0.0f;
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  const float sw[5] = {1.0f, 2.0f, 0.0f, -2.0f, -1.0f};
  float ex = 0.0f, ey = 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, wgt = 0.0f;
  const int lx = 2 * threadIdx.x, ly = 2 * threadIdx.y; // tile coords of tap (-2,-2)
#pragma unroll
  for(int jj = 0; jj < 5; jj++)
#pragma unroll
    for(int ii = 0; ii < 5; ii++)
    {
      const float4 m = tile[ly + jj][HR_COL(lx + ii)];
      ex += w[jj] * sw[ii] * m.w;
      ey += w[ii] * sw[jj] * m.w;
      const float u = w[ii] * w[jj];
      if(okay[ly + jj][HR_COL(lx + ii)] != 0.0f)
      {
        cr += m.x * u; cg += m.y * u; cb += m.z * u;
        wgt += u;
      }
    }
  float4 o;
  if(wgt == 0.0f) { o.x = o.y = o.z = 1.0f; }
  else { o.x = div_g(cr, wgt); o.y = div_g(cg, wgt); o.z = div_g(cb, wgt); }
  o.w = sqrt_g(ex * ex + ey * ey);
  st_rgba(out, ow, x, y, o);
}

// ---- assemble: expand coarse, rescale to the fine level's unclipped channels, blend (assemble.comp:23-118) ----
__global__ void __launch_bounds__(256) k_hilite_assemble(const uint2 *__restrict__ fine_img, const uint2 *__restrict__ coarse,
    int cw, int ch, uint2 *__restrict__ out, int ow, int oh, hilite_params_t p, const band_t bd)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  const int ix = x / 2, iy = y / 2, dx = x & 1, dy = y & 1;
  float ur = 0.0f, ug = 0.0f, ub = 0.0f, wgt = 0.0f;
  // same taps in the same order (ii outer, jj inner), unrolled over the full -1..1 range with the odd pixels' missing
  // first tap predicated off: the weights become selects between constants instead of indexed loads from a local array
#pragma unroll
  for(int ii = -1; ii <= 1; ii++)
  {
    if(ii < 0 && dx) continue;
    const float wx = ii < 0 ? w[0] : (ii == 0 ? (dx ? w[1] : w[2]) : (dx ? w[3] : w[4]));
    const int cx = mirror1(ix + ii, cw);
#pragma unroll
    for(int jj = -1; jj <= 1; jj++)
    {
      if(jj < 0 && dy) continue;
      const float wy = jj < 0 ? w[0] : (jj == 0 ? (dy ? w[1] : w[2]) : (dy ? w[3] : w[4]));
      const float4 rgb = ld_rgba(coarse, cw, cx, mirror1(iy + jj, ch));
      ur += rgb.x * wy * wx; ug += rgb.y * wy * wx; ub += rgb.z * wy * wx;
      wgt += wy * wx;
    }
  }
  if(wgt == 0.0f) { ur = 0.0f; ug = 1.0f; ub = 1.0f; }
  else { ur = div_f(ur, wgt); ug = div_f(ug, wgt); ub = div_f(ub, wgt); }   // a sum of the positive tap weights
  float4 fine = ld_rgba(fine_img, ow, x, y);
  const float white = p.white;
  const float sr = div_f(fine.x, fmaxf(0.001f, ur));
  const float sg = div_f(fine.y, fmaxf(0.001f, ug));
  const float sb = div_f(fine.z, fmaxf(0.001f, ub));
  // blend weights: libm's exponential bit for bit (strict) or the SFU one (fast, 2 ulp), three times per pixel
  const float wr = m_exp(ur - fmaxf(ug, ub));
  const float wg = m_exp(ug - fmaxf(ur, ub));
  const float wb = m_exp(ub - fmaxf(ur, ug));
  const float scale = div_g(sr * wr + sg * wg + sb * wb, wr + wg + wb);
  float t = p.soft;
  if(fine.x >= white || fine.y >= white || fine.z >= white) t = 1.0f;
  if(isnan(fine.w)) fine.w = 0.0f;
  t = clampf(mixf(t, 1.0f, sqrt_g(fmaxf(0.0f, fine.w))), 0.0f, 1.0f);
  float4 o;
  o.x = mixf(fine.x, clampf(ur * scale, -65535.0f, 65535.0f), t);
  o.y = mixf(fine.y, clampf(ug * scale, -65535.0f, 65535.0f), t);
  o.z = mixf(fine.z, clampf(ub * scale, -65535.0f, 65535.0f), t);
  o.w = 1.0f;
  st_rgba(out, ow, x, y, o);
}

// ---- doub: write reconstructed values back into the mosaic where it clips (doub.comp:22-121) ----
__global__ void __launch_bounds__(256) k_hilite_doub(const __half *__restrict__ in, int iw, int ih,
    const uint2 *__restrict__ coarse, int cw, int ch, __half *__restrict__ out, int ow, int oh, float white, int xtrans, const band_t bd)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  const int bw = xtrans ? ow / 3 : ow / 2, bh = xtrans ? oh / 3 : oh / 2;
  if(x >= bw || y >= bh || BAND_SKIP(y)) return;
  const float softw = 0.97f * white;
  // the reconstructed colour only enters a block whose brightest site is above the soft threshold (doub.comp:88, :113): every
  // other block is copied, and neither the coarse texel nor the three exponentials and four quotients behind `scale` are needed
  if(xtrans)
  {
    float c[9];
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
      for(int j = 0; j < 3; j++) c[3 * i + j] = ld_h_clamp(in, iw, ih, 3 * x + i, 3 * y + j);
    const float maxr = fmaxf(c[1], c[7]), maxb = fmaxf(c[3], c[5]);
    const float maxg = fmaxf(fmaxf(fmaxf(c[0], c[2]), c[4]), fmaxf(c[6], c[8]));
    const float maxrgb = fmaxf(maxr, fmaxf(maxg, maxb));
    if(maxrgb > softw)
    {
    float4 upsm = ld_rgba_clamp(coarse, cw, ch, x, y);
    if(((x + y) & 1) == 0) { const float t = upsm.x; upsm.x = upsm.z; upsm.z = t; }
    const float minr = fminf(c[1], c[7]), minb = fminf(c[3], c[5]);
    const float ming = fminf(fminf(fminf(c[0], c[2]), c[4]), fminf(c[6], c[8]));
    const float sr = div_g(minr, fmaxf(0.001f, upsm.x));
    const float sg = div_g(ming, fmaxf(0.001f, upsm.y));
    const float sb = div_g(minb, fmaxf(0.001f, upsm.z));
    const float wr = m_exp(upsm.x - fmaxf(upsm.y, upsm.z));
    const float wg = m_exp(upsm.y - fmaxf(upsm.x, upsm.z));
    const float wb = m_exp(upsm.z - fmaxf(upsm.x, upsm.y));
    const float scale = div_g(sr * wr + sg * wg + sb * wb, wr + wg + wb);
    {
      float t = smoothstepf(softw, white, maxrgb);
      c[1] = mixf(c[1], upsm.x * scale, t); c[7] = mixf(c[7], upsm.x * scale, t);
      c[3] = mixf(c[3], upsm.z * scale, t); c[5] = mixf(c[5], upsm.z * scale, t);
      t = smoothstepf(softw, white, ming);
      c[0] = mixf(c[0], upsm.y * scale, t); c[2] = mixf(c[2], upsm.y * scale, t);
      c[4] = mixf(c[4], upsm.y * scale, t); c[6] = mixf(c[6], upsm.y * scale, t);
      c[8] = mixf(c[8], upsm.y * scale, t);
    }
    }
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
      for(int j = 0; j < 3; j++)
        if(3 * x + i < ow && 3 * y + j < oh) out[(size_t)(3 * y + j) * ow + 3 * x + i] = __float2half_rn(c[3 * i + j]);
  }
  else
  {
    const int x0 = mirror1(2 * x, iw), x1 = mirror1(2 * x + 1, iw), y0 = mirror1(2 * y, ih), y1 = mirror1(2 * y + 1, ih);
    float c[4] = { ld_h(in, iw, x0, y1), ld_h(in, iw, x1, y1), ld_h(in, iw, x1, y0), ld_h(in, iw, x0, y0) };
    const float maxrgb = fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3]));
    if(maxrgb > softw)
    {
    const float4 upsm = ld_rgba_clamp(coarse, cw, ch, x, y);
    const float ming = fminf(c[0], c[2]);
    const float sr = div_g(c[3], fmaxf(0.001f, upsm.x));
    const float sg = div_g(ming, fmaxf(0.001f, upsm.y));
    const float sb = div_g(c[1], fmaxf(0.001f, upsm.z));
    const float wr = m_exp(upsm.x - fmaxf(upsm.y, upsm.z));
    const float wg = m_exp(upsm.y - fmaxf(upsm.x, upsm.z));
    const float wb = m_exp(upsm.z - fmaxf(upsm.x, upsm.y));
    const float scale = div_g(sr * wr + sg * wg + sb * wb, wr + wg + wb);
    const float t = smoothstepf(softw, white, maxrgb);
    c[0] = mixf(c[0], scale * upsm.y, t); c[1] = mixf(c[1], scale * upsm.z, t);
    c[2] = mixf(c[2], scale * upsm.y, t); c[3] = mixf(c[3], scale * upsm.x, t);
    }
    // two texels per row -> one 4 byte store each
    const __half2 top = __floats2half2_rn(c[3], c[2]), bot = __floats2half2_rn(c[0], c[1]);
    if((ow & 1) == 0)
    {
      *reinterpret_cast<__half2 *>(out + (size_t)(2 * y) * ow + 2 * x) = top;
      *reinterpret_cast<__half2 *>(out + (size_t)(2 * y + 1) * ow + 2 * x) = bot;
    }
    else
    {
      out[(size_t)(2 * y) * ow + 2 * x] = __low2half(top);     out[(size_t)(2 * y) * ow + 2 * x + 1] = __high2half(top);
      out[(size_t)(2 * y + 1) * ow + 2 * x] = __low2half(bot); out[(size_t)(2 * y + 1) * ow + 2 * x + 1] = __high2half(bot);
    }
  }
}

static inline dim3 grid2d(unsigned w, unsigned h) { return dim3(vkb_cdiv(w, 32), vkb_cdiv(h, 8)); }
static const dim3 blk2d(32, 8);

// conn: [0] input mosaic f16 1ch, [1] output rgba f16
static int launch_hilite_half(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= sizeof(hilite_push_t) && l->params_size >= 4);
  const hilite_push_t *pc = (const hilite_push_t *)l->push;
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->format == VKB_TOKEN_F16 && in->chan == 1 && out->format == VKB_TOKEN_F16 && out->chan == 4);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_hilite_half<<<grid, blk2d, 0, l->stream>>>((const __half *)in->data, in->wd, in->ht,
      (uint2 *)out->data, out->wd, out->ht, ((const float *)l->params)[0], pc->filters == 9, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("hilite", "half", launch_hilite_half);

// conn: [0] input rgba f16, [1] output rgba f16
static int launch_hilite_reduce(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= sizeof(hilite_push_t) && l->params_size >= sizeof(hilite_params_t));
  const hilite_push_t *pc = (const hilite_push_t *)l->push;
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && out->chan == 4 && in->format == VKB_TOKEN_F16 && out->format == VKB_TOKEN_F16);
  float wb[3] = { pc->wb[0], pc->wb[1], pc->wb[2] };
  if(!(wb[0] * wb[0] + wb[1] * wb[1] + wb[2] * wb[2] > 1e-3f)) wb[0] = wb[1] = wb[2] = 1.0f;
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_hilite_reduce<<<grid, blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht,
      (uint2 *)out->data, out->wd, out->ht, *(const hilite_params_t *)l->params, wb[0], wb[1], wb[2], bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("hilite", "reduce", launch_hilite_reduce);

// conn: [0] fine rgba f16, [1] coarse rgba f16, [2] output rgba f16 (fine dims)
static int launch_hilite_assemble(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 3 && l->params_size >= sizeof(hilite_params_t));
  const vkb_image_t *fine = l->conn, *coarse = l->conn + 1, *out = l->conn + 2;
  VKB_REQUIRE(fine->chan == 4 && coarse->chan == 4 && out->chan == 4 && fine->wd == out->wd && fine->ht == out->ht);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_hilite_assemble<<<grid, blk2d, 0, l->stream>>>((const uint2 *)fine->data, (const uint2 *)coarse->data,
      coarse->wd, coarse->ht, (uint2 *)out->data, out->wd, out->ht, *(const hilite_params_t *)l->params, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("hilite", "assemble", launch_hilite_assemble);

// conn: [0] input mosaic f16, [1] coarse rgba f16, [2] output mosaic f16
static int launch_hilite_doub(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 3 && l->push_size >= sizeof(hilite_push_t) && l->params_size >= 4);
  const hilite_push_t *pc = (const hilite_push_t *)l->push;
  const vkb_image_t *in = l->conn, *coarse = l->conn + 1, *out = l->conn + 2;
  VKB_REQUIRE(in->chan == 1 && coarse->chan == 4 && out->chan == 1 && out->format == VKB_TOKEN_F16);
  const int xt = pc->filters == 9;
  dim3 grid = grid2d(out->wd / (xt ? 3 : 2), out->ht / (xt ? 3 : 2));
  const band_t bd = band_of(l, xt ? 3 : 2, 8, out->ht / (xt ? 3 : 2), &grid.y); // band image: the output mosaic, one cfa block row per thread row
  if(!grid.y) return VKB_OK;
  k_hilite_doub<<<grid, blk2d, 0, l->stream>>>((const __half *)in->data, in->wd, in->ht,
      (const uint2 *)coarse->data, coarse->wd, coarse->ht, (__half *)out->data, out->wd, out->ht, ((const float *)l->params)[0], xt, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("hilite", "doub", launch_hilite_doub);

VKB_NS_END

// bayer fast path of demosaic/splat.comp and demosaic/fix.comp: one thread per SHIFTED 2x2 block.
// both shaders fetch their gaussian at cov[(x+1)/2, (y+1)/2] (splat.comp:129, fix.comp:122), so the four pixels
// (2b-1..2b, 2c-1..2c) share one covariance and therefore the same tap weights: 4 (splat) / 4 (fix) exponentials per
// block instead of ~9 per pixel, no divergence between green and red/blue sites, mosaic and green fetched as 4x4
// neighbourhoods with 4-byte loads.  arithmetic and accumulation order per pixel are exactly those of the shaders
// (loop j outer, i inner), so results are bit identical to the per-pixel kernels in k_demosaic.cu.
#include "common.cuh"

// splat.comp:27-30: the border "hack" instead of mirroring
VKB_DEV float splat_fetch(const __half *__restrict__ in, int w, int h, int px, int py)
{
  if(px < 0) px += 6;
  if(py < 0) py += 6;
  if(px >= w) px -= 6;
  if(py >= h) py -= 6;
  return ld_h_clamp(in, w, h, px, py);
}
#if VKB_FAST
typedef float ediv_t;                      // the eigenvalue itself
VKB_DEV ediv_t ediv(float e) { return e; }
VKB_DEV float edivide(float x, ediv_t e) { return div_f(x, e); }   // clamped eigenvalue: the call free exact quotient
#else
typedef double ediv_t;                     // its reciprocal in double: a block's eight / ten quotients share two divisors (div_rd)
VKB_DEV ediv_t ediv(float e) { return rcp_dn(e); }   // clamped to [0.01, 98]: normal
VKB_DEV float edivide(float x, ediv_t e) { return div_rd(x, e); }
#endif
VKB_DEV float splat_weight(ediv_t e0, ediv_t e1, float cz, float cw, int i, int j)
{ // splat.comp:33-37
  const float of0 = cz * (float)i + cw * (float)j;
  const float of1 = -cw * (float)i + cz * (float)j;
  return clampf(m_exp(-0.5f * (edivide(of0, e0) * of0 + edivide(of1, e1) * of1)), 1e-4f, 1.0f);
}

__global__ void __launch_bounds__(256, 6) k_bayer_splat(const __half *__restrict__ in, int w, int h,
    const uint2 *__restrict__ gauss, int gw, int gh, __half *__restrict__ out, const band_t bd)
{
  const int b = blockIdx.x * 32 + threadIdx.x, c = BAND_BY * 8 + threadIdx.y;
  if(2 * b - 1 >= w || 2 * c - 1 >= h) return;
  const float4 cov = ld_rgba_clamp(gauss, gw, gh, b, c);
  const ediv_t e0 = ediv(clampf(cov.x, 0.01f, 25.0f)), e1 = ediv(clampf(cov.y, 0.01f, 25.0f));
  const float w11 = splat_weight(e0, e1, cov.z, cov.w, 1, 1), w1m = splat_weight(e0, e1, cov.z, cov.w, 1, -1);
  const float w10 = splat_weight(e0, e1, cov.z, cov.w, 1, 0), w01 = splat_weight(e0, e1, cov.z, cov.w, 0, 1);
  // 4x4 neighbourhood m[j][i] = mosaic(2b-2+i, 2c-2+j)
  float m[4][4];
  const int x0 = 2 * b - 2, y0 = 2 * c - 2;
  const bool interior = x0 >= 0 && y0 >= 0 && x0 + 3 < w && y0 + 3 < h && (w & 1) == 0;
  if(interior)
  {
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      const __half2 *row = reinterpret_cast<const __half2 *>(in + (size_t)(y0 + j) * w + x0);
      const float2 a = __half22float2(__ldg(row)), bb = __half22float2(__ldg(row + 1));
      m[j][0] = a.x; m[j][1] = a.y; m[j][2] = bb.x; m[j][3] = bb.y;
    }
  }
#pragma unroll
  for(int q = 0; q < 4; q++)
  { // the four pixels of the shifted block: (lx,ly) in {1,2}^2 of the 4x4 neighbourhood
    const int lx = 1 + (q & 1), ly = 1 + (q >> 1);
    const int x = x0 + lx, y = y0 + ly;
    if(x < 0 || y < 0 || x >= w || y >= h || BAND_SKIP(y)) continue; // band rows are output rows here
    float g = 0.0f, wg = 0.0f;
    const bool green = ((x & 1) != (y & 1));
#define M(I, J) (interior ? m[ly + (J)][lx + (I)] : splat_fetch(in, w, h, x + (I), y + (J)))
    if(green)
    { // taps in shader order: (-1,-1) (1,-1) (0,0) (-1,1) (1,1); w(-i,-j) == w(i,j)
      float col;
      col = M(-1, -1) * w11; g += col; wg += w11;
      col = M( 1, -1) * w1m; g += col; wg += w1m;
      col = M( 0,  0) * 666.0f; g += col; wg += 666.0f;
      col = M(-1,  1) * w1m; g += col; wg += w1m;
      col = M( 1,  1) * w11; g += col; wg += w11;
    }
    else
    { // (0,-1) (-1,0) (1,0) (0,1)
      float col;
      col = M( 0, -1) * w01; g += col; wg += w01;
      col = M(-1,  0) * w10; g += col; wg += w10;
      col = M( 1,  0) * w10; g += col; wg += w10;
      col = M( 0,  1) * w01; g += col; wg += w01;
    }
#undef M
    out[(size_t)y * w + x] = __float2half_rn(div_f(g, fmaxf(1e-8f, wg)));
  }
}

VKB_DEV float fixw(ediv_t e0, ediv_t e1, float cz, float cw, int i, int j)
{ // fix.comp:16-23
  const float of0 = cz * (float)i + cw * (float)j;
  const float of1 = -cw * (float)i + cz * (float)j;
  return clampf(m_exp(-0.5f * (edivide(of0, e0) * of0 + edivide(of1, e1) * of1)), 1e-3f, 1.0f);
}

__global__ void __launch_bounds__(256, 5) k_bayer_fix(const __half *__restrict__ in, const __half *__restrict__ green, int w, int h,
    const uint2 *__restrict__ covimg, int gw, int gh, uint2 *__restrict__ out, const band_t bd)
{
  const int b = blockIdx.x * 32 + threadIdx.x, c = BAND_BY * 8 + threadIdx.y;
  if(2 * b - 1 >= w || 2 * c - 1 >= h) return;
  float4 cov = ld_rgba_clamp(covimg, gw, gh, b, c);
  cov.x = clampf(cov.x, 1.0f, 49.f); cov.y = clampf(cov.y, 1.0f, 49.f);
  const ediv_t e0 = ediv(2.0f * cov.x), e1 = ediv(2.0f * cov.y);
  const float w00 = fixw(e0, e1, cov.z, cov.w, 0, 0);
  const float w11 = fixw(e0, e1, cov.z, cov.w, 1, 1), w1m = fixw(e0, e1, cov.z, cov.w, 1, -1);
  const float w10 = fixw(e0, e1, cov.z, cov.w, 1, 0), w01 = fixw(e0, e1, cov.z, cov.w, 0, 1);
  float m[4][4], g[4][4];
  const int x0 = 2 * b - 2, y0 = 2 * c - 2;
  const bool interior = x0 >= 0 && y0 >= 0 && x0 + 3 < w && y0 + 3 < h && (w & 1) == 0;
  if(interior)
  {
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      const __half2 *rm = reinterpret_cast<const __half2 *>(in + (size_t)(y0 + j) * w + x0);
      const __half2 *rg = reinterpret_cast<const __half2 *>(green + (size_t)(y0 + j) * w + x0);
      const float2 a = __half22float2(__ldg(rm)), bb = __half22float2(__ldg(rm + 1));
      const float2 ga = __half22float2(__ldg(rg)), gb = __half22float2(__ldg(rg + 1));
      m[j][0] = a.x; m[j][1] = a.y; m[j][2] = bb.x; m[j][3] = bb.y;
      g[j][0] = ga.x; g[j][1] = ga.y; g[j][2] = gb.x; g[j][3] = gb.y;
    }
  }
  else
  { // texture(): mirrored repeat
#pragma unroll
    for(int j = 0; j < 4; j++)
#pragma unroll
      for(int i = 0; i < 4; i++)
      {
        const int xx = mirrori(x0 + i, w), yy = mirrori(y0 + j, h);
        m[j][i] = ld_h(in, w, xx, yy); g[j][i] = ld_h(green, w, xx, yy);
      }
  }
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    const int lx = 1 + (q & 1), ly = 1 + (q >> 1);
    const int x = x0 + lx, y = y0 + ly;
    if(x < 0 || y < 0 || x >= w || y >= h || BAND_SKIP(y)) continue; // band rows are output rows here
    const float gc = g[ly][lx];
    float r = 0.0f, bl = 0.0f, wr = 0.0f, wb = 0.0f;
// (1e-4f + an f16 value is never zero: -1e-4f is no f16 value.  div_f: as double multiplies these 18 quotients per block
// made the kernel XU bound, 0.41 -> 0.52 ms)
#define TAP(ACC, WACC, I, J, WGT) { ACC += div_f(m[ly + (J)][lx + (I)] * (1e-4f + gc), 1e-4f + g[ly + (J)][lx + (I)]) * (WGT); WACC += (WGT); }
    const int ex = (x & 1) == 0, ey = (y & 1) == 0;
    if(ex && ey)
    { // red site: blue on the diagonals, red at the centre; shader order j outer, i inner
      TAP(bl, wb, -1, -1, w11) TAP(bl, wb, 1, -1, w1m) TAP(r, wr, 0, 0, w00) TAP(bl, wb, -1, 1, w1m) TAP(bl, wb, 1, 1, w11)
    }
    else if(!ex && !ey)
    { // blue site
      TAP(r, wr, -1, -1, w11) TAP(r, wr, 1, -1, w1m) TAP(bl, wb, 0, 0, w00) TAP(r, wr, -1, 1, w1m) TAP(r, wr, 1, 1, w11)
    }
    else if(!ex && ey)
    { // green in a red row: red left/right, blue above/below
      TAP(bl, wb, 0, -1, w01) TAP(r, wr, -1, 0, w10) TAP(r, wr, 1, 0, w10) TAP(bl, wb, 0, 1, w01)
    }
    else
    { // green in a blue row: blue left/right, red above/below
      TAP(r, wr, 0, -1, w01) TAP(bl, wb, -1, 0, w10) TAP(bl, wb, 1, 0, w10) TAP(r, wr, 0, 1, w01)
    }
#undef TAP
    st_rgba(out, w, x, y, make_float4(div_f(r, fmaxf(1e-8f, wr)), gc / fmaxf(1e-8f, 1.0f), div_f(bl, fmaxf(1e-8f, wb)), 1.0f));
  }
}

// band of the shifted block kernels: thread row c writes output rows 2c-1 and 2c, so rows [y0, y1) need c in
// [y0 / 2, y1 / 2]; bd.y0 / bd.y1 stay OUTPUT rows and are tested per pixel
static inline band_t band_shifted(const vkb_launch_t *l, int out_ht, unsigned *grid_y)
{
  band_t b = { 0, 0, out_ht };
  if(l->band_y0 < 0) return b;
  b.y0 = l->band_y0; b.y1 = l->band_y1 < out_ht ? l->band_y1 : out_ht;
  const int c0 = b.y0 / 2, c1 = b.y1 / 2 + 1; // thread rows [c0, c1)
  b.by0 = c0 / 8;
  *grid_y = b.y1 > b.y0 ? (unsigned)((c1 + 7) / 8 - b.by0) : 0u;
  return b;
}
int launch_bayer_splat(const vkb_launch_t *l)
{
  const vkb_image_t *in = l->conn, *g = l->conn + 1, *out = l->conn + 2;
  dim3 grid(vkb_cdiv(out->wd / 2 + 1, 32), vkb_cdiv(out->ht / 2 + 1, 8));
  const band_t bd = band_shifted(l, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_bayer_splat<<<grid, dim3(32, 8), 0, l->stream>>>((const __half *)in->data, in->wd, in->ht, (const uint2 *)g->data, g->wd, g->ht, (__half *)out->data, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
int launch_bayer_fix(const vkb_launch_t *l)
{
  const vkb_image_t *in = l->conn, *g = l->conn + 1, *cov = l->conn + 2, *out = l->conn + 3;
  dim3 grid(vkb_cdiv(out->wd / 2 + 1, 32), vkb_cdiv(out->ht / 2 + 1, 8));
  const band_t bd = band_shifted(l, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  k_bayer_fix<<<grid, dim3(32, 8), 0, l->stream>>>((const __half *)in->data, (const __half *)g->data, in->wd, in->ht,
      (const uint2 *)cov->data, cov->wd, cov->ht, (uint2 *)out->data, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

VKB_NS_END

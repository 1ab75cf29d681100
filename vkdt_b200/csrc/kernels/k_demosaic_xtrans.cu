// x-trans fast path of demosaic/splat.comp and demosaic/fix.comp (radius 2): one thread per 3x3 block.
// both shaders pick their gaussian per 3x3 block (splat.comp: cov[x/3, y/3], fix.comp: cov[(x+1)/3, (y+1)/3]), so the
// nine pixels of a block share the 25 tap weights: 13 exponentials per block (w(-i,-j) == w(i,j) bit for bit) instead
// of 25 per pixel, and the cfa colour of every tap is a compile time property of its position in the block's 7x7
// window (red and blue swap with the parity of the block, so they are accumulated by "base colour" and swapped at
// the end).  per pixel the arithmetic and the accumulation order (j outer, i inner) are those of the shaders.
// blocks whose window leaves the image take the per pixel path with the shaders' border rules.
// compiled with --fmad=false like the other demosaic kernels.
#include "common.cuh"

VKB_DEV float xt_weight(float e0, float e1, float cz, float cw, int i, int j, float lo)
{ // splat.comp:33-37 (lo = 1e-4), fix.comp:16-23 (lo = 1e-3)
  const float of0 = cz * (float)i + cw * (float)j;
  const float of1 = -cw * (float)i + cz * (float)j;
  return clampf(m_exp(-0.5f * (div_f(of0, e0) * of0 + div_f(of1, e1) * of1)), lo, 1.0f);   // the eigenvalues are clamped: div_f
}
// the 13 distinct weights of a 5x5 window: index (j+2)*5 + (i+2), mirrored index 24 - idx
#define XT_WEIGHTS(W, E0, E1, CZ, CW, LO) \
  float W[25]; \
  _Pragma("unroll") for(int k = 0; k < 13; k++) { W[k] = xt_weight(E0, E1, CZ, CW, k % 5 - 2, k / 5 - 2, LO); W[24 - k] = W[k]; }

// ---- splat: green by anisotropic gaussian weights (splat.comp:16-140) ----
VKB_DEV float xt_splat_px(const __half *__restrict__ in, int w, int h, const uint2 *__restrict__ gauss, int gw, int gh, int x, int y)
{ // per pixel path with the shader's border hack
  const float4 cov = ld_rgba_clamp(gauss, gw, gh, x / 3, y / 3);
  const float e0 = clampf(cov.x, 0.01f, 25.0f), e1 = clampf(cov.y, 0.01f, 25.0f);
  float g = 0.0f, wg = 0.0f;
  for(int j = -2; j <= 2; j++) for(int i = -2; i <= 2; i++)
  {
    if(xtrans_colour(x + i + 6, y + j + 6) != 1) continue;
    int px = x + i, py = y + j;
    if(px < 0) px += 6;
    if(py < 0) py += 6;
    if(px >= w) px -= 6;
    if(py >= h) py -= 6;
    const float col = ld_h_clamp(in, w, h, px, py);
    float weight = xt_weight(e0, e1, cov.z, cov.w, i, j, 1e-4f);
    if(i == 0 && j == 0) weight = 666.0f;
    g += col * weight; wg += weight;
  }
  return div_f(g, fmaxf(1e-8f, wg));
}

__global__ void __launch_bounds__(128, 8) k_xtrans_splat(const __half *__restrict__ in, int w, int h,
    const uint2 *__restrict__ gauss, int gw, int gh, __half *__restrict__ out)
{
  const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 4 + threadIdx.y;
  const int x0 = 3 * X, y0 = 3 * Y;
  if(x0 >= w || y0 >= h) return;
  if(x0 - 2 < 0 || y0 - 2 < 0 || x0 + 4 >= w || y0 + 4 >= h)
  {
    for(int dy = 0; dy < 3; dy++) for(int dx = 0; dx < 3; dx++)
      if(x0 + dx < w && y0 + dy < h) out[(size_t)(y0 + dy) * w + x0 + dx] = __float2half_rn(xt_splat_px(in, w, h, gauss, gw, gh, x0 + dx, y0 + dy));
    return;
  }
  const float4 cov = ld_rgba(gauss, gw, min(X, gw - 1), min(Y, gh - 1));
  const float e0 = clampf(cov.x, 0.01f, 25.0f), e1 = clampf(cov.y, 0.01f, 25.0f);
  XT_WEIGHTS(wt, e0, e1, cov.z, cov.w, 1e-4f)
  // window m[v][u] = mosaic(x0 - 2 + u, y0 - 2 + v); only its green sites are read.  green iff ((u+1)%3 + (v+1)%3) is even
  float m[7][7];
#pragma unroll
  for(int v = 0; v < 7; v++)
#pragma unroll
    for(int u = 0; u < 7; u++)
      if((((u + 1) % 3 + (v + 1) % 3) & 1) == 0) m[v][u] = ld_h(in, w, x0 - 2 + u, y0 - 2 + v);
#pragma unroll
  for(int dy = 0; dy < 3; dy++)
#pragma unroll
    for(int dx = 0; dx < 3; dx++)
    {
      float g = 0.0f, wg = 0.0f;
#pragma unroll
      for(int j = -2; j <= 2; j++)
#pragma unroll
        for(int i = -2; i <= 2; i++)
        {
          const int u = dx + i + 2, v = dy + j + 2;
          if((((u + 1) % 3 + (v + 1) % 3) & 1) != 0) continue;
          const float weight = (i == 0 && j == 0) ? 666.0f : wt[(j + 2) * 5 + i + 2];
          g += m[v][u] * weight; wg += weight;
        }
      out[(size_t)(y0 + dy) * w + x0 + dx] = __float2half_rn(div_f(g, fmaxf(1e-8f, wg)));
    }
}

// ---- fix: red and blue by green-ratio weighted taps (fix.comp:25-135), radius 2 ----
VKB_DEV float4 xt_fix_px(const __half *__restrict__ in, const __half *__restrict__ green, int w, int h,
    const uint2 *__restrict__ covimg, int gw, int gh, int x, int y)
{ // per pixel path: mirrored texture() fetches, colour without the +6 margin (fix.comp:38-40)
  const float gc = ld_h(green, w, x, y);
  float4 cov = ld_rgba_clamp(covimg, gw, gh, (x + 1) / 3, (y + 1) / 3);
  cov.x = clampf(cov.x, 1.f, 10.f); cov.y = clampf(cov.y, 1.f, 10.f);
  float rgb[3] = {0, 0, 0}, wt[3] = {0, 0, 0};
  for(int j = -2; j <= 2; j++) for(int i = -2; i <= 2; i++)
  {
    const int px = x + i, py = y + j;
    const int c = xtrans_colour(px, py);
    if(c == 1) { rgb[1] = gc; wt[1] = 1.0f; continue; }
    const float gh_ = ld_h_mirror(green, w, h, px, py);
    const float col = ld_h_mirror(in, w, h, px, py);
    const float weight = xt_weight(3.0f * cov.x, 3.0f * cov.y, cov.z, cov.w, i, j, 1e-3f);
    const float corr = div_f(1e-4f + gc, 1e-4f + gh_);   // 1e-4f + an f16 value is never zero
    rgb[c] += col * corr * weight;
    wt[c] += weight;
  }
  return make_float4(div_f(rgb[0], fmaxf(1e-8f, wt[0])), div_f(rgb[1], fmaxf(1e-8f, wt[1])), div_f(rgb[2], fmaxf(1e-8f, wt[2])), 1.0f);
}

__global__ void __launch_bounds__(128, 6) k_xtrans_fix(const __half *__restrict__ in, const __half *__restrict__ green, int w, int h,
    const uint2 *__restrict__ covimg, int gw, int gh, uint2 *__restrict__ out)
{
  const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 4 + threadIdx.y;
  const int x0 = 3 * X - 1, y0 = 3 * Y - 1; // first pixel of the block: (x+1)/3 == X for x0..x0+2
  if(x0 >= w || y0 >= h) return;
  if(x0 - 2 < 0 || y0 - 2 < 0 || x0 + 4 >= w || y0 + 4 >= h)
  {
    for(int dy = 0; dy < 3; dy++) for(int dx = 0; dx < 3; dx++)
    {
      const int x = x0 + dx, y = y0 + dy;
      if(x >= 0 && y >= 0 && x < w && y < h) st_rgba(out, w, x, y, xt_fix_px(in, green, w, h, covimg, gw, gh, x, y));
    }
    return;
  }
  float4 cov = ld_rgba(covimg, gw, min(X, gw - 1), min(Y, gh - 1));
  cov.x = clampf(cov.x, 1.f, 10.f); cov.y = clampf(cov.y, 1.f, 10.f);
  XT_WEIGHTS(wt, 3.0f * cov.x, 3.0f * cov.y, cov.z, cov.w, 1e-3f)
  // window [v][u] = (x0 - 2 + u, y0 - 2 + v) = (3(X-1) + u, 3(Y-1) + v): u % 3, u / 3 are the cfa coordinates inside /
  // of the 3x3 cell.  green iff (u%3 + v%3) is even; otherwise "base" blue iff ((u/3 + v/3) & 1) ^ (v%3 == 1), and the
  // true colour is the base colour for even X+Y, the other one for odd X+Y.
  float m[7][7], g[7][7];
#pragma unroll
  for(int v = 0; v < 7; v++)
#pragma unroll
    for(int u = 0; u < 7; u++)
    {
      const bool is_green = ((u % 3 + v % 3) & 1) == 0;
      const bool centre = u >= 2 && u <= 4 && v >= 2 && v <= 4;
      if(!is_green) m[v][u] = ld_h(in, w, x0 - 2 + u, y0 - 2 + v);
      if(!is_green || centre) g[v][u] = ld_h(green, w, x0 - 2 + u, y0 - 2 + v);
    }
  const bool swap = (X + Y) & 1;
#pragma unroll
  for(int dy = 0; dy < 3; dy++)
#pragma unroll
    for(int dx = 0; dx < 3; dx++)
    {
      const float gc = g[dy + 2][dx + 2];
      float a0 = 0.0f, a2 = 0.0f, w0 = 0.0f, w2 = 0.0f; // base red / base blue
#pragma unroll
      for(int j = -2; j <= 2; j++)
#pragma unroll
        for(int i = -2; i <= 2; i++)
        {
          const int u = dx + i + 2, v = dy + j + 2;
          if(((u % 3 + v % 3) & 1) == 0) continue;
          const bool base_blue = (((u / 3 + v / 3) & 1) != 0) != (v % 3 == 1);
          const float weight = wt[(j + 2) * 5 + i + 2];
          const float corr = div_f(1e-4f + gc, 1e-4f + g[v][u]);
          if(base_blue) { a2 += m[v][u] * corr * weight; w2 += weight; }
          else          { a0 += m[v][u] * corr * weight; w0 += weight; }
        }
      const float r0 = div_f(a0, fmaxf(1e-8f, w0)), r2 = div_f(a2, fmaxf(1e-8f, w2));
      // a 5x5 window always holds green sites: rgb[1] = gc, w[1] = 1
      st_rgba(out, w, x0 + dx, y0 + dy, make_float4(swap ? r2 : r0, gc / fmaxf(1e-8f, 1.0f), swap ? r0 : r2, 1.0f));
    }
}

static inline dim3 grid3(unsigned nx, unsigned ny) { return dim3(vkb_cdiv(nx, 32), vkb_cdiv(ny, 4)); }

// conn: [0] input mosaic f16 [1] gauss rgba f16 [2] output green f16
int launch_xtrans_splat(const vkb_launch_t *l)
{
  const vkb_image_t *in = l->conn, *g = l->conn + 1, *out = l->conn + 2;
  VKB_REQUIRE(in->format == VKB_TOKEN_F16 && out->format == VKB_TOKEN_F16 && g->format == VKB_TOKEN_F16);
  k_xtrans_splat<<<grid3(vkb_cdiv(out->wd, 3), vkb_cdiv(out->ht, 3)), dim3(32, 4), 0, l->stream>>>((const __half *)in->data, in->wd, in->ht,
      (const uint2 *)g->data, g->wd, g->ht, (__half *)out->data);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

// conn: [0] input mosaic f16 [1] green f16 [2] cov rgba f16 [3] output rgba f16
int launch_xtrans_fix(const vkb_launch_t *l)
{
  const vkb_image_t *in = l->conn, *g = l->conn + 1, *cov = l->conn + 2, *out = l->conn + 3;
  VKB_REQUIRE(in->format == VKB_TOKEN_F16 && g->format == VKB_TOKEN_F16 && cov->format == VKB_TOKEN_F16 && out->format == VKB_TOKEN_F16);
  k_xtrans_fix<<<grid3(out->wd / 3 + 1, out->ht / 3 + 1), dim3(32, 4), 0, l->stream>>>((const __half *)in->data, (const __half *)g->data, in->wd, in->ht,
      (const uint2 *)cov->data, cov->wd, cov->ht, (uint2 *)out->data);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

VKB_NS_END

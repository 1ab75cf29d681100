// gaussian splatting demosaic for bayer and x-trans.
// replaces src/pipe/modules/demosaic/{gauss,splat,fix}.comp (wired by demosaic/main.c:159-202, method 0).
// `down.comp` is not needed: gauss.comp declares its output as img_in but never samples it.
// the 1:1 shared/resample node vkdt-cli appends (demosaic/main.c:193-201) is the identity and is elided.
#include "common.cuh"

struct demosaic_push_t { float wb[4]; uint32_t filters; };
int launch_bayer_splat(const vkb_launch_t *l);
int launch_bayer_fix(const vkb_launch_t *l);
int launch_xtrans_splat(const vkb_launch_t *l);
int launch_xtrans_fix(const vkb_launch_t *l);

// ---- gauss: green-only structure tensor per block -> (eval.xy, axis snapped evec) (gauss.comp:17-126) ----
// the tap pattern is a compile time property of the cfa: both loops unroll completely and the taps stay in registers.
// i / p and j / p are evaluated as i * (1 / p): i, j are in {-1, 0, 1, 2}, and scaling a correctly rounded quotient
// by 0, +-1 or 2 is exact, so one division per tap serves all three sums bit for bit.
template <bool xtrans>
__global__ void __launch_bounds__(256, 5) k_demosaic_gauss(const __half *__restrict__ orig, int iw, int ih,
    uint2 *__restrict__ out, int ow, int oh, const band_t bd)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  constexpr int blk = xtrans ? 3 : 2, lo = xtrans ? 0 : -1;
  // the "white" (weights p, p^2) and "black" (weights 1/p, 1/p^2) estimates go through identical arithmetic: they run as
  // the two lanes of packed fp32 pairs (FMUL2 / FFMA2; lane lo = white, hi = black), like in denoise's downcov.
  // first moments: i, j are 0, +-1, 2, the products are exact, so the fused multiply-add ptxas makes of them rounds like
  // the separate operations.  second moments: inexact products, multiplied packed and summed as scalars, unfused.
  f2 px[16];
  f2 MX = pk2(0.0f, 0.0f), MY = MX, SM = MX;
  int xi[4], yi[4];
#pragma unroll
  for(int i = lo; i < 3; i++) { xi[i - lo] = mirror1(blk * x + i, iw); yi[i - lo] = mirror1(blk * y + i, ih); }
#pragma unroll
  for(int j = lo; j < 3; j++)
#pragma unroll
  for(int i = lo; i < 3; i++)
  {
    if(xtrans ? (((j + i) & 1) == 1) : (((j + i) & 1) != 1)) continue;
    const float p = ld_h(orig, iw, xi[i - lo], yi[j - lo]);
    const f2 L = pk2(p, div_g(1.0f, p));
    px[4 * (j - lo) + (i - lo)] = L;
    MX = add2(MX, mul2(pk2((float)i, (float)i), L));
    MY = add2(MY, mul2(pk2((float)j, (float)j), L));
    SM = add2(SM, L);
  }
  const float smw = lo2(SM), smb = hi2(SM);
  const float mwx = div_g(lo2(MX), smw), mwy = div_g(lo2(MY), smw), mbx = div_g(hi2(MX), smb), mby = div_g(hi2(MY), smb);
  float Sw0 = 0, Sw1 = 0, Sw2 = 0, Sw3 = 0, Sb0 = 0, Sb1 = 0, Sb2 = 0, Sb3 = 0;
  f2 SS = pk2(0.0f, 0.0f);
#pragma unroll
  for(int j = lo; j < 3; j++)
#pragma unroll
  for(int i = lo; i < 3; i++)
  {
    if(xtrans ? (((j + i) & 1) == 1) : (((j + i) & 1) != 1)) continue;
    const float p = lo2(px[4 * (j - lo) + (i - lo)]);
    const float p2 = p * p;
    const f2 Q = pk2(p2, div_g(1.0f, p2));
    const f2 P0 = pk2((float)i - mwx, (float)i - mbx), P1 = pk2((float)j - mwy, (float)j - mby);
    const f2 T0 = mul2(Q, P0), T1 = mul2(Q, P1);
    const f2 A = mul2(T0, P0), B = mul2(T0, P1), C = mul2(T1, P0), D = mul2(T1, P1);
    Sw0 += lo2(A); Sw1 += lo2(B); Sw2 += lo2(C); Sw3 += lo2(D);
    Sb0 += hi2(A); Sb1 += hi2(B); Sb2 += hi2(C); Sb3 += hi2(D);
    SS = add2(SS, Q);
  }
  const float sw = lo2(SS), sb = hi2(SS);
  Sw0 = div_g(Sw0, sw); Sw1 = div_g(Sw1, sw); Sw2 = div_g(Sw2, sw); Sw3 = div_g(Sw3, sw);
  Sb0 = div_g(Sb0, sb); Sb1 = div_g(Sb1, sb); Sb2 = div_g(Sb2, sb); Sb3 = div_g(Sb3, sb);
  const bool usew = (Sw0 * Sw3 - Sw1 * Sw2) < (Sb0 * Sb3 - Sb1 * Sb2);
  float e0, e1, v0x, v0y, v1x, v1y;
  evd2x2(usew ? Sw0 : Sb0, usew ? Sw2 : Sb2, usew ? Sw3 : Sb3, e0, e1, v0x, v0y, v1x, v1y);
  if(!xtrans)
  {
    e0 *= 0.2f; e1 *= 0.2f;
    if(fabsf(v0x) > fabsf(v0y)) { v0x = 1; v0y = 0; }
    else                        { v0x = 0; v0y = 1; }
  }
  else
  {
    e0 *= 0.4f; e1 *= 0.4f;
    if     (fabsf(v0x) > 2.f * fabsf(v0y)) { v0x = 1; v0y = 0; }
    else if(fabsf(v0y) > 2.f * fabsf(v0x)) { v0x = 0; v0y = 1; }
    else e0 = e1 = .1f;
  }
  st_rgba(out, ow, x, y, make_float4(e0, e1, v0x, v0y));
}

// ---- splat: green by anisotropic gaussian weights (splat.comp:16-140) ----
__global__ void __launch_bounds__(256) k_demosaic_splat(const __half *__restrict__ in, int w, int h,
    const uint2 *__restrict__ gauss, int gw, int gh, __half *__restrict__ out, int xtrans)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= w || y >= h) return;
  const float4 cov = xtrans ? ld_rgba_clamp(gauss, gw, gh, x / 3, y / 3) : ld_rgba_clamp(gauss, gw, gh, (x + 1) / 2, (y + 1) / 2);
  const float e0 = clampf(cov.x, 0.01f, 25.0f), e1 = clampf(cov.y, 0.01f, 25.0f);
  const int r = xtrans ? 2 : 1;
  float g = 0.0f, wg = 0.0f;
  for(int j = -r; j <= r; j++) for(int i = -r; i <= r; i++)
  {
    const int c = xtrans ? xtrans_colour(x + i + 6, y + j + 6) : bayer_colour(x + i, y + j);
    if(c != 1) continue; // only the green accumulator reaches the output
    int px = x + i, py = y + j;
    if(px < 0) px += 6;
    if(py < 0) py += 6;
    if(px >= w) px -= 6;
    if(py >= h) py -= 6;
    const float col = ld_h_clamp(in, w, h, px, py);
    const float of0 = cov.z * (float)i + cov.w * (float)j;
    const float of1 = -cov.w * (float)i + cov.z * (float)j;
    float weight = clampf(m_exp(-0.5f * (of0 / e0 * of0 + of1 / e1 * of1)), 1e-4f, 1.0f);
    if(i == 0 && j == 0) weight = 666.0f;
    g += col * weight;
    wg += weight;
  }
  out[(size_t)y * w + x] = __float2half_rn(g / fmaxf(1e-8f, wg));
}

VKB_DEV float fix_gauss(float e0, float e1, float cz, float cw, int i, int j)
{
  const float of0 = cz * (float)i + cw * (float)j;
  const float of1 = -cw * (float)i + cz * (float)j;
  return clampf(m_exp(-0.5f * (of0 / e0 * of0 + of1 / e1 * of1)), 1e-3f, 1.0f);
}

// ---- fix: red/blue by interpolating the ratio to green (fix.comp:25-135) ----
__global__ void __launch_bounds__(256) k_demosaic_fix(const __half *__restrict__ in, const __half *__restrict__ green, int w, int h,
    const uint2 *__restrict__ covimg, int gw, int gh, uint2 *__restrict__ out, int xtrans, int fixup)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= w || y >= h) return;
  const float gc = ld_h(green, w, x, y);
  float4 cov = xtrans ? ld_rgba_clamp(covimg, gw, gh, (x + 1) / 3, (y + 1) / 3) : ld_rgba_clamp(covimg, gw, gh, (x + 1) / 2, (y + 1) / 2);
  const int r = xtrans ? clampi(fixup + 2, 2, 3) : clampi(fixup + 1, 1, 2);
  if(xtrans) { cov.x = clampf(cov.x, 1.f, 10.f); cov.y = clampf(cov.y, 1.f, 10.f); }
  else       { cov.x = clampf(cov.x, 1.0f, 49.f); cov.y = clampf(cov.y, 1.0f, 49.f); }
  const float ks = xtrans ? 3.0f : 2.0f;
  float rgb[3] = {0, 0, 0}, wt[3] = {0, 0, 0};
  for(int j = -r; j <= r; j++) for(int i = -r; i <= r; i++)
  {
    const int px = x + i, py = y + j;
    const int c = xtrans ? xtrans_colour(px, py) : bayer_colour(px, py);
    if(c == 1) { rgb[1] = gc; wt[1] = 1.0f; continue; }
    const float gh_ = ld_h_mirror(green, w, h, px, py);
    const float col = ld_h_mirror(in, w, h, px, py);
    const float weight = fix_gauss(ks * cov.x, ks * cov.y, cov.z, cov.w, i, j);
    if(xtrans) { const float corr = (1e-4f + gc) / (1e-4f + gh_); rgb[c] += col * corr * weight; }
    else rgb[c] += col * (1e-4f + gc) / (1e-4f + gh_) * weight;
    wt[c] += weight;
  }
  st_rgba(out, w, x, y, make_float4(rgb[0] / fmaxf(1e-8f, wt[0]), rgb[1] / fmaxf(1e-8f, wt[1]), rgb[2] / fmaxf(1e-8f, wt[2]), 1.0f));
}

static inline dim3 grid2d(unsigned w, unsigned h) { return dim3(vkb_cdiv(w, 32), vkb_cdiv(h, 8)); }
static const dim3 blk2d(32, 8);

// conn: [0] (unused `down` output, may be null) [1] orig mosaic f16 [2] output rgba f16
static int launch_demosaic_gauss(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 3 && l->push_size >= sizeof(demosaic_push_t));
  const demosaic_push_t *pc = (const demosaic_push_t *)l->push;
  const vkb_image_t *orig = l->conn + 1, *out = l->conn + 2;
  VKB_REQUIRE(orig->chan == 1 && orig->format == VKB_TOKEN_F16 && out->chan == 4 && out->format == VKB_TOKEN_F16);
  dim3 grid = grid2d(out->wd, out->ht);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  if(pc->filters == 9)
    k_demosaic_gauss<true><<<grid, blk2d, 0, l->stream>>>((const __half *)orig->data, orig->wd, orig->ht,
        (uint2 *)out->data, out->wd, out->ht, bd);
  else
    k_demosaic_gauss<false><<<grid, blk2d, 0, l->stream>>>((const __half *)orig->data, orig->wd, orig->ht,
        (uint2 *)out->data, out->wd, out->ht, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("demosaic", "gauss", launch_demosaic_gauss);

// conn: [0] input mosaic f16 [1] gauss rgba f16 [2] output green f16
static int launch_demosaic_splat(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 3 && l->push_size >= sizeof(demosaic_push_t));
  const demosaic_push_t *pc = (const demosaic_push_t *)l->push;
  const vkb_image_t *in = l->conn, *g = l->conn + 1, *out = l->conn + 2;
  VKB_REQUIRE(in->chan == 1 && g->chan == 4 && out->chan == 1 && in->wd == out->wd && in->ht == out->ht);
  if(pc->filters != 9) return launch_bayer_splat(l); // one thread per shifted 2x2 block, k_demosaic_bayer.cu
  return launch_xtrans_splat(l); // one thread per 3x3 block, k_demosaic_xtrans.cu
  k_demosaic_splat<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const __half *)in->data, in->wd, in->ht,
      (const uint2 *)g->data, g->wd, g->ht, (__half *)out->data, pc->filters == 9);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("demosaic", "splat", launch_demosaic_splat);

// conn: [0] input mosaic f16 [1] green f16 [2] cov rgba f16 [3] output rgba f16.  params: { int colour(fixup); int method }
static int launch_demosaic_fix(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 4 && l->push_size >= sizeof(demosaic_push_t) && l->params_size >= 4);
  const demosaic_push_t *pc = (const demosaic_push_t *)l->push;
  const vkb_image_t *in = l->conn, *g = l->conn + 1, *cov = l->conn + 2, *out = l->conn + 3;
  VKB_REQUIRE(in->chan == 1 && g->chan == 1 && cov->chan == 4 && out->chan == 4 && in->wd == out->wd && g->wd == out->wd);
  if(pc->filters != 9 && ((const int32_t *)l->params)[0] <= 0) return launch_bayer_fix(l); // radius 1 fast path
  if(pc->filters == 9 && ((const int32_t *)l->params)[0] <= 0) return launch_xtrans_fix(l); // radius 2: one thread per 3x3 block, k_demosaic_xtrans.cu
  k_demosaic_fix<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const __half *)in->data, (const __half *)g->data, in->wd, in->ht,
      (const uint2 *)cov->data, cov->wd, cov->ht, (uint2 *)out->data, pc->filters == 9, ((const int32_t *)l->params)[0]);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("demosaic", "fix", launch_demosaic_fix);

VKB_NS_END

// (b200, llapr0): llap/curve.comp + the first llap/reduce.comp, two gamma layers per instruction.
// the ten remapping curves of a pixel are the same arithmetic on different constants, and so are the eleven layers of
// the reduce: pairs of layers travel as the two lanes of Blackwell's packed fp32 instructions (FMUL2 / FFMA2 / FADD2)
// and share one half2 slot in the shared memory tile, one f16x2 conversion and one 32-bit shared store / load.
// per lane the operations are those of llap_curve_k / the reduce of k_llap.cu (comparisons, selects, |x|, saturation and
// the ex2 stay scalar, the packed pipe has no such forms); the reduce only multiplies by 1/2 and 1/4, exact in any
// contraction.  replaces the scalar k_llap_reduce0 (0.85 ms at 61 MP).
#include "common.cuh"

#define NUM_GAMMA 10
#define NL (NUM_GAMMA + 1)
#define NP ((NL + 1) / 2)      // layer pairs: (0,1) (2,3) (4,5) (6,7) (8,9) (grey, -)
#define R0_TW 65
#define R0_TH 17

struct llap_params_t { float sigma, shadows, hilights, clarity; };

VKB_DEV f2 sub2(f2 a, f2 b) { return add2(a, pk2(-lo2(b), -hi2(b))); }
VKB_DEV f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

// curve(x, g) of llap/curve.comp:40-63 for the two gamma values (ga, gb): see llap_curve_k in k_llap.cu
template <bool CLARITY>
VKB_DEV f2 curve2(float x, float ga, float gb, const llap_params_t &p, float inv2s, float invd)
{
  const float ca = x - ga, cb = x - gb;
  const f2 c = pk2(ca, cb), g = pk2(ga, gb);
  const f2 ssig = pk2(ca > 0.0f ? p.sigma : -p.sigma, cb > 0.0f ? p.sigma : -p.sigma);
  const f2 shhi = pk2(ca > 0.0f ? p.shadows : p.hilights, cb > 0.0f ? p.shadows : p.hilights);
  // far branch: g + ssigma + shadhi * (c - ssigma)
  const f2 far = fma2(shhi, sub2(c, ssig), add2(g, ssig));
  // near branch: g + ssigma * 2 * mt * t + t2 * (ssigma + ssigma * shadhi)
  const f2 t = pk2(__saturatef(fabsf(ca) * inv2s), __saturatef(fabsf(cb) * inv2s));
  const f2 t2 = mul2(t, t), mt = sub2(pk2(1.0f, 1.0f), t);
  const f2 near = fma2(t2, fma2(ssig, shhi, ssig), add2(g, mul2(mul2(mul2(ssig, pk2(2.0f, 2.0f)), mt), t)));
  const float lim = 2 * p.sigma;
  f2 val = pk2(fabsf(ca) > lim ? lo2(far) : lo2(near), fabsf(cb) > lim ? hi2(far) : hi2(near));
  if(CLARITY)
  { // + clarity * c * exp(-c * c * invd), exp(a) = ex2(a * log2 e): the same three roundings as the scalar code
    const f2 a = mul2(mul2(mul2(c, c), pk2(-invd, -invd)), pk2(1.4426950408889634f, 1.4426950408889634f));
    const f2 e = pk2(ex2_ftz(lo2(a)), ex2_ftz(hi2(a)));
    val = fma2(mul2(pk2(p.clarity, p.clarity), c), e, val);
  }
  else val = fma2(pk2(0.0f, 0.0f), c, val);
  return val;
}

VKB_DEV f2 avg2(f2 a, f2 b) { const f2 h = pk2(0.5f, 0.5f); return add2(mul2(a, h), mul2(b, h)); }

template <bool CLARITY>
__global__ void __launch_bounds__(256, 5) k_llap_reduce0_p(const uint2 *__restrict__ in, int iw, int ih,
    __half *__restrict__ out, int ow, int oh, const llap_params_t p, const band_t bd)
{
  __shared__ __align__(16) __half2 tile[NP][R0_TH][R0_TW + 1];
  const int tx0 = blockIdx.x * 64 - 1, ty0 = BAND_BY * 16 - 1;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const float inv2s = 1.0f / (2.0f * p.sigma), invd = 1.0f / (2.0f * p.sigma * p.sigma / 3.0f);
  const bool big = iw >= 66 && ih >= 18; // the tile overhangs the image by < one tile: the cheap mirror is enough
  for(int t = tid; t < R0_TW * R0_TH; t += 256)
  {
    const int lx = t % R0_TW, ly = t / R0_TW;
    const int gx = big ? mirror1(tx0 + lx, iw) : mirrori(tx0 + lx, iw), gy = big ? mirror1(ty0 + ly, ih) : mirrori(ty0 + ly, ih);
    const float4 px = ld_rgba(in, iw, gx, gy);
    const float y = lum2020(clampf(px.x, -1000.0f, 1000.0f), clampf(px.y, -1000.0f, 1000.0f), clampf(px.z, -1000.0f, 1000.0f)); // curve.comp:72
#pragma unroll
    for(int q = 0; q < NUM_GAMMA / 2; q++)
    {
      const f2 v = curve2<CLARITY>(y, (float)(2 * q) / (NUM_GAMMA - 1.0f), (float)(2 * q + 1) / (NUM_GAMMA - 1.0f), p, inv2s, invd);
      tile[q][ly][lx] = __floats2half2_rn(lo2(v), hi2(v));
    }
    tile[NP - 1][ly][lx] = __floats2half2_rn(y, 0.0f);
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x, y = BAND_BY * 8 + threadIdx.y;
  if(x >= ow || y >= oh || BAND_SKIP(y)) return;
  const int lx = 2 * threadIdx.x, ly = 2 * threadIdx.y; // tile coords of texel (2x-1, 2y-1)
  const size_t plane = (size_t)ow * oh;
  __half *o = out + (size_t)y * ow + x;
#pragma unroll
  for(int q = 0; q < NP; q++)
  {
    f2 t[3][3];
#pragma unroll
    for(int j = 0; j < 3; j++)
#pragma unroll
      for(int i = 0; i < 3; i++)
      {
        const float2 f = __half22float2(tile[q][ly + j][lx + i]);
        t[j][i] = pk2(f.x, f.y);
      }
    // sample_semisoft: four bilinear taps with weights 1/2, summed, / 4
    const f2 b00 = avg2(avg2(t[0][0], t[0][1]), avg2(t[1][0], t[1][1]));
    const f2 b10 = avg2(avg2(t[0][1], t[0][2]), avg2(t[1][1], t[1][2]));
    const f2 b01 = avg2(avg2(t[1][0], t[1][1]), avg2(t[2][0], t[2][1]));
    const f2 b11 = avg2(avg2(t[1][1], t[1][2]), avg2(t[2][1], t[2][2]));
    const f2 r = mul2(add2(add2(add2(b00, b10), b01), b11), pk2(0.25f, 0.25f));
    o[(size_t)(2 * q) * plane] = __float2half_rn(lo2(r));
    if(2 * q + 1 < NL) o[(size_t)(2 * q + 1) * plane] = __float2half_rn(hi2(r));
  }
}

// conn: [0] input rgba f16 (level 0), [1] output y f16 x 11 layers (level 1).  params: llap params
int launch_llapr0_packed(const vkb_launch_t *l)
{
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  const llap_params_t *lp = (const llap_params_t *)l->params;
  dim3 grid(vkb_cdiv(out->wd, 32), vkb_cdiv(out->ht, 8)), block(32, 8);
  const band_t bd = band_of(l, 1, 8, out->ht, &grid.y);
  if(!grid.y) return VKB_OK;
  if(lp->clarity == 0.0f)
    k_llap_reduce0_p<false><<<grid, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, (__half *)out->data, out->wd, out->ht, *lp, bd);
  else
    k_llap_reduce0_p<true><<<grid, block, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, (__half *)out->data, out->wd, out->ht, *lp, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

VKB_NS_END

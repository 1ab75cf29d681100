// RCD (ratio corrected demosaicing) for bayer, demosaic:method 1.
// replaces src/pipe/modules/demosaic/rcd_conv.comp:11-50 and rcd_fill.comp:26-163 (wired by demosaic/main.c:116-156).
// rcd_fill keeps the reference's structure — native colours staged in shared memory, then green at r/b, the missing one of
// r/b at b/r (diagonal discriminator), r and b at green (v/h discriminator), f16 rounding at every shared-memory store
// like the shader's float16_t planes — but on 32x16 output tiles with a 6 px halo: the four steps reach 6 / 5 / 3 px, so
// the result does not depend on where tile seams fall (the shader's 64x32 tiles with a 3 px border read stale and
// out-of-bounds shared memory next to seams; see oracle/o_rcd.c).  planes clamp to the image edge.
// --fmad=false: the discriminators are ratios of squared sums compared against each other.
#include "common.cuh"

struct demosaic_push_t { float wb[4]; uint32_t filters; };
VKB_DEV int rcd_col(int x, int y) { return (((x + y) & 1) == 1) ? 1 : ((y & 1) == 0 ? 0 : 2); }

__global__ void __launch_bounds__(256) k_rcd_conv(const __half *__restrict__ cfa, int w, int h,
    __half *__restrict__ vh, __half *__restrict__ pq, __half *__restrict__ lp)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= w || y >= h) return;
  const int hw = w / 2;
#define C(X, Y) ld_h(cfa, w, mirrori(x + (X), w), mirrori(y + (Y), h))
  float v = 0.0f, hh = 0.0f;
#pragma unroll
  for(int i = -1; i <= 1; i++)
  {
    v  += C(i,-3) - 3.0f * C(i,-2) - C(i,-1) + 6.0f * C(i,0) - C(i,1) - 3.0f * C(i,2) + C(i,3);
    hh += C(-3,i) - 3.0f * C(-2,i) - C(-1,i) + 6.0f * C(0,i) - C(1,i) - 3.0f * C(2,i) + C(3,i);
  }
  v *= v; hh *= hh;
  vh[(size_t)y * w + x] = __float2half_rn(v / (1e-5f + v + hh));
  if(x / 2 >= hw) return;
  if(((x + y) & 1) == 0)
  {
    float p = 1e-5f, q = 1e-5f;
#pragma unroll
    for(int i = -1; i <= 1; i++)
    {
      p += C(-3+i,-3+i) - C(1+i,-1+i) - C( 1+i,1+i) + C( 3+i,3+i) - 3.0f * (C(-2+i,-2+i) + C( 2+i,2+i)) + 6.0f * C(i, i);
      q += C( 3+i,-3-i) - C(1+i,-1-i) - C(-1+i,1-i) + C(-3+i,3-i) - 3.0f * (C( 2+i,-2-i) + C(-2+i,2-i)) + 6.0f * C(i,-i);
    }
    p *= p; q *= q;
    pq[(size_t)y * hw + x / 2] = __float2half_rn(p / (p + q));
  }
  else
  {
    float l = 0.0f;
    const int off = ((x & 1) == 1) ? -1 : 1;
    const float wt[3] = {0.5f, 1.0f, 0.5f};
#pragma unroll
    for(int j = -1; j <= 1; j++)
#pragma unroll
      for(int i = -1; i <= 1; i++) l += wt[j + 1] * wt[i + 1] * C(i + off, j);
    lp[(size_t)y * hw + x / 2] = __float2half_rn(fmaxf(1e-6f, l));
  }
#undef C
}

#define RT_OW 32
#define RT_OH 16
#define RT_HALO 6
#define RT_W (RT_OW + 2 * RT_HALO)
#define RT_H (RT_OH + 2 * RT_HALO)

__global__ void __launch_bounds__(256) k_rcd_fill(const __half *__restrict__ cfa, const __half *__restrict__ vh, const __half *__restrict__ pq,
    const __half *__restrict__ lp, int w, int h, uint2 *__restrict__ out, float wbr, float wbg, float wbb)
{
  __shared__ float P[3][RT_H][RT_W];
  const int gx0 = blockIdx.x * RT_OW - RT_HALO, gy0 = blockIdx.y * RT_OH - RT_HALO;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int hw = w / 2;
  const float eps = 1e-5f;
  const float wb[3] = { wbr, wbg, wbb };
#define CFA(X, Y) ld_h_clamp(cfa, w, h, (X), (Y))
#define VH(X, Y)  ld_h(vh, w, mirrori((X), w), mirrori((Y), h))
#define HALF(IMG, X, Y) ld_h(IMG, hw, mirrori(((X) + ((((Y) & 1) == 1) ? 0 : 1)) / 2, hw), mirrori((Y), h))
  // planes clamp to the image edge; the clamped position is always inside the tile for the reaches used below
#define S(K, X, Y) P[K][clampi((Y), 0, h - 1) - gy0][clampi((X), 0, w - 1) - gx0]
  // 1) native colours
  for(int t = tid; t < RT_W * RT_H; t += 256)
  {
    const int ly = t / RT_W, lx = t - ly * RT_W, x = gx0 + lx, y = gy0 + ly;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    if(x >= 0 && y >= 0 && x < w && y < h)
    {
      const int c = rcd_col(x, y);
      const float v = clampf(CFA(x, y), 0.0f, 65535.0f);
      if(c == 0) r = f16r(wb[0] * v); else if(c == 1) g = f16r(wb[1] * v); else b = f16r(wb[2] * v);
    }
    P[0][ly][lx] = r; P[1][ly][lx] = g; P[2][ly][lx] = b;
  }
  __syncthreads();
  // 2) green at r/b sites, halo 5 (reads global only)
  for(int t = tid; t < (RT_W - 2) * (RT_H - 2); t += 256)
  {
    const int ly = 1 + t / (RT_W - 2), lx = 1 + t % (RT_W - 2), x = gx0 + lx, y = gy0 + ly;
    if(x < 0 || y < 0 || x >= w || y >= h || rcd_col(x, y) == 1) continue;
    const float vhc = VH(x, y);
    const float vhn = 0.25f * (VH(x-1, y-1) + VH(x+1, y-1) + VH(x-1, y+1) + VH(x+1, y+1));
    const float vh_discr = fabsf(0.5f - vhc) < fabsf(0.5f - vhn) ? vhn : vhc;
    const float N_grad = eps + fabsf(CFA(x,y-1) - CFA(x,y+1)) + fabsf(CFA(x,y) - CFA(x,y-2)) + fabsf(CFA(x,y-1) - CFA(x,y-3)) + fabsf(CFA(x,y-2) - CFA(x,y-4));
    const float S_grad = eps + fabsf(CFA(x,y-1) - CFA(x,y+1)) + fabsf(CFA(x,y) - CFA(x,y+2)) + fabsf(CFA(x,y+1) - CFA(x,y+3)) + fabsf(CFA(x,y+2) - CFA(x,y+4));
    const float W_grad = eps + fabsf(CFA(x-1,y) - CFA(x+1,y)) + fabsf(CFA(x,y) - CFA(x-2,y)) + fabsf(CFA(x-1,y) - CFA(x-3,y)) + fabsf(CFA(x-2,y) - CFA(x-4,y));
    const float E_grad = eps + fabsf(CFA(x-1,y) - CFA(x+1,y)) + fabsf(CFA(x,y) - CFA(x+2,y)) + fabsf(CFA(x+1,y) - CFA(x+3,y)) + fabsf(CFA(x+2,y) - CFA(x+4,y));
    const float l0 = HALF(lp, x, y);
    const float N_est = CFA(x,y-1) * 2.0f * l0 / (eps + l0 + HALF(lp, x, y-2));
    const float S_est = CFA(x,y+1) * 2.0f * l0 / (eps + l0 + HALF(lp, x, y+2));
    const float W_est = CFA(x-1,y) * 2.0f * l0 / (eps + l0 + HALF(lp, x-2, y));
    const float E_est = CFA(x+1,y) * 2.0f * l0 / (eps + l0 + HALF(lp, x+2, y));
    const float v_est = clampf((S_grad * N_est + N_grad * S_est) / (N_grad + S_grad), 0.0f, 65534.0f);
    const float h_est = clampf((W_grad * E_est + E_grad * W_est) / (E_grad + W_grad), 0.0f, 65534.0f);
    P[1][ly][lx] = f16r(mixf(v_est, h_est, vh_discr));
  }
  __syncthreads();
  // 3) the missing one of r/b at b/r sites, halo 3: compute everything first, commit after a barrier (sites whose
  //    neighbours clamp onto another r/b site must see that site's pre-step value, like the whole-image restatement)
  float res3[4]; int n3 = 0;
  for(int t = tid; t < (RT_W - 6) * (RT_H - 6); t += 256, n3++)
  {
    const int ly = 3 + t / (RT_W - 6), lx = 3 + t % (RT_W - 6), x = gx0 + lx, y = gy0 + ly;
    res3[n3] = 0.0f;
    if(x < 0 || y < 0 || x >= w || y >= h || rcd_col(x, y) == 1) continue;
    const int k = rcd_col(x, y) == 0 ? 2 : 0; // plane to fill
    const float pqc = HALF(pq, x, y);
    const float pqn = 0.25f * (HALF(pq, x-1, y-1) + HALF(pq, x+1, y-1) + HALF(pq, x-1, y+1) + HALF(pq, x+1, y+1));
    const float pq_discr = fabsf(0.5f - pqc) < fabsf(0.5f - pqn) ? pqn : pqc;
#define sc(X, Y) S(k, X, Y)
#define sg(X, Y) S(1, X, Y)
    const float NW_grad = eps + fabsf(sc(x-1,y-1) - sc(x+1,y+1)) + fabsf(sc(x-1,y-1) - sc(x-3,y-3)) + fabsf(sg(x,y) - sg(x-2,y-2));
    const float NE_grad = eps + fabsf(sc(x+1,y-1) - sc(x-1,y+1)) + fabsf(sc(x+1,y-1) - sc(x+3,y-3)) + fabsf(sg(x,y) - sg(x+2,y-2));
    const float SW_grad = eps + fabsf(sc(x+1,y-1) - sc(x-1,y+1)) + fabsf(sc(x-1,y+1) - sc(x-3,y+3)) + fabsf(sg(x,y) - sg(x-2,y+2));
    const float SE_grad = eps + fabsf(sc(x-1,y-1) - sc(x+1,y+1)) + fabsf(sc(x+1,y+1) - sc(x+3,y+3)) + fabsf(sg(x,y) - sg(x+2,y+2));
    const float NW_est = sc(x-1,y-1) - sg(x-1,y-1), NE_est = sc(x+1,y-1) - sg(x+1,y-1);
    const float SW_est = sc(x-1,y+1) - sg(x-1,y+1), SE_est = sc(x+1,y+1) - sg(x+1,y+1);
    const float p_est = (NW_grad * SE_est + SE_grad * NW_est) / (NW_grad + SE_grad);
    const float q_est = (NE_grad * SW_est + SW_grad * NE_est) / (NE_grad + SW_grad);
    res3[n3] = f16r(clampf(sg(x,y) + mixf(p_est, q_est, pq_discr), 0.0f, 65535.0f));
#undef sc
  }
  __syncthreads();
  n3 = 0;
  for(int t = tid; t < (RT_W - 6) * (RT_H - 6); t += 256, n3++)
  {
    const int ly = 3 + t / (RT_W - 6), lx = 3 + t % (RT_W - 6), x = gx0 + lx, y = gy0 + ly;
    if(x < 0 || y < 0 || x >= w || y >= h || rcd_col(x, y) == 1) continue;
    P[rcd_col(x, y) == 0 ? 2 : 0][ly][lx] = res3[n3];
  }
  __syncthreads();
  // 4) r and b at green sites + output (results are final: straight to global memory)
  for(int t = tid; t < RT_OW * RT_OH; t += 256)
  {
    const int ly = RT_HALO + t / RT_OW, lx = RT_HALO + t % RT_OW, x = gx0 + lx, y = gy0 + ly;
    if(x >= w || y >= h) continue;
    float r = P[0][ly][lx], b = P[2][ly][lx];
    const float g = P[1][ly][lx];
    if(rcd_col(x, y) == 1)
    {
      const float vhc = VH(x, y);
      const float vhn = 0.25f * (VH(x-1, y-1) + VH(x+1, y-1) + VH(x-1, y+1) + VH(x+1, y+1));
      const float vh_discr = fabsf(0.5f - vhc) < fabsf(0.5f - vhn) ? vhn : vhc;
      const float N1 = eps + fabsf(sg(x,y) - sg(x,y-2)), S1 = eps + fabsf(sg(x,y) - sg(x,y+2));
      const float W1 = eps + fabsf(sg(x,y) - sg(x-2,y)), E1 = eps + fabsf(sg(x,y) - sg(x+2,y));
#pragma unroll
      for(int c = 0; c < 2; c++)
      {
        const int k = c == 0 ? 0 : 2;
#define sc(X, Y) S(k, X, Y)
        const float SNabs = fabsf(sc(x,y-1) - sc(x,y+1)), EWabs = fabsf(sc(x-1,y) - sc(x+1,y));
        const float N_grad = N1 + SNabs + fabsf(sc(x,y-1) - sc(x,y-3));
        const float S_grad = S1 + SNabs + fabsf(sc(x,y+1) - sc(x,y+3));
        const float W_grad = W1 + EWabs + fabsf(sc(x-1,y) - sc(x-3,y));
        const float E_grad = E1 + EWabs + fabsf(sc(x+1,y) - sc(x+3,y));
        const float N_est = sc(x,y-1) - sg(x,y-1), S_est = sc(x,y+1) - sg(x,y+1);
        const float W_est = sc(x-1,y) - sg(x-1,y), E_est = sc(x+1,y) - sg(x+1,y);
        const float v_est = (N_grad * S_est + S_grad * N_est) / (N_grad + S_grad);
        const float h_est = (E_grad * W_est + W_grad * E_est) / (E_grad + W_grad);
        const float val = f16r(clampf(sg(x,y) + mixf(v_est, h_est, vh_discr), 0.0f, 65535.0f));
        if(c == 0) r = val; else b = val;
#undef sc
      }
    }
    st_rgba(out, w, x, y, make_float4(r / wbr, g / wbg, b / wbb, 1.0f));
  }
#undef sg
#undef S
#undef CFA
#undef VH
#undef HALF
}

// conn: [0] cfa mosaic f16, [1] vh f16 w x h, [2] pq f16 (w/2) x h, [3] lp f16 (w/2) x h
static int launch_rcd_conv(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 4);
  const vkb_image_t *cfa = l->conn, *vh = l->conn + 1, *pq = l->conn + 2, *lp = l->conn + 3;
  VKB_REQUIRE(cfa->chan == 1 && cfa->format == VKB_TOKEN_F16 && vh->chan == 1 && vh->wd == cfa->wd && vh->ht == cfa->ht);
  VKB_REQUIRE(pq->wd == cfa->wd / 2 && lp->wd == cfa->wd / 2 && pq->ht == cfa->ht && lp->ht == cfa->ht);
  k_rcd_conv<<<dim3(vkb_cdiv(cfa->wd, 32), vkb_cdiv(cfa->ht, 8)), dim3(32, 8), 0, l->stream>>>((const __half *)cfa->data, cfa->wd, cfa->ht,
      (__half *)vh->data, (__half *)pq->data, (__half *)lp->data);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("demosaic", "rcd_conv", launch_rcd_conv);

// conn: [0] cfa, [1] vh, [2] pq, [3] lp, [4] output rgba f16.  push: { vec4 wb; uint filters }
static int launch_rcd_fill(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 5 && l->push_size >= sizeof(demosaic_push_t));
  const demosaic_push_t *pc = (const demosaic_push_t *)l->push;
  const vkb_image_t *cfa = l->conn, *vh = l->conn + 1, *pq = l->conn + 2, *lp = l->conn + 3, *out = l->conn + 4;
  VKB_REQUIRE(cfa->chan == 1 && out->chan == 4 && out->format == VKB_TOKEN_F16 && out->wd == cfa->wd && out->ht == cfa->ht);
  VKB_REQUIRE(pc->filters != 9);
  k_rcd_fill<<<dim3(vkb_cdiv(out->wd, RT_OW), vkb_cdiv(out->ht, RT_OH)), dim3(32, 8), 0, l->stream>>>((const __half *)cfa->data,
      (const __half *)vh->data, (const __half *)pq->data, (const __half *)lp->data, cfa->wd, cfa->ht, (uint2 *)out->data, pc->wb[0], pc->wb[1], pc->wb[2]);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("demosaic", "rcd_fill", launch_rcd_fill);

VKB_NS_END

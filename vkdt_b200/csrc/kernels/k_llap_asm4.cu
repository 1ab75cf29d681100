// coarse levels of the local laplacian collapse (llap/assemble.comp:52-88), one thread per 2x2 output pixels.
// the four pixels (2k, 2k+1) x (2m, 2m+1) expand the same 5x5 coarse texels around (k, m), and which of those texels a
// sample_soft tap reads (and with which of the weights 0, 1/2, 1) is a property of the pixel's parity: with the parity a
// template parameter the shader's 3x3 bilinear taps become straight-line code in the shader's order, without the per tap
// weight tests and dynamic shared memory indexing of the per pixel kernel (k_llap.cu), and a window is read once for four
// pixels.  multiplications by 0, 1/2 and 1 are exact, so the results are bit identical to k_llap_assemble[_tiled]: this
// level feeds the next one and has to stay in lockstep with the restatement.
#include "common.cuh"

#define NUM_GAMMA 10
#define NL (NUM_GAMMA + 1)
#define A4_W 36
#define A4_H 12

VKB_DEV float gamma_from_i(int i) { return div_c((float)i, NUM_GAMMA - 1.0f); }
VKB_DEV int gamma_hi_from_v(float v)
{ // llap.glsl:17-22: 1 + #{ i in 1..8 : i/9 <= v }
  int hi = 1;
#pragma unroll
  for(int i = 1; i < NUM_GAMMA - 1; i++) hi += ((float)i / (NUM_GAMMA - 1.0f) <= v) ? 1 : 0;
  return hi;
}

// one axis of sample_soft for an output of parity D inside the 5 texel window k-2..k+2:
// even: taps {0|1 (1/2), 2, 3|4 (1/2)}, odd: {1, 2|3 (1/2), 4}
VKB_DEV constexpr int  tap_i0(int d, int t)   { return d ? (t == 0 ? 1 : (t == 1 ? 2 : 4)) : (t == 0 ? 0 : (t == 1 ? 2 : 3)); }
VKB_DEV constexpr bool tap_half(int d, int t) { return d ? t == 1 : t != 1; }

template <int DX, int DY>
VKB_DEV float expand_q(const float (&W)[5][5])
{ // gauss_expand() of k_llap.cu with the weights resolved: top/bot = a (1-ax) + b ax, v = top (1-ay) + bot ay, r += v, / 9
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

__global__ void __launch_bounds__(256, 4) k_llap_assemble4(const __half *__restrict__ coarse, const __half *__restrict__ l0,
    const __half *__restrict__ l1, int cw, int ch, __half *__restrict__ out, int ow, int oh, int first, const band_t bd)
{
  __shared__ float tile[NL + 1][A4_H][A4_W + 1];
  __shared__ int s_pmin, s_pmax;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if(tid == 0) { s_pmin = NUM_GAMMA; s_pmax = 0; }
  const int kx = blockIdx.x * 32 + threadIdx.x, ky = BAND_BY * 8 + threadIdx.y;
  const int cx0 = blockIdx.x * 32 - 2, cy0 = BAND_BY * 8 - 2;
  const size_t p0 = (size_t)ow * oh, p1 = (size_t)cw * ch;
  float v[4]; int hi[4];
  int mylo = NUM_GAMMA, myhi = 0;
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    const int x = 2 * kx + (q & 1), y = 2 * ky + (q >> 1);
    hi[q] = -1;
    if(x < ow && y < oh && !BAND_SKIP(ky))
    {
      v[q] = ld_h(l0 + NUM_GAMMA * p0, ow, x, y);
      hi[q] = gamma_hi_from_v(v[q]);
      mylo = min(mylo, hi[q] - 1); myhi = max(myhi, hi[q]);
    }
  }
  __syncthreads();
  mylo = __reduce_min_sync(0xffffffffu, mylo); myhi = __reduce_max_sync(0xffffffffu, myhi);
  if(threadIdx.x == 0) { atomicMin(&s_pmin, mylo); atomicMax(&s_pmax, myhi); }
  __syncthreads();
  const int pmin = s_pmin, pmax = s_pmax;
  const bool big = cw >= 40 && ch >= 16;
#pragma unroll
  for(int e = 0; e < 2; e++)
  {
    const int t = tid + e * 256;
    if(t >= A4_H * A4_W) break;
    const int r = t / A4_W, c = t - r * A4_W;
    const int gx = big ? mirror1(cx0 + c, cw) : mirrori(cx0 + c, cw), gy = big ? mirror1(cy0 + r, ch) : mirrori(cy0 + r, ch);
    const size_t off = (size_t)gy * cw + gx;
    if(pmax >= pmin) tile[NL][r][c] = __half2float(__ldg((first ? l1 + NUM_GAMMA * p1 : coarse) + off));
    for(int pl = pmin; pl <= pmax; pl++) tile[pl][r][c] = __half2float(__ldg(l1 + pl * p1 + off));
  }
  __syncthreads();
  if(hi[0] < 0) return;
  const int lx = kx - cx0 - 2, ly = ky - cy0 - 2; // window origin in the tile
  float res[4], e0[4], e1[4];
  int hmin = NUM_GAMMA, hmax = 0;
#pragma unroll
  for(int q = 0; q < 4; q++) if(hi[q] >= 0) { hmin = min(hmin, hi[q]); hmax = max(hmax, hi[q]); }
  // the collapsed coarse level, then every gamma layer one of the four pixels brackets (two, next to a boundary three)
  for(int pl = hmin - 2; pl <= hmax; pl++)
  {
    const float (*T)[A4_W + 1] = tile[pl == hmin - 2 ? NL : pl];
    float W[5][5];
#pragma unroll
    for(int r = 0; r < 5; r++)
#pragma unroll
      for(int c = 0; c < 5; c++) W[r][c] = T[ly + r][lx + c];
    const float t0 = expand_q<0, 0>(W), t1 = expand_q<1, 0>(W), t2 = expand_q<0, 1>(W), t3 = expand_q<1, 1>(W);
    const float t[4] = { t0, t1, t2, t3 };
#pragma unroll
    for(int q = 0; q < 4; q++)
    {
      if(pl == hmin - 2)  res[q] = t[q];
      else
      {
        if(pl == hi[q] - 1) e0[q] = t[q];
        if(pl == hi[q])     e1[q] = t[q];
      }
    }
  }
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    if(hi[q] < 0) continue;
    const int x = 2 * kx + (q & 1), y = 2 * ky + (q >> 1);
    const int lo = hi[q] - 1;
    const float glo = gamma_from_i(lo), ghi = gamma_from_i(hi[q]);
    const float a = clampf(div_f(v[q] - glo, ghi - glo), 0.0f, 1.0f);   // the spacing of two gamma values: 1 / 9
    const float lap0 = ld_h(l0 + lo * p0, ow, x, y) - e0[q];
    const float lap1 = ld_h(l0 + hi[q] * p0, ow, x, y) - e1[q];
    // explicit _rn ops: this blend feeds the next pyramid level, keep it unfused whatever the compiler flags say
    out[(size_t)y * ow + x] = __float2half_rn(__fadd_rn(__fadd_rn(res[q], __fmul_rn(lap0, 1.0f - a)), __fmul_rn(lap1, a)));
  }
}

// conn as (llap, assemble): [0] coarse y f16, [1] fine stack x11, [2] coarse stack x11, [3] fine out y f16
int launch_llap_assemble4(const vkb_launch_t *l, int first)
{
  const vkb_image_t *coarse = l->conn, *l0 = l->conn + 1, *l1 = l->conn + 2, *out = l->conn + 3;
  dim3 grid(vkb_cdiv(out->wd, 64), vkb_cdiv(out->ht, 16));
  const band_t bd = band_of(l, 2, 8, (out->ht + 1) / 2, &grid.y); // band image: the fine output, two rows per thread row
  if(!grid.y) return VKB_OK;
  k_llap_assemble4<<<grid, dim3(32, 8), 0, l->stream>>>((const __half *)coarse->data, (const __half *)l0->data,
      (const __half *)l1->data, l1->wd, l1->ht, (__half *)out->data, out->wd, out->ht, first, bd);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

VKB_NS_END

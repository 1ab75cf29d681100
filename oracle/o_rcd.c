/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/demosaic/rcd_conv.comp:11-50 and rcd_fill.comp:26-163 (bayer, demosaic:method 1).
 *
 * The reference evaluates rcd_fill on 64x32 shared-memory tiles with a 3 px border although the four in-place steps
 * reach up to 6 px (native colours) / 5 px (green at r/b) / 3 px (r/b at b/r) away from an output pixel, and reads
 * `shm[ind(x-3,y-3)]` below index 0 at the tile edge: results within ~3 px of a tile seam depend on stale / out of
 * bounds shared memory.  This restatement is the tiling-independent ideal: the same four steps on whole-image planes
 * (f16 rounding at every shared-memory store, like the float16_t planes of the shader). */
#include "o_common.h"
#include "vkdt_oracle.h"

static inline int rcd_col(int x, int y) { return (((x + y) & 1) == 1) ? 1 : ((y & 1) == 0 ? 0 : 2); }
static inline float cfa_tex(const oimg_t *c, int x, int y) { return c->p[(size_t)o_mirror(y, c->h) * c->w + o_mirror(x, c->w)]; }

/* rcd_conv.comp:11-50.  vh: w x h, pq and lp: (w/2) x h, all f16 */
void o_rcd_conv(const oimg_t *cfa, oimg_t *vh, oimg_t *pq, oimg_t *lp)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < cfa->h; y++) for(int x = 0; x < cfa->w; x++)
  {
#define C(X, Y) cfa_tex(cfa, x + (X), y + (Y))
    float v = 0.0f, h = 0.0f;
    for(int i = -1; i <= 1; i++)
    {
      v += C(i,-3) - 3.0f * C(i,-2) - C(i,-1) + 6.0f * C(i,0) - C(i,1) - 3.0f * C(i,2) + C(i,3);
      h += C(-3,i) - 3.0f * C(-2,i) - C(-1,i) + 6.0f * C(0,i) - C(1,i) - 3.0f * C(2,i) + C(3,i);
    }
    v *= v; h *= h;
    o_store1(vh, x, y, v / (1e-5f + v + h), 1);
    if(((x + y) & 1) == 0)
    {
      float p = 1e-5f, q = 1e-5f;
      for(int i = -1; i <= 1; i++)
      {
        p += C(-3+i,-3+i) - C(1+i,-1+i) - C( 1+i,1+i) + C( 3+i,3+i) - 3.0f * (C(-2+i,-2+i) + C( 2+i,2+i)) + 6.0f * C(i, i);
        q += C( 3+i,-3-i) - C(1+i,-1-i) - C(-1+i,1-i) + C(-3+i,3-i) - 3.0f * (C( 2+i,-2-i) + C(-2+i,2-i)) + 6.0f * C(i,-i);
      }
      p *= p; q *= q;
      o_store1(pq, x / 2, y, p / (p + q), 1);
    }
    else
    {
      float l = 0.0f;
      const int off = ((x & 1) == 1) ? -1 : 1;
      static const float w[3] = {0.5f, 1.0f, 0.5f};
      for(int j = -1; j <= 1; j++) for(int i = -1; i <= 1; i++) l += w[j+1] * w[i+1] * C(i + off, j);
      o_store1(lp, x / 2, y, o_max(1e-6f, l), 1);
    }
#undef C
  }
}

static inline float tex_c(const oimg_t *im, int x, int y) { return im->p[(size_t)o_mirror(y, im->h) * im->w + o_mirror(x, im->w)]; }
/* pq / lp live at (x + (y odd ? 0 : 1)) / 2: rcd_fill.comp:33-34 (C integer division, like glsl) */
static inline float tex_half(const oimg_t *im, int x, int y) { return tex_c(im, (x + (((y & 1) == 1) ? 0 : 1)) / 2, y); }

/* rcd_fill.comp:36-163 on whole-image planes */
void o_rcd_fill(const oimg_t *cfa, const oimg_t *vh, const oimg_t *pq, const oimg_t *lp, oimg_t *out, const float *wb)
{
  const int W = cfa->w, H = cfa->h;
  const float eps = 1e-5f;
  float *R = (float *)calloc((size_t)W * H, sizeof(float)), *G = (float *)calloc((size_t)W * H, sizeof(float)), *B = (float *)calloc((size_t)W * H, sizeof(float));
  float *P[3] = { R, G, B };
#define IDX(X, Y) ((size_t)o_clampi((Y), 0, H - 1) * W + o_clampi((X), 0, W - 1))
#define CFA(X, Y) o_fetch1(cfa, (X), (Y))
  /* fill */
#pragma omp parallel for schedule(static)
  for(int y = 0; y < H; y++) for(int x = 0; x < W; x++)
  {
    const int c = rcd_col(x, y);
    const float v = o_clamp(CFA(x, y), 0.0f, 65535.0f);
    for(int k = 0; k < 3; k++) P[k][(size_t)y * W + x] = k == c ? o_f16r(wb[k] * v) : 0.0f;
  }
  /* g@rb */
#pragma omp parallel for schedule(static)
  for(int y = 0; y < H; y++) for(int x = 0; x < W; x++)
  {
    if(rcd_col(x, y) == 1) continue;
    const float vhc = tex_c(vh, x, y);
    const float vhn = 0.25f * (tex_c(vh, x-1, y-1) + tex_c(vh, x+1, y-1) + tex_c(vh, x-1, y+1) + tex_c(vh, x+1, y+1));
    const float vh_discr = fabsf(0.5f - vhc) < fabsf(0.5f - vhn) ? vhn : vhc;
    const float N_grad = eps + fabsf(CFA(x,y-1) - CFA(x,y+1)) + fabsf(CFA(x,y) - CFA(x,y-2)) + fabsf(CFA(x,y-1) - CFA(x,y-3)) + fabsf(CFA(x,y-2) - CFA(x,y-4));
    const float S_grad = eps + fabsf(CFA(x,y-1) - CFA(x,y+1)) + fabsf(CFA(x,y) - CFA(x,y+2)) + fabsf(CFA(x,y+1) - CFA(x,y+3)) + fabsf(CFA(x,y+2) - CFA(x,y+4));
    const float W_grad = eps + fabsf(CFA(x-1,y) - CFA(x+1,y)) + fabsf(CFA(x,y) - CFA(x-2,y)) + fabsf(CFA(x-1,y) - CFA(x-3,y)) + fabsf(CFA(x-2,y) - CFA(x-4,y));
    const float E_grad = eps + fabsf(CFA(x-1,y) - CFA(x+1,y)) + fabsf(CFA(x,y) - CFA(x+2,y)) + fabsf(CFA(x+1,y) - CFA(x+3,y)) + fabsf(CFA(x+2,y) - CFA(x+4,y));
    const float l0 = tex_half(lp, x, y);
    const float N_est = CFA(x,y-1) * 2.0f * l0 / (eps + l0 + tex_half(lp, x, y-2));
    const float S_est = CFA(x,y+1) * 2.0f * l0 / (eps + l0 + tex_half(lp, x, y+2));
    const float W_est = CFA(x-1,y) * 2.0f * l0 / (eps + l0 + tex_half(lp, x-2, y));
    const float E_est = CFA(x+1,y) * 2.0f * l0 / (eps + l0 + tex_half(lp, x+2, y));
    const float v_est = o_clamp((S_grad * N_est + N_grad * S_est) / (N_grad + S_grad), 0.0f, 65534.0f);
    const float h_est = o_clamp((W_grad * E_est + E_grad * W_est) / (E_grad + W_grad), 0.0f, 65534.0f);
    G[(size_t)y * W + x] = o_f16r(o_mix(v_est, h_est, vh_discr));
  }
  /* rb@br: the missing one of r/b at b/r sites, diagonal discriminator */
  float *N3 = (float *)calloc((size_t)W * H, sizeof(float));
#pragma omp parallel for schedule(static)
  for(int y = 0; y < H; y++) for(int x = 0; x < W; x++)
  {
    if(rcd_col(x, y) == 1) continue;
    const int red = rcd_col(x, y) == 0;
    const float pqc = tex_half(pq, x, y);
    const float pqn = 0.25f * (tex_half(pq, x-1, y-1) + tex_half(pq, x+1, y-1) + tex_half(pq, x-1, y+1) + tex_half(pq, x+1, y+1));
    const float pq_discr = fabsf(0.5f - pqc) < fabsf(0.5f - pqn) ? pqn : pqc;
    const float *S = red ? B : R;
#define sc(X, Y) S[IDX(X, Y)]
#define sg(X, Y) G[IDX(X, Y)]
    const float NW_grad = eps + fabsf(sc(x-1,y-1) - sc(x+1,y+1)) + fabsf(sc(x-1,y-1) - sc(x-3,y-3)) + fabsf(sg(x,y) - sg(x-2,y-2));
    const float NE_grad = eps + fabsf(sc(x+1,y-1) - sc(x-1,y+1)) + fabsf(sc(x+1,y-1) - sc(x+3,y-3)) + fabsf(sg(x,y) - sg(x+2,y-2));
    const float SW_grad = eps + fabsf(sc(x+1,y-1) - sc(x-1,y+1)) + fabsf(sc(x-1,y+1) - sc(x-3,y+3)) + fabsf(sg(x,y) - sg(x-2,y+2));
    const float SE_grad = eps + fabsf(sc(x-1,y-1) - sc(x+1,y+1)) + fabsf(sc(x+1,y+1) - sc(x+3,y+3)) + fabsf(sg(x,y) - sg(x+2,y+2));
    const float NW_est = sc(x-1,y-1) - sg(x-1,y-1), NE_est = sc(x+1,y-1) - sg(x+1,y-1);
    const float SW_est = sc(x-1,y+1) - sg(x-1,y+1), SE_est = sc(x+1,y+1) - sg(x+1,y+1);
    const float p_est = (NW_grad * SE_est + SE_grad * NW_est) / (NW_grad + SE_grad);
    const float q_est = (NE_grad * SW_est + SW_grad * NE_est) / (NE_grad + SW_grad);
    N3[(size_t)y * W + x] = o_f16r(o_clamp(sg(x,y) + o_mix(p_est, q_est, pq_discr), 0.0f, 65535.0f));
#undef sc
  }
  /* the in-place writes of the shader touch only planes nobody reads in this step: commit them now */
  for(int y = 0; y < H; y++) for(int x = 0; x < W; x++) if(rcd_col(x, y) != 1)
  { if(rcd_col(x, y) == 0) B[(size_t)y * W + x] = N3[(size_t)y * W + x]; else R[(size_t)y * W + x] = N3[(size_t)y * W + x]; }
  free(N3);
  /* rb@g */
  float *NR = (float *)calloc((size_t)W * H, sizeof(float)), *NB = (float *)calloc((size_t)W * H, sizeof(float));
#pragma omp parallel for schedule(static)
  for(int y = 0; y < H; y++) for(int x = 0; x < W; x++)
  {
    if(rcd_col(x, y) != 1) continue;
    const float vhc = tex_c(vh, x, y);
    const float vhn = 0.25f * (tex_c(vh, x-1, y-1) + tex_c(vh, x+1, y-1) + tex_c(vh, x-1, y+1) + tex_c(vh, x+1, y+1));
    const float vh_discr = fabsf(0.5f - vhc) < fabsf(0.5f - vhn) ? vhn : vhc;
    const float N1 = eps + fabsf(sg(x,y) - sg(x,y-2)), S1 = eps + fabsf(sg(x,y) - sg(x,y+2));
    const float W1 = eps + fabsf(sg(x,y) - sg(x-2,y)), E1 = eps + fabsf(sg(x,y) - sg(x+2,y));
    for(int c = 0; c < 2; c++)
    {
      const float *S = c == 0 ? R : B;
#define sc(X, Y) S[IDX(X, Y)]
      const float SNabs = fabsf(sc(x,y-1) - sc(x,y+1)), EWabs = fabsf(sc(x-1,y) - sc(x+1,y));
      const float N_grad = N1 + SNabs + fabsf(sc(x,y-1) - sc(x,y-3));
      const float S_grad = S1 + SNabs + fabsf(sc(x,y+1) - sc(x,y+3));
      const float W_grad = W1 + EWabs + fabsf(sc(x-1,y) - sc(x-3,y));
      const float E_grad = E1 + EWabs + fabsf(sc(x+1,y) - sc(x+3,y));
      const float N_est = sc(x,y-1) - sg(x,y-1), S_est = sc(x,y+1) - sg(x,y+1);
      const float W_est = sc(x-1,y) - sg(x-1,y), E_est = sc(x+1,y) - sg(x+1,y);
      const float v_est = (N_grad * S_est + S_grad * N_est) / (N_grad + S_grad);
      const float h_est = (E_grad * W_est + W_grad * E_est) / (E_grad + W_grad);
      (c == 0 ? NR : NB)[(size_t)y * W + x] = o_f16r(o_clamp(sg(x,y) + o_mix(v_est, h_est, vh_discr), 0.0f, 65535.0f));
#undef sc
    }
  }
  for(int y = 0; y < H; y++) for(int x = 0; x < W; x++) if(rcd_col(x, y) == 1)
  { R[(size_t)y * W + x] = NR[(size_t)y * W + x]; B[(size_t)y * W + x] = NB[(size_t)y * W + x]; }
  free(NR); free(NB);
#undef sg
#pragma omp parallel for schedule(static)
  for(int y = 0; y < H; y++) for(int x = 0; x < W; x++)
  {
    const size_t i = (size_t)y * W + x;
    const float o[4] = { R[i] / wb[0], G[i] / wb[1], B[i] / wb[2], 1.0f };
    o_store4(out, x, y, o, 1);
  }
#undef IDX
#undef CFA
  free(R); free(G); free(B);
}

/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/crop/main.c:175-345 (host side) and crop/main.comp:20-57,
 * plus shared.glsl:47-96 (sample_catmull_rom). */
#include "o_common.h"
#include "vkdt_oracle.h"

/* core/gaussian_elimination.h:43-112, partial pivoting, row major */
static int gauss_tri(double *A, int *p, int n)
{
  p[n - 1] = n - 1;
  for(int k = 0; k < n; ++k)
  {
    int m = k;
    for(int i = k + 1; i < n; ++i) if(fabs(A[k + n * i]) > fabs(A[k + n * m])) m = i;
    p[k] = m;
    double t1 = A[k + n * m];
    A[k + n * m] = A[k + n * k];
    A[k + n * k] = t1;
    if(t1 == 0) return 0;
    for(int i = k + 1; i < n; ++i) A[k + n * i] /= -t1;
    if(k != m) for(int i = k + 1; i < n; ++i)
    {
      double t2 = A[i + n * m];
      A[i + n * m] = A[i + n * k];
      A[i + n * k] = t2;
    }
    for(int j = k + 1; j < n; ++j)
      for(int i = k + 1; i < n; ++i) A[i + n * j] += A[k + j * n] * A[i + k * n];
  }
  return 1;
}
static void gauss_back(const double *A, const int *p, double *b, int n)
{
  for(int k = 0; k < n - 1; ++k)
  {
    int m = p[k];
    double t = b[m];
    b[m] = b[k];
    b[k] = t;
    for(int i = k + 1; i < n; ++i) b[i] += A[k + n * i] * t;
  }
  for(int k = n - 1; k > 0; --k)
  {
    b[k] /= A[k + n * k];
    double t = b[k];
    for(int i = 0; i < k; ++i) b[i] -= A[k + n * i] * t;
  }
  b[0] /= A[0];
}
int o_gauss_solve(double *A, double *b, int n)
{
  int p[64];
  if(n > 64) return 0;
  int ok = gauss_tri(A, p, n);
  if(ok) gauss_back(A, p, b, n);
  return ok;
}

/* crop/main.c:175-226 */
void o_crop_get_crop_rot(uint32_t orientation, double wd, double ht, const float *p_crop, const float *p_rot, float *crop, float *rot)
{
  float rotation = p_rot[0];
  rot[0] = rotation;
  for(int k = 0; k < 4; k++) crop[k] = p_crop[k];
  if(rotation == 1337.0f)
  {
    if(orientation == 3)      rot[0] = 180.0f;
    else if(orientation == 8) rot[0] = 90.0f;
    else if(orientation == 6) rot[0] = 270.0f;
    else                      rot[0] = 0.0f;
  }
  if(crop[0] == 1.0 && crop[1] == 3.0 && crop[2] == 3.0 && crop[3] == 7.0)
  {
    double crw = wd > 400 ? 3.0 / wd : 0.0, crh = ht > 400 ? 3.0 / ht : 0.0;
    if((rot[0] >= 45 && rot[0] < 135) || (rot[0] >= 225 && rot[0] < 315))
    {
      crop[0] = 0.5 - (.5 - crh) * ht / wd;
      crop[2] = 0.5 - (.5 - crw) * wd / ht;
      crop[1] = 0.5 + (.5 - crh) * ht / wd;
      crop[3] = 0.5 + (.5 - crw) * wd / ht;
    }
    else
    {
      crop[0] = crw;
      crop[2] = crh;
      crop[1] = 1.0 - crw;
      crop[3] = 1.0 - crh;
    }
  }
}

/* crop/main.c:256-275: output full size from input full size */
void o_crop_roi_out(uint32_t orientation, uint32_t in_w, uint32_t in_h, const float *p_crop, const float *p_rot, uint32_t *out_w, uint32_t *out_h)
{
  float crop[4], rot;
  float w = in_w, h = in_h;
  o_crop_get_crop_rot(orientation, w, h, p_crop, p_rot, crop, &rot);
  float wd = crop[1] - crop[0];
  float ht = crop[3] - crop[2];
  float fw = in_w * wd, fh = in_h * ht;
  *out_w = (uint32_t)(32768 < fw ? 32768 : fw);
  *out_h = (uint32_t)(32768 < fh ? 32768 : fh);
}

/* crop/main.c:277-345: committed params f[0..11] H (3 x vec4 columns), f[12..15] rotation, f[16..19] crop window */
void o_crop_commit(uint32_t orientation, uint32_t in_w, uint32_t in_h, const float *p_perspect, const float *p_crop, const float *p_rot, float *f)
{
  float p[8];
  for(int k = 0; k < 4; k++)
  {
    p[2*k+0] = in_w * p_perspect[2*k+0];
    p[2*k+1] = in_h * p_perspect[2*k+1];
  }
  const float a = p[0], A = p[2], b = p[1], B = p[7];
  const float u[] = {a, b, A, b, A, B, a, B};
  double M[] = {
    u[0], u[1], 1, 0, 0, 0, -p[0]*u[0], -p[0]*u[1],
    u[2], u[3], 1, 0, 0, 0, -p[2]*u[2], -p[2]*u[3],
    u[4], u[5], 1, 0, 0, 0, -p[4]*u[4], -p[4]*u[5],
    u[6], u[7], 1, 0, 0, 0, -p[6]*u[6], -p[6]*u[7],
    0, 0, 0, u[0], u[1], 1, -p[1]*u[0], -p[1]*u[1],
    0, 0, 0, u[2], u[3], 1, -p[3]*u[2], -p[3]*u[3],
    0, 0, 0, u[4], u[5], 1, -p[5]*u[4], -p[5]*u[5],
    0, 0, 0, u[6], u[7], 1, -p[7]*u[6], -p[7]*u[7],
  };
  double r[] = {p[0], p[2], p[4], p[6], p[1], p[3], p[5], p[7], 1.0};
  o_gauss_solve(M, r, 8);
  f[ 0] = r[0]; f[ 1] = r[3]; f[ 2] = r[6]; f[ 3] = 0.0f;
  f[ 4] = r[1]; f[ 5] = r[4]; f[ 6] = r[7]; f[ 7] = 0.0f;
  f[ 8] = r[2]; f[ 9] = r[5]; f[10] = r[8]; f[11] = 0.0f;
  float crop[4], rot;
  float wd = in_w, ht = in_h;
  o_crop_get_crop_rot(orientation, wd, ht, p_crop, p_rot, crop, &rot);
  float rad = rot * 3.1415629 / 180.0f; /* sic, crop/main.c:329 */
  f[12] =  cosf(rad); f[13] = sinf(rad);
  f[14] = -sinf(rad); f[15] = cosf(rad);
  for(int k = 0; k < 4; k++) f[16+k] = crop[k];
}

/* shared.glsl:47-96 */
static void catmull_rom(const oimg_t *tex, float u, float v, float *res)
{
  const float sx = (float)tex->w, sy = (float)tex->h;
  const float spx = u * sx, spy = v * sy;
  const float t1x = floorf(spx - 0.5f) + 0.5f, t1y = floorf(spy - 0.5f) + 0.5f;
  const float f[2] = { spx - t1x, spy - t1y };
  float w0[2], w1[2], w2[2], w3[2], w12[2], o12[2];
  for(int k = 0; k < 2; k++)
  {
    w0[k] = f[k] * (-0.5f + f[k] * (1.0f - 0.5f * f[k]));
    w1[k] = 1.0f + f[k] * f[k] * (-2.5f + 1.5f * f[k]);
    w2[k] = f[k] * (0.5f + f[k] * (2.0f - 1.5f * f[k]));
    w3[k] = f[k] * f[k] * (-0.5f + 0.5f * f[k]);
    w12[k] = w1[k] + w2[k];
    o12[k] = w2[k] / (w1[k] + w2[k]);
  }
  const float px[3] = { (t1x - 1.0f) / sx, (t1x + o12[0]) / sx, (t1x + 2.0f) / sx };
  const float py[3] = { (t1y - 1.0f) / sy, (t1y + o12[1]) / sy, (t1y + 2.0f) / sy };
  const float wx[3] = { w0[0], w12[0], w3[0] };
  const float wy[3] = { w0[1], w12[1], w3[1] };
  res[0] = res[1] = res[2] = res[3] = 0.0f;
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++)
  {
    float t[4];
    o_tex4(tex, px[i], py[j], t);
    for(int k = 0; k < 4; k++) res[k] += t[k] * wx[i] * wy[j];
  }
}
void o_sample_catmull_rom(const oimg_t *tex, float u, float v, float *res) { catmull_rom(tex, u, v, res); }

/* crop/main.comp:20-57 */
void o_crop_main(const oimg_t *in, oimg_t *out, const float *f)
{
  const float tsx = (float)in->w, tsy = (float)in->h;
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float xx = (float)x + 0.5f, yy = (float)y + 0.5f;
    xx += f[16] * tsx; yy += f[18] * tsy;
    /* T = mat2(r0, r1, r2, r3) columns (r0,r1),(r2,r3) */
    const float dx = xx - tsx * .5f, dy = yy - tsy * .5f;
    xx = f[12] * dx + f[14] * dy + tsx * .5f;
    yy = f[13] * dx + f[15] * dy + tsy * .5f;
    /* H columns f[0..2], f[4..6], f[8..10] */
    const float hx = f[0] * xx + f[4] * yy + f[8];
    const float hy = f[1] * xx + f[5] * yy + f[9];
    const float hz = f[2] * xx + f[6] * yy + f[10];
    float rdx = hx / hz, rdy = hy / hz;
    rdx /= tsx; rdy /= tsy;
    float rgba[4];
    if(rdx < 0.f || rdy < 0.f || rdx >= 1.f || rdy >= 1.f) rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0.0f;
    else if(f[12] != 1.0f) catmull_rom(in, rdx, rdy, rgba);
    else o_fetch4(in, (int)(rdx * tsx), (int)(rdy * tsy), rgba);
    rgba[3] = 1.0f;
    o_store4(out, x, y, rgba, 1);
  }
}

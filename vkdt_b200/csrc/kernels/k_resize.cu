// the export side's resize module (src/pipe/modules/resize, what vkdt-cli --width / --height inserts in front of the sink):
//   (resize, main)   resize/main.comp:22-35: nearest slice / sample_flower (magnify) / sample_catmull_rom (minify)
//   (shared, blurh), (shared, blurv)   shared/blurh.comp, blurv.comp:7-66: the separable gaussian in front of a slice beyond 3x
// per pixel kernels: an export runs them once per frame on the OUTPUT size, they are nowhere near the frame's cost.
// texture coordinates are carried in double like the restatement's ideal sampler (oracle/o_common.h o_tex4) in both builds.
#include "pointwise.cuh"

VKB_DEV float4 tex_rgba_d(const uint2 *__restrict__ img, int w, int h, double u, double v)
{
  double x = u * (double)w - 0.5, y = v * (double)h - 0.5;
  if(fabs(x - rint(x)) < 1.0 / 4096.0) x = rint(x);
  if(fabs(y - rint(y)) < 1.0 / 4096.0) y = rint(y);
  const double fx = floor(x), fy = floor(y);
  return bilin_rgba(img, w, h, (int)fx, (int)fy, (float)(x - fx), (float)(y - fy));
}
VKB_DEV void st_px(void *__restrict__ out, int f32, int ow, int x, int y, float4 v)
{
  if(f32) reinterpret_cast<float4 *>(out)[(size_t)y * ow + x] = v;
  else st_rgba(reinterpret_cast<uint2 *>(out), ow, x, y, v);
}

// shared.glsl:199-221
VKB_DEV float4 sample_flower(const uint2 *__restrict__ tex, int w, int h, double tcx, double tcy)
{
  const double sx = (double)w, sy = (double)h;
  const float t = 36.0f / 256.0f, wq = (1.0f - t) / 4.0f;
  const double ox[5] = { 0.0, (double)1.2f, -(double)1.2f, -(double)0.4f, (double)0.4f };
  const double oy[5] = { 0.0, (double)0.4f, -(double)0.4f, (double)1.2f, -(double)1.2f };
  float4 res = make_float4(0, 0, 0, 0);
#pragma unroll
  for(int k = 0; k < 5; k++)
  {
    const float4 v = tex_rgba_d(tex, w, h, (tcx + ox[k]) / sx, (tcy + oy[k]) / sy);
    const float W = k ? wq : t;
    const float l = fmaxf(v.x, fmaxf(v.y, v.z));
    res.x += W * v.x; res.y += W * v.y; res.z += W * v.z; res.w += W * l * l;
  }
  return res;
}

__global__ void __launch_bounds__(256) k_resize_main(const uint2 *__restrict__ in, int iw, int ih, void *__restrict__ out, int ow, int oh, int f32, int mode)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  float4 rgb;
  if(mode == 1)
  {
    const int fx = (int)((float)iw * ((float)x + 0.5f) / (float)ow), fy = (int)((float)ih * ((float)y + 0.5f) / (float)oh);
    rgb = ld_rgba_clamp(in, iw, ih, fx, fy);
  }
  else if(mode < 1) rgb = sample_flower(in, iw, ih, ((double)x + 0.5) / (double)ow * (double)iw, ((double)y + 0.5) / (double)oh * (double)ih);
  else rgb = catmull_rom_rgba(in, iw, ih, ((float)x + 0.5f) / (float)ow, ((float)y + 0.5f) / (float)oh);
  st_px(out, f32, ow, x, y, rgb);
}

template <bool VERT>
__global__ void __launch_bounds__(256) k_blur_sep(const uint2 *__restrict__ in, int w, int h, void *__restrict__ out, int f32, float radius, float c0, float w0, float v0)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= w || y >= h) return;
  const int sp = (int)floorf(radius);
  const double u = ((double)x + 0.5) / (double)w, vv = ((double)y + 0.5) / (double)h;
  float4 color = tex_rgba_d(in, w, h, u, vv);
  float wgt = 1.0f;
  if(sp > 0)
  {
    float wi = w0, v = v0;
    int i = 1;
    for(; i <= sp - 1; i += 2)
    {
      const double d1 = (double)i, d2 = (double)(i + 1);
      // (d / size as one double quotient per tap, like the restatement)
      const float4 p1 = tex_rgba_d(in, w, h, VERT ? u : u + d1 / (double)w, VERT ? vv + d1 / (double)h : vv);
      const float4 m1 = tex_rgba_d(in, w, h, VERT ? u : u - d1 / (double)w, VERT ? vv - d1 / (double)h : vv);
      const float4 p2 = tex_rgba_d(in, w, h, VERT ? u : u + d2 / (double)w, VERT ? vv + d2 / (double)h : vv);
      const float4 m2 = tex_rgba_d(in, w, h, VERT ? u : u - d2 / (double)w, VERT ? vv - d2 / (double)h : vv);
      const float w2 = wi * v;
      const float vn = v * c0;
      wgt += 2.0f * (wi + w2);
      color.x += wi * (p1.x + m1.x) + w2 * (p2.x + m2.x);
      color.y += wi * (p1.y + m1.y) + w2 * (p2.y + m2.y);
      color.z += wi * (p1.z + m1.z) + w2 * (p2.z + m2.z);
      color.w += wi * (p1.w + m1.w) + w2 * (p2.w + m2.w);
      wi = w2 * vn;
      v = vn * c0;
    }
    if(i == sp)
    {
      const double d1 = (double)i;
      const float4 p1 = tex_rgba_d(in, w, h, VERT ? u : u + d1 / (double)w, VERT ? vv + d1 / (double)h : vv);
      const float4 m1 = tex_rgba_d(in, w, h, VERT ? u : u - d1 / (double)w, VERT ? vv - d1 / (double)h : vv);
      wgt += 2.0f * wi;
      color.x += wi * (p1.x + m1.x); color.y += wi * (p1.y + m1.y); color.z += wi * (p1.z + m1.z); color.w += wi * (p1.w + m1.w);
    }
  }
  const float iw_ = 1.0f / wgt;
  st_px(out, f32, w, x, y, make_float4(color.x * iw_, color.y * iw_, color.z * iw_, color.w * iw_));
}

static inline dim3 grid2d(unsigned w, unsigned h) { return dim3(vkb_cdiv(w, 32), vkb_cdiv(h, 8)); }
static const dim3 blk2d(32, 8);

// conn: [0] input rgba f16, [1] output rgba f16 | f32.  push: { i32 mode }
static int launch_resize(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= 4 && l->band_y0 < 0);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16 && out->chan == 4 && (out->format == VKB_TOKEN_F16 || out->format == VKB_TOKEN_F32));
  if(!out->wd || !out->ht) return VKB_OK;
  k_resize_main<<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->wd, out->ht,
      out->format == VKB_TOKEN_F32, *(const int32_t *)l->push);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("resize", "main", launch_resize);

// conn: [0] input rgba f16, [1] output of the same size.  push: { f32 radius = 3 sigma }
template <bool VERT>
static int launch_blur(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= 4 && l->band_y0 < 0);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16 && out->chan == 4 && in->wd == out->wd && in->ht == out->ht && in->layers <= 1);
  VKB_REQUIRE(out->format == VKB_TOKEN_F16 || out->format == VKB_TOKEN_F32);
  if(!out->wd || !out->ht) return VKB_OK;
  const float radius = *(const float *)l->push;
  // the three exponentials of blurh.comp:22-28 are functions of the push constant: libm's, on the host, like the restatement's
  const volatile float sigma = radius / 3.0f;
  const volatile float ss = sigma * sigma;
  const volatile float a = 0.5f / ss;
  const volatile float am2 = -2.0f * a, am1 = -a, am3 = -3.0f * a;
  k_blur_sep<VERT><<<grid2d(out->wd, out->ht), blk2d, 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht, out->data, out->format == VKB_TOKEN_F32,
      radius, expf(am2), expf(am1), expf(am3));
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
static int launch_blurh(const vkb_launch_t *l) { return launch_blur<false>(l); }
static int launch_blurv(const vkb_launch_t *l) { return launch_blur<true>(l); }
VKB_REGISTER("shared", "blurh", launch_blurh);
VKB_REGISTER("shared", "blurv", launch_blurv);

VKB_NS_END

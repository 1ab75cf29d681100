// half-size demosaic (demosaic:method 2 or strong downscale) and the generic catmull-rom resampler.
// replaces src/pipe/modules/demosaic/halfsize.comp:15-52 and src/pipe/modules/shared/resample.comp:26-40
// (the shader returns right after sample_catmull_rom(); the quadric fit below it is dead code).
#include "pointwise.cuh"

struct demosaic_push_t { float wb[4]; uint32_t filters; };

__global__ void __launch_bounds__(256) k_demosaic_halfsize(const __half *__restrict__ in, int iw, int ih,
    uint2 *__restrict__ out, int ow, int oh, int xtrans)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  float4 rgba;
  if(xtrans)
  {
    float c[9];
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
      for(int j = 0; j < 3; j++) c[3 * i + j] = ld_h_clamp(in, iw, ih, 3 * x + i, 3 * y + j);
    rgba.y = (c[0] + c[2] + c[4] + c[6] + c[8]) * 1.0f / 5.0f;
    const float col0 = (c[1] + c[7]) * 0.5f, col1 = (c[3] + c[5]) * .5f;
    if(((x + y) & 1) > 0) { rgba.x = col0; rgba.z = col1; }
    else                  { rgba.z = col0; rgba.x = col1; }
    rgba.w = 1.0f; // the shader leaves alpha undefined on this branch
  }
  else
  { // textureGather at the block centre: x=(0,1) y=(1,1) z=(1,0) w=(0,0), mirrored repeat
    const int x0 = mirrori(2 * x, iw), x1 = mirrori(2 * x + 1, iw), y0 = mirrori(2 * y, ih), y1 = mirrori(2 * y + 1, ih);
    const float cx = ld_h(in, iw, x0, y1), cy = ld_h(in, iw, x1, y1), cz = ld_h(in, iw, x1, y0), cw = ld_h(in, iw, x0, y0);
    rgba = make_float4(cw, (cx + cz) / 2.0f, cy, 1.0f);
  }
  st_rgba(out, ow, x, y, rgba);
}

__global__ void __launch_bounds__(256) k_resample(const uint2 *__restrict__ in, int iw, int ih, uint2 *__restrict__ out, int ow, int oh)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  st_rgba(out, ow, x, y, catmull_rom_rgba(in, iw, ih, ((float)x + 0.5f) / (float)ow, ((float)y + 0.5f) / (float)oh));
}

// conn: [0] input mosaic f16, [1] output rgba f16 (1/block size)
static int launch_halfsize(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= sizeof(demosaic_push_t));
  const demosaic_push_t *pc = (const demosaic_push_t *)l->push;
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 1 && in->format == VKB_TOKEN_F16 && out->chan == 4 && out->format == VKB_TOKEN_F16);
  k_demosaic_halfsize<<<dim3(vkb_cdiv(out->wd, 32), vkb_cdiv(out->ht, 8)), dim3(32, 8), 0, l->stream>>>((const __half *)in->data, in->wd, in->ht,
      (uint2 *)out->data, out->wd, out->ht, pc->filters == 9);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("demosaic", "halfsize", launch_halfsize);

// conn: [0] input rgba f16, [1] output rgba f16 (any size)
static int launch_resample(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F16 && out->chan == 4 && out->format == VKB_TOKEN_F16 && in->layers == 1);
  k_resample<<<dim3(vkb_cdiv(out->wd, 32), vkb_cdiv(out->ht, 8)), dim3(32, 8), 0, l->stream>>>((const uint2 *)in->data, in->wd, in->ht,
      (uint2 *)out->data, out->wd, out->ht);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("shared", "resample", launch_resample);

VKB_NS_END

// raw front end: MLV bit unpack (integer, bit exact) and denoise/noop (black/white normalisation + crop).
//  - (i-mlv, unpack)   replaces the CPU loop of src/pipe/modules/i-mlv/video_mlv.c:261-273
//  - (denoise, noop)   replaces src/pipe/modules/denoise/noop.comp:36-57
//  - (b200, rawnoop)   both fused: packed stream -> normalised f16 mosaic, 1.75 B in + 2 B out per pixel
// HBM-bound: every pixel is read once with 16-byte vectors staged through shared memory and written once.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// pixel i occupies stream bits [bpp*i, bpp*i+bpp), stream bit k = bit 15-(k%16) of little endian word k/16.
// pix = (((w[a] << 16) | w[a+1]) >> (32 - bpp - s)) & mask,  a = bpp*i/16, s = bpp*i%16   
// a block of 256 threads handles 2048 pixels = 128*bpp words = 16*bpp uint4.
template <int BPP, bool NORMALISE>
__global__ void __launch_bounds__(256) k_unpack(const uint16_t *__restrict__ in, void *__restrict__ outv,
    uint64_t npix, uint64_t nwords, float black, float white)
{
  __shared__ __align__(16) uint16_t s[128 * BPP + 8];
  const uint64_t blk_word0 = (uint64_t)blockIdx.x * (128 * BPP);
  const uint4 *in4 = reinterpret_cast<const uint4 *>(in + blk_word0);
  for(int i = threadIdx.x; i < 16 * BPP + 1; i += 256)
  {
    const uint64_t w0 = blk_word0 + 8ull * i;
    if(w0 + 8 <= nwords && i < 16 * BPP) reinterpret_cast<uint4 *>(s)[i] = __ldg(in4 + i);
    else for(int k = 0; k < 8; k++) s[8 * i + k] = (w0 + k < nwords) ? __ldg(in + w0 + k) : (uint16_t)0;
  }
  __syncthreads();
  const uint64_t p0 = (uint64_t)blockIdx.x * 2048 + 8ull * threadIdx.x;
  if(p0 >= npix) return;
  const uint16_t *w = s + threadIdx.x * (BPP / 2);
  uint16_t px[8];
#pragma unroll
  for(int k = 0; k < 8; k++)
  {
    const int bits = k * BPP, a = bits >> 4, sh = bits & 15;
    const uint32_t v = ((uint32_t)w[a] << 16) | w[a + 1];
    px[k] = (uint16_t)((v >> (32 - BPP - sh)) & ((1u << BPP) - 1u));
  }
  if(!NORMALISE)
  {
    uint16_t *out = reinterpret_cast<uint16_t *>(outv);
    if(p0 + 8 <= npix)
    {
      uint4 o;
      o.x = px[0] | ((uint32_t)px[1] << 16); o.y = px[2] | ((uint32_t)px[3] << 16);
      o.z = px[4] | ((uint32_t)px[5] << 16); o.w = px[6] | ((uint32_t)px[7] << 16);
      *reinterpret_cast<uint4 *>(out + p0) = o;
    }
    else for(int k = 0; k < 8 && p0 + k < npix; k++) out[p0 + k] = px[k];
  }
  else
  {
    __half *out = reinterpret_cast<__half *>(outv);
    __half h[8];
#pragma unroll
    for(int k = 0; k < 8; k++)
    { // ui16 sampled as UNORM, then denoise/noop.comp:43-44
      const float c = (float)px[k] / 65535.0f;
      h[k] = __float2half_rn(fmaxf(0.0f, (c - black) / (white - black)));
    }
    if(p0 + 8 <= npix) *reinterpret_cast<uint4 *>(out + p0) = *reinterpret_cast<uint4 *>(h);
    else for(int k = 0; k < 8 && p0 + k < npix; k++) out[p0 + k] = h[k];
  }
}

template <bool NORM>
static int launch_unpack_t(const vkb_launch_t *l, int bpp, float black, float white)
{
  const vkb_image_t *in = l->conn + 0, *out = l->conn + 1;
  const uint64_t npix = (uint64_t)out->wd * out->ht;
  const uint64_t nwords = (npix * bpp + 15) / 16;
  VKB_REQUIRE(((uintptr_t)in->data & 15) == 0 && ((uintptr_t)out->data & 15) == 0);
  const unsigned grid = (unsigned)((npix + 2047) / 2048);
  if(!grid) return VKB_OK;
  const uint16_t *ip = (const uint16_t *)in->data;
  switch(bpp)
  {
    case 10: k_unpack<10, NORM><<<grid, 256, 0, l->stream>>>(ip, out->data, npix, nwords, black, white); break;
    case 12: k_unpack<12, NORM><<<grid, 256, 0, l->stream>>>(ip, out->data, npix, nwords, black, white); break;
    case 14: k_unpack<14, NORM><<<grid, 256, 0, l->stream>>>(ip, out->data, npix, nwords, black, white); break;
    default: return vkb_set_error(VKB_ERR_BAD_ARG, "mlv unpack: unsupported bits per pixel %d", bpp);
  }
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}

// conn: [0] packed words (ui16, any wd/ht), [1] output ui16 wd x ht.  push: { int bpp }
static int launch_mlv_unpack(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= 4);
  return launch_unpack_t<false>(l, ((const int32_t *)l->push)[0], 0.0f, 1.0f);
}
VKB_REGISTER("i-mlv", "unpack", launch_mlv_unpack);

// ---------------------------------------------------------------------------------------------------------
// denoise/noop.comp push block (denoise/main.c:208-214): ivec4 crop; vec4 black; vec4 white; vec4 map_os; int filters; int gainmap
struct noop_push_t { int32_t crop[4]; float black[4]; float white[4]; float map_os[4]; int32_t filters, gainmap; };

// the reference stores (v,0,0,1) into an rgba f16 image that every consumer reads as .r; we store .r only.
__global__ void __launch_bounds__(256) k_denoise_noop(const uint16_t *__restrict__ in, int iw, int ih,
    __half *__restrict__ out, int ow, int oh, int cx, int cy, float black, float white)
{
  const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8, y = blockIdx.y * 8 + threadIdx.y;
  if(x0 >= ow || y >= oh) return;
  const int sy = clampi(y + cy, 0, ih - 1);
  const uint16_t *row = in + (size_t)sy * iw;
  __half *orow = out + (size_t)y * ow;
  const float rng = white - black;
  if(x0 + 8 <= ow && x0 + cx + 8 <= iw && (((uintptr_t)(row + x0 + cx)) & 15) == 0 && (((uintptr_t)(orow + x0)) & 15) == 0)
  {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + x0 + cx));
    const uint32_t u[4] = { v.x, v.y, v.z, v.w };
    __half h[8];
#pragma unroll
    for(int k = 0; k < 4; k++)
    {
      const float a = (float)(u[k] & 0xffffu) / 65535.0f, b = (float)(u[k] >> 16) / 65535.0f;
      h[2 * k]     = __float2half_rn(fmaxf(0.0f, (a - black) / rng));
      h[2 * k + 1] = __float2half_rn(fmaxf(0.0f, (b - black) / rng));
    }
    *reinterpret_cast<uint4 *>(orow + x0) = *reinterpret_cast<uint4 *>(h);
  }
  else
  {
    for(int k = 0; k < 8 && x0 + k < ow; k++)
    {
      const float a = (float)__ldg(row + clampi(x0 + k + cx, 0, iw - 1)) / 65535.0f;
      orow[x0 + k] = __float2half_rn(fmaxf(0.0f, (a - black) / rng));
    }
  }
}

// the same with a DNG gain map (noop.comp:48-57): one pixel per thread
__global__ void __launch_bounds__(256) k_denoise_noop_gm(const uint16_t *__restrict__ in, int iw, int ih,
    __half *__restrict__ out, int ow, int oh, int cx, int cy, float black, float white, const gainmap_t G)
{
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if(x >= ow || y >= oh) return;
  const float a = (float)__ldg(in + (size_t)clampi(y + cy, 0, ih - 1) * iw + clampi(x + cx, 0, iw - 1)) / 65535.0f;
  float col = fmaxf(0.0f, (a - black) / (white - black));
  col *= gainmap_gain(G, x, y, cx, cy, iw, ih, 1);
  out[(size_t)y * ow + x] = __float2half_rn(col);
}

// conn: [0] input ui16 1ch, [1] output f16 1ch, [2] gain map rgba f32 (a dummy binding unless push.gainmap, denoise/main.c:216-221)
// params: the denoise parameter block (its last int switches the gain map off)
static int launch_denoise_noop(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= sizeof(noop_push_t));
  const noop_push_t *p = (const noop_push_t *)l->push;
  const vkb_image_t *in = l->conn + 0, *out = l->conn + 1;
  VKB_REQUIRE(in->format == VKB_TOKEN_UI16 && out->format == VKB_TOKEN_F16 && in->chan == 1 && out->chan == 1);
  const int par_gainmap = l->params_size >= 36 ? ((const int32_t *)l->params)[8] : 1;   // strength luma detail pad edges[4] gainmap
  if(p->filters != 0 && p->filters != 9 && p->gainmap == 1 && par_gainmap == 1)
  {
    VKB_REQUIRE(l->num_conn >= 3 && l->conn[2].format == VKB_TOKEN_F32 && l->conn[2].chan == 4 && l->conn[2].data && l->band_y0 < 0);
    gainmap_t G = { (const float4 *)l->conn[2].data, (int)l->conn[2].wd, (int)l->conn[2].ht, { p->map_os[0], p->map_os[1], p->map_os[2], p->map_os[3] } };
    k_denoise_noop_gm<<<dim3(vkb_cdiv(out->wd, 32), vkb_cdiv(out->ht, 8)), dim3(32, 8), 0, l->stream>>>((const uint16_t *)in->data, in->wd, in->ht,
        (__half *)out->data, out->wd, out->ht, p->crop[0], p->crop[1], p->black[0], p->white[0], G);
    VKB_CHECK_LAUNCH();
    return VKB_OK;
  }
  dim3 block(32, 8), grid(vkb_cdiv(out->wd, 256), vkb_cdiv(out->ht, 8));
  k_denoise_noop<<<grid, block, 0, l->stream>>>((const uint16_t *)in->data, in->wd, in->ht,
      (__half *)out->data, out->wd, out->ht, p->crop[0], p->crop[1], p->black[0], p->white[0]);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("denoise", "noop", launch_denoise_noop);

// fused unpack + noop for uncropped frames.  conn: [0] packed words, [1] output f16 wd x ht.
// push: { int bpp; float black; float white }  (black/white already divided by 65535 like denoise/main.c:165-168)
static int launch_rawnoop(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2 && l->push_size >= 12);
  const int32_t *pi = (const int32_t *)l->push;
  const float *pf = (const float *)l->push;
  VKB_REQUIRE(l->conn[1].format == VKB_TOKEN_F16 && l->conn[1].chan == 1);
  return launch_unpack_t<true>(l, pi[0], pf[1], pf[2]);
}
VKB_REGISTER("b200", "rawnoop", launch_rawnoop);

VKB_NS_END

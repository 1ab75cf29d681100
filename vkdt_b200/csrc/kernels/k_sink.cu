// sink side helpers of the executor.
// (b200, pfmpack): rgba f32 -> packed rgb f32, the PFM payload (o-pfm/main.c:36-40).  only launched when the producer of
// the sink image is not one of the fused kernels that store r g b themselves (k_llap_fin.cu, k_pointwise.cu).
#include "common.cuh"

__global__ void __launch_bounds__(256) k_pfmpack(const float4 *__restrict__ in, float *__restrict__ out, size_t n)
{
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(i >= n) return;
  const float4 v = __ldg(in + i);
  out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
}

// conn: [0] rgba f32, [1] rgb f32 (chan 3)
static int launch_pfmpack(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->chan == 4 && in->format == VKB_TOKEN_F32 && out->chan == 3 && out->format == VKB_TOKEN_F32);
  VKB_REQUIRE(in->wd == out->wd && in->ht == out->ht);
  const size_t n = (size_t)in->wd * in->ht;
  k_pfmpack<<<(unsigned)((n + 255) / 256), 256, 0, l->stream>>>((const float4 *)in->data, (float *)out->data, n);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("b200", "pfmpack", launch_pfmpack);

// (b200, cvt16): f32 -> f16 element by element, behind an f32 source (i-pfm) whose consumers read f16 edges
__global__ void __launch_bounds__(256) k_cvt16(const float2 *__restrict__ in, __half2 *__restrict__ out, size_t n2, const float *__restrict__ in1, __half *__restrict__ out1, size_t n)
{
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(i < n2) { const float2 v = __ldg(in + i); out[i] = __floats2half2_rn(v.x, v.y); }
  if(i == 0 && (n & 1)) out1[n - 1] = __float2half_rn(in1[n - 1]);
}

// conn: [0] f32 image, [1] f16 image of the same shape
static int launch_cvt16(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 2);
  const vkb_image_t *in = l->conn, *out = l->conn + 1;
  VKB_REQUIRE(in->format == VKB_TOKEN_F32 && out->format == VKB_TOKEN_F16 && in->wd == out->wd && in->ht == out->ht && in->chan == out->chan);
  const size_t n = (size_t)in->wd * in->ht * in->chan;
  if(!n) return VKB_OK;
  k_cvt16<<<(unsigned)((n / 2 + 256) / 256), 256, 0, l->stream>>>((const float2 *)in->data, (__half2 *)out->data, n / 2, (const float *)in->data, (__half *)out->data, n);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("b200", "cvt16", launch_cvt16);

// (b200, libm): the mode's transcendental functions element by element, for the tests that compare them with libm
// (tests/test_libm_exact_gpu.py).  push: { u32 op }: 0 exp, 1 exp2, 2 log2, 3 pow(a, b).  conn: [0] a f32, [1] b f32, [2] out f32
__global__ void __launch_bounds__(256) k_libm(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, size_t n, int op)
{
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(i >= n) return;
  const float x = a[i], y = b[i];
  out[i] = op == 0 ? m_exp(x) : (op == 1 ? m_exp2(x) : (op == 2 ? m_log2(x) : m_pow(x, y)));
}
static int launch_libm(const vkb_launch_t *l)
{
  VKB_REQUIRE(l->num_conn >= 3 && l->push_size >= 4);
  const vkb_image_t *a = l->conn, *b = l->conn + 1, *out = l->conn + 2;
  VKB_REQUIRE(a->format == VKB_TOKEN_F32 && b->format == VKB_TOKEN_F32 && out->format == VKB_TOKEN_F32);
  const size_t n = (size_t)out->wd * out->ht * out->chan;
  if(!n) return VKB_OK;
  k_libm<<<(unsigned)((n + 255) / 256), 256, 0, l->stream>>>((const float *)a->data, (const float *)b->data, (float *)out->data, n, *(const int *)l->push);
  VKB_CHECK_LAUNCH();
  return VKB_OK;
}
VKB_REGISTER("b200", "libm", launch_libm);

VKB_NS_END

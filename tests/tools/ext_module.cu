// a module OUTSIDE the library, as a third party would write it against include/vkdt_b200.h: "invert" (rgba f16 -> rgba f16,
// out = amount - in) with one kernel for its default node.  built by tests/test_external_module_gpu.py with nvcc.
#include "../../include/vkdt_b200.h"
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__global__ void k_invert(const __half *in, __half *out, size_t n, float amount)
{
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(i >= n) return;
  out[i] = (i & 3) == 3 ? __float2half(1.0f) : __float2half_rn(amount - __half2float(in[i]));
}
extern "C" int invert_main(const vkb_kernel_args_t *a)
{
  if(a->num_conn < 2 || a->params_size < 4) return VKB_ERR_BAD_ARG;
  const vkb_image_t *in = a->conn, *out = a->conn + 1;
  const size_t n = (size_t)out->wd * out->ht * 4;
  k_invert<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)a->stream>>>((const __half *)in->data, (__half *)out->data, n, *(const float *)a->params);
  return cudaGetLastError() == cudaSuccess ? 0 : VKB_ERR_CUDA;
}
extern "C" const char *invert_connectors = "input:read:rgba:f16\noutput:write:rgba:f16\n";
extern "C" const char *invert_params = "amount:float:1:1.0\n";

// a second outside module with a FEEDBACK input: "iir" (out = (1 - keep) * input + keep * back, back = its own output of the
// frame before, wired with a `feedback:` line).  exercises the double buffered connectors of the executor.
__global__ void k_iir(const __half *in, const __half *back, __half *out, size_t n, float keep)
{
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(i >= n) return;
  out[i] = (i & 3) == 3 ? __float2half(1.0f) : __float2half_rn(__fadd_rn(__fmul_rn(1.0f - keep, __half2float(in[i])), __fmul_rn(keep, __half2float(back[i]))));
}
extern "C" int iir_main(const vkb_kernel_args_t *a)
{
  if(a->num_conn < 3 || a->params_size < 4) return VKB_ERR_BAD_ARG;
  const vkb_image_t *in = a->conn, *back = a->conn + 1, *out = a->conn + 2;
  if(back->wd != out->wd || back->ht != out->ht || back->data == out->data) return VKB_ERR_BAD_ARG;
  const size_t n = (size_t)out->wd * out->ht * 4;
  k_iir<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)a->stream>>>((const __half *)in->data, (const __half *)back->data, (__half *)out->data, n, *(const float *)a->params);
  return cudaGetLastError() == cudaSuccess ? 0 : VKB_ERR_CUDA;
}
extern "C" const char *iir_connectors = "input:read:rgba:f16\nback:read:rgba:f16\noutput:write:rgba:f16\n";
extern "C" const char *iir_params = "keep:float:1:0.5\n";

// internal glue between the C-ABI (include/vkdt_b200.h), the kernel registry and the graph executor.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <cuda_runtime.h>
#include "../../include/vkdt_b200.h"

// one node dispatch as the kernels see it (mirrors what record_command_buffer binds per node,
// src/pipe/graph-run-nodes-record-cmd.h:333-416)
struct vkb_launch_t
{
  uint32_t wd, ht, dp;
  const void *push;   uint32_t push_size;
  const void *params; uint32_t params_size;
  const vkb_image_t *conn; uint32_t num_conn;
  cudaStream_t stream;
  // band split (executor.cpp): only rows [band_y0, band_y1) of the launcher's band image (the image
  // its grid covers) are computed; every coordinate, mirror rule and size stays that of the whole image.  -1: everything.
  int32_t band_y0, band_y1;
};
typedef int (*vkb_kernel_fn)(const vkb_launch_t *);

// kernels exist in two builds, see kernels/common.cuh: mode 0 strict (default), 1 fast
#ifndef VKB_FAST
#define VKB_FAST 0
#endif
#if VKB_FAST
#define VKB_NS_BEGIN namespace vkb_fast {
#else
#define VKB_NS_BEGIN namespace vkb_strict {
#endif
#define VKB_NS_END }
void vkb_register_kernel(const char *name, const char *kernel, vkb_kernel_fn fn, int mode);
vkb_kernel_fn vkb_find_kernel(vkb_token_t name, vkb_token_t kernel, int mode = -1); // mode < 0: the process default (vkb_set_mode)
int vkb_default_mode(void);
// what vkb_dispatch does, with an explicit mode and band (the graph executor's entry point)
int vkb_dispatch_launch(vkb_token_t name, vkb_token_t kernel, int mode, const vkb_launch_t *l);
int  vkb_set_error(int code, const char *fmt, ...);
void vkb_count_launch(int n);

#define VKB_TOKEN_F16  0x363166ull        /* "f16"  */
#define VKB_TOKEN_F32  0x323366ull        /* "f32"  */
#define VKB_TOKEN_UI16 0x36316975ull      /* "ui16" */
#define VKB_TOKEN_UI8  0x386975ull        /* "ui8"  */

struct vkb_registrar_t { vkb_registrar_t(const char *n, const char *k, vkb_kernel_fn f) { vkb_register_kernel(n, k, f, VKB_FAST); } };
#define VKB_REGISTER(name, kernel, fn) static vkb_registrar_t vkb_reg_##fn(name, kernel, fn)

#define VKB_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); \
  if(e_ != cudaSuccess) return vkb_set_error(VKB_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
  vkb_count_launch(1); } while(0)
#define VKB_REQUIRE(cond) do { if(!(cond)) return vkb_set_error(VKB_ERR_BAD_ARG, "%s:%d: requirement failed: %s", __FILE__, __LINE__, #cond); } while(0)

static inline unsigned vkb_cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

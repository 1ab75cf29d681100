// runtime + dispatch layers of the C-ABI (include/vkdt_b200.h §1, §2).
// replaces src/qvk (device selection, queues) and the (name, kernel) -> pipeline lookup of src/pipe/graph.c:304-343
// with a static registry of CUDA launchers.  there is no CPU fallback: without a device every call fails.
#include "vkb_internal.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>

#define VKB_MAX_KERNELS 256
struct kernel_entry_t { vkb_token_t name, kernel; vkb_kernel_fn fn; int mode; };
static kernel_entry_t *g_kernels() { static kernel_entry_t k[VKB_MAX_KERNELS]; return k; }
static int &g_kernel_cnt() { static int n = 0; return n; }
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static int g_device = -1;

extern "C" vkb_token_t vkb_token(const char *str)
{ // src/pipe/token.h:39-56: up to 8 chars, little endian, zero padded
  vkb_token_t t = 0;
  for(int i = 0; i < 8 && str && str[i]; i++) t |= (vkb_token_t)(uint8_t)str[i] << (8 * i);
  return t;
}

static int &g_mode()
{ // process default: strict, unless the environment asks for the fast build of the kernels
  static int m = (getenv("VKB_FAST") && atoi(getenv("VKB_FAST")) != 0) ? VKB_MODE_FAST : VKB_MODE_STRICT;
  return m;
}
int vkb_default_mode(void) { return g_mode(); }
void vkb_register_kernel(const char *name, const char *kernel, vkb_kernel_fn fn, int mode)
{
  if(g_kernel_cnt() >= VKB_MAX_KERNELS) return;
  g_kernels()[g_kernel_cnt()++] = { vkb_token(name), vkb_token(kernel), fn, mode };
}
vkb_kernel_fn vkb_find_kernel(vkb_token_t name, vkb_token_t kernel, int mode)
{
  if(mode < 0) mode = g_mode();
  for(int i = 0; i < g_kernel_cnt(); i++)
    if(g_kernels()[i].name == name && g_kernels()[i].kernel == kernel && g_kernels()[i].mode == mode) return g_kernels()[i].fn;
  return 0;
}
int vkb_set_error(int code, const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void vkb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int vkb_dispatch_launch(vkb_token_t name, vkb_token_t kernel, int mode, const vkb_launch_t *l)
{
  if(g_device < 0) { const int r = vkb_init(0); if(r) return r; }
  vkb_kernel_fn fn = vkb_find_kernel(name, kernel, mode);
  if(!fn)
  {
    char a[9] = {0}, b[9] = {0};
    memcpy(a, &name, 8); memcpy(b, &kernel, 8);
    return vkb_set_error(VKB_ERR_UNKNOWN_KERNEL, "no kernel registered for (%s, %s)", a, b);
  }
  for(uint32_t i = 0; i < l->num_conn; i++)
    if(l->conn[i].data && l->conn[i].layers == 0) return vkb_set_error(VKB_ERR_BAD_ARG, "connector %u has zero layers", i);
  return fn(l);
}

extern "C" {

const char *vkb_last_error(void) { return g_err; }
const char *vkb_version(void) { return "vkdt_b200 0.1 (sm_100a)"; }

int vkb_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int vkb_init(int device_id)
{
  const int n = vkb_device_count();
  if(n <= 0) return vkb_set_error(VKB_ERR_NO_DEVICE, "no CUDA device: vkdt_b200 has no CPU fallback");
  if(device_id < 0) device_id = 0;
  if(device_id >= n) return vkb_set_error(VKB_ERR_BAD_ARG, "device %d out of range (%d devices)", device_id, n);
  cudaError_t e = cudaSetDevice(device_id);
  if(e != cudaSuccess) return vkb_set_error(VKB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaFree(0);
  g_device = device_id;
  return VKB_OK;
}
void vkb_cleanup(void) { g_device = -1; }

static int need_device(void)
{
  if(g_device >= 0) return VKB_OK;
  return vkb_init(0);
}

#define CU(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) return vkb_set_error(VKB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while(0)

int vkb_malloc(void **dptr, size_t bytes) { int r = need_device(); if(r) return r; CU(cudaMalloc(dptr, bytes)); return VKB_OK; }
int vkb_free(void *dptr) { CU(cudaFree(dptr)); return VKB_OK; }
int vkb_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{ CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream)); return VKB_OK; }
int vkb_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{ CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream)); return VKB_OK; }
int vkb_stream_sync(void *stream) { CU(cudaStreamSynchronize((cudaStream_t)stream)); return VKB_OK; }
int vkb_host_alloc(void **hptr, size_t bytes) { int r = need_device(); if(r) return r; CU(cudaHostAlloc(hptr, bytes, cudaHostAllocDefault)); return VKB_OK; }
int vkb_host_free(void *hptr) { CU(cudaFreeHost(hptr)); return VKB_OK; }

int vkb_dispatch(vkb_token_t name, vkb_token_t kernel, uint32_t wd, uint32_t ht, uint32_t dp,
                 const void *push, uint32_t push_size, const void *params, uint32_t params_size,
                 const vkb_image_t *conn, uint32_t num_conn, void *stream)
{
  vkb_launch_t l = { wd, ht, dp, push, push_size, params, params_size, conn, num_conn, (cudaStream_t)stream, -1, -1 };
  return vkb_dispatch_launch(name, kernel, -1, &l);
}

int vkb_event_create(void **ev) { int r = need_device(); if(r) return r; cudaEvent_t e; CU(cudaEventCreate(&e)); *ev = (void *)e; return VKB_OK; }
int vkb_event_record(void *ev, void *stream) { CU(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream)); return VKB_OK; }
int vkb_event_sync(void *ev) { CU(cudaEventSynchronize((cudaEvent_t)ev)); return VKB_OK; }
int vkb_event_elapsed_ms(void *ev0, void *ev1, float *ms) { CU(cudaEventElapsedTime(ms, (cudaEvent_t)ev0, (cudaEvent_t)ev1)); return VKB_OK; }
int vkb_event_destroy(void *ev) { CU(cudaEventDestroy((cudaEvent_t)ev)); return VKB_OK; }

// the strict set is what introspection lists (the fast set has the same names)
int vkb_kernel_count(void) { int n = 0; for(int i = 0; i < g_kernel_cnt(); i++) n += g_kernels()[i].mode == VKB_MODE_STRICT; return n; }
int vkb_kernel_name(int idx, vkb_token_t *name, vkb_token_t *kernel)
{
  for(int i = 0; i < g_kernel_cnt(); i++) if(g_kernels()[i].mode == VKB_MODE_STRICT && idx-- == 0)
  {
    *name = g_kernels()[i].name; *kernel = g_kernels()[i].kernel;
    return VKB_OK;
  }
  return VKB_ERR_BAD_ARG;
}
int vkb_set_mode(int mode)
{
  if(mode != VKB_MODE_STRICT && mode != VKB_MODE_FAST) return vkb_set_error(VKB_ERR_BAD_ARG, "mode %d: VKB_MODE_STRICT or VKB_MODE_FAST", mode);
  g_mode() = mode;
  return VKB_OK;
}
int vkb_get_mode(void) { return g_mode(); }
uint64_t vkb_launch_count(void) { return g_launches.load(); }
void vkb_launch_count_reset(void) { g_launches.store(0); }

} // extern "C"

// a caller's kernel (vkb_register_kernel) sees the launch through the public struct: same layout as the internal one
static_assert(sizeof(vkb_kernel_args_t) == sizeof(vkb_launch_t) && offsetof(vkb_kernel_args_t, stream) == offsetof(vkb_launch_t, stream) &&
              offsetof(vkb_kernel_args_t, band_y0) == offsetof(vkb_launch_t, band_y0) && offsetof(vkb_kernel_args_t, conn) == offsetof(vkb_launch_t, conn),
              "vkb_kernel_args_t and vkb_launch_t have to agree");

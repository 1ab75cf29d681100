/* vkdt_b200 — C-ABI of the B200-native raw-development engine.
 *
 * One shared library (libvkdt_b200.so), plain pointers and sizes only.  Three layers, each replacing a
 * named interface of the reference (hanatos/vkdt, paths relative to its tree):
 *
 *  1. runtime      vkb_init / vkb_cleanup / vkb_malloc ...     replaces src/qvk/qvk.h:96-146 (global `qvk`),
 *                                                              qvk_init src/qvk/qvk.c:138, qvk_cleanup
 *  2. dispatch     vkb_dispatch(name, kernel, push, params,    replaces the per-node body of record_command_buffer,
 *                  connectors)                                 src/pipe/graph-run-nodes-record-cmd.h:333-416: pipeline
 *                                                              lookup by (node->name, node->kernel) (src/pipe/graph.c:312-314),
 *                                                              push constants, params uniform, one binding per connector in
 *                                                              connector order, vkCmdDispatch
 *  3. graph        vkb_graph_*                                 replaces dt_graph_init/run/cleanup (src/pipe/graph.h:176-194),
 *                                                              dt_graph_read_config_ascii (src/pipe/graph-io.c:282),
 *                                                              dt_graph_read_config_line (graph-io.c:232),
 *                                                              dt_graph_export (src/pipe/graph-export.c:108)
 *
 * There is no CPU fallback: every compute entry point returns VKB_ERR_NO_DEVICE when no CUDA device is present.
 * All functions return 0 on success, a negative vkb_err_t otherwise (the reference returns VkResult, 0 = ok).
 * A graph is owned by one thread at a time (src/pipe/graph.h:66-69).
 */
#ifndef VKDT_B200_H
#define VKDT_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKB_API __attribute__((visibility("default")))

typedef enum vkb_err_t
{
  VKB_OK              =  0,
  VKB_ERR_NO_DEVICE   = -1,  /* no CUDA device / driver: the product path refuses to run */
  VKB_ERR_UNKNOWN_KERNEL = -2, /* no kernel registered for (name, kernel) */
  VKB_ERR_BAD_ARG     = -3,
  VKB_ERR_CUDA        = -4,  /* a CUDA call failed; see vkb_last_error() */
  VKB_ERR_IO          = -5,
  VKB_ERR_GRAPH       = -6,  /* graph incomplete (the reference's VK_INCOMPLETE, src/pipe/graph.c:745-793) */
  VKB_ERR_OOM         = -7,
} vkb_err_t;

/* ---- tokens: <= 8 chars packed little endian into a u64, identical to dt_token_t (src/pipe/token.h:15,39-56) ---- */
typedef uint64_t vkb_token_t;
VKB_API vkb_token_t vkb_token(const char *str);

/* ---- 1. runtime ---- */
VKB_API int  vkb_init(int device_id);            /* qvk_init(): pick device, create the pool + streams */
VKB_API void vkb_cleanup(void);                  /* qvk_cleanup() */
VKB_API int  vkb_device_count(void);
VKB_API const char *vkb_last_error(void);
VKB_API const char *vkb_version(void);
VKB_API int  vkb_malloc(void **dptr, size_t bytes);   /* plain device allocation helpers for callers without their own */
VKB_API int  vkb_free(void *dptr);
VKB_API int  vkb_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
VKB_API int  vkb_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
VKB_API int  vkb_stream_sync(void *stream);
VKB_API int  vkb_host_alloc(void **hptr, size_t bytes);  /* pinned host staging, the "mapped" pointer of read_source/write_sink */
VKB_API int  vkb_host_free(void *hptr);
/* device timestamps on a stream (the reference times nodes with vulkan timestamp queries, src/pipe/graph.c:96-109) */
VKB_API int  vkb_event_create(void **ev);
VKB_API int  vkb_event_record(void *ev, void *stream);
VKB_API int  vkb_event_sync(void *ev);
VKB_API int  vkb_event_elapsed_ms(void *ev0, void *ev1, float *ms);
VKB_API int  vkb_event_destroy(void *ev);

/* ---- 2. dispatch ---- */
/* one connector image as a kernel sees it (dt_connector_t + dt_connector_image_t, src/pipe/connector.h:156-210) */
typedef struct vkb_image_t
{
  void       *data;     /* device pointer; array layers are contiguous, layer stride = wd*ht*chan*sizeof(format) */
  uint32_t    wd, ht;   /* roi.wd, roi.ht */
  uint32_t    chan;     /* channels stored per texel: 1 or 4 (src/pipe/connector.h:281-291) */
  uint32_t    layers;   /* array_length, >= 1 */
  vkb_token_t format;   /* "ui16" | "f16" | "f32" */
} vkb_image_t;

/* launch the kernel registered for (name, kernel) on `stream` (cudaStream_t or 0).
 * wd/ht/dp: the node's dispatch extent; push: the node's push constant blob (same layout as the reference's
 * `pc[]` arrays, ints are bit-cast floats); params: the module's (committed) parameter blob;
 * conn[]: connector images in the node's connector order.  Asynchronous. */
VKB_API int vkb_dispatch(vkb_token_t name, vkb_token_t kernel, uint32_t wd, uint32_t ht, uint32_t dp,
                         const void *push, uint32_t push_size, const void *params, uint32_t params_size,
                         const vkb_image_t *conn, uint32_t num_conn, void *stream);
/* number of registered (name, kernel) pairs and their names, for introspection / tests */
VKB_API int vkb_kernel_count(void);
VKB_API int vkb_kernel_name(int idx, vkb_token_t *name, vkb_token_t *kernel);
/* kernels launched through vkb_dispatch / vkb_graph_run since the last reset (the bench's gpu_launches) */
/* arithmetic mode of the kernels.  every kernel exists in two builds:
 *   VKB_MODE_STRICT (default)  the arithmetic of the reference's shaders operation for operation: unfused multiply-adds, IEEE
 *                              division and square root, exp / exp2 / log2 / pow with libm's results bit for bit (what the CPU
 *                              oracle and the reference's shaders compiled as C++ compute)
 *   VKB_MODE_FAST              the SFU's ex2 / lg2 / rcp approximations (~2 ulp) and fused multiply-adds, what a GPU driver makes
 *                              of GLSL: within the path's tolerance (max abs 1e-3 on all but a 1e-5 tail), not bit compatible
 * vkb_set_mode sets the process default (also: environment VKB_FAST=1), used by vkb_dispatch and by graphs created after it. */
#define VKB_MODE_STRICT 0
#define VKB_MODE_FAST   1
VKB_API int  vkb_set_mode(int mode);
VKB_API int  vkb_get_mode(void);
VKB_API uint64_t vkb_launch_count(void);
VKB_API void     vkb_launch_count_reset(void);

/* ---- module registry (src/pipe/global.c:86-415, dt_pipe_global_init :442) ---- */
/* name a vkdt installation (<dir>/modules/<name>/connectors, params) or checkout (<vkdt>/src/pipe): its module files then are
 * the source of truth for connectors and parameters of every module, the tables built into the library only the fallback, and
 * every module directory found there can be named in a cfg (modules without kernels here parse, and fail only when a sink
 * reaches them).  also: environment VKDT_B200_BASEDIR.  call before creating graphs. */
VKB_API int  vkb_set_basedir(const char *dir);
/* the registered connectors and parameters of a module: "connector name:type:chan:format" and
 * "param name:type:cnt:offset:<default value blob, hex>" lines.  host only. */
VKB_API int  vkb_module_describe(const char *name, char *buf, size_t bufsize);

/* ---- module authors (src/pipe/modules/api.h, the dlopen'd lib<name>.so of src/pipe/global.c:106-119) ----
 * a module outside this library = the text of its `connectors` and `params` files + one kernel per node.  modules registered
 * this way get the reference's default callbacks (graph-run-modules.h:65-107, :289-322, :456-490): one node (name, "main")
 * with the module's connectors copied 1:1, dispatched over the `output` connector's size, output roi = input roi; the raw
 * parameter block is what the kernel receives as `params`.  (modules with their own create_nodes / commit_params / roi
 * callbacks are the ones compiled into this library; their C++ side mirrors api.h: dt_node_add, dt_connector_copy,
 * dt_connector_bypass, dt_node_connect in csrc/pipe/pipe.h.)
 * the kernel is a host function that launches CUDA work on args->stream and returns 0: the same contract as this library's
 * own launchers (csrc/vkb_internal.h), images in connector order, asynchronous.  band_y0 / band_y1 >= 0 only ever reach
 * kernels the band split knows (never a caller's).  mode: VKB_MODE_STRICT, VKB_MODE_FAST or -1 for both.
 * register before creating the graphs that use the module. */
typedef struct vkb_kernel_args_t
{
  uint32_t wd, ht, dp;                         /* the node's dispatch extent */
  const void *push;   uint32_t push_size;      /* push constants (none for default nodes) */
  const void *params; uint32_t params_size;    /* the module's parameter block */
  const vkb_image_t *conn; uint32_t num_conn;  /* connector images in connector order (device pointers) */
  void *stream;                                /* cudaStream_t */
  int32_t band_y0, band_y1;                    /* -1 */
} vkb_kernel_args_t;
typedef int (*vkb_kernel_fn_t)(const vkb_kernel_args_t *);
VKB_API int  vkb_register_module(const char *name, const char *connectors, const char *params);
VKB_API int  vkb_register_kernel(const char *name, const char *kernel, vkb_kernel_fn_t fn, int mode);

/* ---- 3. graph ---- */
typedef struct vkb_graph_t vkb_graph_t;

/* run flags, same bits as dt_graph_run_constants_t (src/pipe/modules/api.h:25-37) */
enum
{
  VKB_RUN_ROI            = 1 << 0,
  VKB_RUN_CREATE_NODES   = 1 << 1,
  VKB_RUN_ALLOC          = 1 << 2,
  VKB_RUN_RECORD_CMD_BUF = 1 << 3,
  VKB_RUN_UPLOAD_SOURCE  = 1 << 4,
  VKB_RUN_DOWNLOAD_SINK  = 1 << 5,
  VKB_RUN_WAIT_DONE      = 1 << 6,
  VKB_RUN_BEFORE_ACTIVE  = 1 << 7,  /* s_graph_run_before_active: a gui notion (active_module), accepted and without effect here */
  VKB_RUN_PERF           = 1 << 16, /* ours, outside the reference's bits and never implied by VKB_RUN_ALL: time every launch with its
                                       own event pair (-d perf, vkb_graph_perf) instead of replaying the captured CUDA graph; costs
                                       ~1 us of stream idle per launch.  vkb_graph_set_perf() switches the same thing on per graph,
                                       like the reference's `-d perf` log mask (src/pipe/graph.c:881) */
  VKB_RUN_ALL            = -1,
};

VKB_API vkb_graph_t *vkb_graph_new(void);                                   /* dt_graph_init */
VKB_API void vkb_graph_free(vkb_graph_t *g);                                /* dt_graph_cleanup */
VKB_API int  vkb_graph_read_config_ascii(vkb_graph_t *g, const char *filename); /* graph-io.c:282 */
VKB_API int  vkb_graph_read_config_line(vkb_graph_t *g, const char *line);  /* graph-io.c:232: module: connect: param: frames: fps: */
VKB_API int  vkb_graph_replace_display(vkb_graph_t *g, const char *sink_module); /* graph-export.c:23-96, e.g. "o-pfm"; bt2020 / linear */
/* the same with the export colour space (cli --colour-prim / --colour-trc: the values of dt_colour_primaries_t / dt_colour_trc_t,
 * src/pipe/module.h:23-62: prim 1 sRGB/rec709, 2 bt2020, 3 AdobeRGB, 4 P3, 5 XYZ; trc 0 linear, 1 rec709, 2 sRGB, 3 PQ, 4 DCI, 5 HLG,
 * 6 gamma 2.2).  a colenc module is put in front of the sink when the sink is 8 bit or prim / trc are not bt2020 / linear */
VKB_API int  vkb_graph_replace_display_ex(vkb_graph_t *g, const char *inst, const char *sink_module, int prim, int trc);
/* the same with an output size limit (cli --width / --height, dt_graph_export's max_width / max_height, graph-export.c:170-180,
 * 54-62, 93-94): a `resize` module (src/pipe/modules/resize) goes in front of the sink and the sink's roi request shrinks to fit */
VKB_API int  vkb_graph_replace_display_sized(vkb_graph_t *g, const char *inst, const char *sink_module, int prim, int trc, int max_width, int max_height);
/* feed a source module from memory instead of a file (what read_source() would have written into the mapped
 * staging buffer): an already decoded u16 mosaic + the dt_image_params_t fields the source module would fill
 * (src/pipe/module.h:72-109).  the pointer must stay valid until the run that uploads it finished. */
typedef struct vkb_raw_params_t
{
  uint32_t width, height;
  uint32_t filters;            /* 0 rgb, 9 x-trans, else bayer (rggb after alignment) */
  uint32_t crop_aabb[4];
  float    black[4], white[4];
  float    whitebalance[4];
  float    cam_to_rec2020[9];
  float    noise_a, noise_b;
  uint32_t orientation;
  uint32_t packed_bpp;         /* 0: `data` is u16 per pixel; 10/12/14: MLV style packed bit stream, unpacked on the device */
} vkb_raw_params_t;
VKB_API int  vkb_graph_set_source(vkb_graph_t *g, const char *inst, const void *data, const vkb_raw_params_t *p);
/* what `param:i-raw:main:filename:<file>.dng` resolves to (uncompressed 16-bit cfa dng only): the image parameters the
 * reference's loader hands to the graph (i-raw/rawloader-c/lib.rs:137-279, i-raw/main.c:138-256) and the cfa offset of
 * the emitted window.  no GPU needed. */
VKB_API int  vkb_dng_info(const char *filename, vkb_raw_params_t *p, uint32_t *cfa_off_x, uint32_t *cfa_off_y);
/* the uniform block a module's commit_params() produces for the current graph (crop/main.c:277-345: 20 floats,
 * colour/main.c:219-365: 242 floats; modules without commit_params: their raw parameter block): runs the host-side
 * module passes like vkb_graph_plan, then commit_params of that module.  *size: bytes available in, bytes written out.
 * no GPU needed; what the kernels are handed as `params`. */
VKB_API int  vkb_graph_committed_params(vkb_graph_t *g, const char *module, const char *inst, void *out, size_t *size);
/* the module and node layer as text: per module on the path its image parameters, parameter block and connectors, and every
 * node its create_nodes() made (dispatch size, push constants, connector formats / sizes / wiring).  the detailed sibling of
 * vkb_graph_dump_nodes (graph-print.h:76); oracle/ref_nodes_driver.h writes the same text from the reference's own
 * <module>/main.c, which is how the tests pin the node graph.  runs the module passes only, no GPU needed. */
VKB_API int  vkb_graph_describe(vkb_graph_t *g, char *buf, size_t bufsize);
/* what the config lines read so far amount to (graph-io.c:232-315): "frames N", then one line per module in the order of
 * creation, "<name>:<inst> <parameter block, hex> <input connector><<module>.<connector> ...".  host only. */
VKB_API int  vkb_graph_state(vkb_graph_t *g, char *buf, size_t bufsize);
/* the lossless jpeg (LJ92) decoder behind lossless MLV clips, replaces lj92_open + lj92_decode of the reference's vendored
 * liblj92 (i-mlv/video_mlv.c:236-250): headers into width/height/bits/components, and, when `out` is not NULL,
 * width*height*components samples in scan order into out[0..count).  host only, bit exact. */
VKB_API int  vkb_lj92_decode(const uint8_t *data, size_t size, uint16_t *out, size_t count, int *width, int *height, int *bits, int *components);
/* the baseline jpeg writer behind the o-jpg sink (replaces libjpeg's use in o-jpg/main.c:102-172): rgba 8 bit in, 4:4:4 ycbcr,
 * annex K tables scaled by `quality` (libjpeg's scale).  host only */
VKB_API int  vkb_jpeg_write(const char *filename, const uint8_t *rgba, int width, int height, float quality);
/* redirect a sink (o-pfm:main ...) into caller memory instead of a file: rgba f32, wd*ht*16 bytes */
VKB_API int  vkb_graph_set_sink_buffer(vkb_graph_t *g, const char *inst, void *dst, size_t bytes);
/* layout of a sink's pixels, on the device and in the caller's buffer.  VKB_SINK_RGBA_F32 (default for memory sinks) is
 * what the reference maps for write_sink (o-pfm/connectors: rgba f32, 16 B/px).  VKB_SINK_RGB_F32 is the PFM payload
 * itself (o-pfm/main.c:36-40 writes r g b per pixel, 12 B/px): the last kernel of the graph stores it directly, which
 * takes a quarter off the device->host transfer that bounds the end-to-end rate.  o-pfm file output always uses it.
 * takes effect at the next vkb_graph_run with VKB_RUN_ALL. */
#define VKB_SINK_RGBA_F32 0
#define VKB_SINK_RGB_F32  1
/* 8 bit sinks (the sink module's input is rgba:ui8 like o-jpg's, i.e. the graph ends in colenc, src/pipe/graph-export.c:66-86):
 * VKB_SINK_RGBA_UI8 is the image the reference maps for write_sink (4 B/px, alpha 255); VKB_SINK_RGB_UI8 packs r g b (3 B/px,
 * what o-jpg hands to the encoder, o-jpg/main.c:161-165): a quarter of the f32 payload's device->host bytes */
#define VKB_SINK_RGBA_UI8 2
#define VKB_SINK_RGB_UI8  3
VKB_API int  vkb_graph_set_sink_layout(vkb_graph_t *g, const char *inst, int layout);
VKB_API int  vkb_graph_sink_size(vkb_graph_t *g, const char *inst, uint32_t *wd, uint32_t *ht);
VKB_API int  vkb_graph_set_frame(vkb_graph_t *g, uint32_t frame);
/* dt_graph_apply_keyframes (src/pipe/graph.c:1025): evaluate the `keyframe:` lines of the config (frame:module:inst:param:beg:end:
 * values; spelled keyframE / Keyframe / KeyframE / keyFRAME for ease out / ease in / smooth / step, graph-io.c:243-247) at the
 * current frame (vkb_graph_set_frame) into the parameters; the export frame loop calls it before every frame
 * (graph-export.c:280-283).  vkb_graph_has_feedback: 1 if the config held a `feedback:` connection: such a graph carries images
 * from frame to frame (connector.h:116), is not frame parallel, and is refused by vkb_graph_run when a sink reaches the connector */
VKB_API int  vkb_graph_apply_keyframes(vkb_graph_t *g);
VKB_API int  vkb_graph_has_feedback(vkb_graph_t *g);
/* graph->frame_cnt: set by a `frames:` config line or by a source that knows its length (i-mlv's modify_roi_out, i.e. after the
 * first run), what dt_graph_export loops over (src/pipe/graph-export.c:251-268) */
VKB_API int  vkb_graph_frame_count(vkb_graph_t *g);
VKB_API int  vkb_graph_run(vkb_graph_t *g, int runflags);                   /* dt_graph_run, src/pipe/graph.c:719 */
/* -d perf equivalent (graph.c:881-933): per-kernel milliseconds of the last run; returns number of entries */
VKB_API int  vkb_graph_perf(vkb_graph_t *g, char *buf, size_t bufsize);
VKB_API int  vkb_graph_set_perf(vkb_graph_t *g, int on);
VKB_API int  vkb_graph_set_mode(vkb_graph_t *g, int mode);                  /* VKB_MODE_STRICT | VKB_MODE_FAST, before the next run */
/* host-side half of a run only (module passes, node rewrite/fusion, liveness + pool layout): needs no device.
 * writes the launch list as text: one line per kernel launch with its connector images and pool offsets */
VKB_API int  vkb_graph_plan(vkb_graph_t *g, char *buf, size_t bufsize);
VKB_API int  vkb_graph_dump_nodes(vkb_graph_t *g, char *buf, size_t bufsize); /* --dump-nodes, graph-print.h:76 */
/* device-resident variant for kernel-only timing: source already in HBM, sink left in HBM */
VKB_API int  vkb_graph_set_source_device(vkb_graph_t *g, const char *inst, const void *d_data, const vkb_raw_params_t *p);
/* an in-memory i-raw source's DNG OpcodeList2 tag (51009; big endian bytes as stored) and the cfa offset of the emitted window:
 * what the reference's loader hands on as s_image_metadata_dngop (i-raw/main.c:62-92).  denoise applies its four Bayer GainMap
 * opcodes (denoise/main.c:12-43, 172-200; noop.comp:48-57, doub.comp:106-114).  bytes == 0 removes it.  dng FILES carry their own. */
VKB_API int  vkb_dng_opcodes_describe(const void *opcode_list, size_t bytes, char *out, size_t out_size); /* the decoded list as text (i-raw/dng_opcode_decode.c) */
VKB_API int  vkb_graph_set_dng_opcodes(vkb_graph_t *g, const char *inst, const void *opcode_list2, size_t bytes, int cfa_off_x, int cfa_off_y);
VKB_API int  vkb_graph_sink_device(vkb_graph_t *g, const char *inst, void **d_ptr);
VKB_API uint64_t vkb_graph_pool_bytes(vkb_graph_t *g);
VKB_API void *vkb_graph_stream(vkb_graph_t *g);                            /* the cudaStream_t the graph launches on (after the first run) */
/* band split: ONE frame developed by n GPUs of the box, each computing a horizontal band of every
 * kernel launch; halo rows (stencils, pyramid neighbours, denoise's quadrant swizzle) are copied between the devices' pools
 * with peer access (NVLink), pyramid levels below 32 rows per device are computed whole on every device.  the result is bit
 * for bit the one GPU frame.  devices[] are CUDA device ordinals (the same ordinal may be listed more than once: the bands
 * then share that GPU, which is how the one-GPU tests exercise the exchange).  needs a u16 mosaic source in memory (host or
 * device) and the default darkroom graph's kernels (bayer, denoise on); anything else fails loudly at the next run.
 * n <= 1 switches it off.  takes effect at the next vkb_graph_run with VKB_RUN_ALL. */
VKB_API int   vkb_graph_set_bands(vkb_graph_t *g, int n, const int *devices);
/* host only (no GPU needed): the band split as text, per launch and device slot the rows computed, the rows pulled and from
 * which slot, the slots waited for; then the source rows uploaded and sink rows downloaded per slot */
VKB_API int   vkb_graph_band_plan(vkb_graph_t *g, char *buf, size_t bufsize);
/* bytes the devices pull from their peers per frame (all devices / the busiest one), number of copies and kernel launches */
VKB_API int   vkb_graph_band_stats(vkb_graph_t *g, uint64_t *bytes_total, uint64_t *bytes_max_device, int *pulls, int *launches);
/* device side timing of banded frames: mark(0) before, mark(1) after a span of runs; elapsed = the slowest device's span */
VKB_API int   vkb_graph_band_mark(vkb_graph_t *g, int which);
VKB_API int   vkb_graph_band_elapsed_ms(vkb_graph_t *g, float *ms);
VKB_API int   vkb_graph_set_device(vkb_graph_t *g, int device);             /* one graph per GPU: frame-parallel / band-split drivers */                      /* -d mem: peak pooled HBM */

#ifdef __cplusplus
}
#endif
#endif

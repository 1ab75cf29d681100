// graph / module / node model of the engine: an API-compatible restatement of vkdt's src/pipe data model
// (token.h, connector.h:77-210, node.h:19-52, module.h:72-185, global.h:45-121, params.h:6-23) with the Vulkan
// handles stripped.  module authors see the same names: dt_module_t, dt_node_t, dt_connector_t, dt_roi_t,
// dt_token(), dt_node_add(), dt_connector_copy(), dt_connector_bypass(), dt_node_connect(), the callback
// table (init, cleanup, modify_roi_out, modify_roi_in, create_nodes, commit_params, read_source, write_sink).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>
#include <vector>
#include <string>
#include "../vkb_internal.h"

typedef uint64_t dt_token_t;
static inline dt_token_t dt_token(const char *s)
{ // token.h:39-56
  dt_token_t t = 0;
  for(int i = 0; i < 8 && s[i]; i++) t |= (dt_token_t)(uint8_t)s[i] << (8 * i);
  return t;
}
static inline std::string dt_token_string(dt_token_t t)
{
  char b[9] = {0};
  memcpy(b, &t, 8);
  return std::string(b);
}

// ---- connector.h:64-92 ----
enum dt_roi_mark_t
{
  s_roi_mark_uninited = 0, s_roi_mark_soft = 1, s_roi_mark_hard = 2, s_roi_mark_bck = 4, s_roi_mark_fwd = 8,
  s_roi_mark_soft_bck = 5, s_roi_mark_hard_bck = 6, s_roi_mark_soft_fwd = 9, s_roi_mark_hard_fwd = 10,
  s_roi_mark_dontcare = 0xff,
};
struct dt_roi_t { uint32_t full_wd, full_ht, wd, ht, marker; };
static inline int dt_roi_stronger(const dt_roi_t *a, const dt_roi_t *b)
{
  if(a->marker == s_roi_mark_dontcare) return 0;
  if(b->marker == s_roi_mark_dontcare) return 1;
  const int mark = (int)(a->marker & 3u) - (int)(b->marker & 3u);
  if(mark == 0) return (a->marker & 12u) > (b->marker & 12u);
  return mark > 0;
}
struct dt_cid_t { int16_t i, c; };
static const dt_cid_t s_cid_unset = { -1, -1 };
static inline int dt_cid_unset(dt_cid_t id) { return id.i == -1 && id.c == -1; }
static inline dt_cid_t dt_cid(int i, int c) { return dt_cid_t{ (int16_t)i, (int16_t)c }; }

enum dt_connector_flags_t
{
  s_conn_none = 0, s_conn_smooth = 1, s_conn_clear = 2, s_conn_feedback = 4, s_conn_dynamic_array = 8,
  s_conn_protected = 16, s_conn_double_buffer = 32, s_conn_mipmap = 64, s_conn_clear_once = 128,
};
#define DT_MAX_CONNECTORS 30

// connector.h:156-210
struct dt_connector_t
{
  dt_token_t name, type, chan, format;
  uint32_t   flags;
  dt_cid_t   connected;   // inputs: (module|node, connector) of the source. outputs: .i is a reference count
  dt_cid_t   associated;  // module <-> node layer link
  dt_cid_t   bypass;
  dt_roi_t   roi;
  int        max_wd, max_ht;
  int        frames;
  int        array_length;
  int        buf;          // executor: index of the pooled HBM buffer backing this (owner) connector, -1 if none
};
static inline int dt_connector_owner(const dt_connector_t *c) { return c->type == dt_token("write") || c->type == dt_token("source"); }
static inline int dt_connector_input(const dt_connector_t *c) { return c->type == dt_token("read") || c->type == dt_token("sink") || c->type == dt_token("modify"); }
static inline int dt_connector_output(const dt_connector_t *c) { return c->type == dt_token("write") || c->type == dt_token("source") || c->type == dt_token("modify"); }
static inline int dt_connected(const dt_connector_t *c)
{
  if(c->type == dt_token("read") || c->type == dt_token("sink")) return c->connected.i >= 0 && c->connected.c >= 0;
  return c->connected.i > 0;
}
// connector.h:281-301
struct dt_module_t;
int dt_module_source_failed(const dt_module_t *mod); // modules.cpp: a file source without a readable file
static inline int dt_connector_channels(const dt_connector_t *c)
{
  if(c->chan == dt_token("ssbo") || c->chan == dt_token("rggb") || c->chan == dt_token("rgbx")) return 1;
  const uint64_t t = c->chan;
  if(t <= 0xff) return 1;
  if(t <= 0xffff) return 2;
  return 4;
}
static inline size_t dt_connector_bytes_per_channel(const dt_connector_t *c)
{
  if(c->format == dt_token("ui32") || c->format == dt_token("u32") || c->format == dt_token("f32")) return 4;
  if(c->format == dt_token("ui16") || c->format == dt_token("f16")) return 2;
  if(c->format == dt_token("ui8")) return 1;
  return 2;
}

// module.h:72-109
#include "dngop.h"
struct dt_image_params_t
{
  float    black[4], white[4], whitebalance[4];
  uint32_t filters;
  uint32_t crop_aabb[4];
  float    cam_to_rec2020[9];
  uint32_t orientation;
  char     datetime[20], maker[32], model[32];
  float    exposure, aperture, iso, focal_length;
  int      colour_primaries, colour_trc;
  int      snd_format, snd_channels, snd_samplerate;
  float    noise_a, noise_b;
  dt_token_t input_name;
  void    *meta;   // metadata.h's list reduced to the one entry this path reads: a dt_image_metadata_dngop_t of the source module, or 0
};

// params.h:6-23 (gui annotations dropped)
struct dt_ui_param_t
{
  dt_token_t name, type;
  int32_t cnt, offset;
  std::vector<uint8_t> def;   // default value blob
};

// module.h dt_keyframe_t: a parameter (or the sub range [beg, end) of it) pinned at a frame, anim.h: how to get there
struct dt_keyframe_t { int frame; int anim; dt_token_t param; int beg, end; std::vector<uint8_t> data; };
enum { s_anim_lerp = 0, s_anim_step = 1, s_anim_ease_in = 2, s_anim_ease_out = 3, s_anim_smooth = 4 };

struct dt_graph_t;
struct dt_module_t;
struct dt_node_t;
struct dt_read_source_params_t { dt_node_t *node; int c; int a; };
struct dt_write_sink_params_t  { dt_node_t *node; int c; int a; };

// global.h:45-121: the module class ("so" for shared object in the reference; statically registered here)
struct dt_module_so_t
{
  dt_token_t name;
  std::vector<dt_connector_t> connector;
  std::vector<dt_ui_param_t>  param;
  int  (*init)(dt_module_t *);
  void (*cleanup)(dt_module_t *);
  void (*modify_roi_out)(dt_graph_t *, dt_module_t *);
  void (*modify_roi_in)(dt_graph_t *, dt_module_t *);
  void (*create_nodes)(dt_graph_t *, dt_module_t *);
  void (*commit_params)(dt_graph_t *, dt_module_t *);
  int  (*read_source)(dt_module_t *, void *mapped, dt_read_source_params_t *);
  void (*write_sink)(dt_module_t *, void *buf, dt_write_sink_params_t *);
  int  has_source_size;   // b200 extension: source staging size/format hook below is valid
};

enum { s_module_request_none = 0, s_module_request_read_source = 1, s_module_request_write_sink = 2, s_module_request_all = 8 };

// module.h:126-185
struct dt_module_t
{
  dt_module_so_t *so;
  dt_token_t name, inst;
  dt_graph_t *graph;
  int disabled;
  dt_connector_t connector[DT_MAX_CONNECTORS];
  int num_connectors;
  dt_image_params_t img_param;
  uint8_t *param;              // points into the graph's param pool
  int      param_size;
  uint8_t *committed_param;
  int      committed_param_size;
  uint32_t flags;
  void    *data;
  float    gui_x, gui_y;
  std::vector<dt_keyframe_t> keyframe;
};

// node.h:19-52
struct dt_node_t
{
  dt_token_t name, kernel;
  dt_module_t *module;
  dt_connector_t connector[DT_MAX_CONNECTORS];
  int num_connectors;
  uint32_t wd, ht, dp;
  uint32_t flags;
  uint8_t  push_constant[256];
  size_t   push_constant_size;
};

// in-memory source / sink redirection (vkb_graph_set_source / vkb_graph_set_sink_buffer)
struct vkb_mem_source_t { const void *data; int on_device; vkb_raw_params_t p; int valid; };
struct vkb_mem_sink_t   { void *dst; size_t bytes; int valid; int layout; }; // layout: VKB_SINK_RGBA_F32 | VKB_SINK_RGB_F32

struct vkb_plan_t; // executor.cpp

struct dt_graph_t
{
  std::vector<dt_module_t> module;       // stable: reserved up front like the reference's fixed arrays (graph.c:45-56)
  std::vector<dt_node_t>   node;
  std::vector<uint8_t>     params_pool;
  size_t                   params_end;
  uint32_t frame, frame_cnt;
  double   frame_rate;
  dt_image_params_t main_img_param;
  uint32_t runflags;
  char     searchpath[1024];
  char     basedir[1024];
  // b200 executor state
  vkb_plan_t *plan;
  std::vector<vkb_mem_source_t> mem_source;  // indexed by module id
  std::vector<vkb_mem_sink_t>   mem_sink;
  std::string perf_text;
  int      device;
  std::vector<int> band_devices;             // > 1 entries: band split over these CUDA devices (vkb_graph_set_bands)
  int      mode;                             // VKB_MODE_STRICT | VKB_MODE_FAST (vkb_graph_set_mode)
  int      perf;                             // time every launch (vkb_graph_set_perf)
};

// ---- graph api (graph.h:176-194, module.c, connector.inc, graph-io.c) ----
dt_graph_t *dt_graph_new();
void dt_graph_cleanup(dt_graph_t *g);
int  dt_module_add(dt_graph_t *g, dt_token_t name, dt_token_t inst);
int  dt_module_get(const dt_graph_t *g, dt_token_t name, dt_token_t inst);
int  dt_module_remove(dt_graph_t *g, int modid);
int  dt_module_get_connector(const dt_module_t *m, dt_token_t conn);
int  dt_module_get_param(const dt_module_so_t *so, dt_token_t name);
int  dt_module_connect(dt_graph_t *g, int m0, int c0, int m1, int c1);
int  dt_node_connect(dt_graph_t *g, int n0, int c0, int n1, int c1);
int  dt_node_connect_named(dt_graph_t *g, int n0, const char *c0, int n1, const char *c1);
int  dt_graph_read_config_line(dt_graph_t *g, char *line);
int  dt_graph_read_config_ascii(dt_graph_t *g, const char *filename);
int  dt_graph_replace_display(dt_graph_t *g, dt_token_t inst, dt_token_t mod, int prim, int trc, int resize = 0, int max_wd = 0, int max_ht = 0);
void dt_graph_disconnect_display_modules(dt_graph_t *g);
int  dt_graph_run(dt_graph_t *g, uint32_t runflags);
void dt_graph_apply_keyframes(dt_graph_t *g);           // graph.c:1025
int  dt_graph_has_feedback(const dt_graph_t *g);         // any connector flagged s_conn_feedback: frames depend on each other
std::string dt_graph_dump_nodes(dt_graph_t *g);
std::string dt_graph_describe(dt_graph_t *g, const std::vector<int> &modid);

dt_module_so_t *dt_module_so_get(dt_token_t name);   // registry (global.c:442)
int dt_pipe_register_module(const char *name, const char *connectors, const char *params); // a caller's module (vkb_register_module)
int dt_pipe_set_basedir(const char *dir);            // <basedir>/modules/<name>/{connectors,params} override the built-in tables
int dt_module_so_describe(dt_token_t name, std::string *text); // connectors and params of a registered module, in the files' grammar

// ---- module author api (modules/api.h) ----
#define dt_no_roi ((const dt_roi_t *)(uintptr_t)-1)
int  dt_node_add(dt_graph_t *g, dt_module_t *m, const char *name, const char *kernel, int wd, int ht, int dp,
                 int pc_size, const void *pc, int nc, ...);
void dt_connector_copy(dt_graph_t *g, dt_module_t *m, int mc, int nid, int nc);
void dt_connector_bypass(dt_graph_t *g, dt_module_t *m, int mc_in, int mc_out);
const dt_image_params_t *dt_module_get_input_img_param(dt_graph_t *g, dt_module_t *m, dt_token_t input);
static inline const float   *dt_module_param_float(const dt_module_t *m, int p) { return (p >= 0 && p < (int)m->so->param.size()) ? (const float *)(m->param + m->so->param[p].offset) : 0; }
static inline const int32_t *dt_module_param_int(const dt_module_t *m, int p)   { return (p >= 0 && p < (int)m->so->param.size()) ? (const int32_t *)(m->param + m->so->param[p].offset) : 0; }
static inline const char    *dt_module_param_string(const dt_module_t *m, int p) { return (p >= 0 && p < (int)m->so->param.size()) ? (const char *)(m->param + m->so->param[p].offset) : 0; }
int  dt_module_set_param_float(dt_module_t *m, dt_token_t p, float v);
int  dt_module_set_param_float_n(dt_module_t *m, dt_token_t p, const float *v, int n);
int  dt_module_set_param_string(dt_module_t *m, dt_token_t p, const char *str);

// core/gaussian_elimination.h
int gauss_solve(double *A, double *b, int n);
int gauss_make_triangular(double *A, int *p, int n);
void gauss_solve_triangular(const double *A, const int *p, double *b, int n);

// graph layer of the C-ABI (include/vkdt_b200.h §3): thin extern "C" wrappers over the pipe model + executor.
#include "pipe.h"
#include "dng.h"
#include "lj92.h"
#include "mlv.h"
#include "jpeg.h"

int vkb_plan_sink(dt_graph_t *g, int modid, uint32_t *wd, uint32_t *ht, void **dptr);
uint64_t vkb_plan_pool_bytes(dt_graph_t *g);
int dt_graph_plan(dt_graph_t *g, std::string *text);
int dt_graph_run_modules(dt_graph_t *g, std::vector<int> &modid);
void *vkb_plan_stream(dt_graph_t *g);
int vkb_plan_band_stats(dt_graph_t *g, uint64_t *total, uint64_t *max_dev, int *pulls, int *launches);
int vkb_plan_band_mark(dt_graph_t *g, int which);
int vkb_plan_band_elapsed(dt_graph_t *g, float *ms);
int dt_graph_band_plan(dt_graph_t *g, std::string *text);

struct vkb_graph_t { dt_graph_t *g; };

static int find_inst(dt_graph_t *g, const char *inst, bool source)
{ // first module with that instance name whose connector 0 is a source / sink ("main" if inst is null)
  const dt_token_t it = dt_token(inst && inst[0] ? inst : "main");
  for(size_t m = 0; m < g->module.size(); m++)
  {
    const dt_module_t *mod = &g->module[m];
    if(!mod->name || mod->inst != it || !mod->num_connectors) continue;
    if(mod->connector[0].type == dt_token(source ? "source" : "sink") && mod->name != dt_token("display")) return (int)m;
  }
  return -1;
}

int dt_iraw_set_dng_opcodes(dt_module_t *mod, const void *blob, size_t len, int ox, int oy);   // modules.cpp

extern "C" {

vkb_graph_t *vkb_graph_new(void)
{
  vkb_graph_t *h = new vkb_graph_t();
  h->g = dt_graph_new();
  return h;
}
void vkb_graph_free(vkb_graph_t *h)
{
  if(!h) return;
  dt_graph_cleanup(h->g);
  delete h;
}
int vkb_graph_read_config_ascii(vkb_graph_t *h, const char *filename)
{
  if(!h || !filename) return VKB_ERR_BAD_ARG;
  return dt_graph_read_config_ascii(h->g, filename) ? vkb_set_error(VKB_ERR_IO, "could not read config '%s'", filename) : VKB_OK;
}
int vkb_graph_read_config_line(vkb_graph_t *h, const char *line)
{ // 0 ok, > 0 warning (ignored by the reference's reader), < 0 fatal (graph-io.c:297-298)
  if(!h || !line) return VKB_ERR_BAD_ARG;
  std::string s(line);
  return dt_graph_read_config_line(h->g, &s[0]);
}
int vkb_graph_replace_display_ex(vkb_graph_t *h, const char *inst, const char *sink_module, int prim, int trc)
{
  if(!h) return VKB_ERR_BAD_ARG;
  const int m = dt_graph_replace_display(h->g, dt_token(inst && inst[0] ? inst : "main"), dt_token(sink_module && sink_module[0] ? sink_module : "o-jpg"), prim, trc);
  if(m < 0) return vkb_set_error(VKB_ERR_GRAPH, "replace display failed (%d)", m);
  dt_graph_disconnect_display_modules(h->g);
  return VKB_OK;
}
int vkb_graph_replace_display_sized(vkb_graph_t *h, const char *inst, const char *sink_module, int prim, int trc, int max_width, int max_height)
{ // graph-export.c:170-180: a resize module is inserted when either limit is given
  if(!h || max_width < 0 || max_height < 0) return VKB_ERR_BAD_ARG;
  const int m = dt_graph_replace_display(h->g, dt_token(inst && inst[0] ? inst : "main"), dt_token(sink_module && sink_module[0] ? sink_module : "o-jpg"), prim, trc,
      max_width > 0 || max_height > 0, max_width, max_height);
  if(m < 0) return vkb_set_error(VKB_ERR_GRAPH, "replace display failed (%d)", m);
  dt_graph_disconnect_display_modules(h->g);
  return VKB_OK;
}
int vkb_graph_replace_display(vkb_graph_t *h, const char *sink_module)
{
  if(!h) return VKB_ERR_BAD_ARG;
  const int m = dt_graph_replace_display(h->g, dt_token("main"), dt_token(sink_module && sink_module[0] ? sink_module : "o-pfm"), 2, 0);
  if(m < 0) return vkb_set_error(VKB_ERR_GRAPH, "replace display failed (%d)", m);
  dt_graph_disconnect_display_modules(h->g);
  return VKB_OK;
}
static int set_source(vkb_graph_t *h, const char *inst, const void *data, const vkb_raw_params_t *p, int on_device)
{
  if(!h || !data || !p) return VKB_ERR_BAD_ARG;
  const int m = find_inst(h->g, inst, true);
  if(m < 0) return vkb_set_error(VKB_ERR_BAD_ARG, "no source module with instance '%s'", inst ? inst : "main");
  if(p->packed_bpp && h->g->module[m].name != dt_token("i-mlv")) return vkb_set_error(VKB_ERR_BAD_ARG, "packed payloads go through i-mlv");
  vkb_mem_source_t *s = &h->g->mem_source[m];
  s->data = data; s->on_device = on_device; s->p = *p; s->valid = 1;
  return VKB_OK;
}
int vkb_graph_set_source(vkb_graph_t *h, const char *inst, const void *data, const vkb_raw_params_t *p) { return set_source(h, inst, data, p, 0); }
int vkb_graph_set_source_device(vkb_graph_t *h, const char *inst, const void *d, const vkb_raw_params_t *p) { return set_source(h, inst, d, p, 1); }
int vkb_dng_opcodes_describe(const void *opcode_list, size_t bytes, char *out, size_t out_size)
{ // what the opcode list decoder makes of a tag, as text (tests compare it with the reference's own decoder)
  if(!opcode_list || !out || !out_size) return VKB_ERR_BAD_ARG;
  dt_dng_opcode_list_t ol;
  std::string t;
  char b[512];
  if(dng_opcode_list_decode((const uint8_t *)opcode_list, bytes, &ol)) t = "none\n";
  else
  {
    snprintf(b, sizeof(b), "count %d\n", (int)ol.ops.size()); t += b;
    for(const dt_dng_opcode_t &op : ol.ops)
    {
      snprintf(b, sizeof(b), "op %u optional %u preview_skip %u\n", op.id, op.optional, op.preview_skip); t += b;
      if(op.gain_map < 0) continue;
      const dt_dng_gain_map_t &g = ol.gain_maps[op.gain_map];
      snprintf(b, sizeof(b), " region %u %u %u %u plane %u %u pitch %u %u points %u %u spacing %.17g %.17g origin %.17g %.17g planes %u\n gains",
          g.top, g.left, g.bottom, g.right, g.plane, g.planes, g.row_pitch, g.col_pitch, g.map_points_v, g.map_points_h,
          g.map_spacing_v, g.map_spacing_h, g.map_origin_v, g.map_origin_h, g.map_planes);
      t += b;
      for(float v : g.map_gain) { uint32_t u; memcpy(&u, &v, 4); snprintf(b, sizeof(b), " %08x", u); t += b; }
      t += "\n";
    }
  }
  snprintf(out, out_size, "%s", t.c_str());
  return VKB_OK;
}
int vkb_graph_set_dng_opcodes(vkb_graph_t *h, const char *inst, const void *opcode_list2, size_t bytes, int cfa_off_x, int cfa_off_y)
{ // what rawler hands the reference beside the pixels (i-raw/main.c:62-92): the OpcodeList2 tag and the cfa offset of the window
  if(!h) return VKB_ERR_BAD_ARG;
  const int m = find_inst(h->g, inst, true);
  if(m < 0) return vkb_set_error(VKB_ERR_BAD_ARG, "no source module with instance '%s'", inst ? inst : "main");
  const int r = dt_iraw_set_dng_opcodes(&h->g->module[m], opcode_list2, bytes, cfa_off_x, cfa_off_y);
  if(r == 1) return vkb_set_error(VKB_ERR_BAD_ARG, "dng opcode lists belong to an i-raw source");
  if(r) return vkb_set_error(VKB_ERR_BAD_ARG, "the opcode list does not decode (sizes that do not add up)");
  return VKB_OK;
}
int vkb_graph_set_sink_buffer(vkb_graph_t *h, const char *inst, void *dst, size_t bytes)
{ // dst == NULL: keep the result on the device (no download, no file)
  if(!h) return VKB_ERR_BAD_ARG;
  const int m = find_inst(h->g, inst, false);
  if(m < 0) return vkb_set_error(VKB_ERR_BAD_ARG, "no sink module with instance '%s'", inst ? inst : "main");
  const int layout = h->g->mem_sink[m].layout;
  h->g->mem_sink[m] = vkb_mem_sink_t{ dst, bytes, 1, layout };
  return VKB_OK;
}
int vkb_graph_set_sink_layout(vkb_graph_t *h, const char *inst, int layout)
{
  if(!h || layout < VKB_SINK_RGBA_F32 || layout > VKB_SINK_RGB_UI8) return VKB_ERR_BAD_ARG;
  const int m = find_inst(h->g, inst, false);
  if(m < 0) return vkb_set_error(VKB_ERR_BAD_ARG, "no sink module with instance '%s'", inst ? inst : "main");
  h->g->mem_sink[m].layout = layout;
  return VKB_OK;
}
int vkb_graph_sink_size(vkb_graph_t *h, const char *inst, uint32_t *wd, uint32_t *ht)
{
  if(!h) return VKB_ERR_BAD_ARG;
  const int m = find_inst(h->g, inst, false);
  if(m < 0) return vkb_set_error(VKB_ERR_BAD_ARG, "no sink module with instance '%s'", inst ? inst : "main");
  return vkb_plan_sink(h->g, m, wd, ht, 0) ? vkb_set_error(VKB_ERR_GRAPH, "graph has not been run yet") : VKB_OK;
}
int vkb_graph_sink_device(vkb_graph_t *h, const char *inst, void **d_ptr)
{
  if(!h) return VKB_ERR_BAD_ARG;
  const int m = find_inst(h->g, inst, false);
  if(m < 0) return vkb_set_error(VKB_ERR_BAD_ARG, "no sink module with instance '%s'", inst ? inst : "main");
  return vkb_plan_sink(h->g, m, 0, 0, d_ptr) ? vkb_set_error(VKB_ERR_GRAPH, "graph has not been run yet") : VKB_OK;
}
int vkb_graph_set_frame(vkb_graph_t *h, uint32_t frame) { if(!h) return VKB_ERR_BAD_ARG; h->g->frame = frame; return VKB_OK; }
int vkb_graph_apply_keyframes(vkb_graph_t *h) { if(!h) return VKB_ERR_BAD_ARG; dt_graph_apply_keyframes(h->g); return VKB_OK; }
int vkb_graph_has_feedback(vkb_graph_t *h) { return h ? dt_graph_has_feedback(h->g) : VKB_ERR_BAD_ARG; }
int vkb_graph_frame_count(vkb_graph_t *h) { return h ? (int)h->g->frame_cnt : VKB_ERR_BAD_ARG; }
int vkb_graph_run(vkb_graph_t *h, int runflags) { if(!h) return VKB_ERR_BAD_ARG; return dt_graph_run(h->g, (uint32_t)runflags); }
int vkb_graph_perf(vkb_graph_t *h, char *buf, size_t bufsize)
{
  if(!h || !buf || !bufsize) return VKB_ERR_BAD_ARG;
  snprintf(buf, bufsize, "%s", h->g->perf_text.c_str());
  int n = 0;
  for(char c : h->g->perf_text) n += c == '\n';
  return n;
}
int vkb_graph_set_mode(vkb_graph_t *h, int mode) { if(!h || (mode != VKB_MODE_STRICT && mode != VKB_MODE_FAST)) return VKB_ERR_BAD_ARG; h->g->mode = mode; return VKB_OK; }
int vkb_graph_set_perf(vkb_graph_t *h, int on) { if(!h) return VKB_ERR_BAD_ARG; h->g->perf = on != 0; return VKB_OK; }
int vkb_graph_dump_nodes(vkb_graph_t *h, char *buf, size_t bufsize)
{
  if(!h || !buf || !bufsize) return VKB_ERR_BAD_ARG;
  const std::string s = dt_graph_dump_nodes(h->g);
  snprintf(buf, bufsize, "%s", s.c_str());
  return (int)s.size();
}
int vkb_graph_state(vkb_graph_t *h, char *buf, size_t bufsize)
{
  if(!h || !buf || !bufsize) return VKB_ERR_BAD_ARG;
  std::string s;
  char b[64];
  snprintf(b, sizeof(b), "frames %u\n", h->g->frame_cnt);
  s += b;
  for(const dt_module_t &m : h->g->module)
  {
    if(!m.name) continue; // removed (module.c:dt_module_remove leaves a hole)
    s += dt_token_string(m.name) + ":" + dt_token_string(m.inst) + " ";
    for(int k = 0; k < m.param_size; k++) { snprintf(b, sizeof(b), "%02x", m.param[k]); s += b; }
    for(int c = 0; c < m.num_connectors; c++) if(dt_connector_input(m.connector + c))
    { snprintf(b, sizeof(b), " %s<%d.%d", dt_token_string(m.connector[c].name).c_str(), m.connector[c].connected.i, m.connector[c].connected.c); s += b; }
    s += "\n";
  }
  if(s.size() + 1 > bufsize) return vkb_set_error(VKB_ERR_BAD_ARG, "buffer too small: %zu < %zu", bufsize, s.size() + 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  return VKB_OK;
}
int vkb_graph_describe(vkb_graph_t *h, char *buf, size_t bufsize)
{
  if(!h || !buf || !bufsize) return VKB_ERR_BAD_ARG;
  std::vector<int> modid;
  const int r = dt_graph_run_modules(h->g, modid); // module passes only: roi out, roi in, create nodes
  if(r) return r;
  const std::string s = dt_graph_describe(h->g, modid);
  if(s.size() + 1 > bufsize) return vkb_set_error(VKB_ERR_BAD_ARG, "buffer too small: %zu < %zu", bufsize, s.size() + 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  return VKB_OK;
}
int vkb_graph_plan(vkb_graph_t *h, char *buf, size_t bufsize)
{
  if(!h) return VKB_ERR_BAD_ARG;
  std::string s;
  const int r = dt_graph_plan(h->g, &s);
  if(r) return r;
  if(buf && bufsize) snprintf(buf, bufsize, "%s", s.c_str());
  return VKB_OK;
}
int vkb_graph_committed_params(vkb_graph_t *h, const char *module, const char *inst, void *out, size_t *size)
{
  if(!h || !module || !inst || !out || !size) return VKB_ERR_BAD_ARG;
  std::string s;
  const int r = dt_graph_plan(h->g, &s); // module passes: roi + img_param of every module on the path
  if(r) return r;
  const int m = dt_module_get(h->g, dt_token(module), dt_token(inst));
  if(m < 0) return vkb_set_error(VKB_ERR_BAD_ARG, "no module %s:%s", module, inst);
  dt_module_t *mod = &h->g->module[m];
  if(mod->so->commit_params) mod->so->commit_params(h->g, mod);
  const uint8_t *src = mod->committed_param_size ? mod->committed_param : mod->param;
  const size_t n = mod->committed_param_size ? (size_t)mod->committed_param_size : (size_t)mod->param_size;
  if(*size < n) return vkb_set_error(VKB_ERR_BAD_ARG, "buffer too small: %zu < %zu", *size, n);
  memcpy(out, src, n);
  *size = n;
  return VKB_OK;
}
void *vkb_graph_stream(vkb_graph_t *h) { return h ? vkb_plan_stream(h->g) : 0; }
int vkb_dng_info(const char *filename, vkb_raw_params_t *p, uint32_t *ox, uint32_t *oy)
{
  if(!filename || !p) return VKB_ERR_BAD_ARG;
  dng_image_t img;
  if(dng_read(filename, &img)) return VKB_ERR_IO;
  return dng_raw_params(&img, p, ox, oy) ? VKB_ERR_BAD_ARG : VKB_OK;
}
int vkb_lj92_decode(const uint8_t *data, size_t size, uint16_t *out, size_t count, int *w, int *h, int *bits, int *comps)
{
  if(!data || !size) return VKB_ERR_BAD_ARG;
  if(lj92_info(data, size, w, h, bits, comps)) return vkb_set_error(VKB_ERR_IO, "not a lossless jpeg stream this decoder handles");
  if(out && lj92_decode(data, size, out, count)) return vkb_set_error(VKB_ERR_IO, "lossless jpeg stream is corrupt or the output buffer too small");
  return VKB_OK;
}
int vkb_graph_set_bands(vkb_graph_t *h, int n, const int *devices)
{
  if(!h || n < 0 || n > 64 || (n > 0 && !devices)) return VKB_ERR_BAD_ARG;
  h->g->band_devices.assign(devices, devices + n);
  if(n == 1) { h->g->device = devices[0]; h->g->band_devices.clear(); }
  return VKB_OK;
}
int vkb_graph_band_plan(vkb_graph_t *h, char *buf, size_t bufsize)
{
  if(!h || !buf || !bufsize) return VKB_ERR_BAD_ARG;
  std::string s;
  const int r = dt_graph_band_plan(h->g, &s);
  if(r) return r;
  if(s.size() + 1 > bufsize) return vkb_set_error(VKB_ERR_BAD_ARG, "buffer too small: %zu < %zu", bufsize, s.size() + 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  return VKB_OK;
}
int vkb_graph_band_stats(vkb_graph_t *h, uint64_t *bytes_total, uint64_t *bytes_max_device, int *pulls, int *launches)
{ if(!h) return VKB_ERR_BAD_ARG; return vkb_plan_band_stats(h->g, bytes_total, bytes_max_device, pulls, launches) ? vkb_set_error(VKB_ERR_GRAPH, "no band split planned") : VKB_OK; }
int vkb_graph_band_mark(vkb_graph_t *h, int which) { if(!h) return VKB_ERR_BAD_ARG; return vkb_plan_band_mark(h->g, which); }
int vkb_graph_band_elapsed_ms(vkb_graph_t *h, float *ms) { if(!h || !ms) return VKB_ERR_BAD_ARG; return vkb_plan_band_elapsed(h->g, ms); }
int vkb_register_module(const char *name, const char *connectors, const char *params)
{
  if(dt_pipe_register_module(name, connectors, params)) return vkb_set_error(VKB_ERR_BAD_ARG, "module name (1..8 chars) and connectors are needed");
  return VKB_OK;
}
int vkb_register_kernel(const char *name, const char *kernel, vkb_kernel_fn_t fn, int mode)
{
  if(!name || !kernel || !fn || mode > VKB_MODE_FAST) return VKB_ERR_BAD_ARG;
  if(mode < 0) { vkb_register_kernel(name, kernel, (vkb_kernel_fn)fn, VKB_MODE_STRICT); vkb_register_kernel(name, kernel, (vkb_kernel_fn)fn, VKB_MODE_FAST); }
  else vkb_register_kernel(name, kernel, (vkb_kernel_fn)fn, mode);
  return VKB_OK;
}
int vkb_set_basedir(const char *dir) { return dt_pipe_set_basedir(dir); }
int vkb_module_describe(const char *name, char *buf, size_t bufsize)
{
  if(!name || !buf || !bufsize) return VKB_ERR_BAD_ARG;
  std::string s;
  if(dt_module_so_describe(dt_token(name), &s)) return vkb_set_error(VKB_ERR_BAD_ARG, "no module '%s'", name);
  if(s.size() + 1 > bufsize) return vkb_set_error(VKB_ERR_BAD_ARG, "buffer too small: %zu < %zu", bufsize, s.size() + 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  return VKB_OK;
}
int vkb_jpeg_write(const char *filename, const uint8_t *rgba, int width, int height, float quality)
{
  if(!filename || !rgba) return VKB_ERR_BAD_ARG;
  return jpeg_write_rgba8(filename, rgba, width, height, quality) ? vkb_set_error(VKB_ERR_IO, "could not write '%s'", filename) : VKB_OK;
}
int vkb_graph_set_device(vkb_graph_t *h, int device) { if(!h) return VKB_ERR_BAD_ARG; h->g->device = device; return VKB_OK; }
uint64_t vkb_graph_pool_bytes(vkb_graph_t *h) { return h ? vkb_plan_pool_bytes(h->g) : 0; }

} // extern "C"

/* ORACLE — test infrastructure.  the MODULE PASS of the reference's own graph on its own default darkroom config:
 * global.c (module classes from modules/<name>/{connectors,params}), module.c (dt_module_add), graph-io.c (the .cfg
 * reader), connector.c, graph-export.c (dt_graph_replace_display: what vkdt-cli does to the display) and
 * graph-run-modules.h (roi out, roi in, create nodes, repointing through the module layer) are compiled where they lie
 * under /root/reference by `make -C oracle ref`, never copied.  no Vulkan: ref_stub/vulkan/vulkan.h has opaque types and
 * the few vk* entry points the pass mentions are stubs below; the pass is run without s_graph_run_alloc, so no device
 * object is ever asked for.
 * module callbacks are bound by hand instead of dlopen(lib<name>.so): denoise, hilite, demosaic, llap, filmcurv
 * (ref_nodes_shim.c), crop, colour (ref_crop_shim.c, ref_colour_shim.c) are the reference's own main.c; every other
 * module on the path (grade, o-pfm, display...) has no host callbacks on this path in the reference either and runs the
 * graph's defaults.  i-mlv is the reference's own main.c too (clip header -> image parameters, for .cfg files that name it).
 * i-raw alone is a stand-in (its main.c needs the un-vendored rawspeed / rawler): it hands over the
 * image parameters and size the caller gives, like the product's in-memory source.
 * the result is written in the text form of ref_nodes_driver.h / vkb_graph_describe. */
#include "pipe/global.c"
#include "pipe/graph-export.c"
#include "pipe/graph-run-modules.h"
#include "ref_nodes_driver.h"

/* symbols the compiled reference files bind to but never reach on this path */
VkResult vkMapMemory() { return VK_SUCCESS; }
void     vkUnmapMemory() {}
VkResult vkQueueWaitIdle() { return VK_SUCCESS; }
void     vkDestroyDescriptorSetLayout() {}
VkResult vkCreateDescriptorSetLayout() { return VK_SUCCESS; }
VkResult vkWaitSemaphores() { return VK_SUCCESS; }
void     vkDestroyBuffer() {}
void     vkDestroyImageView() {}
void     vkDestroyImage() {}
void     vkDestroyPipelineLayout() {}
void     vkDestroyPipeline() {}
void     vkDestroyFramebuffer() {}
void     vkDestroyRenderPass() {}
void     vkDestroySampler() {}
void     vkDestroySamplerYcbcrConversion() {}

/* ...and of graph.c / raytrace.c / log.c, which need a device: only dt_graph_export (unused here) and error paths call these */
dt_log_t dt_log_global;
const char *qvk_result_to_string(VkResult r) { return "vk"; }
VkResult dt_graph_run(dt_graph_t *g, dt_graph_run_t run) { return VK_INCOMPLETE; }
void dt_graph_apply_keyframes(dt_graph_t *g) {}
dt_connector_image_t *dt_graph_connector_image(dt_graph_t *g, int nid, int cid, int array, int dbuf) { return 0; }
void dt_raytrace_node_cleanup(dt_node_t *node) {}
void dt_raytrace_graph_cleanup(dt_graph_t *graph) {}

#define REF_DECL(M) \
  void M##_ref_create_nodes(dt_graph_t *, dt_module_t *); void M##_ref_modify_roi_out(dt_graph_t *, dt_module_t *); \
  void M##_ref_modify_roi_in(dt_graph_t *, dt_module_t *); int M##_ref_init(dt_module_t *); void M##_ref_cleanup(dt_module_t *); \
  void M##_ref_commit_params(dt_graph_t *, dt_module_t *);
REF_DECL(imlv) REF_DECL(ipfm) REF_DECL(ilut) REF_DECL(denoise) REF_DECL(hilite) REF_DECL(demosaic) REF_DECL(llap) REF_DECL(filmcurv) REF_DECL(crop) REF_DECL(colour) REF_DECL(resize)

static const ref_nodes_in_t *ref_src; /* what the stand-in source hands over */
static void src_modify_roi_out(dt_graph_t *graph, dt_module_t *mod)
{ /* the contract of i-raw/main.c:138-259: image parameters, full size, channels by mosaic type */
  const ref_nodes_in_t *in = ref_src;
  for(int k = 0; k < 4; k++) { mod->img_param.black[k] = in->black[k]; mod->img_param.white[k] = in->white[k]; mod->img_param.whitebalance[k] = in->wb[k]; mod->img_param.crop_aabb[k] = in->crop_aabb[k]; }
  mod->img_param.filters = in->filters; mod->img_param.noise_a = in->noise_a; mod->img_param.noise_b = in->noise_b;
  mod->img_param.cam_to_rec2020[0] = mod->img_param.cam_to_rec2020[4] = mod->img_param.cam_to_rec2020[8] = 1.0f;
  mod->connector[0].chan = in->filters ? dt_token("rggb") : dt_token("rgba");
  mod->connector[0].roi.full_wd = in->in_full_wd; mod->connector[0].roi.full_ht = in->in_full_ht;
}

static dt_module_so_t *so_get(const char *name)
{
  for(int i = 0; i < dt_pipe.num_modules; i++) if(dt_pipe.module[i].name == dt_token(name)) return dt_pipe.module + i;
  return 0;
}

static int ref_pipe_init(const char *basedir)
{
  static int inited = 0;
  if(!inited)
  { /* dt_pipe_global_init (global.c:442-481) with the base directory given instead of the one of the executable */
    memset(&dt_pipe, 0, sizeof(dt_pipe));
    (void)setlocale(LC_ALL, "C");
    snprintf(dt_pipe.basedir, sizeof(dt_pipe.basedir), "%s", basedir);
    snprintf(dt_pipe.homedir, sizeof(dt_pipe.homedir), "/nonexistent");
    static const char *names[] = { "i-raw", "denoise", "hilite", "demosaic", "colour", "filmcurv", "llap", "grade", "hist", "zones", "crop", "lens", "pick",
      "display", "o-pfm", "colenc", "resize", "i-mlv", "contrast", "i-pfm", "i-lut", "o-jpg", 0 };
    int n = 0; while(names[n]) n++;
    dt_pipe.module = malloc(sizeof(dt_module_so_t) * n);
    int i = 0;
    for(int k = 0; k < n; k++) if(!dt_module_so_load(dt_pipe.module + i, names[k])) i++;
    dt_pipe.num_modules = i;
    qsort(dt_pipe.module, dt_pipe.num_modules, sizeof(dt_pipe.module[0]), &compare_module_name);
#define BIND(M, RO, RI, IN, CP) { dt_module_so_t *so = so_get(#M); if(!so) return -10; so->create_nodes = M##_ref_create_nodes; \
    if(RO) so->modify_roi_out = M##_ref_modify_roi_out; if(RI) so->modify_roi_in = M##_ref_modify_roi_in; \
    if(IN) so->init = M##_ref_init; if(CP) so->commit_params = M##_ref_commit_params; }
    BIND(denoise, 1, 1, 1, 0) BIND(hilite, 0, 0, 0, 0) BIND(demosaic, 1, 1, 0, 0) BIND(llap, 0, 0, 0, 0) BIND(filmcurv, 1, 0, 0, 0) BIND(resize, 1, 1, 0, 0)
    { dt_module_so_t *so = so_get("denoise"); so->cleanup = denoise_ref_cleanup; }
    { dt_module_so_t *so = so_get("crop"); if(!so) return -10; so->modify_roi_out = crop_ref_modify_roi_out; so->modify_roi_in = crop_ref_modify_roi_in; so->init = crop_ref_init; so->commit_params = crop_ref_commit_params; }
    { dt_module_so_t *so = so_get("colour"); if(!so) return -10; so->modify_roi_out = colour_ref_modify_roi_out; so->modify_roi_in = colour_ref_modify_roi_in; so->init = colour_ref_init; so->commit_params = colour_ref_commit_params; so->create_nodes = colour_ref_create_nodes; }
    { dt_module_so_t *so = so_get("i-raw"); if(!so) return -10; so->modify_roi_out = src_modify_roi_out; }
    { dt_module_so_t *so = so_get("i-pfm"); if(!so) return -10; so->init = ipfm_ref_init; so->cleanup = ipfm_ref_cleanup; so->modify_roi_out = ipfm_ref_modify_roi_out; }
    { dt_module_so_t *so = so_get("i-lut"); if(!so) return -10; so->init = ilut_ref_init; so->cleanup = ilut_ref_cleanup; so->modify_roi_out = ilut_ref_modify_roi_out; }
    { dt_module_so_t *so = so_get("i-mlv"); if(!so) return -10; so->init = imlv_ref_init; so->cleanup = imlv_ref_cleanup; so->modify_roi_out = imlv_ref_modify_roi_out; }
    inited = 1;
  }
  return 0;
}

/* cfgfile: a .cfg of the reference (bin/default-darkroom.i-raw); extra: more config lines, '\n' separated (may be 0);
 * basedir: <reference>/src/pipe (holds modules/); sink: module the display is replaced by, e.g. "o-pfm".
 * writes one block per module on the path, execution order.  returns bytes written, < 0 on failure. */
int ref_graph_describe(const char *basedir, const char *cfgfile, const char *extra, const char *sink, const ref_nodes_in_t *in, char *out, int outsize)
{
  { const int r = ref_pipe_init(basedir); if(r) return r; }
  ref_src = in;
  /* dt_graph_init (graph.c:34-56) without the device objects */
  dt_graph_t *g = calloc(1, sizeof(*g));
  g->frame_cnt = 1;
  g->max_modules = 100; g->module = calloc(sizeof(dt_module_t), g->max_modules);
  g->max_nodes = 4000;  g->node = calloc(sizeof(dt_node_t), g->max_nodes);
  g->params_max = 16u << 20; g->params_pool = calloc(1, g->params_max);
  g->conn_image_max = 30*2*2000; g->conn_image_pool = calloc(sizeof(dt_connector_image_t), g->conn_image_max);
  int ret = -20;
  int max_wd = 0, max_ht = 0, prim = s_colour_primaries_2020, trc = s_colour_trc_linear;
  char sinkname[16]; snprintf(sinkname, sizeof(sinkname), "%s", sink);
  if(dt_graph_read_config_ascii(g, cfgfile)) goto done;
  if(extra && extra[0])
  {
    char *copy = strdup(extra), *c = copy;
    while(c && *c)
    {
      char *e = strchr(c, '\n'); if(e) *e++ = 0;
      char line[4096]; snprintf(line, sizeof(line), "%s", c);
      /* "#export:max:<w>:<h>": vkdt-cli --width / --height (cli/main.c:68-71 -> dt_graph_export -> replace_display's resize) */
      if(!strncmp(line, "#export:max:", 12)) { sscanf(line + 12, "%d:%d", &max_wd, &max_ht); c = e; continue; }
      /* "#export:colour:<prim>:<trc>": --colour-prim / --colour-trc (cli/main.c:58-59, :80-81; the cli's own default is 1:1);
       * "#export:sink:<module>": --format */
      if(!strncmp(line, "#export:colour:", 15)) { sscanf(line + 15, "%d:%d", &prim, &trc); c = e; continue; }
      if(!strncmp(line, "#export:sink:", 13)) { snprintf(sinkname, sizeof(sinkname), "%s", line + 13); c = e; continue; }
      if(line[0] && dt_graph_read_config_line(g, line) < 0) { free(copy); ret = -21; goto done; }
      c = e;
    }
    free(copy);
  }
  /* what vkdt-cli does (cli/main.c, graph-export.c:160-230): the main display becomes the output module, linear rec2020 */
  if(dt_graph_replace_display(g, dt_token("main"), 0, dt_token(sinkname), max_wd > 0 || max_ht > 0, max_wd, max_ht, prim, trc) < 0) { ret = -22; goto done; }
  dt_graph_disconnect_display_modules(g);
  {
    dt_graph_run_t run = s_graph_run_roi | s_graph_run_create_nodes;
    uint32_t modid[100];
    dt_module_flags_t flags = 0;
    if(dt_graph_run_modules(g, &run, modid, &flags) != VK_SUCCESS) { ret = -23; goto done; }
    /* commit_params of every module on the path (graph-run-modules.h:8-30 without the mapped uniform memory) */
    int cnt = 0;
    {
      dt_module_t *const arr = g->module;
      const int arr_cnt = g->num_modules;
      uint32_t order[100];
#define TRAVERSE_POST order[cnt++] = curr;
#include "pipe/graph-traverse.inc"
      char *o = out; int left = outsize; char b0[9], b1[9], b2[9], b3[9];
      for(int mi = 0; mi < cnt; mi++)
      {
        dt_module_t *mod = g->module + order[mi];
        if(mod->connector[0].roi.full_wd == 0) continue;
        const dt_image_params_t *ip = &mod->img_param;
        ref_out(&o, &left, "module %s filters=%u black=%08x,%08x,%08x,%08x white=%08x,%08x,%08x,%08x wb=%08x,%08x,%08x,%08x crop=%u,%u,%u,%u noise=%08x,%08x\n", ref_tkn(mod->name, b0), ip->filters,
            ref_fbits(ip->black[0]), ref_fbits(ip->black[1]), ref_fbits(ip->black[2]), ref_fbits(ip->black[3]),
            ref_fbits(ip->white[0]), ref_fbits(ip->white[1]), ref_fbits(ip->white[2]), ref_fbits(ip->white[3]),
            ref_fbits(ip->whitebalance[0]), ref_fbits(ip->whitebalance[1]), ref_fbits(ip->whitebalance[2]), ref_fbits(ip->whitebalance[3]),
            ip->crop_aabb[0], ip->crop_aabb[1], ip->crop_aabb[2], ip->crop_aabb[3], ref_fbits(ip->noise_a), ref_fbits(ip->noise_b));
        ref_out(&o, &left, " params ");
        for(int p = 0; p < mod->so->num_params; p++) if(mod->so->param[p]->type == dt_token("string"))
        { /* what follows a string's terminator is whatever the allocation held: not part of the value */
          char *str = (char *)mod->param + mod->so->param[p]->offset;
          const int len = strnlen(str, mod->so->param[p]->cnt);
          memset(str + len, 0, mod->so->param[p]->cnt - len);
        }
        for(int k = 0; k < mod->param_size; k++) ref_out(&o, &left, "%02x", mod->param[k]);
        ref_out(&o, &left, "\n");
        for(int i = 0; i < mod->num_connectors; i++)
        {
          const dt_connector_t *c = mod->connector + i;
          ref_out(&o, &left, " mconn %d %s:%s:%s:%s roi=%ux%u/%ux%u m=%u bypass=%d\n", i, ref_tkn(c->name, b0), ref_tkn(c->type, b1), ref_tkn(c->chan, b2), ref_tkn(c->format, b3),
              c->roi.full_wd, c->roi.full_ht, c->roi.wd, c->roi.ht, c->roi.marker, dt_cid_unset(c->bypass) ? -1 : c->bypass.c);
        }
        int first = -1;
        for(uint32_t n = 0; n < g->num_nodes; n++) if(g->node[n].module == mod) { first = n; break; }
        for(uint32_t n = 0; n < g->num_nodes; n++)
        {
          const dt_node_t *nd = g->node + n;
          if(nd->module != mod) continue;
          ref_out(&o, &left, " node %d %s:%s %ux%ux%u pc=%d:", (int)n - first, ref_tkn(nd->name, b0), ref_tkn(nd->kernel, b1), nd->wd, nd->ht, nd->dp, (int)nd->push_constant_size);
          for(size_t k = 0; k < nd->push_constant_size / 4; k++) ref_out(&o, &left, "%s%08x", k ? "," : "", nd->push_constant[k]);
          ref_out(&o, &left, "\n");
          for(int i = 0; i < nd->num_connectors; i++)
          {
            const dt_connector_t *c = nd->connector + i;
            ref_out(&o, &left, "  conn %d %s:%s:%s:%s roi=%ux%u/%ux%u al=%d ", i, ref_tkn(c->name, b0), ref_tkn(c->type, b1), ref_tkn(c->chan, b2), ref_tkn(c->format, b3),
                c->roi.full_wd, c->roi.full_ht, c->roi.wd, c->roi.ht, c->array_length);
            const int copied = !dt_cid_unset(c->associated) && (dt_connector_input(c) ||
                (c->associated.i == (int)order[mi] && mod->connector[c->associated.c].associated.i == (int)n && mod->connector[c->associated.c].associated.c == i));
            if(copied)                                            ref_out(&o, &left, "mod.%d\n", c->associated.c);
            else if(dt_connector_input(c) && c->connected.i >= 0) ref_out(&o, &left, "n%d.%d\n", c->connected.i - first, c->connected.c);
            else if(dt_connector_input(c))                        ref_out(&o, &left, "open\n");
            else                                                  ref_out(&o, &left, "own\n");
          }
        }
        if(mod->so->commit_params && mod->committed_param_size)
        {
          mod->so->commit_params(g, mod);
          ref_out(&o, &left, " committed ");
          for(int k = 0; k < mod->committed_param_size; k++) ref_out(&o, &left, "%02x", mod->committed_param[k]);
          ref_out(&o, &left, "\n");
        }
      }
      ret = left > 0 ? (int)(o - out) : -1;
    }
  }
done:
  for(uint32_t m = 0; m < g->num_modules; m++) if(g->module[m].name && g->module[m].so && g->module[m].so->cleanup) g->module[m].so->cleanup(g->module + m);
  free(g->conn_image_pool); free(g->params_pool); free(g->node); free(g->module); free(g);
  return ret;
}

/* the reference's config grammar (graph-io.c:232-252 and the readers above it) line by line on top of a loaded .cfg: return
 * code of dt_graph_read_config_line per line (0 ok, > 0 warning, < 0 fatal) into codes[], and afterwards frame count and the
 * parameter block of every module as text ("<name>:<inst> <hex>"), so that what a line DID can be compared too. */
int ref_config_lines(const char *basedir, const char *cfgfile, const char *lines, int *codes, int maxcodes, char *out, int outsize)
{
  { const int r = ref_pipe_init(basedir); if(r) return r; }
  dt_graph_t *g = calloc(1, sizeof(*g));
  g->frame_cnt = 1;
  g->max_modules = 100; g->module = calloc(sizeof(dt_module_t), g->max_modules);
  g->max_nodes = 16;    g->node = calloc(sizeof(dt_node_t), g->max_nodes);
  g->params_max = 16u << 20; g->params_pool = calloc(1, g->params_max);
  int n = -20;
  if(!dt_graph_read_config_ascii(g, cfgfile))
  {
    n = 0;
    char *copy = strdup(lines), *c = copy;
    while(c && *c && n < maxcodes)
    {
      char *e = strchr(c, '\n'); if(e) *e++ = 0;
      static char line[300000];
      snprintf(line, sizeof(line), "%s", c);
      codes[n++] = dt_graph_read_config_line(g, line);
      c = e;
    }
    free(copy);
    char *o = out; int left = outsize; char b0[9], b1[9];
    ref_out(&o, &left, "frames %u\n", g->frame_cnt);
    for(uint32_t m = 0; m < g->num_modules; m++)
    {
      dt_module_t *mod = g->module + m;
      if(!mod->name) continue;
      for(int p = 0; p < mod->so->num_params; p++) if(mod->so->param[p]->type == dt_token("string"))
      {
        char *str = (char *)mod->param + mod->so->param[p]->offset;
        const int len = strnlen(str, mod->so->param[p]->cnt);
        memset(str + len, 0, mod->so->param[p]->cnt - len);
      }
      ref_out(&o, &left, "%s:%s ", ref_tkn(mod->name, b0), ref_tkn(mod->inst, b1));
      for(int k = 0; k < mod->param_size; k++) ref_out(&o, &left, "%02x", mod->param[k]);
      for(int cc = 0; cc < mod->num_connectors; cc++) if(dt_connector_input(mod->connector + cc))
        ref_out(&o, &left, " %s<%d.%d", ref_tkn(mod->connector[cc].name, b0), mod->connector[cc].connected.i, mod->connector[cc].connected.c);
      ref_out(&o, &left, "\n");
    }
    if(left <= 0) n = -1;
  }
  for(uint32_t m = 0; m < g->num_modules; m++) if(g->module[m].name && g->module[m].so && g->module[m].so->cleanup) g->module[m].so->cleanup(g->module + m);
  free(g->params_pool); free(g->node); free(g->module); free(g);
  return n;
}

/* the reference's own o-pfm write_sink (o-pfm/main.c:8-42) on a width x height rgba f32 buffer: writes <basename>.pfm */
void opfm_ref_write_sink(dt_module_t *, void *, dt_write_sink_params_t *);
int ref_write_pfm(const char *basename, const float *rgba, int width, int height)
{
  static dt_ui_param_t par;
  static dt_module_so_t so;
  dt_module_t *mod = calloc(1, sizeof(*mod));
  char name[256];
  snprintf(name, sizeof(name), "%s", basename);
  memset(&so, 0, sizeof(so)); memset(&par, 0, sizeof(par));
  par.name = dt_token("filename"); par.type = dt_token("string"); par.cnt = 256; par.offset = 0;
  so.param[0] = &par; so.num_params = 1;
  mod->so = &so; mod->param = (uint8_t *)name; mod->param_size = 256;
  mod->num_connectors = 1;
  mod->connector[0].roi.wd = mod->connector[0].roi.full_wd = width;
  mod->connector[0].roi.ht = mod->connector[0].roi.full_ht = height;
  opfm_ref_write_sink(mod, (void *)rgba, 0);
  free(mod);
  return 0;
}

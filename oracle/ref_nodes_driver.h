/* ORACLE — test infrastructure.  runs the HOST side of one of the reference's modules (modify_roi_out, modify_roi_in,
 * create_nodes of <module>/main.c, compiled in place from /root/reference by `make -C oracle ref`, never copied) on a
 * two module graph (a source standing in for whatever feeds the module, and the module) the way the reference's graph
 * passes do (graph-run-modules.h:200-330 roi out, :440-500 roi in, :30-110 create nodes), and writes the nodes the
 * module created as text: the same text the product writes for its own modules (vkb_graph_describe), so that the node
 * lists, dispatch sizes, push constants, connector formats and wiring can be compared line by line.
 * the module's connector and parameter tables are read from the reference's own `connectors` / `params` files. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>

typedef struct ref_nodes_in_t
{
  uint32_t in_full_wd, in_full_ht;  /* size of what the source hands to connector 0 */
  uint32_t filters;
  float    black[4], white[4], wb[4];
  uint32_t crop_aabb[4];
  float    noise_a, noise_b;
  const char *in_chan, *in_format;  /* channel / format of the source's output connector (wildcards of the module resolve to it) */
  const char *out_chan, *out_format; /* what the output connector ends up as after the consumer's roi pass negotiated it (graph-run-modules.h:495-548) */
  uint32_t out_marker;              /* ...and the strength of the consumer's size request (connector.h:64-75); the size asked for is the full size */
  const char *moddir;               /* <reference>/src/pipe/modules/<name> */
  const uint8_t *param; uint32_t param_size; /* the module's parameter block, laid out as the `params` file says */
} ref_nodes_in_t;

typedef void (*ref_cb_t)(dt_graph_t *, dt_module_t *);
typedef int  (*ref_init_t)(dt_module_t *);
typedef void (*ref_cleanup_t)(dt_module_t *);

static int ref_out(char **o, int *left, const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt);
  const int n = vsnprintf(*o, *left, fmt, ap);
  va_end(ap);
  if(n < 0 || n >= *left) { *left = 0; return 1; }
  *o += n; *left -= n;
  return 0;
}
static uint32_t ref_fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static const char *ref_tkn(dt_token_t t, char *b) { memcpy(b, &t, 8); b[8] = 0; return b; }

static int ref_nodes_run(const char *name, const ref_nodes_in_t *in, ref_init_t init, ref_cleanup_t cleanup, ref_cb_t roi_out, ref_cb_t roi_in, ref_cb_t create,
    char *out, int outsize)
{
  static dt_ui_param_t par[32];
  static dt_module_so_t so;
  dt_graph_t *graph = calloc(1, sizeof(*graph));
  graph->module = calloc(2, sizeof(dt_module_t)); graph->num_modules = graph->max_modules = 2;
  graph->node = calloc(256, sizeof(dt_node_t));   graph->max_nodes = 256;
  dt_module_t *src = graph->module, *mod = graph->module + 1;
  memset(&so, 0, sizeof(so));
  so.name = dt_token(name);
  char path[1024], line[4096];
  /* params: name:type:cnt:defaults (global.c:41-84), offsets packed back to back (:147-150) */
  snprintf(path, sizeof(path), "%s/params", in->moddir);
  FILE *f = fopen(path, "rb");
  int off = 0;
  if(f)
  {
    while(fgets(line, sizeof(line), f) && so.num_params < 32)
    {
      char *c0 = line, *c1 = strchr(c0, ':'); if(!c1) continue; *c1++ = 0;
      char *c2 = strchr(c1, ':'); if(!c2) continue; *c2++ = 0;
      dt_ui_param_t *p = par + so.num_params;
      memset(p, 0, sizeof(*p));
      p->name = dt_token(c0); p->type = dt_token(c1); p->cnt = atoi(c2); p->offset = off;
      off += p->cnt * (p->type == dt_token("string") ? 1 : 4);
      so.param[so.num_params++] = p;
    }
    fclose(f);
  }
  if((uint32_t)off != in->param_size) { free(graph->node); free(graph->module); free(graph); return -2; } /* the product lays its parameters out differently */
  uint8_t *params = malloc(off + 16);
  memcpy(params, in->param, off);
  mod->so = &so; mod->graph = graph; mod->name = so.name; mod->inst = dt_token("01"); mod->param = params; mod->param_size = off;
  src->graph = graph; src->name = dt_token("i-src"); src->inst = dt_token("main");
  src->num_connectors = 1;
  src->connector[0] = (dt_connector_t){ .name = dt_token("output"), .type = dt_token("source"), .chan = dt_token(in->in_chan), .format = dt_token(in->in_format) };
  src->connector[0].array_length = 1;
  src->connector[0].roi = (dt_roi_t){ .full_wd = in->in_full_wd, .full_ht = in->in_full_ht, .wd = in->in_full_wd, .ht = in->in_full_ht };
  for(int k = 0; k < 4; k++) { src->img_param.black[k] = in->black[k]; src->img_param.white[k] = in->white[k]; src->img_param.whitebalance[k] = in->wb[k]; src->img_param.crop_aabb[k] = in->crop_aabb[k]; }
  src->img_param.filters = in->filters; src->img_param.noise_a = in->noise_a; src->img_param.noise_b = in->noise_b;
  /* connectors: name:type:chan:format (global.c:27-38) */
  snprintf(path, sizeof(path), "%s/connectors", in->moddir);
  f = fopen(path, "rb");
  if(!f) { free(params); free(graph->node); free(graph->module); free(graph); return -3; }
  while(fgets(line, sizeof(line), f) && mod->num_connectors < DT_MAX_CONNECTORS)
  {
    char *tok[4] = {0}, *c = line;
    for(int k = 0; k < 4 && c; k++) { tok[k] = c; c = strpbrk(c, ":\n"); if(c) *c++ = 0; }
    if(!tok[3]) continue;
    dt_connector_t *cn = mod->connector + mod->num_connectors++;
    *cn = (dt_connector_t){ .name = dt_token(tok[0]), .type = dt_token(tok[1]), .chan = dt_token(tok[2]), .format = dt_token(tok[3]) };
    cn->connected = s_cid_unset; cn->associated = s_cid_unset; cn->bypass = s_cid_unset; cn->array_length = 1; /* module.c:73 */
    if(cn->type == dt_token("write")) cn->connected.i = cn->connected.c = 0;
  }
  fclose(f);
  /* the module layer connection source -> input, by the reference's own dt_module_connect (connector.c) */
  const int cerr = dt_module_connect(graph, 0, 0, 1, 0);
  if(cerr) { free(params); free(graph->node); free(graph->module); free(graph); return -100 - cerr; }
  if(init) init(mod);
  /* pass 1, roi out (graph-run-modules.h:200-330) */
  mod->img_param = src->img_param;
  for(int i = 0; i < mod->num_connectors; i++)
  {
    dt_connector_t *c = mod->connector + i;
    char b[9]; ref_tkn(c->chan, b);
    if(b[0] == '&') { const dt_token_t ref = c->chan >> 8; for(int j = 0; j < mod->num_connectors; j++) if(mod->connector[j].name == ref) { c->chan = mod->connector[j].chan; break; } }
    if(c->type == dt_token("write") && c->format == dt_token("*")) c->format = dt_token(in->out_format);
  }
  mod->connector[0].roi = src->connector[0].roi;
  if(roi_out) roi_out(graph, mod);
  else for(int i = 0; i < mod->num_connectors; i++) if(mod->connector[i].type == dt_token("write"))
  { mod->connector[i].roi.full_wd = mod->connector[0].roi.full_wd; mod->connector[i].roi.full_ht = mod->connector[0].roi.full_ht; }
  for(int i = 0; i < mod->num_connectors; i++) if(dt_connector_owner(mod->connector + i))
  { mod->connector[i].roi.wd = mod->connector[i].roi.full_wd; mod->connector[i].roi.ht = mod->connector[i].roi.full_ht; }
  char *o = out; int left = outsize; char b0[9], b1[9], b2[9], b3[9];
  const dt_image_params_t *ip = &mod->img_param;
  /* what modules further down copy in their own pass 1, before this module's create_nodes touches it; not in the product's text */
  ref_out(&o, &left, "imgout filters=%u black=%08x,%08x,%08x,%08x white=%08x,%08x,%08x,%08x wb=%08x,%08x,%08x,%08x crop=%u,%u,%u,%u noise=%08x,%08x\n", ip->filters,
      ref_fbits(ip->black[0]), ref_fbits(ip->black[1]), ref_fbits(ip->black[2]), ref_fbits(ip->black[3]),
      ref_fbits(ip->white[0]), ref_fbits(ip->white[1]), ref_fbits(ip->white[2]), ref_fbits(ip->white[3]),
      ref_fbits(ip->whitebalance[0]), ref_fbits(ip->whitebalance[1]), ref_fbits(ip->whitebalance[2]), ref_fbits(ip->whitebalance[3]),
      ip->crop_aabb[0], ip->crop_aabb[1], ip->crop_aabb[2], ip->crop_aabb[3], ref_fbits(ip->noise_a), ref_fbits(ip->noise_b));
  /* pass 2, roi in: the consumer asks for the full output (:440-500) */
  mod->connector[1].roi.marker = in->out_marker;
  mod->connector[1].chan = dt_token(in->out_chan);
  if(roi_in) roi_in(graph, mod);
  else for(int i = 0; i < mod->num_connectors; i++) if(dt_connector_input(mod->connector + i)) mod->connector[i].roi = mod->connector[1].roi;
  src->connector[0].roi = mod->connector[0].roi;
  /* pass 3, create nodes (:30-110): the module's own, every module here has one */
  create(graph, mod);

  ref_out(&o, &left, "module %s filters=%u black=%08x,%08x,%08x,%08x white=%08x,%08x,%08x,%08x wb=%08x,%08x,%08x,%08x crop=%u,%u,%u,%u noise=%08x,%08x\n", name, ip->filters,
      ref_fbits(ip->black[0]), ref_fbits(ip->black[1]), ref_fbits(ip->black[2]), ref_fbits(ip->black[3]),
      ref_fbits(ip->white[0]), ref_fbits(ip->white[1]), ref_fbits(ip->white[2]), ref_fbits(ip->white[3]),
      ref_fbits(ip->whitebalance[0]), ref_fbits(ip->whitebalance[1]), ref_fbits(ip->whitebalance[2]), ref_fbits(ip->whitebalance[3]),
      ip->crop_aabb[0], ip->crop_aabb[1], ip->crop_aabb[2], ip->crop_aabb[3], ref_fbits(ip->noise_a), ref_fbits(ip->noise_b));
  for(int i = 0; i < mod->num_connectors; i++)
  {
    const dt_connector_t *c = mod->connector + i;
    ref_out(&o, &left, " mconn %d %s:%s:%s:%s roi=%ux%u/%ux%u m=%u bypass=%d\n", i, ref_tkn(c->name, b0), ref_tkn(c->type, b1), ref_tkn(c->chan, b2), ref_tkn(c->format, b3),
        c->roi.full_wd, c->roi.full_ht, c->roi.wd, c->roi.ht, c->roi.marker, dt_cid_unset(c->bypass) ? -1 : c->bypass.c);
  }
  for(uint32_t n = 0; n < graph->num_nodes; n++)
  {
    const dt_node_t *nd = graph->node + n;
    ref_out(&o, &left, " node %u %s:%s %ux%ux%u pc=%d:", n, ref_tkn(nd->name, b0), ref_tkn(nd->kernel, b1), nd->wd, nd->ht, nd->dp, (int)nd->push_constant_size);
    for(size_t k = 0; k < nd->push_constant_size / 4; k++) ref_out(&o, &left, "%s%08x", k ? "," : "", nd->push_constant[k]);
    ref_out(&o, &left, "\n");
    for(int i = 0; i < nd->num_connectors; i++)
    {
      const dt_connector_t *c = nd->connector + i;
      ref_out(&o, &left, "  conn %d %s:%s:%s:%s roi=%ux%u/%ux%u al=%d ", i, ref_tkn(c->name, b0), ref_tkn(c->type, b1), ref_tkn(c->chan, b2), ref_tkn(c->format, b3),
          c->roi.full_wd, c->roi.full_ht, c->roi.wd, c->roi.ht, c->array_length);
      /* dt_node_add zero-fills `associated`, which reads as (module 0, connector 0): an output counts as the module's only when
       * the module connector points back at it (dt_connector_copy sets both directions, modules/api.h:87-121) */
      const int copied = !dt_cid_unset(c->associated) && (dt_connector_input(c) ||
          (c->associated.i == 1 && mod->connector[c->associated.c].associated.i == (int)n && mod->connector[c->associated.c].associated.c == i));
      if(copied)                                ref_out(&o, &left, "mod.%d\n", c->associated.c);
      else if(dt_connector_input(c) && c->connected.i >= 0) ref_out(&o, &left, "n%d.%d\n", c->connected.i, c->connected.c);
      else if(dt_connector_input(c))            ref_out(&o, &left, "open\n");
      else                                      ref_out(&o, &left, "own\n");
    }
  }
  if(cleanup) cleanup(mod);
  free(params); free(graph->node); free(graph->module); free(graph);
  return left > 0 ? (int)(o - out) : -1;
}

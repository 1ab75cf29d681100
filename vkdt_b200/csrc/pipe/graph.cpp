// graph model, config io and the module-level passes of dt_graph_run (topological order, ROI negotiation,
// node creation, bypass repointing).  semantics restated from src/pipe/module.c:7-110, connector.inc:1-130,
// graph-io.c:20-315, graph-export.c:23-106, graph-traverse.inc:12-152 and graph-run-modules.h:36-822.
#include "pipe.h"
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>

#define MAX_MODULES 100
#define MAX_NODES   4000

// ------------------------------------------------------------------------------------------------
// core/gaussian_elimination.h:43-112
int gauss_make_triangular(double *A, int *p, int n)
{
  p[n - 1] = n - 1;
  for(int k = 0; k < n; ++k)
  {
    int m = k;
    for(int i = k + 1; i < n; ++i) if(fabs(A[k + n * i]) > fabs(A[k + n * m])) m = i;
    p[k] = m;
    double t1 = A[k + n * m];
    A[k + n * m] = A[k + n * k];
    A[k + n * k] = t1;
    if(t1 == 0) return 0;
    for(int i = k + 1; i < n; ++i) A[k + n * i] /= -t1;
    if(k != m) for(int i = k + 1; i < n; ++i) std::swap(A[i + n * m], A[i + n * k]);
    for(int j = k + 1; j < n; ++j) for(int i = k + 1; i < n; ++i) A[i + n * j] += A[k + j * n] * A[i + k * n];
  }
  return 1;
}
void gauss_solve_triangular(const double *A, const int *p, double *b, int n)
{
  for(int k = 0; k < n - 1; ++k)
  {
    const int m = p[k];
    const double t = b[m];
    b[m] = b[k]; b[k] = t;
    for(int i = k + 1; i < n; ++i) b[i] += A[k + n * i] * t;
  }
  for(int k = n - 1; k > 0; --k)
  {
    b[k] /= A[k + n * k];
    const double t = b[k];
    for(int i = 0; i < k; ++i) b[i] -= A[k + n * i] * t;
  }
  b[0] /= A[0];
}
int gauss_solve(double *A, double *b, int n)
{
  std::vector<int> p(n);
  const int ok = gauss_make_triangular(A, p.data(), n);
  if(ok) gauss_solve_triangular(A, p.data(), b, n);
  return ok;
}

// ------------------------------------------------------------------------------------------------
dt_graph_t *dt_graph_new()
{ // graph.c:35-120: fixed capacity arrays, 16 MiB param pool
  dt_graph_t *g = new dt_graph_t();
  g->module.reserve(MAX_MODULES);
  g->node.reserve(MAX_NODES);
  g->params_pool.resize(16u << 20);
  g->params_end = 0;
  g->frame = 0; g->frame_cnt = 1; g->frame_rate = 24.0;
  memset(&g->main_img_param, 0, sizeof(g->main_img_param));
  g->runflags = 0;
  g->searchpath[0] = 0; g->basedir[0] = 0;
  g->plan = 0;
  g->device = 0;
  g->perf = 0;
  g->mode = vkb_default_mode();
  return g;
}

int dt_module_get(const dt_graph_t *g, dt_token_t name, dt_token_t inst)
{
  for(size_t i = 0; i < g->module.size(); i++) if(g->module[i].name == name && g->module[i].inst == inst) return (int)i;
  return -1;
}
int dt_module_get_connector(const dt_module_t *m, dt_token_t conn)
{
  for(int c = 0; c < m->num_connectors; c++) if(m->connector[c].name == conn) return c;
  return -1;
}
int dt_module_get_param(const dt_module_so_t *so, dt_token_t name)
{
  for(size_t i = 0; i < so->param.size(); i++) if(so->param[i].name == name) return (int)i;
  return -1;
}
int dt_module_set_param_float(dt_module_t *m, dt_token_t p, float v) { return dt_module_set_param_float_n(m, p, &v, 1); }
int dt_module_set_param_float_n(dt_module_t *m, dt_token_t p, const float *v, int n)
{
  const int id = dt_module_get_param(m->so, p);
  if(id < 0 || m->so->param[id].cnt < n) return 1;
  memcpy(m->param + m->so->param[id].offset, v, sizeof(float) * n);
  return 0;
}
int dt_module_set_param_string(dt_module_t *m, dt_token_t p, const char *str)
{
  const int id = dt_module_get_param(m->so, p);
  if(id < 0) return 1;
  snprintf((char *)m->param + m->so->param[id].offset, m->so->param[id].cnt, "%s", str);
  return 0;
}

// module.c:7-110
int dt_module_add(dt_graph_t *g, dt_token_t name, dt_token_t inst)
{
  const int ex = dt_module_get(g, name, inst);
  if(ex >= 0) return ex;
  dt_module_so_t *so = dt_module_so_get(name);
  if(!so) return -1;
  if(g->module.size() >= MAX_MODULES) return -1;
  g->module.emplace_back();
  dt_module_t *mod = &g->module.back();
  memset(mod, 0, sizeof(*mod));
  mod->so = so; mod->name = name; mod->inst = inst; mod->graph = g;
  mod->num_connectors = (int)so->connector.size();
  for(int c = 0; c < mod->num_connectors; c++) { mod->connector[c] = so->connector[c]; mod->connector[c].array_length = 1; } // module.c:73
  int psize = 0;
  for(const dt_ui_param_t &p : so->param) psize += (int)p.def.size();
  mod->param = g->params_pool.data() + g->params_end;
  mod->param_size = psize;
  g->params_end += (psize + 15) & ~15;
  for(const dt_ui_param_t &p : so->param) memcpy(mod->param + p.offset, p.def.data(), p.def.size());
  if(so->init) so->init(mod);
  if(mod->committed_param_size)
  {
    mod->committed_param = g->params_pool.data() + g->params_end;
    g->params_end += (mod->committed_param_size + 15) & ~15;
    memset(mod->committed_param, 0, mod->committed_param_size);
  }
  g->mem_source.resize(g->module.size());
  g->mem_sink.resize(g->module.size());
  return (int)g->module.size() - 1;
}

// connector.inc:1-130, shared by the module and the node layer
template <typename T, bool IS_MODULE>
static int connect_generic(std::vector<T> &el, int m0, int c0, int m1, int c1)
{
  const int num = (int)el.size();
  if(m1 < 0 || m1 >= num) return 1;
  if(c1 < 0 || c1 >= el[m1].num_connectors) return 2;
  dt_connector_t *cn1 = el[m1].connector + c1;
  if(cn1->connected.i == m0 && cn1->connected.c == c0) return 0;
  if(cn1->type != dt_token("read") && cn1->type != dt_token("sink") && cn1->type != dt_token("modify")) return 3;
  const int old_mod = cn1->connected.i;
  if(old_mod >= 0)
  {
    const int old_con = cn1->connected.c;
    if(old_mod >= num) return 4;
    cn1->connected.i = cn1->connected.c = -1;
    if constexpr(IS_MODULE)
    {
      cn1->format = el[m1].so->connector[c1].format;
      cn1->frames = el[m1].so->connector[c1].frames;
      cn1->chan   = el[m1].so->connector[c1].chan;
    }
    dt_connector_t *oc = el[old_mod].connector + old_con;
    if(oc->connected.i > 0)
    {
      oc->connected.i--;
      if constexpr(IS_MODULE) if(oc->connected.i == 0)
      {
        oc->format = el[old_mod].so->connector[old_con].format;
        oc->frames = el[old_mod].so->connector[old_con].frames;
        oc->chan   = el[old_mod].so->connector[old_con].chan;
        oc->flags &= ~s_conn_feedback;
      }
    }
    else return 6;
  }
  cn1->associated = s_cid_unset;
  if(c0 < 0 || m0 < 0) return 0;
  if(m0 >= num) return 7;
  if(c0 >= el[m0].num_connectors) return 8;
  el[m0].connector[c0].associated = s_cid_unset;
  dt_connector_t *cn0 = el[m0].connector + c0;
  if(cn0->type != dt_token("write") && cn0->type != dt_token("source") && cn0->type != dt_token("modify")) return 9;
  const int c0ref = ((const char *)&cn0->chan)[0] == '&';
  if(cn1->chan == dt_token("*") && !c0ref) cn1->chan = cn0->chan;
  if(cn0->chan == dt_token("*")) cn0->chan = cn1->chan;
  if(cn1->chan == dt_token("*")) cn1->chan = dt_token("rgba");
  if(cn0->chan == dt_token("*")) cn0->chan = dt_token("rgba");
  if(!c0ref && cn1->chan != cn0->chan) return 10;
  if(cn1->format == dt_token("*")) cn1->format = cn0->format;
  if(cn0->format == dt_token("*")) cn0->format = cn1->format;
  if(cn1->format == dt_token("*")) cn1->format = dt_token("f16");
  if(cn0->format == dt_token("*")) cn0->format = dt_token("f16");
  if(cn1->format != cn0->format) return 11;
  cn1->connected = dt_cid(m0, c0);
  cn1->array_length = cn0->array_length;
  cn1->flags = cn0->flags & ~s_conn_feedback;
  cn1->roi = cn0->roi;
  if(cn0->type == dt_token("write") || cn0->type == dt_token("source")) cn0->connected.i++;
  return 0;
}
// would feeding m1 from m0 close a loop?  (cycles.h:47-66: refused with code 12 before anything is touched.)  the graph so far
// has none, so only the new edge can: it does iff m0 already depends on m1, found by walking m0's inputs upstream
static bool connection_is_cyclic(const dt_graph_t *g, int m0, int m1)
{
  const int num = (int)g->module.size();
  if(m0 < 0 || m1 < 0 || m0 >= num || m1 >= num) return false;
  std::vector<char> seen(num, 0);
  std::vector<int> todo(1, m0);
  while(!todo.empty())
  {
    const int m = todo.back(); todo.pop_back();
    if(m == m1) return true;
    if(seen[m]) continue;
    seen[m] = 1;
    for(int c = 0; c < g->module[m].num_connectors; c++)
    {
      const dt_connector_t *cn = g->module[m].connector + c;
      if(!dt_connector_input(cn) || (cn->flags & s_conn_feedback)) continue; // feedback edges carry last frame's data: no cycle
      if(cn->connected.i >= 0 && cn->connected.i < num) todo.push_back(cn->connected.i);
    }
  }
  return false;
}
int dt_module_connect(dt_graph_t *g, int m0, int c0, int m1, int c1)
{
  if(connection_is_cyclic(g, m0, m1)) return 12;
  return connect_generic<dt_module_t, true>(g->module, m0, c0, m1, c1);
}
int dt_module_feedback(dt_graph_t *g, int m0, int c0, int m1, int c1)
{ // connector.c:30-38: connect without the cycle test (a feedback edge closes one on purpose) and flag the input
  const int err = connect_generic<dt_module_t, true>(g->module, m0, c0, m1, c1);
  if(!err && m1 >= 0 && c1 >= 0)
  { // both ends are double buffered: written in frame f, read through this edge in frame f + 1
    g->module[m1].connector[c1].flags |= s_conn_feedback;
    g->module[m1].connector[c1].frames = 2;
    if(m0 >= 0 && c0 >= 0) g->module[m0].connector[c0].frames = 2;
  }
  return err;
}
// the node layer (connector.c node flavour): wildcards on either side take the other side's channels / format
// (connector.inc:94-106); a mismatch of declared formats is not an error here, nodes declare what they read
int dt_node_connect(dt_graph_t *g, int n0, int c0, int n1, int c1)
{
  if(n1 < 0 || n1 >= (int)g->node.size() || n0 < 0 || n0 >= (int)g->node.size()) return 1;
  if(c1 < 0 || c1 >= g->node[n1].num_connectors) return 2;
  if(c0 < 0 || c0 >= g->node[n0].num_connectors) return 8;
  dt_connector_t *cn1 = g->node[n1].connector + c1, *cn0 = g->node[n0].connector + c0;
  if(!dt_connector_input(cn1)) return 3;
  if(!dt_connector_output(cn0)) return 9;
  if(cn1->chan == dt_token("*")) cn1->chan = cn0->chan;
  if(cn0->chan == dt_token("*")) cn0->chan = cn1->chan;
  if(cn1->chan == dt_token("*")) cn1->chan = dt_token("rgba");
  if(cn0->chan == dt_token("*")) cn0->chan = dt_token("rgba");
  if(cn1->format == dt_token("*")) cn1->format = cn0->format;
  if(cn0->format == dt_token("*")) cn0->format = cn1->format;
  if(cn1->format == dt_token("*")) cn1->format = dt_token("f16");
  if(cn0->format == dt_token("*")) cn0->format = dt_token("f16");
  cn1->connected = dt_cid(n0, c0);
  cn1->associated = s_cid_unset;
  cn0->associated = s_cid_unset;
  cn1->array_length = cn0->array_length;
  cn1->roi = cn0->roi;
  if(dt_connector_owner(cn0)) cn0->connected.i++;
  return 0;
}
int dt_node_connect_named(dt_graph_t *g, int n0, const char *c0, int n1, const char *c1)
{
  int i0 = -1, i1 = -1;
  for(int c = 0; c < g->node[n0].num_connectors; c++) if(g->node[n0].connector[c].name == dt_token(c0)) i0 = c;
  for(int c = 0; c < g->node[n1].num_connectors; c++) if(g->node[n1].connector[c].name == dt_token(c1)) i1 = c;
  if(i0 < 0) return -100;
  if(i1 < 0) return -101;
  return dt_node_connect(g, n0, i0, n1, i1);
}

int dt_module_remove(dt_graph_t *g, int modid)
{ // module.c dt_module_remove: disconnect everything, mark deleted (name = 0)
  if(modid < 0 || modid >= (int)g->module.size()) return 1;
  dt_module_t *m = &g->module[modid];
  for(int c = 0; c < m->num_connectors; c++)
  {
    if(dt_connector_input(m->connector + c)) dt_module_connect(g, -1, -1, modid, c);
    else for(size_t k = 0; k < g->module.size(); k++) for(int cc = 0; cc < g->module[k].num_connectors; cc++)
      if(dt_connector_input(g->module[k].connector + cc) && g->module[k].connector[cc].connected.i == modid && g->module[k].connector[cc].connected.c == c)
        dt_module_connect(g, -1, -1, (int)k, cc);
  }
  if(m->so->cleanup) m->so->cleanup(m);
  m->name = 0; m->inst = 0; m->num_connectors = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// modules/api.h
int dt_node_add(dt_graph_t *g, dt_module_t *m, const char *name, const char *kernel, int wd, int ht, int dp,
                int pc_size, const void *pc, int nc, ...)
{
  if(g->node.size() >= MAX_NODES) return -1;
  g->node.emplace_back();
  const int id = (int)g->node.size() - 1;
  dt_node_t *n = &g->node[id];
  memset(n, 0, sizeof(*n));
  n->name = dt_token(name); n->kernel = dt_token(kernel); n->module = m; n->num_connectors = nc;
  n->flags = m->flags; n->wd = wd; n->ht = ht; n->dp = dp; n->push_constant_size = pc_size;
  if(pc) memcpy(n->push_constant, pc, pc_size);
  va_list args;
  va_start(args, nc);
  for(int c = 0; c < nc; c++)
  {
    const char *t0 = va_arg(args, const char *), *t1 = va_arg(args, const char *);
    const char *t2 = va_arg(args, const char *), *t3 = va_arg(args, const char *);
    const dt_roi_t *roi = va_arg(args, const dt_roi_t *);
    dt_connector_t *cn = n->connector + c;
    cn->name = dt_token(t0); cn->type = dt_token(t1); cn->chan = dt_token(t2); cn->format = dt_token(t3);
    cn->associated = s_cid_unset; cn->bypass = s_cid_unset; cn->buf = -1;
    cn->connected = dt_cid(0, 0);
    if(dt_connector_owner(cn) && roi != dt_no_roi) cn->roi = *roi;
    if(dt_connector_input(cn)) cn->connected = s_cid_unset;
  }
  va_end(args);
  return id;
}
void dt_connector_copy(dt_graph_t *g, dt_module_t *m, int mc, int nid, int nc)
{ // api.h:87-120
  m->connector[mc].associated = dt_cid(nid, nc);
  dt_connector_t *c0 = m->connector + mc, *c1 = g->node[nid].connector + nc;
  if(c1->name == 0) *c1 = *c0;
  else
  {
    c1->frames = std::max(c1->frames, c0->frames);
    c1->flags  = c0->flags;
    c1->format = c0->format;
    if(!(c1->type == dt_token("write") && c1->roi.marker != s_roi_mark_uninited && c0->roi.marker == s_roi_mark_uninited))
      c1->roi = c0->roi;
    c1->connected = c0->connected;
    c1->array_length = c0->array_length;
  }
  c1->associated = dt_cid((int)(m - g->module.data()), mc);
}
void dt_connector_bypass(dt_graph_t *g, dt_module_t *m, int mc_in, int mc_out)
{
  m->connector[mc_out].bypass = dt_cid((int)(m - g->module.data()), mc_in);
  m->connector[mc_in].bypass  = dt_cid((int)(m - g->module.data()), mc_out);
}
const dt_image_params_t *dt_module_get_input_img_param(dt_graph_t *g, dt_module_t *m, dt_token_t input)
{
  const int c = dt_module_get_connector(m, input);
  if(c < 0) return 0;
  const dt_cid_t id = m->connector[c].connected;
  if(id.i < 0 || id.i >= (int)g->module.size()) return 0;
  return &g->module[id.i].img_param;
}

// ------------------------------------------------------------------------------------------------
// config io (graph-io.c:20-315, asciiio.h)
static dt_token_t io_token(char *&c)
{
  char b[9] = {0};
  int i = 0;
  while(i < 8 && *c && *c != ':' && *c != '\n') b[i++] = *c++;
  if(*c == ':' || *c == '\n') c++;
  return dt_token(b);
}
static int io_int(char *&c)
{
  if(!*c) return 0;
  char *e; const int r = (int)strtol(c, &e, 10);
  if(*e && e != c) e++;
  c = e; return r;
}
static float io_float(char *&c)
{
  if(!*c) return 0.0f;
  char *e; const float r = strtof(c, &e);
  if(*e && e != c) e++;
  c = e; return r;
}
static int read_param_values(dt_graph_t *g, char *line, dt_token_t name, dt_token_t inst, dt_token_t parm, int beg, int end, int mode, int frame = -1, int anim = 0)
{
  const int modid = dt_module_get(g, name, inst);
  if(modid < 0) { fprintf(stderr, "[vkdt_b200] no such module/instance %s/%s\n", dt_token_string(name).c_str(), dt_token_string(inst).c_str()); return 1; }
  dt_module_t *m = &g->module[modid];
  const int parid = dt_module_get_param(m->so, parm);
  if(parid < 0) return 2;
  const dt_ui_param_t *p = &m->so->param[parid];
  const int cnt = p->cnt;
  uint8_t *data = m->param + p->offset;
  if(beg < 0 || beg >= cnt || end < 0 || end > cnt) return 4;
  if(end == 0) end = cnt;
  if(frame >= 0)
  { // graph-io.c:56-76: the values go into a keyframe's own copy of the parameter instead (one per parameter and frame)
    dt_keyframe_t *kf = 0;
    for(dt_keyframe_t &k : m->keyframe) if(!kf && k.param == parm && (int)k.frame == frame) kf = &k;
    if(!kf) { m->keyframe.push_back(dt_keyframe_t()); kf = &m->keyframe.back(); }
    kf->frame = frame; kf->anim = anim; kf->param = parm; kf->beg = beg; kf->end = end;
    kf->data.assign(p->def.size(), 0);   // a fresh block from the (zeroed) parameter pool
    data = kf->data.data();
  }
  if(p->type == dt_token("float"))
  {
    float *block = (float *)data + beg;
    for(int i = beg; i < end; i++, block++)
    {
      const float v = io_float(line);
      *block = mode == 0 ? v : (mode == 1 ? *block + v : *block - v);
    }
  }
  else if(p->type == dt_token("int"))
  {
    int32_t *block = (int32_t *)data + beg;
    for(int i = beg; i < end; i++, block++)
    {
      const int v = io_int(line);
      *block = mode == 0 ? v : (mode == 1 ? *block + v : *block - v);
    }
  }
  else if(p->type == dt_token("string"))
  {
    char *str = (char *)data;
    int i = beg;
    do str[i++] = *(line++); while(line[0] && (i < end - 1));
    str[i] = 0;
  }
  return 0;
}
int dt_graph_read_config_line(dt_graph_t *g, char *c)
{ // graph-io.c:232-254
  if(c[0] == '#') return 0;
  const dt_token_t cmd = io_token(c); // an empty line is an unknown command like any other: warning 1
  if(cmd == dt_token("module"))
  {
    const dt_token_t name = io_token(c), inst = io_token(c);
    const float x = io_float(c), y = io_float(c);
    const int modid = dt_module_add(g, name, inst);
    if(modid < 0) { fprintf(stderr, "[vkdt_b200] failed to add module %s %s (not part of the raw->display path)\n", dt_token_string(name).c_str(), dt_token_string(inst).c_str()); return 1; }
    g->module[modid].gui_x = x; g->module[modid].gui_y = y;
    return 0;
  }
  if(cmd == dt_token("param"))
  {
    const dt_token_t name = io_token(c), inst = io_token(c), parm = io_token(c);
    return read_param_values(g, c, name, inst, parm, 0, 0, 0);
  }
  if(cmd == dt_token("paramsub") || cmd == dt_token("paraminc") || cmd == dt_token("paramdec"))
  {
    const dt_token_t name = io_token(c), inst = io_token(c), parm = io_token(c);
    const int beg = io_int(c), end = io_int(c);
    return read_param_values(g, c, name, inst, parm, beg, end, cmd == dt_token("paramsub") ? 0 : (cmd == dt_token("paraminc") ? 1 : 2));
  }
  if(cmd == dt_token("connect") || cmd == dt_token("feedback"))
  {
    const dt_token_t mod0 = io_token(c), inst0 = io_token(c), conn0 = io_token(c);
    const dt_token_t mod1 = io_token(c), inst1 = io_token(c), conn1 = io_token(c);
    int modid0 = dt_module_get(g, mod0, inst0);
    const int modid1 = dt_module_get(g, mod1, inst1);
    if((mod0 != dt_token("-1") && modid0 <= -1) || modid1 <= -1) return 1;
    int conid0 = -1;
    if(mod0 == dt_token("-1")) modid0 = -1;
    else conid0 = dt_module_get_connector(&g->module[modid0], conn0);
    const int conid1 = dt_module_get_connector(&g->module[modid1], conn1);
    const int err = cmd == dt_token("feedback") ? dt_module_feedback(g, modid0, conid0, modid1, conid1) : dt_module_connect(g, modid0, conid0, modid1, conid1);
    // graph-io.c:163-186: a feedback connection is an ordinary one flagged on its input; the traversal does not follow it
    // (graph-traverse.inc), the data comes from the frame before.  the executor refuses to run a graph that reaches one
    // (double buffered connectors are not built), frame parallel drivers fall back to one gpu (dt_graph_has_feedback)

    if(err) fprintf(stderr, "[vkdt_b200] connect %s:%s:%s -> %s:%s:%s failed: error %d\n", dt_token_string(mod0).c_str(), dt_token_string(inst0).c_str(),
        dt_token_string(conn0).c_str(), dt_token_string(mod1).c_str(), dt_token_string(inst1).c_str(), dt_token_string(conn1).c_str(), err);
    return err;
  }
  if(cmd == dt_token("frames")) { g->frame_cnt = atol(c); return 0; }
  if(cmd == dt_token("fps"))    { g->frame_rate = atof(c); return 0; }
  if(cmd == dt_token("keyframe") || cmd == dt_token("keyframE") || cmd == dt_token("Keyframe") || cmd == dt_token("KeyframE") || cmd == dt_token("keyFRAME"))
  { // graph-io.c:139-153, :243-247: frame:module:instance:param:beg:end:values, the spelling picks the easing
    const int anim = cmd == dt_token("keyframe") ? s_anim_lerp : cmd == dt_token("keyframE") ? s_anim_ease_out : cmd == dt_token("Keyframe") ? s_anim_ease_in :
                     cmd == dt_token("KeyframE") ? s_anim_smooth : s_anim_step;
    const int frame = io_int(c);
    const dt_token_t name = io_token(c), inst = io_token(c), parm = io_token(c);
    const int beg = io_int(c), end = io_int(c);
    return read_param_values(g, c, name, inst, parm, beg, end, 0, frame, anim);
  }
  return 1;
}

static float anim_warp(float t, int mode)
{ // anim.h:34-47
  switch(mode)
  {
    default:
    case s_anim_lerp:     return t;
    case s_anim_step:     return t > 0.5f ? 1.0f : 0.0f;
    case s_anim_ease_in:  return t * t * t;
    case s_anim_ease_out: { const float t1 = 1.0f - t; return 1.0f - t1 * t1 * t1; }
    case s_anim_smooth:   return 3.0f * t * t - 2.0f * t * t * t;
  }
}
// graph.c:1025-1100: for every parameter with keyframes take the latest one at or before the current frame (the earliest if
// all lie ahead) and, for floats, interpolate towards the next one with that keyframe's easing.  kept as the reference has it,
// including that the sub range values are read from the start of the keyframe's block and written to the start of the parameter
void dt_graph_apply_keyframes(dt_graph_t *g)
{
  for(dt_module_t &m : g->module)
  {
    if(!m.name || m.keyframe.empty()) continue;
    std::vector<dt_keyframe_t> &kf = m.keyframe;
    for(const dt_ui_param_t &p : m.so->param)
    {
      int ki = -1, kiM = -1;
      for(int i = 0; i < (int)kf.size(); i++) if(kf[i].param == p.name)
      {
        if(ki == -1) ki = i;
        else if(kf[ki].frame >  (int)g->frame && kf[i].frame < kf[ki].frame) ki = i;
        else if(kf[ki].frame <= (int)g->frame && kf[i].frame <= (int)g->frame && kf[i].frame > kf[ki].frame) ki = i;
      }
      if(ki == -1) continue;
      for(int i = 0; i < (int)kf.size(); i++) if(kf[i].param == p.name)
        if(kf[i].frame > (int)g->frame && (kiM == -1 || kf[kiM].frame > kf[i].frame)) kiM = i;
      if(kiM == ki) kiM = -1;
      uint8_t *pdat = m.param + p.offset;
      const uint8_t *fdat = kf[ki].data.data();
      const size_t els = p.type == dt_token("string") ? 1 : 4;
      if(kiM >= 0 && p.type == dt_token("float"))
      {
        const float t = anim_warp(((int)g->frame - kf[ki].frame) / (float)(kf[kiM].frame - kf[ki].frame), kf[kiM].anim);
        float *dst = (float *)pdat; const float *src0 = (const float *)fdat, *src1 = (const float *)kf[kiM].data.data();
        for(int i = kf[ki].beg; i < kf[ki].end; i++) dst[i] = t * src1[i - kf[ki].beg] + (1.0f - t) * src0[i - kf[ki].beg];
      }
      else memcpy(pdat, fdat + els * kf[ki].beg, els * (kf[ki].end - kf[ki].beg));
    }
  }
}
int dt_graph_has_feedback(const dt_graph_t *g)
{
  for(const dt_module_t &m : g->module) if(m.name) for(int c = 0; c < m.num_connectors; c++) if(m.connector[c].flags & s_conn_feedback) return 1;
  return 0;
}
int dt_graph_read_config_ascii(dt_graph_t *g, const char *filename)
{
  FILE *f = fopen(filename, "rb");
  if(!f) return 1;
  snprintf(g->searchpath, sizeof(g->searchpath), "%s", filename);
  char *slash = strrchr(g->searchpath, '/');
  if(slash) *slash = 0; else snprintf(g->searchpath, sizeof(g->searchpath), ".");
  std::vector<char> line(300000);
  uint32_t lno = 0;
  while(fgets(line.data(), (int)line.size(), f))
  {
    lno++;
    size_t n = strlen(line.data());
    while(n && (line[n-1] == '\n' || line[n-1] == '\r')) line[--n] = 0;
    if(dt_graph_read_config_line(g, line.data()) < 0)
    {
      fprintf(stderr, "[vkdt_b200] failed in line %u: '%s'\n", lno, line.data());
      fclose(f);
      return 1;
    }
  }
  fclose(f);
  return 0;
}

// graph-export.c:23-96
int dt_graph_replace_display(dt_graph_t *g, dt_token_t inst, dt_token_t mod, int prim, int trc, int resize, int max_wd, int max_ht)
{
  if(inst == 0) inst = dt_token("main");
  const int mid = dt_module_get(g, dt_token("display"), inst);
  if(mid < 0) return -1;
  const int cid = dt_module_get_connector(&g->module[mid], dt_token("input"));
  int m0 = g->module[mid].connector[cid].connected.i, o0 = g->module[mid].connector[cid].connected.c;
  if(m0 < 0) return -2;
  if(mod == 0) mod = dt_token("o-pfm");
  if(resize)
  { // :54-62: a resize module between the graph and the output
    const int m1 = dt_module_add(g, dt_token("resize"), inst);
    if(m1 < 0) return -3;
    if(dt_module_connect(g, m0, o0, m1, 0)) return -4;
    m0 = m1; o0 = 1;
  }
  const int m2 = dt_module_add(g, mod, inst);
  if(m2 < 0) return -3;
  const int i2 = dt_module_get_connector(&g->module[m2], dt_token("input"));
  g->module[m2].connector[i2].max_wd = max_wd;   // :93-94
  g->module[m2].connector[i2].max_ht = max_ht;
  if(g->module[m2].connector[i2].format == dt_token("ui8") || prim != 2 || trc != 0)
  { // :66-86: 8 bit sinks and other colour spaces than linear bt2020 get a colenc module in front
    const int m1 = dt_module_add(g, dt_token("colenc"), inst);
    if(m1 < 0) return -3;
    dt_module_t *ce = &g->module[m1];
    const int i1 = dt_module_get_connector(ce, dt_token("input")), o1 = dt_module_get_connector(ce, dt_token("output"));
    if(prim == 0xffff) prim = g->module[m0].img_param.colour_primaries;
    if(trc  == 0xffff) trc  = g->module[m0].img_param.colour_trc;
    *(int32_t *)dt_module_param_int(ce, dt_module_get_param(ce->so, dt_token("prim"))) = prim;
    *(int32_t *)dt_module_param_int(ce, dt_module_get_param(ce->so, dt_token("trc")))  = trc;
    ce->connector[o1].format = g->module[m2].connector[i2].format;
    ce->connector[o1].chan   = g->module[m2].connector[i2].chan;
    if(dt_module_connect(g, m0, o0, m1, i1)) return -4;
    if(dt_module_connect(g, m1, o1, m2, i2)) return -4;
    return m2;
  }
  if(g->module[m2].connector[i2].format != dt_token("*")) g->module[m0].connector[o0].format = g->module[m2].connector[i2].format;
  if(g->module[m2].connector[i2].chan != dt_token("*"))   g->module[m0].connector[o0].chan   = g->module[m2].connector[i2].chan;
  const int err = dt_module_connect(g, m0, o0, m2, i2);
  if(err) return -4;
  return m2;
}
void dt_graph_disconnect_display_modules(dt_graph_t *g)
{
  for(size_t m = 0; m < g->module.size(); m++) if(g->module[m].name == dt_token("display")) dt_module_remove(g, (int)m);
}

// ------------------------------------------------------------------------------------------------
// graph-traverse.inc:12-152: iterative post-order dfs from the sinks, "main" instances first
template <typename T, typename F>
static void traverse_post(std::vector<T> &arr, F is_main, std::vector<int> &order)
{
  const int cnt = (int)arr.size();
  std::vector<uint8_t> mark(cnt, 0);
  std::vector<int> stack; std::vector<uint8_t> done;
  for(int i = 0; i < cnt; i++)
    if(arr[i].num_connectors && arr[i].connector[0].type == dt_token("sink") && dt_connected(arr[i].connector) && is_main(arr[i]))
    { mark[i] = 1; stack.push_back(i); done.push_back(0); }
  for(int i = cnt - 1; i >= 0; i--)
    if(arr[i].num_connectors && arr[i].connector[0].type == dt_token("sink") && dt_connected(arr[i].connector) && !mark[i])
    { mark[i] = 1; stack.push_back(i); done.push_back(0); }
  // graph-traverse.inc:71-148: elements only reached through a feedback edge are not dependencies of this frame; they are
  // collected and traversed in a second round (they still have to run: the next frame reads what they write)
  std::vector<int> feedback_stack;
  for(;;)
  {
  while(!stack.empty())
  {
    const int curr = stack.back();
    if(mark[curr] == 3) { stack.pop_back(); done.pop_back(); }
    else if(done.back())
    {
      order.push_back(curr);
      mark[curr] = 3;
      stack.pop_back(); done.pop_back();
    }
    else
    {
      if(mark[curr] < 2) mark[curr] = 2;
      done.back() = 1;
      for(int i = 0; i < arr[curr].num_connectors; i++)
      {
        const int el = arr[curr].connector[i].connected.i;
        if(el < 0 || el >= cnt) continue;
        if(!dt_connector_input(arr[curr].connector + i)) continue;
        if(mark[el] != 3 && !(arr[curr].connector[i].flags & s_conn_feedback))
        {
          if(stack.size() > 100000) { order.clear(); return; } // cyclic
          stack.push_back(el); done.push_back(0);
          if(mark[el] < 1) mark[el] = 1;
        }
        if(mark[el] != 3 && (arr[curr].connector[i].flags & s_conn_feedback))
        {
          if(feedback_stack.size() > 100000) { order.clear(); return; }
          feedback_stack.push_back(el);
          mark[el] = 1;
        }
      }
    }
  }
  if(feedback_stack.empty()) break;
  stack = feedback_stack;
  done.assign(stack.size(), 0);
  feedback_stack.clear();
  }
}

// ------------------------------------------------------------------------------------------------
// graph-run-modules.h:200-337
static void modify_roi_out(dt_graph_t *g, dt_module_t *module)
{
  int input = dt_module_get_connector(module, dt_token("input"));
  dt_connector_t *c = 0;
  if(input >= 0 && module->connector[input].connected.i >= 0 &&
     g->module[module->connector[input].connected.i].img_param.input_name != dt_token("main"))
    for(int i = 0; i < module->num_connectors; i++)
    {
      if(!dt_connector_input(module->connector + i)) continue;
      const int mid = module->connector[i].connected.i;
      if(mid < 0) continue;
      if(g->module[mid].img_param.input_name == dt_token("main")) input = i;
    }
  const std::string nm = dt_token_string(module->name);
  if(module->inst != dt_token("main") || nm.compare(0, 2, "i-")) memset(&module->img_param, 0, sizeof(module->img_param));
  if(input >= 0)
  {
    c = module->connector + input;
    if(c->connected.i != -1) module->img_param = g->module[c->connected.i].img_param;
  }
  for(int i = 0; i < module->num_connectors; i++)
  { // &input style channel references
    dt_connector_t *cc = module->connector + i;
    if(((const char *)&cc->chan)[0] == '&')
    {
      const dt_token_t tmp = cc->chan >> 8;
      for(int j = 0; j < module->num_connectors; j++) if(module->connector[j].name == tmp) { cc->chan = module->connector[j].chan; break; }
    }
  }
  for(int i = 0; i < module->num_connectors; i++)
  {
    dt_connector_t *cc = module->connector + i;
    if(dt_connector_input(cc) && cc->connected.i >= 0 && cc->connected.c >= 0)
      cc->roi = g->module[cc->connected.i].connector[cc->connected.c].roi;
  }
  if(!module->disabled && module->so->modify_roi_out)
  {
    for(int i = 0; i < module->num_connectors; i++) if(dt_connector_owner(module->connector + i)) module->connector[i].roi.marker = s_roi_mark_uninited;
    module->so->modify_roi_out(g, module);
    for(int i = 0; i < module->num_connectors; i++)
    {
      if(module->connector[i].type == dt_token("source") && module->connector[i].roi.marker == s_roi_mark_uninited)
        module->connector[i].roi.marker = s_roi_mark_hard_fwd;
      else if(module->connector[i].roi.marker == s_roi_mark_uninited)
        module->connector[i].roi.marker = s_roi_mark_soft_fwd;
    }
    if(module->inst == dt_token("main") && dt_connector_owner(module->connector)) g->main_img_param = module->img_param;
    if(module->connector[0].type == dt_token("source")) module->img_param.input_name = module->inst;
  }
  else
  {
    for(int i = 0; i < module->num_connectors; i++) if(dt_connector_owner(module->connector + i)) module->connector[i].roi.marker = s_roi_mark_uninited;
    if(module->connector[0].type == dt_token("source")) module->img_param.input_name = module->inst;
    dt_roi_t roi = {0, 0, 0, 0, s_roi_mark_soft_fwd};
    if(input < 0) { roi.full_wd = 1024; roi.full_ht = 1024; }
    else if(c->connected.i != -1)
    {
      roi = g->module[c->connected.i].connector[c->connected.c].roi;
      c->roi = roi;
    }
    for(int i = 0; i < module->num_connectors; i++) if(module->connector[i].type == dt_token("write"))
    {
      module->connector[i].roi.full_wd = roi.full_wd;
      module->connector[i].roi.full_ht = roi.full_ht;
      module->connector[i].roi.marker = s_roi_mark_soft_fwd;
    }
  }
  for(int i = 0; i < module->num_connectors; i++)
  {
    dt_connector_t *cc = module->connector + i;
    if(dt_connector_owner(cc)) { cc->roi.wd = cc->roi.full_wd; cc->roi.ht = cc->roi.full_ht; }
  }
}

// graph-run-modules.h:338-447
static void propagate_roi_in(dt_graph_t *g, dt_module_t *module)
{
  for(int i = 0; i < module->num_connectors; i++)
  {
    dt_connector_t *c = module->connector + i;
    if(!dt_connector_input(c) || dt_cid_unset(c->connected)) continue;
    dt_connector_t *src = &g->module[c->connected.i].connector[c->connected.c];
    dt_roi_t *roi = &src->roi;
    if(roi->wd == 0)
    {
      if(roi->full_wd == 0) roi->full_wd = 2;
      if(roi->full_ht == 0) roi->full_ht = 2;
      roi->wd = roi->full_wd; roi->ht = roi->full_ht;
      if(roi->marker == s_roi_mark_uninited) roi->marker = s_roi_mark_soft_fwd;
    }
    if(!dt_roi_stronger(&c->roi, roi)) { c->roi = *roi; c->format = src->format; c->chan = src->chan; }
    else { *roi = c->roi; src->format = c->format; src->chan = c->chan; }
    c->array_length = src->array_length;
  }
  const std::string nm = dt_token_string(module->name);
  if(!nm.compare(0, 2, "i-"))
  {
    for(int i = 0; i < module->num_connectors; i++) if(dt_connector_owner(module->connector + i))
    { module->connector[i].roi.wd = module->connector[i].roi.full_wd; module->connector[i].roi.ht = module->connector[i].roi.full_ht; }
  }
  else if(!module->disabled && module->so->modify_roi_out)
  {
    dt_module_t tmp = *module;
    for(int i = 0; i < module->num_connectors; i++)
    { tmp.connector[i].roi.full_wd = tmp.connector[i].roi.wd; tmp.connector[i].roi.full_ht = tmp.connector[i].roi.ht; }
    module->so->modify_roi_out(g, &tmp);
    for(int i = 0; i < module->num_connectors; i++) if(dt_connector_owner(module->connector + i))
    {
      module->connector[i].roi.wd = tmp.connector[i].roi.full_wd;
      module->connector[i].roi.ht = tmp.connector[i].roi.full_ht;
      module->connector[i].roi.marker = tmp.connector[i].roi.marker;
    }
  }
  else
  {
    const int input = dt_module_get_connector(module, dt_token("input"));
    dt_roi_t roi = {0, 0, 0, 0, 0};
    if(input >= 0) roi = module->connector[input].roi;
    for(int i = 0; i < module->num_connectors; i++) if(module->connector[i].type == dt_token("write"))
    { module->connector[i].roi.wd = roi.wd; module->connector[i].roi.ht = roi.ht; }
  }
}

// graph-run-modules.h:449-545
static void modify_roi_in(dt_graph_t *g, dt_module_t *module)
{
  if(!module->disabled && module->so->modify_roi_in) module->so->modify_roi_in(g, module);
  else
  {
    int output = dt_module_get_connector(module, dt_token("output"));
    if(output == -1 && module->connector[0].type == dt_token("sink"))
    {
      output = 0;
      dt_roi_t *r = &module->connector[0].roi;
      const float max_wd = module->connector[0].max_wd, max_ht = module->connector[0].max_ht;
      const float scalex = max_wd > 0 ? r->full_wd / (float)max_wd : 1.0f;
      const float scaley = max_ht > 0 ? r->full_ht / (float)max_ht : 1.0f;
      const float scale = std::max(scalex, scaley);
      r->wd = (uint32_t)(r->full_wd / scale + 0.5f);
      r->ht = (uint32_t)(r->full_ht / scale + 0.5f);
      r->marker = module->inst == dt_token("main") ? s_roi_mark_hard_bck : s_roi_mark_soft_bck;
    }
    if(output < 0) return;
    const dt_roi_t roi = module->connector[output].roi;
    for(int i = 0; i < module->num_connectors; i++) if(dt_connector_input(module->connector + i)) module->connector[i].roi = roi;
  }
  for(int i = 0; i < module->num_connectors; i++)
  {
    dt_connector_t *c = module->connector + i;
    if(!dt_connector_input(c) || dt_cid_unset(c->connected)) continue;
    dt_connector_t *src = &g->module[c->connected.i].connector[c->connected.c];
    dt_roi_t *roi = &src->roi;
    if(src->type == dt_token("source"))
    {
      roi->wd = roi->full_wd; roi->ht = roi->full_ht;
      if(roi->marker == s_roi_mark_uninited) roi->marker = s_roi_mark_hard_bck;
      c->roi = *roi;
    }
    if(!dt_roi_stronger(&c->roi, roi)) { c->roi = *roi; c->format = src->format; c->chan = src->chan; }
    else { *roi = c->roi; src->format = c->format; src->chan = c->chan; }
    src->flags |= c->flags;
    c->array_length = src->array_length;
  }
}

// graph-run-modules.h:36-110: default node = copy of the module
static void create_nodes(dt_graph_t *g, dt_module_t *module)
{
  for(int i = 0; i < module->num_connectors; i++) module->connector[i].bypass = s_cid_unset;
  if(module->disabled)
  {
    int mc_in = -1, mc_out = -1;
    for(int i = 0; i < module->num_connectors; i++)
    {
      if(dt_connector_output(module->connector + i) && module->connector[i].name == dt_token("output")) mc_out = i;
      if(dt_connector_input(module->connector + i) && module->connector[i].name == dt_token("input")) mc_in = i;
    }
    if(mc_in >= 0 && mc_out >= 0) dt_connector_bypass(g, module, mc_in, mc_out);
  }
  else if(module->so->create_nodes) module->so->create_nodes(g, module);
  else
  {
    g->node.emplace_back();
    const int nodeid = (int)g->node.size() - 1;
    dt_node_t *node = &g->node[nodeid];
    memset(node, 0, sizeof(*node));
    node->name = module->name; node->kernel = dt_token("main");
    node->num_connectors = module->num_connectors; node->module = module; node->flags = module->flags;
    const int output = dt_module_get_connector(module, dt_token("output"));
    if(output >= 0) { node->wd = module->connector[output].roi.wd; node->ht = module->connector[output].roi.ht; node->dp = 1; }
    for(int i = 0; i < module->num_connectors; i++) dt_connector_copy(g, module, i, nodeid, i);
  }
}

// the module layer portion of dt_graph_run (graph-run-modules.h:549-822)
int dt_graph_run_modules(dt_graph_t *g, std::vector<int> &modid)
{
  modid.clear();
  traverse_post(g->module, [](const dt_module_t &m) { return m.inst == dt_token("main"); }, modid);
  if(modid.empty()) return vkb_set_error(VKB_ERR_GRAPH, "no module reaches a main sink (display:main / o-*:main): nothing to run, or the connections form a cycle");
  const int cnt = (int)modid.size();
  int main_input_module = -1;
  for(int i = 0; i < cnt; i++)
  {
    const std::string nm = dt_token_string(g->module[modid[i]].name);
    if(!nm.compare(0, 2, "i-") && (main_input_module == -1 || g->module[modid[i]].inst == dt_token("main"))) main_input_module = modid[i];
  }
  // pass 1: roi out, source -> sink
  for(int i = 0; i < cnt; i++) for(int c = 0; c < g->module[modid[i]].num_connectors; c++)
    memset(&g->module[modid[i]].connector[c].roi, 0, sizeof(dt_roi_t));
  if(main_input_module >= 0) modify_roi_out(g, &g->module[main_input_module]);
  for(int i = 0; i < cnt; i++) modify_roi_out(g, &g->module[modid[i]]);
  for(int i = 0; i < cnt; i++) if(g->module[modid[i]].connector[0].roi.full_wd == 0) modify_roi_out(g, &g->module[modid[i]]);
  for(int i = cnt - 1; i >= 0; i--) if(g->module[modid[i]].connector[0].type == dt_token("sink"))
  {
    if(g->module[modid[i]].connector[0].roi.full_wd == 0)
      return vkb_set_error(VKB_ERR_GRAPH, "module %s %s connector %s has uninited size", dt_token_string(g->module[modid[i]].name).c_str(),
          dt_token_string(g->module[modid[i]].inst).c_str(), dt_token_string(g->module[modid[i]].connector[0].name).c_str());
    break;
  }
  // pass 2: roi in (sink -> source), propagate (source -> sink), create nodes
  g->node.clear();
  for(int i = 0; i < cnt; i++) for(int j = 0; j < g->module[modid[i]].num_connectors; j++) g->module[modid[i]].connector[j].associated = s_cid_unset;
  for(int i = cnt - 1; i >= 0; i--) if(g->module[modid[i]].connector[0].roi.full_wd > 0) modify_roi_in(g, &g->module[modid[i]]);
  for(int i = 0; i < cnt; i++) propagate_roi_in(g, &g->module[modid[i]]);
  for(int i = 0; i < cnt; i++) if(g->module[modid[i]].connector[0].roi.full_wd > 0) create_nodes(g, &g->module[modid[i]]);
  // repoint node connectors through the module layer, following bypass chains (:766-812)
  for(size_t ni = 0; ni < g->node.size(); ni++)
  {
    dt_node_t *n = &g->node[ni];
    for(int i = 0; i < n->num_connectors; i++)
    {
      if(dt_cid_unset(n->connector[i].associated)) continue;
      dt_cid_t id;
      const dt_cid_t m0 = n->connector[i].associated;
      if(dt_connector_input(n->connector + i))
      {
        dt_cid_t m1 = g->module[m0.i].connector[m0.c].connected;
        if(dt_cid_unset(m1) || m1.i < 0) { n->connector[i].connected = s_cid_unset; continue; }
        int guard = 0;
        while(!dt_cid_unset(g->module[m1.i].connector[m1.c].bypass) && guard++ < 1000)
        {
          const dt_cid_t m2 = g->module[m1.i].connector[m1.c].bypass;
          const dt_cid_t m3 = g->module[m2.i].connector[m2.c].connected;
          if(dt_cid_unset(m3) || m3.i < 0) break;
          m1 = m3;
        }
        id = g->module[m1.i].connector[m1.c].associated;
      }
      else id = g->module[m0.i].connector[m0.c].associated;
      n->connector[i].connected = id;
    }
  }
  if(main_input_module >= 0 && g->module[main_input_module].connector[0].roi.wd == 0) return vkb_set_error(VKB_ERR_GRAPH, "main input has zero size");
  return VKB_OK;
}

// node-level post order (graph.c:753-768)
void dt_graph_node_order(dt_graph_t *g, std::vector<int> &nodeid)
{
  nodeid.clear();
  traverse_post(g->node, [](const dt_node_t &n) { return n.module->inst == dt_token("main"); }, nodeid);
}

std::string dt_graph_dump_nodes(dt_graph_t *g)
{ // graph-print.h:76: graphviz of the node layer
  std::string s = "digraph nodes {\nnode [shape=record]\n";
  char b[512];
  for(size_t i = 0; i < g->node.size(); i++)
  {
    const dt_node_t *n = &g->node[i];
    snprintf(b, sizeof(b), "n%zu [label=\"%s_%s|%ux%ux%u\"];\n", i, dt_token_string(n->name).c_str(), dt_token_string(n->kernel).c_str(), n->wd, n->ht, n->dp);
    s += b;
  }
  for(size_t i = 0; i < g->node.size(); i++) for(int c = 0; c < g->node[i].num_connectors; c++)
  {
    const dt_connector_t *cn = g->node[i].connector + c;
    if(dt_connector_input(cn) && cn->connected.i >= 0)
    {
      snprintf(b, sizeof(b), "n%d -> n%zu [label=\"%s\"];\n", cn->connected.i, i, dt_token_string(cn->name).c_str());
      s += b;
    }
  }
  s += "}\n";
  return s;
}

// the module and node layer as text, one block per module on the path in execution order: image parameters as they leave
// the module, its connectors, and every node it created with dispatch size, push constants and the connectors' formats,
// sizes and wiring (node ids local to the module; "mod.c": copied from module connector c, "nK.c": node connection).
// what create_nodes of the reference builds for the same module is written in the same form by oracle/ref_nodes_driver.h,
// and tests/test_host_ref_cpu.py compares the two line by line.  also lists each module's parameter block.
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
std::string dt_graph_describe(dt_graph_t *g, const std::vector<int> &modid)
{
  std::string s;
  char b[1024];
  for(int mi : modid)
  {
    const dt_module_t *m = &g->module[mi];
    const dt_image_params_t *ip = &m->img_param;
    snprintf(b, sizeof(b), "module %s filters=%u black=%08x,%08x,%08x,%08x white=%08x,%08x,%08x,%08x wb=%08x,%08x,%08x,%08x crop=%u,%u,%u,%u noise=%08x,%08x\n",
        dt_token_string(m->name).c_str(), ip->filters, f2u(ip->black[0]), f2u(ip->black[1]), f2u(ip->black[2]), f2u(ip->black[3]),
        f2u(ip->white[0]), f2u(ip->white[1]), f2u(ip->white[2]), f2u(ip->white[3]),
        f2u(ip->whitebalance[0]), f2u(ip->whitebalance[1]), f2u(ip->whitebalance[2]), f2u(ip->whitebalance[3]),
        ip->crop_aabb[0], ip->crop_aabb[1], ip->crop_aabb[2], ip->crop_aabb[3], f2u(ip->noise_a), f2u(ip->noise_b));
    s += b;
    s += " params ";
    for(int k = 0; k < m->param_size; k++) { snprintf(b, sizeof(b), "%02x", m->param[k]); s += b; }
    s += "\n";
    for(int i = 0; i < m->num_connectors; i++)
    {
      const dt_connector_t *c = m->connector + i;
      snprintf(b, sizeof(b), " mconn %d %s:%s:%s:%s roi=%ux%u/%ux%u m=%u bypass=%d\n", i, dt_token_string(c->name).c_str(), dt_token_string(c->type).c_str(),
          dt_token_string(c->chan).c_str(), dt_token_string(c->format).c_str(), c->roi.full_wd, c->roi.full_ht, c->roi.wd, c->roi.ht, c->roi.marker, dt_cid_unset(c->bypass) ? -1 : c->bypass.c);
      s += b;
    }
    int first = -1;
    for(size_t n = 0; n < g->node.size(); n++) if(g->node[n].module == m) { first = (int)n; break; }
    for(size_t n = 0; n < g->node.size(); n++)
    {
      const dt_node_t *nd = &g->node[n];
      if(nd->module != m) continue;
      snprintf(b, sizeof(b), " node %d %s:%s %ux%ux%u pc=%d:", (int)n - first, dt_token_string(nd->name).c_str(), dt_token_string(nd->kernel).c_str(), nd->wd, nd->ht, nd->dp, (int)nd->push_constant_size);
      s += b;
      for(size_t k = 0; k < nd->push_constant_size / 4; k++) { uint32_t w; memcpy(&w, nd->push_constant + 4 * k, 4); snprintf(b, sizeof(b), "%s%08x", k ? "," : "", w); s += b; }
      s += "\n";
      for(int i = 0; i < nd->num_connectors; i++)
      {
        const dt_connector_t *c = nd->connector + i;
        snprintf(b, sizeof(b), "  conn %d %s:%s:%s:%s roi=%ux%u/%ux%u al=%d ", i, dt_token_string(c->name).c_str(), dt_token_string(c->type).c_str(),
            dt_token_string(c->chan).c_str(), dt_token_string(c->format).c_str(), c->roi.full_wd, c->roi.full_ht, c->roi.wd, c->roi.ht, c->array_length);
        s += b;
        const bool copied = !dt_cid_unset(c->associated) && (dt_connector_input(c) ||
            (m->connector[c->associated.c].associated.i == (int)n && m->connector[c->associated.c].associated.c == i));
        if(copied) snprintf(b, sizeof(b), "mod.%d\n", c->associated.c);
        else if(dt_connector_input(c) && c->connected.i >= 0) snprintf(b, sizeof(b), "n%d.%d\n", c->connected.i - first, c->connected.c);
        else if(dt_connector_input(c)) snprintf(b, sizeof(b), "open\n");
        else snprintf(b, sizeof(b), "own\n");
        s += b;
      }
    }
    if(m->so->commit_params && m->committed_param_size)
    { // the uniform block the kernels of this module are handed
      dt_module_t *mw = &g->module[mi];
      mw->so->commit_params(g, mw);
      s += " committed ";
      for(int k = 0; k < mw->committed_param_size; k++) { snprintf(b, sizeof(b), "%02x", mw->committed_param[k]); s += b; }
      s += "\n";
    }
  }
  return s;
}

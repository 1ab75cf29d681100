// the graph executor: replaces dt_graph_run's node half (src/pipe/graph.c:719-936, graph-run-nodes-allocate.h,
// graph-run-nodes-upload.h, graph-run-nodes-record-cmd.h, graph-run-nodes-download.h) and the sub-allocator
// (src/pipe/alloc.c) with a B200 schedule:
//   node DAG -> rewrite pass (dead nodes dropped, identity resample aliased, llap restructured, pointwise chains
//   and unpack+noop fused) -> launch list -> liveness analysis -> offsets into ONE pooled HBM allocation
//   (buffers whose lifetimes do not overlap share memory, same idea as nid_last_ref in allocate.h:598-617,966-971)
//   -> asynchronous launches on one stream, pinned staging for source upload and sink download.
// per-launch cudaEvent timing reproduces `-d perf` (graph.c:881-933).
#include "pipe.h"
#include "mlv.h"
#include <algorithm>
#include <map>
#include <set>

int dt_graph_run_modules(dt_graph_t *g, std::vector<int> &modid);
void dt_graph_node_order(dt_graph_t *g, std::vector<int> &nodeid);

struct plan_buf_t
{
  size_t bytes = 0, offset = 0;
  int first = 1 << 30, last = -1;   // launch indices of first write / last read
  void *external = 0;               // device pointer owned by the caller (vkb_graph_set_source_device)
  int pinned_live = 0;              // must survive the whole run (sink input)
};
struct plan_img_t { int buf; uint32_t wd, ht, chan, layers; dt_token_t format; };
struct plan_launch_t
{
  dt_token_t name, kernel;
  uint32_t wd, ht, dp;
  std::vector<uint8_t> push;
  std::vector<int> param_mods;      // modules whose (committed) params are concatenated at launch time
  std::vector<plan_img_t> conn;
  std::string label;
  float ms = 0.0f;
  std::vector<uint8_t> arg_params;  // this run's arguments (filled by dt_graph_run)
  std::vector<vkb_image_t> arg_conn;
};
struct plan_graph_t { cudaGraphExec_t exec = 0; uint64_t hash = 0; int launches = 0; }; // one captured frame
struct plan_source_t { int modid; int nodeid; int buf_upload; size_t bytes; int packed_bpp; int external; };
struct plan_sink_t   { int modid; int nodeid; int buf; size_t bytes; uint32_t wd, ht; int rgb; };

struct vkb_plan_t
{
  std::vector<plan_buf_t> buf;
  std::vector<plan_launch_t> launch;
  std::vector<plan_source_t> source;
  std::vector<plan_sink_t> sink;
  std::vector<int> modid;
  void *pool = 0; size_t pool_bytes = 0;
  void *staging_up = 0; size_t staging_up_bytes = 0;     // pinned host
  void *staging_down = 0; size_t staging_down_bytes = 0; // pinned host
  cudaStream_t stream = 0;
  std::vector<cudaEvent_t> ev;
  std::vector<plan_graph_t> graphs = std::vector<plan_graph_t>(4); // small cache keyed by the fingerprint of the launch arguments
  unsigned graph_next = 0;
};

static void plan_free(vkb_plan_t *p)
{
  if(!p) return;
  if(p->pool) cudaFree(p->pool);
  if(p->staging_up) cudaFreeHost(p->staging_up);
  if(p->staging_down) cudaFreeHost(p->staging_down);
  for(cudaEvent_t e : p->ev) cudaEventDestroy(e);
  for(plan_graph_t &cg : p->graphs) if(cg.exec) cudaGraphExecDestroy(cg.exec);
  if(p->stream) cudaStreamDestroy(p->stream);
  delete p;
}
void dt_graph_cleanup(dt_graph_t *g)
{
  if(!g) return;
  plan_free(g->plan);
  for(dt_module_t &m : g->module) if(m.name && m.so && m.so->cleanup) m.so->cleanup(&m);
  delete g;
}

static inline bool is_node(const dt_node_t *n, const char *name, const char *kernel) { return n->name == dt_token(name) && n->kernel == dt_token(kernel); }
static inline bool is_pointwise(const dt_node_t *n)
{
  return is_node(n, "crop", "main") || is_node(n, "colour", "main") || is_node(n, "filmcurv", "main") || is_node(n, "grade", "main");
}
static inline uint32_t pw_op(const dt_node_t *n)
{
  if(is_node(n, "crop", "main")) return 1;
  if(is_node(n, "colour", "main")) return 2;
  if(is_node(n, "filmcurv", "main")) return 3;
  return 4;
}
static size_t conn_bytes(const dt_connector_t *c)
{
  const size_t layers = c->array_length > 0 ? c->array_length : 1;
  return (size_t)c->roi.wd * c->roi.ht * dt_connector_channels(c) * dt_connector_bytes_per_channel(c) * layers;
}

struct builder_t
{
  dt_graph_t *g;
  vkb_plan_t *p;
  std::vector<int> order;                       // reachable nodes, topological
  std::vector<uint8_t> reach, consumed;
  std::map<std::pair<int,int>, std::vector<std::pair<int,int>>> consumers; // (node, out conn) -> [(node, in conn)]

  int out_buf(int n, int c)
  { // buffer of an owner connector, created on demand
    dt_connector_t *cn = &g->node[n].connector[c];
    if(cn->buf < 0)
    {
      plan_buf_t b;
      b.bytes = ((conn_bytes(cn) + 255) / 256) * 256 + 256;
      p->buf.push_back(b);
      cn->buf = (int)p->buf.size() - 1;
    }
    return cn->buf;
  }
  plan_img_t img_out(int n, int c)
  {
    dt_connector_t *cn = &g->node[n].connector[c];
    return plan_img_t{ out_buf(n, c), cn->roi.wd, cn->roi.ht, (uint32_t)dt_connector_channels(cn), (uint32_t)(cn->array_length > 0 ? cn->array_length : 1), cn->format };
  }
  plan_img_t img_in(int n, int c)
  { // an input sees the image its owner declared
    const dt_cid_t src = g->node[n].connector[c].connected;
    if(src.i < 0 || src.i >= (int)g->node.size() || src.c < 0) return plan_img_t{ -1, 0, 0, 1, 1, dt_token("f16") };
    return img_out(src.i, src.c);
  }
  const std::vector<std::pair<int,int>> &cons(int n, int c) { return consumers[{n, c}]; }
  void add_launch(plan_launch_t &l)
  {
    const int idx = (int)p->launch.size();
    for(size_t i = 0; i < l.conn.size(); i++) if(l.conn[i].buf >= 0)
    {
      plan_buf_t &b = p->buf[l.conn[i].buf];
      b.first = std::min(b.first, idx);
      b.last = std::max(b.last, idx);
    }
    p->launch.push_back(l);
  }
  plan_launch_t node_launch(int n)
  {
    const dt_node_t *nd = &g->node[n];
    plan_launch_t l;
    l.name = nd->name; l.kernel = nd->kernel; l.wd = nd->wd; l.ht = nd->ht; l.dp = nd->dp;
    l.push.assign(nd->push_constant, nd->push_constant + nd->push_constant_size);
    l.param_mods.push_back((int)(nd->module - g->module.data()));
    l.label = dt_token_string(nd->module->name) + ":" + dt_token_string(nd->module->inst) + " " + dt_token_string(nd->name) + "_" + dt_token_string(nd->kernel);
    for(int c = 0; c < nd->num_connectors; c++)
      l.conn.push_back(dt_connector_owner(nd->connector + c) ? img_out(n, c) : img_in(n, c));
    return l;
  }
};

static int find_conn(const dt_node_t *n, const char *name)
{
  for(int c = 0; c < n->num_connectors; c++) if(n->connector[c].name == dt_token(name)) return c;
  return -1;
}

// llap/main.c:31-105 wired: curve -> reduce[1..nl-1], assemble[nl-1..1], colour.  restructured launches, see kernels/k_llap.cu
static int plan_llap(builder_t &B, int n_curve)
{
  dt_graph_t *g = B.g;
  const dt_module_t *mod = g->node[n_curve].module;
  const int modid = (int)(mod - g->module.data());
  std::vector<int> reduce, assemble; int n_colour = -1;
  for(int n : B.order) if(g->node[n].module == mod)
  {
    if(is_node(&g->node[n], "llap", "reduce")) reduce.push_back(n);
    else if(is_node(&g->node[n], "llap", "assemble")) assemble.push_back(n);
    else if(is_node(&g->node[n], "llap", "colour")) n_colour = n;
  }
  // order by level: reduce by decreasing size, assemble by the level they write
  std::sort(reduce.begin(), reduce.end(), [&](int a, int b) { return g->node[a].wd > g->node[b].wd || (g->node[a].wd == g->node[b].wd && g->node[a].ht > g->node[b].ht) || (g->node[a].wd == g->node[b].wd && g->node[a].ht == g->node[b].ht && a < b); });
  std::sort(assemble.begin(), assemble.end(), [&](int a, int b) { return g->node[a].wd > g->node[b].wd || (g->node[a].wd == g->node[b].wd && g->node[a].ht > g->node[b].ht) || (g->node[a].wd == g->node[b].wd && g->node[a].ht == g->node[b].ht && a < b); });
  if(reduce.empty() || assemble.size() != reduce.size() || n_colour < 0) return vkb_set_error(VKB_ERR_GRAPH, "llap: unexpected node structure");
  const int nl = (int)reduce.size() + 1;
  const std::string lab = "llap:" + dt_token_string(mod->inst) + " ";
  // 1) curve + reduce level 1
  {
    plan_launch_t l;
    l.name = dt_token("b200"); l.kernel = dt_token("llapr0"); l.wd = g->node[reduce[0]].wd; l.ht = g->node[reduce[0]].ht; l.dp = 1;
    l.param_mods.push_back(modid);
    l.conn.push_back(B.img_in(n_curve, 0));
    l.conn.push_back(B.img_out(reduce[0], 1));
    l.label = lab + "b200_llapr0 (curve+reduce)";
    B.add_launch(l);
  }
  // 2) coarser reduces
  for(int l = 1; l < nl - 1; l++)
  {
    plan_launch_t L = B.node_launch(reduce[l]);
    L.conn[0] = B.img_out(reduce[l-1], 1);
    B.add_launch(L);
  }
  // 3) assembles from coarse to level 1 (assemble[k] writes level k: k = nl-2 .. 1), the finest (k = 0) is fused below
  for(int k = nl - 2; k >= 1; k--)
  {
    plan_launch_t L = B.node_launch(assemble[k]);
    const int first = (k == nl - 2);
    L.push.resize(8);
    ((uint32_t *)L.push.data())[0] = 10; ((uint32_t *)L.push.data())[1] = first;
    L.conn[0] = first ? B.img_out(reduce[k], 1) : B.img_out(assemble[k+1], 3);
    L.conn[1] = B.img_out(reduce[k-1], 1);
    L.conn[2] = B.img_out(reduce[k], 1);
    L.conn[3] = B.img_out(assemble[k], 3);
    B.add_launch(L);
  }
  // 4) finest assemble + colour (+ grade when it is the only consumer)
  {
    int n_out = n_colour, c_out = 2, have_grade = 0, grade_mod = -1;
    const auto &cs = B.cons(n_colour, 2);
    if(cs.size() == 1 && is_node(&g->node[cs[0].first], "grade", "main") && cs[0].second == 0)
    {
      have_grade = 1; n_out = cs[0].first; c_out = find_conn(&g->node[n_out], "output");
      grade_mod = (int)(g->node[n_out].module - g->module.data());
      B.consumed[n_out] = 1;
    }
    plan_launch_t l;
    l.name = dt_token("b200"); l.kernel = dt_token("llapfin"); l.wd = g->node[n_colour].wd; l.ht = g->node[n_colour].ht; l.dp = 1;
    const int first = nl == 2;
    l.push.resize(8);
    ((uint32_t *)l.push.data())[0] = first; ((uint32_t *)l.push.data())[1] = have_grade;
    l.param_mods.push_back(modid);
    if(have_grade) l.param_mods.push_back(grade_mod);
    l.conn.push_back(B.img_in(n_colour, 1));
    l.conn.push_back(first ? B.img_out(reduce[0], 1) : B.img_out(assemble[1], 3));
    l.conn.push_back(B.img_out(reduce[0], 1));
    l.conn.push_back(B.img_out(n_out, c_out));
    l.label = lab + (have_grade ? "b200_llapfin (assemble+colour+grade)" : "b200_llapfin (assemble+colour)");
    B.add_launch(l);
  }
  B.consumed[n_curve] = 1; B.consumed[n_colour] = 1;
  for(int n : reduce) B.consumed[n] = 1;
  for(int n : assemble) B.consumed[n] = 1;
  return VKB_OK;
}

static int build_plan(dt_graph_t *g, bool with_device)
{
  plan_free(g->plan);
  g->plan = 0;
  vkb_plan_t *p = new vkb_plan_t();
  int r = dt_graph_run_modules(g, p->modid);
  if(r) { delete p; return r; }
  builder_t B;
  B.g = g; B.p = p;
  dt_graph_node_order(g, B.order);
  if(B.order.empty()) { delete p; return vkb_set_error(VKB_ERR_GRAPH, "no node is reachable from a sink"); }
  B.reach.assign(g->node.size(), 0); B.consumed.assign(g->node.size(), 0);
  for(int n : B.order) B.reach[n] = 1;
  for(dt_node_t &n : g->node) for(int c = 0; c < n.num_connectors; c++) n.connector[c].buf = -1;
  // all inputs connected? (graph.c:776-793)
  for(int n : B.order) for(int c = 0; c < g->node[n].num_connectors; c++)
  {
    dt_connector_t *cn = &g->node[n].connector[c];
    if(!dt_connector_input(cn)) continue;
    if(cn->connected.i < 0 || cn->connected.i >= (int)g->node.size())
    {
      delete p;
      return vkb_set_error(VKB_ERR_GRAPH, "kernel %s_%s:%s is not connected", dt_token_string(g->node[n].name).c_str(),
          dt_token_string(g->node[n].kernel).c_str(), dt_token_string(cn->name).c_str());
    }
    // dummy bindings (unconnected lut / gainmap inputs are wired to `input` in the reference) are not consumers
    const dt_node_t *nd = &g->node[n];
    if((is_node(nd, "colour", "main") && c >= 2) || (is_node(nd, "denoise", "noop") && c == 2) || (is_node(nd, "denoise", "doub") && c == 4)) continue;
    B.consumers[{cn->connected.i, cn->connected.c}].push_back({n, c});
  }
  // identity resample: alias output to input (demosaic/main.c:193-201 appends it in cli exports)
  for(int n : B.order)
  {
    dt_node_t *nd = &g->node[n];
    if(is_node(nd, "shared", "resample"))
    {
      const dt_cid_t src = nd->connector[0].connected;
      const dt_connector_t *so = &g->node[src.i].connector[src.c];
      if(so->roi.wd == nd->connector[1].roi.wd && so->roi.ht == nd->connector[1].roi.ht)
      {
        for(auto &cs : B.consumers[{n, 1}])
        {
          g->node[cs.first].connector[cs.second].connected = src;
          B.consumers[{src.i, src.c}].push_back(cs);
        }
        auto &v = B.consumers[{src.i, src.c}];
        v.erase(std::remove(v.begin(), v.end(), std::make_pair(n, 0)), v.end());
        B.consumers[{n, 1}].clear();
        B.consumed[n] = 1;
      }
      // otherwise a real resample: generic (shared, resample) launch below
    }
    if(is_node(nd, "demosaic", "down"))
    { // dead: gauss.comp never samples it
      B.consumed[n] = 1;
      const dt_cid_t src = nd->connector[0].connected;
      auto &v = B.consumers[{src.i, src.c}];
      v.erase(std::remove(v.begin(), v.end(), std::make_pair(n, 0)), v.end());
    }
  }
  for(int n : B.order)
  {
    if(B.consumed[n]) continue;
    dt_node_t *nd = &g->node[n];
    const int modid = (int)(nd->module - g->module.data());
    if(nd->connector[0].type == dt_token("source"))
    { // upload target; packed mlv payloads are unpacked on the device
      plan_source_t s;
      s.modid = modid; s.nodeid = n; s.packed_bpp = 0; s.external = 0;
      const vkb_mem_source_t *ms = modid < (int)g->mem_source.size() && g->mem_source[modid].valid ? &g->mem_source[modid] : 0;
      const uint32_t wd = nd->connector[0].roi.wd, ht = nd->connector[0].roi.ht;
      if(nd->module->name == dt_token("i-mlv"))
      {
        if(ms) s.packed_bpp = ms->p.packed_bpp;
        else s.packed_bpp = ((mlv_clip_t *)nd->module->data)->lossless ? 0 : ((mlv_clip_t *)nd->module->data)->bpp;
      }
      s.external = ms && ms->on_device;
      const int out = B.out_buf(n, 0);
      if(s.packed_bpp)
      {
        plan_buf_t pb; pb.bytes = mlv_packed_bytes(wd, ht, s.packed_bpp) + 256;
        p->buf.push_back(pb);
        s.buf_upload = (int)p->buf.size() - 1;
        s.bytes = mlv_packed_bytes(wd, ht, s.packed_bpp);
        plan_img_t packed{ s.buf_upload, (uint32_t)(s.bytes / 2), 1, 1, 1, dt_token("ui16") };
        // fuse with denoise/noop when it is the only consumer and does not crop
        const auto &cs = B.cons(n, 0);
        bool fused = false;
        if(cs.size() == 1 && is_node(&g->node[cs[0].first], "denoise", "noop"))
        {
          const dt_node_t *nn = &g->node[cs[0].first];
          const int32_t *pc = (const int32_t *)nn->push_constant;
          const int oc = find_conn(nn, "output");
          if(pc[0] == 0 && pc[1] == 0 && nn->connector[oc].roi.wd == wd && nn->connector[oc].roi.ht == ht)
          {
            plan_launch_t l;
            l.name = dt_token("b200"); l.kernel = dt_token("rawnoop"); l.wd = wd; l.ht = ht; l.dp = 1;
            l.push.resize(12);
            ((int32_t *)l.push.data())[0] = s.packed_bpp;
            memcpy(l.push.data() + 4, nn->push_constant + 16, 4); // black.r
            memcpy(l.push.data() + 8, nn->push_constant + 32, 4); // white.r
            l.conn.push_back(packed);
            l.conn.push_back(B.img_out(cs[0].first, oc));
            l.label = dt_token_string(nd->module->name) + "+denoise b200_rawnoop (unpack+noop)";
            B.add_launch(l);
            B.consumed[cs[0].first] = 1;
            fused = true;
          }
        }
        if(!fused)
        {
          plan_launch_t l;
          l.name = dt_token("i-mlv"); l.kernel = dt_token("unpack"); l.wd = wd; l.ht = ht; l.dp = 1;
          l.push.resize(4);
          ((int32_t *)l.push.data())[0] = s.packed_bpp;
          l.conn.push_back(packed);
          l.conn.push_back(B.img_out(n, 0));
          l.label = dt_token_string(nd->module->name) + " i-mlv_unpack";
          B.add_launch(l);
        }
      }
      else if(nd->connector[0].format == dt_token("f32") || (nd->module->num_connectors > 0 && nd->module->connector[0].format == dt_token("f32")))
      { // f32 sources (i-pfm): the kernels of this path read f16 edges.  upload as is, convert once on the device and let
        // every consumer see the f16 image (lossless for images that were f16 before they became a pfm file)
        const uint32_t chan = (uint32_t)dt_connector_channels(nd->connector);
        s.bytes = (size_t)wd * ht * chan * 4;
        plan_buf_t ub; ub.bytes = s.bytes + 256;
        p->buf.push_back(ub);
        s.buf_upload = (int)p->buf.size() - 1;
        g->node[n].connector[0].format = dt_token("f16");
        plan_launch_t l;
        l.name = dt_token("b200"); l.kernel = dt_token("cvt16"); l.wd = wd; l.ht = ht; l.dp = 1;
        l.conn.push_back(plan_img_t{ s.buf_upload, wd, ht, chan, 1, dt_token("f32") });
        l.conn.push_back(B.img_out(n, 0));
        l.label = dt_token_string(nd->module->name) + " b200_cvt16 (f32 -> f16)";
        B.add_launch(l);
      }
      else { s.buf_upload = out; s.bytes = conn_bytes(nd->connector); }
      p->buf[s.buf_upload].first = -1; // written by the upload, before launch 0
      // a run without UPLOAD_SOURCE (parameters changed, same image) reads the source again: the reference keeps source
      // connectors s_conn_protected for that (graph-run-nodes-allocate.h:959-963), here the upload buffer is never recycled
      p->buf[s.buf_upload].pinned_live = 1;
      if(s.external) p->buf[s.buf_upload].external = (void *)ms->data;
      p->source.push_back(s);
      continue;
    }
    if(nd->connector[0].type == dt_token("sink"))
    {
      plan_sink_t s;
      const plan_img_t in = B.img_in(n, 0);
      s.modid = modid; s.nodeid = n; s.buf = in.buf; s.wd = in.wd; s.ht = in.ht;
      s.bytes = (size_t)in.wd * in.ht * in.chan * (in.format == dt_token("f32") ? 4 : 2);
      s.rgb = 0;
      if(in.buf >= 0 && (in.wd == 0 || in.ht == 0 || (uint64_t)in.wd * in.ht > (1ull << 32)))
      { // e.g. a crop window with no area: the reference signals failure by an empty roi (graph-run-modules.h:652-666)
        plan_free(p);
        return vkb_set_error(VKB_ERR_GRAPH, "sink %s:%s would receive a %ux%u image", dt_token_string(nd->module->name).c_str(),
            dt_token_string(nd->module->inst).c_str(), in.wd, in.ht);
      }
      // packed rgb f32 (the PFM payload): asked for by the caller, or implied by o-pfm writing a file
      const vkb_mem_sink_t *msk = modid < (int)g->mem_sink.size() ? &g->mem_sink[modid] : 0;
      const bool to_file = !(msk && msk->valid) && nd->module->name == dt_token("o-pfm");
      if(in.buf >= 0 && in.chan == 4 && in.format == dt_token("f32") && (to_file || (msk && msk->layout == VKB_SINK_RGB_F32)))
      {
        // the launch that writes this buffer: if it is one of the fused kernels it stores r g b directly
        bool fused = false;
        for(int li = (int)p->launch.size() - 1; li >= 0 && !fused; li--)
        {
          plan_launch_t &pl = p->launch[li];
          int ci = -1;
          for(size_t c = 0; c < pl.conn.size(); c++) if(pl.conn[c].buf == in.buf) ci = (int)c;
          if(ci < 0) continue;
          if(pl.name == dt_token("b200") && ((pl.kernel == dt_token("llapfin") && ci == 3) || (pl.kernel == dt_token("pointw") && ci == 1)))
          { pl.conn[ci].chan = 3; fused = true; }
          break;
        }
        if(!fused)
        { // any other producer: one repack launch behind it
          plan_buf_t b;
          b.bytes = (((size_t)in.wd * in.ht * 12 + 255) / 256) * 256 + 256;
          p->buf.push_back(b);
          plan_launch_t l;
          l.name = dt_token("b200"); l.kernel = dt_token("pfmpack"); l.wd = in.wd; l.ht = in.ht; l.dp = 1;
          l.conn.push_back(in);
          l.conn.push_back(plan_img_t{ (int)p->buf.size() - 1, in.wd, in.ht, 3, 1, dt_token("f32") });
          l.label = dt_token_string(nd->module->name) + " b200_pfmpack (rgba -> rgb)";
          B.add_launch(l);
          s.buf = l.conn[1].buf;
        }
        s.bytes = (size_t)in.wd * in.ht * 12;
        s.rgb = 1;
      }
      if(s.buf >= 0) p->buf[s.buf].pinned_live = 1;
      p->sink.push_back(s);
      continue;
    }
    if(is_node(nd, "llap", "curve")) { r = plan_llap(B, n); if(r) { plan_free(p); return r; } continue; }
    if(is_pointwise(nd))
    { // grow a chain while the edge has a single pointwise consumer (crop can only lead)
      std::vector<int> chain{ n };
      int cur = n;
      for(;;)
      {
        const int oc = find_conn(&g->node[cur], "output");
        const auto &cs = B.cons(cur, oc);
        if(cs.size() != 1) break;
        const dt_node_t *nx = &g->node[cs[0].first];
        if(!is_pointwise(nx) || is_node(nx, "crop", "main") || cs[0].second != 0 || B.consumed[cs[0].first]) break;
        if(nx->connector[find_conn(nx, "output")].roi.wd != g->node[cur].connector[oc].roi.wd) break;
        // the fused kernel holds one parameter block per op type: a second instance of a module starts a new launch
        bool repeated = false;
        for(int cn : chain) if(pw_op(&g->node[cn]) == pw_op(nx)) repeated = true;
        if(repeated) break;
        chain.push_back(cs[0].first);
        cur = cs[0].first;
        if(chain.size() == 7) break;
      }
      if(chain.size() == 1) { plan_launch_t l = B.node_launch(n); l.conn.resize(2); B.add_launch(l); continue; }
      plan_launch_t l;
      l.name = dt_token("b200"); l.kernel = dt_token("pointw"); l.dp = 1;
      l.push.resize(4 * (1 + chain.size()));
      ((uint32_t *)l.push.data())[0] = (uint32_t)chain.size();
      l.label = "b200_pointw (";
      for(size_t k = 0; k < chain.size(); k++)
      {
        ((uint32_t *)l.push.data())[1 + k] = pw_op(&g->node[chain[k]]);
        l.param_mods.push_back((int)(g->node[chain[k]].module - g->module.data()));
        l.label += (k ? "+" : "") + dt_token_string(g->node[chain[k]].name);
        if(k) B.consumed[chain[k]] = 1;
      }
      l.label += ")";
      const int last = chain.back(), oc = find_conn(&g->node[last], "output");
      l.wd = g->node[last].wd; l.ht = g->node[last].ht;
      l.conn.push_back(B.img_in(n, 0));
      l.conn.push_back(B.img_out(last, oc));
      B.add_launch(l);
      continue;
    }
    if(!vkb_find_kernel(nd->name, nd->kernel))
    {
      plan_free(p);
      return vkb_set_error(VKB_ERR_UNKNOWN_KERNEL, "no CUDA kernel for node %s_%s (module %s)", dt_token_string(nd->name).c_str(),
          dt_token_string(nd->kernel).c_str(), dt_token_string(nd->module->name).c_str());
    }
    plan_launch_t l = B.node_launch(n);
    if(is_node(nd, "demosaic", "gauss")) l.conn[0].buf = -1; // the dropped `down` output
    B.add_launch(l);
  }
  if(p->sink.empty()) { plan_free(p); return vkb_set_error(VKB_ERR_GRAPH, "graph has no sink"); }
  // ---- liveness -> offsets (first fit over a free list, in launch order) ----
  const int nl = (int)p->launch.size();
  for(plan_buf_t &b : p->buf) if(b.pinned_live) b.last = nl;
  struct seg_t { size_t off, size; };
  std::vector<seg_t> free_list{ { 0, (size_t)1 << 62 } };
  size_t peak = 0;
  auto alloc = [&](size_t bytes) {
    for(size_t i = 0; i < free_list.size(); i++) if(free_list[i].size >= bytes)
    {
      const size_t off = free_list[i].off;
      free_list[i].off += bytes; free_list[i].size -= bytes;
      if(!free_list[i].size) free_list.erase(free_list.begin() + i);
      peak = std::max(peak, off + bytes);
      return off;
    }
    return (size_t)0;
  };
  auto release = [&](size_t off, size_t bytes) {
    size_t i = 0;
    while(i < free_list.size() && free_list[i].off < off) i++;
    free_list.insert(free_list.begin() + i, seg_t{ off, bytes });
    if(i + 1 < free_list.size() && free_list[i].off + free_list[i].size == free_list[i+1].off)
    { free_list[i].size += free_list[i+1].size; free_list.erase(free_list.begin() + i + 1); }
    if(i > 0 && free_list[i-1].off + free_list[i-1].size == free_list[i].off)
    { free_list[i-1].size += free_list[i].size; free_list.erase(free_list.begin() + i); }
  };
  for(int t = -1; t <= nl; t++)
  {
    for(plan_buf_t &b : p->buf) if(!b.external && b.last >= 0 && b.first == t) b.offset = alloc(b.bytes);
    for(plan_buf_t &b : p->buf) if(!b.external && b.last == t && b.first <= t && b.last >= 0) release(b.offset, b.bytes);
  }
  p->pool_bytes = peak + 256;
  for(const plan_source_t &s : p->source) if(!s.external) p->staging_up_bytes = std::max(p->staging_up_bytes, s.bytes);
  for(const plan_sink_t &s : p->sink) p->staging_down_bytes = std::max(p->staging_down_bytes, s.bytes);
  if(!with_device) { g->plan = p; return VKB_OK; } // host-side planning only (vkb_graph_plan)
  cudaError_t e = cudaSetDevice(g->device);
  if(e == cudaSuccess) e = cudaMalloc(&p->pool, p->pool_bytes);
  if(e != cudaSuccess) { const char *msg = cudaGetErrorString(e); plan_free(p); return vkb_set_error(e == cudaErrorMemoryAllocation ? VKB_ERR_OOM : VKB_ERR_NO_DEVICE, "pool allocation of %zu bytes failed: %s", peak, msg); }
  cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if(p->staging_up_bytes && cudaHostAlloc(&p->staging_up, p->staging_up_bytes, cudaHostAllocDefault) != cudaSuccess)
  { plan_free(p); return vkb_set_error(VKB_ERR_OOM, "pinned upload staging allocation failed"); }
  p->ev.resize(nl + 1);
  for(cudaEvent_t &ev : p->ev) cudaEventCreate(&ev);
  g->plan = p;
  return VKB_OK;
}

static void *buf_ptr(const vkb_plan_t *p, int b)
{
  if(b < 0) return 0;
  if(p->buf[b].external) return p->buf[b].external;
  return (uint8_t *)p->pool + p->buf[b].offset;
}

int dt_graph_run(dt_graph_t *g, uint32_t run)
{
  int r;
  if(vkb_device_count() <= 0) return vkb_set_error(VKB_ERR_NO_DEVICE, "no CUDA device: vkdt_b200 has no CPU fallback");
  // per-launch timing is a property of the graph (vkb_graph_set_perf, the reference's -d perf log mask) or an explicit flag
  // outside the reference's bits; s_graph_run_all (-1u, every bit set) never implies it
  const bool perf = g->perf || (run != (uint32_t)VKB_RUN_ALL && (run & VKB_RUN_PERF));
  if((run & (VKB_RUN_ROI | VKB_RUN_CREATE_NODES | VKB_RUN_ALLOC)) || !g->plan || !g->plan->pool)
  {
    r = build_plan(g, true);
    if(r) return r;
    run |= VKB_RUN_UPLOAD_SOURCE | VKB_RUN_RECORD_CMD_BUF;
  }
  vkb_plan_t *p = g->plan;
  cudaSetDevice(g->device);
  // sources with s_module_request_read_source re-upload every run (graph-run-modules.h:573-587)
  for(const plan_source_t &s : p->source)
  {
    dt_module_t *mod = &g->module[s.modid];
    const bool want = (run & VKB_RUN_UPLOAD_SOURCE) || ((run & VKB_RUN_RECORD_CMD_BUF) && (mod->flags & s_module_request_read_source));
    if(!want) continue;
    if(s.external)
    { // caller owned device memory: just repoint
      p->buf[s.buf_upload].external = (void *)g->mem_source[s.modid].data;
      continue;
    }
    {
      const vkb_mem_source_t *ms = s.modid < (int)g->mem_source.size() && g->mem_source[s.modid].valid ? &g->mem_source[s.modid] : 0;
      const uint32_t swd = g->node[s.nodeid].connector[0].roi.wd;
      if(ms && (s.packed_bpp || swd == ms->p.width))
      { // the caller's buffer already has the staging layout: copy straight from it (pinned if it came from vkb_host_alloc)
        const size_t payload = s.packed_bpp ? ((size_t)ms->p.width * ms->p.height * s.packed_bpp + 7) / 8 : s.bytes;
        cudaMemcpyAsync(buf_ptr(p, s.buf_upload), ms->data, payload, cudaMemcpyHostToDevice, p->stream);
        continue;
      }
    }
    if(!mod->so->read_source) return vkb_set_error(VKB_ERR_GRAPH, "source module %s has no read_source", dt_token_string(mod->name).c_str());
    cudaStreamSynchronize(p->stream); // staging is reused
    dt_read_source_params_t rp = { &g->node[s.nodeid], 0, 0 };
    if(mod->so->read_source(mod, p->staging_up, &rp)) return vkb_set_error(VKB_ERR_IO, "read_source failed for %s:%s", dt_token_string(mod->name).c_str(), dt_token_string(mod->inst).c_str());
    cudaMemcpyAsync(buf_ptr(p, s.buf_upload), p->staging_up, s.bytes, cudaMemcpyHostToDevice, p->stream);
  }
  if(run & VKB_RUN_RECORD_CMD_BUF)
  {
    // commit_params for every module in traversal order (graph-run-modules.h:5-31)
    for(int m : p->modid) if(g->module[m].so->commit_params) g->module[m].so->commit_params(g, &g->module[m]);
    // launch arguments of this run, and their fingerprint: every byte a kernel can see (parameters, push constants,
    // image pointers and shapes)
    uint64_t hash = 1469598103934665603ull ^ (uint64_t)g->mode;
    auto mix = [&hash](const void *d, size_t n) { const uint8_t *b = (const uint8_t *)d; for(size_t k = 0; k < n; k++) { hash ^= b[k]; hash *= 1099511628211ull; } };
    for(size_t i = 0; i < p->launch.size(); i++)
    {
      plan_launch_t &l = p->launch[i];
      l.arg_params.clear();
      for(int m : l.param_mods)
      {
        const dt_module_t *mod = &g->module[m];
        const uint8_t *src = mod->committed_param_size ? mod->committed_param : mod->param;
        const int sz = mod->committed_param_size ? mod->committed_param_size : mod->param_size;
        l.arg_params.insert(l.arg_params.end(), src, src + sz);
      }
      l.arg_conn.clear();
      for(const plan_img_t &im : l.conn) l.arg_conn.push_back(vkb_image_t{ buf_ptr(p, im.buf), im.wd, im.ht, im.chan, im.layers, im.format });
      mix(l.arg_params.data(), l.arg_params.size());
      mix(l.push.data(), l.push.size());
      for(const vkb_image_t &im : l.arg_conn) { mix(&im.data, sizeof(im.data)); mix(&im.wd, 4); mix(&im.ht, 4); mix(&im.chan, 4); mix(&im.layers, 4); mix(&im.format, sizeof(im.format)); }
    }
    auto dispatch_all = [&](bool events) -> int {
      for(size_t i = 0; i < p->launch.size(); i++)
      {
        plan_launch_t &l = p->launch[i];
        if(events) cudaEventRecord(p->ev[i], p->stream);
        const vkb_launch_t kl = { l.wd, l.ht, l.dp, l.push.data(), (uint32_t)l.push.size(), l.arg_params.data(), (uint32_t)l.arg_params.size(),
            l.arg_conn.data(), (uint32_t)l.arg_conn.size(), p->stream, -1, -1 };
        const int rr = vkb_dispatch_launch(l.name, l.kernel, g->mode, &kl);
        if(rr) return rr;
      }
      if(events) cudaEventRecord(p->ev[p->launch.size()], p->stream);
      return 0;
    };
    // frame loops replay a CUDA graph: the launch sequence of a frame is captured once per distinct set of arguments
    // (a clip ping-ponging between two device buffers has two) and re-launched as one unit, which takes the ~60 kernel
    // launches and their dependency latencies off the host and the stream front end.  -d perf runs stay plain launches.
    static const bool no_graph = getenv("VKB_NO_CUDA_GRAPH") != 0;
    if(perf || no_graph) { r = dispatch_all(perf); if(r) return r; }
    else
    {
      plan_graph_t *hit = 0;
      for(plan_graph_t &cg : p->graphs) if(cg.exec && cg.hash == hash) hit = &cg;
      if(!hit)
      {
        const uint64_t before = vkb_launch_count();
        if(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        { cudaGetLastError(); r = dispatch_all(false); if(r) return r; }
        else
        {
          r = dispatch_all(false);
          cudaGraph_t cap = 0;
          const cudaError_t e = cudaStreamEndCapture(p->stream, &cap);
          if(r) { if(cap) cudaGraphDestroy(cap); return r; }
          if(e != cudaSuccess || !cap) return vkb_set_error(VKB_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(e));
          plan_graph_t &slot = p->graphs[p->graph_next++ % p->graphs.size()];
          if(slot.exec) cudaGraphExecDestroy(slot.exec);
          slot.exec = 0;
          const cudaError_t e2 = cudaGraphInstantiate(&slot.exec, cap, 0);
          cudaGraphDestroy(cap);
          if(e2 != cudaSuccess) { slot.exec = 0; return vkb_set_error(VKB_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e2)); }
          slot.hash = hash;
          slot.launches = (int)(vkb_launch_count() - before);
          vkb_count_launch(-slot.launches); // counted again below, when the captured launches actually run
          hit = &slot;
        }
      }
      if(hit)
      {
        const cudaError_t e = cudaGraphLaunch(hit->exec, p->stream);
        if(e != cudaSuccess) return vkb_set_error(VKB_ERR_CUDA, "graph launch failed: %s", cudaGetErrorString(e));
        vkb_count_launch(hit->launches);
      }
    }
  }
  if(run & VKB_RUN_DOWNLOAD_SINK)
  {
    for(const plan_sink_t &s : p->sink)
    {
      dt_module_t *mod = &g->module[s.modid];
      const vkb_mem_sink_t *ms = s.modid < (int)g->mem_sink.size() && g->mem_sink[s.modid].valid ? &g->mem_sink[s.modid] : 0;
      if(ms && ms->dst)
      {
        if(ms->bytes < s.bytes) return vkb_set_error(VKB_ERR_BAD_ARG, "sink buffer too small: %zu < %zu", ms->bytes, s.bytes);
        cudaMemcpyAsync(ms->dst, buf_ptr(p, s.buf), s.bytes, cudaMemcpyDeviceToHost, p->stream);
      }
      else if(mod->so->write_sink && !(ms && !ms->dst))
      {
        if(!p->staging_down && cudaHostAlloc(&p->staging_down, p->staging_down_bytes, cudaHostAllocDefault) != cudaSuccess)
          return vkb_set_error(VKB_ERR_OOM, "pinned download staging allocation failed");
        cudaMemcpyAsync(p->staging_down, buf_ptr(p, s.buf), s.bytes, cudaMemcpyDeviceToHost, p->stream);
        cudaStreamSynchronize(p->stream);
        dt_write_sink_params_t wp = { &g->node[s.nodeid], 0, 0 };
        mod->so->write_sink(mod, p->staging_down, &wp);
      }
    }
  }
  // downloads into caller memory are asynchronous until a run carries WAIT_DONE (two graphs can then ping-pong:
  // the D2H of one frame overlaps the upload + kernels of the next)
  if(run & VKB_RUN_WAIT_DONE)
  {
    cudaError_t e = cudaStreamSynchronize(p->stream);
    if(e != cudaSuccess) return vkb_set_error(VKB_ERR_CUDA, "graph run failed: %s", cudaGetErrorString(e));
    if((run & VKB_RUN_RECORD_CMD_BUF) && perf)
    { // -d perf (graph.c:881-933)
      char b[256];
      g->perf_text.clear();
      float total = 0.0f;
      for(size_t i = 0; i < p->launch.size(); i++)
      {
        cudaEventElapsedTime(&p->launch[i].ms, p->ev[i], p->ev[i+1]);
        total += p->launch[i].ms;
        size_t bytes = 0; // unique bytes in + out of this launch
        std::set<int> seen;
        for(const plan_img_t &im : p->launch[i].conn) if(im.buf >= 0 && seen.insert(im.buf).second)
          bytes += (size_t)im.wd * im.ht * im.chan * im.layers * (im.format == dt_token("f32") ? 4 : 2);
        snprintf(b, sizeof(b), "[perf] %-60s:\t%8.3f ms\t%12zu B\n", p->launch[i].label.c_str(), p->launch[i].ms, bytes);
        g->perf_text += b;
      }
      snprintf(b, sizeof(b), "[perf] total time:\t%8.3f ms\n", total);
      g->perf_text += b;
    }
  }
  return VKB_OK;
}

// host-side half only: module passes, rewrite, liveness.  no device needed (used by tests and --dump-nodes)
int dt_graph_plan(dt_graph_t *g, std::string *text)
{
  const int r = build_plan(g, false);
  if(r) return r;
  const vkb_plan_t *p = g->plan;
  char b[512];
  for(size_t i = 0; i < p->launch.size(); i++)
  {
    const plan_launch_t &l = p->launch[i];
    snprintf(b, sizeof(b), "launch %2zu %s_%s [%s]", i, dt_token_string(l.name).c_str(), dt_token_string(l.kernel).c_str(), l.label.c_str());
    *text += b;
    for(const plan_img_t &im : l.conn)
    {
      if(im.buf < 0) { *text += " -"; continue; }
      snprintf(b, sizeof(b), " b%d:%ux%ux%ux%u:%s@%zu", im.buf, im.wd, im.ht, im.chan, im.layers, dt_token_string(im.format).c_str(), p->buf[im.buf].offset);
      *text += b;
    }
    *text += "\n";
  }
  for(const plan_sink_t &s : p->sink) { snprintf(b, sizeof(b), "sink %s %ux%u b%d\n", dt_token_string(g->module[s.modid].name).c_str(), s.wd, s.ht, s.buf); *text += b; }
  for(const plan_source_t &s : p->source) { snprintf(b, sizeof(b), "source %s bytes %zu packed %d b%d\n", dt_token_string(g->module[s.modid].name).c_str(), s.bytes, s.packed_bpp, s.buf_upload); *text += b; }
  snprintf(b, sizeof(b), "pool %zu bytes, %zu buffers\n", p->pool_bytes, p->buf.size());
  *text += b;
  return VKB_OK;
}

// accessors used by the C-ABI
int vkb_plan_sink(dt_graph_t *g, int modid, uint32_t *wd, uint32_t *ht, void **dptr)
{
  if(!g->plan) return VKB_ERR_GRAPH;
  for(const plan_sink_t &s : g->plan->sink) if(s.modid == modid)
  {
    if(wd) *wd = s.wd; if(ht) *ht = s.ht;
    if(dptr) *dptr = buf_ptr(g->plan, s.buf);
    return VKB_OK;
  }
  return VKB_ERR_BAD_ARG;
}
uint64_t vkb_plan_pool_bytes(dt_graph_t *g) { return g->plan ? g->plan->pool_bytes : 0; }
int vkb_plan_launches(dt_graph_t *g) { return g->plan ? (int)g->plan->launch.size() : 0; }
void *vkb_plan_stream(dt_graph_t *g) { return g->plan ? (void *)g->plan->stream : 0; }

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
#include "bands.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <set>
#include <thread>

int dt_graph_run_modules(dt_graph_t *g, std::vector<int> &modid);
void dt_graph_node_order(dt_graph_t *g, std::vector<int> &nodeid);

struct plan_buf_t
{
  size_t bytes = 0, offset = 0;
  int first = 1 << 30, last = -1;   // launch indices of first write / last read
  void *external = 0;               // device pointer owned by the caller (vkb_graph_set_source_device)
  int pinned_live = 0;              // must survive the whole run (sink input)
  int frames = 1;                   // 2: double buffered (an s_conn_feedback input reads it): two copies of `bytes`, back to back,
                                    // never recycled.  frame f writes copy f & 1, feedback inputs read copy 1 - (f & 1)
};
struct plan_img_t { int buf; uint32_t wd, ht, chan, layers; dt_token_t format; int fb = 0; }; // fb: reads the other frame's copy
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

// ---- band split: one frame over several GPUs (build_bands below) ----
struct band_pull_t { int src; int buf; int r0, r1; };              // rows [r0, r1) of buffer `buf` copied from device slot `src`
struct band_step_t                                                  // what one device does for one launch of the plan
{
  std::vector<std::pair<int,int>> compute;                          // row intervals of the launch's band image (empty: nothing)
  std::vector<band_pull_t> pulls;                                   // halo rows fetched from peers before the launch
  std::vector<int> waits;                                           // device slots whose previous step must have completed
};
struct band_copy_t { int buf; int r0, r1; };
struct band_dev_t
{
  int device = 0;
  void *pool = 0;
  cudaStream_t stream = 0;
  std::vector<cudaEvent_t> ev;                                      // one per launch
  std::vector<band_step_t> step;
  std::vector<band_copy_t> upload, download;                        // source rows this device needs / sink rows it owns
  size_t pulled_bytes = 0;                                          // per frame, over NVLink (or device to device in tests)
  std::thread worker;
  int rc = 0; char err[256] = {0};
};
struct vkb_bands_t
{
  std::vector<band_dev_t> dev;
  std::mutex mtx; std::condition_variable cv;
  uint64_t frame = 0, go = 0; int pending = 0, quit = 0; uint32_t runflags = 0;
  std::vector<std::atomic<uint64_t>> issued;                        // per device: frame * 65536 + launches recorded so far
  cudaEvent_t t0 = 0, t1 = 0;
};

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
  // upload-only runs (VKB_RUN_UPLOAD_SOURCE alone) prefetch the next frame's source on a stream of their own: the copy waits
  // for the launches that still read the old source, not for the download behind them, and the next recorded run waits for it
  cudaStream_t up_stream = 0;
  cudaEvent_t ev_src_free = 0, ev_src_ready = 0;
  int src_prefetched = 0;
  std::vector<cudaEvent_t> ev;
  std::vector<plan_graph_t> graphs = std::vector<plan_graph_t>(4); // small cache keyed by the fingerprint of the launch arguments
  unsigned graph_next = 0;
  vkb_bands_t *bands = 0;
};

static void bands_free(vkb_bands_t *b);
static void plan_free(vkb_plan_t *p)
{
  if(!p) return;
  if(p->bands) { bands_free(p->bands); p->bands = 0; }
  if(p->pool) cudaFree(p->pool);
  if(p->staging_up) cudaFreeHost(p->staging_up);
  if(p->staging_down) cudaFreeHost(p->staging_down);
  for(cudaEvent_t e : p->ev) cudaEventDestroy(e);
  for(plan_graph_t &cg : p->graphs) if(cg.exec) cudaGraphExecDestroy(cg.exec);
  if(p->ev_src_free) cudaEventDestroy(p->ev_src_free);
  if(p->ev_src_ready) cudaEventDestroy(p->ev_src_ready);
  if(p->up_stream) cudaStreamDestroy(p->up_stream);
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
static inline bool colour_lut_input(const dt_node_t *n, int c)
{ // (colour, main) connectors 2.. are dummies wired to `input` unless the push constants { have_clut, have_pick, have_abney } say otherwise
  if(!is_node(n, "colour", "main") || n->push_constant_size < 12) return false;
  const int32_t *pc = (const int32_t *)n->push_constant;
  return ((c == 2 || c == 6) && pc[0]) || ((c == 4 || c == 5) && pc[2]);   // 6: the autotemp node's answer, there with a clut
}
static inline bool is_pointwise(const dt_node_t *n)
{
  if(colour_lut_input(n, 2) || colour_lut_input(n, 4)) return false; // reads luts: a launch of its own (k_colour_lut)
  return is_node(n, "crop", "main") || is_node(n, "colour", "main") || is_node(n, "filmcurv", "main") || is_node(n, "grade", "main") || is_node(n, "colenc", "main");
}
static inline uint32_t pw_op(const dt_node_t *n)
{
  if(is_node(n, "crop", "main")) return 1;
  if(is_node(n, "colour", "main")) return 2;
  if(is_node(n, "filmcurv", "main")) return 3;
  if(is_node(n, "colenc", "main")) return 5;
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
  std::map<std::pair<int,int>, plan_img_t> converted;                      // (node, out conn) -> its f16 copy, see img_in

  int out_buf(int n, int c)
  { // buffer of an owner connector, created on demand
    dt_connector_t *cn = &g->node[n].connector[c];
    if(cn->buf < 0)
    {
      plan_buf_t b;
      b.bytes = ((conn_bytes(cn) + 255) / 256) * 256 + 256;
      // double buffered when an s_conn_feedback input is wired to it (dt_module_feedback / dt_connector_copy carry frames = 2 to
      // the owner, init_connector_images graph-run-modules.h:169-196); such a buffer outlives the frame
      if(cn->frames == 2) { b.frames = 2; b.pinned_live = 1; b.first = -1; }
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
    plan_img_t im = img_out(src.i, src.c);
    // an rgba image that became f32 because an f32 sink hangs on it (dt_graph_replace_display pushes the sink's format onto
    // its producer, graph-export.c:88-91) and that ANOTHER module reads as well: the reference samples it as it is, the
    // kernels of this path read f16 edges.  the other reader gets a converted copy (what the edge held before the sink came)
    if(im.format == dt_token("f32") && im.chan == 4 && im.layers == 1 && g->node[n].connector[0].type != dt_token("sink") &&
       g->node[src.i].connector[0].type != dt_token("source"))
    {
      auto it = converted.find({src.i, src.c});
      if(it == converted.end())
      {
        plan_buf_t b;
        b.bytes = (((size_t)im.wd * im.ht * 4 * 2 + 255) / 256) * 256 + 256;
        p->buf.push_back(b);
        plan_img_t cv = im;
        cv.buf = (int)p->buf.size() - 1; cv.format = dt_token("f16");
        plan_launch_t l;
        l.name = dt_token("b200"); l.kernel = dt_token("cvt16"); l.wd = im.wd; l.ht = im.ht; l.dp = 1;
        l.conn.push_back(im);
        l.conn.push_back(cv);
        l.label = dt_token_string(g->node[src.i].module->name) + " b200_cvt16 (f32 -> f16 for a second reader)";
        add_launch(l);
        it = converted.insert({{src.i, src.c}, cv}).first;
      }
      return it->second;
    }
    // feedback inputs cross the frame wires (graph-run-nodes-allocate.h:222-229): they see what the owner wrote one frame ago
    if((g->node[n].connector[c].flags & s_conn_feedback) && p->buf[im.buf].frames == 2) im.fb = 1;
    return im;
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
    // (a gain map input that is wired to the module's (denoise, gainmap) source node is real)
    const bool real_gainmap = is_node(&g->node[cn->connected.i], "denoise", "gainmap");
    if((is_node(nd, "colour", "main") && c >= 2 && !colour_lut_input(nd, c)) || (is_node(nd, "colour", "autotemp") && c == 2) || (!real_gainmap && ((is_node(nd, "denoise", "noop") && c == 2) || (is_node(nd, "denoise", "doub") && c == 4)))) continue;
    B.consumers[{cn->connected.i, cn->connected.c}].push_back({n, c});
  }
  // a module between the graph and an f32 sink that bypassed itself (resize at 1:1): the sink then hangs on an f16 edge.  the
  // image a sink downloads has the sink's format (dt_graph_replace_display pushes it onto the producer, graph-export.c:88-91):
  // do the same here when the sink is that edge's only reader
  for(int n : B.order)
  {
    dt_connector_t *sc = &g->node[n].connector[0];
    if(sc->type != dt_token("sink") || sc->connected.i < 0 || sc->connected.i >= (int)g->node.size()) continue;
    dt_connector_t *oc = &g->node[sc->connected.i].connector[sc->connected.c];
    if(sc->format == dt_token("f32") && oc->format == dt_token("f16") && B.consumers[{sc->connected.i, sc->connected.c}].size() == 1)
      oc->format = dt_token("f32");
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
      if(!wd || !ht || dt_module_source_failed(nd->module))
      { // a file that could not be read (the reference shows a 32x32 placeholder in the gui; an export has nothing to show for it)
        const std::string who = dt_token_string(nd->module->name) + ":" + dt_token_string(nd->module->inst);
        delete p;
        return vkb_set_error(VKB_ERR_IO, "source %s has no image (file missing or unreadable)", who.c_str());
      }
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
          if(pc[0] == 0 && pc[1] == 0 && pc[17] == 0 && nn->connector[oc].roi.wd == wd && nn->connector[oc].roi.ht == ht)   // uncropped, no gain map
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
      else if(is_node(nd, "denoise", "gainmap"))
      { // the dng gain maps of denoise (denoise/main.c:181-196): a small rgba f32 texture, sampled as it is
        s.buf_upload = out; s.bytes = conn_bytes(nd->connector);
      }
      else if(nd->module->name == dt_token("i-lut"))
      { // lookup tables are sampled as stored (f16 or f32, 1 / 2 / 4 channels)
        s.buf_upload = out; s.bytes = conn_bytes(nd->connector);
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
      s.bytes = (size_t)in.wd * in.ht * in.chan * (in.format == dt_token("f32") ? 4 : (in.format == dt_token("ui8") ? 1 : 2));
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
      if(in.buf >= 0 && in.chan == 4 && in.format == dt_token("ui8") && msk && msk->valid && msk->layout == VKB_SINK_RGB_UI8)
      { // packed 8 bit rgb (3 B/px) into caller memory: the pointwise kernel that ends in colenc stores it directly
        for(int li = (int)p->launch.size() - 1; li >= 0; li--)
        {
          plan_launch_t &pl = p->launch[li];
          int ci = -1;
          for(size_t c = 0; c < pl.conn.size(); c++) if(pl.conn[c].buf == in.buf) ci = (int)c;
          if(ci < 0) continue;
          if((pl.name == dt_token("b200") && pl.kernel == dt_token("pointw") && ci == 1) || (pl.name == dt_token("colenc") && ci == 1))
          { pl.conn[ci].chan = 3; s.bytes = (size_t)in.wd * in.ht * 3; s.rgb = 1; }
          break;
        }
      }
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
    for(plan_buf_t &b : p->buf) if(!b.external && b.last >= 0 && b.first == t) b.offset = alloc(b.bytes * b.frames);
    for(plan_buf_t &b : p->buf) if(!b.external && b.last == t && b.first <= t && b.last >= 0) release(b.offset, b.bytes * b.frames);
  }
  p->pool_bytes = peak + 256;
  for(const plan_source_t &s : p->source) if(!s.external) p->staging_up_bytes = std::max(p->staging_up_bytes, s.bytes);
  for(const plan_sink_t &s : p->sink) p->staging_down_bytes = std::max(p->staging_down_bytes, s.bytes);
  if(!with_device) { g->plan = p; return VKB_OK; } // host-side planning only (vkb_graph_plan)
  cudaError_t e = cudaSetDevice(g->device);
  if(e == cudaSuccess) e = cudaMalloc(&p->pool, p->pool_bytes);
  if(e != cudaSuccess) { const char *msg = cudaGetErrorString(e); plan_free(p); return vkb_set_error(e == cudaErrorMemoryAllocation ? VKB_ERR_OOM : VKB_ERR_NO_DEVICE, "pool allocation of %zu bytes failed: %s", peak, msg); }
  // double buffered connectors are read one frame before they are first written: frame 0 sees zeros (the reference leaves
  // fresh device memory there unless the connector asks for s_conn_clear_once)
  for(const plan_buf_t &b : p->buf) if(b.frames == 2 && !b.external) cudaMemset((uint8_t *)p->pool + b.offset, 0, b.bytes * 2);
  cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&p->up_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&p->ev_src_free, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_src_ready, cudaEventDisableTiming);
  if(p->staging_up_bytes && cudaHostAlloc(&p->staging_up, p->staging_up_bytes, cudaHostAllocDefault) != cudaSuccess)
  { plan_free(p); return vkb_set_error(VKB_ERR_OOM, "pinned upload staging allocation failed"); }
  p->ev.resize(nl + 1);
  for(cudaEvent_t &ev : p->ev) cudaEventCreate(&ev);
  g->plan = p;
  return VKB_OK;
}


// ================================================================================================================
// band split (BASELINE.json config 5): ONE frame developed by several GPUs.
//
// every device gets the same plan and the same pool layout for the WHOLE image (7.6 GB of 180 at 201 MP), so a buffer has
// the same offset everywhere and every kernel keeps the coordinates, sizes and mirror rules of the one GPU run.  a device
// computes only its rows of every launch (vkb_launch_t::band_y0/y1); rows of an input it does not hold (the halo of a
// stencil, a pyramid level's neighbours, the other half of denoise's quadrant swizzle) are copied from the device that
// computed them, over NVLink with peer access, into the same offset of its own pool right before the launch.  pyramid
// levels with fewer than 32 rows per device are computed whole on every device (the first such level all-gathers its
// input through the same mechanism).  all of it is planned here on the host: which rows a device computes per launch
// (band_rows), which rows of which connector a launch reads for them (band_describe), hence the pulls; the result is
// bit for bit the one GPU frame.  execution: one host thread per device issuing its steps, devices ordered by events:
// a device waits for the peers it pulls from (read after write) and for the peers that pulled from it in the step before
// (its pool recycles memory: write after read).
// ================================================================================================================
enum { BA_NONE = 0, BA_READ, BA_READ_SWZ, BA_WRITE, BA_WRITE_SWZ };
struct band_acc_t { int kind, num, den, lo, hi; };     // BA_READ: rows [floor(y0 num/den) + lo, ceil(y1 num/den) + hi); BA_READ_SWZ: lo = depth
struct band_desc_t
{
  int band_conn = -1;      // connector whose rows are the launch's band coordinate
  int unit = 1;            // band boundaries are multiples of this
  int from_owned = -1;     // >= 0: compute the rows of this input connector that the device computed itself (swizzled levels)
  int edge = 0;            // rows within `edge` of the image border read up to edge + 6 rows from it (splat.comp:27-30)
  band_acc_t acc[DT_MAX_CONNECTORS];
};

static bool band_describe(const plan_launch_t &l, band_desc_t &D)
{
  auto is = [&](const char *n, const char *k) { return l.name == dt_token(n) && l.kernel == dt_token(k); };
  for(auto &a : D.acc) a = band_acc_t{ BA_NONE, 1, 1, 0, 0 };
  auto R = [&](int c, int num, int den, int lo, int hi) { D.acc[c] = band_acc_t{ BA_READ, num, den, lo, hi }; };
  auto W = [&](int c) { D.acc[c] = band_acc_t{ BA_WRITE, 1, 1, 0, 0 }; };
  if(is("denoise", "half"))
  { // half.comp: block (x, y) reads mosaic rows crop.y + 2y, + 2y + 1
    const int cy = l.push.size() >= 56 ? ((const int32_t *)l.push.data())[13] : 0;
    D.band_conn = 1; R(0, 2, 1, cy, cy); W(1);
  }
  else if(is("denoise", "downcov")) { D.band_conn = 0; R(0, 1, 1, -2, 2); D.acc[1] = band_acc_t{ BA_WRITE_SWZ, 1, 1, 0, 0 }; W(2); }
  else if(is("denoise", "down"))    { D.band_conn = 0; D.from_owned = 0; R(0, 1, 1, -2, 2); D.acc[1] = band_acc_t{ BA_WRITE_SWZ, 1, 1, 0, 0 }; }
  else if(is("denoise", "assemble"))
  { // assemble.comp:56-70: scale l is read at the l times swizzled position
    D.band_conn = 5; R(0, 1, 1, 0, 0);
    for(int k = 1; k <= 4; k++) D.acc[k] = band_acc_t{ BA_READ_SWZ, 1, 1, k, 0 };
    W(5);
  }
  else if(is("denoise", "doub"))
  { // band image: the output mosaic.  bayer kernel: block Y = y / 2 reads coarse rows Y - 1 .. Y + 1
    const int cy = l.push.size() >= 56 ? ((const int32_t *)l.push.data())[13] : 0;
    D.band_conn = 3; D.unit = 2; R(0, 1, 1, cy, cy); R(1, 1, 2, -1, 1); R(2, 1, 2, -1, 1); W(3);
  }
  else if(is("hilite", "half"))     { D.band_conn = 1; R(0, 2, 1, 0, 0); W(1); }
  else if(is("hilite", "reduce"))   { D.band_conn = 1; R(0, 2, 1, -2, 1); W(1); }
  else if(is("hilite", "assemble")) { D.band_conn = 2; R(0, 1, 1, 0, 0); R(1, 1, 2, -1, 1); W(2); }
  else if(is("hilite", "doub"))     { D.band_conn = 2; D.unit = 2; R(0, 1, 1, 0, 0); R(1, 1, 2, 0, 1); W(2); }
  else if(is("demosaic", "gauss"))  { D.band_conn = 2; R(1, 2, 1, -1, 1); W(2); }
  else if(is("demosaic", "splat"))  { D.band_conn = 2; D.unit = 2; D.edge = 2; R(0, 1, 1, -2, 2); R(1, 1, 2, 0, 1); W(2); }
  else if(is("demosaic", "fix"))    { D.band_conn = 3; D.unit = 2; R(0, 1, 1, -2, 2); R(1, 1, 1, -2, 2); R(2, 1, 2, 0, 1); W(3); }
  else if(is("b200", "pointw"))
  { // crop resolved to an integer shift (the launcher refuses anything else in a band): the shift is at most the size difference
    D.band_conn = 1; R(0, 1, 1, 0, (int)l.conn[0].ht - (int)l.conn[1].ht); W(1);
  }
  else if(is("b200", "llapr0"))     { D.band_conn = 1; R(0, 2, 1, -1, 0); W(1); }
  else if(is("llap", "reduce"))     { D.band_conn = 1; R(0, 2, 1, -1, 0); W(1); }
  else if(is("llap", "assemble"))   { D.band_conn = 3; D.unit = 2; R(0, 1, 2, -2, 2); R(1, 1, 1, 0, 0); R(2, 1, 2, -2, 2); W(3); }
  else if(is("b200", "llapfin"))    { D.band_conn = 3; D.unit = 2; R(0, 1, 1, 0, 0); R(1, 1, 2, -2, 2); R(2, 1, 2, -2, 2); W(3); }
  else return false;
  return true;
}

static void bands_free(vkb_bands_t *b)
{
  if(!b) return;
  { std::lock_guard<std::mutex> lk(b->mtx); b->quit = 1; }
  b->cv.notify_all();
  for(band_dev_t &d : b->dev) if(d.worker.joinable()) d.worker.join();
  for(band_dev_t &d : b->dev)
  {
    cudaSetDevice(d.device);
    for(cudaEvent_t e : d.ev) cudaEventDestroy(e);
    if(d.stream) cudaStreamDestroy(d.stream);
    if(d.pool) cudaFree(d.pool);
  }
  delete b;
}

// rows of an image of height hb that device slot d of n computes: the cut positions are those of the full resolution image
// (multiples of 64 rows of the source) scaled to the image's pyramid level, so that a device's bands nest across levels
static std::pair<int,int> band_rows(int hb, int hfull, int d, int n, int unit)
{
  if(hb < 32 * n) return { 0, hb };                                  // too small to split: every device computes it whole
  int k = 0;
  while(k < 20 && ((hfull >> (k + 1)) >= hb || std::abs((hfull >> (k + 1)) - hb) < std::abs((hfull >> k) - hb))) k++;
  auto cut = [&](int i) -> int {
    if(i <= 0) return 0;
    if(i >= n) return hb;
    const int yfull = (int)(((int64_t)hfull * i / n) / 64) * 64;
    int c = yfull >> k;
    c -= c % unit;
    return std::min(std::max(c, 0), hb);
  };
  return { cut(d), cut(d + 1) };
}

static void band_worker(dt_graph_t *g, int d);

// host half: what every device computes, pulls and waits for, per launch.  needs no device (vkb_graph_band_plan)
static int plan_bands(dt_graph_t *g, vkb_bands_t **out)
{
  vkb_plan_t *p = g->plan;
  const int n = (int)g->band_devices.size();
  const int nl = (int)p->launch.size();
  if(p->source.empty()) return vkb_set_error(VKB_ERR_GRAPH, "band split: no source");
  for(const plan_buf_t &b : p->buf) if(b.frames == 2)
    return vkb_set_error(VKB_ERR_GRAPH, "band split: the graph has double buffered (feedback) connectors, which are kept per device: run it on one GPU");
  const int hfull = (int)g->node[p->source[0].nodeid].connector[0].roi.ht;
  vkb_bands_t *B = new vkb_bands_t();
  B->dev.resize(n);
  B->issued = std::vector<std::atomic<uint64_t>>(n);
  for(int d = 0; d < n; d++) { B->dev[d].device = g->band_devices[d]; B->dev[d].step.resize(nl); B->issued[d].store(0); }
  // buffer heights (rows) and row strides, from the images the launches bind
  const int nb = (int)p->buf.size();
  std::vector<int> bh(nb, 0);
  for(const plan_launch_t &l : p->launch) for(const plan_img_t &im : l.conn) if(im.buf >= 0) bh[im.buf] = (int)im.ht;
  std::vector<uint8_t> is_src(nb, 0);
  for(const plan_source_t &s : p->source) { is_src[s.buf_upload] = s.external ? 2 : 1; bh[s.buf_upload] = std::max(bh[s.buf_upload], 1); }
  // valid[b][d]: rows of buffer b device d holds; own[b][d]: rows it computed itself
  std::vector<std::vector<rows_t>> valid(nb, std::vector<rows_t>(n)), own(nb, std::vector<rows_t>(n));
  for(int b = 0; b < nb; b++) if(is_src[b] == 2) for(int d = 0; d < n; d++) valid[b][d].add(0, 1 << 30); // caller's device memory, read over peer access
  std::vector<std::vector<rows_t>> up(nb, std::vector<rows_t>(n));
  std::vector<std::vector<std::vector<int>>> pulled_from(nl, std::vector<std::vector<int>>(n)); // [l][src] -> devices that pulled from src at launch l
  for(int li = 0; li < nl; li++)
  {
    const plan_launch_t &l = p->launch[li];
    band_desc_t D;
    if(!band_describe(l, D)) { bands_free(B); return vkb_set_error(VKB_ERR_GRAPH, "band split: kernel %s_%s has no band description (label %s)",
        dt_token_string(l.name).c_str(), dt_token_string(l.kernel).c_str(), l.label.c_str()); }
    const int hb = (int)l.conn[D.band_conn].ht;
    // rows every device computes
    std::vector<rows_t> comp(n);
    for(int d = 0; d < n; d++)
    {
      if(D.from_owned >= 0 && hb >= 32 * n) comp[d] = own[l.conn[D.from_owned].buf][d];
      else { const auto r = band_rows(hb, hfull, d, n, D.unit); comp[d].add(r.first, r.second); }
    }
    // reads -> pulls (against the state before this launch)
    for(int d = 0; d < n; d++)
    {
      band_step_t &st = B->dev[d].step[li];
      st.compute = comp[d].v;
      for(size_t c = 0; c < l.conn.size(); c++)
      {
        const band_acc_t &a = D.acc[c];
        const int b = l.conn[c].buf;
        if(b < 0 || (a.kind != BA_READ && a.kind != BA_READ_SWZ)) continue;
        const int hc = (int)l.conn[c].ht;
        rows_t need;
        if(a.kind == BA_READ)
        {
          for(auto &iv : comp[d].v)
          {
            const int r0 = (int)(((int64_t)iv.first * a.num) / a.den) + a.lo, r1 = (int)(((int64_t)iv.second * a.num + a.den - 1) / a.den) + a.hi;
            need.add(r0, r1);
            if(D.edge && iv.first <= D.edge) need.add(0, D.edge + 7);
            if(D.edge && iv.second >= hb - D.edge) need.add(hc - D.edge - 7, hc);
          }
        }
        else { need = comp[d]; for(int k = 0; k < a.lo; k++) need = band_swz_rows(need, hc); }
        need = need.clipped(0, hc);
        rows_t missing = need.minus(valid[b][d]);
        if(missing.empty()) continue;
        if(is_src[b] == 1) { up[b][d].add(missing); valid[b][d].add(missing); continue; } // host source: uploaded straight from the caller's buffer
        for(int e = 0; e < n && !missing.empty(); e++)
        {
          if(e == d) continue;
          const rows_t take = missing.intersect(own[b][e]);   // from the device that computed them (always complete one step earlier)
          for(auto &iv : take.v)
          {
            st.pulls.push_back(band_pull_t{ e, b, iv.first, iv.second });
            const plan_img_t &im = l.conn[c];
            B->dev[d].pulled_bytes += (size_t)(iv.second - iv.first) * im.wd * im.chan * im.layers * (im.format == dt_token("f32") ? 4 : 2);
          }
          if(!take.empty())
          {
            if(std::find(st.waits.begin(), st.waits.end(), e) == st.waits.end()) st.waits.push_back(e);
            if(std::find(pulled_from[li][e].begin(), pulled_from[li][e].end(), d) == pulled_from[li][e].end()) pulled_from[li][e].push_back(d);
            missing = missing.minus(take);
            valid[b][d].add(take);
          }
        }
        if(!missing.empty())
        {
          bands_free(B);
          return vkb_set_error(VKB_ERR_GRAPH, "band split: launch %d (%s) on device slot %d needs rows [%d, %d) of connector %zu that no device computed",
              li, l.label.c_str(), d, missing.v[0].first, missing.v[0].second, c);
        }
      }
      // write after read: whoever pulled from this device in the step before must be done before this device moves on
      if(li > 0) for(int e : pulled_from[li - 1][d]) if(std::find(st.waits.begin(), st.waits.end(), e) == st.waits.end()) st.waits.push_back(e);
    }
    // writes
    for(int d = 0; d < n; d++) for(size_t c = 0; c < l.conn.size(); c++)
    {
      const band_acc_t &a = D.acc[c];
      const int b = l.conn[c].buf;
      if(b < 0 || (a.kind != BA_WRITE && a.kind != BA_WRITE_SWZ)) continue;
      const int hc = (int)l.conn[c].ht;
      const rows_t w = (a.kind == BA_WRITE ? comp[d] : band_swz_rows(comp[d], hc)).clipped(0, hc);
      valid[b][d].add(w); own[b][d].add(w);
    }
  }
  for(int d = 0; d < n; d++)
  {
    for(const plan_source_t &s : p->source) if(!s.external) for(auto &iv : up[s.buf_upload][d].v) B->dev[d].upload.push_back(band_copy_t{ s.buf_upload, iv.first, iv.second });
    for(const plan_sink_t &s : p->sink) if(s.buf >= 0) for(auto &iv : own[s.buf][d].v) B->dev[d].download.push_back(band_copy_t{ s.buf, iv.first, iv.second });
  }
  *out = B;
  return VKB_OK;
}

static int build_bands(dt_graph_t *g)
{
  vkb_plan_t *p = g->plan;
  vkb_bands_t *B = 0;
  const int r = plan_bands(g, &B);
  if(r) return r;
  const int n = (int)B->dev.size(), nl = (int)p->launch.size();
  // device side: pools, streams, events, peer access, workers
  for(int d = 0; d < n; d++)
  {
    band_dev_t &bd = B->dev[d];
    cudaError_t e = cudaSetDevice(bd.device);
    if(e == cudaSuccess) e = cudaMalloc(&bd.pool, p->pool_bytes);
    if(e != cudaSuccess) { const char *msg = cudaGetErrorString(e); bands_free(B); return vkb_set_error(VKB_ERR_OOM, "band split: pool of %zu bytes on device %d: %s", p->pool_bytes, bd.device, msg); }
    cudaStreamCreateWithFlags(&bd.stream, cudaStreamNonBlocking);
    bd.ev.resize(nl);
    for(cudaEvent_t &ev : bd.ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for(int o = 0; o < n; o++) if(B->dev[o].device != bd.device)
    {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, bd.device, B->dev[o].device);
      if(!can) { bands_free(B); return vkb_set_error(VKB_ERR_CUDA, "band split: device %d cannot access device %d (no NVLink / PCIe peer path)", bd.device, B->dev[o].device); }
      const cudaError_t pe = cudaDeviceEnablePeerAccess(B->dev[o].device, 0);
      if(pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { bands_free(B); return vkb_set_error(VKB_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe)); }
      cudaGetLastError();
    }
  }
  cudaSetDevice(B->dev[0].device);
  cudaEventCreate(&B->t0); cudaEventCreate(&B->t1);
  p->bands = B;
  for(int d = 0; d < n; d++) B->dev[d].worker = std::thread(band_worker, g, d);
  return VKB_OK;
}

static size_t band_row_bytes(const vkb_plan_t *p, int buf, uint32_t *layers, size_t *plane)
{ // geometry of a buffer as its launches see it
  for(const plan_launch_t &l : p->launch) for(const plan_img_t &im : l.conn) if(im.buf == buf)
  {
    const size_t rb = (size_t)im.wd * im.chan * (im.format == dt_token("f32") ? 4 : 2);
    *layers = im.layers; *plane = rb * im.ht;
    return rb;
  }
  *layers = 1; *plane = 0;
  return 0;
}

static void band_worker(dt_graph_t *g, int d)
{
  vkb_plan_t *p = g->plan;
  vkb_bands_t *B = p->bands;
  band_dev_t &me = B->dev[d];
  const int n = (int)B->dev.size(), nl = (int)p->launch.size();
  uint64_t seen = 0;
  cudaSetDevice(me.device);
  for(;;)
  {
    uint32_t run; uint64_t frame;
    {
      std::unique_lock<std::mutex> lk(B->mtx);
      B->cv.wait(lk, [&] { return B->quit || B->go > seen; });
      if(B->quit) return;
      seen = B->go; run = B->runflags; frame = B->frame;
    }
    me.rc = 0;
    // the previous frame's last pulls out of this pool must have landed before the pool is written again
    if(frame > 1) for(int e = 0; e < n; e++) if(e != d) cudaStreamWaitEvent(me.stream, B->dev[e].ev[nl - 1], 0);
    if(run & VKB_RUN_UPLOAD_SOURCE) for(const plan_source_t &s : p->source) if(!s.external)
    {
      const vkb_mem_source_t *ms = s.modid < (int)g->mem_source.size() && g->mem_source[s.modid].valid ? &g->mem_source[s.modid] : 0;
      if(!ms || s.packed_bpp) { me.rc = VKB_ERR_GRAPH; snprintf(me.err, sizeof(me.err), "band split: the source has to be a u16 mosaic in memory"); break; }
      const size_t rb = (size_t)ms->p.width * 2;
      for(const band_copy_t &c : me.upload) if(c.buf == s.buf_upload)
        cudaMemcpyAsync((uint8_t *)me.pool + p->buf[c.buf].offset + rb * c.r0, (const uint8_t *)ms->data + rb * c.r0, rb * (c.r1 - c.r0), cudaMemcpyHostToDevice, me.stream);
    }
    if(run & VKB_RUN_RECORD_CMD_BUF) for(int li = 0; li < nl && !me.rc; li++)
    {
      const plan_launch_t &l = p->launch[li];
      const band_step_t &st = me.step[li];
      for(int e : st.waits)
      { // the peer's record of its step li - 1 has to be issued before the wait on it is
        const uint64_t want = frame * 65536 + (uint64_t)li;
        while(B->issued[e].load(std::memory_order_acquire) < want) std::this_thread::yield();
        cudaStreamWaitEvent(me.stream, B->dev[e].ev[li - 1], 0);
      }
      for(const band_pull_t &q : st.pulls)
      {
        uint32_t layers; size_t plane;
        const size_t rb = band_row_bytes(p, q.buf, &layers, &plane);
        const size_t off = p->buf[q.buf].offset + rb * q.r0, bytes = rb * (q.r1 - q.r0);
        if(layers <= 1) cudaMemcpyAsync((uint8_t *)me.pool + off, (const uint8_t *)B->dev[q.src].pool + off, bytes, cudaMemcpyDefault, me.stream);
        else cudaMemcpy2DAsync((uint8_t *)me.pool + off, plane, (const uint8_t *)B->dev[q.src].pool + off, plane, bytes, layers, cudaMemcpyDefault, me.stream);
      }
      std::vector<vkb_image_t> conn;
      for(const plan_img_t &im : l.conn)
      {
        void *ptr = 0;
        if(im.buf >= 0) ptr = p->buf[im.buf].external ? p->buf[im.buf].external : (void *)((uint8_t *)me.pool + p->buf[im.buf].offset);
        conn.push_back(vkb_image_t{ ptr, im.wd, im.ht, im.chan, im.layers, im.format });
      }
      for(const auto &iv : st.compute)
      {
        const vkb_launch_t kl = { l.wd, l.ht, l.dp, l.push.data(), (uint32_t)l.push.size(), l.arg_params.data(), (uint32_t)l.arg_params.size(),
            conn.data(), (uint32_t)conn.size(), me.stream, iv.first, iv.second };
        const int rr = vkb_dispatch_launch(l.name, l.kernel, g->mode, &kl);
        if(rr) { me.rc = rr; snprintf(me.err, sizeof(me.err), "%s", vkb_last_error()); break; }
      }
      cudaEventRecord(me.ev[li], me.stream);
      B->issued[d].store(frame * 65536 + (uint64_t)li + 1, std::memory_order_release);
    }
    if(me.rc) B->issued[d].store(frame * 65536 + 65535, std::memory_order_release); // do not leave peers spinning
    if(!me.rc && (run & VKB_RUN_DOWNLOAD_SINK)) for(const plan_sink_t &s : p->sink)
    {
      const vkb_mem_sink_t *ms = s.modid < (int)g->mem_sink.size() && g->mem_sink[s.modid].valid ? &g->mem_sink[s.modid] : 0;
      if(!ms || !ms->dst) continue;
      const size_t rb = s.bytes / s.ht;
      for(const band_copy_t &c : me.download) if(c.buf == s.buf)
        cudaMemcpyAsync((uint8_t *)ms->dst + rb * c.r0, (const uint8_t *)me.pool + p->buf[c.buf].offset + rb * c.r0, rb * (c.r1 - c.r0), cudaMemcpyDeviceToHost, me.stream);
    }
    if(!me.rc && (run & VKB_RUN_WAIT_DONE))
    {
      const cudaError_t e = cudaStreamSynchronize(me.stream);
      if(e != cudaSuccess) { me.rc = VKB_ERR_CUDA; snprintf(me.err, sizeof(me.err), "band split: device %d: %s", me.device, cudaGetErrorString(e)); }
    }
    {
      std::lock_guard<std::mutex> lk(B->mtx);
      B->pending--;
    }
    B->cv.notify_all();
  }
}

static int run_bands(dt_graph_t *g, uint32_t run)
{
  vkb_plan_t *p = g->plan;
  vkb_bands_t *B = p->bands;
  {
    std::unique_lock<std::mutex> lk(B->mtx);
    B->frame++; B->go++; B->runflags = run; B->pending = (int)B->dev.size();
  }
  B->cv.notify_all();
  {
    std::unique_lock<std::mutex> lk(B->mtx);
    B->cv.wait(lk, [&] { return B->pending == 0; });
  }
  for(band_dev_t &d : B->dev) if(d.rc) return vkb_set_error(d.rc, "%s", d.err);
  return VKB_OK;
}

static void *buf_ptr(const vkb_plan_t *p, int b, int copy = 0)
{
  if(b < 0) return 0;
  if(p->buf[b].external) return p->buf[b].external;
  return (uint8_t *)p->pool + p->buf[b].offset + (p->buf[b].frames == 2 && copy ? p->buf[b].bytes : 0);
}

int dt_graph_run(dt_graph_t *g, uint32_t run)
{
  int r;
  if(vkb_device_count() <= 0) return vkb_set_error(VKB_ERR_NO_DEVICE, "no CUDA device: vkdt_b200 has no CPU fallback");
  // per-launch timing is a property of the graph (vkb_graph_set_perf, the reference's -d perf log mask) or an explicit flag
  // outside the reference's bits; s_graph_run_all (-1u, every bit set) never implies it
  const bool perf = g->perf || (run != (uint32_t)VKB_RUN_ALL && (run & VKB_RUN_PERF));
  const bool banded = g->band_devices.size() > 1;   // one frame over several GPUs (vkb_graph_set_bands)
  if((run & (VKB_RUN_ROI | VKB_RUN_CREATE_NODES | VKB_RUN_ALLOC)) || !g->plan || (banded ? !g->plan->bands : !g->plan->pool))
  {
    r = build_plan(g, !banded);
    if(r) return r;
    if(banded) { r = build_bands(g); if(r) { plan_free(g->plan); g->plan = 0; return r; } }
    run |= VKB_RUN_UPLOAD_SOURCE | VKB_RUN_RECORD_CMD_BUF;
  }
  vkb_plan_t *p = g->plan;
  if(banded)
  { // parameters are committed here, everything that touches a device happens in the per device workers
    if(run & VKB_RUN_RECORD_CMD_BUF)
    {
      for(int m : p->modid) if(g->module[m].so->commit_params) g->module[m].so->commit_params(g, &g->module[m]);
      for(plan_launch_t &l : p->launch)
      {
        l.arg_params.clear();
        for(int m : l.param_mods)
        {
          const dt_module_t *mod = &g->module[m];
          const uint8_t *src = mod->committed_param_size ? mod->committed_param : mod->param;
          const int sz = mod->committed_param_size ? mod->committed_param_size : mod->param_size;
          l.arg_params.insert(l.arg_params.end(), src, src + sz);
        }
      }
    }
    return run_bands(g, run);
  }
  cudaSetDevice(g->device);
  const int parity = g->frame & 1;
  // sources with s_module_request_read_source re-upload every run (graph-run-modules.h:573-587)
  // an upload on its own (no record, no download, no wait) is a prefetch: see up_stream above
  const bool prefetch = (run & VKB_RUN_UPLOAD_SOURCE) && !(run & (VKB_RUN_RECORD_CMD_BUF | VKB_RUN_DOWNLOAD_SINK | VKB_RUN_WAIT_DONE));
  const bool was_prefetched = p->src_prefetched && (run & VKB_RUN_RECORD_CMD_BUF) && !(run & VKB_RUN_UPLOAD_SOURCE);
  if(was_prefetched) { cudaStreamWaitEvent(p->stream, p->ev_src_ready, 0); p->src_prefetched = 0; }
  cudaStream_t up = p->stream;
  if(prefetch) { up = p->up_stream; cudaStreamWaitEvent(up, p->ev_src_free, 0); }
  else if((run & VKB_RUN_UPLOAD_SOURCE) && p->src_prefetched)
  { cudaStreamWaitEvent(p->stream, p->ev_src_ready, 0); p->src_prefetched = 0; }   // a later plain upload overwrites the prefetched one, in order
  for(const plan_source_t &s : p->source)
  {
    dt_module_t *mod = &g->module[s.modid];
    const bool want = (run & VKB_RUN_UPLOAD_SOURCE) || ((run & VKB_RUN_RECORD_CMD_BUF) && (mod->flags & s_module_request_read_source) && !was_prefetched);
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
        cudaMemcpyAsync(buf_ptr(p, s.buf_upload), ms->data, payload, cudaMemcpyHostToDevice, up);
        continue;
      }
    }
    if(!mod->so->read_source) return vkb_set_error(VKB_ERR_GRAPH, "source module %s has no read_source", dt_token_string(mod->name).c_str());
    if(prefetch) { cudaStreamSynchronize(up); up = p->stream; } // file sources go through the one staging buffer: in stream order
    cudaStreamSynchronize(p->stream); // staging is reused
    dt_read_source_params_t rp = { &g->node[s.nodeid], 0, 0 };
    if(mod->so->read_source(mod, p->staging_up, &rp)) return vkb_set_error(VKB_ERR_IO, "read_source failed for %s:%s", dt_token_string(mod->name).c_str(), dt_token_string(mod->inst).c_str());
    cudaMemcpyAsync(buf_ptr(p, s.buf_upload), p->staging_up, s.bytes, cudaMemcpyHostToDevice, p->stream);
  }
  if(prefetch && up == p->up_stream) { cudaEventRecord(p->ev_src_ready, up); p->src_prefetched = 1; }
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
      // graph->double_buffer of the reference's frame loop (graph-export.c:294): frame f writes copy f & 1
      for(const plan_img_t &im : l.conn) l.arg_conn.push_back(vkb_image_t{ buf_ptr(p, im.buf, im.fb ? 1 - parity : parity), im.wd, im.ht, im.chan, im.layers, im.format });
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
  if(run & VKB_RUN_RECORD_CMD_BUF) cudaEventRecord(p->ev_src_free, p->stream);   // the launches above were the source's last readers
  if(run & VKB_RUN_DOWNLOAD_SINK)
  {
    for(const plan_sink_t &s : p->sink)
    {
      dt_module_t *mod = &g->module[s.modid];
      const vkb_mem_sink_t *ms = s.modid < (int)g->mem_sink.size() && g->mem_sink[s.modid].valid ? &g->mem_sink[s.modid] : 0;
      if(ms && ms->dst)
      {
        if(ms->bytes < s.bytes) return vkb_set_error(VKB_ERR_BAD_ARG, "sink buffer too small: %zu < %zu", ms->bytes, s.bytes);
        cudaMemcpyAsync(ms->dst, buf_ptr(p, s.buf, parity), s.bytes, cudaMemcpyDeviceToHost, p->stream);
      }
      else if(mod->so->write_sink && !(ms && !ms->dst))
      {
        if(!p->staging_down && cudaHostAlloc(&p->staging_down, p->staging_down_bytes, cudaHostAllocDefault) != cudaSuccess)
          return vkb_set_error(VKB_ERR_OOM, "pinned download staging allocation failed");
        cudaMemcpyAsync(p->staging_down, buf_ptr(p, s.buf, parity), s.bytes, cudaMemcpyDeviceToHost, p->stream);
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
          bytes += (size_t)im.wd * im.ht * im.chan * im.layers * (im.format == dt_token("f32") ? 4 : (im.format == dt_token("ui8") ? 1 : 2));
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
      snprintf(b, sizeof(b), " b%d:%ux%ux%ux%u:%s@%zu%s", im.buf, im.wd, im.ht, im.chan, im.layers, dt_token_string(im.format).c_str(), p->buf[im.buf].offset,
          p->buf[im.buf].frames == 2 ? (im.fb ? "[fb:last frame]" : "[x2]") : "");
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

// host only: the band split as text, one line per launch and device slot (rows computed, rows pulled and from whom)
int dt_graph_band_plan(dt_graph_t *g, std::string *text)
{
  if(g->band_devices.size() < 2) return vkb_set_error(VKB_ERR_BAD_ARG, "no band split set (vkb_graph_set_bands)");
  int r = build_plan(g, false);
  if(r) return r;
  vkb_bands_t *B = 0;
  r = plan_bands(g, &B);
  if(r) return r;
  char b[256];
  const vkb_plan_t *p = g->plan;
  for(size_t li = 0; li < p->launch.size(); li++) for(size_t d = 0; d < B->dev.size(); d++)
  {
    const band_step_t &st = B->dev[d].step[li];
    snprintf(b, sizeof(b), "launch %2zu dev %zu %s_%s rows", li, d, dt_token_string(p->launch[li].name).c_str(), dt_token_string(p->launch[li].kernel).c_str());
    *text += b;
    for(auto &iv : st.compute) { snprintf(b, sizeof(b), " [%d,%d)", iv.first, iv.second); *text += b; }
    for(auto &q : st.pulls) { snprintf(b, sizeof(b), " pull b%d[%d,%d)<-%d", q.buf, q.r0, q.r1, q.src); *text += b; }
    for(int e : st.waits) { snprintf(b, sizeof(b), " wait %d", e); *text += b; }
    *text += "\n";
  }
  for(size_t d = 0; d < B->dev.size(); d++)
  {
    for(auto &c : B->dev[d].upload)   { snprintf(b, sizeof(b), "upload dev %zu b%d[%d,%d)\n", d, c.buf, c.r0, c.r1); *text += b; }
    for(auto &c : B->dev[d].download) { snprintf(b, sizeof(b), "download dev %zu b%d[%d,%d)\n", d, c.buf, c.r0, c.r1); *text += b; }
    snprintf(b, sizeof(b), "dev %zu pulls %zu bytes per frame\n", d, B->dev[d].pulled_bytes); *text += b;
  }
  bands_free(B);
  return VKB_OK;
}

// accessors used by the C-ABI
int vkb_plan_sink(dt_graph_t *g, int modid, uint32_t *wd, uint32_t *ht, void **dptr)
{
  if(!g->plan) return VKB_ERR_GRAPH;
  for(const plan_sink_t &s : g->plan->sink) if(s.modid == modid)
  {
    if(wd) *wd = s.wd; if(ht) *ht = s.ht;
    if(dptr) *dptr = g->plan->bands ? (void *)((uint8_t *)g->plan->bands->dev[0].pool + g->plan->buf[s.buf].offset) : buf_ptr(g->plan, s.buf);
    return VKB_OK;
  }
  return VKB_ERR_BAD_ARG;
}
uint64_t vkb_plan_pool_bytes(dt_graph_t *g) { return g->plan ? g->plan->pool_bytes : 0; }
int vkb_plan_launches(dt_graph_t *g) { return g->plan ? (int)g->plan->launch.size() : 0; }
void *vkb_plan_stream(dt_graph_t *g) { return g->plan ? (g->plan->bands ? (void *)g->plan->bands->dev[0].stream : (void *)g->plan->stream) : 0; }
// band split statistics: bytes pulled from peers per frame (sum over devices, and the largest single device)
int vkb_plan_band_stats(dt_graph_t *g, uint64_t *total, uint64_t *max_dev, int *pulls, int *launches)
{
  if(!g->plan || !g->plan->bands) return VKB_ERR_GRAPH;
  uint64_t t = 0, m = 0; int np = 0, nk = 0;
  for(const band_dev_t &d : g->plan->bands->dev)
  {
    t += d.pulled_bytes; m = std::max<uint64_t>(m, d.pulled_bytes);
    for(const band_step_t &st : d.step) { np += (int)st.pulls.size(); nk += (int)st.compute.size(); }
  }
  if(total) *total = t; if(max_dev) *max_dev = m; if(pulls) *pulls = np; if(launches) *launches = nk;
  return VKB_OK;
}
// device side time of a span of banded frames: mark(0) / mark(1) record an event on every device's stream, elapsed() is the
// largest per device span (all devices start together after a synchronising run)
int vkb_plan_band_mark(dt_graph_t *g, int which)
{
  if(!g->plan || !g->plan->bands) return VKB_ERR_GRAPH;
  vkb_bands_t *B = g->plan->bands;
  static thread_local std::vector<cudaEvent_t> dummy;
  for(band_dev_t &d : B->dev)
  {
    cudaSetDevice(d.device);
    if(d.ev.size() < (size_t)g->plan->launch.size() + 2)
    { // two timing events behind the per launch ones
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      d.ev.push_back(a); d.ev.push_back(b);
    }
    cudaEventRecord(d.ev[g->plan->launch.size() + (which ? 1 : 0)], d.stream);
  }
  return VKB_OK;
}
int vkb_plan_band_elapsed(dt_graph_t *g, float *ms)
{
  if(!g->plan || !g->plan->bands) return VKB_ERR_GRAPH;
  float mx = 0.0f;
  const size_t nl = g->plan->launch.size();
  for(band_dev_t &d : g->plan->bands->dev)
  {
    if(d.ev.size() < nl + 2) return VKB_ERR_GRAPH;
    cudaSetDevice(d.device);
    cudaEventSynchronize(d.ev[nl + 1]);
    float t = 0.0f;
    if(cudaEventElapsedTime(&t, d.ev[nl], d.ev[nl + 1]) != cudaSuccess) return VKB_ERR_CUDA;
    mx = std::max(mx, t);
  }
  *ms = mx;
  return VKB_OK;
}

// vkdt-cli compatible driver for the raw->display path: argument surface of src/cli/main.c:37-126, export flow of
// src/pipe/graph-export.c:108-325 (display -> sink module, colenc in front of 8 bit sinks / other colour spaces, frame loop),
// on top of the C-ABI only.  like the reference it writes o-jpg in sRGB primaries with the rec709 curve unless told otherwise
// (cli/main.c:58-59: trc = parse_prim("sRGB") = 1 = rec709, kept).
//   vkdt-b200-cli -g <graph.cfg> [--format o-jpg|o-pfm|o-null] [--filename out] [--output main] [--quality q]
//                 [--colour-prim sRGB|bt2020|AdobeRGB|P3|XYZ] [--colour-trc linear|709|sRGB|PQ|DCI|HLG|gamma2.2]
//                 [--device-id N] [-d perf|mem] [--dump-nodes] [--last-frame-only] [--progress]
//                 [--gpus N] [--bands] [--fast] [--config <cfg lines...>]
// ours: --gpus N develops a sequence frame parallel (frame f on GPU f mod N: one host thread and one graph per GPU, every
// frame writes its own file) or, with --bands, splits every frame into N horizontal bands with halo exchange over NVLink
// (large stills); --fast selects the fast build of the kernels.
#include "../../../include/vkdt_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

static int usage()
{
  fprintf(stderr, "usage: vkdt-b200-cli -g <graph.cfg>\n"
      "    [-d perf|mem]                 print per kernel timings / pool size\n"
      "    [--last-frame-only]           only write the last frame, not the intermediates\n"
      "    [--progress]                  print some progress information\n"
      "    [--dump-modules|--dump-nodes] write graphviz of the node layer to stdout\n"
      "    [--quality <0-100>]           (jpg) output quality\n"
      "    [--filename <f>]              output filename (without extension or frame number)\n"
      "    [--format <fm>]               output format (o-jpg, o-pfm, o-null)\n"
      "    [--width <x>]                 max output width\n"
      "    [--height <y>]                max output height\n"
      "    [--colour-prim <prim-id>]     colour primaries to use for encoding, one of: sRGB, bt2020, AdobeRGB, P3, XYZ\n"
      "    [--colour-trc <trc-id>]       tone response curve for encoding, one of: linear, 709, sRGB, PQ, DCI, HLG, gamma2.2\n"
      "    [--output <inst>]             name the instance of the display to replace (default: main)\n"
      "    [--device-id <gpu id>]        cuda device\n"
      "    [--gpus <n>]                  sequences: frame f on gpu f mod n; with --bands: every frame split into n bands\n"
      "    [--bands]                     band split over the gpus given by --gpus (stills of 100 MP and above)\n"
      "    [--fast]                      the fast build of the kernels (SFU exp / pow) instead of the strict one\n"
      "    [--config]                    everything after this will be interpreted as additional cfg lines\n");
  return 1;
}
static int parse_prim(const char *s)
{ // cli/main.c:13-23
  if(!strcasecmp(s, "sRGB")) return 1;
  if(!strcasecmp(s, "bt2020") || !strcasecmp(s, "2020")) return 2;
  if(!strcasecmp(s, "adobergb")) return 3;
  if(!strcasecmp(s, "P3")) return 4;
  if(!strcasecmp(s, "XYZ")) return 5;
  return 0xffff;
}
static int parse_trc(const char *s)
{ // cli/main.c:25-35 (its usage text also names bt709 and HLG)
  if(!strcasecmp(s, "linear")) return 0;
  if(!strcasecmp(s, "709") || !strcasecmp(s, "bt709")) return 1;
  if(!strcasecmp(s, "sRGB")) return 2;
  if(!strcasecmp(s, "PQ")) return 3;
  if(!strcasecmp(s, "DCI")) return 4;
  if(!strcasecmp(s, "HLG")) return 5;
  if(!strcasecmp(s, "gamma2.2")) return 6;
  return 0xffff;
}

struct opts_t
{
  const char *cfg = 0, *format = "o-jpg", *filename = "output", *inst = "main";
  int device = 0, perf = 0, mem = 0, dump = 0, last_only = 0, progress = 0, gpus = 1, bands = 0, fast = 0;
  int prim = 1, trc = 1; float quality = -1.0f;
  int max_width = 0, max_height = 0;
  int config_start = 0, argc = 0; char **argv = 0;
};

// dt_graph_export up to the first run: read the cfg, swap the display for the sink, extra config lines, output file name
static vkb_graph_t *make_graph(const opts_t &o, int device)
{
  vkb_graph_t *g = vkb_graph_new();
  vkb_graph_set_device(g, device);
  vkb_graph_set_perf(g, o.perf);
  if(o.fast) vkb_graph_set_mode(g, VKB_MODE_FAST);
  if(vkb_graph_read_config_ascii(g, o.cfg)) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); vkb_graph_free(g); return 0; }
  if(vkb_graph_replace_display_sized(g, o.inst, o.format, o.prim, o.trc, o.max_width, o.max_height)) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); vkb_graph_free(g); return 0; }
  if(o.config_start) for(int i = o.config_start; i < o.argc; i++) vkb_graph_read_config_line(g, o.argv[i]);
  char line[1024];
  snprintf(line, sizeof(line), "param:%s:%s:filename:%s", o.format, o.inst, o.filename);
  vkb_graph_read_config_line(g, line);
  if(o.quality >= 0.0f && !strcmp(o.format, "o-jpg"))
  { snprintf(line, sizeof(line), "param:o-jpg:%s:quality:%g", o.inst, o.quality); vkb_graph_read_config_line(g, line); }
  return g;
}
static void frame_name(vkb_graph_t *g, const opts_t &o, int f)
{ // graph-export.c:213-237: sequences number their files
  char fn[1024];
  snprintf(fn, sizeof(fn), "param:%s:%s:filename:%s_%04d", o.format, o.inst, o.filename, f);
  vkb_graph_read_config_line(g, fn);
}

int main(int argc, char *argv[])
{
  opts_t o; o.argc = argc; o.argv = argv;
  for(int i = 1; i < argc; i++)
  {
    if(!strcmp(argv[i], "-g") && i + 1 < argc) o.cfg = argv[++i];
    else if(!strcmp(argv[i], "-d") && i + 1 < argc) { i++; if(!strcmp(argv[i], "perf")) o.perf = 1; else if(!strcmp(argv[i], "mem")) o.mem = 1; }
    else if(!strcmp(argv[i], "-D") && i + 1 < argc) i++;
    else if(!strcmp(argv[i], "--dump-nodes") || !strcmp(argv[i], "--dump-modules")) o.dump = 1;
    else if(!strcmp(argv[i], "--format") && i + 1 < argc) o.format = argv[++i];
    else if(!strcmp(argv[i], "--filename") && i + 1 < argc) o.filename = argv[++i];
    else if(!strcmp(argv[i], "--output") && i + 1 < argc) o.inst = argv[++i];
    else if(!strcmp(argv[i], "--quality") && i + 1 < argc) o.quality = (float)atof(argv[++i]);
    else if(!strcmp(argv[i], "--colour-prim") && i + 1 < argc) o.prim = parse_prim(argv[++i]);
    else if(!strcmp(argv[i], "--colour-trc") && i + 1 < argc) o.trc = parse_trc(argv[++i]);
    else if((!strcmp(argv[i], "--device-id") || !strcmp(argv[i], "--device")) && i + 1 < argc) o.device = atoi(argv[++i]);
    else if(!strcmp(argv[i], "--gpus") && i + 1 < argc) o.gpus = atoi(argv[++i]);
    else if(!strcmp(argv[i], "--bands")) o.bands = 1;
    else if(!strcmp(argv[i], "--fast")) o.fast = 1;
    else if(!strcmp(argv[i], "--last-frame-only")) o.last_only = 1;
    else if(!strcmp(argv[i], "--progress")) o.progress = 1;
    else if(!strcmp(argv[i], "--width") && i + 1 < argc) o.max_width = (int)atof(argv[++i]);    // cli/main.c:68-71
    else if(!strcmp(argv[i], "--height") && i + 1 < argc) o.max_height = (int)atof(argv[++i]);
    else if(!strcmp(argv[i], "--audio") && i + 1 < argc) i++;
    else if(!strcmp(argv[i], "--config")) { o.config_start = i + 1; break; }
    else return usage();
  }
  if(!o.cfg) return usage();
  if(o.gpus < 1) o.gpus = 1;
  if(!o.dump && vkb_init(o.device)) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); return 2; }
  if(!o.dump && o.gpus > vkb_device_count()) { fprintf(stderr, "[cli] --gpus %d: only %d cuda devices\n", o.gpus, vkb_device_count()); return 2; }
  vkb_graph_t *g = make_graph(o, o.device);
  if(!g) return 3;
  std::vector<char> buf(1 << 20);
  if(o.dump)
  {
    if(vkb_graph_plan(g, buf.data(), buf.size())) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); return 5; }
    vkb_graph_dump_nodes(g, buf.data(), buf.size());
    fputs(buf.data(), stdout);
    vkb_graph_free(g);
    return 0;
  }
  if(o.bands && o.gpus > 1)
  {
    std::vector<int> devs(o.gpus);
    for(int d = 0; d < o.gpus; d++) devs[d] = d;
    vkb_graph_set_bands(g, o.gpus, devs.data());
  }
  // frame 0 decides how many frames there are (graph-export.c:251-268: graph->frame_cnt, set by a `frames:` line or by the source)
  int frames = vkb_graph_frame_count(g);
  vkb_graph_apply_keyframes(g);
  if(frames > 1) frame_name(g, o, 0);
  int flags0 = VKB_RUN_ALL;
  if(o.last_only && frames > 1) flags0 &= ~VKB_RUN_DOWNLOAD_SINK;
  int err = vkb_graph_run(g, flags0);
  if(err) { fprintf(stderr, "[cli] frame 0: %s\n", vkb_last_error()); vkb_graph_free(g); return 6; }
  if(o.perf && vkb_graph_perf(g, buf.data(), buf.size()) > 0) fputs(buf.data(), stdout);
  if(vkb_graph_frame_count(g) > frames)
  { // a source that only knows its length after the first run (i-mlv): frame 0 went to the unnumbered name, write it again
    frames = vkb_graph_frame_count(g);
    frame_name(g, o, 0);
    if(!o.last_only) vkb_graph_run(g, VKB_RUN_RECORD_CMD_BUF | VKB_RUN_DOWNLOAD_SINK | VKB_RUN_WAIT_DONE);
  }
  if(o.mem) printf("[mem] pooled HBM: %.1f MB\n", vkb_graph_pool_bytes(g) / 1e6);
  const int frame_flags = VKB_RUN_RECORD_CMD_BUF | VKB_RUN_UPLOAD_SOURCE | VKB_RUN_DOWNLOAD_SINK | VKB_RUN_WAIT_DONE;
  std::atomic<int> failed{0};
  std::mutex print_mtx;
  auto develop = [&](vkb_graph_t *gr, int f) {
    vkb_graph_set_frame(gr, f);
    vkb_graph_apply_keyframes(gr);   // graph-export.c:280-283
    frame_name(gr, o, f);
    int flags = frame_flags;
    if(o.last_only && f < frames - 1) flags &= ~VKB_RUN_DOWNLOAD_SINK;
    const int e = vkb_graph_run(gr, flags);
    std::lock_guard<std::mutex> lk(print_mtx);
    if(e) { fprintf(stderr, "[cli] frame %d: %s\n", f, vkb_last_error()); failed = 1; }
    if(o.progress) fprintf(stderr, "[cli] frame %d / %d\n", f + 1, frames);
    if(o.perf && !e && vkb_graph_perf(gr, buf.data(), buf.size()) > 0) fputs(buf.data(), stdout);
  };
  if(frames > 1 && o.gpus > 1 && !o.bands && vkb_graph_has_feedback(g))
  { fprintf(stderr, "[cli] the graph has feedback connectors (frames depend on each other): one gpu\n"); o.gpus = 1; }
  if(frames > 1 && o.gpus > 1 && !o.bands)
  { // frame parallel: frames are independent units, no exchange between the gpus.  gpu 0 keeps the graph
    // that developed frame 0; every other gpu builds its own from the same cfg (a graph belongs to one thread, graph.h:66-69)
    std::vector<std::thread> th;
    for(int d = 0; d < o.gpus; d++) th.emplace_back([&, d] {
      vkb_graph_t *gr = d == 0 ? g : make_graph(o, d);
      if(!gr) { failed = 1; return; }
      bool first = d != 0;
      for(int f = d == 0 ? o.gpus : d; f < frames && !failed; f += o.gpus)
      {
        if(first)
        { // this gpu's first frame builds its plan and pool
          vkb_graph_set_frame(gr, f); vkb_graph_apply_keyframes(gr); frame_name(gr, o, f);
          int fl = VKB_RUN_ALL; if(o.last_only && f < frames - 1) fl &= ~VKB_RUN_DOWNLOAD_SINK;
          if(vkb_graph_run(gr, fl)) { std::lock_guard<std::mutex> lk(print_mtx); fprintf(stderr, "[cli] frame %d: %s\n", f, vkb_last_error()); failed = 1; }
          first = false;
          continue;
        }
        develop(gr, f);
      }
      if(d != 0) vkb_graph_free(gr);
    });
    for(std::thread &t : th) t.join();
  }
  else for(int f = 1; f < frames && !failed; f++) develop(g, f);
  vkb_graph_free(g);
  vkb_cleanup();
  return failed ? 6 : 0;
}

// vkdt-cli compatible driver for the raw->display path (argument surface of src/cli/main.c:62-98, export flow of
// src/pipe/graph-export.c:108-325), on top of the C-ABI only.
//   vkdt-b200-cli -g <graph.cfg> [--format o-pfm] [--filename out] [--output main] [--config <cfg lines...>]
//                 [--device-id N] [-d perf|mem] [--dump-nodes] [--last-frame-only]
// export colour space is linear rec2020 f32 (o-pfm); colenc/resize/o-jpg are the next rows of SURVEY §8f.
#include "../../../include/vkdt_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

static int usage()
{
  fprintf(stderr, "usage: vkdt-b200-cli -g <graph.cfg>\n"
      "    [-d perf|mem]                 print per kernel timings / pool size\n"
      "    [--dump-nodes]                write graphviz of the node layer to stdout\n"
      "    [--format o-pfm]              output module (o-pfm, o-null)\n"
      "    [--filename <basename>]       output file basename (default: output)\n"
      "    [--output <inst>]             display instance to replace (default: main)\n"
      "    [--device-id <n>]             cuda device\n"
      "    [--last-frame-only]           for sequences: write the last frame only\n"
      "    [--config <cfg lines...>]     extra config lines, must come last\n");
  return 1;
}

int main(int argc, char *argv[])
{
  const char *cfg = 0, *format = "o-pfm", *filename = "output";
  int device = 0, perf = 0, mem = 0, dump = 0, last_only = 0, config_start = 0;
  for(int i = 1; i < argc; i++)
  {
    if(!strcmp(argv[i], "-g") && i + 1 < argc) cfg = argv[++i];
    else if(!strcmp(argv[i], "-d") && i + 1 < argc) { i++; if(!strcmp(argv[i], "perf")) perf = 1; else if(!strcmp(argv[i], "mem")) mem = 1; }
    else if(!strcmp(argv[i], "-D") && i + 1 < argc) i++;
    else if(!strcmp(argv[i], "--dump-nodes") || !strcmp(argv[i], "--dump-modules")) dump = 1;
    else if(!strcmp(argv[i], "--format") && i + 1 < argc) format = argv[++i];
    else if(!strcmp(argv[i], "--filename") && i + 1 < argc) filename = argv[++i];
    else if(!strcmp(argv[i], "--output") && i + 1 < argc) i++;  // only `main` exists on this path
    else if((!strcmp(argv[i], "--device-id") || !strcmp(argv[i], "--device")) && i + 1 < argc) device = atoi(argv[++i]);
    else if(!strcmp(argv[i], "--last-frame-only")) last_only = 1;
    else if((!strcmp(argv[i], "--width") || !strcmp(argv[i], "--height") || !strcmp(argv[i], "--quality") ||
             !strcmp(argv[i], "--colour-prim") || !strcmp(argv[i], "--colour-trc") || !strcmp(argv[i], "--audio")) && i + 1 < argc)
    {
      const char *flag = argv[i], *val = argv[++i];
      if((!strcmp(flag, "--colour-prim") && strcmp(val, "bt2020") && strcmp(val, "2020")) || (!strcmp(flag, "--colour-trc") && strcmp(val, "linear")))
        fprintf(stderr, "[cli] %s %s: only linear rec2020 export is built (no colenc module yet)\n", flag, val);
      else if(!strcmp(flag, "--width") || !strcmp(flag, "--height"))
        fprintf(stderr, "[cli] %s ignored: resize is not on the hot path\n", flag);
    }
    else if(!strcmp(argv[i], "--config")) { config_start = i + 1; break; }
    else if(!strcmp(argv[i], "--progress")) {}
    else return usage();
  }
  if(!cfg) return usage();
  if(!dump && vkb_init(device)) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); return 2; }
  vkb_graph_t *g = vkb_graph_new();
  vkb_graph_set_device(g, device);
  vkb_graph_set_perf(g, perf);
  if(vkb_graph_read_config_ascii(g, cfg)) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); return 3; }
  if(vkb_graph_replace_display(g, format)) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); return 4; }
  if(config_start) for(int i = config_start; i < argc; i++) vkb_graph_read_config_line(g, argv[i]);
  std::string line = std::string("param:") + format + ":main:filename:" + filename;
  vkb_graph_read_config_line(g, line.c_str());
  std::vector<char> buf(1 << 20);
  if(dump)
  {
    if(vkb_graph_plan(g, buf.data(), buf.size())) { fprintf(stderr, "[cli] %s\n", vkb_last_error()); return 5; }
    vkb_graph_dump_nodes(g, buf.data(), buf.size());
    fputs(buf.data(), stdout);
    vkb_graph_free(g);
    return 0;
  }
  // frame loop (graph-export.c:251-315): run_all for frame 0, then record + download per frame.  frames:N comes from
  // `--config frames:N` like in the reference (i-mlv learns its frame count after export tested it)
  int frames = 1;
  if(config_start) for(int i = config_start; i < argc; i++) if(!strncmp(argv[i], "frames:", 7)) frames = atoi(argv[i] + 7);
  if(frames < 1) frames = 1;
  int err = 0;
  for(int f = 0; f < frames && !err; f++)
  {
    vkb_graph_set_frame(g, f);
    if(frames > 1)
    {
      char fn[1024];
      snprintf(fn, sizeof(fn), "param:%s:main:filename:%s_%04d", format, filename, f);
      vkb_graph_read_config_line(g, fn);
    }
    int flags = f == 0 ? VKB_RUN_ALL : (VKB_RUN_RECORD_CMD_BUF | VKB_RUN_UPLOAD_SOURCE | VKB_RUN_DOWNLOAD_SINK | VKB_RUN_WAIT_DONE);
    if(last_only && f < frames - 1) flags &= ~VKB_RUN_DOWNLOAD_SINK;
    err = vkb_graph_run(g, flags);
    if(err) fprintf(stderr, "[cli] frame %d: %s\n", f, vkb_last_error());
    if(perf && !err && vkb_graph_perf(g, buf.data(), buf.size()) > 0) fputs(buf.data(), stdout);
  }
  if(mem) printf("[mem] pooled HBM: %.1f MB\n", vkb_graph_pool_bytes(g) / 1e6);
  vkb_graph_free(g);
  vkb_cleanup();
  return err ? 6 : 0;
}

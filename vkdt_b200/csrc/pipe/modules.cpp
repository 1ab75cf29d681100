// module classes of the raw->display path, statically registered (the reference scans modules/<name>/ and
// dlopens lib<name>.so, src/pipe/global.c:86-415; SURVEY.md §2: "callbacks may be statically registered").
// connector and param tables restate the modules' `connectors` / `params` files; the callbacks restate the
// host side main.c of each module (cited per function).  ROI arithmetic keeps the reference's types and
// operation order because integer sizes fall out of float math (crop/main.c:270-274).
#include "pipe.h"
#include "mlv.h"
#include "dng.h"
#include "jpeg.h"
#include <strings.h>
#include <math.h>
#include <stdlib.h>
#include <stdarg.h>
#include <map>

// ------------------------------------------------------------------------------------------------
// registry
struct module_def_t { const char *name; const char *connectors; const char *params; };

static const module_def_t g_defs[] = {
  { "i-raw",    "output:source:*:ui16", "filename:string:256:test.cr2\nnoise a:float:1:0.0\nnoise b:float:1:0.0\nstartid:int:1:0" },
  { "i-mlv",    "output:source:rggb:ui16", "filename:string:256:test.mlv" },
  { "i-pfm",    "output:source:rgba:f32", "filename:string:256:test.pfm\nstartid:int:1:0\nnoise a:float:1:0.0\nnoise b:float:1:0.0" },
  { "i-lut",    "output:source:*:*", "filename:string:256:test.lut" },
  { "denoise",  "input:read:*:*\noutput:write:&input:*",
                "strength:float:1:0.0\nluma:float:1:0.6\ndetail:float:1:1.0\npad:float:1:0\nedges:float:4:0:0:0:0\ngainmap:int:1:1" },
  { "hilite",   "input:read:*:*\noutput:write:&input:*", "white:float:1:0.985\ndesat:float:1:0.3\nsoft:float:1:0.6" },
  { "demosaic", "input:read:rggb:*\noutput:write:rgba:f16", "colour:int:1:0\nmethod:int:1:0" },
  { "crop",     "input:read:*:*\noutput:write:&input:f16",
                "perspect:float:8:0.25:0.25:0.75:0.25:0.75:0.75:0.25:0.75\ncrop:float:4:1.0:3.0:3.0:7.0\nrotate:float:1:1337" },
  { "colour",   "input:read:rgba:*\noutput:write:rgba:*\nclut:read:rg:f16\npicked:read:r:*\nabney:read:rg:f16\nspectra:read:rgba:*",
                "exposure:float:1:0.0\nsat:float:1:1.0\npicked:int:1:0\nmatrix:int:1:1\ngamut:int:1:0\nclip:int:1:0\nclipmax:float:1:1\n"
                "temp:float:1:6504\nwhite:float:4:0.0:0.0:0.0:0.0\nmat:float:9:-1.0\nmode:int:1:0\ncnt:int:1:4\n"
                "rbmap:float:144:0.3333:0.3333:0.3333:0.3333:0.3333:0.3333:0.5:0.25:0.25:0.5:0.25:0.25:0.25:0.5:0.25:0.25:0.5:0.25:0.25:0.25:0.5:0.25:0.25:0.5\n"
                "import:string:8:01" },
  { "filmcurv", "input:read:rgba:*\noutput:write:rgba:f16\ndspy:write:rgba:f16",
                "light:float:1:3\ncontrast:float:1:1.2\nbias:float:1:0.0\ncolour:int:1:3\nchroma:float:1:1.0\nrolloff:float:1:0.0\n"
                "red:float:1:0.0\nyellow:float:1:0.0\nblue:float:1:0.0\nshadows:float:1:0.0" },
  { "llap",     "input:read:rgba:f16\noutput:write:rgba:f16", "sigma:float:1:0.12\nshadows:float:1:1.0\nhilights:float:1:1.0\nclarity:float:1:0.0" },
  { "grade",    "input:read:*:*\noutput:write:*:*",
                "lift:float:4:0.0:0.0:0.0:0\ngamma:float:4:1.0:1.0:1.0:0\ngain:float:4:1.0:1.0:1.0:0\noffset:float:4:0.0:0.0:0.0:0\n"
                "mode:int:1:0\nsh_pivot:float:1:0.3\nhi_pivot:float:1:0.4" },
  { "colenc",   "input:read:rgba:*\noutput:write:rgba:*", "prim:int:1:1\ntrc:int:1:0" },
  { "resize",   "input:read:*:*\noutput:write:*:*", "width:int:1:0\nheight:int:1:0" },
  { "o-pfm",    "input:sink:rgba:f32", "filename:string:256:output" },
  { "o-jpg",    "input:sink:rgba:ui8", "filename:string:256:output\nquality:float:1:95\nexif:int:1:1" },
  { "o-null",   "input:sink:*:*", "" },
  { "display",  "input:sink:rgba:*", "" },
  // present in the default darkroom graph but never reachable from the exported sink: parse only
  { "hist",     "input:read:rgba:f16\noutput:write:rgba:*", "" },
  { "zones",    "input:read:rgba:f16\noutput:write:rgba:f16\ndspy:write:rgba:f16",
                "radius:float:1:0.01\nepsilon:float:1:0.06\ngamma:float:1:1.0\nnzones:int:1:3\nzone:float:7:0:0:0:0:0:0:0" },
  { "lens",     "input:read:*:*\noutput:write:*:*", "center:float:2:0.0:0.0\nscale:float:2:1.0:1.0\nsquish0:float:1:0.0\nsquish1:float:1:0.0\nca red:float:1:1.0\nca blue:float:1:1.0" },
  { "pick",     "input:sink:*:*\nspectra:read:rgba:f32\ndsp_:write:rgba:f16\npicked:write:r:atom",
                "nspots:int:1:0\npad:int:3:0\nspots:float:96:0\npicked:float:72:0\nref:float:72:0\nshow:int:1:0\ngrab:int:1:0\nde76:float:3:0\nfreeze:int:1:0" },
};

static dt_token_t read_token(const char *&c)
{ // asciiio.h:6-19
  char b[9] = {0};
  int i = 0;
  while(i < 8 && *c && *c != ':' && *c != '\n') b[i++] = *c++;
  while(*c && *c != ':' && *c != '\n') c++;  // tokens longer than 8 chars are truncated
  if(*c == ':' || *c == '\n') c++;
  return dt_token(b);
}
static int read_int(const char *&c)
{
  if(!*c) return 0;
  char *e; const int r = (int)strtol(c, &e, 10);
  if(*e && e != c) e++;
  c = e; return r;
}
static float read_float(const char *&c)
{
  if(!*c) return 0.0f;
  char *e; const float r = strtof(c, &e);
  if(*e && e != c) e++;
  c = e; return r;
}

static void parse_def(const module_def_t &d, dt_module_so_t *so)
{
  so->name = dt_token(d.name);
  const char *c = d.connectors;
  while(*c)
  { // global.c:27-38
    dt_connector_t cn;
    memset(&cn, 0, sizeof(cn));
    cn.name = read_token(c); cn.type = read_token(c); cn.chan = read_token(c); cn.format = read_token(c);
    cn.connected = s_cid_unset; cn.associated = s_cid_unset; cn.bypass = s_cid_unset; cn.buf = -1;
    if(cn.type == dt_token("write") || cn.type == dt_token("source")) cn.connected.i = 0; // reference counter
    so->connector.push_back(cn);
    while(*c == '\n') c++;
  }
  c = d.params;
  int offset = 0;
  while(*c)
  { // global.c:41-84, :147-150
    std::string line;
    while(*c && *c != '\n') line.push_back(*c++);
    while(*c == '\n') c++;
    const char *l = line.c_str();
    dt_ui_param_t p;
    p.name = read_token(l); p.type = read_token(l); p.cnt = read_int(l);
    p.offset = offset;
    if(p.type == dt_token("float"))
    { p.def.resize(4 * p.cnt); for(int i = 0; i < p.cnt; i++) { const float v = read_float(l); memcpy(&p.def[4 * i], &v, 4); } }
    else if(p.type == dt_token("int"))
    { p.def.resize(4 * p.cnt); for(int i = 0; i < p.cnt; i++) { const int v = read_int(l); memcpy(&p.def[4 * i], &v, 4); } }
    else
    { p.def.assign(p.cnt, 0); int i = 0; while(*l && i < p.cnt - 1) p.def[i++] = *l++; }
    offset += (int)p.def.size();
    so->param.push_back(p);
  }
}

// callbacks, defined below
#define CB(m) \
  int m##_init(dt_module_t *); void m##_cleanup(dt_module_t *); void m##_roi_out(dt_graph_t *, dt_module_t *); \
  void m##_roi_in(dt_graph_t *, dt_module_t *); void m##_create_nodes(dt_graph_t *, dt_module_t *); void m##_commit(dt_graph_t *, dt_module_t *);

static int  iraw_init(dt_module_t *);
static void iraw_cleanup(dt_module_t *);
static void iraw_roi_out(dt_graph_t *, dt_module_t *);
static int  iraw_read_source(dt_module_t *, void *, dt_read_source_params_t *);
static int  ipfm_init(dt_module_t *);
static void ipfm_cleanup(dt_module_t *);
static void ipfm_roi_out(dt_graph_t *, dt_module_t *);
static int  ipfm_read_source(dt_module_t *, void *, dt_read_source_params_t *);
static int  imlv_init(dt_module_t *);
static void imlv_cleanup(dt_module_t *);
static void imlv_roi_out(dt_graph_t *, dt_module_t *);
static int  imlv_read_source(dt_module_t *, void *, dt_read_source_params_t *);
static void denoise_roi_in(dt_graph_t *, dt_module_t *);
static void denoise_roi_out(dt_graph_t *, dt_module_t *);
static void denoise_create_nodes(dt_graph_t *, dt_module_t *);
static int  denoise_init(dt_module_t *);
static void denoise_cleanup(dt_module_t *);
static int  denoise_read_source(dt_module_t *, void *, dt_read_source_params_t *);
static void hilite_create_nodes(dt_graph_t *, dt_module_t *);
static void demosaic_roi_in(dt_graph_t *, dt_module_t *);
static void demosaic_roi_out(dt_graph_t *, dt_module_t *);
static void demosaic_create_nodes(dt_graph_t *, dt_module_t *);
static int  crop_init(dt_module_t *);
static void crop_roi_in(dt_graph_t *, dt_module_t *);
static void crop_roi_out(dt_graph_t *, dt_module_t *);
static void crop_commit(dt_graph_t *, dt_module_t *);
static int  ilut_init(dt_module_t *);
static void ilut_cleanup(dt_module_t *);
static void ilut_roi_out(dt_graph_t *, dt_module_t *);
static int  ilut_read_source(dt_module_t *, void *, dt_read_source_params_t *);
static int  colour_init(dt_module_t *);
static void colour_roi_in(dt_graph_t *, dt_module_t *);
static void colour_roi_out(dt_graph_t *, dt_module_t *);
static void colour_commit(dt_graph_t *, dt_module_t *);
static void colour_create_nodes(dt_graph_t *, dt_module_t *);
static void filmcurv_roi_out(dt_graph_t *, dt_module_t *);
static void filmcurv_create_nodes(dt_graph_t *, dt_module_t *);
static void llap_create_nodes(dt_graph_t *, dt_module_t *);
static void opfm_write_sink(dt_module_t *, void *, dt_write_sink_params_t *);
static void ojpg_write_sink(dt_module_t *, void *, dt_write_sink_params_t *);
static void colenc_roi_out(dt_graph_t *, dt_module_t *);
static void resize_roi_out(dt_graph_t *, dt_module_t *);
static void resize_roi_in(dt_graph_t *, dt_module_t *);
static void resize_create_nodes(dt_graph_t *, dt_module_t *);

// ---- module discovery (global.c:86-415, :442): <basedir>/modules/<name>/{connectors,params} are the source of truth when a
// vkdt installation (or checkout: basedir = <vkdt>/src/pipe) is named by vkb_set_basedir() / VKDT_B200_BASEDIR; the built-in
// tables above are the fallback for a stand-alone library.  callbacks stay statically registered (the reference dlopens
// lib<name>.so); a directory of a module that has no callbacks here registers with the defaults (one node "main", roi copied
// through), so that every cfg of the installation parses: such a module fails at run time only if a sink reaches it.
static std::string &basedir()
{
  static std::string d = getenv("VKDT_B200_BASEDIR") ? getenv("VKDT_B200_BASEDIR") : "";
  return d;
}
static bool read_text(const std::string &fn, std::string *out)
{
  FILE *f = fopen(fn.c_str(), "rb");
  if(!f) return false;
  char b[4096]; size_t n;
  out->clear();
  while((n = fread(b, 1, sizeof(b), f)) > 0) out->append(b, n);
  fclose(f);
  return true;
}
static std::vector<dt_module_so_t> &registry_storage() { static std::vector<dt_module_so_t> r; return r; }
// modules registered by a caller (vkb_register_module): name, connectors, params as the text of the reference's module files
struct external_module_t { std::string name, connectors, params; };
static std::vector<external_module_t> &external_modules() { static std::vector<external_module_t> e; return e; }
int dt_pipe_register_module(const char *name, const char *connectors, const char *params)
{
  if(!name || !name[0] || strlen(name) > 8 || !connectors) return 1;
  for(external_module_t &e : external_modules()) if(e.name == name) { e.connectors = connectors; e.params = params ? params : ""; registry_storage().clear(); return 0; }
  external_modules().push_back(external_module_t{ name, connectors, params ? params : "" });
  registry_storage().clear();   // rebuilt at the next lookup
  return 0;
}
static std::vector<std::string> &registry_strings() { static std::vector<std::string> s; return s; }
int dt_pipe_set_basedir(const char *dir)
{ // takes effect for graphs created afterwards
  basedir() = dir ? dir : "";
  registry_storage().clear();
  return 0;
}
#include <dirent.h>
static std::vector<dt_module_so_t> &registry()
{
  std::vector<dt_module_so_t> &r = registry_storage();
  if(!r.empty()) return r;
  std::vector<module_def_t> defs(g_defs, g_defs + sizeof(g_defs) / sizeof(g_defs[0]));
  std::vector<std::string> &keep = registry_strings();
  keep.clear(); keep.reserve(1024);
  if(!basedir().empty())
  {
    const std::string mdir = basedir() + "/modules";
    for(module_def_t &d : defs)
    { // a known module: the installation's files override the built-in tables
      std::string c, p;
      if(read_text(mdir + "/" + d.name + "/connectors", &c)) { keep.push_back(c); d.connectors = keep.back().c_str(); }
      else continue;
      if(read_text(mdir + "/" + d.name + "/params", &p)) { keep.push_back(p); d.params = keep.back().c_str(); } else d.params = "";
    }
    if(DIR *dp = opendir(mdir.c_str()))
    { // every other module directory: parse only
      while(struct dirent *ep = readdir(dp))
      {
        if(ep->d_name[0] == '.' || strlen(ep->d_name) > 8) continue;
        bool known = false;
        for(const module_def_t &d : defs) if(!strcmp(d.name, ep->d_name)) known = true;
        std::string c, p;
        if(known || !read_text(mdir + "/" + ep->d_name + "/connectors", &c)) continue;
        read_text(mdir + "/" + ep->d_name + "/params", &p);
        keep.push_back(ep->d_name); const char *nm = keep.back().c_str();
        keep.push_back(c); const char *cs = keep.back().c_str();
        keep.push_back(p); const char *ps = keep.back().c_str();
        defs.push_back(module_def_t{ nm, cs, ps });
      }
      closedir(dp);
    }
  }
  for(const external_module_t &e : external_modules())
  { // a caller's module replaces a built-in of the same name
    bool replaced = false;
    for(module_def_t &d : defs) if(e.name == d.name) { d.connectors = e.connectors.c_str(); d.params = e.params.c_str(); replaced = true; }
    if(!replaced) defs.push_back(module_def_t{ e.name.c_str(), e.connectors.c_str(), e.params.c_str() });
  }
  for(const module_def_t &d : defs)
  {
    dt_module_so_t so = {};
    parse_def(d, &so);
    const std::string n = d.name;
    if(n == "i-raw")    { so.init = iraw_init; so.cleanup = iraw_cleanup; so.modify_roi_out = iraw_roi_out; so.read_source = iraw_read_source; }
    if(n == "i-pfm")    { so.init = ipfm_init; so.cleanup = ipfm_cleanup; so.modify_roi_out = ipfm_roi_out; so.read_source = ipfm_read_source; }
    if(n == "i-lut")    { so.init = ilut_init; so.cleanup = ilut_cleanup; so.modify_roi_out = ilut_roi_out; so.read_source = ilut_read_source; }
    if(n == "i-mlv")    { so.init = imlv_init; so.cleanup = imlv_cleanup; so.modify_roi_out = imlv_roi_out; so.read_source = imlv_read_source; }
    if(n == "denoise")  { so.init = denoise_init; so.cleanup = denoise_cleanup; so.modify_roi_in = denoise_roi_in; so.modify_roi_out = denoise_roi_out;
                          so.create_nodes = denoise_create_nodes; so.read_source = denoise_read_source; }
    if(n == "hilite")   { so.create_nodes = hilite_create_nodes; }
    if(n == "demosaic") { so.modify_roi_in = demosaic_roi_in; so.modify_roi_out = demosaic_roi_out; so.create_nodes = demosaic_create_nodes; }
    if(n == "crop")     { so.init = crop_init; so.modify_roi_in = crop_roi_in; so.modify_roi_out = crop_roi_out; so.commit_params = crop_commit; }
    if(n == "colour")   { so.init = colour_init; so.modify_roi_in = colour_roi_in; so.modify_roi_out = colour_roi_out; so.commit_params = colour_commit; so.create_nodes = colour_create_nodes; }
    if(n == "filmcurv") { so.modify_roi_out = filmcurv_roi_out; so.create_nodes = filmcurv_create_nodes; }
    if(n == "llap")     { so.create_nodes = llap_create_nodes; }
    if(n == "o-pfm")    { so.write_sink = opfm_write_sink; }
    if(n == "o-jpg")    { so.write_sink = ojpg_write_sink; }
    if(n == "colenc")   { so.modify_roi_out = colenc_roi_out; }
    if(n == "resize")   { so.modify_roi_out = resize_roi_out; so.modify_roi_in = resize_roi_in; so.create_nodes = resize_create_nodes; }
    r.push_back(so);
  }
  return r;
}

dt_module_so_t *dt_module_so_get(dt_token_t name)
{
  for(dt_module_so_t &so : registry()) if(so.name == name) return &so;
  return 0;
}

int dt_module_so_describe(dt_token_t name, std::string *text)
{ // "connector <name>:<type>:<chan>:<format>" and "param <name>:<type>:<cnt>:<offset>:<default blob, hex>" lines
  const dt_module_so_t *so = dt_module_so_get(name);
  if(!so) return 1;
  char b[128];
  for(const dt_connector_t &c : so->connector)
  { *text += "connector " + dt_token_string(c.name) + ":" + dt_token_string(c.type) + ":" + dt_token_string(c.chan) + ":" + dt_token_string(c.format) + "\n"; }
  for(const dt_ui_param_t &p : so->param)
  {
    snprintf(b, sizeof(b), ":%d:%d:", p.cnt, p.offset);
    *text += "param " + dt_token_string(p.name) + ":" + dt_token_string(p.type) + b;
    for(uint8_t v : p.def) { snprintf(b, sizeof(b), "%02x", v); *text += b; }
    *text += "\n";
  }
  return 0;
}

static inline int param_id(const dt_module_t *m, const char *name) { return dt_module_get_param(m->so, dt_token(name)); }
#define CONN(A) do { int err_ = (A); if(err_) fprintf(stderr, "[vkdt_b200] %s:%d connection failed: error %d\n", __FILE__, __LINE__, err_); } while(0)

// ------------------------------------------------------------------------------------------------
// i-raw: in-memory source (already decoded u16 mosaic + dt_image_params_t, the contract rawler/rawspeed fulfil,
// i-raw/main.c:138-299).  file decoding of camera formats is out of scope.
static void fill_img_param(dt_module_t *mod, const vkb_raw_params_t *p)
{
  dt_image_params_t *ip = &mod->img_param;
  memset(ip, 0, sizeof(*ip));
  for(int k = 0; k < 4; k++) { ip->black[k] = p->black[k]; ip->white[k] = p->white[k]; ip->whitebalance[k] = p->whitebalance[k]; ip->crop_aabb[k] = p->crop_aabb[k]; }
  for(int k = 0; k < 9; k++) ip->cam_to_rec2020[k] = p->cam_to_rec2020[k];
  ip->filters = p->filters;
  ip->orientation = p->orientation;
  ip->noise_a = p->noise_a; ip->noise_b = p->noise_b;
  ip->colour_primaries = 0; // s_colour_primaries_custom: use cam_to_rec2020
  ip->colour_trc = 0;       // linear
}
// file source: uncompressed 16-bit cfa dng (pipe/dng.cpp); camera formats proper stay with rawler / rawspeed
struct iraw_file_t
{
  std::string filename; dng_image_t img; vkb_raw_params_t p; uint32_t ox = 0, oy = 0; bool loaded = false;
  // i-raw/main.c:62-92: OpcodeList2 decoded into what denoise reads, handed on through img_param.meta.  `file` from the dng,
  // `mem` for an in-memory source (vkb_graph_set_dng_opcodes: the tag's bytes as rawler hands them to the reference)
  dt_image_metadata_dngop_t dngop_file, dngop_mem; bool have_dngop_file = false, have_dngop_mem = false;
};
int dt_iraw_set_dng_opcodes(dt_module_t *mod, const void *blob, size_t len, int ox, int oy)
{
  if(!mod || mod->name != dt_token("i-raw") || !mod->data) return 1;
  iraw_file_t *d = (iraw_file_t *)mod->data;
  d->have_dngop_mem = false;
  if(!blob || !len) return 0;
  if(dng_opcode_list_decode((const uint8_t *)blob, len, &d->dngop_mem.list2)) return 2;
  d->dngop_mem.ox = ox; d->dngop_mem.oy = oy;
  d->have_dngop_mem = true;
  return 0;
}
static std::string resource_path(const dt_module_t *mod, const char *fname)
{ // dt_graph_get_resource_filename: relative to the cfg's directory first
  std::string path = fname;
  if(fname[0] != '/' && mod->graph->searchpath[0])
  {
    std::string p2 = std::string(mod->graph->searchpath) + "/" + fname;
    FILE *t = fopen(p2.c_str(), "rb");
    if(t) { fclose(t); path = p2; }
  }
  return path;
}
static int iraw_load(dt_module_t *mod)
{ // i-raw/main.c:60-100 (load_raw): decode once per filename
  iraw_file_t *d = (iraw_file_t *)mod->data;
  const char *fname = dt_module_param_string(mod, 0);
  if(d->loaded && d->filename == fname) return 0;
  d->loaded = false;
  const std::string path = resource_path(mod, fname);
  const size_t len = path.size();
  if(len < 4 || strcasecmp(path.c_str() + len - 4, ".dng"))
  { fprintf(stderr, "[i-raw] %s: only uncompressed dng files and in-memory sources are decoded here\n", path.c_str()); return 1; }
  const int err = dng_read(path.c_str(), &d->img);
  if(err) { fprintf(stderr, "[i-raw] failed to load raw file %s (%d)\n", path.c_str(), err); return 1; }
  if(dng_raw_params(&d->img, &d->p, &d->ox, &d->oy)) return 1;
  d->have_dngop_file = !d->img.opcode_list2.empty() && !dng_opcode_list_decode(d->img.opcode_list2.data(), d->img.opcode_list2.size(), &d->dngop_file.list2);
  d->dngop_file.ox = (int)d->ox; d->dngop_file.oy = (int)d->oy;
  d->filename = fname;
  d->loaded = true;
  return 0;
}
static int iraw_init(dt_module_t *mod) { mod->data = new iraw_file_t(); mod->flags = s_module_request_read_source; return 0; }
static void iraw_cleanup(dt_module_t *mod) { delete (iraw_file_t *)mod->data; mod->data = 0; }
static void iraw_roi_out(dt_graph_t *g, dt_module_t *mod)
{
  const int mid = (int)(mod - g->module.data());
  if(mid < 0 || mid >= (int)g->mem_source.size() || !g->mem_source[mid].valid)
  {
    if(iraw_load(mod)) return; // leaves full_wd == 0: graph run fails
    iraw_file_t *d = (iraw_file_t *)mod->data;
    fill_img_param(mod, &d->p);
    dt_image_params_t *ip = &mod->img_param;
    snprintf(ip->maker, sizeof(ip->maker), "%s", d->img.make);
    snprintf(ip->model, sizeof(ip->model), "%s", d->img.model);
    ip->iso = d->img.iso;
    // i-raw/main.c:172-199: noise profile from the parameters, else from nprof/<maker>-<model>-<iso>.nprof
    const float na = dt_module_param_float(mod, 1)[0], nb = dt_module_param_float(mod, 2)[0];
    ip->noise_a = na; ip->noise_b = nb;
    if(na == 0.0f && nb == 0.0f)
    {
      char pname[512];
      snprintf(pname, sizeof(pname), "nprof/%s-%s-%d.nprof", ip->maker, ip->model, (int)ip->iso);
      FILE *f = fopen(resource_path(mod, pname).c_str(), "rb");
      if(f) { float a = 0.0f, b = 0.0f; if(fscanf(f, "%g %g", &a, &b) == 2) { ip->noise_a = a; ip->noise_b = b; } fclose(f); }
    }
    mod->connector[0].roi.full_wd = d->p.width;  // already rounded to the cfa block
    mod->connector[0].roi.full_ht = d->p.height;
    mod->connector[0].chan = dt_token("rggb");
    ip->meta = d->have_dngop_file ? &d->dngop_file : 0;
    return;
  }
  const vkb_raw_params_t *p = &g->mem_source[mid].p;
  fill_img_param(mod, p);
  { iraw_file_t *d = (iraw_file_t *)mod->data; mod->img_param.meta = d && d->have_dngop_mem ? &d->dngop_mem : 0; }
  // i-raw/main.c:156-157: dimensions rounded down to the cfa block
  const int block = p->filters == 9u ? 3 : (p->filters ? 2 : 1);
  mod->connector[0].roi.full_wd = (p->width / block) * block;
  mod->connector[0].roi.full_ht = (p->height / block) * block;
  // i-raw/main.c:221-226: one channel for any cfa (bayer and x-trans), rgba only for 3-component raws
  mod->connector[0].chan = p->filters ? dt_token("rggb") : dt_token("rgba");
}
static int iraw_read_source(dt_module_t *mod, void *mapped, dt_read_source_params_t *p)
{ // i-raw/main.c:281-297: row copy of the aligned window into the mapped staging buffer
  dt_graph_t *g = mod->graph;
  const int mid = (int)(mod - g->module.data());
  if(mid >= (int)g->mem_source.size() || !g->mem_source[mid].valid)
  { // file: window starting at the cfa offset (i-raw/main.c:283-288)
    if(iraw_load(mod)) return 1;
    const iraw_file_t *d = (const iraw_file_t *)mod->data;
    const uint32_t wd = mod->connector[0].roi.full_wd, ht = mod->connector[0].roi.full_ht;
    for(uint32_t j = 0; j < ht; j++)
      memcpy((uint16_t *)mapped + (size_t)j * wd, d->img.pix.data() + (size_t)(j + d->oy) * d->img.width + d->ox, sizeof(uint16_t) * wd);
    return 0;
  }
  if(g->mem_source[mid].on_device) return 1;
  const vkb_raw_params_t *rp = &g->mem_source[mid].p;
  const uint32_t wd = mod->connector[0].roi.full_wd, ht = mod->connector[0].roi.full_ht;
  const uint16_t *src = (const uint16_t *)g->mem_source[mid].data;
  if(wd == rp->width) memcpy(mapped, src, sizeof(uint16_t) * (size_t)wd * ht);
  else for(uint32_t j = 0; j < ht; j++) memcpy((uint16_t *)mapped + (size_t)j * wd, src + (size_t)j * rp->width, sizeof(uint16_t) * wd);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// i-pfm (i-pfm/main.c:38-166): rgb ("PF") or single channel ("Pf") float images, e.g. an intermediate image another
// vkdt wrote with o-pfm: lets every stage of the path be run and compared in isolation.
// the executor converts the uploaded f32 image to the f16 edge format the kernels read (b200:cvt16).
struct ipfm_t { std::string filename; FILE *f = 0; uint32_t width = 0, height = 0; int channels = 3; long data_begin = 0; };
static int ipfm_init(dt_module_t *mod) { mod->data = new ipfm_t(); return 0; }
static void ipfm_cleanup(dt_module_t *mod)
{
  ipfm_t *p = (ipfm_t *)mod->data;
  if(p) { if(p->f) fclose(p->f); delete p; }
  mod->data = 0;
}
static int ipfm_read_header(dt_module_t *mod)
{ // i-pfm/main.c:38-101
  ipfm_t *p = (ipfm_t *)mod->data;
  const char *fname = dt_module_param_string(mod, 0);
  if(p->f && p->filename == fname) return 0;
  if(p->f) fclose(p->f);
  p->filename.clear();
  p->f = fopen(resource_path(mod, fname).c_str(), "rb");
  if(!p->f) { fprintf(stderr, "[i-pfm] could not load file `%s'!\n", fname); return 1; }
  int wd = 0, ht = 0;
  p->channels = 3;
  if(fscanf(p->f, "PF\n%d %d\n%*[^\n]", &wd, &ht) != 2)
  {
    rewind(p->f);
    if(fscanf(p->f, "Pf\n%d %d\n%*[^\n]", &wd, &ht) != 2 || wd <= 0 || ht <= 0) { fclose(p->f); p->f = 0; fprintf(stderr, "[i-pfm] `%s' is not a pfm file\n", fname); return 1; }
    p->channels = 1;
  }
  if(wd <= 0 || ht <= 0 || fgetc(p->f) == EOF) { fclose(p->f); p->f = 0; return 1; }
  p->width = wd; p->height = ht;
  p->data_begin = ftell(p->f);
  p->filename = fname;
  return 0;
}
static void ipfm_roi_out(dt_graph_t *g, dt_module_t *mod)
{
  dt_image_params_t *ip = &mod->img_param;
  if(ipfm_read_header(mod))
  { // i-pfm/main.c:147-152: a 32x32 placeholder keeps the graph alive in the gui; a cli run has nothing to show for it
    mod->connector[0].roi.full_wd = 0; mod->connector[0].roi.full_ht = 0;
    return;
  }
  const ipfm_t *p = (const ipfm_t *)mod->data;
  memset(ip, 0, sizeof(*ip));
  for(int k = 0; k < 4; k++) { ip->black[k] = 0.0f; ip->white[k] = 65535.0f; ip->whitebalance[k] = 1.0f; }
  ip->crop_aabb[2] = p->width; ip->crop_aabb[3] = p->height;
  ip->noise_a = dt_module_param_float(mod, 2)[0];
  ip->noise_b = dt_module_param_float(mod, 3)[0];
  ip->filters = 0;
  ip->colour_primaries = 2; // s_colour_primaries_2020
  ip->colour_trc = 0;       // linear; cam_to_rec2020 stays zero like in the reference (i-pfm/main.c:69-83): rec2020 input does not use it
  mod->connector[0].chan = p->channels == 1 ? dt_token("y") : dt_token("rgba");
  mod->connector[0].roi.full_wd = p->width;
  mod->connector[0].roi.full_ht = p->height;
}
static int ipfm_read_source(dt_module_t *mod, void *mapped, dt_read_source_params_t *)
{ // i-pfm/main.c:103-118 (read_plain): rgb -> rgba with alpha 1, rows in file order
  if(ipfm_read_header(mod)) return 1;
  ipfm_t *p = (ipfm_t *)mod->data;
  fseek(p->f, p->data_begin, SEEK_SET);
  float *out = (float *)mapped;
  const size_t n = (size_t)p->width * p->height;
  if(p->channels == 1) return fread(out, sizeof(float), n, p->f) == n ? 0 : 1;
  std::vector<float> row((size_t)p->width * 3);
  for(uint32_t j = 0; j < p->height; j++)
  {
    if(fread(row.data(), sizeof(float), row.size(), p->f) != row.size()) return 1;
    float *o = out + (size_t)4 * j * p->width;
    for(uint32_t i = 0; i < p->width; i++) { o[4*i] = row[3*i]; o[4*i+1] = row[3*i+1]; o[4*i+2] = row[3*i+2]; o[4*i+3] = 1.0f; }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// i-lut (i-lut/main.c:23-263, core/lut.h): the small tables colour reads (camera clut, abney, spectra), one .lut file each:
// { u32 magic 1234, u16 version 2, u8 channels, u8 datatype (0 f16, 1 f32), u32 wd, u32 ht } + texels.  single files only
// (lists of textures and ssbo payloads belong to other modules' inputs)
struct ilut_t { std::string filename; FILE *f = 0; uint32_t wd = 0, ht = 0; int channels = 0, datatype = 0; long data_begin = 0; };
static int ilut_init(dt_module_t *mod) { mod->data = new ilut_t(); return 0; }
static void ilut_cleanup(dt_module_t *mod)
{
  ilut_t *p = (ilut_t *)mod->data;
  if(p) { if(p->f) fclose(p->f); delete p; }
  mod->data = 0;
}
static std::string ilut_expand(const dt_module_t *mod, const char *pattern)
{ // ${maker} ${model} ${flen} of the main input (i-lut/main.c:33-37, core/strexpand.h)
  char flen[16];
  snprintf(flen, sizeof(flen), "%g", mod->graph->main_img_param.focal_length);
  const char *key[] = { "${maker}", "${model}", "${flen}" };
  const char *val[] = { mod->graph->main_img_param.maker, mod->graph->main_img_param.model, flen };
  std::string out = pattern;
  for(int k = 0; k < 3; k++)
    for(size_t pos = out.find(key[k]); pos != std::string::npos; pos = out.find(key[k], pos + strlen(val[k])))
      out.replace(pos, strlen(key[k]), val[k]);
  return out;
}
static int ilut_read_header(dt_module_t *mod)
{
  ilut_t *p = (ilut_t *)mod->data;
  const std::string fname = ilut_expand(mod, dt_module_param_string(mod, 0));
  if(p->f && p->filename == fname) return 0;
  if(p->f) fclose(p->f);
  p->f = 0; p->filename.clear();
  std::string path = resource_path(mod, fname.c_str());
  p->f = fopen(path.c_str(), "rb");
  if(!p->f && fname[0] != '/' && !basedir().empty()) p->f = fopen((basedir() + "/" + fname).c_str(), "rb"); // dt_graph_open_resource: the installation's data
  if(!p->f) { fprintf(stderr, "[i-lut] %s could not load file `%s'!\n", dt_token_string(mod->inst).c_str(), fname.c_str()); return 1; }
  uint8_t h[16];
  uint32_t magic, wd, ht; uint16_t version;
  if(fread(h, 1, 16, p->f) != 16) { fclose(p->f); p->f = 0; return 1; }
  memcpy(&magic, h, 4); memcpy(&version, h + 4, 2); memcpy(&wd, h + 8, 4); memcpy(&ht, h + 12, 4);
  if(magic != 1234 || version != 2 || !wd || !ht || wd > 65536 || ht > 65536 || h[6] < 1 || h[6] > 4 || h[7] > 1)
  {
    fprintf(stderr, "[i-lut] `%s' is not a lut file this path reads (magic %u version %u channels %u datatype %u)\n", fname.c_str(), magic, version, h[6], h[7]);
    fclose(p->f); p->f = 0; return 1;
  }
  { // the payload the header promises has to be there (a truncated or hostile file must not size a buffer)
    fseek(p->f, 0, SEEK_END);
    const long have = ftell(p->f);
    const uint64_t need = 16 + (uint64_t)wd * ht * h[6] * (h[7] == 0 ? 2 : 4);
    if(have < 0 || (uint64_t)have < need)
    {
      fprintf(stderr, "[i-lut] `%s' is truncated: %ld bytes, its header asks for %llu\n", fname.c_str(), have, (unsigned long long)need);
      fclose(p->f); p->f = 0; return 1;
    }
  }
  p->wd = wd; p->ht = ht; p->channels = h[6]; p->datatype = h[7];
  p->data_begin = 16;
  p->filename = fname;
  return 0;
}
static void ilut_roi_out(dt_graph_t *, dt_module_t *mod)
{
  if(ilut_read_header(mod)) { mod->connector[0].roi.full_wd = 0; mod->connector[0].roi.full_ht = 0; return; } // no 32x32 placeholder: fail at planning
  const ilut_t *p = (const ilut_t *)mod->data;
  mod->connector[0].roi.full_wd = p->wd;
  mod->connector[0].roi.full_ht = p->ht;
  mod->connector[0].chan = p->channels == 1 ? dt_token("r") : p->channels == 2 ? dt_token("rg") : dt_token("rgba");
  mod->connector[0].format = p->datatype == 0 ? dt_token("f16") : dt_token("f32");
  dt_image_params_t *ip = &mod->img_param;
  for(int k = 0; k < 4; k++) { ip->black[k] = 0.0f; ip->white[k] = 1.0f; ip->whitebalance[k] = 1.0f; }
  ip->colour_primaries = 2; ip->colour_trc = 0; ip->filters = 0;
}
int dt_module_source_failed(const dt_module_t *mod)
{ // a file source whose header could not be read (the roi pass keeps such a graph alive with a placeholder size)
  if(mod->name == dt_token("i-lut") && mod->data) return ((const ilut_t *)mod->data)->f == 0;
  if(mod->name == dt_token("i-pfm") && mod->data) return ((const ipfm_t *)mod->data)->f == 0;
  return 0;
}
static int ilut_read_source(dt_module_t *mod, void *mapped, dt_read_source_params_t *)
{ // i-lut/main.c:87-113 (read_plain): three channels are padded to four (the pad is all ones bits there; here a proper 1.0)
  if(ilut_read_header(mod)) return 1;
  ilut_t *p = (ilut_t *)mod->data;
  fseek(p->f, p->data_begin, SEEK_SET);
  const size_t sz = p->datatype == 0 ? 2 : 4, n = (size_t)p->wd * p->ht;
  if(p->channels != 3) return fread(mapped, sz * p->channels, n, p->f) == n ? 0 : 1;
  uint8_t *o = (uint8_t *)mapped;
  const uint16_t one16 = 0x3c00; const float one32 = 1.0f;
  for(size_t k = 0; k < n; k++)
  {
    if(fread(o + 4 * sz * k, sz, 3, p->f) != 3) return 1;
    memcpy(o + sz * (4 * k + 3), sz == 2 ? (const void *)&one16 : (const void *)&one32, sz);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// i-mlv (i-mlv/main.c:60-229): file or in-memory packed frames
static int imlv_init(dt_module_t *mod)
{
  mod->data = new mlv_clip_t();
  mod->flags = s_module_request_read_source;
  return 0;
}
static void imlv_cleanup(dt_module_t *mod)
{
  mlv_clip_t *c = (mlv_clip_t *)mod->data;
  if(c) { mlv_close(c); delete c; }
  mod->data = 0;
}
static int imlv_open(dt_module_t *mod)
{
  mlv_clip_t *c = (mlv_clip_t *)mod->data;
  const char *fname = dt_module_param_string(mod, 0);
  if(c->file && c->filename == fname) return 0;
  std::string path = fname;
  if(fname[0] != '/' && mod->graph->searchpath[0])
  { // dt_graph_get_resource_filename: relative to the cfg's directory first
    std::string p2 = std::string(mod->graph->searchpath) + "/" + fname;
    FILE *t = fopen(p2.c_str(), "rb");
    if(t) { fclose(t); path = p2; }
  }
  if(mlv_open(c, path.c_str())) return 1;
  c->filename = fname;
  return 0;
}
// dt_dcraw_adobe_coeff (i-mlv/adobe_coeff.h): camera name -> xyz_to_cam / 10000, from the table of the vkdt installation named
// with vkb_set_basedir() (<basedir>/modules/i-mlv/adobe_coeff.h, parsed as text: { "name", { n, n, ... } }).  1: not found
static int imlv_adobe_coeff(const char *name, float *xyz_to_cam)
{
  for(int j = 0; j < 12; j++) xyz_to_cam[j] = -1.0f;
  if(!name[0] || basedir().empty()) return 1;
  std::string text;
  if(!read_text(basedir() + "/modules/i-mlv/adobe_coeff.h", &text)) return 1;
  size_t pos = 0;
  while((pos = text.find("{ \"", pos)) != std::string::npos)
  {
    const size_t n0 = pos + 3, n1 = text.find('"', n0);
    if(n1 == std::string::npos) break;
    pos = n1;
    if(strcasecmp(text.substr(n0, n1 - n0).c_str(), name)) continue;
    const size_t b0 = text.find('{', n1), b1 = text.find('}', n1);
    if(b0 == std::string::npos || b1 == std::string::npos || b1 < b0) return 1;
    short trans[12] = { 0 };
    const char *c = text.c_str() + b0 + 1;
    for(int j = 0; j < 12 && c < text.c_str() + b1; j++)
    {
      char *e; trans[j] = (short)strtol(c, &e, 10);
      if(e == c) break;
      c = e; while(*c == ',' || *c == ' ') c++;
    }
    for(int j = 0; j < 12; j++) xyz_to_cam[j] = (float)(trans[j] / 10000.0);
    return 0;
  }
  return 1;
}
static void imlv_roi_out(dt_graph_t *g, dt_module_t *mod)
{
  const int mid = (int)(mod - g->module.data());
  if(mid < (int)g->mem_source.size() && g->mem_source[mid].valid)
  { // in-memory clip frame (bench / frame-parallel driver feeds packed payloads directly)
    const vkb_raw_params_t *p = &g->mem_source[mid].p;
    fill_img_param(mod, p);
    mod->connector[0].roi.full_wd = p->width;
    mod->connector[0].roi.full_ht = p->height;
    return;
  }
  if(imlv_open(mod)) return;
  mlv_clip_t *c = (mlv_clip_t *)mod->data;
  mod->connector[0].roi.full_wd = c->width;
  mod->connector[0].roi.full_ht = c->height;
  dt_image_params_t *ip = &mod->img_param;
  memset(ip, 0, sizeof(*ip));
  const float b = c->black, w = c->white;
  for(int k = 0; k < 4; k++) { ip->black[k] = b; ip->white[k] = w; ip->whitebalance[k] = 1.0f; }
  ip->filters = 0x5d5d5d5d; // i-mlv/main.c:127
  ip->crop_aabb[2] = c->width; ip->crop_aabb[3] = c->height;
  // :165-201: camera rgb -> xyz comes from dcraw's adobe_coeff table by camera name (xyz_to_cam, inverted); a camera that is
  // not in it gets the identity, the white balance is the row sums of cam_to_xyz normalised to green, and the matrix handed on
  // is xyz_to_rec2020 * cam_to_xyz.  the table is the vkdt installation's own file (imlv_adobe_coeff below): without a
  // basedir every clip takes the identity branch (a named camera is told so once).
  static const float xyz_to_rec2020[9] = {
     1.7166511880f, -0.3556707838f, -0.2533662814f,
    -0.6666843518f,  1.6164812366f,  0.0157685458f,
     0.0176398574f, -0.0427706133f,  0.9421031212f };
  float xyz_to_cam[12], mat[9] = { 0 };
  const int unknown = imlv_adobe_coeff(c->camera_name, xyz_to_cam);
  if(unknown) mat[0] = mat[4] = mat[8] = 1.0f;
  else
  { // core/mat3.h:21-46 mat3inv, operation for operation
    const float *A = xyz_to_cam;
#define A_(y, x) A[(y - 1) * 3 + (x - 1)]
    const float det = A_(1, 1) * (A_(3, 3) * A_(2, 2) - A_(3, 2) * A_(2, 3)) - A_(2, 1) * (A_(3, 3) * A_(1, 2) - A_(3, 2) * A_(1, 3)) + A_(3, 1) * (A_(2, 3) * A_(1, 2) - A_(2, 2) * A_(1, 3));
    if(!(det > -1e-7f && det < 1e-7f))
    {
      const float idet = 1.f / det;
      mat[0] =  idet * (A_(3, 3) * A_(2, 2) - A_(3, 2) * A_(2, 3)); mat[1] = -idet * (A_(3, 3) * A_(1, 2) - A_(3, 2) * A_(1, 3)); mat[2] =  idet * (A_(2, 3) * A_(1, 2) - A_(2, 2) * A_(1, 3));
      mat[3] = -idet * (A_(3, 3) * A_(2, 1) - A_(3, 1) * A_(2, 3)); mat[4] =  idet * (A_(3, 3) * A_(1, 1) - A_(3, 1) * A_(1, 3)); mat[5] = -idet * (A_(2, 3) * A_(1, 1) - A_(2, 1) * A_(1, 3));
      mat[6] =  idet * (A_(3, 2) * A_(2, 1) - A_(3, 1) * A_(2, 2)); mat[7] = -idet * (A_(3, 2) * A_(1, 1) - A_(3, 1) * A_(1, 2)); mat[8] =  idet * (A_(2, 2) * A_(1, 1) - A_(2, 1) * A_(1, 2));
    }
#undef A_
  }
  const double cam_to_xyz[9] = { mat[0], mat[1], mat[2], mat[3], mat[4], mat[5], mat[6], mat[7], mat[8] };
  ip->whitebalance[0] = (float)(cam_to_xyz[0] + cam_to_xyz[1] + cam_to_xyz[2]);
  ip->whitebalance[1] = (float)(cam_to_xyz[3] + cam_to_xyz[4] + cam_to_xyz[5]);
  ip->whitebalance[2] = (float)(cam_to_xyz[6] + cam_to_xyz[7] + cam_to_xyz[8]);
  ip->whitebalance[0] /= ip->whitebalance[1];
  ip->whitebalance[2] /= ip->whitebalance[1];
  ip->whitebalance[1]  = 1.0f;
  float cam_to_rec2020[9] = { 0.0f };
  for(int j = 0; j < 3; j++) for(int i = 0; i < 3; i++) for(int k = 0; k < 3; k++)
    cam_to_rec2020[3 * j + i] += xyz_to_rec2020[3 * j + k] * cam_to_xyz[3 * k + i];
  for(int k = 0; k < 9; k++) ip->cam_to_rec2020[k] = cam_to_rec2020[k];
  static int told = 0;
  if(unknown && c->camera_name[0] && basedir().empty() && !told++)
    fprintf(stderr, "[i-mlv] no vkdt installation named (vkb_set_basedir): `%s' is developed with camera rgb = xyz\n", c->camera_name);
  ip->noise_a = 1.0f; ip->noise_b = 1.0f; // :140-141; the nprof lookup is outside the hot path
  snprintf(ip->model, sizeof(ip->model), "%s", c->camera_name);
  snprintf(ip->maker, sizeof(ip->maker), "%s", c->camera_name);
  for(size_t i = 0; i < sizeof(ip->maker); i++) if(ip->maker[i] == ' ') ip->maker[i] = 0;
  g->frame_cnt = c->frame_count;               // :150
  g->frame_rate = c->fps_denom ? c->fps_nom / (double)c->fps_denom : 24.0;
}
// writes the PACKED payload (bpp/8 bytes per pixel + padding) into mapped; the device unpacks (kernels/k_raw.cu)
static int imlv_read_source(dt_module_t *mod, void *mapped, dt_read_source_params_t *p)
{
  dt_graph_t *g = mod->graph;
  const int mid = (int)(mod - g->module.data());
  if(mid < (int)g->mem_source.size() && g->mem_source[mid].valid)
  {
    if(g->mem_source[mid].on_device) return 1;
    const vkb_raw_params_t *rp = &g->mem_source[mid].p;
    const size_t bytes = rp->packed_bpp ? ((size_t)rp->width * rp->height * rp->packed_bpp + 7) / 8 : (size_t)rp->width * rp->height * 2;
    memcpy(mapped, g->mem_source[mid].data, bytes);
    return 0;
  }
  if(imlv_open(mod)) return 1;
  mlv_clip_t *c = (mlv_clip_t *)mod->data;
  uint32_t frame = g->frame;
  if(frame >= c->frame_count) frame = c->frame_count - 1; // i-mlv/main.c:82
  if(c->lossless) return mlv_read_lossless(c, frame, (uint16_t *)mapped); // host decode like the reference, uploaded as plain u16
  return mlv_read_packed(c, frame, mapped);
}

// ------------------------------------------------------------------------------------------------
// denoise (denoise/main.c:80-333)
static void denoise_roi_in(dt_graph_t *graph, dt_module_t *module)
{
  const dt_image_params_t *img_param = dt_module_get_input_img_param(graph, module, dt_token("input"));
  if(!img_param) return;
  if(img_param->filters)
  {
    module->connector[0].roi.wd = module->connector[0].roi.full_wd;
    module->connector[0].roi.ht = module->connector[0].roi.full_ht;
    module->connector[0].roi.marker = s_roi_mark_hard_bck;
  }
  else module->connector[0].roi = module->connector[1].roi;
}
static void denoise_roi_out(dt_graph_t *graph, dt_module_t *module)
{
  const dt_image_params_t *img_param = dt_module_get_input_img_param(graph, module, dt_token("input"));
  if(!img_param) return;
  const uint32_t *b = img_param->crop_aabb;
  module->connector[1].roi = module->connector[0].roi;
  if(img_param->filters)
  {
    module->connector[1].roi.full_wd = b[2] - b[0];
    module->connector[1].roi.full_ht = b[3] - b[1];
  }
}
static inline int32_t fbits(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
// denoise/main.c:12-43 get_gain_maps_bayer: the list has to hold four GainMap opcodes up front, one per site of the 2x2 cfa block
// (plane 0, pitch 2 x 2, the same grid and, up to the site's offset, the same region); gm[filter] by the site the region starts on
struct denoise_data_t { const dt_dng_gain_map_t *gm[4] = { 0, 0, 0, 0 }; };
static int denoise_init(dt_module_t *mod) { mod->data = new denoise_data_t(); return 0; }
static void denoise_cleanup(dt_module_t *mod) { delete (denoise_data_t *)mod->data; mod->data = 0; }
static int get_gain_maps_bayer(const dt_dng_opcode_list_t *ol, int index, denoise_data_t *dat, uint32_t ox, uint32_t oy)
{
  if((int)ol->ops.size() - index < 4) return 0;
  for(int i = 0; i < 4; i++) if(ol->ops[i].id != 9) return 0;   // (the reference looks at ops[0..3] whatever `index` is)
  for(int i = 0; i < 4; i++) dat->gm[i] = 0;
  for(int i = 0; i < 4; i++)
  {
    const dt_dng_gain_map_t *gm = &ol->gain_maps[ol->ops[i].gain_map];
    if(!(gm->plane == 0 && gm->planes == 1 && gm->map_planes == 1 && gm->row_pitch == 2 && gm->col_pitch == 2)) return 0;
    if(gm->map_points_h < 2 && gm->map_points_v < 2) return 0;
    const int filter = (((gm->top + oy) & 1) << 1) + ((gm->left + ox) & 1);
    dat->gm[filter] = gm;
  }
  for(int i = 0; i < 4; i++) if(!dat->gm[i]) return 0;
  for(int i = 1; i < 4; i++)
    if(dat->gm[0]->map_points_h != dat->gm[i]->map_points_h || dat->gm[0]->map_points_v != dat->gm[i]->map_points_v ||
       dat->gm[0]->map_spacing_h != dat->gm[i]->map_spacing_h || dat->gm[0]->map_spacing_v != dat->gm[i]->map_spacing_v ||
       dat->gm[0]->map_origin_h != dat->gm[i]->map_origin_h || dat->gm[0]->map_origin_v != dat->gm[i]->map_origin_v ||
       dat->gm[0]->top / 2 != dat->gm[i]->top / 2 || dat->gm[0]->left / 2 != dat->gm[i]->left / 2 ||
       dat->gm[0]->bottom / 2 != dat->gm[i]->bottom / 2 || dat->gm[0]->right / 2 != dat->gm[i]->right / 2) return 0;
  return 1;
}
static int denoise_read_source(dt_module_t *mod, void *mapped, dt_read_source_params_t *p)
{ // denoise/main.c:61-80: the four gain planes interleaved into the rgba f32 texture of the (denoise, gainmap) node
  const denoise_data_t *dat = (const denoise_data_t *)mod->data;
  if(p->node->kernel != dt_token("gainmap") || !dat || !dat->gm[0]) return 0;
  const int wd = dat->gm[0]->map_points_h, ht = dat->gm[0]->map_points_v;
  for(int j = 0; j < ht; j++) for(int i = 0; i < wd; i++) for(int c = 0; c < 4; c++)
    ((float *)mapped)[(j * wd + i) * 4 + c] = dat->gm[c]->map_gain[j * wd + i];
  return 0;
}
static void denoise_create_nodes(dt_graph_t *graph, dt_module_t *module)
{
  const dt_image_params_t *img_param = dt_module_get_input_img_param(graph, module, dt_token("input"));
  if(!img_param) return;
  for(int k = 0; k < 4; k++) { module->img_param.black[k] = 0.0f; module->img_param.white[k] = 1.0f; }
  module->img_param.crop_aabb[0] = 0; module->img_param.crop_aabb[1] = 0;
  module->img_param.crop_aabb[2] = module->connector[1].roi.full_wd;
  module->img_param.crop_aabb[3] = module->connector[1].roi.full_ht;
  const float nowb[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  const float *wb = (!img_param->filters) ? nowb : img_param->whitebalance;
  int32_t wbi[4], blacki[4], whitei[4];
  const uint32_t *caf = img_param->crop_aabb;
  const float cs = module->connector[0].roi.wd / (float)module->connector[0].roi.full_wd;
  const uint32_t crop_aabb[4] = { (uint32_t)(caf[0] * cs), (uint32_t)(caf[1] * cs), (uint32_t)(caf[2] * cs), (uint32_t)(caf[3] * cs) };
  for(int k = 0; k < 4; k++) { wbi[k] = fbits(wb[k]); blacki[k] = fbits(img_param->black[k] / 65535.0f); whitei[k] = fbits(img_param->white[k] / 65535.0f); }
  const int32_t noisei[2] = { fbits(img_param->noise_a), fbits(img_param->noise_b) };
  // denoise/main.c:172-200: a source node for the gain maps of a dng, and where they sit in the image
  int32_t gainmap = 0, gainmap_sx = 0, gainmap_sy = 0, gainmap_ox = 0, gainmap_oy = 0;
  denoise_data_t *dat = (denoise_data_t *)module->data;
  const dt_image_metadata_dngop_t *dngop = (const dt_image_metadata_dngop_t *)module->img_param.meta;
  if(dngop && dat) for(int op = 0; op < (int)dngop->list2.ops.size() && !gainmap; op++) gainmap = get_gain_maps_bayer(&dngop->list2, op, dat, dngop->ox, dngop->oy);
  int id_gmdata = -1;
  if(gainmap)
  {
    const int map_wd = dat->gm[0]->map_points_h, map_ht = dat->gm[0]->map_points_v;
    dt_roi_t gmdata_roi = {}; gmdata_roi.wd = map_wd; gmdata_roi.ht = map_ht;
    id_gmdata = dt_node_add(graph, module, "denoise", "gainmap", map_wd, map_ht, 1, 0, 0, 1, "source", "source", "rgba", "f32", &gmdata_roi);
    const float ox = dat->gm[0]->map_origin_h, oy = dat->gm[0]->map_origin_v;
    const float sx = 1.0 / (dat->gm[0]->map_spacing_h * (map_wd - 1)), sy = 1.0 / (dat->gm[0]->map_spacing_v * (map_ht - 1));
    gainmap_ox = fbits(ox); gainmap_oy = fbits(oy); gainmap_sx = fbits(sx); gainmap_sy = fbits(sy);
  }
  else if(dat) dat->gm[0] = dat->gm[1] = dat->gm[2] = dat->gm[3] = 0;
  const float strength = dt_module_param_float(module, param_id(module, "strength"))[0];
  if(strength <= 0.0f)
  {
    if(img_param->filters == 0) return dt_connector_bypass(graph, module, 0, 1);
    const int32_t pc[] = { (int32_t)crop_aabb[0], (int32_t)crop_aabb[1], (int32_t)crop_aabb[2], (int32_t)crop_aabb[3],
      blacki[0], blacki[1], blacki[2], blacki[3], whitei[0], whitei[1], whitei[2], whitei[3],
      gainmap_ox, gainmap_oy, gainmap_sx, gainmap_sy, (int32_t)img_param->filters, gainmap };
    // the reference declares the output rgba and stores (v,0,0,1); every consumer reads .r: we keep one channel
    const int id_noop = dt_node_add(graph, module, "denoise", "noop", module->connector[1].roi.wd, module->connector[1].roi.ht, 1, sizeof(pc), pc, 3,
        "input",   "read",  "rgba", "f16", dt_no_roi,
        "output",  "write", "rggb", "f16", &module->connector[1].roi,
        "gainmap", "read",  "rgba", "*",   dt_no_roi);
    dt_connector_copy(graph, module, 0, id_noop, 0);
    dt_connector_copy(graph, module, 0, id_noop, 2);
    if(gainmap) CONN(dt_node_connect(graph, id_gmdata, 0, id_noop, 2));
    dt_connector_copy(graph, module, 1, id_noop, 1);
    return;
  }
  const int block = (img_param->filters == 0) ? 1 : (img_param->filters == 9u ? 3 : 2);
  dt_roi_t roi_half = module->connector[1].roi;
  roi_half.full_wd /= block; roi_half.full_ht /= block; roi_half.wd /= block; roi_half.ht /= block;
  const int wd = roi_half.wd, ht = roi_half.ht;
  int id_down[4] = {0};
  for(int i = 0; i < 4; i++)
  {
    const int c0 = (i == 0 && block == 1);
    const int32_t pc[] = { wbi[0], wbi[1], wbi[2], wbi[3], blacki[0], blacki[1], blacki[2], blacki[3], whitei[0], whitei[1], whitei[2], whitei[3],
      c0 ? (int32_t)crop_aabb[0] : 0, c0 ? (int32_t)crop_aabb[1] : 0, c0 ? (int32_t)crop_aabb[2] : 0, c0 ? (int32_t)crop_aabb[3] : 0,
      noisei[0], noisei[1], i, block };
    const int cov = (img_param->filters) && (i == 0);
    id_down[i] = dt_node_add(graph, module, "denoise", cov ? "downcov" : "down", wd, ht, 1, sizeof(pc), pc, cov ? 3 : 2,
        "input",  "read",  "rgba", "f16", dt_no_roi,
        "output", "write", "rgba", "f16", &roi_half,
        "cov",    "write", "rgba", "f16", &roi_half);
  }
  for(int i = 1; i < 4; i++) CONN(dt_node_connect(graph, id_down[i-1], 1, id_down[i], 0));
  const int c1 = block == 1;
  const int32_t pcas[] = { wbi[0], wbi[1], wbi[2], wbi[3], blacki[0], blacki[1], blacki[2], blacki[3], whitei[0], whitei[1], whitei[2], whitei[3],
    c1 ? (int32_t)crop_aabb[0] : 0, c1 ? (int32_t)crop_aabb[1] : 0, c1 ? (int32_t)crop_aabb[2] : 0, c1 ? (int32_t)crop_aabb[3] : 0,
    noisei[0], noisei[1], (int32_t)img_param->filters };
  const int id_assemble = dt_node_add(graph, module, "denoise", "assemble", wd, ht, 1, sizeof(pcas), pcas, 6,
      "s0", "read", "rgba", "f16", dt_no_roi, "s1", "read", "rgba", "f16", dt_no_roi, "s2", "read", "rgba", "f16", dt_no_roi,
      "s3", "read", "rgba", "f16", dt_no_roi, "s4", "read", "rgba", "f16", dt_no_roi, "output", "write", "rgba", "f16", &roi_half);
  for(int i = 0; i < 4; i++) CONN(dt_node_connect(graph, id_down[i], 1, id_assemble, i + 1));
  if(img_param->filters)
  {
    const int32_t pch[] = { wbi[0], wbi[1], wbi[2], wbi[3], blacki[0], blacki[1], blacki[2], blacki[3], whitei[0], whitei[1], whitei[2], whitei[3],
      (int32_t)crop_aabb[0], (int32_t)crop_aabb[1], (int32_t)crop_aabb[2], (int32_t)crop_aabb[3], (int32_t)img_param->filters };
    const int id_half = dt_node_add(graph, module, "denoise", "half", roi_half.full_wd, roi_half.full_ht, 1, sizeof(pch), pch, 2,
        "input",  "read",  "rggb", "ui16", dt_no_roi,
        "output", "write", "rgba", "f16", &roi_half);
    const int32_t pc[] = { wbi[0], wbi[1], wbi[2], wbi[3], blacki[0], blacki[1], blacki[2], blacki[3], whitei[0], whitei[1], whitei[2], whitei[3],
      (int32_t)crop_aabb[0], (int32_t)crop_aabb[1], (int32_t)crop_aabb[2], (int32_t)crop_aabb[3],
      (int32_t)img_param->filters, noisei[0], noisei[1], gainmap, gainmap_ox, gainmap_oy, gainmap_sx, gainmap_sy };
    const int id_doub = dt_node_add(graph, module, "denoise", "doub", module->connector[1].roi.wd, module->connector[1].roi.ht, 1, sizeof(pc), pc, 5,
        "orig", "read", "rggb", "f16", dt_no_roi, "crs0", "read", "rgba", "f16", dt_no_roi, "crs1", "read", "rgba", "f16", dt_no_roi,
        "output", "write", "rggb", "f16", &module->connector[1].roi, "gainmap", "read", "rgba", "*", dt_no_roi);
    CONN(dt_node_connect(graph, id_assemble, 5, id_doub, 1));
    CONN(dt_node_connect(graph, id_half, 1, id_doub, 2));
    if(gainmap) CONN(dt_node_connect(graph, id_gmdata, 0, id_doub, 4));
    else        CONN(dt_node_connect(graph, id_half, 1, id_doub, 4)); // connect dummy gainmap
    dt_connector_copy(graph, module, 0, id_doub, 0);
    dt_connector_copy(graph, module, 1, id_doub, 3);
    CONN(dt_node_connect(graph, id_half, 1, id_down[0], 0));
    CONN(dt_node_connect(graph, id_half, 1, id_assemble, 0));
    dt_connector_copy(graph, module, 0, id_half, 0);
  }
  else
  {
    dt_connector_copy(graph, module, 0, id_down[0], 0);
    dt_connector_copy(graph, module, 0, id_assemble, 0);
    dt_connector_copy(graph, module, 1, id_assemble, 5);
  }
}

// ------------------------------------------------------------------------------------------------
// hilite (hilite/main.c:5-90)
static void hilite_create_nodes(dt_graph_t *graph, dt_module_t *module)
{
  const dt_image_params_t *img_param = dt_module_get_input_img_param(graph, module, dt_token("input"));
  if(!img_param) return;
  const uint32_t filters = img_param->filters;
  if(!filters) return dt_connector_bypass(graph, module, 0, 1);
  const float *wb = module->img_param.whitebalance;
  const int wd = module->connector[0].roi.wd, ht = module->connector[0].roi.ht;
  dt_roi_t roic = module->connector[0].roi;
  const int block = filters == 9 ? 3 : 2;
  roic.wd /= block; roic.ht /= block;
  const int32_t pc[] = { fbits(wb[0]), fbits(wb[1]), fbits(wb[2]), fbits(wb[3]), (int32_t)filters };
  const int id_half = dt_node_add(graph, module, "hilite", "half", wd / block, ht / block, 1, sizeof(pc), pc, 2,
      "input",  "read",  "rggb", "ui16", dt_no_roi,
      "output", "write", "rgba", "f16",  &roic);
  const int id_doub = dt_node_add(graph, module, "hilite", "doub", wd / block, ht / block, 1, sizeof(pc), pc, 3,
      "input",  "read",  "rggb", "ui16", dt_no_roi,
      "coarse", "read",  "rgba", "f16",  dt_no_roi,
      "output", "write", "rggb", "ui16", &module->connector[0].roi);
  dt_connector_copy(graph, module, 0, id_half, 0);
  dt_connector_copy(graph, module, 0, id_doub, 0);
  dt_connector_copy(graph, module, 1, id_doub, 2);
  dt_roi_t rf = roic, rc = roic;
  rc.wd = (rc.wd - 1) / 2 + 1; rc.ht = (rc.ht - 1) / 2 + 1;
  rc.full_wd = (rc.full_wd - 1) / 2 + 1; rc.full_ht = (rc.full_ht - 1) / 2 + 1;
  int node_in = id_half, conn_in = 1, node_up = id_doub, conn_up = 1;
  const int max_nl = 15;
  for(int l = 1; l < max_nl; l++)
  {
    const int id_reduce = dt_node_add(graph, module, "hilite", "reduce", rc.wd, rc.ht, 1, sizeof(pc), pc, 2,
        "input",  "read",  "rgba", "f16", dt_no_roi,
        "output", "write", "rgba", "f16", &rc);
    const int id_assemble = dt_node_add(graph, module, "hilite", "assemble", rf.wd, rf.ht, 1, sizeof(pc), pc, 3,
        "fine",   "read",  "rgba", "f16", dt_no_roi,
        "coarse", "read",  "rgba", "f16", dt_no_roi,
        "output", "write", "rgba", "f16", &rf);
    CONN(dt_node_connect(graph, node_in, conn_in, id_reduce, 0));
    CONN(dt_node_connect(graph, node_in, conn_in, id_assemble, 0));
    node_in = id_reduce; conn_in = 1;
    CONN(dt_node_connect(graph, id_assemble, 2, node_up, conn_up));
    node_up = id_assemble; conn_up = 1;
    rf = rc;
    rc.wd = (rc.wd - 1) / 2 + 1; rc.ht = (rc.ht - 1) / 2 + 1;
    rc.full_wd = (rc.full_wd - 1) / 2 + 1; rc.full_ht = (rc.full_ht - 1) / 2 + 1;
    if(rc.wd <= 1 || rc.ht <= 1 || l + 1 == max_nl)
    {
      CONN(dt_node_connect(graph, id_reduce, 1, id_assemble, 1));
      break;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// demosaic (demosaic/main.c:8-203): method 0 (gaussian splats); rcd / halfsize are later rows of SURVEY §8
static void demosaic_roi_in(dt_graph_t *, dt_module_t *module)
{
  dt_roi_t *ri = &module->connector[0].roi;
  ri->wd = ri->full_wd; ri->ht = ri->full_ht;
  ri->marker = s_roi_mark_hard_bck;
}
static void demosaic_roi_out(dt_graph_t *, dt_module_t *module)
{
  dt_roi_t *ri = &module->connector[0].roi, *ro = &module->connector[1].roi;
  const int method = dt_module_param_int(module, 1)[0];
  const int block = module->img_param.filters == 9u ? 3 : 2;
  const float scale = ro->full_wd > 0 ? (float)ri->full_wd / (float)ro->full_wd : 1.0f;
  const int halfsize = (method == 2) || (scale >= 1.5 * block);
  ro->marker = ri->marker;
  if(halfsize)
  {
    ro->full_wd = (ri->full_wd + 1) / 2; ro->full_ht = (ri->full_ht + 1) / 2;
    if(scale >= block) ro->marker = s_roi_mark_soft_fwd;
  }
  else { ro->full_wd = ri->full_wd; ro->full_ht = ri->full_ht; }
  module->img_param.filters = 0u;
}
static void demosaic_create_nodes(dt_graph_t *graph, dt_module_t *module)
{
  const dt_image_params_t *img_param = dt_module_get_input_img_param(graph, module, dt_token("input"));
  if(!img_param) return;
  if(!img_param->filters)
  {
    if(module->connector[0].roi.wd == module->connector[1].roi.wd && module->connector[0].roi.ht == module->connector[1].roi.ht)
      return dt_connector_bypass(graph, module, 0, 1);
    fprintf(stderr, "[vkdt_b200] demosaic: resize of non-mosaic input is outside the hot path\n");
    return;
  }
  const int block = img_param->filters == 9u ? 3 : 2;
  module->img_param.filters = 0u;
  const int wd = module->connector[0].roi.wd, ht = module->connector[0].roi.ht;
  dt_roi_t roi_full = module->connector[0].roi, roi_half = module->connector[0].roi;
  roi_half.full_wd /= block; roi_half.full_ht /= block; roi_half.wd /= block; roi_half.ht /= block;
  const float *wb = img_param->whitebalance;
  const int32_t pc[] = { fbits(wb[0]), fbits(wb[1]), fbits(wb[2]), fbits(wb[3]), (int32_t)img_param->filters };
  const int method = dt_module_param_int(module, 1)[0];
  const float scale = (float)module->connector[0].roi.wd / (float)module->connector[1].roi.wd;
  const int halfsize = (scale >= 1.5 * block) || (method == 2);
  if(halfsize)
  { // demosaic/main.c:93-112
    const int id_half = dt_node_add(graph, module, "demosaic", "halfsize", roi_half.wd, roi_half.ht, 1, sizeof(pc), pc, 2,
        "input",  "read",  "rggb", "*",   dt_no_roi,
        "output", "write", "rgba", "f16", &roi_half);
    dt_connector_copy(graph, module, 0, id_half, 0);
    if(block != scale)
    { // resample to get to the rest of the resolution, only if block != scale
      const int id_resample = dt_node_add(graph, module, "shared", "resample", module->connector[1].roi.wd, module->connector[1].roi.ht, 1, 0, 0, 2,
          "input",  "read",  "rgba", "f16", dt_no_roi,
          "output", "write", "rgba", "f16", &module->connector[1].roi);
      CONN(dt_node_connect(graph, id_half, 1, id_resample, 0));
      dt_connector_copy(graph, module, 1, id_resample, 1);
    }
    else dt_connector_copy(graph, module, 1, id_half, 1);
    return;
  }
  if(method == 1 && block == 2)
  { // bayer with RCD (demosaic/main.c:116-156).  the reference sizes rcd_fill's dispatch in 64x32 shared memory tiles;
    // our kernel tiles internally, the node keeps the image extent
    dt_roi_t hr = module->connector[0].roi;
    hr.wd /= 2;
    const int id_conv = dt_node_add(graph, module, "demosaic", "rcd_conv", wd, ht, 1, 0, 0, 4,
        "cfa", "read",  "*", "*",   dt_no_roi,
        "vh",  "write", "r", "f16", &module->connector[0].roi,
        "pq",  "write", "r", "f16", &hr,
        "lp",  "write", "r", "f16", &hr);
    const int id_fill = dt_node_add(graph, module, "demosaic", "rcd_fill", wd, ht, 1, sizeof(pc), pc, 5,
        "cfa", "read", "*", "*", dt_no_roi,
        "vh",  "read", "*", "*", dt_no_roi,
        "pq",  "read", "*", "*", dt_no_roi,
        "lp",  "read", "*", "*", dt_no_roi,
        "output", "write", "rgba", "f16", &module->connector[0].roi);
    CONN(dt_node_connect_named(graph, id_conv, "vh", id_fill, "vh"));
    CONN(dt_node_connect_named(graph, id_conv, "pq", id_fill, "pq"));
    CONN(dt_node_connect_named(graph, id_conv, "lp", id_fill, "lp"));
    dt_connector_copy(graph, module, 0, id_conv, 0);
    dt_connector_copy(graph, module, 0, id_fill, 0);
    if(module->connector[1].roi.marker & s_roi_mark_hard)
    {
      const int id_resample = dt_node_add(graph, module, "shared", "resample", module->connector[1].roi.wd, module->connector[1].roi.ht, 1, 0, 0, 2,
          "input",  "read",  "rgba", "f16", dt_no_roi,
          "output", "write", "rgba", "f16", &module->connector[1].roi);
      CONN(dt_node_connect(graph, id_fill, 4, id_resample, 0));
      dt_connector_copy(graph, module, 1, id_resample, 1);
    }
    else dt_connector_copy(graph, module, 1, id_fill, 4);
    return;
  }
  const int id_down = dt_node_add(graph, module, "demosaic", "down", wd / block, ht / block, 1, sizeof(pc), pc, 2,
      "input", "read", "rggb", "*", dt_no_roi,
      "output", "write", "y", "f16", &roi_half);
  const int id_gauss = dt_node_add(graph, module, "demosaic", "gauss", wd / block, ht / block, 1, sizeof(pc), pc, 3,
      "input",  "read",  "y",    "f16", dt_no_roi,
      "orig",   "read",  "rggb", "*",   dt_no_roi,
      "output", "write", "rgba", "f16", &roi_half);
  CONN(dt_node_connect(graph, id_down, 1, id_gauss, 0));
  dt_connector_copy(graph, module, 0, id_gauss, 1);
  const int id_splat = dt_node_add(graph, module, "demosaic", "splat", wd, ht, 1, sizeof(pc), pc, 3,
      "input",  "read",  "rggb", "*",   dt_no_roi,
      "gauss",  "read",  "rgba", "f16", dt_no_roi,
      "output", "write", "g",    "f16", &roi_full);
  dt_connector_copy(graph, module, 0, id_splat, 0);
  CONN(dt_node_connect(graph, id_gauss, 2, id_splat, 1));
  dt_connector_copy(graph, module, 0, id_down, 0);
  const int id_fix = dt_node_add(graph, module, "demosaic", "fix", wd, ht, 1, sizeof(pc), pc, 4,
      "input",  "read",  "rggb", "*",   dt_no_roi,
      "green",  "read",  "g",    "*",   dt_no_roi,
      "cov",    "read",  "rgba", "f16", dt_no_roi,
      "output", "write", "rgba", "f16", &roi_full);
  dt_connector_copy(graph, module, 0, id_fix, 0);
  CONN(dt_node_connect(graph, id_splat, 2, id_fix, 1));
  CONN(dt_node_connect(graph, id_gauss, 2, id_fix, 2));
  if(module->connector[1].roi.marker & s_roi_mark_hard)
  {
    const int id_resample = dt_node_add(graph, module, "shared", "resample", module->connector[1].roi.wd, module->connector[1].roi.ht, 1, 0, 0, 2,
        "input",  "read",  "rgba", "f16", dt_no_roi,
        "output", "write", "rgba", "f16", &module->connector[1].roi);
    CONN(dt_node_connect(graph, id_fix, 3, id_resample, 0));
    dt_connector_copy(graph, module, 1, id_resample, 1);
  }
  else dt_connector_copy(graph, module, 1, id_fix, 3);
}

// ------------------------------------------------------------------------------------------------
// crop (crop/main.c:175-345)
static void get_crop_rot(uint32_t orient, double wd, double ht, const float *p_crop, const float *p_rot, float *crop, float *rot)
{
  const float rotation = p_rot[0];
  rot[0] = rotation;
  crop[0] = p_crop[0]; crop[1] = p_crop[1]; crop[2] = p_crop[2]; crop[3] = p_crop[3];
  if(rotation == 1337.0f)
  {
    if(orient == 3)      rot[0] = 180.0f;
    else if(orient == 8) rot[0] = 90.0f;
    else if(orient == 6) rot[0] = 270.0f;
    else                 rot[0] = 0.0f;
  }
  if(crop[0] == 1.0 && crop[1] == 3.0 && crop[2] == 3.0 && crop[3] == 7.0)
  {
    const double crw = wd > 400 ? 3.0 / wd : 0.0, crh = ht > 400 ? 3.0 / ht : 0.0;
    const bool quarter = (rot[0] >= 45 && rot[0] < 135) || (!(rot[0] < 225) && rot[0] < 315);
    if(quarter)
    {
      crop[0] = 0.5 - (.5 - crh) * ht / wd; crop[2] = 0.5 - (.5 - crw) * wd / ht;
      crop[1] = 0.5 + (.5 - crh) * ht / wd; crop[3] = 0.5 + (.5 - crw) * wd / ht;
    }
    else { crop[0] = crw; crop[2] = crh; crop[1] = 1.0 - crw; crop[3] = 1.0 - crh; }
  }
}
static int crop_init(dt_module_t *mod) { mod->committed_param_size = sizeof(float) * 20; return 0; }
static void crop_roi_in(dt_graph_t *, dt_module_t *module)
{
  float crop[4], rot;
  const float *p_crop = dt_module_param_float(module, 1), *p_rot = dt_module_param_float(module, 2);
  const float w = module->connector[0].roi.full_wd, h = module->connector[0].roi.full_ht;
  get_crop_rot(module->img_param.orientation, w, h, p_crop, p_rot, crop, &rot);
  const float wd = crop[1] - crop[0], ht = crop[3] - crop[2];
  if(module->connector[1].roi.full_wd == module->connector[1].roi.wd)
  {
    module->connector[0].roi.wd = module->connector[0].roi.full_wd;
    module->connector[0].roi.ht = module->connector[0].roi.full_ht;
  }
  else
  {
    module->connector[0].roi.wd = module->connector[1].roi.wd / wd;
    module->connector[0].roi.ht = module->connector[1].roi.ht / ht;
  }
  module->connector[0].roi.marker = module->connector[1].roi.marker;
}
static void crop_roi_out(dt_graph_t *, dt_module_t *module)
{
  float crop[4], rot;
  const float *p_crop = dt_module_param_float(module, 1), *p_rot = dt_module_param_float(module, 2);
  const float w = module->connector[0].roi.full_wd, h = module->connector[0].roi.full_ht;
  get_crop_rot(module->img_param.orientation, w, h, p_crop, p_rot, crop, &rot);
  module->connector[1].roi = module->connector[0].roi;
  const float wd = crop[1] - crop[0], ht = crop[3] - crop[2];
  const float fw = module->connector[0].roi.full_wd * wd, fh = module->connector[0].roi.full_ht * ht;
  module->connector[1].roi.full_wd = (uint32_t)(32768 < fw ? 32768 : fw);
  module->connector[1].roi.full_ht = (uint32_t)(32768 < fh ? 32768 : fh);
}
static void crop_commit(dt_graph_t *, dt_module_t *module)
{
  const float *inp = dt_module_param_float(module, 0);
  float p[8];
  for(int k = 0; k < 4; k++)
  {
    p[2*k+0] = module->connector[0].roi.wd * inp[2*k+0];
    p[2*k+1] = module->connector[0].roi.ht * inp[2*k+1];
  }
  const float a = p[0], A = p[2], b = p[1], B = p[7];
  const float u[] = {a, b, A, b, A, B, a, B};
  double M[] = {
    u[0], u[1], 1, 0, 0, 0, -p[0]*u[0], -p[0]*u[1],
    u[2], u[3], 1, 0, 0, 0, -p[2]*u[2], -p[2]*u[3],
    u[4], u[5], 1, 0, 0, 0, -p[4]*u[4], -p[4]*u[5],
    u[6], u[7], 1, 0, 0, 0, -p[6]*u[6], -p[6]*u[7],
    0, 0, 0, u[0], u[1], 1, -p[1]*u[0], -p[1]*u[1],
    0, 0, 0, u[2], u[3], 1, -p[3]*u[2], -p[3]*u[3],
    0, 0, 0, u[4], u[5], 1, -p[5]*u[4], -p[5]*u[5],
    0, 0, 0, u[6], u[7], 1, -p[7]*u[6], -p[7]*u[7],
  };
  double r[] = {p[0], p[2], p[4], p[6], p[1], p[3], p[5], p[7], 1.0};
  gauss_solve(M, r, 8);
  float *f = (float *)module->committed_param;
  f[ 0] = r[0]; f[ 1] = r[3]; f[ 2] = r[6]; f[ 3] = 0.0f;
  f[ 4] = r[1]; f[ 5] = r[4]; f[ 6] = r[7]; f[ 7] = 0.0f;
  f[ 8] = r[2]; f[ 9] = r[5]; f[10] = r[8]; f[11] = 0.0f;
  f += 12;
  float crop[4], rot;
  const float *p_crop = dt_module_param_float(module, 1), *p_rot = dt_module_param_float(module, 2);
  const float wd = module->connector[0].roi.wd, ht = module->connector[0].roi.ht;
  get_crop_rot(module->img_param.orientation, wd, ht, p_crop, p_rot, crop, &rot);
  const float rad = rot * 3.1415629 / 180.0f; // sic
  f[0] = cosf(rad); f[1] = sinf(rad); f[2] = -sinf(rad); f[3] = cosf(rad);
  f += 4;
  f[0] = crop[0]; f[1] = crop[1]; f[2] = crop[2]; f[3] = crop[3];
  if(p_rot[0] == 1337.0f)
  { // write back the resolved values (crop/main.c:339-344)
    dt_module_set_param_float_n(module, dt_token("crop"), crop, 4);
    dt_module_set_param_float(module, dt_token("rotate"), rot);
  }
}

// ------------------------------------------------------------------------------------------------
// colour (colour/main.c:88-465), lut inputs (clut/picked/abney/spectra) unconnected
static int colour_init(dt_module_t *mod) { mod->committed_param_size = sizeof(float) * (4+12+4+12+4*24+4*24+5+8+5); return 0; }
static void colour_roi_in(dt_graph_t *, dt_module_t *module)
{
  module->connector[0].roi = module->connector[1].roi;
  for(int k = 2; k <= 5; k++) module->connector[k].roi.marker = s_roi_mark_uninited;
}
static void colour_roi_out(dt_graph_t *graph, dt_module_t *module)
{
  module->connector[1].roi = module->connector[0].roi;
  const dt_image_params_t *img_param = dt_module_get_input_img_param(graph, module, dt_token("input"));
  if(!img_param) return;
  for(int k = 0; k < 9; k++) module->img_param.cam_to_rec2020[k] = (k % 4 == 0) ? 1.0f : 0.0f;
  module->img_param.whitebalance[0] = module->img_param.whitebalance[1] = module->img_param.whitebalance[2] = 1.0f;
  module->img_param.colour_primaries = 2; // s_colour_primaries_2020
  module->img_param.colour_trc = 0;       // s_colour_trc_linear
}
static void rbf_coefficients(const int N, const float *source, const float *target, float *coef)
{ // colour/main.c:88-186
  const int N2 = N + 3;
  if(N == 0) { for(int co = 0; co < 3; co++) coef[co*4+co] = 1.0f; return; }
  if(N == 1) { for(int co = 0; co < 3; co++) coef[co*4+co] = target[co] / source[co]; return; }
  std::vector<double> A(N2 * N2), b(N2);
  std::vector<int> pivot(N2);
  for(int j = 0; j < N; j++) for(int i = j; i < N; i++)
  {
    const float *x = source + 3*i, *y = source + 3*j;
    const double r2 = (x[0]-y[0])*(x[0]-y[0]) + (x[1]-y[1])*(x[1]-y[1]) + (x[2]-y[2])*(x[2]-y[2]);
    A[j*N2+i] = A[i*N2+j] = sqrt(r2);
  }
  for(int k = 0; k < 3; k++) for(int i = 0; i < N; i++) A[i*N2+N+k] = A[(N+k)*N2+i] = source[3*i+k];
  for(int j = N; j < N2; j++) for(int i = N; i < N2; i++) A[j*N2+i] = 0;
  if(!gauss_make_triangular(A.data(), pivot.data(), N2)) return;
  for(int ch = 0; ch < 3; ch++)
  {
    for(int i = 0; i < N; i++) b[i] = target[3*i+ch];
    for(int i = N; i < N2; i++) b[i] = 0;
    gauss_solve_triangular(A.data(), pivot.data(), b.data(), N2);
    for(int i = 0; i < N; i++) coef[12 + 4*i + ch] = b[i];
    for(int i = 0; i < 3; i++) coef[4*i + ch] = b[N+i];
  }
}
static void colour_commit(dt_graph_t *graph, dt_module_t *module)
{
  const dt_image_params_t *img_param = dt_module_get_input_img_param(graph, module, dt_token("input"));
  if(!img_param) return;
  float *f = (float *)module->committed_param;
  uint32_t *i = (uint32_t *)module->committed_param;
  float *p_wb = (float *)dt_module_param_float(module, param_id(module, "white"));
  const float  p_tmp = dt_module_param_float(module, param_id(module, "temp"))[0];
  const int    p_cnt = dt_module_param_int(module, param_id(module, "cnt"))[0];
  const float *p_map = dt_module_param_float(module, param_id(module, "rbmap"));
  const int    p_mat = dt_module_param_int(module, param_id(module, "matrix"))[0];
  const float *p_mtx = dt_module_param_float(module, param_id(module, "mat"));
  const int    p_gam = dt_module_param_int(module, param_id(module, "gamut"))[0];
  const int    p_mod = dt_module_param_int(module, param_id(module, "mode"))[0];
  const float  p_sat = dt_module_param_float(module, param_id(module, "sat"))[0];
  const int    p_pck = dt_module_param_int(module, param_id(module, "picked"))[0];
  const int    p_clp = dt_module_param_int(module, param_id(module, "clip"))[0];
  const float  p_clm = dt_module_param_float(module, param_id(module, "clipmax"))[0];
  if(p_wb[0] == 0.0f && p_wb[1] == 0.0f && p_wb[2] == 0.0f)
  {
    float w0[3] = {0}, w[] = { img_param->whitebalance[0], img_param->whitebalance[1], img_param->whitebalance[2] };
    for(int j = 0; j < 3; j++) for(int k = 0; k < 3; k++) w0[j] += img_param->cam_to_rec2020[3*j+k] / w[k];
    w0[0] /= w0[1]; w0[2] /= w0[1]; w0[1] = 1.0f;
    p_wb[0] = 1 / w0[0]; p_wb[1] = 1; p_wb[2] = 1 / w0[2];
  }
  if(!(p_wb[0] == p_wb[0]) || p_wb[0] == 0.0f || p_wb[1] == 0.0f || p_wb[2] == 0.0f) p_wb[0] = p_wb[1] = p_wb[2] = 1.0f;
  f[0] = p_wb[0] / p_wb[1]; f[1] = 1.0f; f[2] = p_wb[2] / p_wb[1];
  f[3] = powf(2.0f, ((float *)module->param)[0]);
  const int off = 4+12+4+12+4*24+4*24;
  // three bands identify the legacy clut layout (colour/main.c:268-292)
  const uint32_t clut_ht = module->connector[2].roi.full_ht;
  const uint32_t nbands  = clut_ht ? module->connector[2].roi.full_wd / clut_ht : 3;
  if(p_tmp <= 0.0f) f[off+0] = -1.0f; // as-shot: resolved by the autotemp node in the reference; the kernel here refuses it
  else if(nbands > 3)
  { // anchors uniform in mired, 2000 .. 15000 K
    const float T_lo = 2000.0f, T_hi = 15000.0f;
    const float m_lo = 1e6f / T_hi, m_hi = 1e6f / T_lo;
    const float m = 1e6f / (p_tmp < T_lo ? T_lo : p_tmp > T_hi ? T_hi : p_tmp);
    const float v = (m - m_lo) / (m_hi - m_lo);
    f[off+0] = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
  }
  else
  {
    float v = tanf(asinhf(46.3407f + p_tmp)) + (-0.0287128f * cosf(0.000798585f * (714.855f - p_tmp))) + 0.942275f;
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    f[off+0] = 1.0f - v;
  }
  i[off+1] = p_mat == 4 ? 1 : 0;
  f[off+2] = p_sat;
  i[off+3] = p_pck;
  i[off+4] = p_gam;
  i[off+5] = img_param->colour_primaries;
  i[off+6] = img_param->colour_trc;
  f[off+7] = p_clp ? p_clm : 0.0;
  float awb[3] = { img_param->whitebalance[0], img_param->whitebalance[1], img_param->whitebalance[2] };
  if(!(awb[0] > 0.0f) || !(awb[1] > 0.0f) || !(awb[2] > 0.0f)) awb[0] = awb[1] = awb[2] = 1.0f;
  f[off+8] = awb[0] / awb[1]; f[off+9] = 1.0f; f[off+10] = awb[2] / awb[1]; f[off+11] = 1.0f;
  if(p_mat == 1) { for(int j = 0; j < 3; j++) for(int k = 0; k < 3; k++) f[4+4*k+j] = img_param->cam_to_rec2020[3*j+k]; }
  else if(p_mat == 2) { i[off+5] = 5; i[off+6] = 0; }
  else if(p_mat == 3) { i[off+5] = 1; i[off+6] = 0; }
  else if(p_mat == 5)
  {
    i[off+5] = 0; i[off+6] = 0;
    for(int j = 0; j < 3; j++) for(int k = 0; k < 3; k++) f[4+4*k+j] = p_mtx[3*j+k];
  }
  else
  {
    i[off+5] = 2; i[off+6] = 0;
    for(int j = 0; j < 3; j++) for(int k = 0; k < 3; k++) f[4+4*j+k] = k == j ? 1.0f : 0.0f;
  }
  if(p_mod == 1)
  {
    const int N = p_cnt < 0 ? 0 : (p_cnt > 24 ? 24 : p_cnt);
    i[16] = N; i[17] = i[18] = i[19] = 0;
    float src[72], tgt[72];
    for(int k = 0; k < N; k++) for(int c = 0; c < 3; c++) { src[3*k+c] = p_map[6*k+c]; tgt[3*k+c] = p_map[6*k+3+c]; }
    memset(f + 20, 0, sizeof(float) * (12 + 24*4 + 24*4));
    for(int k = 0; k < N; k++) { f[128+4*k+0] = src[3*k+0]; f[128+4*k+1] = src[3*k+1]; f[128+4*k+2] = src[3*k+2]; f[128+4*k+3] = 0.0f; }
    rbf_coefficients(N, src, tgt, f + 20);
  }
  else i[16] = i[17] = i[18] = i[19] = 0;
}
static void colour_create_nodes(dt_graph_t *graph, dt_module_t *module)
{ // colour/main.c:416-465.  the colour picker input is not part of the path (a connected picker is ignored).  of the autotemp /
  // sink pair behind a clut the autotemp node is there: with temp <= 0 every run blends the clut's anchors where the as-shot
  // white balance comes out neutral, like the reference's first run; the sink that writes this back into the parameter as
  // an editable temperature (write_sink, :395-413) is a gui matter and is left out
  const int have_clut = dt_connected(module->connector + 2);
  const int have_abney = dt_connected(module->connector + 4) && dt_connected(module->connector + 5);
  if(dt_connected(module->connector + 3)) fprintf(stderr, "[vkdt_b200] colour: the colour picker input is ignored (outside the hot path)\n");
  const int pc[] = { have_clut, 0, have_abney };
  int id_auto = -1;
  if(have_clut)
  {
    const int pc_auto[] = { 0 };
    dt_roi_t tiny = {};
    tiny.wd = tiny.ht = 1;    // colour/main.c:427: only the extent is set
    id_auto = dt_node_add(graph, module, "colour", "autotemp", 1, 1, 1, sizeof(pc_auto), pc_auto, 3,
        "clut",   "read",  "rg", "f16", dt_no_roi,
        "temp",   "write", "y",  "f32", &tiny,
        "picked", "read",  "r",  "f16", dt_no_roi);
    dt_connector_copy(graph, module, 2, id_auto, 0);
    dt_connector_copy(graph, module, 0, id_auto, 2); // dummy
  }
  const int nodeid = dt_node_add(graph, module, "colour", "main", module->connector[0].roi.wd, module->connector[0].roi.ht, 1, sizeof(pc), pc, 7,
      "input",   "read",  "rgba", "f16", dt_no_roi,
      "output",  "write", "rgba", "f16", &module->connector[0].roi,
      "clut",    "read",  "rgba", "f16", dt_no_roi,
      "picked",  "read",  "r",    "f16", dt_no_roi,
      "abney",   "read",  "rg",   "f16", dt_no_roi,
      "spectra", "read",  "rgba", "f16", dt_no_roi,
      "autotemp", "read", "y",    "f32", dt_no_roi);
  dt_connector_copy(graph, module, 0, nodeid, 0);
  dt_connector_copy(graph, module, 1, nodeid, 1);
  dt_connector_copy(graph, module, have_clut ? 2 : 0, nodeid, 2);
  dt_connector_copy(graph, module, 0, nodeid, 3);
  dt_connector_copy(graph, module, have_abney ? 4 : 0, nodeid, 4);
  dt_connector_copy(graph, module, have_abney ? 5 : 0, nodeid, 5);
  if(id_auto >= 0) dt_node_connect(graph, id_auto, 1, nodeid, 6);
  else             dt_connector_copy(graph, module, 0, nodeid, 6);
}

// ------------------------------------------------------------------------------------------------
// filmcurv (filmcurv/main.c:3-38); the histogram/dspy nodes feed a gui widget only and are never reachable in cli
static void filmcurv_roi_out(dt_graph_t *, dt_module_t *module)
{
  module->connector[2].roi.full_wd = 1024;
  module->connector[2].roi.full_ht = 512;
  module->connector[1].roi = module->connector[0].roi;
}
static void filmcurv_create_nodes(dt_graph_t *graph, dt_module_t *module)
{ // filmcurv/main.c:12-38.  hist and dspy draw the curve widget of the gui: they hang off the `dspy` connector, which nothing
  // on an export path reads, so the executor never reaches them (no kernel is built for them); they exist so that the node
  // list is the reference's
  const int wd = module->connector[0].roi.wd, ht = module->connector[0].roi.ht;
  dt_roi_t hroi = {};
  hroi.wd = 16 + 1; hroi.ht = 16;
  const int id_hist = dt_node_add(graph, module, "OpenDRT", "hist", (wd + 15) / 16 * 8, (ht + 15) / 16 * 8, 1, 0, 0, 2,
      "input",  "read",  "*",    "*",   dt_no_roi,
      "hist",   "write", "ssbo", "u32", &hroi);
  const int id_main = dt_node_add(graph, module, "filmcurv", "main", module->connector[0].roi.wd, module->connector[0].roi.ht, 1, 0, 0, 2,
      "input",  "read",  "*",    "*",   dt_no_roi,
      "output", "write", "rgba", "f16", &module->connector[1].roi);
  const int id_dspy = dt_node_add(graph, module, "filmcurv", "dspy", module->connector[2].roi.wd, module->connector[2].roi.ht, 1, 0, 0, 2,
      "hist",   "read",  "ssbo", "u32", dt_no_roi,
      "output", "write", "rgba", "f16", &module->connector[2].roi);
  graph->node[id_hist].connector[1].flags |= s_conn_clear;
  CONN(dt_node_connect_named(graph, id_hist, "hist", id_dspy, "hist"));
  dt_connector_copy(graph, module, 0, id_hist, 0);
  dt_connector_copy(graph, module, 0, id_main, 0);
  dt_connector_copy(graph, module, 1, id_main, 1);
  dt_connector_copy(graph, module, 2, id_dspy, 1);
}

// ------------------------------------------------------------------------------------------------
// llap (llap/main.c:6-106)
static void llap_create_nodes(dt_graph_t *graph, dt_module_t *module)
{
  const int wd = module->connector[0].roi.wd, ht = module->connector[0].roi.ht, dp = 1;
  const int num_gamma = 10;
  int pc[] = { num_gamma };
  const int id_curve = dt_node_add(graph, module, "llap", "curve", wd, ht, dp, sizeof(pc), pc, 2,
      "input",  "read",  "rgba", "f16", dt_no_roi,
      "output", "write", "y",    "f16", &module->connector[0].roi);
  graph->node[id_curve].connector[1].array_length = num_gamma + 1;
  dt_roi_t rf = module->connector[0].roi, rc = module->connector[0].roi;
  rc.wd = (rc.wd - 1) / 2 + 1; rc.ht = (rc.ht - 1) / 2 + 1;
  rc.full_wd = (rc.full_wd - 1) / 2 + 1; rc.full_ht = (rc.full_ht - 1) / 2 + 1;
  const int max_nl = 12;
  int nl = max_nl;
  int id_reduce[12] = { -1 }, id_assemble[12] = { -1 };
  id_reduce[0] = id_curve;
  for(int l = 1; l < nl; l++)
  {
    id_reduce[l] = dt_node_add(graph, module, "llap", "reduce", rc.wd, rc.ht, num_gamma + 1, 0, 0, 2,
        "inhi",  "read",  "y", "f16", dt_no_roi,
        "outlo", "write", "y", "f16", &rc);
    graph->node[id_reduce[l]].connector[1].array_length = num_gamma + 1;
    CONN(dt_node_connect(graph, id_reduce[l-1], 1, id_reduce[l], 0));
    int pca[] = { num_gamma, 0 };
    id_assemble[l] = dt_node_add(graph, module, "llap", "assemble", rf.wd, rf.ht, dp, sizeof(pca), pca, 4,
        "coarse", "read",  "y", "f16", dt_no_roi,
        "currlo", "read",  "y", "f16", dt_no_roi,
        "currhi", "read",  "y", "f16", dt_no_roi,
        "fine",   "write", "y", "f16", &rf);
    CONN(dt_node_connect(graph, id_reduce[l-1], 1, id_assemble[l], 1));
    CONN(dt_node_connect(graph, id_reduce[l  ], 1, id_assemble[l], 2));
    if(l > 1) CONN(dt_node_connect(graph, id_assemble[l], 3, id_assemble[l-1], 0));
    rf = rc;
    rc.wd = (rc.wd - 1) / 2 + 1; rc.ht = (rc.ht - 1) / 2 + 1;
    rc.full_wd = (rc.full_wd - 1) / 2 + 1; rc.full_ht = (rc.full_ht - 1) / 2 + 1;
    if(rc.wd <= 1 || rc.ht <= 1) { nl = l + 1; break; }
  }
  ((int32_t *)graph->node[id_assemble[nl-1]].push_constant)[1] = 1;
  CONN(dt_node_connect(graph, id_curve, 1, id_assemble[nl-1], 0));
  const int id_col = dt_node_add(graph, module, "llap", "colour", wd, ht, dp, 0, 0, 3,
      "lum",    "read",  "y",    "f16", dt_no_roi,
      "input",  "read",  "rgba", "f16", dt_no_roi,
      "output", "write", "rgba", "f16", &module->connector[0].roi);
  CONN(dt_node_connect(graph, id_assemble[1], 3, id_col, 0));
  dt_connector_copy(graph, module, 0, id_curve, 0);
  dt_connector_copy(graph, module, 0, id_col, 1);
  dt_connector_copy(graph, module, 1, id_col, 2);
}

// ------------------------------------------------------------------------------------------------
// o-pfm (o-pfm/main.c:8-42)
// colenc/main.c: the image downstream is in the encoded colour space
static void colenc_roi_out(dt_graph_t *, dt_module_t *module)
{
  module->img_param.colour_primaries = dt_module_param_int(module, dt_module_get_param(module->so, dt_token("prim")))[0];
  module->img_param.colour_trc       = dt_module_param_int(module, dt_module_get_param(module->so, dt_token("trc")))[0];
  module->connector[1].roi.full_wd = module->connector[0].roi.full_wd;
  module->connector[1].roi.full_ht = module->connector[0].roi.full_ht;
}

// ------------------------------------------------------------------------------------------------
// resize (resize/main.c): what `vkdt-cli --width / --height` puts in front of the sink (graph-export.c:54-62).  the sink's
// max_wd / max_ht shrink its roi request (graph-run-modules.h:458-476), this module asks its input for the full image
// and rescales: catmull-rom when minifying (behind a separable gaussian above a factor of three), flower taps when magnifying
static void resize_roi_out(dt_graph_t *, dt_module_t *module)
{ // resize/main.c:6-27
  module->connector[1].roi = module->connector[0].roi;
  const int wd = dt_module_param_int(module, 0)[0], ht = dt_module_param_int(module, 1)[0];
  if(!wd || !ht) { module->connector[1].roi.marker = s_roi_mark_soft_fwd; return; }
  const double scale = std::min(1.0, std::min(wd / (double)module->connector[1].roi.full_wd, ht / (double)module->connector[1].roi.full_ht));
  if(scale < 1.0)
  {
    module->connector[1].roi.full_wd = (int)(module->connector[0].roi.full_wd * scale + 0.5);
    module->connector[1].roi.full_ht = (int)(module->connector[0].roi.full_ht * scale + 0.5);
  }
}
static void resize_roi_in(dt_graph_t *, dt_module_t *module)
{ // resize/main.c:29-39: request the full thing, we'll rescale
  module->connector[0].roi.wd = module->connector[0].roi.full_wd;
  module->connector[0].roi.ht = module->connector[0].roi.full_ht;
  module->connector[0].roi.marker = (module->connector[1].roi.marker & ~s_roi_mark_hard) | s_roi_mark_soft;
}
// modules/api.h:506-578 dt_api_blur_sep: two nodes (shared, blurh) -> (shared, blurv) with the input's channels, format and roi;
// returns blurv, *id_in = blurh.  nodeid_input < 0: the module's connector is the input (wired by the caller with dt_connector_copy)
static int api_blur_sep(dt_graph_t *g, dt_module_t *module, int nodeid_input, int connid_input, int *id_in, float radius)
{
  const dt_connector_t ci = nodeid_input >= 0 ? g->node[nodeid_input].connector[connid_input] : module->connector[connid_input];
  const int dp = ci.array_length > 0 ? ci.array_length : 1;
  const std::string chan = dt_token_string(ci.chan), format = dt_token_string(ci.format);
  int id[2];
  for(int k = 0; k < 2; k++)
  {
    id[k] = dt_node_add(g, module, "shared", k ? "blurv" : "blurh", ci.roi.wd, ci.roi.ht, dp, sizeof(float), &radius, 2,
        "input", "read", chan.c_str(), format.c_str(), dt_no_roi,
        "output", "write", chan.c_str(), format.c_str(), &ci.roi);
    for(int c = 0; c < 2; c++)
    { // the reference fills both connectors from the input's (roi, array length), connected = unset
      g->node[id[k]].connector[c].roi = ci.roi;
      g->node[id[k]].connector[c].array_length = ci.array_length;
    }
    g->node[id[k]].connector[0].connected = s_cid_unset;
  }
  if(nodeid_input >= 0) CONN(dt_node_connect(g, nodeid_input, connid_input, id[0], 0));
  CONN(dt_node_connect(g, id[0], 1, id[1], 0));
  if(id_in) *id_in = id[0];
  return id[1];
}
static void resize_create_nodes(dt_graph_t *g, dt_module_t *module)
{ // resize/main.c:57-83
  if(module->connector[0].roi.wd == module->connector[1].roi.wd && module->connector[0].roi.ht == module->connector[1].roi.ht)
    return dt_connector_bypass(g, module, 0, 1);
  const float scale = module->connector[0].roi.wd / (float)module->connector[1].roi.wd;   // downscaling factor
  int mode = scale < 0.99f ? 0 : scale > 1.01f ? 2 : 1;                                    // magnify / 1:1 / minify
  if(scale > 3) mode = 1;                                                                  // slice after blur
  const int32_t pc[] = { mode };
  const int id_resize = dt_node_add(g, module, "resize", "main", module->connector[1].roi.wd, module->connector[1].roi.ht, 1, sizeof(pc), pc, 2,
      "input", "read", "*", "*", dt_no_roi,
      "output", "write", "rgba", "f16", &module->connector[1].roi);
  if(scale > 3)
  { // 3x3 is the natural support of the catmull-rom spline: blur above that only.  radius in px = 3 sigma
    const float blur = scale + 0.5f;
    int id_blur_in = -1;
    const int id_blur = api_blur_sep(g, module, -1, 0, &id_blur_in, blur);
    CONN(dt_node_connect_named(g, id_blur, "output", id_resize, "input"));
    dt_connector_copy(g, module, 0, id_blur_in, 0);
  }
  else dt_connector_copy(g, module, 0, id_resize, 0);
  dt_connector_copy(g, module, 1, id_resize, 1);
}

// o-jpg/main.c:102-172 with an own baseline encoder (pipe/jpeg.cpp); no icc profile, no exif copy (needs exiftool)
static void ojpg_write_sink(dt_module_t *module, void *buf, dt_write_sink_params_t *)
{
  const char *basename = dt_module_param_string(module, 0);
  fprintf(stderr, "[o-jpg] writing '%s'\n", basename);
  char filename[512];
  snprintf(filename, sizeof(filename), "%s.jpg", basename);
  const float quality = dt_module_param_float(module, 1)[0];
  if(jpeg_write_rgba8(filename, (const uint8_t *)buf, module->connector[0].roi.wd, module->connector[0].roi.ht, quality))
    fprintf(stderr, "[o-jpg] could not write '%s'\n", filename);
}

static void opfm_write_sink(dt_module_t *module, void *buf, dt_write_sink_params_t *)
{
  const char *basename = dt_module_param_string(module, 0);
  fprintf(stderr, "[o-pfm] writing '%s'\n", basename);
  const float *pf = (const float *)buf;
  const int width = module->connector[0].roi.wd, height = module->connector[0].roi.ht;
  char filename[512];
  snprintf(filename, sizeof(filename), "%s.pfm", basename);
  FILE *f = fopen(filename, "wb");
  if(!f) return;
  char header[1024];
  snprintf(header, sizeof(header), "PF\n%d %d\n-1.0", width, height);
  const size_t len = strlen(header);
  fputs(header, f);
  long off = 0;
  while((len + 1 + off) & 0xf) off++;
  while(off-- > 0) fputc('0', f);
  fputc('\n', f);
  // the executor hands a file sink the payload as the file wants it (r g b per pixel, VKB_SINK_RGB_F32): the last
  // kernel stored it that way, so this is one write instead of the reference's fwrite per pixel (o-pfm/main.c:36-40)
  fwrite(pf, sizeof(float), (size_t)3 * width * height, f);
  fclose(f);
}

/* ORACLE — test infrastructure.  drives the reference's own colour/main.c commit_params (:219-365, incl. the RBF
 * coefficient solve compute_coefficients :88-186), compiled in place from /root/reference by `make -C oracle ref`, never
 * copied, so that o_colour_commit and the product's colour module can be pinned against it. */
#define ui_callback    colour_ref_ui_callback
#define modify_roi_out colour_ref_modify_roi_out
#define modify_roi_in  colour_ref_modify_roi_in
#define commit_params  colour_ref_commit_params
#define init           colour_ref_init
#define animate        colour_ref_animate
#define check_params   colour_ref_check_params
#define write_sink     colour_ref_write_sink
#define create_nodes   colour_ref_create_nodes
#include "pipe/modules/colour/main.c"
#include <stdlib.h>
#include <string.h>

/* a symbol write_sink of the module would bind to in a real vkdt: never reached from commit_params.
 * (dt_node_connect comes from the reference's connector.c, linked into the same library) */
qvk_t qvk;

/* p: the module's parameter block in the order of colour/params (exposure sat picked matrix gamut clip clipmax temp
 * white[4] mat[9] mode cnt rbmap[144] import[8]) = o_colour_params_t.  clut / picked / abney / spectra unconnected. */
int ref_colour_commit(const void *p, uint32_t psize, const float *img_wb4, const float *img_cam_to_rec2020, int primaries, int trc,
    float *p_wb_out4, float *committed /* 242 floats */)
{
  static const char *names[14] = { "exposure", "sat", "picked", "matrix", "gamut", "clip", "clipmax", "temp", "white", "mat", "mode", "cnt", "rbmap", "import" };
  static const char *types[14] = { "float", "float", "int", "int", "int", "int", "float", "float", "float", "float", "int", "int", "float", "string" };
  static const int   cnts[14]  = { 1, 1, 1, 1, 1, 1, 1, 1, 4, 9, 1, 1, 144, 8 };
  static dt_ui_param_t par[14];
  static dt_module_so_t so;
  dt_graph_t *graph = calloc(1, sizeof(*graph));
  graph->module = calloc(2, sizeof(dt_module_t));
  graph->num_modules = graph->max_modules = 2;
  dt_module_t *src = graph->module, *mod = graph->module + 1;
  memset(&so, 0, sizeof(so));
  int off = 0;
  for(int k = 0; k < 14; k++)
  {
    memset(par + k, 0, sizeof(par[k]));
    par[k].name = dt_token(names[k]); par[k].type = dt_token(types[k]); par[k].cnt = cnts[k]; par[k].offset = off;
    off += cnts[k] * (k == 13 ? 1 : 4);
    so.param[k] = par + k;
  }
  so.num_params = 14;
  if((uint32_t)off > psize) { free(graph->module); free(graph); return 1; }
  uint8_t *params = malloc(off);
  memcpy(params, p, off);
  for(int k = 0; k < 4; k++) src->img_param.whitebalance[k] = img_wb4[k];
  for(int k = 0; k < 9; k++) src->img_param.cam_to_rec2020[k] = img_cam_to_rec2020[k];
  src->img_param.colour_primaries = primaries; src->img_param.colour_trc = trc;
  mod->so = &so; mod->graph = graph; mod->param = params; mod->param_size = off;
  static const char *cn[6] = { "input", "output", "clut", "picked", "abney", "spectra" };
  mod->num_connectors = 6;
  for(int c = 0; c < 6; c++) { mod->connector[c].name = dt_token(cn[c]); mod->connector[c].connected.i = -1; mod->connector[c].connected.c = -1; mod->connector[c].type = dt_token(c == 1 ? "write" : "read"); }
  mod->connector[0].connected.i = 0; mod->connector[0].connected.c = 0;
  colour_ref_init(mod);
  float *cp = calloc(1, mod->committed_param_size + 64);
  mod->committed_param = (uint8_t *)cp;
  colour_ref_commit_params(graph, mod);
  memcpy(committed, cp, mod->committed_param_size);
  memcpy(p_wb_out4, params + par[8].offset, 16); /* commit_params writes the derived white back into the parameter */
  const int n = mod->committed_param_size / 4;
  free(cp); free(params); free(graph->module); free(graph);
  return n;
}

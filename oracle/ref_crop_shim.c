/* ORACLE — test infrastructure.  drives the reference's own crop/main.c (compiled in place from /root/reference by
 * `make -C oracle ref`, never copied) so that the oracle's o_crop_roi_out / o_crop_commit and the product's crop module
 * can be pinned against it: modify_roi_out (crop/main.c:256-275) and commit_params (:277-345). */
#define modify_roi_out crop_ref_modify_roi_out
#define modify_roi_in  crop_ref_modify_roi_in
#define commit_params  crop_ref_commit_params
#define init           crop_ref_init
#define ui_callback    crop_ref_ui_callback
#include "pipe/modules/crop/main.c"
#include <stdlib.h>
#include <string.h>

/* crop/params: perspect:float:8, crop:float:4, rotate:float:1 */
int ref_crop(uint32_t orientation, uint32_t in_w, uint32_t in_h, const float *perspect8, const float *crop4, float rotate,
    uint32_t *out_w, uint32_t *out_h, float *committed20)
{
  static dt_ui_param_t par[3];
  static dt_module_so_t so;
  dt_module_t *mod = calloc(1, sizeof(*mod));
  dt_graph_t *graph = calloc(1, sizeof(*graph));
  float values[13];
  memcpy(values, perspect8, 32); memcpy(values + 8, crop4, 16); values[12] = rotate;
  const int cnt[3] = { 8, 4, 1 }, off[3] = { 0, 32, 48 };
  memset(&so, 0, sizeof(so));
  for(int k = 0; k < 3; k++) { memset(par + k, 0, sizeof(par[k])); par[k].type = dt_token("float"); par[k].cnt = cnt[k]; par[k].offset = off[k]; so.param[k] = par + k; }
  par[0].name = dt_token("perspect"); par[1].name = dt_token("crop"); par[2].name = dt_token("rotate");
  so.num_params = 3;
  mod->so = &so; mod->graph = graph; mod->param = (uint8_t *)values; mod->param_size = sizeof(values);
  mod->num_connectors = 2;
  mod->connector[0].name = dt_token("input"); mod->connector[1].name = dt_token("output");
  mod->connector[0].roi.full_wd = mod->connector[0].roi.wd = in_w;
  mod->connector[0].roi.full_ht = mod->connector[0].roi.ht = in_h;
  mod->img_param.orientation = orientation;
  float committed[64] = {0};
  mod->committed_param = (uint8_t *)committed; mod->committed_param_size = 20 * sizeof(float);
  crop_ref_modify_roi_out(graph, mod);
  *out_w = mod->connector[1].roi.full_wd; *out_h = mod->connector[1].roi.full_ht;
  /* the graph would now run modify_roi_in with the sink's request: full size */
  mod->connector[1].roi.wd = mod->connector[1].roi.full_wd; mod->connector[1].roi.ht = mod->connector[1].roi.full_ht;
  crop_ref_commit_params(graph, mod);
  memcpy(committed20, committed, 20 * sizeof(float));
  free(mod); free(graph);
  return 0;
}

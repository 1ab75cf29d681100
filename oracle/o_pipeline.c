/* ORACLE — test infrastructure only (see o_common.h).
 * Module wiring as done by the reference's create_nodes callbacks, the default darkroom graph
 * (bin/default-darkroom.i-raw:1-32 after vkdt-cli replaced display by o-pfm, graph-export.c:23-96),
 * the MLV bit unpack and the PFM writer. */
#include "o_common.h"
#include "vkdt_oracle.h"
#include <stdio.h>

/* i-mlv/video_mlv.c:261-273.  the reference reads an unaligned u32 at word `addr` and rotates right by
 * 16 + (32-bpp) - shift (>= 32 for small shifts: formally UB, mod-32 on x86).  equivalent well defined form: */
void o_mlv_unpack(const uint16_t *w, uint64_t pixel_cnt, int bpp, uint16_t *out)
{
  const uint32_t mask = (1u << bpp) - 1u;
#pragma omp parallel for schedule(static)
  for(int64_t i = 0; i < (int64_t)pixel_cnt; i++)
  {
    const uint64_t bits = (uint64_t)i * bpp;
    const uint64_t addr = bits / 16;
    const uint32_t shift = (uint32_t)(bits % 16);
    const uint32_t v = ((uint32_t)w[addr] << 16) | w[addr + 1];
    out[i] = (uint16_t)((v >> (32 - bpp - shift)) & mask);
  }
}

/* hilite/main.c:5-90 */
void o_hilite_module(const oimg_t *in, oimg_t *out, const o_hilite_params_t *p, const float *wb4, uint32_t filters)
{
  enum { MAXL = 15 };
  const int block = filters == 9 ? 3 : 2;
  oimg_t lvl[MAXL];
  lvl[0] = o_img_alloc(in->w / block, in->h / block, 4);
  o_hilite_half(in, &lvl[0], p, filters);
  int L = 0, cw = lvl[0].w, ch = lvl[0].h;
  cw = (cw - 1) / 2 + 1; ch = (ch - 1) / 2 + 1;
  for(int l = 1; l < MAXL; l++)
  {
    lvl[l] = o_img_alloc(cw, ch, 4);
    o_hilite_reduce(&lvl[l-1], &lvl[l], p, wb4);
    L = l;
    cw = (cw - 1) / 2 + 1; ch = (ch - 1) / 2 + 1;
    if(cw <= 1 || ch <= 1 || l + 1 == MAXL) break;
  }
  oimg_t coarse = lvl[L]; int own = 0;
  for(int l = L; l >= 1; l--)
  {
    oimg_t a = o_img_alloc(lvl[l-1].w, lvl[l-1].h, 4);
    o_hilite_assemble(&lvl[l-1], &coarse, &a, p);
    if(own) o_img_free(&coarse);
    coarse = a; own = 1;
  }
  o_hilite_doub(in, &coarse, out, p, filters);
  if(own) o_img_free(&coarse);
  for(int l = 0; l <= L; l++) o_img_free(&lvl[l]);
}

/* white balance pushed to rcd_fill (img_param->whitebalance of demosaic's input, demosaic/main.c:89-90); set by the caller */
float o_demosaic_wb[4] = { 1.0f, 1.0f, 1.0f, 1.0f };

/* demosaic/main.c:159-202, method 0 (1:1 shared/resample node is the identity and elided) */
void o_demosaic_module(const oimg_t *in, oimg_t *out, const o_demosaic_params_t *p, uint32_t filters)
{
  const int block = filters == 9 ? 3 : 2;
  if(p->method == 2)
  { /* demosaic/main.c:93-112: half size; out is ((w+1)/2, (h+1)/2), resampled when the block is not 2 */
    if(out->w == in->w / block && out->h == in->h / block) { o_demosaic_halfsize(in, out, filters); return; }
    oimg_t half = o_img_alloc(in->w / block, in->h / block, 4);
    o_demosaic_halfsize(in, &half, filters);
    o_resample(&half, out);
    o_img_free(&half);
    return;
  }
  if(p->method == 1 && filters != 9)
  { /* demosaic/main.c:116-156: RCD.  wb travels in the unused tail of the params struct (see o_darkroom_run) */
    oimg_t vh = o_img_alloc(in->w, in->h, 1), pq = o_img_alloc(in->w / 2, in->h, 1), lp = o_img_alloc(in->w / 2, in->h, 1);
    o_rcd_conv(in, &vh, &pq, &lp);
    o_rcd_fill(in, &vh, &pq, &lp, out, o_demosaic_wb);
    o_img_free(&vh); o_img_free(&pq); o_img_free(&lp);
    return;
  }
  oimg_t cov = o_img_alloc(in->w / block, in->h / block, 4);
  oimg_t green = o_img_alloc(in->w, in->h, 1);
  o_demosaic_gauss(in, &cov, filters);
  o_demosaic_splat(in, &cov, &green, filters);
  o_demosaic_fix(in, &green, &cov, out, filters, p->colour);
  o_img_free(&cov); o_img_free(&green);
}

/* the dng gain maps of the source (denoise/main.c:172-200), set by the caller before o_denoise_module / o_darkroom_run:
 * rgba f32 texture + { origin x, origin y, 1 / extent x, 1 / extent y }; 0 = none */
/* the lut inputs of colour (i-lut modules wired to its clut / abney / spectra connectors), same convention as the gain map */
static const oimg_t *o_lut_clut = 0, *o_lut_abney = 0, *o_lut_spectra = 0;
void o_set_colour_luts(const oimg_t *clut, const oimg_t *abney, const oimg_t *spectra)
{
  o_lut_clut = clut; o_lut_abney = abney; o_lut_spectra = spectra;
}
static const oimg_t *o_gainmap_img = 0;
static float o_gainmap_os[4];
void o_set_gainmap(const oimg_t *gm, const float *map_os)
{
  o_gainmap_img = gm;
  if(gm) for(int k = 0; k < 4; k++) o_gainmap_os[k] = map_os[k];
}

/* denoise/main.c:134-333 for mosaic input */
void o_denoise_module(const oimg_t *in, oimg_t *out, const o_denoise_params_t *p, const int *crop, const float *wb4,
    const float *black4, const float *white4, float noise_a, float noise_b, uint32_t filters)
{
  float black[4], white[4];
  for(int k = 0; k < 4; k++) { black[k] = black4[k] / 65535.0f; white[k] = white4[k] / 65535.0f; }
  const oimg_t *gm = (o_gainmap_img && filters != 9 && p->gainmap == 1) ? o_gainmap_img : 0;
  if(p->strength <= 0.0f)
  {
    if(gm) o_denoise_noop_gm(in, out, crop, black, white, gm, o_gainmap_os);
    else o_denoise_noop(in, out, crop, black, white);
    return;
  }
  const int block = filters == 9 ? 3 : 2;
  const int hw = out->w / block, hh = out->h / block;
  oimg_t half = o_img_alloc(hw, hh, 4), cov = o_img_alloc(hw, hh, 4), assembled = o_img_alloc(hw, hh, 4);
  oimg_t dn[4];
  for(int i = 0; i < 4; i++) dn[i] = o_img_alloc(hw, hh, 4);
  o_denoise_half(in, &half, crop, white, filters);
  o_denoise_downcov(&half, &dn[0], &cov);
  for(int i = 1; i < 4; i++) o_denoise_down(&dn[i-1], &dn[i], p, black, white, noise_a, noise_b, i, block);
  o_denoise_assemble(&half, &dn[0], &dn[1], &dn[2], &dn[3], &assembled, p, wb4, black, white, noise_a, noise_b, filters);
  o_denoise_doub_gm(in, &assembled, &half, out, p, crop, black, white, noise_a, noise_b, filters, gm, o_gainmap_os);
  for(int i = 0; i < 4; i++) o_img_free(&dn[i]);
  o_img_free(&half); o_img_free(&cov); o_img_free(&assembled);
}

void o_darkroom_defaults(o_darkroom_t *d, uint32_t width, uint32_t height)
{
  memset(d, 0, sizeof(*d));
  d->width = width; d->height = height;
  d->filters = 0x5d5d5d5d; /* i-mlv/main.c:127 */
  d->crop_aabb[2] = width; d->crop_aabb[3] = height;
  for(int k = 0; k < 4; k++) { d->black[k] = 2048; d->white[k] = 15000; d->whitebalance[k] = 1.0f; }
  d->cam_to_rec2020[0] = d->cam_to_rec2020[4] = d->cam_to_rec2020[8] = 1.0f;
  d->noise_a = 1.0f; d->noise_b = 1.0f; /* i-mlv/main.c:140-141 */
  d->denoise = (o_denoise_params_t){ 0.0f, 0.6f, 1.0f, 0.0f, {0, 0, 0, 0}, 1 };
  d->hilite = (o_hilite_params_t){ 0.985f, 0.3f, 0.6f };
  d->demosaic = (o_demosaic_params_t){ 0, 0 };
  const float persp[8] = {0.25f, 0.25f, 0.75f, 0.25f, 0.75f, 0.75f, 0.25f, 0.75f};
  memcpy(d->crop.perspect, persp, sizeof(persp));
  d->crop.crop[0] = 1.0f; d->crop.crop[1] = 3.0f; d->crop.crop[2] = 3.0f; d->crop.crop[3] = 7.0f;
  d->crop.rotate = 1337.0f;
  d->colour.exposure = 0.0f; d->colour.sat = 1.0f; d->colour.matrix = 1; d->colour.clipmax = 1.0f; d->colour.temp = 6504.0f;
  d->colour.mat[0] = -1.0f; /* missing defaults read as 0 (asciiio.h:28-36) */
  d->colour.cnt = 4;
  {
    const float rb[24] = {0.3333f, 0.3333f, 0.3333f, 0.3333f, 0.3333f, 0.3333f, 0.5f, 0.25f, 0.25f, 0.5f, 0.25f, 0.25f,
      0.25f, 0.5f, 0.25f, 0.25f, 0.5f, 0.25f, 0.25f, 0.25f, 0.5f, 0.25f, 0.25f, 0.5f};
    for(int k = 0; k < 144; k++) d->colour.rbmap[k] = k < 24 ? rb[k] : 0.0f;
  }
  d->filmcurv = (o_filmcurv_params_t){ 3.0f, 1.2f, 0.0f, 3, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
  d->llap = (o_llap_params_t){ 0.12f, 1.0f, 1.0f, 0.2f }; /* default-darkroom.i-raw:27-30 */
  d->grade = (o_grade_params_t){ {0, 0, 0, 0}, {1, 1, 1, 0}, {1, 1, 1, 0}, {0, 0, 0, 0}, 0, 0.3f, 0.4f };
  d->enable_llap = 1; d->enable_grade = 1;
}

void o_darkroom_out_size(const o_darkroom_t *d, uint32_t *out_w, uint32_t *out_h)
{
  uint32_t w = d->crop_aabb[2] - d->crop_aabb[0], h = d->crop_aabb[3] - d->crop_aabb[1];
  if(d->demosaic.method == 2) { w = (w + 1) / 2; h = (h + 1) / 2; } /* demosaic/main.c:29-33 */
  o_crop_roi_out(d->orientation, w, h, d->crop.crop, &d->crop.rotate, out_w, out_h);
}

static void copy_out(const oimg_t *im, float *dst) { memcpy(dst, im->p, sizeof(float) * (size_t)im->w * im->h * im->c); }

int o_darkroom_run(const o_darkroom_t *d, const uint16_t *raw, float *out, int stage, float *stage_out)
{
  const int W = d->width, H = d->height;
  oimg_t src = o_img_alloc(W, H, 1);
#pragma omp parallel for schedule(static)
  for(int64_t i = 0; i < (int64_t)W * H; i++) src.p[i] = (float)raw[i] / 65535.0f; /* R16_UNORM */
  const int cw = d->crop_aabb[2] - d->crop_aabb[0], ch = d->crop_aabb[3] - d->crop_aabb[1];
  const int crop[4] = { (int)d->crop_aabb[0], (int)d->crop_aabb[1], (int)d->crop_aabb[2], (int)d->crop_aabb[3] };
  int ret = 0;

  oimg_t den = o_img_alloc(cw, ch, 1);
  o_denoise_module(&src, &den, &d->denoise, crop, d->whitebalance, d->black, d->white, d->noise_a, d->noise_b, d->filters);
  o_img_free(&src);
  if(stage == 1) { copy_out(&den, stage_out); o_img_free(&den); return 0; }

  oimg_t hil = o_img_alloc(cw, ch, 1);
  o_hilite_module(&den, &hil, &d->hilite, d->whitebalance, d->filters);
  o_img_free(&den);
  if(stage == 2) { copy_out(&hil, stage_out); o_img_free(&hil); return 0; }

  const int dw = d->demosaic.method == 2 ? (cw + 1) / 2 : cw, dh = d->demosaic.method == 2 ? (ch + 1) / 2 : ch;
  oimg_t dem = o_img_alloc(dw, dh, 4);
  for(int k = 0; k < 4; k++) o_demosaic_wb[k] = d->whitebalance[k];
  o_demosaic_module(&hil, &dem, &d->demosaic, d->filters);
  o_img_free(&hil);
  if(stage == 3) { copy_out(&dem, stage_out); o_img_free(&dem); return 0; }

  uint32_t ow, oh;
  o_darkroom_out_size(d, &ow, &oh);
  float fc[20];
  o_crop_commit(d->orientation, dw, dh, d->crop.perspect, d->crop.crop, &d->crop.rotate, fc);
  oimg_t crp = o_img_alloc(ow, oh, 4);
  o_crop_main(&dem, &crp, fc);
  o_img_free(&dem);
  if(stage == 4) { copy_out(&crp, stage_out); o_img_free(&crp); return 0; }

  float fcol[O_COLOUR_COMMITTED_FLOATS];
  float p_wb[4] = { d->colour.white[0], d->colour.white[1], d->colour.white[2], d->colour.white[3] };
  o_colour_commit(&d->colour, p_wb, d->whitebalance, d->cam_to_rec2020, d->colour_primaries, d->colour_trc, fcol);
  oimg_t col = o_img_alloc(ow, oh, 4);
  if(o_lut_clut && d->colour.temp > 0.0f)
  { /* colour/main.c:268-292: the anchor blend of a clut with more than three bands is uniform in mired, 2000 .. 15000 K */
    const int nbands = o_lut_clut->w / o_lut_clut->h;
    if(nbands > 3)
    {
      const float T_lo = 2000.0f, T_hi = 15000.0f, m_lo = 1e6f / T_hi, m_hi = 1e6f / T_lo;
      const float m = 1e6f / o_clamp(d->colour.temp, T_lo, T_hi);
      fcol[224] = o_clamp((m - m_lo) / (m_hi - m_lo), 0.0f, 1.0f);
    }
  }
  /* temp <= 0: the autotemp node's answer for this frame (colour/main.c:425-441; its write-back into the parameter is a gui matter) */
  const float auto_temp = (o_lut_clut && fcol[224] < 0.0f) ? o_colour_autotemp(o_lut_clut, fcol) : 0.0f;
  o_colour_main_lut(&crp, &col, fcol, 1, o_lut_clut, o_lut_abney, o_lut_spectra, auto_temp);
  o_img_free(&crp);
  if(stage == 5) { copy_out(&col, stage_out); o_img_free(&col); return 0; }

  /* the last module before the sink stores f32 (graph-export.c:89-91) */
  const int film_last = !d->enable_llap && !d->enable_grade && !d->enable_colenc;
  oimg_t flm = o_img_alloc(ow, oh, 4);
  o_filmcurv_main(&col, &flm, &d->filmcurv, !film_last);
  o_img_free(&col);
  if(stage == 6) { copy_out(&flm, stage_out); o_img_free(&flm); return 0; }

  oimg_t cur = flm;
  if(d->enable_llap)
  {
    oimg_t ll = o_img_alloc(ow, oh, 4);
    o_llap_module(&cur, &ll, &d->llap, (d->enable_grade || d->enable_colenc) ? 1 : 0);
    o_img_free(&cur);
    cur = ll;
    if(stage == 7) { copy_out(&cur, stage_out); o_img_free(&cur); return 0; }
  }
  if(d->enable_grade)
  {
    oimg_t gr = o_img_alloc(ow, oh, 4);
    o_grade_main(&cur, &gr, &d->grade, d->enable_colenc ? 1 : 0);
    o_img_free(&cur);
    cur = gr;
  }
  if(d->enable_colenc)
  { /* graph-export.c:66-86: colenc's output takes the sink's format */
    oimg_t ce = o_img_alloc(ow, oh, 4);
    o_colenc_main(&cur, &ce, d->colenc_prim, d->colenc_trc, d->sink_unorm8 ? 2 : 0);
    o_img_free(&cur);
    cur = ce;
  }
  if(out) copy_out(&cur, out);
  o_img_free(&cur);
  return ret;
}

/* o-pfm/main.c:8-42: "PF\n%d %d\n-1.0" padded with '0' so that the payload starts 16-byte aligned, rgb only, no flip */
int o_write_pfm(const char *filename, const float *rgba, int width, int height)
{
  FILE *f = fopen(filename, "wb");
  if(!f) return 1;
  char header[1024];
  snprintf(header, sizeof(header), "PF\n%d %d\n-1.0", width, height);
  size_t len = strlen(header);
  fputs(header, f);
  long off = 0;
  while((len + 1 + off) & 0xf) off++;
  while(off-- > 0) fputc('0', f);
  fputc('\n', f);
  for(size_t k = 0; k < (size_t)width * height; k++) fwrite(rgba + 4 * k, sizeof(float), 3, f);
  fclose(f);
  return 0;
}

/* ORACLE — TEST INFRASTRUCTURE ONLY (see o_common.h for the parity statement).
 * Public surface of liboracle.so, loaded by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py through ctypes.  Never linked into libvkdt_b200.so. */
#pragma once
#include <stdint.h>
#include "o_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* raw `params` blobs in the order of the modules' params files (src/pipe/module.c:52-65) */
typedef struct { float strength, luma, detail, pad; float edges[4]; int gainmap; } o_denoise_params_t; /* denoise/params */
typedef struct { float white, desat, soft; } o_hilite_params_t;                                        /* hilite/params */
typedef struct { int colour, method; } o_demosaic_params_t;                                            /* demosaic/params */
typedef struct { float perspect[8]; float crop[4]; float rotate; } o_crop_params_t;                    /* crop/params */
typedef struct
{ /* colour/params */
  float exposure, sat; int picked, matrix, gamut, clip; float clipmax, temp; float white[4]; float mat[9];
  int mode, cnt; float rbmap[144]; char import[8];
} o_colour_params_t;
#define O_COLOUR_COMMITTED_FLOATS (4+12+4+12+4*24+4*24+5+8+5) /* colour/main.c:369 */
typedef struct { float light, contrast, bias; int colour; float chroma, rolloff, red, yellow, blue, shadows; } o_filmcurv_params_t; /* filmcurv/params */
typedef struct { float sigma, shadows, hilights, clarity; } o_llap_params_t;                            /* llap/params */
typedef struct { float lift[4], gamma[4], gain[4], offset[4]; int mode; float sh_pivot, hi_pivot; } o_grade_params_t; /* grade/params */

/* i-mlv: video_mlv.c:261-273 restated with the well-defined shift */
void o_mlv_unpack(const uint16_t *packed_words, uint64_t pixel_cnt, int bpp, uint16_t *out);

/* denoise */
void o_denoise_noop(const oimg_t *in, oimg_t *out, const int *crop, const float *black, const float *white);
void o_denoise_half(const oimg_t *in, oimg_t *out, const int *crop, const float *white4, uint32_t filters);
void o_denoise_downcov(const oimg_t *in, oimg_t *out, oimg_t *covimg);
void o_denoise_down(const oimg_t *in, oimg_t *out, const o_denoise_params_t *p, const float *black4, const float *white4,
    float noise_a, float noise_b, int level, uint32_t block);
void o_denoise_assemble(const oimg_t *s0, const oimg_t *s1, const oimg_t *s2, const oimg_t *s3, const oimg_t *s4,
    oimg_t *out, const o_denoise_params_t *p, const float *wb, const float *black, const float *white,
    float noise_a, float noise_b, uint32_t filters);
void o_denoise_doub(const oimg_t *in, const oimg_t *crs0, const oimg_t *crs1, oimg_t *out, const o_denoise_params_t *p,
    const int *crop, const float *black4, const float *white4, float noise_a, float noise_b, uint32_t filters);
/* the same with a DNG gain map (rgba f32 texture, map_os = origin x, y, 1 / extent x, y): noop.comp:48-57, doub.comp:106-114 */
void o_denoise_noop_gm(const oimg_t *in, oimg_t *out, const int *crop, const float *black, const float *white, const oimg_t *gm, const float *map_os);
void o_denoise_doub_gm(const oimg_t *in, const oimg_t *crs0, const oimg_t *crs1, oimg_t *out, const o_denoise_params_t *p,
    const int *crop, const float *black4, const float *white4, float noise_a, float noise_b, uint32_t filters, const oimg_t *gm, const float *map_os);
int  o_xtrans_colour(int x, int y);

/* hilite */
void o_hilite_half(const oimg_t *in, oimg_t *out, const o_hilite_params_t *p, uint32_t filters);
void o_hilite_reduce(const oimg_t *in, oimg_t *out, const o_hilite_params_t *p, const float *wb4);
void o_hilite_assemble(const oimg_t *fine, const oimg_t *coarse, oimg_t *out, const o_hilite_params_t *p);
void o_hilite_doub(const oimg_t *in, const oimg_t *coarse, oimg_t *out, const o_hilite_params_t *p, uint32_t filters);

/* demosaic */
void o_demosaic_down(const oimg_t *in, oimg_t *out, uint32_t filters);
void o_demosaic_halfsize(const oimg_t *in, oimg_t *out, uint32_t filters);
void o_resample(const oimg_t *in, oimg_t *out);
/* resize/main.comp (mode 0 magnify / 1 slice / 2 minify), shared/blurh.comp + blurv.comp */
void o_resize_main(const oimg_t *in, oimg_t *out, int mode, int out_f16);
void o_blur_sep(const oimg_t *in, oimg_t *out, float radius, int vertical, int out_f16);
void o_rcd_conv(const oimg_t *cfa, oimg_t *vh, oimg_t *pq, oimg_t *lp);
void o_rcd_fill(const oimg_t *cfa, const oimg_t *vh, const oimg_t *pq, const oimg_t *lp, oimg_t *out, const float *wb);
void o_demosaic_gauss(const oimg_t *orig, oimg_t *out, uint32_t filters);
void o_demosaic_splat(const oimg_t *in, const oimg_t *gauss, oimg_t *out, uint32_t filters);
void o_demosaic_fix(const oimg_t *in, const oimg_t *green, const oimg_t *cov, oimg_t *out, uint32_t filters, int fixup);

/* crop */
int  o_gauss_solve(double *A, double *b, int n);
void o_crop_get_crop_rot(uint32_t orientation, double wd, double ht, const float *p_crop, const float *p_rot, float *crop, float *rot);
void o_crop_roi_out(uint32_t orientation, uint32_t in_w, uint32_t in_h, const float *p_crop, const float *p_rot, uint32_t *out_w, uint32_t *out_h);
void o_crop_commit(uint32_t orientation, uint32_t in_w, uint32_t in_h, const float *p_perspect, const float *p_crop, const float *p_rot, float *f20);
void o_crop_main(const oimg_t *in, oimg_t *out, const float *f20);
void o_sample_catmull_rom(const oimg_t *tex, float u, float v, float *res);

/* colour */
void o_colour_commit(const o_colour_params_t *p, float *p_wb, const float *img_wb, const float *img_cam_to_rec2020,
    int img_primaries, int img_trc, float *f);
void o_colour_main(const oimg_t *in, oimg_t *out, const float *f, int out_f16);
float o_colour_autotemp(const oimg_t *clut, const float *f);
/* with the lut inputs (clut: rg; abney: rg; spectra: rgba; any of them null = not connected) */
void o_colour_main_lut(const oimg_t *in, oimg_t *out, const float *f, int out_f16, const oimg_t *clut, const oimg_t *abney, const oimg_t *spectra, float auto_temp);
void o_xyY_to_dt_UCS_JCH(const float *xyY, float L_white, float *JCH);
void o_dt_UCS_JCH_to_xyY(const float *JCH, float L_white, float *xyY);

/* filmcurv */
void o_adjust_colour_dng(const float *col0, float *col1);
void o_filmcurv_px(const float *col_in, float *col1, const o_filmcurv_params_t *p);
void o_filmcurv_main(const oimg_t *in, oimg_t *out, const o_filmcurv_params_t *p, int out_f16);

/* llap + grade */
void o_llap_curve(const oimg_t *in, oimg_t *out11, const o_llap_params_t *p);
void o_llap_reduce(const oimg_t *in, oimg_t *out);
void o_llap_assemble(const oimg_t *coarse, const oimg_t *l0, const oimg_t *l1, oimg_t *out, int first);
void o_llap_colour(const oimg_t *lum, const oimg_t *org, oimg_t *out, int out_f16);
void o_llap_module(const oimg_t *in, oimg_t *out, const o_llap_params_t *p, int out_f16);
void o_grade_main(const oimg_t *in, oimg_t *out, const o_grade_params_t *q, int out_f16);

/* whole modules as wired by the reference's create_nodes */
void o_hilite_module(const oimg_t *in, oimg_t *out, const o_hilite_params_t *p, const float *wb4, uint32_t filters);
void o_demosaic_module(const oimg_t *in, oimg_t *out, const o_demosaic_params_t *p, uint32_t filters);
void o_denoise_module(const oimg_t *in_unorm, oimg_t *out, const o_denoise_params_t *p, const int *crop, const float *wb4,
    const float *black4, const float *white4, float noise_a, float noise_b, uint32_t filters);

/* the default darkroom graph (bin/default-darkroom.i-raw): everything a frame needs, flat so that ctypes can fill it */
typedef struct o_darkroom_t
{
  uint32_t width, height;        /* decoded mosaic dimensions */
  uint32_t filters;              /* 9 = x-trans, other non-zero = bayer rggb */
  uint32_t crop_aabb[4];
  float black[4], white[4];      /* raw units (u16 scale) */
  float whitebalance[4];
  float cam_to_rec2020[9];
  float noise_a, noise_b;
  uint32_t orientation;
  int colour_primaries, colour_trc;
  o_denoise_params_t denoise;
  o_hilite_params_t hilite;
  o_demosaic_params_t demosaic;
  o_crop_params_t crop;
  o_colour_params_t colour;
  o_filmcurv_params_t filmcurv;
  o_llap_params_t llap;
  o_grade_params_t grade;
  int enable_llap, enable_grade; /* graph variants without these modules */
  /* export side (graph-export.c:66-86): colenc behind the last module (prim / trc of colenc/params), sink image format */
  int enable_colenc, colenc_prim, colenc_trc;
  int sink_unorm8;               /* 1: the sink image is rgba ui8 (o-jpg): values 0..255 returned as floats */
} o_darkroom_t;

void o_darkroom_defaults(o_darkroom_t *d, uint32_t width, uint32_t height);
/* dng gain maps for the denoise module of the next runs (rgba f32 texture + origin / inverse extent); gm = 0 removes them */
void o_set_gainmap(const oimg_t *gm, const float *map_os);
/* the tables o_darkroom_run's colour step reads (null: not connected); they stay set until cleared */
void o_set_colour_luts(const oimg_t *clut, const oimg_t *abney, const oimg_t *spectra);
void o_colenc_px(float *rgb, int prim, int trc);
float o_unorm8(float v);
void o_colenc_main(const oimg_t *in, oimg_t *out, int prim, int trc, int fmt);
/* output dimensions of the sink for this configuration */
void o_darkroom_out_size(const o_darkroom_t *d, uint32_t *out_w, uint32_t *out_h);
/* raw: width*height u16.  out: out_w*out_h*4 floats (rgba f32, what o-pfm receives).
 * stage_out (optional, may be 0): if stage >= 0, a copy of an intermediate image is written instead:
 *   1 denoise out (1ch)  2 hilite out (1ch)  3 demosaic out (4ch)  4 crop  5 colour  6 filmcurv  7 llap */
int o_darkroom_run(const o_darkroom_t *d, const uint16_t *raw, float *out, int stage, float *stage_out);

/* o-pfm/main.c:8-42 */
int o_write_pfm(const char *filename, const float *rgba, int width, int height);

#ifdef __cplusplus
}
#endif

/* ORACLE — test infrastructure only (see o_common.h).
 * CPU restatement of src/pipe/modules/colenc/main.comp:17-81: rec2020 -> output primaries, then the output transfer curve.
 * vkdt-cli inserts this module in front of every 8 bit sink and whenever --colour-prim / --colour-trc differ from
 * bt2020 / linear (src/pipe/graph-export.c:66-86); its default (sRGB primaries = 1, trc = parse_prim("sRGB") = 1 = the
 * rec709 curve, cli/main.c:58-59) is what `vkdt-cli -g x.cfg` writes into a jpg.
 * constants: glslang folds constant expressions in double and rounds once ((float)(1.0/2.2), not 1.0f/2.2f).
 * store: f32 / f16 like every image, or UNORM8: clamp to [0,1], scale by 255, round to nearest even (what Mesa's
 * float_to_ubyte and the hardware conversion do; the Vulkan spec leaves ties open); the 8 bit value is returned as a float. */
#include "o_common.h"
#include "vkdt_oracle.h"

static const float M_2020_to_709[9]   = {1.66022677f, -0.58754761f, -0.07283825f, -0.12455334f, 1.13292605f, -0.00834963f, -0.01815514f, -0.10060303f, 1.11899817f};
static const float M_2020_to_adobe[9] = {1.15194302f, -0.09753232f, -0.05448118f, -0.12454585f, 1.13290963f, -0.00837122f, -0.02253539f, -0.04979918f, 1.07275365f};
static const float M_2020_to_p3d65[9] = {1.34353337f, -0.28218904f, -0.06142427f, -0.06530851f, 1.07578268f, -0.01048453f, 0.00282971f, -0.01961215f, 1.01717851f};
static const float M_2020_to_xyz[9]   = {0.636958048301290991f, 0.144616903586208406f, 0.168880975164172054f, 0.26270021201126692f, 0.677998071518871148f, 0.0593017164698619384f, 4.9999999999999999e-17f, 0.0280726930490874452f, 1.06098505771079066f};
static const float M_2020_to_ap0[9]   = {6.68685575e-01f, 1.51817679e-01f, 1.77189677e-01f, 4.49002044e-02f, 8.62145497e-01f, 1.01922441e-01f, -2.66851927e-09f, 2.78271109e-02f, 1.05170358f};
static const float M_2020_to_ap1[9]   = {9.62918591e-01f, 1.16137050e-02f, 2.55863361e-02f, 4.16800770e-04f, 9.99378426e-01f, -8.82457347e-05f, 5.31123331e-03f, 2.18655328e-02f, 9.75907920e-01f};
static const float M_2020_to_redwg[9] = {0.853263f, 0.079695f, 0.067042f, 0.029375f, 0.809195f, 0.161430f, 0.051575f, 0.208097f, 0.740329f};

void o_colenc_px(float *rgb, int prim, int trc)
{
  const float *M = prim == 1 ? M_2020_to_709 : prim == 3 ? M_2020_to_adobe : prim == 4 ? M_2020_to_p3d65 : prim == 5 ? M_2020_to_xyz :
                   prim == 6 ? M_2020_to_ap0 : prim == 7 ? M_2020_to_ap1 : prim == 10 ? M_2020_to_redwg : 0;
  if(M) o_mat3mulv(M, rgb, rgb);
  for(int k = 0; k < 3; k++)
  {
    float v = rgb[k];
    if(trc == 1)
    {
      const float a = 1.09929682680944f, b = 0.018053968510807f;
      v = v > b ? powf(v, (float)(1.0 / 2.2)) * a - (float)(1.09929682680944 - 1.0) : v * 4.5f; /* a - 1: a constant, folded in double */
    }
    else if(trc == 2) v = v > 0.0031308f ? powf(v, (float)(1.0 / 2.4)) * 1.055f - 0.055f : v * 12.92f;
    else if(trc == 3)
    {
      const float c3 = (float)(2392.0 / 128.0), c2 = (float)(2413.0 / 128.0), c1 = c3 - c2 + 1.0f;
      const float m1 = (float)(1305.0 / 8192.0), m2 = (float)(2523.0 / 32.0);
      v = o_max(0.0f, v);
      v = powf(v, m1);
      const float num = (c1 - 1.0f) + (c2 - c3) * v, den = 1.0f + c3 * v;
      v = powf(1.0f + num / den, m2);
    }
    else if(trc == 4) v = powf(v, (float)(1.0 / 2.6));
    else if(trc == 5)
    {
      const float a = 0.17883277f, b = 1.0f - 4.0f * a, c = 0.5f - a * logf(4.0f * a);
      v = v > (float)(1.0 / 12.0) ? a * logf(12.0f * v - b) + c : sqrtf(3.0f * v);
    }
    else if(trc == 6) v = powf(v, (float)(1.0 / 2.2));
    rgb[k] = v;
  }
}

float o_unorm8(float v)
{ /* NaN -> 0 */
  if(!(v > 0.0f)) return 0.0f;
  if(v > 1.0f) v = 1.0f;
  return rintf(v * 255.0f);
}

/* fmt: 0 f32, 1 f16, 2 unorm8 (values 0..255 as floats, alpha 255) */
void o_colenc_main(const oimg_t *in, oimg_t *out, int prim, int trc, int fmt)
{
#pragma omp parallel for schedule(static)
  for(int y = 0; y < out->h; y++) for(int x = 0; x < out->w; x++)
  {
    float rgb[4];
    o_fetch4(in, x, y, rgb);
    o_colenc_px(rgb, prim, trc);
    rgb[3] = 1.0f;
    if(fmt == 2) { for(int k = 0; k < 4; k++) rgb[k] = o_unorm8(rgb[k]); o_store4(out, x, y, rgb, 0); }
    else o_store4(out, x, y, rgb, fmt == 1);
  }
}

/* ORACLE — test infrastructure.  the reference's own DNG opcode list decoder (src/pipe/modules/i-raw/dng_opcode_decode.c),
 * compiled where it lies under /root/reference by `make -C oracle ref` (never copied): decodes an OpcodeList tag and writes what
 * it found as text, in the form of the product's vkb_dng_opcodes_describe. */
#include <stdio.h>
#include "pipe/modules/i-raw/dng_opcode_decode.c"

int ref_dngop_describe(const unsigned char *blob, int len, char *out, int outsize)
{
  dt_dng_opcode_list_t *ol = dng_opcode_list_decode((uint8_t *)blob, (size_t)len);
  int n = 0;
  if(!ol) { n = snprintf(out, outsize, "none\n"); return n; }
  n += snprintf(out + n, outsize - n, "count %d\n", ol->count);
  for(int i = 0; i < ol->count && n < outsize - 512; i++)
  {
    const dt_dng_opcode_t *op = ol->ops + i;
    n += snprintf(out + n, outsize - n, "op %u optional %u preview_skip %u\n", (unsigned)op->id, op->optional, op->preview_skip);
    if(op->id == s_dngop_gain_map)
    {
      const dt_dng_gain_map_t *g = (const dt_dng_gain_map_t *)op->data;
      n += snprintf(out + n, outsize - n, " region %u %u %u %u plane %u %u pitch %u %u points %u %u spacing %.17g %.17g origin %.17g %.17g planes %u\n gains",
          g->region.top, g->region.left, g->region.bottom, g->region.right, g->region.plane, g->region.planes, g->region.row_pitch, g->region.col_pitch,
          g->map_points_v, g->map_points_h, g->map_spacing_v, g->map_spacing_h, g->map_origin_v, g->map_origin_h, g->map_planes);
      const size_t cnt = (size_t)g->map_points_h * g->map_points_v * g->map_planes;
      for(size_t k = 0; k < cnt && n < outsize - 32; k++) { unsigned u; memcpy(&u, g->map_gain + k, 4); n += snprintf(out + n, outsize - n, " %08x", u); }
      n += snprintf(out + n, outsize - n, "\n");
    }
  }
  dng_opcode_list_free(ol);
  return n;
}

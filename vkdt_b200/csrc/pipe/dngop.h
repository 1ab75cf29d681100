// DNG opcode lists as far as the path reads them (src/pipe/dng_opcode.h, i-raw/dng_opcode_decode.c, metadata.h
// dt_image_metadata_dngop_t): the list structure of an OpcodeList tag is validated like the reference does, GainMap opcodes
// are decoded, every other opcode is kept by id only (the reference decodes them for modules outside this path).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector>

struct dt_dng_gain_map_t
{ // dng_opcode.h:52-64
  uint32_t top, left, bottom, right, plane, planes, row_pitch, col_pitch;   // dt_dng_region_t
  uint32_t map_points_v, map_points_h;
  double   map_spacing_v, map_spacing_h, map_origin_v, map_origin_h;
  uint32_t map_planes;
  std::vector<float> map_gain;
};
struct dt_dng_opcode_t { uint32_t id, optional, preview_skip; int gain_map; };   // gain_map: index into gain_maps, -1 for other opcodes
struct dt_dng_opcode_list_t { std::vector<dt_dng_opcode_t> ops; std::vector<dt_dng_gain_map_t> gain_maps; };
// what img_param.meta carries from the source module to denoise: OpcodeList2 and the cfa offset of the emitted window
struct dt_image_metadata_dngop_t { int ox = 0, oy = 0; dt_dng_opcode_list_t list2; };

// dng_opcode_list_decode (i-raw/dng_opcode_decode.c:311-346) for the tag's bytes (big endian).  0 on success; an empty or
// malformed list (sizes that do not add up, a GainMap whose length disagrees with its point counts) decodes to nothing: 1
int dng_opcode_list_decode(const uint8_t *data, size_t len, dt_dng_opcode_list_t *out);

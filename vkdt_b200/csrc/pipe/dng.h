// minimal reader for uncompressed 16-bit CFA DNG files: the one container in which a (synthetic) still can be handed to
// both engines.  TIFF/EP + DNG 1.4 tags only; camera raw formats proper stay with rawspeed / rawler.
#pragma once
#include <stdint.h>
#include <vector>

struct dng_image_t
{
  std::vector<uint16_t> pix;       // width*height, as stored
  uint32_t width = 0, height = 0;
  uint32_t cfa_dim = 0;            // 2 (bayer) or 6 (x-trans)
  uint8_t  cfa[36] = {0};          // CFAPattern, 0 r 1 g 2 b, row major cfa_dim x cfa_dim
  float    black[4] = {0, 0, 0, 0};
  float    white = 65535.0f;
  float    neutral[3] = {1, 1, 1}; // AsShotNeutral
  float    color_matrix[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; // xyz -> camera (ColorMatrix2 if present, else ColorMatrix1)
  int      illuminant = 21;        // EXIF LightSource of that matrix (21 = D65)
  uint32_t active[4] = {0, 0, 0, 0}; // ActiveArea top left bottom right (0 0 h w when absent)
  char     make[32] = {0}, model[32] = {0};
  float    iso = 0.0f;
  uint32_t orientation = 0;
  std::vector<uint8_t> opcode_list2; // the OpcodeList2 tag (51009) as stored: big endian whatever the file's byte order
};

// 0 on success
int dng_read(const char *filename, dng_image_t *img);

struct vkb_raw_params_t;
// what the reference's loader derives from the decoded file (i-raw/rawloader-c/lib.rs:137-279 + i-raw/main.c:138-256):
// aligned dimensions, cfa offset (ox, oy) of the emitted window, crop box, levels, normalised white balance,
// camera -> rec2020 matrix.  0 on success.
int dng_raw_params(const dng_image_t *img, vkb_raw_params_t *p, uint32_t *ox, uint32_t *oy);

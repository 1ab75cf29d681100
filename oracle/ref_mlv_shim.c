/* ORACLE — test infrastructure only.
 * Thin shim around the REFERENCE's own MLV decoder (src/pipe/modules/i-mlv/video_mlv.c, compiled in place
 * from /root/reference by `make ref`; no reference source is copied into this repo).  It calls the same two
 * functions i-mlv/main.c:60-85 calls: mlv_open_clip() and mlv_get_frame().  Used to pin o_mlv_unpack and the
 * CUDA unpack kernel bit-exactly, and to generate tests/golden/mlv_*.bin. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "video_mlv.h"

int ref_mlv_info(const char *filename, int *info /* w h bpp black white frames */)
{
  mlv_header_t v;
  memset(&v, 0, sizeof(v));
  if(mlv_open_clip(&v, filename, 0)) return 1;
  info[0] = v.RAWI.xRes; info[1] = v.RAWI.yRes; info[2] = v.RAWI.raw_info.bits_per_pixel;
  info[3] = v.RAWI.raw_info.black_level; info[4] = v.RAWI.raw_info.white_level;
  info[5] = v.MLVI.videoFrameCount;
  mlv_header_cleanup(&v);
  return 0;
}

int ref_mlv_decode(const char *filename, uint64_t frame, uint16_t *out)
{
  mlv_header_t v;
  memset(&v, 0, sizeof(v));
  if(mlv_open_clip(&v, filename, 0)) return 1;
  const int r = mlv_get_frame(&v, frame, out);
  mlv_header_cleanup(&v);
  return r;
}

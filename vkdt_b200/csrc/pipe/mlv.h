// minimal Magic Lantern Video (MLV v2) container reader for the i-mlv source: block index + raw payload access.
// own implementation from the block layout (reference behaviour: src/pipe/modules/i-mlv/
// video_mlv.c:298-634 mlv_open_clip, :197-278 mlv_get_frame).  uncompressed clips: the payload is handed to the
// GPU still packed (bits_per_pixel/8 bytes per pixel) and unpacked there.  lossless (LJ92) clips: decoded on the host, lj92.cpp.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

struct mlv_frame_t { uint64_t timestamp; uint64_t payload_offset; uint32_t payload_size; uint32_t frame_number; };

struct mlv_clip_t
{
  FILE *file = 0;
  std::string filename;
  uint32_t width = 0, height = 0, bpp = 0;
  int32_t  black = 0, white = 0;
  uint32_t frame_count = 0;      // MLVI.videoFrameCount (drives graph->frame_cnt, i-mlv/main.c:150)
  uint32_t fps_nom = 0, fps_denom = 0;
  uint16_t video_class = 0;
  bool     lossless = false;     // frames are lossless jpeg streams (decoded on the host), not packed bits
  char camera_name[32] = {0};
  std::vector<mlv_frame_t> frames; // sorted by timestamp like the reference's index
};

int  mlv_open(mlv_clip_t *c, const char *filename);
void mlv_close(mlv_clip_t *c);
// bytes of one packed frame incl. the 16-byte tail padding the unpack kernel may touch
size_t mlv_packed_bytes(uint32_t width, uint32_t height, uint32_t bpp);
// reads frame `idx` (index into the timestamp sorted list) into dst (mlv_packed_bytes() bytes available)
int  mlv_read_packed(mlv_clip_t *c, uint32_t idx, void *dst);
// lossless clips: decodes frame `idx` into width*height u16
int  mlv_read_lossless(mlv_clip_t *c, uint32_t idx, uint16_t *dst);

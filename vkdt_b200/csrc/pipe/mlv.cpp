#include "mlv.h"
#include "lj92.h"
#include <string.h>
#include <algorithm>

static uint32_t rd32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

size_t mlv_packed_bytes(uint32_t width, uint32_t height, uint32_t bpp)
{
  const size_t payload = ((size_t)width * height * bpp + 7) / 8;
  return ((payload + 15) / 16) * 16 + 16;
}

void mlv_close(mlv_clip_t *c)
{
  if(c->file) fclose(c->file);
  c->file = 0;
  c->frames.clear();
}

int mlv_open(mlv_clip_t *c, const char *filename)
{
  mlv_close(c);
  FILE *f = fopen(filename, "rb");
  if(!f) return 1;
  uint8_t hdr[16];
  bool have_mlvi = false, have_rawi = false;
  uint64_t pos = 0;
  fseek(f, 0, SEEK_END);
  const uint64_t fsize = (uint64_t)ftell(f);
  while(pos + 8 <= fsize)
  { // every block: 4 byte type, u32 blockSize; all but MLVI carry a u64 timestamp next
    fseek(f, (long)pos, SEEK_SET);
    if(fread(hdr, 1, 8, f) != 8) break;
    const uint32_t bsize = rd32(hdr + 4);
    if(bsize < 8 || pos + bsize > fsize) break;
    if(!memcmp(hdr, "MLVI", 4))
    { // fileMagic blockSize versionString[8] fileGuid fileNum fileCount fileFlags videoClass audioClass videoFrameCount audioFrameCount fpsNom fpsDenom
      uint8_t b[52];
      fseek(f, (long)pos, SEEK_SET);
      if(bsize < 52 || fread(b, 1, 52, f) != 52) break;
      c->video_class = rd16(b + 32);
      c->frame_count = rd32(b + 36);
      c->fps_nom = rd32(b + 44); c->fps_denom = rd32(b + 48);
      have_mlvi = true;
    }
    else if(!memcmp(hdr, "RAWI", 4))
    { // type size timestamp xRes(u16) yRes(u16) raw_info{api_version, buffer, height, width, pitch, frame_size, bits_per_pixel, black_level, white_level, ...}
      uint8_t b[20 + 36];
      fseek(f, (long)pos, SEEK_SET);
      if(bsize < sizeof(b) || fread(b, 1, sizeof(b), f) != sizeof(b)) break;
      c->width = rd16(b + 16); c->height = rd16(b + 18);
      c->bpp = rd32(b + 20 + 24);
      c->black = (int32_t)rd32(b + 20 + 28);
      c->white = (int32_t)rd32(b + 20 + 32);
      have_rawi = true;
    }
    else if(!memcmp(hdr, "IDNT", 4))
    {
      uint8_t b[16 + 32];
      fseek(f, (long)pos, SEEK_SET);
      if(bsize >= sizeof(b) && fread(b, 1, sizeof(b), f) == sizeof(b)) { memcpy(c->camera_name, b + 16, 31); c->camera_name[31] = 0; }
    }
    else if(!memcmp(hdr, "VIDF", 4))
    { // type size timestamp frameNumber cropPosX cropPosY panPosX panPosY frameSpace, then frameSpace pad bytes, then payload
      uint8_t b[32];
      fseek(f, (long)pos, SEEK_SET);
      if(bsize < 32 || fread(b, 1, 32, f) != 32) break;
      mlv_frame_t fr;
      fr.timestamp = rd64(b + 8);
      fr.frame_number = rd32(b + 16);
      const uint32_t space = rd32(b + 28);
      if(space > bsize - 32) { pos += bsize; continue; } // frameSpace larger than the block: not a frame we can trust
      fr.payload_offset = pos + 32 + space;
      fr.payload_size = bsize - 32 - space;
      c->frames.push_back(fr);
    }
    pos += bsize;
  }
  if(!have_mlvi || !have_rawi || c->frames.empty() || !c->width || !c->height) { fclose(f); c->frames.clear(); return 1; }
  c->lossless = (c->video_class & 0x20) != 0; // MLV_VIDEO_CLASS_FLAG_LJ92 (video_mlv.c:227)
  if(c->lossless && (c->bpp < 8 || c->bpp > 16)) { fclose(f); c->frames.clear(); fprintf(stderr, "[i-mlv] implausible bit depth %u\n", c->bpp); return 1; }
  if(!c->lossless && c->bpp != 10 && c->bpp != 12 && c->bpp != 14) { fclose(f); c->frames.clear(); fprintf(stderr, "[i-mlv] unsupported bit depth %u\n", c->bpp); return 1; }
  std::stable_sort(c->frames.begin(), c->frames.end(), [](const mlv_frame_t &a, const mlv_frame_t &b) { return a.timestamp < b.timestamp; });
  if(c->frame_count == 0 || c->frame_count > c->frames.size()) c->frame_count = (uint32_t)c->frames.size();
  c->file = f;
  return 0;
}

int mlv_read_packed(mlv_clip_t *c, uint32_t idx, void *dst)
{
  if(!c->file || idx >= c->frames.size()) return 1;
  const size_t payload = ((size_t)c->width * c->height * c->bpp + 7) / 8;
  const size_t total = mlv_packed_bytes(c->width, c->height, c->bpp);
  if(c->frames[idx].payload_size < payload) return 1;
  fseek(c->file, (long)c->frames[idx].payload_offset, SEEK_SET);
  if(fread(dst, 1, payload, c->file) != payload) return 1;
  memset((uint8_t *)dst + payload, 0, total - payload);
  return 0;
}

// lossless clips: the frame is one lossless jpeg stream; decoded on the host (video_mlv.c:224-251 does the same with
// liblj92) straight into the staging buffer as width*height u16
int mlv_read_lossless(mlv_clip_t *c, uint32_t idx, uint16_t *dst)
{
  if(!c->file || idx >= c->frames.size() || !c->lossless) return 1;
  std::vector<uint8_t> buf(c->frames[idx].payload_size);
  fseek(c->file, (long)c->frames[idx].payload_offset, SEEK_SET);
  if(buf.empty() || fread(buf.data(), 1, buf.size(), c->file) != buf.size()) return 1;
  int w = 0, h = 0, bits = 0, comps = 0;
  if(lj92_info(buf.data(), buf.size(), &w, &h, &bits, &comps)) return 1;
  if((size_t)w * h * comps != (size_t)c->width * c->height) { fprintf(stderr, "[i-mlv] lossless frame is %dx%dx%d, clip says %ux%u\n", w, h, comps, c->width, c->height); return 1; }
  return lj92_decode(buf.data(), buf.size(), dst, (size_t)c->width * c->height);
}

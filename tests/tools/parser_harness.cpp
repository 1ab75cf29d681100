// address/UB-sanitizer harness for the host-side parsers (built and driven by tests/test_parsers_fuzz_cpu.py).
// usage: parser_harness dng|lj92|mlv <file>...   exit code 0 unless a sanitizer aborts.
#include "../../vkdt_b200/csrc/pipe/dng.h"
#include "../../vkdt_b200/csrc/pipe/lj92.h"
#include "../../vkdt_b200/csrc/pipe/mlv.h"
#include "../../include/vkdt_b200.h"
#include <stdio.h>
#include <string.h>
#include <vector>

static std::vector<uint8_t> slurp(const char *fn)
{
  std::vector<uint8_t> d;
  FILE *f = fopen(fn, "rb");
  if(!f) return d;
  fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
  d.resize(n > 0 ? n : 0);
  if(n > 0 && fread(d.data(), 1, n, f) != (size_t)n) d.clear();
  fclose(f);
  return d;
}

int main(int argc, char **argv)
{
  if(argc < 3) return 2;
  for(int a = 2; a < argc; a++)
  {
    if(!strcmp(argv[1], "dng"))
    {
      dng_image_t img; vkb_raw_params_t p; uint32_t ox, oy;
      if(!dng_read(argv[a], &img)) dng_raw_params(&img, &p, &ox, &oy);
    }
    else if(!strcmp(argv[1], "lj92"))
    {
      const std::vector<uint8_t> d = slurp(argv[a]);
      int w = 0, h = 0, b = 0, c = 0;
      if(!d.empty() && !lj92_info(d.data(), d.size(), &w, &h, &b, &c) && (size_t)w * h * c < (1u << 24))
      {
        std::vector<uint16_t> out((size_t)w * h * c);
        lj92_decode(d.data(), d.size(), out.data(), out.size());
      }
    }
    else if(!strcmp(argv[1], "mlv"))
    {
      mlv_clip_t c;
      if(!mlv_open(&c, argv[a]))
      {
        if((size_t)c.width * c.height < (1u << 24))
        {
          std::vector<uint8_t> buf(mlv_packed_bytes(c.width, c.height, c.bpp ? c.bpp : 16) + (size_t)c.width * c.height * 2 + 64);
          for(uint32_t f = 0; f < c.frames.size() && f < 4; f++)
          {
            if(c.lossless) mlv_read_lossless(&c, f, (uint16_t *)buf.data());
            else mlv_read_packed(&c, f, buf.data());
          }
        }
        mlv_close(&c);
      }
    }
  }
  return 0;
}

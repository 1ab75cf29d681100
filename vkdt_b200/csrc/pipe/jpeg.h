// baseline JPEG writer behind the o-jpg sink (pipe/jpeg.cpp)
#pragma once
#include <stdint.h>
// rgba: width * height * 4 bytes (alpha ignored).  quality 1..100 (libjpeg's scale).  0 on success
int jpeg_write_rgba8(const char *filename, const uint8_t *rgba, int width, int height, float quality);

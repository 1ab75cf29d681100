// lossless JPEG (ITU-T T.81 annex H, "LJ92") decoder for MLV clips recorded with lossless compression
// (MLV_VIDEO_CLASS_FLAG_LJ92; the reference decodes them on the host with liblj92, i-mlv/video_mlv.c:224-251).
// own implementation from the standard: SOF3 frames, one scan, 1..4 interleaved components, predictors 1..7,
// point transform, no restart intervals.  entropy decoding is serial by nature and stays on the host like in the
// reference; the decoded u16 mosaic takes the same upload path as any other unpacked source.
#pragma once
#include <stddef.h>
#include <stdint.h>

// parses the headers: 0 on success
int lj92_info(const uint8_t *data, size_t size, int *width, int *height, int *bits, int *components);
// decodes width*height*components samples in scan order (components interleaved) into out[0..count): 0 on success
int lj92_decode(const uint8_t *data, size_t size, uint16_t *out, size_t count);

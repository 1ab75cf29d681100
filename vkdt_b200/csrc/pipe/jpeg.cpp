// baseline JPEG writer for the o-jpg sink (the reference's o-jpg/main.c hands its 8 bit sink image to libjpeg on the CPU;
// libjpeg's headers are not in this image, and entropy coding is host work either way).
// ISO/IEC 10918-1 baseline sequential DCT, 8 bit, YCbCr (JFIF 1.01, 300 dpi like o-jpg/main.c:141-143), the standard
// quantisation tables of Annex K scaled like libjpeg's jpeg_set_quality, the standard Huffman tables of Annex K.3, no
// chroma subsampling (what o-jpg selects above quality 92, :134-135; the default quality is 95).  every decoder reads it;
// the bytes are not libjpeg's (it optimises the Huffman tables), the decoded pixels agree to the quantisation step.
#include "jpeg.h"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <vector>

namespace {
const uint8_t zigzag[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
const uint8_t qlum[64] = { 16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
  18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };
const uint8_t qchr[64] = { 17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 };
// Annex K.3: number of codes per length 1..16, then the symbols
const uint8_t dc_lum_n[16] = { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 }, dc_chr_n[16] = { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 };
const uint8_t dc_sym[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
const uint8_t ac_lum_n[16] = { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d }, ac_chr_n[16] = { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 };
const uint8_t ac_lum_sym[162] = { 0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
  0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
  0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
  0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
  0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
  0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
  0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };
const uint8_t ac_chr_sym[162] = { 0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
  0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
  0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
  0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
  0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
  0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
  0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };

struct huff_t { uint16_t code[256]; uint8_t len[256]; };
void build(const uint8_t *n, const uint8_t *sym, huff_t *h)
{ // Annex C: canonical codes in order of length
  memset(h, 0, sizeof(*h));
  int code = 0, k = 0;
  for(int l = 1; l <= 16; l++) { for(int i = 0; i < n[l - 1]; i++, k++) { h->code[sym[k]] = (uint16_t)code++; h->len[sym[k]] = (uint8_t)l; } code <<= 1; }
}
struct bits_t
{
  std::vector<uint8_t> out; uint32_t acc = 0; int cnt = 0;
  void put(uint32_t v, int n)
  {
    acc = (acc << n) | (v & ((1u << n) - 1)); cnt += n;
    while(cnt >= 8) { const uint8_t b = (uint8_t)(acc >> (cnt - 8)); out.push_back(b); if(b == 0xff) out.push_back(0); cnt -= 8; }
  }
  void flush() { if(cnt) put(0x7f, 8 - cnt); }
};
void fdct8x8(float *b)
{ // separable DCT-II with the JPEG normalisation, double accumulation: exact to rounding, 4 MB/s per core does not matter here
  static float c[8][8]; static bool init = false;
  if(!init) { for(int u = 0; u < 8; u++) for(int x = 0; x < 8; x++) c[u][x] = (float)((u ? 1.0 : sqrt(0.5)) * 0.5 * cos((2 * x + 1) * u * M_PI / 16.0)); init = true; }
  float t[64];
  for(int y = 0; y < 8; y++) for(int u = 0; u < 8; u++) { double s = 0; for(int x = 0; x < 8; x++) s += c[u][x] * b[8 * y + x]; t[8 * y + u] = (float)s; }
  for(int u = 0; u < 8; u++) for(int v = 0; v < 8; v++) { double s = 0; for(int y = 0; y < 8; y++) s += c[v][y] * t[8 * y + u]; b[8 * v + u] = (float)s; }
}
void block(bits_t &bw, float *b, const uint8_t *q, const huff_t &dc, const huff_t &ac, int *pred)
{
  fdct8x8(b);
  int z[64];
  for(int i = 0; i < 64; i++) z[i] = (int)lrintf(b[zigzag[i]] / (float)q[zigzag[i]]);
  auto mag = [](int v, int *nb, uint32_t *bits) { int a = v < 0 ? -v : v, n = 0; while(a >> n) n++; *nb = n; *bits = (uint32_t)(v < 0 ? v - 1 : v); };
  int nb; uint32_t bits;
  mag(z[0] - *pred, &nb, &bits); *pred = z[0];
  bw.put(dc.code[nb], dc.len[nb]); if(nb) bw.put(bits, nb);
  int run = 0;
  for(int i = 1; i < 64; i++)
  {
    if(!z[i]) { run++; continue; }
    while(run > 15) { bw.put(ac.code[0xf0], ac.len[0xf0]); run -= 16; }
    mag(z[i], &nb, &bits);
    bw.put(ac.code[(run << 4) | nb], ac.len[(run << 4) | nb]); bw.put(bits, nb);
    run = 0;
  }
  if(run) bw.put(ac.code[0], ac.len[0]);
}
} // namespace

int jpeg_write_rgba8(const char *filename, const uint8_t *rgba, int width, int height, float quality)
{
  if(width <= 0 || height <= 0 || width > 65535 || height > 65535) return 1;
  FILE *f = fopen(filename, "wb");
  if(!f) return 2;
  int q = (int)quality; if(q < 1) q = 1; if(q > 100) q = 100;
  const int scale = q < 50 ? 5000 / q : 200 - 2 * q;                           // jpeg_quality_scaling
  uint8_t ql[64], qc[64];
  for(int i = 0; i < 64; i++)
  {
    int a = (qlum[i] * scale + 50) / 100, b = (qchr[i] * scale + 50) / 100;
    ql[i] = (uint8_t)(a < 1 ? 1 : (a > 255 ? 255 : a)); qc[i] = (uint8_t)(b < 1 ? 1 : (b > 255 ? 255 : b));
  }
  std::vector<uint8_t> h;
  auto w16 = [&](int v) { h.push_back((uint8_t)(v >> 8)); h.push_back((uint8_t)v); };
  h.push_back(0xff); h.push_back(0xd8);
  h.push_back(0xff); h.push_back(0xe0); w16(16); for(char c : { 'J', 'F', 'I', 'F', '\0' }) h.push_back((uint8_t)c); h.push_back(1); h.push_back(1); h.push_back(1); w16(300); w16(300); h.push_back(0); h.push_back(0);
  for(int t = 0; t < 2; t++) { h.push_back(0xff); h.push_back(0xdb); w16(67); h.push_back((uint8_t)t); for(int i = 0; i < 64; i++) h.push_back((t ? qc : ql)[zigzag[i]]); }
  h.push_back(0xff); h.push_back(0xc0); w16(17); h.push_back(8); w16(height); w16(width); h.push_back(3);
  for(int c = 0; c < 3; c++) { h.push_back((uint8_t)(c + 1)); h.push_back(0x11); h.push_back(c ? 1 : 0); }
  const uint8_t *ns[4] = { dc_lum_n, ac_lum_n, dc_chr_n, ac_chr_n }; const uint8_t *sy[4] = { dc_sym, ac_lum_sym, dc_sym, ac_chr_sym };
  const int cls[4] = { 0x00, 0x10, 0x01, 0x11 };
  for(int t = 0; t < 4; t++)
  {
    int cnt = 0; for(int i = 0; i < 16; i++) cnt += ns[t][i];
    h.push_back(0xff); h.push_back(0xc4); w16(19 + cnt); h.push_back((uint8_t)cls[t]);
    for(int i = 0; i < 16; i++) h.push_back(ns[t][i]);
    for(int i = 0; i < cnt; i++) h.push_back(sy[t][i]);
  }
  h.push_back(0xff); h.push_back(0xda); w16(12); h.push_back(3); h.push_back(1); h.push_back(0x00); h.push_back(2); h.push_back(0x11); h.push_back(3); h.push_back(0x11);
  h.push_back(0); h.push_back(63); h.push_back(0);
  fwrite(h.data(), 1, h.size(), f);
  huff_t hd[2], ha[2];
  build(dc_lum_n, dc_sym, &hd[0]); build(ac_lum_n, ac_lum_sym, &ha[0]); build(dc_chr_n, dc_sym, &hd[1]); build(ac_chr_n, ac_chr_sym, &ha[1]);
  bits_t bw;
  int pred[3] = { 0, 0, 0 };
  for(int by = 0; by < height; by += 8) for(int bx = 0; bx < width; bx += 8)
  {
    float Y[64], Cb[64], Cr[64];
    for(int j = 0; j < 8; j++) for(int i = 0; i < 8; i++)
    { // edge blocks repeat the last row / column
      const int x = bx + i < width ? bx + i : width - 1, y = by + j < height ? by + j : height - 1;
      const uint8_t *p = rgba + ((size_t)y * width + x) * 4;
      const float r = p[0], g = p[1], b = p[2];
      Y[8 * j + i]  =  0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
      Cb[8 * j + i] = -0.168735892f * r - 0.331264108f * g + 0.5f * b;
      Cr[8 * j + i] =  0.5f * r - 0.418687589f * g - 0.081312411f * b;
    }
    block(bw, Y, ql, hd[0], ha[0], &pred[0]);
    block(bw, Cb, qc, hd[1], ha[1], &pred[1]);
    block(bw, Cr, qc, hd[1], ha[1], &pred[2]);
    if(bw.out.size() > (1u << 20)) { fwrite(bw.out.data(), 1, bw.out.size(), f); bw.out.clear(); }
  }
  bw.flush();
  bw.out.push_back(0xff); bw.out.push_back(0xd9);
  fwrite(bw.out.data(), 1, bw.out.size(), f);
  fclose(f);
  return 0;
}

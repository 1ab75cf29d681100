#include "lj92.h"
#include <string.h>

namespace {
struct huff_t
{ // canonical huffman table in the decoding form of T.81 F.2.2.3: per code length the smallest / largest code and the
  // index of its first symbol, plus a 9-bit prefix lookup for the common short codes
  int     maxcode[18]; // -1: no code of this length
  int     mincode[17];
  int     valptr[17];
  uint8_t symbol[256];
  uint8_t look_len[512], look_sym[512];
  bool    present = false;
};

struct frame_t
{
  int precision = 0, width = 0, height = 0, ncomp = 0;
  int comp_id[4] = {0}, comp_table[4] = {0};
  int predictor = 1, pt = 0;
  huff_t huff[4];
  size_t scan_begin = 0;
};

bool build_table(huff_t *h, const uint8_t *counts, const uint8_t *symbols, int nsym)
{
  int code = 0, k = 0;
  memset(h->look_len, 0, sizeof(h->look_len));
  for(int len = 1; len <= 16; len++)
  {
    h->valptr[len] = k;
    h->mincode[len] = code;
    for(int i = 0; i < counts[len - 1]; i++, k++, code++)
    {
      if(k >= nsym) return false;
      h->symbol[k] = symbols[k];
      if(len <= 9)
      { // every 9-bit word that starts with this code
        const int lo = code << (9 - len), n = 1 << (9 - len);
        for(int j = 0; j < n; j++) { h->look_len[lo + j] = (uint8_t)len; h->look_sym[lo + j] = symbols[k]; }
      }
    }
    h->maxcode[len] = counts[len - 1] ? code - 1 : -1;
    if(code > (1 << len)) return false; // over-subscribed
    code <<= 1;
  }
  h->maxcode[17] = 0x7fffffff;
  h->present = true;
  return true;
}

int parse(const uint8_t *d, size_t size, frame_t *f)
{
  if(size < 4 || d[0] != 0xff || d[1] != 0xd8) return 1; // SOI
  size_t p = 2;
  bool have_sof = false;
  while(p + 4 <= size)
  {
    if(d[p] != 0xff) return 1;
    const int marker = d[p + 1];
    if(marker == 0xff) { p++; continue; }      // fill byte
    const size_t len = ((size_t)d[p + 2] << 8) | d[p + 3];
    if(len < 2 || p + 2 + len > size) return 1;
    const uint8_t *s = d + p + 4;
    const size_t n = len - 2;
    if(marker == 0xc4)
    { // DHT: any number of tables
      size_t q = 0;
      while(q + 17 <= n)
      {
        const int tc = s[q] >> 4, th = s[q] & 15;
        int nsym = 0;
        for(int i = 0; i < 16; i++) nsym += s[q + 1 + i];
        if(tc != 0 || th > 3 || nsym > 256 || q + 17 + nsym > n) return 1;
        if(!build_table(&f->huff[th], s + q + 1, s + q + 17, nsym)) return 1;
        q += 17 + nsym;
      }
    }
    else if(marker == 0xc3)
    { // SOF3: lossless, huffman
      if(n < 6) return 1;
      f->precision = s[0]; f->height = (s[1] << 8) | s[2]; f->width = (s[3] << 8) | s[4]; f->ncomp = s[5];
      if(f->precision < 2 || f->precision > 16 || f->ncomp < 1 || f->ncomp > 4 || n < (size_t)(6 + 3 * f->ncomp)) return 1;
      for(int c = 0; c < f->ncomp; c++) f->comp_id[c] = s[6 + 3 * c];
      have_sof = true;
    }
    else if(marker == 0xdd) { if(n >= 2 && (((int)s[0] << 8) | s[1]) != 0) return 2; } // DRI: restart intervals are not supported
    else if(marker == 0xda)
    { // SOS
      if(!have_sof || n < 1 || s[0] != f->ncomp || n < (size_t)(4 + 2 * f->ncomp)) return 1;
      bool assigned[4] = { false, false, false, false };
      for(int c = 0; c < f->ncomp; c++)
      {
        int idx = -1;
        for(int k = 0; k < f->ncomp; k++) if(f->comp_id[k] == s[1 + 2 * c]) idx = k;
        if(idx < 0 || idx >= 4 || assigned[idx]) return 1;   // unknown component, or one listed twice (another would keep no table)
        assigned[idx] = true;
        f->comp_table[idx] = s[2 + 2 * c] >> 4;
        if(f->comp_table[idx] > 3 || !f->huff[f->comp_table[idx]].present) return 1;
      }
      for(int k = 0; k < f->ncomp; k++) if(!assigned[k]) return 1;
      f->predictor = s[1 + 2 * f->ncomp];
      f->pt = s[3 + 2 * f->ncomp] & 15;
      if(f->predictor < 1 || f->predictor > 7 || f->pt >= f->precision) return 1;
      f->scan_begin = p + 2 + len;
      return 0;
    }
    else if((marker >= 0xc0 && marker <= 0xcf) && marker != 0xc8 && marker != 0xcc) return 2; // another process: not lossless huffman
    p += 2 + len;
  }
  return 1;
}

struct bits_t
{ // msb first bit reader over the entropy coded segment: 0xff00 is a stuffed 0xff, any other marker ends the data
  const uint8_t *d; size_t size, pos; uint64_t acc = 0; int cnt = 0; bool end = false;
  void fill()
  {
    while(cnt <= 56)
    {
      uint64_t b = 0;
      if(!end && pos < size)
      {
        b = d[pos];
        if(b == 0xff)
        {
          if(pos + 1 < size && d[pos + 1] == 0x00) pos += 2;
          else { end = true; b = 0; }
        }
        else pos++;
      }
      else end = true;
      acc |= b << (56 - cnt);
      cnt += 8;
    }
  }
  inline uint32_t peek(int n) { return (uint32_t)(acc >> (64 - n)); }
  inline void skip(int n) { acc <<= n; cnt -= n; }
};

inline int decode_ssss(bits_t &b, const huff_t &h)
{
  if(b.cnt < 32) b.fill();
  const uint32_t w = b.peek(9);
  if(h.look_len[w]) { b.skip(h.look_len[w]); return h.look_sym[w]; }
  int code = (int)b.peek(10), len = 10;
  while(len <= 16 && (h.maxcode[len] < 0 || code > h.maxcode[len])) { len++; code = (int)b.peek(len); }
  if(len > 16) return -1;
  b.skip(len);
  return h.symbol[h.valptr[len] + code - h.mincode[len]];
}
} // namespace

int lj92_info(const uint8_t *data, size_t size, int *width, int *height, int *bits, int *components)
{
  frame_t f;
  const int r = parse(data, size, &f);
  if(r) return r;
  if(width) *width = f.width;
  if(height) *height = f.height;
  if(bits) *bits = f.precision;
  if(components) *components = f.ncomp;
  return 0;
}

int lj92_decode(const uint8_t *data, size_t size, uint16_t *out, size_t count)
{
  frame_t f;
  const int r = parse(data, size, &f);
  if(r) return r;
  const int nc = f.ncomp;
  const size_t row = (size_t)f.width * nc;
  if(count < row * f.height) return 3;
  bits_t b{ data, size, f.scan_begin };
  const int first = 1 << (f.precision - f.pt - 1);
  for(int y = 0; y < f.height; y++)
  {
    uint16_t *cur = out + (size_t)y * row;
    const uint16_t *up = cur - row;
    for(int x = 0; x < f.width; x++) for(int c = 0; c < nc; c++)
    {
      const size_t i = (size_t)x * nc + c;
      int pred;
      if(y == 0) pred = x == 0 ? first : cur[i - nc];      // H.1.2.1: first line predicts from the left
      else if(x == 0) pred = up[i];                        // first sample of a line: from above
      else
      {
        const int ra = cur[i - nc], rb = up[i], rc = up[i - nc];
        switch(f.predictor)
        {
          case 1: pred = ra; break;
          case 2: pred = rb; break;
          case 3: pred = rc; break;
          case 4: pred = ra + rb - rc; break;
          case 5: pred = ra + ((rb - rc) >> 1); break;
          case 6: pred = rb + ((ra - rc) >> 1); break;
          default: pred = (ra + rb) >> 1; break;
        }
      }
      const int ssss = decode_ssss(b, f.huff[f.comp_table[c]]);
      if(ssss < 0 || ssss > 16) return 4;
      int diff = 0;
      if(ssss == 16) diff = 32768;
      else if(ssss)
      {
        if(b.cnt < 32) b.fill();
        const int v = (int)b.peek(ssss);
        b.skip(ssss);
        diff = v < (1 << (ssss - 1)) ? v - (1 << ssss) + 1 : v;
      }
      cur[i] = (uint16_t)((pred + diff) & 0xffff);
    }
  }
  if(f.pt) for(size_t i = 0; i < row * f.height; i++) out[i] = (uint16_t)(out[i] << f.pt);
  return 0;
}

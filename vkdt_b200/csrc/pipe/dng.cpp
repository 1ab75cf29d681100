#include "dng.h"
#include "../../../include/vkdt_b200.h"
#include <stdio.h>
#include <string.h>
#include <string>

namespace {
struct tiff_t
{
  std::vector<uint8_t> d;
  bool be = false;
  uint16_t u16(size_t o) const { if(o + 2 > d.size()) return 0; return be ? (uint16_t)((d[o] << 8) | d[o+1]) : (uint16_t)(d[o] | (d[o+1] << 8)); }
  uint32_t u32(size_t o) const { if(o + 4 > d.size()) return 0; return be ? ((uint32_t)d[o] << 24) | (d[o+1] << 16) | (d[o+2] << 8) | d[o+3] : ((uint32_t)d[o+3] << 24) | (d[o+2] << 16) | (d[o+1] << 8) | d[o]; }
};
struct entry_t { uint16_t tag, type; uint32_t count; size_t value_off; };
const int type_size[13] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8};

bool find(const tiff_t &t, size_t ifd, uint16_t tag, entry_t *e)
{
  const int n = t.u16(ifd);
  for(int i = 0; i < n; i++)
  {
    const size_t o = ifd + 2 + 12 * (size_t)i;
    if(t.u16(o) != tag) continue;
    e->tag = tag; e->type = t.u16(o + 2); e->count = t.u32(o + 4);
    const size_t bytes = (e->type < 13 ? type_size[e->type] : 1) * (size_t)e->count;
    e->value_off = bytes <= 4 ? o + 8 : t.u32(o + 8);
    return true;
  }
  return false;
}
double value(const tiff_t &t, const entry_t &e, uint32_t i)
{
  const size_t o = e.value_off + (size_t)i * type_size[e.type < 13 ? e.type : 1];
  switch(e.type)
  {
    case 1: case 6: case 7: return o < t.d.size() ? t.d[o] : 0;
    case 3: return t.u16(o);
    case 8: return (int16_t)t.u16(o);
    case 4: return t.u32(o);
    case 9: return (int32_t)t.u32(o);
    case 5: { const double den = t.u32(o + 4); return den ? t.u32(o) / den : 0.0; }
    case 10: { const double den = (int32_t)t.u32(o + 4); return den ? (int32_t)t.u32(o) / den : 0.0; }
    case 11: { uint32_t v = t.u32(o); float f; memcpy(&f, &v, 4); return f; }
    default: return 0.0;
  }
}
bool is_raw_ifd(const tiff_t &t, size_t ifd)
{
  entry_t e;
  return find(t, ifd, 262, &e) && (int)value(t, e, 0) == 32803; // PhotometricInterpretation = CFA
}
} // namespace

int dng_read(const char *filename, dng_image_t *img)
{
  FILE *f = fopen(filename, "rb");
  if(!f) return 1;
  tiff_t t;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if(sz < 16) { fclose(f); return 1; }
  t.d.resize(sz);
  if(fread(t.d.data(), 1, sz, f) != (size_t)sz) { fclose(f); return 1; }
  fclose(f);
  if(t.d[0] == 'M' && t.d[1] == 'M') t.be = true;
  else if(!(t.d[0] == 'I' && t.d[1] == 'I')) return 2;
  if(t.u16(2) != 42) return 2;
  const size_t ifd0 = t.u32(4);
  size_t raw = 0;
  entry_t e;
  if(is_raw_ifd(t, ifd0)) raw = ifd0;
  else if(find(t, ifd0, 330, &e)) // SubIFDs
    for(uint32_t i = 0; i < e.count && !raw; i++) { const size_t s = (size_t)value(t, e, i); if(is_raw_ifd(t, s)) raw = s; }
  if(!raw) return 3;
  if(!find(t, raw, 256, &e)) return 4; img->width = (uint32_t)value(t, e, 0);
  if(!find(t, raw, 257, &e)) return 4; img->height = (uint32_t)value(t, e, 0);
  if(!img->width || !img->height || (uint64_t)img->width * img->height > (1ull << 31)) return 4; // nothing a sensor produces
  if(find(t, raw, 258, &e) && (int)value(t, e, 0) != 16) { fprintf(stderr, "[i-raw] dng: only 16 bits per sample are supported\n"); return 5; }
  if(find(t, raw, 259, &e) && (int)value(t, e, 0) != 1)  { fprintf(stderr, "[i-raw] dng: only uncompressed data is supported\n"); return 5; }
  if(find(t, raw, 277, &e) && (int)value(t, e, 0) != 1)  { fprintf(stderr, "[i-raw] dng: only one sample per pixel (cfa) is supported\n"); return 5; }
  entry_t so, sc;
  if(!find(t, raw, 273, &so) || !find(t, raw, 279, &sc)) { fprintf(stderr, "[i-raw] dng: tiled files are not supported\n"); return 5; }
  img->pix.assign((size_t)img->width * img->height, 0);
  size_t filled = 0;
  for(uint32_t s = 0; s < so.count; s++)
  {
    const size_t off = (size_t)value(t, so, s), cnt = (size_t)value(t, sc, s);
    if(off + cnt > t.d.size()) return 6;
    for(size_t k = 0; k + 1 < cnt && filled < img->pix.size(); k += 2) img->pix[filled++] = t.u16(off + k);
  }
  if(filled != img->pix.size()) return 6;
  if(find(t, raw, 33421, &e) && e.count == 2) img->cfa_dim = (uint32_t)value(t, e, 0);
  if(img->cfa_dim != 2 && img->cfa_dim != 6) { fprintf(stderr, "[i-raw] dng: cfa pattern dimension %u not supported\n", img->cfa_dim); return 7; }
  if(!find(t, raw, 33422, &e) || e.count != img->cfa_dim * img->cfa_dim) return 7;
  for(uint32_t i = 0; i < e.count; i++) img->cfa[i] = (uint8_t)value(t, e, i);
  if(find(t, raw, 50714, &e)) for(uint32_t i = 0; i < 4; i++) img->black[i] = (float)value(t, e, i < e.count ? i : e.count - 1);
  if(find(t, raw, 50717, &e)) img->white = (float)value(t, e, 0);
  img->active[0] = 0; img->active[1] = 0; img->active[2] = img->height; img->active[3] = img->width;
  if(find(t, raw, 50829, &e) && e.count == 4) for(int i = 0; i < 4; i++) img->active[i] = (uint32_t)value(t, e, i);
  img->opcode_list2.clear();
  if(find(t, raw, 51009, &e) && e.count >= 4 && e.value_off + (size_t)e.count <= t.d.size())
    img->opcode_list2.assign(t.d.begin() + e.value_off, t.d.begin() + e.value_off + e.count);
  // colour / identification tags live in IFD0
  if(find(t, ifd0, 50728, &e) && e.count >= 3) for(int i = 0; i < 3; i++) img->neutral[i] = (float)value(t, e, i);
  bool have2 = false;
  if(find(t, ifd0, 50722, &e) && e.count >= 9) { for(int i = 0; i < 9; i++) img->color_matrix[i] = (float)value(t, e, i); have2 = true; }
  else if(find(t, ifd0, 50721, &e) && e.count >= 9) for(int i = 0; i < 9; i++) img->color_matrix[i] = (float)value(t, e, i);
  if(find(t, ifd0, have2 ? 50779 : 50778, &e)) img->illuminant = (int)value(t, e, 0);
  if(find(t, ifd0, 271, &e)) { for(uint32_t i = 0; i < e.count && i < 31; i++) img->make[i] = (char)value(t, e, i); }
  if(find(t, ifd0, 272, &e)) { for(uint32_t i = 0; i < e.count && i < 31; i++) img->model[i] = (char)value(t, e, i); }
  if(find(t, ifd0, 274, &e)) img->orientation = (uint32_t)value(t, e, 0);
  if(find(t, ifd0, 34855, &e)) img->iso = (float)value(t, e, 0);
  else if(find(t, ifd0, 34665, &e)) { const size_t exif = (size_t)value(t, e, 0); entry_t i2; if(find(t, exif, 34855, &i2)) img->iso = (float)value(t, i2, 0); }
  return 0;
}

// ------------------------------------------------------------------------------------------------
namespace {
inline int cfa_at(const dng_image_t *img, uint32_t row, uint32_t col)
{ return img->cfa[(row % img->cfa_dim) * img->cfa_dim + (col % img->cfa_dim)]; }

void mul3(float *dst, const float *a, const float *b)
{ for(int k = 0; k < 3; k++) for(int i = 0; i < 3; i++) { float s = 0.0f; for(int j = 0; j < 3; j++) s += a[3*k+j] * b[3*j+i]; dst[3*k+i] = s; } }

int inv3(float *dst, const float *m)
{
  const float c00 = m[4]*m[8] - m[5]*m[7], c01 = m[5]*m[6] - m[3]*m[8], c02 = m[3]*m[7] - m[4]*m[6];
  const float det = m[0]*c00 + m[1]*c01 + m[2]*c02;
  if(det > -1e-7f && det < 1e-7f) return 1;
  const float id = 1.0f / det;
  dst[0] = id * c00; dst[1] = id * (m[2]*m[7] - m[1]*m[8]); dst[2] = id * (m[1]*m[5] - m[2]*m[4]);
  dst[3] = id * c01; dst[4] = id * (m[0]*m[8] - m[2]*m[6]); dst[5] = id * (m[2]*m[3] - m[0]*m[5]);
  dst[6] = id * c02; dst[7] = id * (m[1]*m[6] - m[0]*m[7]); dst[8] = id * (m[0]*m[4] - m[1]*m[3]);
  return 0;
}

// CAT16 adaptation of XYZ from the calibration illuminant to D65 (values of src/pipe/modules/matrices.h:37-42),
// keyed by EXIF LightSource
const float *cat16_to_d65(int light_source)
{
  static const float a[9]   = { 9.50674182e-01f, -1.87430902e-01f,  2.62831155e-01f, -2.56724729e-02f, 1.03231456e+00f, -1.15608371e-02f, -2.74089665e-03f,  9.09809774e-02f, 2.81290019e+00f };
  static const float d50[9] = { 9.89482020e-01f, -3.99715078e-02f,  4.40103520e-02f, -5.39737108e-03f, 1.00665157e+00f, -1.75423920e-03f, -4.03659609e-04f,  1.50625995e-02f, 1.30181385e+00f };
  static const float d55[9] = { 9.93809109e-01f, -2.36805157e-02f,  2.51833773e-02f, -3.19186083e-03f, 1.00396313e+00f, -9.86707812e-04f, -2.25775304e-04f,  8.60286147e-03f, 1.17257778e+00f };
  static const float d75[9] = { 1.00410545e+00f,  1.61206883e-02f, -1.57697661e-02f,  2.16403273e-03f, 9.97220529e-01f,  5.90558927e-04f,  1.33059688e-04f, -5.36122439e-03f, 8.92084722e-01f };
  switch(light_source)
  {
    case 17: case 3: return a;   // standard light A, tungsten
    case 23: return d50;
    case 20: return d55;
    case 22: return d75;
    default: return 0;           // 21 = D65, and everything the reference's loader does not adapt either
  }
}
} // namespace

int dng_raw_params(const dng_image_t *img, vkb_raw_params_t *p, uint32_t *pox, uint32_t *poy)
{
  memset(p, 0, sizeof(*p));
  uint32_t ox = 0, oy = 0;
  // lib.rs:206-256: move the window so that it starts on the canonical cfa phase
  if(img->cfa_dim == 6)
  {
    p->filters = 9;
    for(uint32_t i = 0; i < 6; i++) if(cfa_at(img, 0, i) == 1) { ox = i; break; }   // first green of row 0
    if(cfa_at(img, 0, ox + 1) != 1 && cfa_at(img, 0, ox + 2) != 1) { oy = 2; ox = (ox + 2) % 3; } // centre of the green x
    if(cfa_at(img, oy + 1, ox) == 1) oy++;                                           // two greens stacked
    if(cfa_at(img, oy, ox + 1) == 1) { if(ox >= 2) ox -= 2; else ox++; }              // two greens side by side
    if(cfa_at(img, oy, ox + 1) == 2) { if(ox < oy) ox += 3; else oy += 3; }           // red/blue swapped: other half block
  }
  else
  {
    p->filters = 0x49494949u; // lib.rs:203: "not 0 and not 9" means bayer, rggb after the shift
    if(cfa_at(img, 0, 0) == 1) { if(cfa_at(img, 0, 1) == 0) ox = 1; if(cfa_at(img, 0, 1) == 2) oy = 1; }
    else if(cfa_at(img, 0, 0) == 2) { ox = 1; oy = 1; }
  }
  if(ox >= img->width || oy >= img->height) return 1;
  const uint32_t block = img->cfa_dim == 6 ? 3 : 2, big = img->cfa_dim;
  // lib.rs:183-270: x y X Y from the active area; only X Y move with the window, x y round up to the full pattern
  uint32_t b[4] = { img->active[1], img->active[0], img->active[3], img->active[2] };
  b[2] -= ox < b[2] ? ox : b[2]; b[3] -= oy < b[3] ? oy : b[3];
  b[0] = ((b[0] + big - 1) / big) * big; b[1] = ((b[1] + big - 1) / big) * big;
  b[2] = (b[2] / block) * block;         b[3] = (b[3] / block) * block;
  for(int k = 0; k < 4; k++) p->crop_aabb[k] = b[k];
  p->width  = ((img->width  - ox) / block) * block;
  p->height = ((img->height - oy) / block) * block;
  // levels are given in cfa pattern order of the stored file; after the shift index k of the 2x2 block moves too
  for(int k = 0; k < 4; k++)
  {
    const uint32_t r = ((k >> 1) + oy) & 1, c = ((k & 1) + ox) & 1;
    p->black[k] = img->black[2 * r + c];
    p->white[k] = img->white;
  }
  // i-raw/main.c:215-219: wb = 1/neutral, normalised to green
  for(int k = 0; k < 3; k++) p->whitebalance[k] = img->neutral[1] / (img->neutral[k] > 0.0f ? img->neutral[k] : 1.0f);
  p->whitebalance[3] = 1.0f; // the reference carries rawler's 4th coefficient (unused by the graph for 3 colour cfas)
  // i-raw/main.c:228-256
  float xyz_to_cam[9], tmp[9], cam_to_xyz[9];
  memcpy(xyz_to_cam, img->color_matrix, sizeof(xyz_to_cam));
  if(const float *M = cat16_to_d65(img->illuminant)) { memcpy(tmp, xyz_to_cam, sizeof(tmp)); mul3(xyz_to_cam, tmp, M); }
  if(inv3(cam_to_xyz, xyz_to_cam)) return 2;
  static const float xyz_to_rec2020[9] = { 1.71665119f, -0.35567078f, -0.25336628f, -0.66668435f, 1.61648124f, 0.01576855f, 0.01763986f, -0.04277061f, 0.94210312f };
  mul3(p->cam_to_rec2020, xyz_to_rec2020, cam_to_xyz);
  p->orientation = img->orientation; // exif value as is (lib.rs:81), crop/main.c:190-193 interprets 3 / 6 / 8
  if(pox) *pox = ox;
  if(poy) *poy = oy;
  return 0;
}

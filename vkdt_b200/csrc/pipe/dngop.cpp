#include "dngop.h"
#include <string.h>

namespace {
struct rd_t
{
  const uint8_t *p, *end; bool ok = true;
  uint32_t u32() { if(end - p < 4) { ok = false; return 0; } const uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; p += 4; return v; }
  float f32() { const uint32_t v = u32(); float f; memcpy(&f, &v, 4); return f; }
  double f64() { const uint64_t hi = u32(), lo = u32(); const uint64_t v = (hi << 32) | lo; double d; memcpy(&d, &v, 8); return d; }
};
}

int dng_opcode_list_decode(const uint8_t *data, size_t len, dt_dng_opcode_list_t *out)
{
  out->ops.clear(); out->gain_maps.clear();
  if(!data || len < 4) return 1;
  rd_t r{ data, data + len };
  const uint32_t count = r.u32();
  if(count == 0) return 1;
  { // :319-329: every opcode's declared size has to fit, and the list has to end where the tag ends
    const uint8_t *p = data + 4;
    for(uint32_t i = 0; i < count; i++)
    {
      if((size_t)(data + len - p) < 16) return 1;
      p += 12;
      const uint32_t sz = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
      p += 4;
      if((size_t)(data + len - p) < sz) return 1;
      p += sz;
    }
    if(p != data + len) return 1;
  }
  for(uint32_t i = 0; i < count; i++)
  { // decode_opcode :261-297
    dt_dng_opcode_t op;
    op.id = r.u32();
    r.u32();                                   // minimum dng version
    const uint32_t flags = r.u32(), sz = r.u32();
    op.optional = (flags & 1) > 0; op.preview_skip = (flags & 2) > 0; op.gain_map = -1;
    const uint8_t *body = r.p;
    if(op.id == 9)
    { // decode_gain_map :113-138
      dt_dng_gain_map_t gm;
      gm.top = r.u32(); gm.left = r.u32(); gm.bottom = r.u32(); gm.right = r.u32();
      gm.plane = r.u32(); gm.planes = r.u32(); gm.row_pitch = r.u32(); gm.col_pitch = r.u32();
      gm.map_points_v = r.u32(); gm.map_points_h = r.u32();
      gm.map_spacing_v = r.f64(); gm.map_spacing_h = r.f64(); gm.map_origin_v = r.f64(); gm.map_origin_h = r.f64();
      gm.map_planes = r.u32();
      const uint64_t cnt = (uint64_t)gm.map_points_h * gm.map_points_v * gm.map_planes;
      if(!r.ok || sz < 76 || (uint64_t)sz != 76 + cnt * 4) { out->ops.clear(); out->gain_maps.clear(); return 1; }
      gm.map_gain.resize((size_t)cnt);
      for(uint64_t k = 0; k < cnt; k++) gm.map_gain[(size_t)k] = r.f32();
      op.gain_map = (int)out->gain_maps.size();
      out->gain_maps.push_back(gm);
    }
    r.p = body + sz;
    out->ops.push_back(op);
  }
  return 0;
}

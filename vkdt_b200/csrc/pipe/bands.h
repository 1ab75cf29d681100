// band split of large stills over several GPUs (BASELINE.json config 5): row interval sets and the
// per kernel footprints the executor plans its halo exchange with.  see executor.cpp (build_bands) and DESIGN.md section 6.
#pragma once
#include <vector>
#include <algorithm>
#include <utility>

// sorted, disjoint, non-adjacent half open row intervals [a, b)
struct rows_t
{
  std::vector<std::pair<int,int>> v;
  bool empty() const { return v.empty(); }
  long count() const { long n = 0; for(auto &p : v) n += p.second - p.first; return n; }
  void add(int a, int b)
  {
    if(b <= a) return;
    std::vector<std::pair<int,int>> o;
    bool placed = false;
    for(auto &p : v)
    {
      if(p.second < a) o.push_back(p);
      else if(p.first > b) { if(!placed) { o.push_back({a, b}); placed = true; } o.push_back(p); }
      else { a = std::min(a, p.first); b = std::max(b, p.second); }
    }
    if(!placed) o.push_back({a, b});
    v.swap(o);
  }
  void add(const rows_t &r) { for(auto &p : r.v) add(p.first, p.second); }
  rows_t clipped(int lo, int hi) const
  {
    rows_t r;
    for(auto &p : v) r.add(std::max(p.first, lo), std::min(p.second, hi));
    return r;
  }
  rows_t minus(const rows_t &o) const
  {
    rows_t r;
    for(auto p : v)
    {
      int a = p.first;
      for(auto &q : o.v)
      {
        if(q.second <= a || q.first >= p.second) continue;
        if(q.first > a) r.add(a, q.first);
        a = std::max(a, q.second);
        if(a >= p.second) break;
      }
      if(a < p.second) r.add(a, p.second);
    }
    return r;
  }
  rows_t intersect(const rows_t &o) const
  {
    rows_t r;
    for(auto &p : v) for(auto &q : o.v) r.add(std::max(p.first, q.first), std::min(p.second, q.second));
    return r;
  }
};

// denoise's quadrant swizzle of one axis (down.comp:103-104): row y of a level is stored at y/2 + (y&1)*(h+1)/2
static inline int band_swz(int y, int h) { return y / 2 + ((y & 1) * (h + 1)) / 2; }
static inline rows_t band_swz_rows(const rows_t &r, int h)
{
  rows_t o;
  for(auto &p : r.v) for(int y = p.first; y < p.second; y++) { const int s = band_swz(y, h); if(s >= 0 && s < h) o.add(s, s + 1); }
  return o;
}

// ORACLE — test infrastructure.  enough of the GLSL language as C++17 types and functions that the reference's compute
// shaders (src/pipe/modules/<module>/<kernel>.comp, preprocessed by comp2cpp.py where they lie under /root/reference, never
// copied) compile with g++ and run on the CPU, one invocation per pixel.  this pins the CPU restatement (oracle/o_*.c)
// against the reference's own shader SOURCE: the arithmetic, its order and its constants are the shader's; what is
// restated here is only the language runtime: vector types with swizzles, the built-in functions (libm, fp32), and the
// image model (texelFetch, imageStore with f16 rounding on f16 images, linear sampling with mirrored repeat).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#include <barrier>
#include <thread>

namespace glsl {
typedef unsigned int uint;
typedef _Float16 float16_t;   // GL_EXT_shader_explicit_arithmetic_types_float16: storage type of rcd_fill's shared memory tiles

// ---- swizzles: a proxy that lives in a union with the vector's storage ------------------------------------------------
template<class V, int P, int... I> struct swz
{ // V: the vector type this reads as, P: components of the parent
  float d[P];
  operator V() const { V v; int k = 0; ((v.d[k++] = d[I]), ...); return v; }
  swz &operator=(const V &v) { int k = 0; ((d[I] = v.d[k++]), ...); return *this; }
  swz &operator=(const swz &o) { return *this = V(o); }
  swz &operator+=(const V &v) { return *this = V(*this) + v; }
  swz &operator-=(const V &v) { return *this = V(*this) - v; }
  swz &operator*=(const V &v) { return *this = V(*this) * v; }
  swz &operator*=(float s)    { return *this = V(*this) * s; }
  swz &operator/=(float s)    { return *this = V(*this) / s; }
  float operator[](int i) const { const int idx[] = { I... }; return d[idx[i]]; }
};

#define GLSL_VEC_OPS(V, N) \
  float &operator[](int i) { return d[i]; } \
  float operator[](int i) const { return d[i]; } \
  friend V operator+(const V &a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; } \
  friend V operator-(const V &a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; } \
  friend V operator*(const V &a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] * b.d[i]; return r; } \
  friend V operator/(const V &a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] / b.d[i]; return r; } \
  friend V operator+(const V &a, float b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] + b; return r; } \
  friend V operator-(const V &a, float b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] - b; return r; } \
  friend V operator*(const V &a, float b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] * b; return r; } \
  friend V operator/(const V &a, float b) { V r; for(int i = 0; i < N; i++) r.d[i] = a.d[i] / b; return r; } \
  friend V operator+(float a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a + b.d[i]; return r; } \
  friend V operator-(float a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a - b.d[i]; return r; } \
  friend V operator*(float a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a * b.d[i]; return r; } \
  friend V operator/(float a, const V &b) { V r; for(int i = 0; i < N; i++) r.d[i] = a / b.d[i]; return r; } \
  friend V operator-(const V &a) { V r; for(int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; } \
  V &operator+=(const V &b) { for(int i = 0; i < N; i++) d[i] += b.d[i]; return *this; } \
  V &operator-=(const V &b) { for(int i = 0; i < N; i++) d[i] -= b.d[i]; return *this; } \
  V &operator*=(const V &b) { for(int i = 0; i < N; i++) d[i] *= b.d[i]; return *this; } \
  V &operator/=(const V &b) { for(int i = 0; i < N; i++) d[i] /= b.d[i]; return *this; } \
  V &operator+=(float b) { for(int i = 0; i < N; i++) d[i] += b; return *this; } \
  V &operator-=(float b) { for(int i = 0; i < N; i++) d[i] -= b; return *this; } \
  V &operator*=(float b) { for(int i = 0; i < N; i++) d[i] *= b; return *this; } \
  V &operator/=(float b) { for(int i = 0; i < N; i++) d[i] /= b; return *this; }

struct vec2; struct vec3; struct vec4; struct ivec2; struct uvec3;
template<class V, int P, int A, int B> struct iswz2 { int d[P]; operator V() const { return V(d[A], d[B]); } };

struct vec2
{
  union { float d[2]; struct { float x, y; }; struct { float r, g; };
    swz<vec2, 2, 0, 1> xy, rg; swz<vec2, 2, 1, 0> yx; swz<vec2, 2, 0, 0> xx; swz<vec2, 2, 1, 1> yy; };
  vec2() : d{0, 0} {}
  explicit vec2(float s) : d{s, s} {}
  vec2(float a, float b) : d{a, b} {}
  vec2(const ivec2 &i);
  vec2(const vec2 &o) { d[0] = o.d[0]; d[1] = o.d[1]; }
  vec2 &operator=(const vec2 &o) { d[0] = o.d[0]; d[1] = o.d[1]; return *this; }
  GLSL_VEC_OPS(vec2, 2)
};
struct vec3
{
  union { float d[3]; struct { float x, y, z; }; struct { float r, g, b; };
    swz<vec3, 3, 0, 1, 2> xyz, rgb; swz<vec3, 3, 2, 1, 0> zyx, bgr; swz<vec3, 3, 0, 0, 0> xxx, rrr; swz<vec3, 3, 1, 1, 1> yyy, ggg; swz<vec3, 3, 2, 2, 2> zzz, bbb;
    swz<vec3, 3, 1, 2, 0> yzx, gbr; swz<vec3, 3, 2, 0, 1> zxy, brg; swz<vec4, 3, 1, 2, 1, 0> gbgr; swz<vec4, 3, 0, 1, 2, 2> rgbb; swz<vec4, 3, 0, 1, 1, 2> rggb; swz<vec2, 3, 2, 0> zx, br; swz<vec2, 3, 1, 0> yx, gr; swz<vec2, 3, 2, 1> zy, bg;
    swz<vec2, 3, 0, 1> xy, rg; swz<vec2, 3, 1, 2> yz, gb; swz<vec2, 3, 0, 2> xz, rb; };
  vec3() : d{0, 0, 0} {}
  explicit vec3(float s) : d{s, s, s} {}
  vec3(float a, float b, float c) : d{a, b, c} {}
  vec3(const vec2 &a, float c) : d{a.d[0], a.d[1], c} {}
  vec3(float a, const vec2 &b) : d{a, b.d[0], b.d[1]} {}
  vec3(const vec3 &o) { for(int i = 0; i < 3; i++) d[i] = o.d[i]; }
  vec3 &operator=(const vec3 &o) { for(int i = 0; i < 3; i++) d[i] = o.d[i]; return *this; }
  GLSL_VEC_OPS(vec3, 3)
};
struct vec4
{
  union { float d[4]; struct { float x, y, z, w; }; struct { float r, g, b, a; };
    swz<vec4, 4, 0, 1, 2, 3> xyzw, rgba;
    swz<vec3, 4, 0, 1, 2> xyz, rgb; swz<vec3, 4, 2, 1, 0> zyx, bgr; swz<vec3, 4, 1, 1, 1> yyy, ggg; swz<vec4, 4, 1, 2, 1, 0> gbgr; swz<vec4, 4, 0, 1, 1, 2> rggb; swz<vec4, 4, 3, 2, 1, 0> wzyx, abgr;
    swz<vec3, 4, 1, 2, 3> yzw, gba; swz<vec3, 4, 0, 0, 0> xxx, rrr;
    swz<vec2, 4, 0, 1> xy, rg; swz<vec2, 4, 2, 0> zx, br; swz<vec2, 4, 1, 3> yw, ga; swz<vec2, 4, 2, 3> zw, ba; swz<vec2, 4, 1, 2> yz, gb; swz<vec2, 4, 0, 2> xz, rb; swz<vec2, 4, 0, 3> xw, ra; };
  vec4() : d{0, 0, 0, 0} {}
  explicit vec4(float s) : d{s, s, s, s} {}
  vec4(float a, float b, float c, float e) : d{a, b, c, e} {}
  vec4(const vec3 &a, float e) : d{a.d[0], a.d[1], a.d[2], e} {}
  vec4(float a, const vec3 &b) : d{a, b.d[0], b.d[1], b.d[2]} {}
  vec4(const vec2 &a, const vec2 &b) : d{a.d[0], a.d[1], b.d[0], b.d[1]} {}
  vec4(const vec2 &a, float c, float e) : d{a.d[0], a.d[1], c, e} {}
  vec4(const vec4 &o) { for(int i = 0; i < 4; i++) d[i] = o.d[i]; }
  vec4 &operator=(const vec4 &o) { for(int i = 0; i < 4; i++) d[i] = o.d[i]; return *this; }
  GLSL_VEC_OPS(vec4, 4)
};
struct uvec3 { union { uint d[3]; struct { uint x, y, z; }; iswz2<ivec2, 3, 0, 1> xy; }; uvec3(uint a = 0, uint b = 0, uint c = 0) : d{a, b, c} {} };
struct bvec2 { bool d[2]; };
struct bvec3 { union { bool d[3]; struct { bool x, y, z; }; }; bvec3() : d{false, false, false} {} explicit bvec3(bool b) : d{b, b, b} {} bvec3(bool a, bool b, bool c) : d{a, b, c} {} };
struct bvec4 { bool d[4]; };
struct ivec2
{
  union { int d[2]; struct { int x, y; }; struct { int r, g; }; iswz2<ivec2, 2, 0, 1> xy; iswz2<ivec2, 2, 1, 0> yx; };
  ivec2() : d{0, 0} {}
  explicit ivec2(int s) : d{s, s} {}
  ivec2(int a, int b) : d{a, b} {}
  explicit ivec2(const uvec3 &u) : d{(int)u.x, (int)u.y} {}
  explicit ivec2(const vec2 &v) : d{(int)v.d[0], (int)v.d[1]} {}   // conversion truncates towards zero
  int &operator[](int i) { return d[i]; }
  int operator[](int i) const { return d[i]; }
  ivec2 &operator+=(const ivec2 &b) { x += b.x; y += b.y; return *this; }
  ivec2 &operator-=(const ivec2 &b) { x -= b.x; y -= b.y; return *this; }
  ivec2 &operator*=(int b) { x *= b; y *= b; return *this; }
  ivec2 &operator/=(int b) { x /= b; y /= b; return *this; }
  friend ivec2 operator+(const ivec2 &a, const ivec2 &b) { return ivec2(a.x + b.x, a.y + b.y); }
  friend ivec2 operator-(const ivec2 &a, const ivec2 &b) { return ivec2(a.x - b.x, a.y - b.y); }
  friend ivec2 operator*(const ivec2 &a, const ivec2 &b) { return ivec2(a.x * b.x, a.y * b.y); }
  friend ivec2 operator/(const ivec2 &a, const ivec2 &b) { return ivec2(a.x / b.x, a.y / b.y); }
  friend ivec2 operator*(int a, const ivec2 &b) { return ivec2(a * b.x, a * b.y); }
  friend ivec2 operator*(const ivec2 &a, int b) { return ivec2(a.x * b, a.y * b); }
  friend ivec2 operator/(const ivec2 &a, int b) { return ivec2(a.x / b, a.y / b); }
  friend ivec2 operator+(const ivec2 &a, int b) { return ivec2(a.x + b, a.y + b); }
  friend ivec2 operator-(const ivec2 &a, int b) { return ivec2(a.x - b, a.y - b); }
  friend ivec2 operator&(const ivec2 &a, int b) { return ivec2(a.x & b, a.y & b); }
  friend ivec2 operator%(const ivec2 &a, int b) { return ivec2(a.x % b, a.y % b); }
  // int vector with a float scalar or vector: the int side converts (GLSL implicit conversion)
  friend vec2 operator*(const ivec2 &a, float b) { return vec2((float)a.x * b, (float)a.y * b); }
  friend vec2 operator*(float a, const ivec2 &b) { return vec2(a * (float)b.x, a * (float)b.y); }
  friend vec2 operator/(float a, const ivec2 &b) { return vec2(a / (float)b.x, a / (float)b.y); }
  friend vec2 operator/(const ivec2 &a, float b) { return vec2((float)a.x / b, (float)a.y / b); }
  friend vec2 operator+(const ivec2 &a, float b) { return vec2((float)a.x + b, (float)a.y + b); }
  friend vec2 operator+(float a, const ivec2 &b) { return vec2(a + (float)b.x, a + (float)b.y); }
  friend vec2 operator-(const ivec2 &a, float b) { return vec2((float)a.x - b, (float)a.y - b); }
  friend vec2 operator-(float a, const ivec2 &b) { return vec2(a - (float)b.x, a - (float)b.y); }
  friend vec2 operator*(const vec2 &a, const ivec2 &b) { return vec2(a.d[0] * (float)b.x, a.d[1] * (float)b.y); }
  friend vec2 operator/(const ivec2 &a, const vec2 &b) { return vec2((float)a.x / b.d[0], (float)a.y / b.d[1]); }
  friend vec2 operator-(const vec2 &a, const ivec2 &b) { return vec2(a.d[0] - (float)b.x, a.d[1] - (float)b.y); }
  friend vec2 operator+(const ivec2 &a, const vec2 &b) { return vec2((float)a.x + b.d[0], (float)a.y + b.d[1]); }
  friend vec2 operator+(const vec2 &a, const ivec2 &b) { return vec2(a.d[0] + (float)b.x, a.d[1] + (float)b.y); }
  friend vec2 operator-(const ivec2 &a, const vec2 &b) { return vec2((float)a.x - b.d[0], (float)a.y - b.d[1]); }
  friend vec2 operator*(const ivec2 &a, const vec2 &b) { return vec2((float)a.x * b.d[0], (float)a.y * b.d[1]); }
  friend vec2 operator/(const vec2 &a, const ivec2 &b) { return vec2(a.d[0] / (float)b.x, a.d[1] / (float)b.y); }
};
inline vec2::vec2(const ivec2 &i) : d{(float)i.x, (float)i.y} {}

// ---- built-ins (fp32, libm) -----------------------------------------------------------------------------------------
#define GLSL_MAP1(F, EXPR) \
  inline float F(float a) { return EXPR; } \
  inline vec2 F(const vec2 &v) { vec2 r; for(int i = 0; i < 2; i++) { const float a = v.d[i]; r.d[i] = EXPR; } return r; } \
  inline vec3 F(const vec3 &v) { vec3 r; for(int i = 0; i < 3; i++) { const float a = v.d[i]; r.d[i] = EXPR; } return r; } \
  inline vec4 F(const vec4 &v) { vec4 r; for(int i = 0; i < 4; i++) { const float a = v.d[i]; r.d[i] = EXPR; } return r; }
GLSL_MAP1(exp, ::expf(a)) GLSL_MAP1(exp2, ::exp2f(a)) GLSL_MAP1(log, ::logf(a)) GLSL_MAP1(log2, ::log2f(a)) GLSL_MAP1(sqrt, ::sqrtf(a))
GLSL_MAP1(abs, ::fabsf(a)) GLSL_MAP1(floor, ::floorf(a)) GLSL_MAP1(ceil, ::ceilf(a)) GLSL_MAP1(fract, a - ::floorf(a)) GLSL_MAP1(sin, ::sinf(a))
GLSL_MAP1(cos, ::cosf(a)) GLSL_MAP1(sign, (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f)) GLSL_MAP1(inversesqrt, 1.0f / ::sqrtf(a))
GLSL_MAP1(isnan_f, std::isnan(a) ? 1.0f : 0.0f)
#define GLSL_MAP2(F, EXPR) \
  inline float F(float a, float b) { return EXPR; } \
  inline vec2 F(const vec2 &u, const vec2 &v) { vec2 r; for(int i = 0; i < 2; i++) { const float a = u.d[i], b = v.d[i]; r.d[i] = EXPR; } return r; } \
  inline vec3 F(const vec3 &u, const vec3 &v) { vec3 r; for(int i = 0; i < 3; i++) { const float a = u.d[i], b = v.d[i]; r.d[i] = EXPR; } return r; } \
  inline vec4 F(const vec4 &u, const vec4 &v) { vec4 r; for(int i = 0; i < 4; i++) { const float a = u.d[i], b = v.d[i]; r.d[i] = EXPR; } return r; } \
  inline vec2 F(const vec2 &u, float b) { vec2 r; for(int i = 0; i < 2; i++) { const float a = u.d[i]; r.d[i] = EXPR; } return r; } \
  inline vec3 F(const vec3 &u, float b) { vec3 r; for(int i = 0; i < 3; i++) { const float a = u.d[i]; r.d[i] = EXPR; } return r; } \
  inline vec4 F(const vec4 &u, float b) { vec4 r; for(int i = 0; i < 4; i++) { const float a = u.d[i]; r.d[i] = EXPR; } return r; }
// min / max: the operand that is not NaN wins, like the GPU's (and like the oracle's o_min / o_max)
GLSL_MAP2(max, (a < b || std::isnan(a)) ? b : a) GLSL_MAP2(min, (b < a || std::isnan(a)) ? b : a)
GLSL_MAP2(pow, ::powf(a, b)) GLSL_MAP2(mod, a - b * ::floorf(a / b)) GLSL_MAP2(step, b < a ? 0.0f : 1.0f) GLSL_MAP2(atan, ::atan2f(a, b))
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline float max(int a, float b) { return max((float)a, b); }
inline float max(float a, int b) { return max(a, (float)b); }
inline float min(int a, float b) { return min((float)a, b); }
inline float min(float a, int b) { return min(a, (float)b); }
inline float max(float a, double b) { return max(a, (float)b); }
inline float max(double a, float b) { return max((float)a, b); }
inline float min(float a, double b) { return min(a, (float)b); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int   clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline uint  clamp(uint x, int lo, int hi) { return x < (uint)lo ? (uint)lo : (x > (uint)hi ? (uint)hi : x); }
inline float clamp(float x, int lo, float hi) { return clamp(x, (float)lo, hi); }
inline float clamp(float x, float lo, int hi) { return clamp(x, lo, (float)hi); }
inline float clamp(float x, int lo, int hi) { return clamp(x, (float)lo, (float)hi); }
inline vec2  clamp(const vec2 &x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3  clamp(const vec3 &x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec4  clamp(const vec4 &x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec2  clamp(const vec2 &x, const vec2 &lo, const vec2 &hi) { return min(max(x, lo), hi); }
inline vec3  clamp(const vec3 &x, const vec3 &lo, const vec3 &hi) { return min(max(x, lo), hi); }
inline vec4  clamp(const vec4 &x, const vec4 &lo, const vec4 &hi) { return min(max(x, lo), hi); }
inline ivec2 clamp(const ivec2 &x, const ivec2 &lo, const ivec2 &hi) { return ivec2(clamp(x.x, lo.x, hi.x), clamp(x.y, lo.y, hi.y)); }
inline ivec2 max(const ivec2 &a, const ivec2 &b) { return ivec2(max(a.x, b.x), max(a.y, b.y)); }
inline ivec2 min(const ivec2 &a, const ivec2 &b) { return ivec2(min(a.x, b.x), min(a.y, b.y)); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float mix(float a, float b, bool t) { return t ? b : a; }
inline vec2  mix(const vec2 &a, const vec2 &b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  mix(const vec3 &a, const vec3 &b, float t) { return a * (1.0f - t) + b * t; }
inline vec4  mix(const vec4 &a, const vec4 &b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  mix(const vec3 &a, const vec3 &b, const vec3 &t) { return a * (1.0f - t) + b * t; }
inline vec4  mix(const vec4 &a, const vec4 &b, const vec4 &t) { return a * (1.0f - t) + b * t; }
inline float smoothstep(float e0, float e1, float x) { const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
inline vec3  smoothstep(float e0, float e1, const vec3 &x) { return vec3(smoothstep(e0, e1, x.d[0]), smoothstep(e0, e1, x.d[1]), smoothstep(e0, e1, x.d[2])); }
inline float dot(const vec2 &a, const vec2 &b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1]; }
inline float dot(const vec3 &a, const vec3 &b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1] + a.d[2] * b.d[2]; }
inline float dot(const vec4 &a, const vec4 &b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1] + a.d[2] * b.d[2] + a.d[3] * b.d[3]; }
inline float length(const vec2 &a) { return ::sqrtf(dot(a, a)); }
inline float length(const vec3 &a) { return ::sqrtf(dot(a, a)); }
inline vec3  normalize(const vec3 &a) { return a / length(a); }
inline vec3  cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline bvec2 greaterThanEqual(const ivec2 &a, const ivec2 &b) { return bvec2{{a.x >= b.x, a.y >= b.y}}; }
inline bvec2 greaterThan(const ivec2 &a, const ivec2 &b) { return bvec2{{a.x > b.x, a.y > b.y}}; }
inline bvec2 lessThan(const ivec2 &a, const ivec2 &b) { return bvec2{{a.x < b.x, a.y < b.y}}; }
#define GLSL_CMP(F, OP) \
  inline bvec3 F(const vec3 &a, const vec3 &b) { return bvec3(a.d[0] OP b.d[0], a.d[1] OP b.d[1], a.d[2] OP b.d[2]); } \
  inline bvec4 F(const vec4 &a, const vec4 &b) { return bvec4{{a.d[0] OP b.d[0], a.d[1] OP b.d[1], a.d[2] OP b.d[2], a.d[3] OP b.d[3]}}; } \
  inline bvec2 F(const vec2 &a, const vec2 &b) { return bvec2{{a.d[0] OP b.d[0], a.d[1] OP b.d[1]}}; }
GLSL_CMP(lessThan, <) GLSL_CMP(lessThanEqual, <=) GLSL_CMP(greaterThan, >) GLSL_CMP(greaterThanEqual, >=) GLSL_CMP(equal, ==)
inline bool any(const bvec3 &b) { return b.d[0] || b.d[1] || b.d[2]; }
inline bool all(const bvec3 &b) { return b.d[0] && b.d[1] && b.d[2]; }
inline bool any(const bvec4 &b) { return b.d[0] || b.d[1] || b.d[2] || b.d[3]; }
inline bool all(const bvec4 &b) { return b.d[0] && b.d[1] && b.d[2] && b.d[3]; }
inline vec3 mix(const vec3 &a, const vec3 &b, const bvec3 &t) { return vec3(t.d[0] ? b.d[0] : a.d[0], t.d[1] ? b.d[1] : a.d[1], t.d[2] ? b.d[2] : a.d[2]); }
inline bool any(const bvec2 &b) { return b.d[0] || b.d[1]; }
inline bool all(const bvec2 &b) { return b.d[0] && b.d[1]; }
inline vec2 unpackHalf2x16(uint v) { uint16_t lo = (uint16_t)(v & 0xffffu), hi = (uint16_t)(v >> 16); _Float16 a, b; memcpy(&a, &lo, 2); memcpy(&b, &hi, 2); return vec2((float)a, (float)b); }
inline bool isnan(float a) { return std::isnan(a); }
inline bool isinf(float a) { return std::isinf(a); }

struct mat2
{ // column major
  vec2 c[2];
  mat2() {}
  mat2(float a0, float a1, float b0, float b1) { c[0] = vec2(a0, a1); c[1] = vec2(b0, b1); }
  mat2(const vec2 &a, const vec2 &b) { c[0] = a; c[1] = b; }
  vec2 &operator[](int i) { return c[i]; }
  friend vec2 operator*(const mat2 &m, const vec2 &v) { return m.c[0] * v.d[0] + m.c[1] * v.d[1]; }
  friend vec2 operator*(const vec2 &v, const mat2 &m) { return vec2(dot(v, m.c[0]), dot(v, m.c[1])); }
  friend mat2 operator+(const mat2 &a, const mat2 &b) { return mat2(a.c[0] + b.c[0], a.c[1] + b.c[1]); }
  friend mat2 operator-(const mat2 &a, const mat2 &b) { return mat2(a.c[0] - b.c[0], a.c[1] - b.c[1]); }
  friend mat2 operator*(const mat2 &a, float b) { return mat2(a.c[0] * b, a.c[1] * b); }
  friend mat2 operator*(float a, const mat2 &b) { return mat2(a * b.c[0], a * b.c[1]); }
  friend mat2 operator/(const mat2 &a, float b) { return mat2(a.c[0] / b, a.c[1] / b); }
  friend mat2 operator*(const mat2 &a, const mat2 &b) { return mat2(a * b.c[0], a * b.c[1]); }
  mat2 &operator+=(const mat2 &b) { c[0] += b.c[0]; c[1] += b.c[1]; return *this; }
  mat2 &operator*=(float b) { c[0] *= b; c[1] *= b; return *this; }
  mat2 &operator/=(float b) { c[0] /= b; c[1] /= b; return *this; }
  explicit mat2(float s) { c[0] = vec2(s, 0); c[1] = vec2(0, s); }
};
inline float determinant(const mat2 &m) { return m.c[0].d[0] * m.c[1].d[1] - m.c[1].d[0] * m.c[0].d[1]; }
inline mat2 outerProduct(const vec2 &c, const vec2 &r) { return mat2(c * r.d[0], c * r.d[1]); }
inline mat2 transpose(const mat2 &m) { return mat2(vec2(m.c[0].d[0], m.c[1].d[0]), vec2(m.c[0].d[1], m.c[1].d[1])); }
struct ivec4 { union { int d[4]; struct { int x, y, z, w; }; iswz2<ivec2, 4, 0, 1> xy; iswz2<ivec2, 4, 2, 3> zw; }; ivec4() : d{0, 0, 0, 0} {} int operator[](int i) const { return d[i]; } };
struct uvec4 { union { uint d[4]; struct { uint x, y, z, w; }; }; uvec4() : d{0, 0, 0, 0} {} uint operator[](int i) const { return d[i]; } };
struct mat3
{ // column major: m[c] is a column, mat3(c0, c1, c2)
  vec3 c[3];
  mat3() {}
  explicit mat3(float s) { c[0] = vec3(s, 0, 0); c[1] = vec3(0, s, 0); c[2] = vec3(0, 0, s); }
  mat3(const vec3 &a, const vec3 &b, const vec3 &e) { c[0] = a; c[1] = b; c[2] = e; }
  mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) { c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2); }
  vec3 &operator[](int i) { return c[i]; }
  const vec3 &operator[](int i) const { return c[i]; }
  friend vec3 operator*(const mat3 &m, const vec3 &v) { return m.c[0] * v.d[0] + m.c[1] * v.d[1] + m.c[2] * v.d[2]; }
  friend vec3 operator*(const vec3 &v, const mat3 &m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }
  friend mat3 operator*(const mat3 &a, const mat3 &b) { return mat3(a * b.c[0], a * b.c[1], a * b.c[2]); }
};
inline mat3 transpose(const mat3 &m) { return mat3(vec3(m.c[0].x, m.c[1].x, m.c[2].x), vec3(m.c[0].y, m.c[1].y, m.c[2].y), vec3(m.c[0].z, m.c[1].z, m.c[2].z)); }

// ---- images ---------------------------------------------------------------------------------------------------------
// one struct for both views of a connector: sampler2D (texelFetch / texture / textureSize) and image2D (imageStore / imageSize).
// data: float, `chan` interleaved channels; f16: stores round to half precision (RNE) like a VK_FORMAT_R16*_SFLOAT image.
struct image_t { float *data; int wd, ht, chan, f16; };
typedef image_t sampler2D;
typedef image_t image2D;

inline float round_f16(float f) { return (float)(_Float16)f; }
inline vec4 image_fetch(const image_t &im, int x, int y)
{ // texelFetch out of range is undefined in Vulkan without robustness; with it, zero.  the oracle clamps, like NVIDIA does
  x = x < 0 ? 0 : (x >= im.wd ? im.wd - 1 : x);
  y = y < 0 ? 0 : (y >= im.ht ? im.ht - 1 : y);
  const float *p = im.data + ((size_t)y * im.wd + x) * im.chan;
  vec4 v(0.0f, 0.0f, 0.0f, 1.0f);
  for(int c = 0; c < im.chan && c < 4; c++) v.d[c] = p[c];
  return v;
}
inline vec4 texelFetch(const sampler2D &s, const ivec2 &p, int) { return image_fetch(s, p.x, p.y); }
inline ivec2 textureSize(const sampler2D &s, int) { return ivec2(s.wd, s.ht); }
inline ivec2 imageSize(const image2D &s) { return ivec2(s.wd, s.ht); }
inline void imageStore(image2D &im, const ivec2 &p, const vec4 &v)
{
  if(p.x < 0 || p.y < 0 || p.x >= im.wd || p.y >= im.ht) return;
  float *o = im.data + ((size_t)p.y * im.wd + p.x) * im.chan;
  for(int c = 0; c < im.chan && c < 4; c++) o[c] = im.f16 ? round_f16(v.d[c]) : v.d[c];
}
inline int mirror_repeat(int i, int n)
{ // VK_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT (qvk.c:596-611)
  const int period = 2 * n;
  int m = i % period; if(m < 0) m += period;
  return m < n ? m : period - 1 - m;
}
inline vec4 texture(const sampler2D &s, const vec2 &uv)
{ // linear filter, normalised coordinates; the ideal sampler of DESIGN.md §4: coordinates in double, exact float weights,
  // taps within 1/4096 of a texel centre snap to it
  const double fx = (double)uv.d[0] * s.wd - 0.5, fy = (double)uv.d[1] * s.ht - 0.5;
  double x0 = std::floor(fx), y0 = std::floor(fy);
  double wx = fx - x0, wy = fy - y0;
  const double eps = 1.0 / 4096.0;
  if(wx < eps) wx = 0.0; else if(wx > 1.0 - eps) { wx = 0.0; x0 += 1.0; }
  if(wy < eps) wy = 0.0; else if(wy > 1.0 - eps) { wy = 0.0; y0 += 1.0; }
  const int ix0 = mirror_repeat((int)x0, s.wd), ix1 = mirror_repeat((int)x0 + 1, s.wd);
  const int iy0 = mirror_repeat((int)y0, s.ht), iy1 = mirror_repeat((int)y0 + 1, s.ht);
  const float ax = (float)wx, ay = (float)wy;
  const vec4 t00 = image_fetch(s, ix0, iy0), t10 = image_fetch(s, ix1, iy0), t01 = image_fetch(s, ix0, iy1), t11 = image_fetch(s, ix1, iy1);
  return (t00 * (1.0f - ax) + t10 * ax) * (1.0f - ay) + (t01 * (1.0f - ax) + t11 * ax) * ay;
}
inline vec4 textureGather(const sampler2D &s, const vec2 &uv, int comp)
{ // the four texels a linear fetch at uv would blend: (i0,j1), (i1,j1), (i1,j0), (i0,j0), component comp of each
  const double fx = (double)uv.d[0] * s.wd - 0.5, fy = (double)uv.d[1] * s.ht - 0.5;
  // the shaders gather at texel corners (2*(ipos+.5)/size): fx, fy are integers + 0.5 up to fp32 noise in uv; snap like texture()
  const int i0 = mirror_repeat((int)std::floor(fx + 1.0 / 4096.0), s.wd), i1 = mirror_repeat((int)std::floor(fx + 1.0 / 4096.0) + 1, s.wd);
  const int j0 = mirror_repeat((int)std::floor(fy + 1.0 / 4096.0), s.ht), j1 = mirror_repeat((int)std::floor(fy + 1.0 / 4096.0) + 1, s.ht);
  return vec4(image_fetch(s, i0, j1).d[comp], image_fetch(s, i1, j1).d[comp], image_fetch(s, i1, j0).d[comp], image_fetch(s, i0, j0).d[comp]);
}
inline vec4 textureLod(const sampler2D &s, const vec2 &uv, float) { return texture(s, uv); }
// ---- workgroups: shaders that call barrier() run one thread per invocation of a workgroup -------------------------------------
inline thread_local std::barrier<> *current_barrier = 0;
inline void barrier() { if(current_barrier) current_barrier->arrive_and_wait(); }
inline void memoryBarrierShared() {}
template<class F> inline void run_workgroup(int lx, int ly, F &&invocation)
{ // invocation(local x, local y) sets the built-in ids and runs main(); a thread that returns early drops out of the barrier
  std::barrier<> bar(lx * ly);
  std::vector<std::thread> th;
  th.reserve((size_t)lx * ly);
  for(int y = 0; y < ly; y++) for(int x = 0; x < lx; x++)
    th.emplace_back([&bar, &invocation, x, y]() { current_barrier = &bar; invocation(x, y); bar.arrive_and_drop(); });
  for(std::thread &t : th) t.join();
}
} // namespace glsl

// TEST INFRASTRUCTURE ONLY — CPU oracle for the ElasticFusion hot path.
// Nothing under oracle/ is linked into, imported by or executed from the product
// (elasticfusion_b200/). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, and only as the checker or the
// timed baseline.
//
// Small fp32 vector helpers restating Core/Cuda/operators.cuh:58-85 of the reference
// (dot / cross / norm / normalized / mat33*vec) with the same operation order.
// Compiled with -ffp-contract=off so every +,-,*,/ and sqrt is a single IEEE op.
// The reference's `normalized` uses the GPU's approximate rsqrtf (operators.cuh:78-81);
// here it is 1/sqrtf (IEEE), so reference-GPU outputs agree to ~2 ulp, not bitwise.
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace efo {

struct f3 {
  float x, y, z;
};
struct f4 {
  float x, y, z, w;
};
struct m33 {
  f3 r[3];  // rows
};

static inline f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
static inline f3 operator-(const f3& a, const f3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 operator+(const f3& a, const f3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 operator*(const f3& a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline f3 cross(const f3& a, const f3& b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float dot(const f3& a, const f3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float norm(const f3& a) { return sqrtf(dot(a, a)); }
static inline f3 normalized(const f3& a) {
  const float rn = 1.0f / sqrtf(dot(a, a));
  return mk3(a.x * rn, a.y * rn, a.z * rn);
}
static inline f3 mul(const m33& m, const f3& a) { return mk3(dot(m.r[0], a), dot(m.r[1], a), dot(m.r[2], a)); }

static inline float qnan() {
  uint32_t u = 0x7fffffffu;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// CUDA __float2int_rn: round-to-nearest-even, NaN -> 0, saturating.
static inline int f2i_rn(float x) {
  if (std::isnan(x)) return 0;
  if (x >= 2147483648.0f) return INT_MAX;
  if (x <= -2147483648.0f) return INT_MIN;
  return (int)nearbyintf(x);
}
// CUDA float -> int truncation (cvt.rzi): NaN -> 0, saturating.
static inline int f2i_rz(float x) {
  if (std::isnan(x)) return 0;
  if (x >= 2147483648.0f) return INT_MAX;
  if (x <= -2147483648.0f) return INT_MIN;
  return (int)x;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

}  // namespace efo

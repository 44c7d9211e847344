// Morton keys of points and leaves (host + device).
//
// pvfmm::MortonId(x,y,z) stores (uint32)floor(c * 2^15) per axis and orders ids along
// the Z-curve with z the most significant axis, then y, then x (restated from PVFMM's
// published mortonid; call sites: reference tree_functor.h:173-184,196,467-479).  That
// order equals unsigned comparison of the bit-interleaved anchor, so a key here is the
// 45-bit interleave.  The depth field of a MortonId only orders ids with equal anchors;
// point ids carry depth 15 >= any leaf depth, so "leaf <= point" never depends on it.
// Anchors outside the 15-bit range (coordinate >= 1 or negative -- the reference's
// float->unsigned conversion wraps to a huge value on x86) sort after every in-domain
// key: they are saturated to UINT64_MAX and land in the last leaf, where the Chebyshev
// basis is zero outside [-1,1] (value 0).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

namespace tb {

TB_HD uint64_t spread3(uint32_t v) {  // bit b -> bit 3b
  uint64_t x = v & 0x1fffffu;
  x = (x | (x << 32)) & 0x001f00000000ffffull;
  x = (x | (x << 16)) & 0x001f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

// tree_functor.h:442-449: one conditional add, then one conditional subtract.
TB_HD double wrap_periodic(double c) {
  if (c < 0.0) c = c + 1.0;
  if (c >= 1.0) c = c - 1.0;
  return c;
}

// tree_functor.h:464-479: a coordinate that is exactly 1.0 moves down by 2^-15 unless
// the boundary is periodic; then floor(c * 2^15) per axis (exact: power-of-two scale).
TB_HD uint64_t point_key(double x, double y, double z, int periodic) {
  const double shift = 1.0 / 32768.0;
  if (!periodic) {
    if (x == 1.0) x = x - shift;
    if (y == 1.0) y = y - shift;
    if (z == 1.0) z = z - shift;
  }
  const double fx = floor(x * 32768.0), fy = floor(y * 32768.0), fz = floor(z * 32768.0);
  const bool in = fx >= 0.0 && fx < 32768.0 && fy >= 0.0 && fy < 32768.0 && fz >= 0.0 &&
                  fz < 32768.0;  // false for NaN
  if (!in) return ~0ull;
  return spread3((uint32_t)fx) | (spread3((uint32_t)fy) << 1) | (spread3((uint32_t)fz) << 2);
}

// The same key from the integer anchors (device fast path: 32-bit logic only).  bit b of
// an anchor goes to bit 3b (+0 x, +1 y, +2 z); the low 10 bits of each axis fill the low
// 30 key bits, the high 5 bits the next 15.
TB_HD uint32_t spread10(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
TB_HD uint64_t anchor_key(uint32_t ix, uint32_t iy, uint32_t iz) {
  const uint32_t lo = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
  const uint32_t hi = spread10(ix >> 10) | (spread10(iy >> 10) << 1) | (spread10(iz >> 10) << 2);
  return ((uint64_t)hi << 30) | lo;
}

// Cheb_Node::GetMortonId() = MortonId(Coord(), Depth()): anchor of the lower corner.
TB_HD uint64_t leaf_key(double cx, double cy, double cz) { return point_key(cx, cy, cz, 1); }

}  // namespace tb

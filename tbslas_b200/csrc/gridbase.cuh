// Arrival points as a function of their index.
//
// In a tree-level step (tbslas::SolveSemilagInSitu, reference src/tree/tree_semilag.h:103-124) the
// start points of the trajectories are the Chebyshev grid points of the advected tree's own leaves
// (CollectChebTreeGridPoints, tree_utils.h:442-498): point i of the call is node (i mod P) of leaf
// leaf0 + i / P.  The RK2 update  x' = x + tau * v  (traj.inc:42) needs x once more after the first
// stage; rebuilding it here -- with the very expressions gridpts.cu uses, so the bits are the same --
// means the 24 B per point of x are never written to or read from HBM.
#pragma once
#include "common.cuh"
#include "keys.cuh"

namespace tb {

// coordinate `axis` (0 x, 1 y, 2 z) of point i; g = geometry {cx, cy, cz, 2*2^depth} of its leaf
__device__ __forceinline__ double grid_base_coord(const GridBase &gb, const double4 &g, unsigned i, int axis) {
  const unsigned P2 = gb.D * gb.D;
  const unsigned r = i % gb.P;
  const unsigned pz = r / P2, rem = r - pz * P2, py = rem / gb.D, px = rem - py * gb.D;
  const unsigned k = axis == 0 ? px : (axis == 1 ? py : pz);
  const double c = axis == 0 ? g.x : (axis == 1 ? g.y : g.z);
  // 2^-depth = 2 / g.w exactly (g.w = 2 * 2^depth is a power of two): exponent arithmetic, no division
  const double len = __longlong_as_double((2047ll << 52) - __double_as_longlong(g.w));
  double b = __dadd_rn(c, __dmul_rn(len, gb.node[k]));  // gridpts.cu
  if (gb.periodic) b = wrap_periodic(b);                // the first evaluation wrapped x in place (tree_functor.h:442-449)
  return b;
}

// The same with the nodes per axis known at compile time (divisions by constants) and the node table in
// shared memory (a lane-divergent index into the kernel's parameter space is replayed per distinct index):
// what the evaluation kernel's RK2 epilogue uses.
template <int D>
__device__ __forceinline__ double grid_base_coord_ct(int periodic, const double *node, const double4 &g, unsigned i,
                                                     int axis) {
  constexpr unsigned P2 = D * D, P = P2 * D;
  const unsigned r = i % P;
  const unsigned pz = r / P2, rem = r - pz * P2, py = rem / D, px = rem - py * D;
  const unsigned k = axis == 0 ? px : (axis == 1 ? py : pz);
  const double c = axis == 0 ? g.x : (axis == 1 ? g.y : g.z);
  const double len = __longlong_as_double((2047ll << 52) - __double_as_longlong(g.w));
  double b = __dadd_rn(c, __dmul_rn(len, node[k]));  // gridpts.cu
  if (periodic) b = wrap_periodic(b);
  return b;
}

}  // namespace tb

// Arrival points of a tree's leaves: tbslas::CollectChebTreeGridPoints (reference
// src/tree/tree_utils.h:442-498) over tbslas::new_nodes (src/utils/cheb.h:41-68).
// out[(leaf*P + i)*3 + a] = coord[leaf][a] + 2^-depth * node[i][a], leaf-major,
// node grid x fastest.  The 1-D node table is computed on the host with the same libm
// cos() the CPU path uses, so the points are bit-identical; the kernel is a pure
// streaming write (24 B/point).
#include <cmath>

#include "common.cuh"

namespace tb {

struct Nodes1D {
  double x[TBSLAS_MAX_CHEB_DEG + 1];
};

void new_nodes_host(int q, double *x) {  // cheb.h:51-58
  const unsigned d = q + 1;
  const double pi = 3.14159265358979323846264338327950288;
  volatile double scal = 1.0 / std::cos(0.5 * pi / d);
  for (unsigned i = 0; i < d; i++) {
    volatile double a = (i + 0.5) * pi;
    volatile double b = a / d;
    volatile double c = -std::cos(b);
    volatile double e = c * scal;
    volatile double f = e * 0.5;
    x[i] = f + 0.5;
  }
}

// One CTA per leaf (grid-stride over leaves), one thread per output scalar of that leaf: all index
// arithmetic is 32-bit and the stores of a warp are one contiguous 256-B run.
__global__ void __launch_bounds__(256)
grid_points_kernel(const double4 *__restrict__ geom, const uint8_t *__restrict__ depth,
                   size_t n_leaf, int d, Nodes1D nodes, double *__restrict__ out) {
  __shared__ double s_node[TBSLAS_MAX_CHEB_DEG + 1];
  if (threadIdx.x <= TBSLAS_MAX_CHEB_DEG) s_node[threadIdx.x] = nodes.x[threadIdx.x];
  __syncthreads();
  const unsigned ud = (unsigned)d, dd = ud * ud, row_len = 3u * ud;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warp = blockDim.x >> 5;
  for (size_t leaf = blockIdx.x; leaf < n_leaf; leaf += gridDim.x) {
    const double4 g = geom[leaf];
    const double len = 1.0 / (double)(1u << depth[leaf]);  // pow(0.5, depth), exact
    double *o = out + leaf * (size_t)(3u * dd * ud);
    // a warp writes one x-row of the grid (3*d contiguous doubles) per iteration
    for (unsigned row = warp; row < dd; row += n_warp) {
      const unsigned iz = row / ud, iy = row - iz * ud;
      const double y = __dadd_rn(g.y, __dmul_rn(len, s_node[iy]));
      const double z = __dadd_rn(g.z, __dmul_rn(len, s_node[iz]));
      double *orow = o + (size_t)row * row_len;
      for (unsigned e = lane; e < row_len; e += 32) {
        const unsigned ix = e / 3u, a = e - 3u * ix;
        orow[e] = (a == 0) ? __dadd_rn(g.x, __dmul_rn(len, s_node[ix])) : (a == 1 ? y : z);
      }
    }
  }
}

// leaves [leaf0, leaf0 + n_leaf) of the tree -> out[n_leaf * P][3]
int launch_grid_points(tbslas_ctx *ctx, const tbslas_tree *t, double *out, size_t leaf0, size_t n_leaf) {
  if (leaf0 > t->n_leaf) leaf0 = t->n_leaf;
  if (n_leaf > t->n_leaf - leaf0) n_leaf = t->n_leaf - leaf0;
  const int d = t->q + 1;
  const size_t total = n_leaf * (size_t)d * d * d * 3;
  if (!total) return TBSLAS_OK;
  StageScope sc(ctx, ST_GRIDPTS, (double)(total / 3), 1);
  Nodes1D nodes;
  new_nodes_host(t->q, nodes.x);
  const size_t want = n_leaf < (size_t)ctx->n_sm * 64 ? n_leaf : (size_t)ctx->n_sm * 64;
  grid_points_kernel<<<(unsigned)want, 256, 0, ctx->stream>>>(t->d_geom + leaf0, t->d_depth + leaf0, n_leaf, d,
                                                             nodes, out);
  TB_CUDA(ctx, cudaGetLastError());
  return TBSLAS_OK;
}

}  // namespace tb

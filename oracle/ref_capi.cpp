// oracle/_ref: the REFERENCE'S OWN hot-path code behind a C ABI.
//
// TEST INFRASTRUCTURE ONLY.  This translation unit #includes the reference's
// headers verbatim from /root/reference/src (they are never copied into this
// repository) over the PVFMM/MPI stand-in in oracle/shim/, and exports thin C
// wrappers so the Python test-suite and bench.py's `--impl reference` arm can
// call tbslas::NodeFieldFunctor / ComputeTrajRK2 / SolveSemilagRK2 / fast_interp /
// FieldSetFunctor / FieldExtrapFunctor / CubicInterpPolicy / new_nodes exactly as
// the reference's drivers do (e.g. src/applications/src/advection.cpp:296).
// Built by oracle/Makefile into oracle/_ref/libtbslas_ref.so (git-ignored; it
// travels to the GPU box as a prebuilt file).  Only tests/, smoke() and
// bench.py's CPU-baseline legs may load it.
#include <mpi.h>
#include <omp.h>

#include <vector>

// Include order follows the reference's drivers (advection.cpp:15-34): the
// headers are not self-contained.
// clang-format off
#include "utils/common.h"
#include "utils/cubic.h"
#include "utils/cheb.h"
#include "tree/tree_functor.h"
#include "tree/tree_set_functor.h"
#include "tree/tree_extrap_functor.h"
#include "semilag/traj.h"
#include "semilag/semilag.h"
// clang-format on

typedef pvfmm::Cheb_Node<double> Node_t;
typedef pvfmm::MPI_Tree<Node_t> Tree_t;

static inline long ncoef(int q) { return (long)(q + 1) * (q + 2) * (q + 3) / 6; }

extern "C" {

void ref_set_num_threads(int n) { omp_set_num_threads(n); }
int ref_get_max_threads() { return omp_get_max_threads(); }

// bc: 0 = FreeSpace, 1 = Periodic (pvfmm::BoundaryType); read implicitly by the
// reference through the SimConfig singleton (tree_functor.h:174,469,803).
void ref_set_bc(int bc) {
  tbslas::SimConfigSingleton::Instance()->bc =
      bc ? pvfmm::Periodic : pvfmm::FreeSpace;
}

// Leaves must be given in Morton order (what PVFMM's traversal yields).
void *ref_tree_create(int q, int dof, long n_leaf, const double *coord,
                      const unsigned char *depth, const double *coeff) {
  Tree_t *t = new Tree_t();
  const long nc = ncoef(q) * dof;
  std::vector<Node_t *> &nodes = t->GetNodeList();
  nodes.resize(n_leaf);
  for (long i = 0; i < n_leaf; i++) {
    Node_t *n = new Node_t();
    n->deg_ = q;
    n->dof_ = dof;
    n->depth_ = depth[i];
    for (int k = 0; k < 3; k++) n->coord_[k] = coord[3 * i + k];
    n->data_.Resize(nc);
    for (long k = 0; k < nc; k++) n->data_[k] = coeff[i * nc + k];
    nodes[i] = n;
  }
  return t;
}
void ref_tree_destroy(void *t) { delete (Tree_t *)t; }

// tbslas::NodeFieldFunctor::operator() (tree_functor.h:800-806).  pos is mutated
// when bc is Periodic, exactly like the reference (const_cast at :803).
void ref_eval_tree(void *tree, double *pos, long n, double *out) {
  tbslas::NodeFieldFunctor<double, Tree_t> f((Tree_t *)tree);
  f(pos, (int)n, out);
}

// Leaf assignment as EvalNodesLocal makes it (tree_functor.h:166-198): keys with
// the x==1 shift, sort, part_indx[j] = lower_bound(sorted keys, key(leaf j)),
// point belongs to leaf j iff part_indx[j] <= rank < part_indx[j+1]; -1 when no
// leaf claims it.  pos must already be wrapped when bc is Periodic.
void ref_leaf_index(void *tree, const double *pos, long n, int *leaf_out) {
  tbslas::SimConfig *cfg = tbslas::SimConfigSingleton::Instance();
  std::vector<Node_t *> &nodes = ((Tree_t *)tree)->GetNodeList();
  typedef pvfmm::par::SortPair<pvfmm::MortonId, size_t> Pair_t;
  std::vector<Pair_t> pk(n);
  const double shift = 1.0 / (1UL << MAX_DEPTH);
#pragma omp parallel for
  for (long i = 0; i < n; i++) {
    double c[3];
    for (int k = 0; k < 3; k++) {
      c[k] = pos[3 * i + k];
      if (c[k] == 1.0 && cfg->bc != pvfmm::Periodic) c[k] -= shift;
    }
    pk[i].key = pvfmm::MortonId(c[0], c[1], c[2]);
    pk[i].data = i;
  }
  pvfmm::omp_par::merge_sort(pk.begin(), pk.end());
  std::vector<pvfmm::MortonId> keys(n);
  for (long i = 0; i < n; i++) keys[i] = pk[i].key;
  std::vector<size_t> part(nodes.size() + 1);
  part[nodes.size()] = n;
  for (size_t j = 0; j < nodes.size(); j++)
    part[j] = std::lower_bound(keys.begin(), keys.end(), nodes[j]->GetMortonId()) -
              keys.begin();
  for (long i = 0; i < n; i++) leaf_out[i] = -1;
  for (size_t j = 0; j < nodes.size(); j++)
    for (size_t s = part[j]; s < part[j + 1]; s++) leaf_out[pk[s].data] = (int)j;
}

// tbslas::ComputeTrajRK2 (traj.inc:49-68) over a NodeFieldFunctor.
void ref_traj_rk2(void *vel, const double *pos, long n, double tinit,
                  double tfinal, int nrk, double *out) {
  std::vector<double> xi(pos, pos + 3 * n), xs(3 * n);
  tbslas::NodeFieldFunctor<double, Tree_t> f((Tree_t *)vel);
  tbslas::ComputeTrajRK2(f, xi, tinit, tfinal, nrk, xs);
  std::copy(xs.begin(), xs.end(), out);
}

// tbslas::SolveSemilagRK2 (semilag.inc:27-45).
void ref_semilag_rk2(void *vel, void *con, int dof_con, const double *pos, long n,
                     int timestep, double dt, int nrk, double *out_vals) {
  std::vector<double> xi(pos, pos + 3 * n), vals(n * dof_con);
  tbslas::NodeFieldFunctor<double, Tree_t> fv((Tree_t *)vel);
  tbslas::NodeFieldFunctor<double, Tree_t> fc((Tree_t *)con);
  tbslas::SolveSemilagRK2(fv, fc, xi, 3, timestep, dt, nrk, vals);
  std::copy(vals.begin(), vals.end(), out_vals);
}

// tbslas::FieldSetFunctor::operator() (tree_set_functor.h:49-79).
void ref_eval_set4(void **trees, const double *times, double t, double *pos,
                   long n, double *out) {
  std::vector<Tree_t *> tv(4);
  std::vector<double> tt(times, times + 4);
  for (int i = 0; i < 4; i++) tv[i] = (Tree_t *)trees[i];
  tbslas::FieldSetFunctor<double, Tree_t> f(tv, tt);
  f(pos, (int)n, t, out);
}

// tbslas::FieldExtrapFunctor::operator() (tree_extrap_functor.h:47-78).
void ref_eval_extrap(void *tp, void *tc, double *pos, long n, double *out) {
  tbslas::FieldExtrapFunctor<double, Tree_t> f((Tree_t *)tp, (Tree_t *)tc);
  f(pos, (int)n, out);
}

// Config-3 path: ComputeTrajRK2 over a FieldSetFunctor (advtv.cpp:301 pattern).
void ref_traj_rk2_set4(void **trees, const double *times, const double *pos,
                       long n, double tinit, double tfinal, int nrk, double *out) {
  std::vector<Tree_t *> tv(4);
  std::vector<double> tt(times, times + 4);
  for (int i = 0; i < 4; i++) tv[i] = (Tree_t *)trees[i];
  tbslas::FieldSetFunctor<double, Tree_t> f(tv, tt);
  std::vector<double> xi(pos, pos + 3 * n), xs(3 * n);
  tbslas::ComputeTrajRK2(f, xi, tinit, tfinal, nrk, xs);
  std::copy(xs.begin(), xs.end(), out);
}

void ref_semilag_rk2_set4(void **trees, const double *times, void *con,
                          int dof_con, const double *pos, long n, int timestep,
                          double dt, int nrk, double *out_vals) {
  std::vector<Tree_t *> tv(4);
  std::vector<double> tt(times, times + 4);
  for (int i = 0; i < 4; i++) tv[i] = (Tree_t *)trees[i];
  tbslas::FieldSetFunctor<double, Tree_t> fv(tv, tt);
  tbslas::NodeFieldFunctor<double, Tree_t> fc((Tree_t *)con);
  std::vector<double> xi(pos, pos + 3 * n), vals(n * dof_con);
  tbslas::SolveSemilagRK2(fv, fc, xi, 3, timestep, dt, nrk, vals);
  std::copy(vals.begin(), vals.end(), out_vals);
}

// Extrapolated variant (traj.inc:95-115): first stage samples tc, second stage
// samples 1.5 tc - 0.5 tp (tree_ns.h:471-483 pattern).
void ref_traj_rk2_extrap(void *tp, void *tc, const double *pos, long n,
                         double tinit, double tfinal, int nrk, double *out) {
  tbslas::NodeFieldFunctor<double, Tree_t> f((Tree_t *)tc);
  tbslas::FieldExtrapFunctor<double, Tree_t> fe((Tree_t *)tp, (Tree_t *)tc);
  std::vector<double> xi(pos, pos + 3 * n), xs(3 * n);
  tbslas::ComputeTrajRK2(f, fe, xi, tinit, tfinal, nrk, xs);
  std::copy(xs.begin(), xs.end(), out);
}

// tbslas::fast_interp (tree_functor.h:89-153).
void ref_fast_interp(const double *grid, int dof, int n_reg, const double *pts,
                     long n, double *out) {
  std::vector<double> g(grid, grid + (size_t)dof * n_reg * n_reg * n_reg);
  std::vector<double> p(pts, pts + 3 * n), v;
  tbslas::fast_interp(g, dof, n_reg, p, v);
  std::copy(v.begin(), v.end(), out);
}

// tbslas::CubicInterpPolicy::InterpCubic1D (cubic.h:42-56).
double ref_interp_cubic1d(double x, const double *xx, const double *pp) {
  double a[4] = {xx[0], xx[1], xx[2], xx[3]};
  double b[4] = {pp[0], pp[1], pp[2], pp[3]};
  return tbslas::CubicInterpPolicy<double>::InterpCubic1D(x, a, b);
}

// tbslas::new_nodes (cheb.h:41-164); out has (q+1)^dim * dim doubles.
long ref_new_nodes(int q, int dim, double *out) {
  std::vector<double> y = tbslas::new_nodes<double>(q, dim);
  if (out) std::copy(y.begin(), y.end(), out);
  return (long)y.size();
}

// tbslas::GetPt2CoeffMatrix (cheb.h:166-196); M is (q+1)^3 x Ncoef row-major.
void ref_pt2coeff(int q, double *M_out) {
  pvfmm::Matrix<double> M;
  tbslas::GetPt2CoeffMatrix<double>(q, M);
  std::memcpy(M_out, M[0], M.Dim(0) * M.Dim(1) * sizeof(double));
}

// pvfmm::MortonId comparator of the shim, for pinning the 48-bit key encoding.
int ref_morton_less(double ax, double ay, double az, int ad, double bx, double by,
                    double bz, int bd) {
  return pvfmm::MortonId(ax, ay, az, (uint8_t)ad) <
         pvfmm::MortonId(bx, by, bz, (uint8_t)bd);
}
void ref_morton_xyz(double x, double y, double z, unsigned *out) {
  pvfmm::MortonId m(x, y, z);
  out[0] = m.x;
  out[1] = m.y;
  out[2] = m.z;
}

}  // extern "C"

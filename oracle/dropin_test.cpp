// oracle/dropin_test.cpp -- the C++ drop-in boundary, exercised by the reference's own code.
//
// TEST INFRASTRUCTURE ONLY (built into oracle/_ref/dropin_test by oracle/Makefile where
// /root/reference exists; the binary travels to the GPU box, no reference source is copied).
//
// This translation unit #includes the reference's headers verbatim over the PVFMM/MPI
// stand-in (oracle/shim) AND the product's header-only adaptors
// (include/tbslas_b200/functors.hpp), links libtbslas_b200.so, and runs the reference's
// call sites three ways on identical trees:
//   (A) reference templates over the reference's CPU functors     (the oracle)
//   (B) reference templates, UNCHANGED, over tbslas::b200 functors (drop-in: every functor
//       call is a GPU evaluation through the C ABI)
//   (C) tbslas::b200 fused overloads with the reference's signatures (one C-ABI call/step)
// for SolveSemilagRK2 (semilag.inc:27-45), ComputeTrajRK2 with a FieldSetFunctor
// (traj.inc:49-68, tree_set_functor.h:49-79), the extrapolated two-functor form
// (traj.inc:95-115), the tree-level SolveSemilagInSitu in its one- and two-functor forms
// (tree_semilag.h:92-135, :137-181) and the Navier-Stokes call pattern (tree_ns.h:466-518: two
// trajectories from the same arrival points, dof-3 advected field, 2/dt*c - 0.5/dt*p, refit).
// Bars: trajectories 1e-12; composed steps 1e-11 of the field scale (see kStepTol).
#include <mpi.h>
#include <omp.h>

#include <cmath>
#include <cstdio>
#include <vector>

// clang-format off
#include "utils/common.h"
#include "utils/cubic.h"
#include "utils/cheb.h"
#include "tree/tree_functor.h"
#include "tree/tree_set_functor.h"
#include "tree/tree_extrap_functor.h"
#include "semilag/traj.h"
#include "semilag/semilag.h"
#include "tree/tree_utils.h"
#include "tree/tree_semilag.h"
#include "tbslas_b200/functors.hpp"
// clang-format on

typedef pvfmm::Cheb_Node<double> Node_t;
typedef pvfmm::MPI_Tree<Node_t> Tree_t;

static long ncoef(int q) { return (long)(q + 1) * (q + 2) * (q + 3) / 6; }
// packed index of degree (i: z, j: y, k: x), order of tree_functor.h:256-266
static long cidx(int q, int i, int j, int k) {
  const int d = q + 1;
  long o = 0;
  for (int a = 0; a < i; a++) o += (long)(d - a) * (d - a + 1) / 2;
  for (int b = 0; b < j; b++) o += d - i - b;
  return o + k;
}

// uniform octree of the given depth in Morton (z-major) order
static void morton_leaves(int depth, double cx, double cy, double cz, int level, std::vector<double> &out) {
  if (level == depth) {
    out.push_back(cx);
    out.push_back(cy);
    out.push_back(cz);
    return;
  }
  const double h = std::pow(0.5, level + 1);
  for (int c = 0; c < 8; c++)
    morton_leaves(depth, cx + (c & 1) * h, cy + ((c >> 1) & 1) * h, cz + ((c >> 2) & 1) * h, level + 1, out);
}

static Tree_t *make_tree(int depth, int q, int dof, int kind, double scale) {
  std::vector<double> c;
  morton_leaves(depth, 0, 0, 0, 0, c);
  const long L = c.size() / 3, nc = ncoef(q);
  const double h = std::pow(0.5, depth);
  Tree_t *t = new Tree_t();
  std::vector<Node_t *> &nodes = t->GetNodeList();
  nodes.resize(L);
  unsigned long long s = 88172645463325252ull + kind;
  for (long l = 0; l < L; l++) {
    Node_t *n = new Node_t();
    n->deg_ = q;
    n->dof_ = dof;
    n->depth_ = depth;
    for (int k = 0; k < 3; k++) n->coord_[k] = c[3 * l + k];
    n->data_.Resize(nc * dof);
    for (long k = 0; k < nc * dof; k++) n->data_[k] = 0;
    if (kind == 0) {  // solid-body rotation (0.5 - y, x - 0.5, 0) * scale, exact in T0/T1
      n->data_[0 * nc + cidx(q, 0, 0, 0)] = scale * (0.5 - c[3 * l + 1] - 0.5 * h);
      n->data_[0 * nc + cidx(q, 0, 1, 0)] = -scale * 0.5 * h;
      n->data_[1 * nc + cidx(q, 0, 0, 0)] = scale * (c[3 * l + 0] + 0.5 * h - 0.5);
      n->data_[1 * nc + cidx(q, 0, 0, 1)] = scale * 0.5 * h;
    } else {  // pseudo-random coefficients decaying with total degree
      for (int d0 = 0; d0 < dof; d0++)
        for (int i = 0; i <= q; i++)
          for (int j = 0; i + j <= q; j++)
            for (int k = 0; i + j + k <= q; k++) {
              s ^= s << 13;
              s ^= s >> 7;
              s ^= s << 17;
              const double u = (double)(s >> 11) / 9007199254740992.0 * 2 - 1;
              n->data_[d0 * nc + cidx(q, i, j, k)] = u * std::pow(0.5, i + j + k);
            }
    }
    nodes[l] = n;
  }
  return t;
}

static double maxabs(const std::vector<double> &a) {
  double m = 0;
  for (size_t i = 0; i < a.size(); i++) m = std::fmax(m, std::fabs(a[i]));
  return m;
}
static double maxdiff(const std::vector<double> &a, const std::vector<double> &b) {
  if (a.size() != b.size()) return 1e300;
  double m = 0;
  for (size_t i = 0; i < a.size(); i++) m = std::fmax(m, std::fabs(a[i] - b[i]));
  return m;
}

// Bars, relative to the field scale.  A composed step (trajectory, then the advected field AT the
// departure point) carries the departure point's last-bit difference times the field's steepness: the
// test fields here are pseudo-random polynomials of degree 6 (gradients of O(10..100)), so the composed
// bar is 1e-11 while trajectories are held to 1e-12; the NS pattern multiplies by 2/dt = 32 on top.
static const double kStepTol = 1e-11, kNsTol = 1e-10;
static int n_fail = 0;
static void report(const char *what, double err, double tol) {
  printf("  %-58s %.3e  %s\n", what, err, err <= tol ? "ok" : "FAIL");
  if (!(err <= tol)) n_fail++;
}

int main() {
  const int q = 6, depth = 2, nrk = 2, timestep = 3;
  const double dt = 0.0628;
  tbslas::SimConfig *cfg = tbslas::SimConfigSingleton::Instance();
  try {
    for (int bc = 0; bc < 2; bc++) {
      cfg->bc = bc ? pvfmm::Periodic : pvfmm::FreeSpace;
      printf("boundary = %s\n", bc ? "Periodic" : "FreeSpace");
      Tree_t *tvel = make_tree(depth, q, 3, 0, 1.0), *tcon = make_tree(depth, q, 1, 7, 1.0);
      std::vector<double> pos;
      tbslas::CollectChebTreeGridPoints(*tcon, pos);
      const size_t n = pos.size() / 3;

      // ---- SolveSemilagRK2 (advection.cpp:296 -> tree_semilag.h:124 -> semilag.inc:27)
      std::vector<double> a(n), b(n), c(n);
      tbslas::NodeFieldFunctor<double, Tree_t> rvel(tvel), rcon(tcon);
      tbslas::SolveSemilagRK2(rvel, rcon, pos, 3, timestep, dt, nrk, a);
      tbslas::b200::NodeFieldFunctor<double, Tree_t> gvel(tvel), gcon(tcon);
      tbslas::SolveSemilagRK2(gvel, gcon, pos, 3, timestep, dt, nrk, b);  // reference template, GPU functors
      tbslas::b200::SolveSemilagRK2(gvel, gcon, pos, 3, timestep, dt, nrk, c);
      const double sc = maxabs(a);
      report("SolveSemilagRK2: reference template over b200 functors", maxdiff(a, b) / sc, kStepTol);
      report("SolveSemilagRK2: b200 fused overload", maxdiff(a, c) / sc, kStepTol);
      report("                 fused == functor-by-functor (bitwise)", maxdiff(b, c), 0.0);

      // ---- time-varying velocity: FieldSetFunctor (advtv.cpp:171-190,301)
      std::vector<Tree_t *> set_r, set_g;
      std::vector<double> times;
      for (int i = 0; i < 4; i++) {
        set_r.push_back(make_tree(depth, q, 3, 0, 0.7 + 0.2 * i));
        set_g.push_back(make_tree(depth, q, 3, 0, 0.7 + 0.2 * i));
        times.push_back(dt * (i - 1));
      }
      tbslas::FieldSetFunctor<double, Tree_t> rset(set_r, times);
      tbslas::b200::FieldSetFunctor<double, Tree_t> gset(set_g, times);
      std::vector<double> xa(3 * n), xb(3 * n), xc(3 * n);
      tbslas::ComputeTrajRK2(rset, pos, dt, 0.0, nrk, xa);
      tbslas::ComputeTrajRK2(gset, pos, dt, 0.0, nrk, xb);
      tbslas::b200::ComputeTrajRK2(gset, pos, dt, 0.0, nrk, xc);
      report("ComputeTrajRK2(FieldSetFunctor): reference template", maxdiff(xa, xb), 1e-12);
      report("ComputeTrajRK2(FieldSetFunctor): b200 fused", maxdiff(xa, xc), 1e-12);

      // ---- extrapolated velocity, two-functor form (tree_ns.h:471-483)
      Tree_t *tp_r = make_tree(depth, q, 3, 0, 0.9), *tc_r = make_tree(depth, q, 3, 0, 1.0);
      Tree_t *tp_g = make_tree(depth, q, 3, 0, 0.9), *tc_g = make_tree(depth, q, 3, 0, 1.0);
      tbslas::FieldExtrapFunctor<double, Tree_t> rext(tp_r, tc_r);
      tbslas::b200::FieldExtrapFunctor<double, Tree_t> gext(tp_g, tc_g);
      tbslas::NodeFieldFunctor<double, Tree_t> rcur(tc_r);
      tbslas::b200::NodeFieldFunctor<double, Tree_t> gcur(tc_g);
      tbslas::ComputeTrajRK2(rcur, rext, pos, dt, 0.0, nrk, xa);
      tbslas::ComputeTrajRK2(gcur, gext, pos, dt, 0.0, nrk, xb);
      tbslas::b200::ComputeTrajRK2(gcur, gext, pos, dt, 0.0, nrk, xc);
      report("ComputeTrajRK2(v, extrap): reference template", maxdiff(xa, xb), 1e-12);
      report("ComputeTrajRK2(v, extrap): b200 fused", maxdiff(xa, xc), 1e-12);

      // ---- tree-level step: SolveSemilagInSitu (advection.cpp:296), in place on the tree
      Tree_t *con_r = make_tree(depth, q, 1, 7, 1.0), *con_g = make_tree(depth, q, 1, 7, 1.0),
             *con_f = make_tree(depth, q, 1, 7, 1.0);
      tbslas::SolveSemilagInSitu(rvel, *con_r, timestep, dt, nrk);
      tbslas::SolveSemilagInSitu(gvel, *con_g, timestep, dt, nrk);        // reference template, GPU velocity
      tbslas::b200::SolveSemilagInSitu(gvel, *con_f, timestep, dt, nrk);  // GPU step, reference refit
      std::vector<double> ca, cb, cc;
      for (size_t l = 0; l < con_r->GetNodeList().size(); l++)
        for (size_t k = 0; k < con_r->GetNodeList()[l]->ChebData().Dim(); k++) {
          ca.push_back(con_r->GetNodeList()[l]->ChebData()[k]);
          cb.push_back(con_g->GetNodeList()[l]->ChebData()[k]);
          cc.push_back(con_f->GetNodeList()[l]->ChebData()[k]);
        }
      report("SolveSemilagInSitu coefficients: reference template", maxdiff(ca, cb) / maxabs(ca), kStepTol);
      report("SolveSemilagInSitu coefficients: b200 in-situ", maxdiff(ca, cc) / maxabs(ca), kStepTol);

      // ---- the two-functor tree-level step (tree_semilag.h:137-181; advtvextrap.cpp, ns.cpp)
      Tree_t *c2_r = make_tree(depth, q, 1, 7, 1.0), *c2_g = make_tree(depth, q, 1, 7, 1.0),
             *c2_f = make_tree(depth, q, 1, 7, 1.0);
      tbslas::SolveSemilagInSitu(rcur, rext, *c2_r, timestep, dt, nrk);
      tbslas::SolveSemilagInSitu(gcur, gext, *c2_g, timestep, dt, nrk);        // reference template, GPU functors
      tbslas::b200::SolveSemilagInSitu(gcur, gext, *c2_f, timestep, dt, nrk);  // one C-ABI call + read-back
      ca.clear(), cb.clear(), cc.clear();
      for (size_t l = 0; l < c2_r->GetNodeList().size(); l++)
        for (size_t k = 0; k < c2_r->GetNodeList()[l]->ChebData().Dim(); k++) {
          ca.push_back(c2_r->GetNodeList()[l]->ChebData()[k]);
          cb.push_back(c2_g->GetNodeList()[l]->ChebData()[k]);
          cc.push_back(c2_f->GetNodeList()[l]->ChebData()[k]);
        }
      report("SolveSemilagInSitu(v, extrap): reference template", maxdiff(ca, cb) / maxabs(ca), kStepTol);
      report("SolveSemilagInSitu(v, extrap): b200 in-situ", maxdiff(ca, cc) / maxabs(ca), kStepTol);

      // ---- the Navier-Stokes call pattern (tree_ns.h:466-518): the advected field is the dof-3
      // velocity itself; two backward trajectories from the same arrival points,
      //   [t, t-dt]  stage 1 = v^n,     stage 2 = extrapolation;  sample v^n     there
      //   [t, t-2dt] stage 1 = v^{n-1}, stage 2 = v^n;            sample v^{n-1} there
      // combined 2/dt * c - 0.5/dt * p, transposed point-major -> dof-major per leaf, refitted.
      {
        const double tcurr = timestep * dt, ccoeff = 2.0 / dt, pcoeff = -0.5 / dt;
        const int P = (q + 1) * (q + 1) * (q + 1), dof = 3;
        Tree_t *vp[2] = {make_tree(depth, q, 3, 9, 1.0), make_tree(depth, q, 3, 9, 1.0)};   // v^{n-1}: rough field
        Tree_t *vc[2] = {make_tree(depth, q, 3, 11, 1.0), make_tree(depth, q, 3, 11, 1.0)}; // v^n
        Tree_t *tn[2] = {make_tree(depth, q, 3, 11, 1.0), make_tree(depth, q, 3, 11, 1.0)}; // tree that receives the result
        std::vector<double> res[2];
        for (int side = 0; side < 2; side++) {  // 0: reference functors (CPU), 1: b200 functors (GPU)
          std::vector<double> arr, dep, cval, pval;
          const int num_leaf = tbslas::CollectChebTreeGridPoints(*tn[side], arr);
          const int np = arr.size() / 3;
          dep.resize(arr.size());
          cval.resize((size_t)np * dof);
          pval.resize((size_t)np * dof);
          if (side == 0) {
            tbslas::NodeFieldFunctor<double, Tree_t> fp(vp[0]), fc(vc[0]);
            tbslas::FieldExtrapFunctor<double, Tree_t> fe(vp[0], vc[0]);
            tbslas::ComputeTrajRK2(fc, fe, arr, tcurr, tcurr - dt, nrk, dep);
            fc(dep.data(), np, cval.data());
            tbslas::ComputeTrajRK2(fp, fc, arr, tcurr, tcurr - dt * 2, nrk, dep);
            fp(dep.data(), np, pval.data());
          } else {
            tbslas::b200::NodeFieldFunctor<double, Tree_t> fp(vp[1]), fc(vc[1]);
            tbslas::b200::FieldExtrapFunctor<double, Tree_t> fe(vp[1], vc[1]);
            tbslas::ComputeTrajRK2(fc, fe, arr, tcurr, tcurr - dt, nrk, dep);   // the reference's template
            fc(dep.data(), np, cval.data());
            tbslas::b200::ComputeTrajRK2(fp, fc, arr, tcurr, tcurr - dt * 2, nrk, dep);  // fused overload
            fp(dep.data(), np, pval.data());
          }
          std::vector<double> val((size_t)np * dof), ml((size_t)np * dof);
          for (size_t i = 0; i < val.size(); i++) val[i] = ccoeff * cval[i] + pcoeff * pval[i];
          for (int nindx = 0; nindx < num_leaf; nindx++) {
            const size_t shift = (size_t)nindx * P * dof;
            for (int j = 0; j < P; j++)
              for (int i = 0; i < dof; i++) ml[shift + j + (size_t)i * P] = val[shift + (size_t)j * dof + i];
          }
          if (side == 0) {
            tbslas::SetTreeGridValues(*tn[0], q, dof, ml);
          } else {  // device refit through the C ABI: the dof-major layout SetTreeGridValues consumes
            tbslas::b200::NodeFieldFunctor<double, Tree_t> fn(tn[1]);
            tbslas::b200::DeviceTree<Tree_t> &dn = fn.device_tree();
            int has = 0;
            dn.context().check(tbslas_b200_has_pt2coeff(dn.context().get(), q, &has));
            if (!has) {
              pvfmm::Matrix<double> M;
              tbslas::GetPt2CoeffMatrix<double>(q, M);
              dn.context().check(tbslas_b200_set_pt2coeff(dn.context().get(), q, &M[0][0]));
            }
            dn.context().check(tbslas_b200_tree_set_grid_values(dn.get(), ml.data(), /*point_major=*/0, TBSLAS_MEM_HOST));
            std::vector<double> coeff((size_t)ncoef(q) * dof * num_leaf);
            dn.context().check(tbslas_b200_tree_get_coeff(dn.get(), coeff.data(), TBSLAS_MEM_HOST));
            for (int l = 0; l < num_leaf; l++)
              for (long k = 0; k < ncoef(q) * dof; k++) tn[1]->GetNodeList()[l]->ChebData()[k] = coeff[(size_t)l * ncoef(q) * dof + k];
          }
          for (size_t l = 0; l < tn[side]->GetNodeList().size(); l++)
            for (size_t k = 0; k < tn[side]->GetNodeList()[l]->ChebData().Dim(); k++)
              res[side].push_back(tn[side]->GetNodeList()[l]->ChebData()[k]);
        }
        report("NS pattern (2 trajectories, dof 3, 2/dt c - 0.5/dt p, refit)", maxdiff(res[0], res[1]) / maxabs(res[0]),
               kNsTol);
      }
    }
  } catch (const std::exception &e) {
    printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  printf("%s\n", n_fail ? "dropin_test: FAILED" : "dropin_test: ALL OK");
  return n_fail ? 1 : 0;
}

// tbslas_b200/functors.hpp -- header-only C++ adaptors that put the C ABI of
// libtbslas_b200.so (tbslas_b200.h) back behind the reference's operator surface.
//
// arashb/tbslas's semi-Lagrangian core is generic over "field functors" -- anything
// callable as
//     f(const real_t* pos_aos, int n, real_t* out_aos);
//     f(const real_t* pos_aos, int n, real_t t, real_t* out_aos);
// (reference src/tree/tree_functor.h:800-811, tree_set_functor.h:49-50,
// tree_extrap_functor.h:47; consumed by src/semilag/traj.inc:33,40,80,86 and
// semilag.inc:43).  The classes below satisfy that concept with the GPU path, so the
// reference's own templates (tbslas::ComputeTrajRK2, tbslas::SolveSemilagRK2,
// tbslas::SolveSemilagInSitu) compile and run over them UNCHANGED:
//
//     tbslas::NodeFieldFunctor<double,Tree_t>        tvel_func(&tvel);   // reference, CPU
//     tbslas::b200::NodeFieldFunctor<double,Tree_t>  tvel_func(&tvel);   // this library, GPU
//     tbslas::SolveSemilagInSitu(tvel_func, tcon, timestep, dt, nrk);    // same call site
//
// With the reference's templates every functor call is one host->device->host round
// trip.  The overloads tbslas::b200::{ComputeTrajRK2, SolveSemilagRK2, SolveSemilagInSitu}
// at the end of this file have the reference's signatures but run the whole step in ONE
// C-ABI call (points stay in HBM between the evaluations).
//
// Tree surface used (exactly what tree_functor.h touches, :161,249-250,283-284,417-427):
//   Tree_t::GetNodeList();  Node: IsLeaf IsGhost Coord Depth ChebDeg DataDOF ChebData.
// Boundary condition: the reference reads it from its SimConfig singleton
// (tree_functor.h:174,469); include the reference's utils/common.h BEFORE this header and
// the adaptors do the same, otherwise set tbslas::b200::DefaultBC().
//
// Real_t must be double (the kernels are FP64).  Not thread safe, like the reference
// (function-static scratch, tree_functor.h:166,519).
#ifndef TBSLAS_B200_FUNCTORS_HPP_
#define TBSLAS_B200_FUNCTORS_HPP_

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../tbslas_b200.h"

namespace tbslas {
namespace b200 {

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

inline int &DefaultBC() {
  static int bc = TBSLAS_FREESPACE;
  return bc;
}

// pvfmm::BoundaryType the path runs under (tree_functor.h:174,469,803).
inline int CurrentBC() {
#ifdef SRC_UTILS_COMMON_H_  // the reference's utils/common.h is in this translation unit
  return tbslas::SimConfigSingleton::Instance()->bc == pvfmm::Periodic ? TBSLAS_PERIODIC
                                                                      : TBSLAS_FREESPACE;
#else
  return DefaultBC();
#endif
}

// One GPU context per process (the reference keeps its communicator and options in a
// process-wide singleton too, utils/common.h:42).
class Context {
 public:
  explicit Context(int device = 0) {
    tbslas_ctx *c = nullptr;
    const int rc = tbslas_b200_init(device, &c);
    if (rc != TBSLAS_OK)
      throw Error(rc, "tbslas_b200_init(device " + std::to_string(device) +
                          ") failed: no usable sm_100 GPU (there is no CPU fallback)");
    p_.reset(c, [](tbslas_ctx *x) { tbslas_b200_finalize(x); });
  }
  static Context &Default(int device = 0) {
    static Context ctx(device);
    return ctx;
  }
  tbslas_ctx *get() const { return p_.get(); }
  void check(int rc) const {
    if (rc != TBSLAS_OK) throw Error(rc, std::string("tbslas_b200: ") + tbslas_b200_last_error(p_.get()));
  }
  // Evaluation shortcuts (both on by default): coefficients of snapshot trees
  // with one leaf list combined in time before a single evaluation, and the first velocity
  // evaluation of a tree-level step by sum factorisation over the arrival grids.
  void SetTimeCombine(bool on) { check(tbslas_b200_set_time_combine(p_.get(), on ? 1 : 0)); }
  void SetTensorGrid(bool on) { check(tbslas_b200_set_tensor_grid(p_.get(), on ? 1 : 0)); }
  // Multi-GPU: `bcast128(void* buf)` must broadcast 128 bytes from rank 0 to all ranks
  // (MPI_Bcast(buf,128,MPI_BYTE,0,comm) in an MPI host such as the reference's drivers).
  template <class Bcast128>
  void CommInit(int nranks, int rank, Bcast128 bcast128) {
    unsigned char id[128];
    std::memset(id, 0, sizeof(id));
    if (rank == 0) check(tbslas_b200_comm_unique_id(id));
    bcast128(static_cast<void *>(id));
    check(tbslas_b200_comm_init(p_.get(), nranks, rank, id));
  }

 private:
  std::shared_ptr<tbslas_ctx> p_;
};

// Frozen, device-resident leaf list of one host tree.
template <class Tree_t>
class DeviceTree {
 public:
  typedef typename Tree_t::Node_t Node_t;
  explicit DeviceTree(Tree_t *tree, Context ctx = Context::Default()) : ctx_(ctx) { Upload(tree); }

  // (Re)build from the host tree: the leaf walk of tree_functor.h:417-427.  When the leaf
  // set is unchanged only the coefficients travel (what SetTreeGridValues rewrote).
  void Upload(Tree_t *tree) {
    std::vector<Node_t *> &all = tree->GetNodeList();
    std::vector<Node_t *> leaves;
    for (size_t i = 0; i < all.size(); i++)
      if (all[i]->IsLeaf() && !all[i]->IsGhost()) leaves.push_back(all[i]);
    const size_t L = leaves.size();
    int q = q_, dof = dof_;
    if (L) {
      q = leaves[0]->ChebDeg();
      dof = leaves[0]->DataDOF();
    }
    const size_t nc = (size_t)(q + 1) * (q + 2) * (q + 3) / 6 * dof;
    std::vector<double> coord(3 * L), coeff(nc * L);
    std::vector<uint8_t> depth(L);
    for (size_t j = 0; j < L; j++) {
      Node_t *n = leaves[j];
      for (int k = 0; k < 3; k++) coord[3 * j + k] = n->Coord()[k];
      depth[j] = (uint8_t)n->Depth();
      for (size_t k = 0; k < nc; k++) coeff[j * nc + k] = n->ChebData()[k];
    }
    const bool same = h_ && q == q_ && dof == dof_ && coord == coord_ && depth == depth_;
    if (same) {
      ctx_.check(tbslas_b200_tree_update_coeff(h_.get(), coeff.data(), TBSLAS_MEM_HOST));
      return;
    }
    tbslas_tree *t = nullptr;
    ctx_.check(tbslas_b200_tree_create(ctx_.get(), q, dof, L, coord.data(), depth.data(), coeff.data(),
                                       TBSLAS_MEM_HOST, &t));
    Context keep = ctx_;  // the context must outlive its trees
    h_.reset(t, [keep](tbslas_tree *x) { tbslas_b200_tree_destroy(x); });
    q_ = q;
    dof_ = dof;
    coord_.swap(coord);
    depth_.swap(depth);
  }
  tbslas_tree *get() const { return h_.get(); }
  int dof() const { return dof_; }
  int cheb_deg() const { return q_; }
  size_t n_leaf() const { return depth_.size(); }
  const Context &context() const { return ctx_; }

 private:
  Context ctx_;
  std::shared_ptr<tbslas_tree> h_;
  int q_ = 0, dof_ = 0;
  std::vector<double> coord_;
  std::vector<uint8_t> depth_;
};

namespace detail {
template <class Real_t>
struct FunctorBase {
  static_assert(std::is_same<Real_t, double>::value, "tbslas_b200 evaluates in FP64: Real_t must be double");
  tbslas_field field;
  Context ctx;
  FunctorBase() : ctx(Context::Default()) { std::memset(&field, 0, sizeof(field)); }
  // pos is const in the reference's signature but wrapped in place when periodic
  // (const_cast at tree_functor.h:803); same here.
  void operator()(const Real_t *points_pos, int num_points, Real_t *out) { (*this)(points_pos, num_points, 0, out); }
  void operator()(const Real_t *points_pos, int num_points, Real_t time, Real_t *out) {
    ctx.check(tbslas_b200_eval_field(&field, time, CurrentBC(), const_cast<Real_t *>(points_pos),
                                     (size_t)num_points, out, TBSLAS_MEM_HOST));
  }
};
}  // namespace detail

// The functor classes live in their own namespace (re-exported into tbslas::b200 below) so
// that argument-dependent lookup from inside the reference's templates -- which call
// ComputeTrajRK2(...) unqualified, semilag.inc:40 -- never sees the fused overloads at the
// end of this file: the reference's code paths stay exactly the reference's.
namespace functors {

// tbslas::NodeFieldFunctor (tree_functor.h:793-815).  Construction uploads the tree's
// leaves; call update() after the host tree changed (refinement, new coefficients).
template <class Real_t, class Tree_t>
class NodeFieldFunctor : public detail::FunctorBase<Real_t> {
 public:
  explicit NodeFieldFunctor(Tree_t *tree) : host_(tree), dev_(std::make_shared<DeviceTree<Tree_t> >(tree)) { bind(); }
  void update() {
    dev_->Upload(host_);
    bind();
  }
  void update(Tree_t *tree) {
    host_ = tree;
    update();
  }
  DeviceTree<Tree_t> &device_tree() { return *dev_; }
  Tree_t *host_tree() const { return host_; }

 private:
  void bind() {
    this->field.kind = TBSLAS_FIELD_STEADY;
    this->field.tree[0] = dev_->get();
  }
  Tree_t *host_;
  std::shared_ptr<DeviceTree<Tree_t> > dev_;
};

// tbslas::FieldSetFunctor (tree_set_functor.h:27-97): four snapshots, cubic in time.
template <class Real_t, class Tree_t>
class FieldSetFunctor : public detail::FunctorBase<Real_t> {
 public:
  FieldSetFunctor(std::vector<Tree_t *> field_set_elems, std::vector<Real_t> field_set_times)
      : host_(field_set_elems), times_(field_set_times) {
    if (host_.size() != 4 || times_.size() != 4) throw Error(TBSLAS_ERR_INVALID, "FieldSetFunctor needs 4 trees");
    for (int i = 0; i < 4; i++) dev_.push_back(std::make_shared<DeviceTree<Tree_t> >(host_[i]));
    bind();
  }
  // Slide the window (tree_set_functor.h:81-90).  Like the reference, the oldest host
  // tree is deleted.
  void update(Tree_t *new_tree, Real_t time) {
    delete host_[0];
    host_.erase(host_.begin());
    dev_.erase(dev_.begin());
    times_.erase(times_.begin());
    host_.push_back(new_tree);
    dev_.push_back(std::make_shared<DeviceTree<Tree_t> >(new_tree));
    times_.push_back(time);
    bind();
  }

 private:
  void bind() {
    this->field.kind = TBSLAS_FIELD_SET4;
    for (int i = 0; i < 4; i++) {
      this->field.tree[i] = dev_[i]->get();
      this->field.times[i] = times_[i];
    }
  }
  std::vector<Tree_t *> host_;
  std::vector<Real_t> times_;
  std::vector<std::shared_ptr<DeviceTree<Tree_t> > > dev_;
};

// tbslas::FieldExtrapFunctor (tree_extrap_functor.h:27-92): 1.5 v(tc) - 0.5 v(tp).
template <class Real_t, class Tree_t>
class FieldExtrapFunctor : public detail::FunctorBase<Real_t> {
 public:
  FieldExtrapFunctor(Tree_t *tp, Tree_t *tc)
      : tp_(tp), tc_(tc), dp_(std::make_shared<DeviceTree<Tree_t> >(tp)), dc_(std::make_shared<DeviceTree<Tree_t> >(tc)) {
    bind();
  }
  void update(Tree_t *new_tree, Real_t /*time*/) {  // tree_extrap_functor.h:80-85
    delete tp_;
    tp_ = tc_;
    dp_ = dc_;
    tc_ = new_tree;
    dc_ = std::make_shared<DeviceTree<Tree_t> >(new_tree);
    bind();
  }

 private:
  void bind() {
    this->field.kind = TBSLAS_FIELD_EXTRAP;
    this->field.tree[0] = dp_->get();
    this->field.tree[1] = dc_->get();
  }
  Tree_t *tp_, *tc_;
  std::shared_ptr<DeviceTree<Tree_t> > dp_, dc_;
};

}  // namespace functors
using functors::FieldExtrapFunctor;
using functors::FieldSetFunctor;
using functors::NodeFieldFunctor;

// ---------------------------------------------------------------------------------
// Fused entry points: the reference's signatures (traj.h:31-44, semilag.h:21-34,
// tree_semilag.h:92-95), one C-ABI call per step.
// ---------------------------------------------------------------------------------
template <class real_t, class Functor>
void ComputeTrajRK2(Functor &field_fn, const std::vector<real_t> &xinit, const real_t tinit,
                    const real_t tfinal, const int num_rk_step, std::vector<real_t> &xsol) {
  xsol.resize(xinit.size());
  field_fn.ctx.check(tbslas_b200_traj_rk2(&field_fn.field, nullptr, CurrentBC(), xinit.data(), xinit.size() / 3,
                                          tinit, tfinal, num_rk_step, xsol.data(), TBSLAS_MEM_HOST));
}

template <class real_t, class Functor, class ExtrapFunctor>
void ComputeTrajRK2(Functor &field_fn, ExtrapFunctor &extrap_fn, const std::vector<real_t> &xinit,
                    const real_t tinit, const real_t tfinal, const int num_rk_step,
                    std::vector<real_t> &xsol) {
  xsol.resize(xinit.size());
  field_fn.ctx.check(tbslas_b200_traj_rk2(&field_fn.field, &extrap_fn.field, CurrentBC(), xinit.data(),
                                          xinit.size() / 3, tinit, tfinal, num_rk_step, xsol.data(),
                                          TBSLAS_MEM_HOST));
}

template <class real_t, class VFunctor, class Tree_t>
void SolveSemilagRK2(VFunctor &vel_evaluator, NodeFieldFunctor<real_t, Tree_t> &con_evaluator,
                     const std::vector<real_t> &points_pos, const int sdim, const int timestep,
                     const real_t dt, const int num_rk_step, std::vector<real_t> &points_vals) {
  const size_t n = points_pos.size() / sdim;
  points_vals.resize(n * con_evaluator.device_tree().dof());
  vel_evaluator.ctx.check(tbslas_b200_semilag_rk2(&vel_evaluator.field, nullptr, con_evaluator.device_tree().get(),
                                                  CurrentBC(), points_pos.data(), n, timestep, dt, num_rk_step,
                                                  points_vals.data(), nullptr, TBSLAS_MEM_HOST));
}

template <class real_t, class VFunctor, class EFunctor, class Tree_t>
void SolveSemilagRK2(VFunctor &vel_evaluator, EFunctor &extrap_evaluator,
                     NodeFieldFunctor<real_t, Tree_t> &con_evaluator, const std::vector<real_t> &points_pos,
                     const int sdim, const int timestep, const real_t dt, const int num_rk_step,
                     std::vector<real_t> &points_vals) {
  const size_t n = points_pos.size() / sdim;
  points_vals.resize(n * con_evaluator.device_tree().dof());
  vel_evaluator.ctx.check(tbslas_b200_semilag_rk2(&vel_evaluator.field, &extrap_evaluator.field,
                                                  con_evaluator.device_tree().get(), CurrentBC(), points_pos.data(),
                                                  n, timestep, dt, num_rk_step, points_vals.data(), nullptr,
                                                  TBSLAS_MEM_HOST));
}

#ifdef SRC_TREE_UTILS_TREE_H_  // the reference's tree/tree_utils.h is in this translation unit
// tbslas::SolveSemilagInSitu (tree_semilag.h:92-135) entirely on the GPU: the arrival points
// are generated in HBM from the uploaded leaf list (CollectChebTreeGridPoints,
// tree_utils.h:442-498), advected, and refitted with the reference's own point-to-coefficient
// matrix (GetPt2CoeffMatrix, cheb.h:166-196, uploaded once per degree) by one tensor-core
// GEMM (SetTreeGridValues, tree_utils.h:500-552); only the new coefficients come back.
namespace detail {
// steps (1)-(3) of tbslas::SolveSemilagInSitu on the device; f2 == nullptr: one-functor form
template <class TreeType>
void SemilagInSituImpl(const tbslas_field *f1, const tbslas_field *f2, TreeType &tree_curr, const int timestep,
                       const typename TreeType::Real_t dt, int num_rk_step) {
  typedef typename TreeType::Real_t RealType;
  typedef typename TreeType::Node_t NodeType;
  NodeFieldFunctor<RealType, TreeType> con(&tree_curr);
  DeviceTree<TreeType> &dcon = con.device_tree();
  const Context &ctx = dcon.context();
  const int q = dcon.cheb_deg(), dof = dcon.dof();
  int has = 0;  // per context, not per process: a second Context needs its own copy of the matrix
  ctx.check(tbslas_b200_has_pt2coeff(ctx.get(), q, &has));
  if (!has) {
    pvfmm::Matrix<RealType> M;
    tbslas::GetPt2CoeffMatrix<RealType>(q, M);
    ctx.check(tbslas_b200_set_pt2coeff(ctx.get(), q, &M[0][0]));
  }
  ctx.check(tbslas_b200_semilag_insitu_update(f1, f2, dcon.get(), CurrentBC(), timestep, dt, num_rk_step));
  const size_t nc = (size_t)(q + 1) * (q + 2) * (q + 3) / 6 * dof;
  std::vector<RealType> coeff(nc * dcon.n_leaf());
  ctx.check(tbslas_b200_tree_get_coeff(dcon.get(), coeff.data(), TBSLAS_MEM_HOST));
  std::vector<NodeType *> &all = tree_curr.GetNodeList();
  size_t j = 0;
  for (size_t i = 0; i < all.size(); i++)
    if (all[i]->IsLeaf() && !all[i]->IsGhost()) {
      std::memcpy(&(all[i]->ChebData()[0]), &coeff[j * nc], nc * sizeof(RealType));
      j++;
    }
}
}  // namespace detail

// NOTE on dof > 1: the values of a step are point-major [point][dof] (SolveSemilagRK2's layout) and the
// refit here reads them as such.  The reference's own SolveSemilagInSitu hands that array to
// SetTreeGridValues, which reads [dof][P] per leaf (tree_semilag.h:124-133 vs tree_utils.h:528-547) --
// right for dof = 1 only; its dof-3 caller (tree_ns.h:502-513) transposes by hand first.  So for dof > 1
// this overload returns what the NS call pattern returns, not what the reference's in-situ template does.
template <class TreeType, class TreeFunc>
void SolveSemilagInSitu(TreeFunc &tvel_func, TreeType &tree_curr, const int timestep,
                        const typename TreeType::Real_t dt, int num_rk_step = 1, bool /*adaptive*/ = true) {
  detail::SemilagInSituImpl(&tvel_func.field, (const tbslas_field *)nullptr, tree_curr, timestep, dt, num_rk_step);
}

// The two-functor form (tree_semilag.h:137-181; drivers advtvextrap.cpp, ns.cpp): stage 1 of every RK2
// sub-step samples tvel_func, stage 2 the extrapolation tvel_extrap (traj.inc:71-92).
template <class TreeType, class TreeFunc, class TreeExtrap>
void SolveSemilagInSitu(TreeFunc &tvel_func, TreeExtrap &tvel_extrap, TreeType &tree_curr, const int timestep,
                        const typename TreeType::Real_t dt, int num_rk_step = 1, bool /*adaptive*/ = true) {
  detail::SemilagInSituImpl(&tvel_func.field, &tvel_extrap.field, tree_curr, timestep, dt, num_rk_step);
}
#endif

}  // namespace b200
}  // namespace tbslas

#endif  // TBSLAS_B200_FUNCTORS_HPP_

/* tbslas_b200.h -- C ABI of the B200-native semi-Lagrangian hot path.
 *
 * Drop-in boundary for arashb/tbslas's src/semilag + src/tree advection path
 * (reference @ 0e66711).  The reference's "operator API" is a C++ functor concept,
 *     void f(const real_t* pos_aos, int n, real_t* out_aos);
 *     void f(const real_t* pos_aos, int n, real_t t, real_t* out_aos);
 * (tree_functor.h:800-811, tree_set_functor.h:49-50, tree_extrap_functor.h:47)
 * consumed by IntegrateRK2 / ComputeTrajRK2 / SolveSemilagRK2 (traj.h:25-44,
 * semilag.h:21-34).  This header is what an FFI for that path binds: plain
 * pointers and sizes, opaque handles, int status (0 = ok), no exceptions, no
 * torch types.  The header-only C++ adaptors in include/tbslas_b200/ wrap these
 * entry points back into the functor concept so the reference's templates and
 * drivers compile against them unchanged.
 *
 * Conventions
 *   - points are AoS [n][3] doubles, values AoS [n][dof] doubles, exactly as the
 *     reference passes them; `mem` says whether the caller's buffers live in host
 *     memory (copied in/out on the context's stream) or are device pointers.
 *   - bc: TBSLAS_FREESPACE / TBSLAS_PERIODIC == pvfmm::BoundaryType, which the
 *     reference reads implicitly from its SimConfig singleton
 *     (tree_functor.h:174,469,803); here it is an explicit argument.
 *   - when bc is periodic the position buffer handed to an eval call is wrapped IN
 *     PLACE ((c<0)->c+1, (c>=1)->c-1, once), as the reference does through a
 *     const_cast (tree_functor.h:442-449,803).
 *   - all work is stream ordered on the context's stream; calls with host buffers
 *     return after the results have landed, calls with device buffers return after
 *     enqueueing (use tbslas_b200_synchronize or your own stream sync).
 *   - one context per GPU per process; a context and its trees are not thread safe
 *     (neither is the reference: function-static scratch, tree_functor.h:166,519).
 *   - there is NO CPU fallback: every entry point fails with TBSLAS_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef TBSLAS_B200_H_
#define TBSLAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tbslas_ctx tbslas_ctx;
typedef struct tbslas_tree tbslas_tree;

enum {
  TBSLAS_OK = 0,
  TBSLAS_ERR_INVALID = 1,     /* bad argument */
  TBSLAS_ERR_CUDA = 2,        /* CUDA runtime / no usable device */
  TBSLAS_ERR_COMM = 3,        /* NCCL */
  TBSLAS_ERR_UNSUPPORTED = 4, /* e.g. Chebyshev degree out of range */
  TBSLAS_ERR_NOMEM = 5
};
enum { TBSLAS_FREESPACE = 0, TBSLAS_PERIODIC = 1 };
enum { TBSLAS_MEM_HOST = 0, TBSLAS_MEM_DEVICE = 1 };

/* Velocity "functor" kinds (what the reference passes as FieldFunctor):
 *   STEADY  tree[0]                      tbslas::NodeFieldFunctor    tree_functor.h:793-815
 *   SET4    tree[0..3] at times[0..3]    tbslas::FieldSetFunctor     tree_set_functor.h:27-97
 *           (cubic Hermite in time, utils/cubic.h:42-56)
 *   EXTRAP  tree[0]=t^{n-1}, tree[1]=t^n tbslas::FieldExtrapFunctor  tree_extrap_functor.h:27-92
 *           (1.5*v(t^n) - 0.5*v(t^{n-1})) */
enum { TBSLAS_FIELD_STEADY = 0, TBSLAS_FIELD_SET4 = 1, TBSLAS_FIELD_EXTRAP = 2 };
typedef struct tbslas_field {
  int kind;
  tbslas_tree *tree[4];
  double times[4];
} tbslas_field;

#define TBSLAS_MAX_CHEB_DEG 19 /* reference asserts deg < 20 (cheb.h:43); scripts use <= 14 */

/* ---- context ------------------------------------------------------------ */
/* Replaces: process start-up of a reference driver (MPI_Init + SimConfig,
 * advection.cpp:50-90).  `device` is the CUDA ordinal this process drives. */
int tbslas_b200_init(int device, tbslas_ctx **ctx);
int tbslas_b200_finalize(tbslas_ctx *ctx);
/* Run on an externally owned cudaStream_t (e.g. the caller's framework stream);
 * NULL restores the context's own stream; the legacy default stream is the CUDA handle
 * cudaStreamLegacy ((void*)1), not NULL. */
int tbslas_b200_set_stream(tbslas_ctx *ctx, void *cuda_stream);
int tbslas_b200_synchronize(tbslas_ctx *ctx);
/* Last error text of this context (never NULL). */
const char *tbslas_b200_last_error(tbslas_ctx *ctx);
const char *tbslas_b200_version(void);

/* ---- multi-GPU: one process per GPU, Morton-range shards ------------------ */
/* Replaces: sim_config->comm / *tree->Comm() (tree_functor.h:407,570).
 * Rank 0 obtains a 128-byte NCCL unique id and ships it to the other ranks by any
 * means the host has (MPI_Bcast, torch.distributed, a file); every rank then calls
 * comm_init.  Without comm_init the context is single-rank. */
int tbslas_b200_comm_unique_id(void *id128);
int tbslas_b200_comm_init(tbslas_ctx *ctx, int nranks, int rank, const void *id128);
int tbslas_b200_comm_rank(tbslas_ctx *ctx, int *rank, int *nranks);
/* Outsider points this rank sent to / received from other ranks during the most recent
 * tree evaluation (the cnt_outside of tree_functor.h:513-517 and its mirror image). */
int tbslas_b200_comm_last_exchange(tbslas_ctx *ctx, size_t *sent, size_t *received);
/* How the outsiders travel (par::SortScatterIndex / ScatterForward / ScatterReverse,
 * tree_functor.h:569-595).
 *   mode 1 (default where every rank can map every peer's memory -- one NVLink/NVSwitch box):
 *     peer-memory mailboxes.  The pack kernel writes each outsider straight into its owner's
 *     receive buffer over NVLink, the owner evaluates what arrived and writes the values straight
 *     back; counts, offsets and the three barriers of an evaluation live on the device, so an
 *     evaluation costs no host synchronisation and no library kernel sits next to the
 *     evaluation kernel.  A rank can receive (and send) at most `mailbox_points` outsiders per
 *     evaluation (grown at tree_create to a quarter of the largest shard's arrival points;
 *     comm_set_mailbox -- collective -- sets it); beyond that the evaluation fails with
 *     TBSLAS_ERR_COMM at the next synchronising call, on every rank.
 *   mode 0: NCCL all-to-all-v (count matrix by ncclAllGather, grouped ncclSend/ncclRecv both
 *     ways): no capacity limit, one host synchronisation per evaluation.
 * All ranks must select the same mode.  comm_exchange_mode reports what is in effect. */
int tbslas_b200_comm_set_exchange(tbslas_ctx *ctx, int mode);
int tbslas_b200_comm_set_mailbox(tbslas_ctx *ctx, size_t points);
int tbslas_b200_comm_exchange_mode(tbslas_ctx *ctx, int *mode, size_t *mailbox_points);

/* ---- trees -------------------------------------------------------------- */
/* Replaces: the leaf walk at tree_functor.h:417-427 (GetNodeList filtered by
 * IsLeaf && !IsGhost) and the per-leaf reads of Coord/Depth/ChebData
 * (:249-266,:283-284).  Leaves must be in Morton (PVFMM preorder) order.  In a
 * multi-rank context every rank passes ITS OWN contiguous Morton range (what an MPI
 * rank of the reference owns; a rank may pass n_leaf = 0); the first-leaf keys are
 * all-gathered here (a collective call: every rank must enter), replacing the per-call
 * MPI_Allgather at tree_functor.h:433-437.  Every evaluation on such a tree is then
 * collective as well, exactly like tbslas::EvalTree.
 *   coord  [n_leaf][3]   depth [n_leaf]   coeff [n_leaf][dof][Ncoef],
 *   Ncoef = (q+1)(q+2)(q+3)/6 in the reference's packed order (:256-266). */
int tbslas_b200_tree_create(tbslas_ctx *ctx, int q, int dof, size_t n_leaf,
                            const double *coord, const uint8_t *depth,
                            const double *coeff, int mem, tbslas_tree **tree);
/* The same, for a tree that every rank of a multi-rank context holds IN FULL (the caller
 * passes all leaves on every rank): evaluations on it are local and not collective -- no
 * point exchange.  The reference has no such mode (an MPI rank only ever holds its own
 * range); it is the "replicate instead of exchange" option for trees that are small next
 * to 180 GB of HBM, e.g. a smooth velocity field.  Identical to tree_create in a
 * single-rank context. */
int tbslas_b200_tree_create_replicated(tbslas_ctx *ctx, int q, int dof, size_t n_leaf,
                                       const double *coord, const uint8_t *depth,
                                       const double *coeff, int mem, tbslas_tree **tree);
/* New coefficients on the same leaves (what SetTreeGridValues writes every step,
 * tree_utils.h:547-550). */
int tbslas_b200_tree_update_coeff(tbslas_tree *tree, const double *coeff, int mem);
/* The same without waiting: the copy runs on the context's copy stream, behind the evaluations
 * already enqueued that still read the old coefficients and concurrently with whatever is enqueued
 * next; the first later call that touches this tree's coefficients waits for it ON THE DEVICE.
 * A host buffer must be pinned for the copy to overlap and must stay valid until a
 * synchronising call on the context returns (any call with host output buffers, or
 * tbslas_b200_synchronize).  In a semi-Lagrangian step only the last of the three evaluations
 * reads the advected tree, so its upload hides behind the two velocity evaluations. */
int tbslas_b200_tree_update_coeff_async(tbslas_tree *tree, const double *coeff, int mem);
/* Read the coefficients back, [n_leaf][dof][Ncoef] (e.g. after semilag_insitu_update, before
 * the host refines the tree). */
int tbslas_b200_tree_get_coeff(tbslas_tree *tree, double *coeff, int mem);
int tbslas_b200_tree_destroy(tbslas_tree *tree);
int tbslas_b200_tree_info(const tbslas_tree *tree, int *q, int *dof, size_t *n_leaf);

/* ---- evaluation (tbslas::EvalTree, tree_functor.h:397-690) ---------------- */
/* out[n][dof] = field(pos[n]).  leaf_idx (optional, same memory space as pos) gets
 * the index of the leaf that evaluated each point, counted in the GLOBAL Morton
 * order (-1: no leaf claims the point; its value is 0).  Out-of-domain points under
 * FREESPACE evaluate to 0 (cheb_poly is zero outside [-1,1]). */
int tbslas_b200_eval(tbslas_tree *tree, int bc, double *pos, size_t n, double *out,
                     int32_t *leaf_idx, int mem);
/* FieldSetFunctor::operator() (tree_set_functor.h:49-79). */
int tbslas_b200_eval_set4(tbslas_tree *const trees[4], const double times[4], double t,
                          int bc, double *pos, size_t n, double *out, int mem);
/* FieldExtrapFunctor::operator() (tree_extrap_functor.h:47-78). */
int tbslas_b200_eval_extrap(tbslas_tree *tp, tbslas_tree *tc, int bc, double *pos,
                            size_t n, double *out, int mem);
/* How SET4 / EXTRAP fields are evaluated when their trees share one leaf list (the usual case:
 * snapshots of one adaptive tree).  mode 1 (default): the trees' coefficients are combined in
 * time on the device and the field is evaluated ONCE -- sum_k w_k tree_k(x) is linear in the
 * coefficients -- so 4 (2) tree evaluations, their point locations and, across ranks, their
 * point exchanges become one; values agree with the reference's order of operations to
 * rounding (~1e-15 of the field scale).  mode 0: every tree is evaluated and the values are
 * combined per point exactly as tree_set_functor.h:55-72 / tree_extrap_functor.h:59-77 do.
 * Trees with different leaf lists always take the mode-0 route. */
int tbslas_b200_set_time_combine(tbslas_ctx *ctx, int mode);
/* The weights of that combination, host only: InterpCubic1D (cubic.h:36-56) is linear in the four
 * snapshot values, out = w[0] p0 + w[1] p1 + w[2] p2 + w[3] p3. */
int tbslas_b200_cubic_time_weights(const double times[4], double t, double w[4]);
/* Any field kind through one entry point (t is ignored by STEADY and EXTRAP). */
int tbslas_b200_eval_field(const tbslas_field *f, double t, int bc, double *pos, size_t n,
                           double *out, int mem);

/* ---- trajectories and the semi-Lagrangian step ---------------------------- */
/* tbslas::ComputeTrajRK2 (traj.inc:49-68; two-functor form :95-115): nrk explicit-
 * midpoint sub-steps from tinit to tfinal.  f2 == NULL: both stages sample f1 (second
 * at t + tau/2); otherwise stage 1 samples f1 and stage 2 samples f2
 * (tree_ns.h:471-483).  pos is not modified; out_pos[n][3]. */
int tbslas_b200_traj_rk2(const tbslas_field *f1, const tbslas_field *f2, int bc,
                         const double *pos, size_t n, double tinit, double tfinal,
                         int nrk, double *out_pos, int mem);
/* tbslas::SolveSemilagRK2 (semilag.inc:27-45, :49-69): departure points over
 * [timestep*dt, timestep*dt - dt], then `con` sampled there.  out_vals[n][dof_con];
 * out_dep (optional) receives the departure points [n][3]. */
int tbslas_b200_semilag_rk2(const tbslas_field *f1, const tbslas_field *f2,
                            tbslas_tree *con, int bc, const double *pos, size_t n,
                            int timestep, double dt, int nrk, double *out_vals,
                            double *out_dep, int mem);

/* Tree-level calls (semilag_insitu, semilag_insitu_update) know that the arrival points are the
 * (q+1)^3 tensor grids of con's leaves.  mode 1 (default): where such a leaf lies inside one leaf
 * of a dof-3 velocity tree held on this rank, the FIRST velocity evaluation of the step runs by
 * sum factorisation (three 1-D passes per leaf: 88 k instead of 2.75 M FMA per leaf and component
 * at q = 14) -- the same polynomial at the same points summed in another order, ~1e-15 of the field
 * scale; grid points on a velocity-leaf face, and leaves not inside one velocity leaf, take the
 * generic path, so every point is evaluated by the leaf the reference assigns it to.  Calls on
 * fewer than 4 Mi points keep the generic path (latency bound either way); mode 2 removes that
 * minimum (tests).  mode 0: every evaluation is point by point. */
int tbslas_b200_set_tensor_grid(tbslas_ctx *ctx, int mode);
/* Tree-level calls whose first velocity evaluation took the sum-factorised path need the arrival
 * points once more, as the base of the RK2 update x' = x + tau*v(x_mid) (traj.inc:42).  on = 1:
 * they are rebuilt there from (leaf geometry, node index) with the expressions that
 * generate them -- the same bits -- instead of being written to and read back from HBM (48 B per
 * point of traffic).  on = 0 (default): always materialised -- on C2 the saving in the first stage
 * (7.6 -> 6.6 ms) is smaller than what the index decode costs the second evaluation (+5.2 ms). */
int tbslas_b200_set_virtual_arrival_points(tbslas_ctx *ctx, int on);
/* Calls with HOST buffers are cut into chunks whose copy-in, kernels and copy-out overlap on three
 * streams.  chunks = 0 (default): chosen from the bytes that cross PCIe (about 16 Mi points per chunk
 * for tree-level calls, 4 Mi with host input, at most 16); > 0: exactly that many.  In a multi-rank
 * context every rank must use the same setting (every chunk is a collective evaluation). */
int tbslas_b200_set_host_chunks(tbslas_ctx *ctx, int chunks);
/* Arrival points of the most recent tree-level call (or of its last chunk, for host buffers) that
 * took the generic path instead (on a velocity-leaf face, or in a leaf not inside one velocity leaf). */
int tbslas_b200_last_grid_exceptions(tbslas_ctx *ctx, size_t *n);
/* Steps (1)+(2) of tbslas::SolveSemilagInSitu (tree_semilag.h:92-130): the arrival points
 * are the Chebyshev grid points of `con`'s own (local) leaves, generated in HBM
 * (CollectChebTreeGridPoints, tree_utils.h:442-498) -- no 24 B/point host->device copy --
 * then semilag_rk2.  out_vals[n_leaf*(q+1)^3][dof_con], leaf-major, ready for
 * SetTreeGridValues (tree_utils.h:500-552). */
int tbslas_b200_semilag_insitu(const tbslas_field *f1, const tbslas_field *f2,
                               tbslas_tree *con, int bc, int timestep, double dt, int nrk,
                               double *out_vals, int mem);

/* The same, also returning the departure points [n_leaf*(q+1)^3][3] as ComputeTrajRK2 left them
 * (before the scalar evaluation wraps them): what the parity tests compare stage by stage. */
int tbslas_b200_semilag_insitu_dep(const tbslas_field *f1, const tbslas_field *f2,
                                   tbslas_tree *con, int bc, int timestep, double dt, int nrk,
                                   double *out_vals, double *out_dep, int mem);

/* ---- values -> coefficients: tbslas::SetTreeGridValues (tree_utils.h:500-552) ------ */
/* The point-to-coefficient matrix of degree q, M[(q+1)^3][Ncoef] row-major (host memory):
 * what tbslas::GetPt2CoeffMatrix builds (cheb.h:166-196: pseudo-inverse of the basis matrix
 * at the new_nodes grid).  Supplied by the caller once per degree so that host and device
 * refits use the very same matrix; kept on the device by the context. */
int tbslas_b200_set_pt2coeff(tbslas_ctx *ctx, int q, const double *M);
/* Whether THIS context already holds the matrix of degree q (the C++ adaptor uploads it on first use). */
int tbslas_b200_has_pt2coeff(tbslas_ctx *ctx, int q, int *has);
/* coeff[leaf][dof][:] = vals[leaf][dof][:] * M  for every local leaf, written into the tree
 * (one FP64 tensor-core GEMM).  point_major = 0: vals is [leaf][dof][P], the layout
 * SetTreeGridValues consumes; 1: vals is [leaf*P][dof], the layout SolveSemilagRK2 produces
 * (tree_ns.h:502-513 transposes between the two). */
int tbslas_b200_tree_set_grid_values(tbslas_tree *tree, const double *vals, int point_major,
                                     int mem);
/* The whole tbslas::SolveSemilagInSitu (tree_semilag.h:92-135) on the device: arrival points
 * generated in HBM, advected, and `con`'s coefficients refitted in place -- nothing crosses
 * PCIe.  Needs set_pt2coeff for con's degree. */
int tbslas_b200_semilag_insitu_update(const tbslas_field *f1, const tbslas_field *f2,
                                      tbslas_tree *con, int bc, int timestep, double dt,
                                      int nrk);

/* ---- uniform-grid cubic variant (tbslas::fast_interp, tree_functor.h:89-153) - */
/* grid [dof][n_reg][n_reg][n_reg] (x fastest), node centred on [0,1]^3. */
int tbslas_b200_cubic_eval(tbslas_ctx *ctx, const double *grid, int n_reg, int dof,
                           const double *pos, size_t n, double *out, int mem);
/* The same with the grid RESIDENT in HBM: the reference's caller samples a field onto the regular
 * grid once and interpolates from it many times (tree_functor.h:89-153 takes the grid by pointer);
 * here the upload (403 MB at 256^3 x dof 3) is paid once per grid, not once per call.  grid_eval with
 * host buffers streams the queries in and the values out in overlapping chunks. */
typedef struct tbslas_grid tbslas_grid;
int tbslas_b200_grid_create(tbslas_ctx *ctx, const double *grid, int n_reg, int dof, int mem,
                            tbslas_grid **out);
int tbslas_b200_grid_update(tbslas_grid *g, const double *grid, int mem);
int tbslas_b200_grid_eval(tbslas_grid *g, const double *pos, size_t n, double *out, int mem);
int tbslas_b200_grid_destroy(tbslas_grid *g);

/* ---- either side of the path ("next" rows) -------------------------------- */
/* tbslas::CollectChebTreeGridPoints (tree_utils.h:442-498): arrival points of the
 * local leaves, leaf-major, [n_leaf*(q+1)^3][3]. */
int tbslas_b200_collect_grid_points(tbslas_tree *tree, double *out_pos, int mem);
/* tbslas::new_nodes (cheb.h:41-68), 1-D nodes, host only: out[q+1]. */
int tbslas_b200_new_nodes(int q, double *out);

/* ---- host-side shard logic (no GPU needed) -------------------------------- */
/* 48-bit Morton key of a query point as EvalTree builds it (tree_functor.h:464-479):
 * z-major interleave of floor(c * 2^15), a coordinate == 1.0 shifted by 2^-15 unless
 * periodic; UINT64_MAX when any anchor leaves the 15-bit range. */
uint64_t tbslas_b200_point_key(double x, double y, double z, int bc);
/* Owner rank of a key given the first-leaf key of every rank (tree_functor.h:491-513
 * + the split-key rule of par::SortScatterIndex, :569): last r with splitter[r] <= key
 * (rank 0 when key < splitter[0]). */
int tbslas_b200_owner_of_key(uint64_t key, const uint64_t *splitters, int nranks);
/* Equal-count Morton-range partition of n_leaf leaves over nranks:
 * first[r] = r*n_leaf/nranks, first[nranks] = n_leaf. */
int tbslas_b200_partition_leaves(size_t n_leaf, int nranks, size_t *first);

/* ---- load balance and refinement hints --------------------------- */
/* Contiguous Morton ranges of about equal total WEIGHT (e.g. last step's points per leaf):
 * what tbslas::SemiMergeTree aims at with its point-count aware break points
 * (tree_utils.h:672-675).  first[r] = first leaf of rank r, first[nranks] = n_leaf.  Host only. */
int tbslas_b200_partition_leaves_weighted(size_t n_leaf, const double *weight, int nranks,
                                          size_t *first);
/* Moves the leaves of a Morton-sharded tree between the ranks -- keys, geometry, integer boxes and
 * coefficient blocks, device to device over NCCL -- so that rank r owns the global leaves
 * [new_first[r], new_first[r+1]) (new_first[0] = 0, new_first[nranks] = total leaves), and rebuilds the
 * split keys: what tbslas::SemiMergeTree does through PVFMM's RedistNodes when it re-balances by last
 * step's point counts (tree_utils.h:609-729).  Collective; every rank passes the same array.  Results of
 * evaluations are unchanged bit for bit (a point is evaluated by the leaf that contains it, wherever
 * that leaf lives).  Typical use: weights = tree_last_point_counts gathered over the ranks ->
 * partition_leaves_weighted -> tree_reshard for the advected tree, and again with the same split keys
 * for the trees evaluated with it. */
int tbslas_b200_tree_reshard(tbslas_tree *tree, const size_t *new_first);
/* Global index of this rank's first leaf and the number of leaves over all ranks. */
int tbslas_b200_tree_global_range(const tbslas_tree *tree, size_t *first, size_t *total);
/* Points that the most recent evaluation of `tree` located in each of its LOCAL leaves -- this
 * rank's own points plus those received from other ranks (the part_indx differences of
 * tree_functor.h:190-198), [n_leaf]: the weights above.  TBSLAS_ERR_INVALID before the first
 * evaluation. */
int tbslas_b200_tree_last_point_counts(tbslas_tree *tree, uint32_t *counts, int mem);
/* l2 norm of every local leaf's highest-degree coefficients (i+j+k == q, all dof), [n_leaf]:
 * the input of a host-side refine/coarsen decision (the reference leaves that decision to
 * PVFMM's RefineTree, tree_utils.h:117-118) without downloading the coefficients. */
int tbslas_b200_tree_tail_norm(tbslas_tree *tree, double *tail, int mem);

/* ---- instrumentation (pvfmm::Profile::Tic/Toc tags, tree_functor.h:463-674) -- */
/* When enabled, every kernel stage is bracketed with CUDA events on the context's
 * stream.  `get` synchronises and returns accumulated milliseconds and launch counts
 * per stage since the last reset.  Stage names: see tbslas_b200_profile_stage_name. */
int tbslas_b200_profile_enable(tbslas_ctx *ctx, int on);
int tbslas_b200_profile_reset(tbslas_ctx *ctx);
int tbslas_b200_profile_num_stages(void);
const char *tbslas_b200_profile_stage_name(int stage);
/* The pvfmm::Profile::Tic tag(s) of the reference's EvalTree that the stage stands for ("" if the
 * reference has none), e.g. Locate -> "LclHQSort" (tree_functor.h:463), ChebEval ->
 * "InEvaluation/OutEvaluation" (:674, :585), Exchange -> "OutScatterForward/OutScatterReverse"
 * (:573, :593).  Every stage is also an NVTX range named "<stage> (<tag>)". */
const char *tbslas_b200_profile_reference_tag(int stage);
int tbslas_b200_profile_get(tbslas_ctx *ctx, int stage, double *ms, long long *launches,
                            double *units);
/* Total kernels launched by this context since init (the bench's gpu_launches). */
long long tbslas_b200_kernel_launches(tbslas_ctx *ctx);
/* Measured FP64 FMA peak of this device (a DFMA-only kernel, best of `reps`),
 * the denominator for the evaluation kernel's roofline. */
int tbslas_b200_fp64_peak(tbslas_ctx *ctx, int reps, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* TBSLAS_B200_H_ */

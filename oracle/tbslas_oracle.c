/* oracle/tbslas_oracle.c -- CPU restatement of the tbslas semi-Lagrangian hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tbslas_b200/ or include/ may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline leg do, and only as the checker.
 *
 * Plain C (gcc -O2 -ffp-contract=off -fopenmp).  Every function cites the
 * reference (arashb/tbslas @ 0e66711) file:line it restates.  Floating-point
 * operations are written in the reference's order with separate multiply and
 * add (the reference's AVX path uses mul/add intrinsics, tree_functor.h:63-76),
 * so on the same inputs this file is expected to be BIT-IDENTICAL to the
 * reference's own code compiled over the PVFMM stand-in (oracle/_ref); the test
 * suite checks exactly that (tests/test_oracle_vs_ref.py) and freezes outputs
 * of the reference build as golden fixtures (tests/golden/).
 *
 * Pinning status: the reference ships no golden vectors or unit tests
 * (SURVEY.md section 4).  This oracle is pinned to the reference's own code for
 * everything that lives in /root/reference.  Two primitives live in the
 * un-vendored, un-pinned PVFMM dependency (github dmalhotra/pvfmm) and are
 * restated from its published algorithm: MortonId (integer anchor at depth 15,
 * z-major comparator) and cheb_poly (three-term recurrence, zero outside
 * [-1,1]); their call sites are tree_functor.h:173-184,196,328-330,467-479.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_DEPTH 15 /* pvfmm MAX_DEPTH; cf. sim_config.h:40 */
#define ORC_MAX_Q 31

typedef struct {
  int q, dof;
  long n_leaf;
  long ncoef;       /* (q+1)(q+2)(q+3)/6 */
  uint64_t *key;    /* 48-bit interleaved anchor of each leaf (Morton order) */
  double *coord;    /* [n_leaf][3] lower corner */
  uint8_t *depth;   /* [n_leaf] */
  double *coeff;    /* [n_leaf][dof][ncoef], reference order (a14) */
} orc_tree;

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- Morton keys ---------------------------------------------------------
 * pvfmm::MortonId(x,y,z): per axis (uint32)floor(x * 2^15); ordering = Z-curve
 * with z most significant, then y, then x.  That order
 * equals unsigned comparison of the bit-interleaved anchor.  The node depth only
 * breaks ties between equal anchors; point keys carry depth 15 >= any leaf depth,
 * so "leaf <= point" never depends on it and the depth is left out of the key.
 * Anchors that do not fit 15 bits (coordinate >= 1, or negative: the reference's
 * float->unsigned conversion wraps to a huge value on x86) compare greater than
 * every in-domain key, so they are saturated to UINT64_MAX. */
static uint64_t spread3(uint32_t v) { /* bit b of v -> bit 3b */
  uint64_t x = v & 0x1fffffu;
  x = (x | (x << 32)) & 0x001f00000000ffffull;
  x = (x | (x << 16)) & 0x001f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

uint64_t orc_anchor_key(uint32_t ix, uint32_t iy, uint32_t iz) {
  if ((ix | iy | iz) >> ORC_MAX_DEPTH) return UINT64_MAX;
  return spread3(ix) | (spread3(iy) << 1) | (spread3(iz) << 2);
}

static uint32_t anchor1(double x) { /* (UINT_T)floor(x*2^15) with x86 wrap */
  double f = floor(x * 32768.0);
  if (!(f > -9.0e18 && f < 9.0e18)) return 0xffffffffu; /* nan/huge: out */
  return (uint32_t)(int64_t)f;
}

/* key of a query point, tree_functor.h:464-479 (and :172-184): a coordinate that
 * is exactly 1.0 is shifted by 2^-15 unless the boundary is periodic. */
uint64_t orc_point_key(double x, double y, double z, int periodic) {
  const double shift = 1.0 / 32768.0;
  if (x == 1.0 && !periodic) x = x - shift;
  if (y == 1.0 && !periodic) y = y - shift;
  if (z == 1.0 && !periodic) z = z - shift;
  return orc_anchor_key(anchor1(x), anchor1(y), anchor1(z));
}

/* ---- tree handle --------------------------------------------------------- */
orc_tree *orc_tree_create(int q, int dof, long n_leaf, const double *coord,
                          const uint8_t *depth, const double *coeff) {
  if (q < 0 || q > ORC_MAX_Q || dof < 1 || n_leaf < 1) return NULL;
  orc_tree *t = (orc_tree *)calloc(1, sizeof(orc_tree));
  t->q = q;
  t->dof = dof;
  t->n_leaf = n_leaf;
  t->ncoef = (long)(q + 1) * (q + 2) * (q + 3) / 6;
  t->key = (uint64_t *)malloc(sizeof(uint64_t) * n_leaf);
  t->coord = (double *)malloc(sizeof(double) * 3 * n_leaf);
  t->depth = (uint8_t *)malloc(n_leaf);
  t->coeff = (double *)malloc(sizeof(double) * n_leaf * dof * t->ncoef);
  memcpy(t->coord, coord, sizeof(double) * 3 * n_leaf);
  memcpy(t->depth, depth, n_leaf);
  memcpy(t->coeff, coeff, sizeof(double) * n_leaf * dof * t->ncoef);
  for (long j = 0; j < n_leaf; j++) /* Cheb_Node::GetMortonId = MortonId(Coord(),Depth()) */
    t->key[j] = orc_anchor_key(anchor1(coord[3 * j]), anchor1(coord[3 * j + 1]),
                               anchor1(coord[3 * j + 2]));
  return t;
}
void orc_tree_destroy(orc_tree *t) {
  if (!t) return;
  free(t->key);
  free(t->coord);
  free(t->depth);
  free(t->coeff);
  free(t);
}
/* returns 1 when leaf keys are strictly increasing (the layout EvalTree assumes,
 * tree_functor.h:417-427: preorder leaf list) */
int orc_tree_is_sorted(const orc_tree *t) {
  for (long j = 1; j < t->n_leaf; j++)
    if (!(t->key[j - 1] < t->key[j])) return 0;
  return 1;
}

/* leaf location, tree_functor.h:190-198: point p belongs to leaf j iff
 * key(leaf j) <= key(p) < key(leaf j+1); the last leaf takes every larger key;
 * -1 when key(p) < key(leaf 0) (the reference never evaluates such a point). */
static long locate(const orc_tree *t, uint64_t k) {
  long lo = 0, hi = t->n_leaf; /* first leaf with key > k */
  while (lo < hi) {
    long mid = (lo + hi) >> 1;
    if (t->key[mid] <= k)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

/* pvfmm::cheb_poly, call site tree_functor.h:328-330: T_0..T_q by the three-term
 * recurrence; all zero when |xi| > 1. */
static void cheb_basis(int q, double xi, double *T) {
  const int in = fabs(xi) <= 1.0;
  const double x = in ? xi : 0.0;
  double y0 = in ? 1.0 : 0.0, y1 = x;
  T[0] = y0;
  if (q >= 1) T[1] = x;
  for (int i = 2; i <= q; i++) {
    double y2 = 2 * x * y1 - y0;
    T[i] = y2;
    y0 = y1;
    y1 = y2;
  }
}

/* one point in one leaf: rescale (tree_functor.h:283-294), basis (:328-330),
 * triangular contraction in vec_eval's order (:32-77), AoS store (:376-383). */
static void eval_point(const orc_tree *t, long j, const double *x, double *out) {
  const int q = t->q, d = q + 1;
  const double *c = t->coord + 3 * j;
  const double s = (double)(1ULL << t->depth[j]);
  double px[ORC_MAX_Q + 1], py[ORC_MAX_Q + 1], pz[ORC_MAX_Q + 1];
  cheb_basis(q, (x[0] - c[0]) * 2.0 * s - 1.0, px);
  cheb_basis(q, (x[1] - c[1]) * 2.0 * s - 1.0, py);
  cheb_basis(q, (x[2] - c[2]) * 2.0 * s - 1.0, pz);
  for (int l = 0; l < t->dof; l++) {
    const double *C = t->coeff + ((size_t)j * t->dof + l) * t->ncoef;
    long idx = 0;
    double u = 0.0;
    for (int i = 0; i < d; i++) {
      double v = 0.0;
      for (int jj = 0; i + jj < d; jj++) {
        double w = 0.0;
        for (int k = 0; i + jj + k < d; k++) {
          w = w + px[k] * C[idx];
          idx++;
        }
        v = v + py[jj] * w;
      }
      u = u + pz[i] * v;
    }
    out[l] = u;
  }
}

/* tbslas::EvalTree for one rank, tree_functor.h:397-690.  pos is wrapped IN PLACE
 * when periodic (:442-449, one conditional add and one conditional subtract, not a
 * modulo).  Sorting (:487) only groups points by leaf, so the restatement works
 * point by point.  A point no leaf claims gets 0 (the reference leaves its output
 * slot untouched).  leaf_idx may be NULL. */
void orc_eval_tree(const orc_tree *t, int periodic, double *pos, long n, double *out,
                   int *leaf_idx) {
#pragma omp parallel for schedule(static)
  for (long p = 0; p < n; p++) {
    double *x = pos + 3 * p;
    if (periodic)
      for (int a = 0; a < 3; a++) {
        if (x[a] < 0.0) x[a] = x[a] + 1.0;
        if (x[a] >= 1.0) x[a] = x[a] - 1.0;
      }
    const long j = locate(t, orc_point_key(x[0], x[1], x[2], periodic));
    if (leaf_idx) leaf_idx[p] = (int)j;
    if (j < 0)
      for (int l = 0; l < t->dof; l++) out[p * t->dof + l] = 0.0;
    else
      eval_point(t, j, x, out + p * t->dof);
  }
}

/* tbslas::CubicInterpPolicy::InterpCubic1D, cubic.h:28-56 (Hermite basis with
 * centred-difference tangents), same expression order. */
double orc_interp_cubic1d(double x, const double *xx, const double *pp) {
  const double mk = (pp[2] - pp[1]) * 0.5 / (xx[2] - xx[1]) +
                    (pp[1] - pp[0]) * 0.5 / (xx[1] - xx[0]);
  const double mk1 = (pp[3] - pp[2]) * 0.5 / (xx[3] - xx[2]) +
                     (pp[2] - pp[1]) * 0.5 / (xx[2] - xx[1]);
  const double t = (x - xx[1]) / (xx[2] - xx[1]);
  const double h00 = 2 * t * t * t - 3 * t * t + 1;
  const double h10 = t * t * t - 2 * t * t + t;
  const double h01 = -2 * t * t * t + 3 * t * t;
  const double h11 = t * t * t - t * t;
  return h00 * pp[1] + h10 * (xx[2] - xx[1]) * mk + h01 * pp[2] +
         h11 * (xx[2] - xx[1]) * mk1;
}

/* velocity "functor": 1 tree (NodeFieldFunctor, tree_functor.h:800-811), 4 trees +
 * cubic in time (FieldSetFunctor, tree_set_functor.h:49-79) or 2 trees
 * extrapolated 1.5 c - 0.5 p (FieldExtrapFunctor, tree_extrap_functor.h:47-78). */
enum { ORC_VEL_STEADY = 0, ORC_VEL_SET4 = 1, ORC_VEL_EXTRAP = 2 };

typedef struct {
  int kind;
  const orc_tree *tree[4]; /* STEADY: [0]; SET4: [0..3]; EXTRAP: [0]=tp, [1]=tc */
  double times[4];
} orc_field;

void orc_eval_set4(const orc_tree *const *trees, const double *times, double tq,
                   int periodic, double *pos, long n, double *out) {
  const int dof = trees[0]->dof;
  double *tmp = (double *)malloc(sizeof(double) * 4 * n * dof);
  for (int m = 0; m < 4; m++)
    orc_eval_tree(trees[m], periodic, pos, n, tmp + (size_t)m * n * dof, NULL);
  for (long i = 0; i < n * dof; i++) {
    double g[4] = {tmp[i], tmp[n * dof + i], tmp[2 * n * dof + i], tmp[3 * n * dof + i]};
    out[i] = orc_interp_cubic1d(tq, times, g);
  }
  free(tmp);
}

void orc_eval_extrap(const orc_tree *tp, const orc_tree *tc, int periodic, double *pos,
                     long n, double *out) {
  const int dof = tp->dof;
  double *vc = (double *)malloc(sizeof(double) * n * dof);
  double *vp = (double *)malloc(sizeof(double) * n * dof);
  orc_eval_tree(tc, periodic, pos, n, vc, NULL); /* T^n first, :59-61 */
  orc_eval_tree(tp, periodic, pos, n, vp, NULL);
  const double cc = 3.0 / 2, pc = 0.5;
  for (long i = 0; i < n * dof; i++) out[i] = cc * vc[i] - pc * vp[i];
  free(vc);
  free(vp);
}

static void field_eval(const orc_field *f, int periodic, double *pos, long n, double tq,
                       double *out) {
  if (f->kind == ORC_VEL_STEADY)
    orc_eval_tree(f->tree[0], periodic, pos, n, out, NULL); /* time ignored, :808-811 */
  else if (f->kind == ORC_VEL_SET4)
    orc_eval_set4(f->tree, f->times, tq, periodic, pos, n, out);
  else
    orc_eval_extrap(f->tree[0], f->tree[1], periodic, pos, n, out);
}

/* tbslas::IntegrateRK2 (traj.inc:21-46; two-functor variant :71-92) and
 * tbslas::ComputeTrajRK2 (:49-68, :95-115).  The solution vector aliases the
 * input of every sub-step (:64), so a periodic wrap done by the first evaluation is
 * visible in the final update x + tau*v, as in the reference.  f2 == NULL means
 * the single-functor form (both stages sample f1, second at t + tau/2). */
void orc_traj_rk2(const orc_field *f1, const orc_field *f2, int periodic,
                  const double *xinit, long n, double tinit, double tfinal, int nrk,
                  double *xsol) {
  const double tau = (tfinal - tinit) / nrk;
  double *xtmp = (double *)malloc(sizeof(double) * 3 * n);
  double *vtmp = (double *)malloc(sizeof(double) * 3 * n);
  for (long i = 0; i < 3 * n; i++) xsol[i] = xinit[i];
  double tcur = tinit;
  for (int s = 0; s < nrk; s++) {
    field_eval(f1, periodic, xsol, n, tcur, vtmp);
    for (long i = 0; i < 3 * n; i++) xtmp[i] = xsol[i] + 0.5 * tau * vtmp[i];
    field_eval(f2 ? f2 : f1, periodic, xtmp, n, tcur + 0.5 * tau, vtmp);
    for (long i = 0; i < 3 * n; i++) xsol[i] = xsol[i] + tau * vtmp[i];
    tcur = tcur + tau;
  }
  free(xtmp);
  free(vtmp);
}

/* tbslas::SolveSemilagRK2, semilag.inc:27-45 (:49-69): departure points over
 * [timestep*dt, timestep*dt - dt], then the advected field sampled there. */
void orc_semilag_rk2(const orc_field *f1, const orc_field *f2, const orc_tree *con,
                     int periodic, const double *pos, long n, int timestep, double dt,
                     int nrk, double *vals, double *dep_or_null) {
  const double tinit = timestep * dt;
  const double tfinal = tinit - dt;
  double *dep = (double *)malloc(sizeof(double) * 3 * n);
  orc_traj_rk2(f1, f2, periodic, pos, n, tinit, tfinal, nrk, dep);
  orc_eval_tree(con, periodic, dep, n, vals, NULL);
  if (dep_or_null) memcpy(dep_or_null, dep, sizeof(double) * 3 * n);
  free(dep);
}

/* tbslas::fast_interp, tree_functor.h:89-153: node-centred N_reg^3 grid
 * [dof][z][y][x]; 0 outside [0,1]^3; 4^3 Lagrange stencil clamped to the grid. */
void orc_fast_interp(const double *grid, int dof, int n_reg, const double *pts, long n,
                     double *out) {
  double den[4];
  for (int i = 0; i < 4; i++) {
    den[i] = 1;
    for (int j = 0; j < 4; j++)
      if (i != j) den[i] /= (double)(i - j);
  }
  const long n3 = (long)n_reg * n_reg * n_reg;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < n; p++) {
    const double *x = pts + 3 * p;
    if (x[0] < 0 || x[0] > 1.0 || x[1] < 0 || x[1] > 1.0 || x[2] < 0 || x[2] > 1.0) {
      for (int k = 0; k < dof; k++) out[p * dof + k] = 0;
      continue;
    }
    double pt[3], M[3][4];
    int g[3];
    for (int a = 0; a < 3; a++) {
      pt[a] = x[a] * (n_reg - 1);
      g[a] = ((int)pt[a]) - 1;
      if (g[a] < 0) g[a] = 0;
      if (g[a] > n_reg - 4) g[a] = n_reg - 4;
      pt[a] -= g[a];
      for (int k = 0; k < 4; k++) {
        M[a][k] = den[k];
        for (int l = 0; l < 4; l++)
          if (k != l) M[a][k] *= (pt[a] - l);
      }
    }
    for (int k = 0; k < dof; k++) {
      double val = 0;
      for (int j2 = 0; j2 < 4; j2++)
        for (int j1 = 0; j1 < 4; j1++) {
          const double m12 = M[1][j1] * M[2][j2];
          const long base = (long)n_reg * ((g[1] + j1) + (long)n_reg * (g[2] + j2));
          for (int j0 = 0; j0 < 4; j0++)
            val += M[0][j0] * m12 * grid[(g[0] + j0) + base + k * n3];
        }
      out[p * dof + k] = val;
    }
  }
}

/* tbslas::new_nodes, cheb.h:41-68: stretched Chebyshev nodes including the end
 * points, tensor grid with x fastest.  out: (q+1)^dim * dim doubles. */
long orc_new_nodes(int q, int dim, double *out) {
  const unsigned d = q + 1;
  double x[ORC_MAX_Q + 1];
  const double pi = 3.14159265358979323846264338327950288;
  const double scal = 1.0 / cos(0.5 * pi / d);
  for (unsigned i = 0; i < d; i++) x[i] = -cos((i + 0.5) * pi / d) * scal * 0.5 + 0.5;
  unsigned n1 = 1;
  for (int i = 0; i < dim; i++) n1 *= d;
  if (out)
    for (int i = 0; i < dim; i++) {
      unsigned n2 = 1;
      for (int k = 0; k < i; k++) n2 *= d;
      for (unsigned j = 0; j < n1; j++) out[j * dim + i] = x[(j / n2) % d];
    }
  return (long)n1 * dim;
}

/* tbslas::CollectChebTreeGridPoints, tree_utils.h:442-498: arrival points,
 * leaf-major, coord + 2^-depth * node. */
void orc_collect_grid_points(const orc_tree *t, double *out) {
  const long P = orc_new_nodes(t->q, 3, NULL) / 3;
  double *nodes = (double *)malloc(sizeof(double) * 3 * P);
  orc_new_nodes(t->q, 3, nodes);
  for (long j = 0; j < t->n_leaf; j++) {
    const double len = pow(0.5, t->depth[j]);
    const double *c = t->coord + 3 * j;
    double *o = out + (size_t)j * 3 * P;
    for (long i = 0; i < P; i++)
      for (int a = 0; a < 3; a++) o[3 * i + a] = c[a] + len * nodes[3 * i + a];
  }
  free(nodes);
}

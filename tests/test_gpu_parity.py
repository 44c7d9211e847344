"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Bars (BASELINE.json north_star):
  * leaf assignment of every point: BIT-EXACT;
  * interpolated values: within 1e-12 relative in FP64.  "Relative" is taken against the
    field scale of the batch (max |value|), because tensor-Chebyshev sums cancel and a
    per-point relative bound is unattainable near zeros (last item);
    the only arithmetic difference to the CPU path is FMA vs mul+add in the accumulation;
  * positions after periodic wrap, cubic-grid values, time interpolation/extrapolation of
    equal inputs, arrival points: BIT-EXACT (same operation order, un-fused).
"""
import numpy as np
import pytest

from conftest import golden
from tbslas_b200 import flat_tree as ftm

pytestmark = pytest.mark.gpu

RTOL = 1e-12
# Composed semi-Lagrangian values (trajectory, then the scalar AT the departure point): the
# north-star bar again.  Where a test's field is steep enough for a last-bit difference of the
# departure point to show above 1e-12, the test checks the two stages separately instead
# (departure points 1e-12 absolute; values at EQUAL points 1e-12) and says so.
STEP_TOL = 1e-12


def rel_err(a, b):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def _tree(g, prefix="tree"):
    return ftm.FlatTree(int(g[prefix + "_q"]), int(g[prefix + "_dof"]), g[prefix + "_coord"],
                        g[prefix + "_depth"], g[prefix + "_coeff"])


def _api():
    from tbslas_b200 import api
    return api


def adaptive_leaves(max_depth=5, min_depth=2):
    def refine(lower, edge, d):
        c = lower + 0.5 * edge[:, None]
        r = np.sqrt(((c - np.array([0.55, 0.5, 0.45])) ** 2).sum(axis=1))
        return np.abs(r - 0.3) < edge
    return ftm.adaptive_leaves(refine, min_depth, max_depth)


# ---------------------------------------------------------------- golden fixtures
def test_gpu_kat_appendix_c(ctx):
    api = _api()
    g = golden("kat_depth1.npz")
    t = ctx.tree(_tree(g))
    f = api.NodeFieldFunctor(t)
    for bc in (0, 1):
        pos = g["pts"].copy()
        v, leaf = f.eval_with_leaf(pos, bc)
        assert np.array_equal(leaf, g["leaf_bc%d" % bc])
        assert np.array_equal(v, g["val_bc%d" % bc])  # single-term sums: exact
        assert np.array_equal(pos, g["pos_bc%d" % bc])  # wrapped in place like the reference


def test_gpu_golden_eval_adaptive(ctx):
    api = _api()
    g = golden("eval_adaptive_q6.npz")
    f = api.NodeFieldFunctor(ctx.tree(_tree(g)))
    for bc in (0, 1):
        pos = g["pts"].copy()
        v, leaf = f.eval_with_leaf(pos, bc)
        assert np.array_equal(leaf, g["leaf_bc%d" % bc])
        assert np.array_equal(pos, g["pos_bc%d" % bc])
        assert rel_err(v, g["val_bc%d" % bc]) < RTOL


def test_gpu_golden_semilag(ctx):
    api = _api()
    g = golden("semilag_rotation_q5.npz")
    vel = api.NodeFieldFunctor(ctx.tree(_tree(g, "vel")))
    con = api.NodeFieldFunctor(ctx.tree(_tree(g, "con")))
    dt, ts, nrk = float(g["dt"]), int(g["timestep"]), int(g["nrk"])
    for bc in (0, 1):
        x = api.ComputeTrajRK2(vel, g["pts"], ts * dt, ts * dt - dt, nrk, bc)
        assert np.abs(x - g["traj_bc%d" % bc]).max() < RTOL
        s = api.SolveSemilagRK2(vel, con, g["pts"], ts, dt, nrk, bc)
        assert rel_err(s, g["semilag_bc%d" % bc]) < STEP_TOL


def test_gpu_golden_timevarying(ctx):
    api = _api()
    g = golden("timevarying_q4.npz")
    trees = [ctx.tree(ftm.FlatTree(int(g["q"]), int(g["dof"]), g["coord"], g["depth"], g["coeff4"][i]))
             for i in range(4)]
    fset = api.FieldSetFunctor(trees, g["times"].tolist())
    fext = api.FieldExtrapFunctor(trees[0], trees[1])
    fcur = api.NodeFieldFunctor(trees[1])
    for bc, combine in ((0, True), (1, True), (0, False), (1, False)):
        # combine: coefficients interpolated in time + one evaluation (default); else every
        # snapshot is evaluated and the values are combined per point, in the reference's order
        ctx.set_time_combine(combine)
        v = fset(g["pts"].copy(), time=float(g["tq"]), bc=bc)
        assert rel_err(v, g["set4_bc%d" % bc]) < RTOL
        e = fext(g["pts"].copy(), bc=bc)
        assert rel_err(e, g["extrap_bc%d" % bc]) < RTOL
        x = api.ComputeTrajRK2(fset, g["pts"], 0.1, 0.0, 1, bc)
        assert np.abs(x - g["traj_set4_bc%d" % bc]).max() < RTOL
        x = api.ComputeTrajRK2(fcur, g["pts"], 0.1, 0.0, 1, bc, extrap_fn=fext)
        assert np.abs(x - g["traj_extrap_bc%d" % bc]).max() < RTOL
    ctx.set_time_combine(True)
    # snapshots on DIFFERENT leaf lists cannot be combined: they take the per-tree route
    c2, d2 = ftm.uniform_leaves(2)
    other = ctx.tree(ftm.random_tree(c2, d2, int(g["q"]), int(g["dof"]), seed=3))
    mixed = api.FieldExtrapFunctor(trees[0], other)
    pts = g["pts"].copy()
    want = 1.5 * api.NodeFieldFunctor(other)(pts.copy(), bc=1) - 0.5 * api.NodeFieldFunctor(trees[0])(pts.copy(), bc=1)
    assert rel_err(mixed(pts.copy(), bc=1), want) < RTOL


def test_gpu_golden_cubic_grid(ctx):
    g = golden("cubic_grid_n12.npz")
    v = ctx.fast_interp(np.ascontiguousarray(g["grid"]), int(g["dof"]), int(g["n_reg"]), g["pts"])
    assert np.array_equal(v, g["val"])  # un-fused, same order: bit-exact


# ---------------------------------------------------------------- seeded parity vs oracle
@pytest.mark.parametrize("q", list(range(1, 20)))
def test_gpu_eval_every_degree(ctx, port, q):
    api = _api()
    coord, dd = ftm.uniform_leaves(2)
    dof = 3 if q % 2 else 1
    ft = ftm.random_tree(coord, dd, q, dof, seed=100 + q)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    rng = np.random.default_rng(q)
    pts = rng.uniform(-0.02, 1.02, size=(6000 + 37 * q, 3))
    for bc in (0, 1):
        vo, lo, po = port.eval_tree(h, dof, pts, bc)
        pos = pts.copy()
        v, leaf = f.eval_with_leaf(pos, bc)
        assert np.array_equal(leaf, lo)
        assert np.array_equal(pos, po)
        assert rel_err(v, vo) < RTOL


@pytest.mark.parametrize("q,dof", [(14, 4), (19, 6), (19, 8), (11, 7)])
def test_gpu_eval_wide_dof(ctx, port, q, dof):
    """Coefficient blocks too large for eight resident workers per SM: the persistent kernel runs with fewer
    workers per SM (q 14 dof 4: 28 KB; q 19 dof 6: 81 KB), beyond 100 KB (q 19 dof 8) the degree-generic kernel
    takes over.  Same leaves and values as the oracle either way."""
    api = _api()
    coord, dd = ftm.uniform_leaves(1)
    ft = ftm.random_tree(coord, dd, q, dof, seed=7 * q + dof)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    pts = np.random.default_rng(q + dof).uniform(-0.01, 1.01, size=(9000, 3))
    vo, lo, _ = port.eval_tree(h, dof, pts, 0)
    v, leaf = f.eval_with_leaf(pts.copy(), 0)
    assert np.array_equal(leaf, lo)
    assert rel_err(v, vo) < RTOL


@pytest.mark.parametrize("variant", ["1", "2"])
def test_gpu_eval_kernel_variants_agree(port, variant, monkeypatch):
    """The A/B kernels behind TBSLAS_EVAL_VARIANT (1: one tile per CTA, q 8 and 14; 2: the degree-generic kernel,
    which keeps T_0 as data and the reference's loop order) against the oracle, in a context of their own (the
    switch is read once, at tbslas_b200_init)."""
    api = _api()
    monkeypatch.setenv("TBSLAS_EVAL_VARIANT", variant)
    own = api.Context(0)
    try:
        for q, dof in ((8, 3), (14, 1)):
            coord, dd = ftm.uniform_leaves(2)
            ft = ftm.random_tree(coord, dd, q, dof, seed=q)
            f = api.NodeFieldFunctor(own.tree(ft))
            h = port.tree_create(ft)
            pts = np.random.default_rng(q).uniform(-0.01, 1.01, size=(7000, 3))
            vo, lo, _ = port.eval_tree(h, dof, pts, 0)
            v, leaf = f.eval_with_leaf(pts.copy(), 0)
            assert np.array_equal(leaf, lo)
            assert rel_err(v, vo) < RTOL
    finally:
        own.close()


@pytest.mark.parametrize("q,dof,bc", [(8, 3, 0), (14, 1, 1), (4, 2, 1), (14, 3, 0)])
def test_gpu_eval_adaptive_tree(ctx, port, q, dof, bc):
    api = _api()
    coord, dd = adaptive_leaves()
    ft = ftm.random_tree(coord, dd, q, dof, seed=q + dof)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-0.1, 1.1, size=(60000, 3)),
                          rng.integers(0, 65, size=(3000, 3)) / 64.0,      # leaf faces and corners
                          0.3 + 0.01 * rng.standard_normal((20000, 3))])  # one crowded region
    vo, lo, po = port.eval_tree(h, dof, pts, bc)
    pos = pts.copy()
    v, leaf = f.eval_with_leaf(pos, bc)
    assert np.array_equal(leaf, lo)
    assert np.array_equal(pos, po)
    assert rel_err(v, vo) < RTOL


def test_gpu_edge_cases(ctx, port):
    api = _api()
    coord, dd = ftm.uniform_leaves(1)
    ft = ftm.random_tree(coord, dd, 5, 3, seed=3)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    # empty input
    v, leaf = f.eval_with_leaf(np.zeros((0, 3)), 0)
    assert v.shape == (0, 3) and leaf.shape == (0,)
    # single point, ragged counts, everything in one leaf, everything outside
    for pts in (np.array([[0.3, 0.6, 0.9]]),
                np.random.default_rng(1).uniform(0, 1, size=(1000 + 13, 3)),
                0.1 + 0.3 * np.random.default_rng(2).uniform(0, 1, size=(5000, 3)),
                1.5 + np.random.default_rng(3).uniform(0, 1, size=(257, 3)),
                -np.random.default_rng(4).uniform(0.1, 1, size=(100, 3))):
        vo, lo, _ = port.eval_tree(h, 3, pts, 0)
        v, leaf = f.eval_with_leaf(pts.copy(), 0)
        assert np.array_equal(leaf, lo)
        assert rel_err(v, vo) < RTOL or np.abs(vo).max() == 0 and np.abs(v).max() == 0
    # NaN coordinates: no leaf arithmetic may trap; value is 0 like any out-of-domain point
    v, leaf = f.eval_with_leaf(np.array([[np.nan, 0.5, 0.5], [0.25, 0.25, 0.25]]), 0)
    assert leaf[1] == 0 and v[0].tolist() == [0, 0, 0]


def test_gpu_partial_tree_null_leaf(ctx, port):
    """A leaf list that does not start at the origin (one rank's shard): points before the
    first leaf get leaf -1 and value 0 (the reference never evaluates them)."""
    api = _api()
    coord, dd = ftm.uniform_leaves(2)
    ft = ftm.random_tree(coord, dd, 4, 1, seed=9).shard(20, 50)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    pts = np.random.default_rng(5).uniform(0, 1, size=(4000, 3))
    vo, lo, _ = port.eval_tree(h, 1, pts, 0)
    v, leaf = f.eval_with_leaf(pts.copy(), 0)
    assert (lo == -1).any()
    assert np.array_equal(leaf, lo)
    assert rel_err(v, vo) < RTOL


@pytest.mark.parametrize("q", [8, 14])
def test_gpu_points_outside_their_leaf_evaluate_to_zero(ctx, port, q):
    """cheb_poly makes every basis value 0 for a coordinate outside [-1, 1], so a point that
    the assignment rule puts into a leaf whose box does not contain it -- every point behind the last leaf
    of a rank's shard -- evaluates to exactly 0, in one, two or three axes alike.  The evaluation kernels
    never read T_0 (inside the leaf it is 1) and zero such points with a select: checked here on
    the unrolled kernels the benchmarks run (q = 8 and 14, three components), through the plain store and
    through the fused RK2 update (x + alpha * 0 must be x itself)."""
    api = _api()
    coord, dd = ftm.uniform_leaves(2)
    ft = ftm.random_tree(coord, dd, q, 3, seed=31 + q).shard(9, 41)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    pts = np.random.default_rng(6).uniform(0, 1, size=(20000, 3))
    vo, lo, _ = port.eval_tree(h, 3, pts, 0)
    v, leaf = f.eval_with_leaf(pts.copy(), 0)
    assert np.array_equal(leaf, lo)
    size = np.power(0.5, dd[9:41].astype(np.float64))
    inside = np.zeros(len(pts), bool)
    ok = lo >= 0
    rel = (pts[ok] - ft.coord[lo[ok]]) / size[lo[ok], None]
    inside[ok] = ((rel >= 0) & (rel <= 1)).all(axis=1)
    outside = ok & ~inside
    assert outside.sum() > 1000 and inside.sum() > 1000
    assert np.all(v[outside] == 0.0) and np.all(vo[outside] == 0.0)
    assert rel_err(v, vo) < RTOL
    xo = port.traj_rk2(h, pts, 0.02, 0.0, 1, 0)
    x = api.ComputeTrajRK2(f, pts, 0.02, 0.0, 1, 0)
    assert np.abs(x - xo).max() < RTOL
    still = outside & (np.abs(xo - pts).max(axis=1) == 0.0)  # never entered the shard: did not move at all
    assert still.sum() > 100 and np.array_equal(x[still], pts[still])


@pytest.mark.parametrize("bc", [0, 1])
def test_gpu_locate_faces_corners_and_leaf_major_streams(ctx, port, bc):
    """The box fast path of the locate kernel: points exactly on leaf faces, edges
    and corners (they belong to the upper neighbour), one ulp either side of them, the domain
    faces 0 and 1, and a long leaf-major stream (the order departure points arrive in) must get
    the oracle's leaf, bit for bit."""
    api = _api()
    coord, dd = adaptive_leaves(6, 2)
    ft = ftm.random_tree(coord, dd, 3, 1, seed=21)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    size = np.power(0.5, dd.astype(np.float64))
    rng = np.random.default_rng(77)
    corners = (coord[:, None, :] + size[:, None, None] *
               np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=np.float64)[None]).reshape(-1, 3)
    mids = coord + 0.5 * size[:, None]
    faces = np.concatenate([np.where(np.arange(3) == a, coord + s * size[:, None], mids)
                            for a in range(3) for s in (0.0, 1.0)])
    special = np.concatenate([corners, faces, np.nextafter(corners, 0.0), np.nextafter(corners, 2.0),
                              np.nextafter(faces, 0.0), np.nextafter(faces, 2.0)])
    stream = ftm.grid_points(coord, dd, 5) - 0.013 * rng.standard_normal((1, 3))  # shifted arrival points
    for pts in (special, stream, np.concatenate([stream[::7], special, stream[::5]])):
        vo, lo, po = port.eval_tree(h, 1, pts, bc)
        pos = pts.copy()
        v, leaf = f.eval_with_leaf(pos, bc)
        assert np.array_equal(leaf, lo)
        assert np.array_equal(pos, po)
        assert rel_err(v, vo) < RTOL


@pytest.mark.parametrize("case", ["root", "depth1", "deep_line", "deep_line_shard", "random"])
def test_gpu_locate_tree_shapes(ctx, port, case):
    """The cell table of the locate kernel against the oracle on tree shapes that
    stress it: leaves coarser than the table cells (a single root leaf), leaves many levels
    finer than the cells (a needle refined to depth 11: long ranges per cell), a shard of it
    (keys before the first and after the last local leaf), and randomly refined trees."""
    api = _api()
    rng = np.random.default_rng({"root": 1, "depth1": 2, "deep_line": 3, "deep_line_shard": 4, "random": 5}[case])
    if case == "root":
        coord, dd = np.zeros((1, 3)), np.zeros(1, dtype=np.uint8)
    elif case == "depth1":
        coord, dd = ftm.uniform_leaves(1)
    elif case in ("deep_line", "deep_line_shard"):
        def refine(lower, edge, d):  # cells cut by the segment x = y = z near 0.3..0.31
            lo, hi = lower, lower + edge[:, None]
            return np.all((lo <= 0.31) & (hi >= 0.3), axis=1)
        coord, dd = ftm.adaptive_leaves(refine, 1, 11)
    else:
        def refine(lower, edge, d):
            return rng.random(lower.shape[0]) < 0.45
        coord, dd = ftm.adaptive_leaves(refine, 1, 7)
    ft = ftm.random_tree(coord, dd, 3, 2, seed=17)
    if case == "deep_line_shard":
        ft = ft.shard(ft.n_leaf // 3, 2 * ft.n_leaf // 3)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    size = np.power(0.5, ft.depth.astype(np.float64))
    pick = rng.integers(0, ft.n_leaf, size=30000)
    inside = ft.coord[pick] + size[pick, None] * rng.random((30000, 3))        # leaf-weighted
    pts = np.concatenate([rng.uniform(-0.05, 1.05, size=(30000, 3)), inside,
                          ft.coord[pick[:5000]], ft.coord[pick[:5000]] + size[pick[:5000], None]])
    for bc in (0, 1):
        vo, lo, po = port.eval_tree(h, 2, pts, bc)
        pos = pts.copy()
        v, leaf = f.eval_with_leaf(pos, bc)
        assert np.array_equal(leaf, lo)
        assert np.array_equal(pos, po)
        assert rel_err(v, vo) < RTOL or np.abs(vo).max() == 0


def test_gpu_locate_overlapping_leaves_fall_back_to_keys(ctx, port):
    """A leaf list that is NOT a set of disjoint octants (a depth-2 octant listed next to the
    depth-1 octant that contains it): 'inside the box' no longer implies 'last leaf with key <=
    key(point)', so the library must drop the box fast path and keep the reference's key rule."""
    api = _api()
    coord, dd = ftm.uniform_leaves(1)
    extra = np.array([[0.25, 0.0, 0.0]])            # second x-child of leaf 0, depth 2
    coord2 = np.concatenate([coord[:1], extra, coord[1:]])
    dd2 = np.concatenate([dd[:1], np.array([2], dtype=dd.dtype), dd[1:]])
    ft = ftm.random_tree(coord2, dd2, 4, 2, seed=5)
    assert np.all(np.diff(ft.keys().astype(np.int64)) > 0)
    f = api.NodeFieldFunctor(ctx.tree(ft))
    h = port.tree_create(ft)
    rng = np.random.default_rng(8)
    pts = np.concatenate([rng.uniform(0, 0.5, size=(20000, 3)), rng.uniform(0, 1, size=(20000, 3))])
    vo, lo, _ = port.eval_tree(h, 2, pts, 0)
    v, leaf = f.eval_with_leaf(pts.copy(), 0)
    assert (lo == 1).any() and (lo == 0).any()
    assert np.array_equal(leaf, lo)
    assert rel_err(v, vo) < RTOL


@pytest.mark.parametrize("bc", [0, 1])
@pytest.mark.parametrize("nrk", [1, 3])
def test_gpu_traj_and_semilag_vs_oracle(ctx, port, bc, nrk):
    api = _api()
    coord, dd = adaptive_leaves(4, 2)
    tv = ftm.fit(coord, dd, 8, 3, lambda p: ftm.vel_rotation(p) + 0.05 * ftm.vel_taylor_green(p))
    tc = ftm.fit(coord, dd, 8, 1, lambda p: ftm.gaussian(p, sigma=0.12))
    vel, con = api.NodeFieldFunctor(ctx.tree(tv)), api.NodeFieldFunctor(ctx.tree(tc))
    hv, hc = port.tree_create(tv), port.tree_create(tc)
    pts = ftm.grid_points(coord, dd, 8)[::7].copy()  # a subset of the real arrival points
    dt = 0.0628
    xo = port.traj_rk2(hv, pts, 2 * dt, dt, nrk, bc)
    x = api.ComputeTrajRK2(vel, pts, 2 * dt, dt, nrk, bc)
    assert np.abs(x - xo).max() < RTOL
    so = port.semilag_rk2(hv, hc, 1, pts, 2, dt, nrk, bc)
    dep = np.empty_like(pts)
    s = api.SolveSemilagRK2(vel, con, pts, 2, dt, nrk, bc, departure_points=dep)
    assert np.abs(dep - xo).max() < RTOL
    assert rel_err(s, so) < STEP_TOL


@pytest.mark.parametrize("bc", [0, 1])
def test_gpu_ns_call_pattern_vs_oracle(ctx, port, bc):
    """The Navier-Stokes stepper's use of the path (tree_ns.h:466-518): the advected field is the dof-3
    velocity itself.  Two backward trajectories from the same arrival points -- [t, t-dt] with stage 1
    = v^n and stage 2 = the extrapolation 1.5 v^n - 0.5 v^{n-1}; [t, t-2dt] with stage 1 = v^{n-1} and
    stage 2 = v^n -- the velocity trees sampled at the two sets of departure points, combined
    2/dt*c - 0.5/dt*p, transposed point-major -> dof-major and refitted.  Stage by stage against the
    oracle: departure points 1e-12, values at equal points 1e-12, refit 1e-12 of the coefficient scale."""
    api = _api()
    q, dt, nrk, ts = 6, 0.02, 2, 3
    coord, dd = adaptive_leaves(4, 2)
    fp_ = ftm.fit(coord, dd, q, 3, lambda p: 0.9 * ftm.vel_rotation(p) + 0.2 * ftm.vel_taylor_green(p))
    fc_ = ftm.fit(coord, dd, q, 3, lambda p: 1.0 * ftm.vel_rotation(p) + 0.25 * ftm.vel_taylor_green(p))
    tp, tc, tn = ctx.tree(fp_), ctx.tree(fc_), ctx.tree(fc_)
    fp, fc, fe = api.NodeFieldFunctor(tp), api.NodeFieldFunctor(tc), api.FieldExtrapFunctor(tp, tc)
    hp, hc = port.tree_create(fp_), port.tree_create(fc_)
    arr = tn.collect_grid_points()
    tcur = ts * dt
    d1 = api.ComputeTrajRK2(fc, arr, tcur, tcur - dt, nrk, bc, extrap_fn=fe)
    c = fc(d1.copy(), bc=bc)
    d2 = api.ComputeTrajRK2(fp, arr, tcur, tcur - 2 * dt, nrk, bc, extrap_fn=fc)
    p = fp(d2.copy(), bc=bc)
    assert np.abs(d1 - port.traj_rk2((hp, hc), arr, tcur, tcur - dt, nrk, bc, kind="extrap")).max() < RTOL
    assert np.abs(d2 - port.traj_rk2((hp, hc), arr, tcur, tcur - 2 * dt, nrk, bc, kind="pair")).max() < RTOL
    assert rel_err(c, port.eval_tree(hc, 3, d1, bc, want_leaf=False)[0]) < RTOL
    assert rel_err(p, port.eval_tree(hp, 3, d2, bc, want_leaf=False)[0]) < RTOL
    val = (2.0 / dt) * c + (-0.5 / dt) * p                       # [L*P, 3] point-major
    P = (q + 1) ** 3
    ml = np.ascontiguousarray(val.reshape(-1, P, 3).transpose(0, 2, 1))  # [L, 3, P] (tree_ns.h:502-513)
    ctx.set_pt2coeff(q)
    tn.set_grid_values(ml, point_major=False)
    want = np.einsum("ldp,pn->ldn", ml, ftm.pt2coeff(q))
    assert np.abs(tn.coefficients() - want).max() < 1e-12 * np.abs(want).max()
    tn.set_grid_values(val, point_major=True)                    # the library's own transpose
    assert np.abs(tn.coefficients() - want).max() < 1e-12 * np.abs(want).max()
    for t in (tp, tc, tn):
        t.destroy()


def test_gpu_device_resident_buffers(ctx, port):
    """Same results when the caller's buffers are device pointers (torch CUDA tensors)."""
    import torch
    api = _api()
    coord, dd = ftm.uniform_leaves(3)
    tv = ftm.fit(coord, dd, 6, 3, ftm.vel_rotation)
    tc = ftm.random_tree(coord, dd, 6, 1, seed=4)
    vel, con = api.NodeFieldFunctor(ctx.tree(tv)), api.NodeFieldFunctor(ctx.tree(tc))
    pts = np.random.default_rng(11).uniform(0, 1, size=(50000, 3))
    host = api.SolveSemilagRK2(vel, con, pts, 1, 0.05, 1, 1)
    ctx.set_stream(torch.cuda.current_stream())
    dpts = torch.from_numpy(pts).cuda()
    dev = api.SolveSemilagRK2(vel, con, dpts, 1, 0.05, 1, 1)
    torch.cuda.synchronize()
    ctx.set_stream(None)
    assert np.array_equal(dev.cpu().numpy(), host)
    assert np.array_equal(dpts.cpu().numpy(), pts)  # arrival points are not modified


def test_gpu_semilag_insitu_equals_explicit_points(ctx):
    """tbslas_b200_semilag_insitu (arrival points generated in HBM) == semilag_rk2 on the
    collected grid points, bit for bit."""
    api = _api()
    coord, dd = adaptive_leaves(4, 2)
    tv = ftm.fit(coord, dd, 5, 3, ftm.vel_rotation)
    tc = ftm.random_tree(coord, dd, 5, 2, seed=21)
    tvel, tcon = ctx.tree(tv), ctx.tree(tc)
    vel, con = api.NodeFieldFunctor(tvel), api.NodeFieldFunctor(tcon)
    pts = tcon.collect_grid_points()
    for bc in (0, 1):
        a = api.SolveSemilagRK2(vel, con, pts, 2, 0.04, 2, bc)
        ctx.set_tensor_grid(False)   # every evaluation point by point: the very same arithmetic
        b = api.SolveSemilagInSitu(vel, tcon, 2, 0.04, 2, bc)
        ctx.set_tensor_grid("always")  # first velocity evaluation by sum factorisation (any size)
        c = api.SolveSemilagInSitu(vel, tcon, 2, 0.04, 2, bc)
        assert np.array_equal(a, b)
        assert rel_err(c, a) < STEP_TOL
    ctx.set_tensor_grid(True)


@pytest.mark.parametrize("q", [5, 8, 14])
@pytest.mark.parametrize("case", ["same_tree", "velocity_coarser", "velocity_finer", "time_varying"])
@pytest.mark.parametrize("bc", [0, 1])
def test_gpu_tensor_grid_velocity_vs_generic_and_oracle(ctx, port, case, bc, q):
    """tensor_eval.cu: the velocity at the arrival grids by sum factorisation against the
    point-by-point path (same library, switch off) and against the oracle's step, for advected
    leaves equal to, finer than and COARSER than the velocity leaves (the last: no containing
    velocity leaf, every point takes the generic path), both boundary conditions (grid points on
    leaf faces and on the domain boundary are the exceptions), and a 4-snapshot velocity -- at
    q = 5, 8 and 14 (14 is the register-blocked instantiation tensor_grid_eval_kernel_t<15> that the
    bench's workloads run).  Stage by stage: departure points 1e-12; values at equal points 1e-12."""
    api = _api()
    coord, dd = adaptive_leaves(4 if q < 14 else 3, 2)
    if case == "same_tree":
        vc, vd = coord, dd
    elif case == "velocity_coarser":
        vc, vd = ftm.uniform_leaves(1)
    else:
        vc, vd = ftm.uniform_leaves(3) if case == "velocity_finer" else ftm.uniform_leaves(2)
    fcon = ftm.fit(coord, dd, q, 1, lambda p: ftm.gaussian(p, (0.5, 0.5, 0.5), 0.25))
    tcon = ctx.tree(fcon)
    times = [-0.05, 0.0, 0.05, 0.1]
    if case == "time_varying":
        fv = [ftm.fit(vc, vd, q, 3, lambda p, s=s: ftm.vel_rotation(p) * s) for s in (0.8, 1.0, 1.2, 0.9)]
        vtrees = [ctx.tree(f) for f in fv]
        vel = api.FieldSetFunctor(vtrees, times)
    else:
        # a velocity that is DIScontinuous across leaves (random coefficients): a grid point on a leaf
        # face evaluated by the wrong leaf would be off by O(0.1)
        fv = [ftm.random_tree(vc, vd, q, 3, seed=41, scale=0.2)]
        vtrees = [ctx.tree(fv[0])]
        vel = api.NodeFieldFunctor(vtrees[0])
    ctx.set_tensor_grid(False)
    ref, ref_dep = api.SolveSemilagInSitu(vel, tcon, 1, 0.05, 1, bc, departure_points=True)
    ctx.set_tensor_grid("always")
    got, dep = api.SolveSemilagInSitu(vel, tcon, 1, 0.05, 1, bc, departure_points=True)
    assert ctx.last_grid_exceptions() > 0
    ctx.set_tensor_grid(True)
    assert np.abs(dep - ref_dep).max() < RTOL
    # oracle, on a strided sample of the arrival points (all of them at q = 5)
    pts = ftm.grid_points(coord, dd, q)
    sl = slice(None, None, 1 if q == 5 else 7)
    hv, hc = [port.tree_create(f) for f in fv], port.tree_create(fcon)
    kind = "set4" if case == "time_varying" else "steady"
    vh = hv if case == "time_varying" else hv[0]
    dep_o = port.traj_rk2(vh, pts[sl], 0.05, 0.0, 1, bc, kind=kind, times=times)
    assert np.abs(dep[sl] - dep_o).max() < RTOL
    want_at_dep, _, _ = port.eval_tree(hc, 1, dep[sl], bc)
    assert rel_err(got[sl], want_at_dep) < RTOL
    want = port.semilag_rk2(vh, hc, 1, pts[sl], 1, 0.05, 1, bc, kind=kind, times=times)
    assert rel_err(got[sl], want) < STEP_TOL  # smooth advected field: the composed step holds the bar too
    for t in vtrees + [tcon]:
        t.destroy()


@pytest.mark.parametrize("q,dof,n_leaf", [(4, 1, 8), (8, 3, 300), (14, 1, 137), (14, 3, 64), (5, 2, 1)])
def test_gpu_refit_values_to_coefficients(ctx, q, dof, n_leaf):
    """tbslas_b200_tree_set_grid_values (SetTreeGridValues, tree_utils.h:500-552) against the
    same GEMM in numpy: 1e-12 of the coefficient scale; both value layouts."""
    coord, dd = ftm.uniform_leaves(3)
    ft = ftm.random_tree(coord[:n_leaf], dd[:n_leaf], q, dof, seed=q)
    t = ctx.tree(ft)
    M = ftm.pt2coeff(q)
    ctx.set_pt2coeff(q, M)
    P = (q + 1) ** 3
    rng = np.random.default_rng(q + dof)
    vals = rng.standard_normal((n_leaf, dof, P))
    want = np.einsum("ldp,pn->ldn", vals, M)
    t.set_grid_values(vals)
    got = t.coefficients()
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
    t.set_grid_values(np.ascontiguousarray(vals.transpose(0, 2, 1)), point_major=True)
    assert np.abs(t.coefficients() - want).max() < 1e-12 * np.abs(want).max()
    # evaluating a (globally continuous) field of total degree <= q on the grid and refitting
    # reproduces its coefficients
    def poly(p):
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        return np.stack([1 + x - 2 * y * z + x * x * y, x * y * z - 0.5 * z ** 3, 0.25 + y ** 2 * z - x ** 4][:dof],
                        axis=1)
    t.destroy()
    fc, fd = ftm.uniform_leaves(1)  # the whole domain, so every grid point has a leaf
    fp = ftm.fit(fc, fd, q, dof, poly)
    t = ctx.tree(fp)
    v = _api().NodeFieldFunctor(t)(ftm.grid_points(fc, fd, q), bc=0)  # [L*P, dof] point-major
    t.set_grid_values(v, point_major=True)
    assert np.abs(t.coefficients() - fp.coeff).max() < 1e-11 * np.abs(fp.coeff).max()
    t.destroy()


def test_gpu_semilag_insitu_update(ctx):
    """The whole tree-level step on the device == step values refitted in numpy."""
    api = _api()
    coord, dd = adaptive_leaves(4, 2)
    q = 6
    tv = ftm.fit(coord, dd, q, 3, ftm.vel_rotation)
    tc = ftm.fit(coord, dd, q, 1, lambda p: ftm.gaussian(p, (0.5, 0.5, 0.5), 0.25))
    tvel, tcon = ctx.tree(tv), ctx.tree(tc)
    vel = api.NodeFieldFunctor(tvel)
    ctx.set_pt2coeff(q)
    vals = api.SolveSemilagInSitu(vel, tcon, 1, 0.05, 1, 0)      # [L*P, 1]
    want = np.einsum("lp,pn->ln", vals.reshape(tc.n_leaf, -1), ftm.pt2coeff(q))[:, None, :]
    api.SolveSemilagInSituUpdate(vel, tcon, 1, 0.05, 1, 0)
    got = tcon.coefficients()
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()


def test_gpu_leaf_point_counts_and_tail_norm(ctx):
    """Row f4 helpers: per-leaf point counts of the last evaluation == histogram of the leaf ids
    it returned; tail norm == numpy on the packed coefficients."""
    coord, dd = adaptive_leaves(5, 2)
    q, dof = 6, 3
    ft = ftm.random_tree(coord, dd, q, dof, seed=31)
    t = ctx.tree(ft)
    f = _api().NodeFieldFunctor(t)
    pts = np.random.default_rng(12).uniform(0, 1, size=(200001, 3)) ** 2
    _, leaf = f.eval_with_leaf(pts, 0)
    assert np.array_equal(t.last_point_counts(), np.bincount(leaf, minlength=ft.n_leaf).astype(np.uint32))
    f(pts[:1000], bc=0)
    assert t.last_point_counts().sum() == 1000
    shell = np.array([i + j + k == q for i in range(q + 1) for j in range(q + 1 - i)
                      for k in range(q + 1 - i - j)])
    want = np.sqrt((ft.coeff[:, :, shell] ** 2).sum(axis=(1, 2)))
    got = t.tail_norm()
    assert np.abs(got - want).max() <= 1e-14 * want.max()
    t.destroy()


def test_gpu_cpp_dropin_binary():
    """The reference's own C++ templates over the product's header-only adaptors
    (oracle/dropin_test.cpp, prebuilt into oracle/_ref/ where /root/reference exists)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                       "dropin_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL OK" in r.stdout


def test_gpu_grid_points_bit_exact(ctx, port):
    coord, dd = adaptive_leaves(4, 1)
    for q in (3, 8, 14):
        ft = ftm.random_tree(coord, dd, q, 1, seed=q)
        t = ctx.tree(ft)
        got = t.collect_grid_points()
        h = port.tree_create(ft)
        want = np.empty_like(got)
        port.lib.orc_collect_grid_points(h, want.ctypes.data_as(
            __import__("ctypes").POINTER(__import__("ctypes").c_double)))
        assert np.array_equal(got, want)


def test_gpu_cubic_grid_vs_oracle(ctx, port):
    rng = np.random.default_rng(8)
    for n_reg, dof in ((4, 1), (33, 3), (64, 1)):
        grid = rng.standard_normal((dof, n_reg, n_reg, n_reg))
        pts = rng.uniform(-0.02, 1.02, size=(40000, 3))
        pts[:64] = rng.integers(0, n_reg, size=(64, 3)) / (n_reg - 1.0)
        assert np.array_equal(ctx.fast_interp(grid, dof, n_reg, pts),
                              port.fast_interp(grid, dof, n_reg, pts))


@pytest.mark.parametrize("bc", [0, 1])
def test_gpu_virtual_arrival_points_same_bits(ctx, bc):
    """tbslas_b200_set_virtual_arrival_points: rebuilding the arrival points in the second stage's
    epilogue (never writing them to HBM) gives the very bits of the materialised path -- steady and
    4-snapshot velocity, nrk 1 and 2, values and departure points."""
    api = _api()
    q = 8
    coord, dd = adaptive_leaves(4, 2)
    vc, vd = ftm.uniform_leaves(2)
    tcon = ctx.tree(ftm.random_tree(coord, dd, q, 1, seed=3))
    fv = [ftm.random_tree(vc, vd, q, 3, seed=50 + k, scale=0.3) for k in range(4)]
    tv = [ctx.tree(f) for f in fv]
    ctx.set_tensor_grid("always")
    try:
        for vel in (api.NodeFieldFunctor(tv[0]), api.FieldSetFunctor(tv, [-0.05, 0.0, 0.05, 0.1])):
            for nrk in (1, 2):
                ctx.set_virtual_arrival_points(False)
                a, da = api.SolveSemilagInSitu(vel, tcon, 1, 0.04, nrk, bc, departure_points=True)
                ctx.set_virtual_arrival_points(True)
                b, db = api.SolveSemilagInSitu(vel, tcon, 1, 0.04, nrk, bc, departure_points=True)
                assert ctx.last_grid_exceptions() > 0
                assert np.array_equal(a, b) and np.array_equal(da, db)
    finally:
        ctx.set_virtual_arrival_points(False)
        ctx.set_tensor_grid(True)
        for t in tv + [tcon]:
            t.destroy()


def test_gpu_resident_cubic_grid_handle(ctx, port):
    """tbslas_b200_grid_create/eval/update/destroy: the same bits as the one-shot call, host and
    device buffers, chunked host pipeline included."""
    import torch
    rng = np.random.default_rng(18)
    n_reg, dof = 21, 3
    grid = rng.standard_normal((dof, n_reg, n_reg, n_reg))
    pts = rng.uniform(-0.02, 1.02, size=(70001, 3))
    want = port.fast_interp(grid, dof, n_reg, pts)
    g = ctx.grid(grid, dof, n_reg)
    assert np.array_equal(g(pts), want)
    ctx.set_host_chunks(5)
    assert np.array_equal(g(pts), want)
    ctx.set_host_chunks(0)
    ctx.set_stream(torch.cuda.current_stream())
    dv = g(torch.from_numpy(pts).cuda())
    torch.cuda.synchronize()
    ctx.set_stream(None)
    assert np.array_equal(dv.cpu().numpy(), want)
    grid2 = 2.0 * grid
    g.update(grid2)
    assert np.array_equal(g(pts), port.fast_interp(grid2, dof, n_reg, pts))
    g.destroy()


def test_gpu_reshard_single_rank_is_a_checked_no_op(ctx):
    coord, dd = ftm.uniform_leaves(2)
    t = ctx.tree(ftm.random_tree(coord, dd, 4, 1, seed=1))
    t.reshard([0, 64])
    assert t.global_range() == (0, 64)
    with pytest.raises(Exception):
        t.reshard([0, 63])
    t.destroy()


# ---------------------------------------------------------------- size-independent properties
def test_gpu_properties_at_config1_size(ctx):
    """BASELINE config 1 size (uniform depth 4, q = 8, N = 2 985 984 arrival points)."""
    api = _api()
    coord, dd = ftm.uniform_leaves(4)
    q = 8
    pts = ftm.grid_points(coord, dd, q)
    assert pts.shape[0] == 2985984
    # (1) polynomial reproduction: a degree-<=q polynomial field is exact at every point
    def poly(p):
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        return np.stack([1 + x - 2 * y * z, x * x * y - z ** 3, 0.5 + x ** 4 * y ** 2 * z ** 2], axis=1)
    tp = ftm.fit(coord, dd, q, 3, poly)
    f = api.NodeFieldFunctor(ctx.tree(tp))
    rng = np.random.default_rng(0)
    moved = pts + 0.01 * rng.standard_normal(pts.shape)
    v, leaf = f.eval_with_leaf(moved, 1)  # periodic: wraps in place
    assert np.abs(v - poly(moved)).max() < 1e-11
    # (2) leaf ids are exactly floor(coordinate * 16) in Morton order
    a = np.floor(moved * 16).astype(np.uint64)
    order = np.argsort(ftm.anchor_key(a[:, 0] << np.uint64(11), a[:, 1] << np.uint64(11),
                                      a[:, 2] << np.uint64(11)), kind="stable")
    lut = {}
    keys = ftm.FlatTree(q, 3, coord, dd, tp.coeff).keys()
    pk = ftm.anchor_key(a[:, 0] << np.uint64(11), a[:, 1] << np.uint64(11), a[:, 2] << np.uint64(11))
    assert np.array_equal(leaf, np.searchsorted(keys, pk, side="right") - 1)
    # (3) linearity: eval(a*A + b*B) == a*eval(A) + b*eval(B) to rounding
    A = ftm.random_tree(coord, dd, q, 1, seed=1)
    B = ftm.random_tree(coord, dd, q, 1, seed=2)
    AB = ftm.FlatTree(q, 1, coord, dd, 0.75 * A.coeff - 1.25 * B.coeff)
    va = api.NodeFieldFunctor(ctx.tree(A))(moved, bc=1)
    vb = api.NodeFieldFunctor(ctx.tree(B))(moved, bc=1)
    vab = api.NodeFieldFunctor(ctx.tree(AB))(moved, bc=1)
    assert np.abs(vab - (0.75 * va - 1.25 * vb)).max() < 1e-12 * max(1.0, np.abs(va).max())
    # (4) semi-Lagrangian round trip: forward then backward rotation returns to the start
    tv = ftm.fit(coord, dd, q, 3, ftm.vel_rotation)
    vel = api.NodeFieldFunctor(ctx.tree(tv))
    inner = pts[np.linalg.norm(pts[:, :2] - 0.5, axis=1) < 0.4]
    fwd = api.ComputeTrajRK2(vel, inner, 0.0, 0.05, 4, 0)
    back = api.ComputeTrajRK2(vel, fwd, 0.05, 0.0, 4, 0)
    assert np.abs(back - inner).max() < 1e-7  # O(dt^3) scheme, reversible to truncation error

"""CPU suite, part 1: the oracle.

(a) the plain-C restatement (oracle/tbslas_oracle.c) reproduces, BIT FOR BIT, the golden
    fixtures frozen from the reference's own code (tests/golden/, made by
    tests/golden/make_golden.py from oracle/_ref);
(b) where oracle/_ref is present, port and reference build agree bit for bit on fresh
    seeded inputs;
(c) the known answers of SURVEY.md Appendix C and analytic properties hold.
"""
import numpy as np
import pytest

from conftest import golden
from tbslas_b200 import flat_tree as ftm


def _tree(g, prefix="tree"):
    return ftm.FlatTree(int(g[prefix + "_q"]), int(g[prefix + "_dof"]), g[prefix + "_coord"],
                        g[prefix + "_depth"], g[prefix + "_coeff"])


# ------------------------------------------------------------------ (a) golden
def test_kat_appendix_c(port):
    g = golden("kat_depth1.npz")
    h = port.tree_create(_tree(g))
    for bc in (0, 1):
        v, leaf, p = port.eval_tree(h, 1, g["pts"], bc)
        assert np.array_equal(v, g["val_bc%d" % bc])
        assert np.array_equal(leaf, g["leaf_bc%d" % bc])
        assert np.array_equal(p, g["pos_bc%d" % bc])
    # the literal table of SURVEY.md Appendix C
    v0 = port.eval_tree(h, 1, g["pts"], 0)[0].ravel()
    assert v0.tolist() == [1, 2, 8, 8, 2, 1, 2, 5, 0, 0, 0]
    v1, _, p1 = port.eval_tree(h, 1, g["pts"], 1)
    assert v1.ravel().tolist() == [1, 2, 1, 8, 2, 1, 1, 1, 7, 8, 3]
    assert p1[8, 0] == pytest.approx(0.03) and p1[9, 0] == 0.99 and p1[10, 1] == 0.5
    assert p1[6, 0] == 0.0  # 1.0 wraps to 0


def test_golden_eval_adaptive(port):
    g = golden("eval_adaptive_q6.npz")
    h = port.tree_create(_tree(g))
    for bc in (0, 1):
        v, leaf, p = port.eval_tree(h, 3, g["pts"], bc)
        assert np.array_equal(leaf, g["leaf_bc%d" % bc])
        assert np.array_equal(v, g["val_bc%d" % bc])
        assert np.array_equal(p, g["pos_bc%d" % bc])


def test_golden_semilag(port):
    g = golden("semilag_rotation_q5.npz")
    hv, hc = port.tree_create(_tree(g, "vel")), port.tree_create(_tree(g, "con"))
    dt, ts, nrk = float(g["dt"]), int(g["timestep"]), int(g["nrk"])
    for bc in (0, 1):
        x = port.traj_rk2(hv, g["pts"], ts * dt, ts * dt - dt, nrk, bc)
        assert np.array_equal(x, g["traj_bc%d" % bc])
        s = port.semilag_rk2(hv, hc, 1, g["pts"], ts, dt, nrk, bc)
        assert np.array_equal(s, g["semilag_bc%d" % bc])


def test_golden_timevarying(port):
    g = golden("timevarying_q4.npz")
    trees = [ftm.FlatTree(int(g["q"]), int(g["dof"]), g["coord"], g["depth"], g["coeff4"][i])
             for i in range(4)]
    hs = [port.tree_create(t) for t in trees]
    for bc in (0, 1):
        v, _ = port.eval_set4(hs, g["times"], float(g["tq"]), 3, g["pts"], bc)
        assert np.array_equal(v, g["set4_bc%d" % bc])
        e, _ = port.eval_extrap(hs[0], hs[1], 3, g["pts"], bc)
        assert np.array_equal(e, g["extrap_bc%d" % bc])
        x = port.traj_rk2(hs, g["pts"], 0.1, 0.0, 1, bc, kind="set4", times=g["times"])
        assert np.array_equal(x, g["traj_set4_bc%d" % bc])
        x = port.traj_rk2((hs[0], hs[1]), g["pts"], 0.1, 0.0, 1, bc, kind="extrap")
        assert np.array_equal(x, g["traj_extrap_bc%d" % bc])


def test_golden_cubic_grid(port):
    g = golden("cubic_grid_n12.npz")
    v = port.fast_interp(g["grid"], int(g["dof"]), int(g["n_reg"]), g["pts"])
    assert np.array_equal(v, g["val"])


def test_golden_scalar_helpers(port):
    g = golden("scalar_helpers.npz")
    for i in range(g["cubic_x"].shape[0]):
        assert port.interp_cubic1d(float(g["cubic_x"][i]), g["cubic_xx"], g["cubic_pp"][i]) == \
            g["cubic_val"][i]
    for q in range(1, 17):
        assert np.array_equal(port.new_nodes(q, 1).ravel(), g["nodes_q%d" % q])
    assert np.array_equal(port.new_nodes(3, 3), g["nodes3_q3"])


# ------------------------------------------------------- (b) port vs reference build
@pytest.mark.parametrize("depth,q,dof,bc", [(3, 8, 3, 0), (2, 14, 1, 1), (3, 5, 2, 1), (1, 1, 1, 0),
                                            (2, 16, 1, 0)])
def test_port_equals_reference_build(port, ref, depth, q, dof, bc):
    coord, dd = ftm.uniform_leaves(depth)
    ft = ftm.random_tree(coord, dd, q, dof, seed=q * 10 + dof)
    hp, hr = port.tree_create(ft), ref.tree_create(ft)
    rng = np.random.default_rng(q)
    pts = rng.uniform(-0.2 if bc else -0.05, 1.2 if bc else 1.05, size=(20000, 3))
    vp, lp, pp = port.eval_tree(hp, dof, pts, bc)
    vr, lr, pr = ref.eval_tree(hr, dof, pts, bc)
    assert np.array_equal(lp, lr)
    assert np.array_equal(vp, vr)
    assert np.array_equal(pp, pr)


def test_morton_key_order_matches_reference_comparator(port, ref):
    """unsigned compare of the interleaved key == pvfmm::MortonId::operator< (depth ties aside)."""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 1 << 15, size=(4000, 3)) / float(1 << 15)
    b = a.copy()
    flip = rng.integers(0, 3, size=4000)
    bit = rng.integers(0, 15, size=4000)
    for i in range(4000):  # neighbours differing in one bit of one axis stress the tie rules
        v = int(b[i, flip[i]] * (1 << 15)) ^ (1 << int(bit[i]))
        b[i, flip[i]] = v / float(1 << 15)
    b[:500] = rng.integers(0, 1 << 15, size=(500, 3)) / float(1 << 15)
    for i in range(4000):
        ka = port.lib.orc_point_key(a[i, 0], a[i, 1], a[i, 2], 1)
        kb = port.lib.orc_point_key(b[i, 0], b[i, 1], b[i, 2], 1)
        less = ref.lib.ref_morton_less(a[i, 0], a[i, 1], a[i, 2], 15, b[i, 0], b[i, 1], b[i, 2], 15)
        assert bool(less) == (ka < kb)


# ------------------------------------------------------------- (c) analytic properties
def test_polynomial_reproduction(port):
    """A polynomial of total degree <= q is represented exactly: evaluation error ~ eps."""
    coord, dd = ftm.uniform_leaves(2)
    q = 6

    def f(p):
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        return (1 + x - 2 * y * y + x * y * z + z ** 3 * x ** 2 + 0.5 * y ** 6)[:, None]
    ft = ftm.fit(coord, dd, q, 1, f)
    h = port.tree_create(ft)
    pts = np.random.default_rng(0).uniform(0, 1, size=(5000, 3))
    v, _, _ = port.eval_tree(h, 1, pts, 0)
    assert np.abs(v - f(pts)).max() < 1e-12


def test_rk2_rotation_third_order_local_error(port):
    """Solid-body rotation: one explicit-midpoint step has O(dt^3) error."""
    coord, dd = ftm.uniform_leaves(1)
    tv = ftm.fit(coord, dd, 3, 3, ftm.vel_rotation)
    h = port.tree_create(tv)
    x0 = np.array([[0.7, 0.5, 0.5]])
    errs = []
    for dt in (0.1, 0.05):
        x = port.traj_rk2(h, x0, 0.0, -dt, 1, 0)
        th = -dt  # backward in time
        exact = np.array([0.5 + 0.2 * np.cos(th), 0.5 + 0.2 * np.sin(th), 0.5])
        errs.append(np.abs(x[0] - exact).max())
    assert errs[0] / errs[1] > 6.0  # ~8 for third order


def test_constant_velocity_trajectory(port):
    """constant velocity 1 in x, x0 = 0.1, t 0 -> -0.01: x = 0.09."""
    coord, dd = ftm.uniform_leaves(0)
    co = np.zeros((1, 3, ftm.ncoef(2)))
    co[0, 0, 0] = 1.0
    h = port.tree_create(ftm.FlatTree(2, 3, coord, dd, co))
    x = port.traj_rk2(h, np.array([[0.1, 0.2, 0.3]]), 0.0, -0.01, 1, 0)
    assert x[0, 0] == pytest.approx(0.09, abs=1e-15) and x[0, 1] == 0.2 and x[0, 2] == 0.3


def test_new_nodes_endpoints_and_cubic_weights(port):
    for q in (3, 8, 14):
        n = port.new_nodes(q, 1).ravel()
        assert abs(n[0]) < 1e-15 and abs(n[-1] - 1.0) < 1e-15
        assert np.all(np.diff(n) > 0)
    # fast_interp reproduces per-axis cubics exactly (Lagrange on 4 nodes) and constants
    n_reg = 9
    g = np.linspace(0, 1, n_reg)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    grid = (1 + X ** 3 - 2 * Y ** 2 + Z * X)[None]
    pts = np.random.default_rng(1).uniform(0, 1, size=(300, 3))
    v = port.fast_interp(grid, 1, n_reg, pts).ravel()
    exact = 1 + pts[:, 0] ** 3 - 2 * pts[:, 1] ** 2 + pts[:, 2] * pts[:, 0]
    assert np.abs(v - exact).max() < 1e-13
    c = port.fast_interp(np.full((1, 8, 8, 8), 2.0), 1, 8, np.array([[0.1, 0.1, 0.1], [1, 1, 1.0]]))
    assert np.allclose(c, 2.0, atol=1e-14)
    assert port.fast_interp(np.full((1, 8, 8, 8), 2.0), 1, 8, np.array([[1.01, 0.1, 0.1]]))[0, 0] == 0


def test_empty_and_single_point(port):
    coord, dd = ftm.uniform_leaves(1)
    ft = ftm.random_tree(coord, dd, 3, 2, seed=1)
    h = port.tree_create(ft)
    v, leaf, _ = port.eval_tree(h, 2, np.zeros((0, 3)), 0)
    assert v.shape == (0, 2) and leaf.shape == (0,)
    v, leaf, _ = port.eval_tree(h, 2, np.array([[0.75, 0.75, 0.75]]), 0)
    assert leaf[0] == 7

"""Generate tests/golden/*.npz from the REFERENCE'S OWN code (oracle/_ref).

Run in the build container (where /root/reference exists and `make -C oracle` has built
oracle/_ref/libtbslas_ref.so):

    python tests/golden/make_golden.py

Each fixture stores the exact inputs and the outputs of tbslas::NodeFieldFunctor /
ComputeTrajRK2 / SolveSemilagRK2 / FieldSetFunctor / FieldExtrapFunctor / fast_interp /
InterpCubic1D / new_nodes as compiled from /root/reference/src over the PVFMM stand-in.
The reference ships no golden vectors of its own; these files are
the frozen pin the oracle port and the CUDA path are checked against on machines that do
not have /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402
from tbslas_b200 import flat_tree as ftm  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
R = Oracle("ref")


def tree_dict(prefix, ft):
    return {prefix + "_q": ft.q, prefix + "_dof": ft.dof, prefix + "_coord": ft.coord,
            prefix + "_depth": ft.depth, prefix + "_coeff": ft.coeff}


def adaptive_demo(max_depth=4, min_depth=1):
    def refine(lower, edge, d):
        c = lower + 0.5 * edge[:, None]
        r = np.sqrt(((c - np.array([0.6, 0.5, 0.45])) ** 2).sum(axis=1))
        return np.abs(r - 0.25) < edge  # refine near a sphere surface
    return ftm.adaptive_leaves(refine, min_depth, max_depth)


def main():
    rng = np.random.default_rng(20240917)
    # 1. known-answer tree of SURVEY.md Appendix C
    coord, depth = ftm.uniform_leaves(1)
    co = np.zeros((8, 1, ftm.ncoef(4)))
    co[:, 0, 0] = 1 + np.arange(8)
    kat = ftm.FlatTree(4, 1, coord, depth, co)
    pts = np.array([(0.1, 0.1, 0.1), (0.9, 0.1, 0.1), (1, 1, 1), (0.5, 0.5, 0.5),
                    (0.5, 0.25, 0.25), (0.4999999999999999, 0.25, 0.25), (1.0, 0.25, 0.25),
                    (0.25, 0.25, 1.0), (1.03, 0.5, 0.5), (-0.01, 0.5, 0.5), (0.25, 1.5, 0.25)])
    h = R.tree_create(kat)
    d = tree_dict("tree", kat)
    d["pts"] = pts
    for bc in (0, 1):
        v, leaf, p = R.eval_tree(h, 1, pts, bc)
        d["val_bc%d" % bc], d["leaf_bc%d" % bc], d["pos_bc%d" % bc] = v, leaf, p
    np.savez_compressed(os.path.join(OUT, "kat_depth1.npz"), **d)

    # 2. adaptive tree, dof 3, random smooth-spectrum coefficients, points in and out
    coord, depth = adaptive_demo()
    ft = ftm.random_tree(coord, depth, 6, 3, seed=11)
    h = R.tree_create(ft)
    pts = rng.uniform(-0.15, 1.15, size=(3000, 3))
    pts[:200] = rng.integers(0, 17, size=(200, 3)) / 16.0  # exact faces / corners / 1.0
    d = tree_dict("tree", ft)
    d["pts"] = pts
    for bc in (0, 1):
        v, leaf, p = R.eval_tree(h, 3, pts, bc)
        d["val_bc%d" % bc], d["leaf_bc%d" % bc], d["pos_bc%d" % bc] = v, leaf, p
    np.savez_compressed(os.path.join(OUT, "eval_adaptive_q6.npz"), **d)

    # 3. trajectories + semi-Lagrangian step, rotation velocity, Gaussian scalar
    coord, depth = ftm.uniform_leaves(2)
    tv = ftm.fit(coord, depth, 5, 3, lambda p: ftm.vel_rotation(p))
    tc = ftm.fit(coord, depth, 5, 1, lambda p: ftm.gaussian(p, sigma=0.15))
    hv, hc = R.tree_create(tv), R.tree_create(tc)
    pts = rng.uniform(0.0, 1.0, size=(800, 3))
    d = {**tree_dict("vel", tv), **tree_dict("con", tc), "pts": pts,
         "dt": 0.0628, "timestep": 3, "nrk": 2}
    for bc in (0, 1):
        d["traj_bc%d" % bc] = R.traj_rk2(hv, pts, 3 * 0.0628, 2 * 0.0628, 2, bc)
        d["semilag_bc%d" % bc] = R.semilag_rk2(hv, hc, 1, pts, 3, 0.0628, 2, bc)
    np.savez_compressed(os.path.join(OUT, "semilag_rotation_q5.npz"), **d)

    # 4. time-varying velocity: 4 snapshots (cubic in time) and extrapolation
    coord, depth = adaptive_demo(3, 1)
    times = np.array([-0.1, 0.0, 0.1, 0.2])
    snaps = [ftm.random_tree(coord, depth, 4, 3, seed=20 + i, scale=0.3) for i in range(4)]
    hs = [R.tree_create(s) for s in snaps]
    pts = rng.uniform(-0.05, 1.05, size=(600, 3))
    d = {"times": times, "tq": 0.037, "pts": pts, "q": 4, "dof": 3, "coord": coord,
         "depth": depth, "coeff4": np.stack([s.coeff for s in snaps])}
    for bc in (0, 1):
        d["set4_bc%d" % bc], _ = R.eval_set4(hs, times, 0.037, 3, pts, bc)
        d["extrap_bc%d" % bc], _ = R.eval_extrap(hs[0], hs[1], 3, pts, bc)
        d["traj_set4_bc%d" % bc] = R.traj_rk2(hs, pts, 0.1, 0.0, 1, bc, kind="set4", times=times)
        d["traj_extrap_bc%d" % bc] = R.traj_rk2((hs[0], hs[1]), pts, 0.1, 0.0, 1, bc, kind="extrap")
    np.savez_compressed(os.path.join(OUT, "timevarying_q4.npz"), **d)

    # 5. uniform-grid cubic interpolation
    n_reg, dof = 12, 2
    grid = rng.standard_normal((dof, n_reg, n_reg, n_reg))
    pts = rng.uniform(-0.05, 1.05, size=(700, 3))
    pts[:50] = rng.integers(0, 12, size=(50, 3)) / 11.0
    np.savez_compressed(os.path.join(OUT, "cubic_grid_n12.npz"), grid=grid, n_reg=n_reg, dof=dof,
                        pts=pts, val=R.fast_interp(grid, dof, n_reg, pts))

    # 6. scalar helpers
    xx = np.array([0.0, 0.1, 0.25, 0.3])
    pp = rng.standard_normal((40, 4))
    xs = rng.uniform(0.1, 0.25, size=40)
    cub = np.array([R.interp_cubic1d(xs[i], xx, pp[i]) for i in range(40)])
    d = {"cubic_xx": xx, "cubic_pp": pp, "cubic_x": xs, "cubic_val": cub}
    for q in range(1, 17):
        d["nodes_q%d" % q] = R.new_nodes(q, 1).ravel()
    d["nodes3_q3"] = R.new_nodes(3, 3)
    np.savez_compressed(os.path.join(OUT, "scalar_helpers.npz"), **d)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()

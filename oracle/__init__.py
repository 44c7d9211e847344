"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

``Oracle("port")``  -> oracle/liboracle.so        (plain-C restatement, tbslas_oracle.c)
``Oracle("ref")``   -> oracle/_ref/libtbslas_ref.so (the reference's own headers compiled
                       over the PVFMM/MPI stand-in; see ref_capi.cpp)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may import
this package.  Nothing under tbslas_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libtbslas_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_ubyte)


def build(quiet: bool = True) -> None:
    """make -C oracle (C port always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _d(a):
    return a.ctypes.data_as(_dp)


class _Field(C.Structure):
    _fields_ = [("kind", C.c_int), ("tree", C.c_void_p * 4), ("times", C.c_double * 4)]


class Oracle:
    """Uniform front for both checkers.  ``bc``: 0 FreeSpace, 1 Periodic."""

    def __init__(self, kind: str = "port"):
        assert kind in ("port", "ref")
        self.kind = kind
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            if kind == "port":
                build()
            else:
                raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        if kind == "port":
            L.orc_tree_create.restype = C.c_void_p
            L.orc_tree_create.argtypes = [C.c_int, C.c_int, C.c_long, _dp, _u8p, _dp]
            L.orc_tree_destroy.argtypes = [C.c_void_p]
            L.orc_tree_is_sorted.argtypes = [C.c_void_p]
            L.orc_eval_tree.argtypes = [C.c_void_p, C.c_int, _dp, C.c_long, _dp, _ip]
            L.orc_interp_cubic1d.restype = C.c_double
            L.orc_interp_cubic1d.argtypes = [C.c_double, _dp, _dp]
            L.orc_eval_set4.argtypes = [C.POINTER(C.c_void_p), _dp, C.c_double, C.c_int, _dp,
                                        C.c_long, _dp]
            L.orc_eval_extrap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, C.c_long, _dp]
            L.orc_traj_rk2.argtypes = [C.POINTER(_Field), C.POINTER(_Field), C.c_int, _dp,
                                       C.c_long, C.c_double, C.c_double, C.c_int, _dp]
            L.orc_semilag_rk2.argtypes = [C.POINTER(_Field), C.POINTER(_Field), C.c_void_p,
                                          C.c_int, _dp, C.c_long, C.c_int, C.c_double, C.c_int,
                                          _dp, _dp]
            L.orc_fast_interp.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_long, _dp]
            L.orc_new_nodes.restype = C.c_long
            L.orc_new_nodes.argtypes = [C.c_int, C.c_int, _dp]
            L.orc_collect_grid_points.argtypes = [C.c_void_p, _dp]
            L.orc_point_key.restype = C.c_uint64
            L.orc_point_key.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        else:
            L.ref_tree_create.restype = C.c_void_p
            L.ref_tree_create.argtypes = [C.c_int, C.c_int, C.c_long, _dp, _u8p, _dp]
            L.ref_tree_destroy.argtypes = [C.c_void_p]
            L.ref_eval_tree.argtypes = [C.c_void_p, _dp, C.c_long, _dp]
            L.ref_leaf_index.argtypes = [C.c_void_p, _dp, C.c_long, _ip]
            L.ref_traj_rk2.argtypes = [C.c_void_p, _dp, C.c_long, C.c_double, C.c_double,
                                       C.c_int, _dp]
            L.ref_semilag_rk2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, C.c_long,
                                          C.c_int, C.c_double, C.c_int, _dp]
            L.ref_eval_set4.argtypes = [C.POINTER(C.c_void_p), _dp, C.c_double, _dp, C.c_long,
                                        _dp]
            L.ref_eval_extrap.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_long, _dp]
            L.ref_traj_rk2_set4.argtypes = [C.POINTER(C.c_void_p), _dp, _dp, C.c_long,
                                            C.c_double, C.c_double, C.c_int, _dp]
            L.ref_semilag_rk2_set4.argtypes = [C.POINTER(C.c_void_p), _dp, C.c_void_p, C.c_int,
                                               _dp, C.c_long, C.c_int, C.c_double, C.c_int, _dp]
            L.ref_traj_rk2_extrap.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_long,
                                              C.c_double, C.c_double, C.c_int, _dp]
            L.ref_fast_interp.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_long, _dp]
            L.ref_interp_cubic1d.restype = C.c_double
            L.ref_interp_cubic1d.argtypes = [C.c_double, _dp, _dp]
            L.ref_new_nodes.restype = C.c_long
            L.ref_new_nodes.argtypes = [C.c_int, C.c_int, _dp]
            L.ref_pt2coeff.argtypes = [C.c_int, _dp]
            L.ref_morton_less.argtypes = [C.c_double] * 3 + [C.c_int] + [C.c_double] * 3 + [C.c_int]
            L.ref_morton_xyz.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_uint)]

    # ------------------------------------------------------------------ misc
    def set_num_threads(self, n: int) -> None:
        (self.lib.orc_set_num_threads if self.kind == "port" else self.lib.ref_set_num_threads)(n)

    def max_threads(self) -> int:
        return (self.lib.orc_get_max_threads if self.kind == "port"
                else self.lib.ref_get_max_threads)()

    # ------------------------------------------------------------------ trees
    def tree_create(self, ft):
        fn = self.lib.orc_tree_create if self.kind == "port" else self.lib.ref_tree_create
        h = fn(ft.q, ft.dof, ft.n_leaf, _d(ft.coord), ft.depth.ctypes.data_as(_u8p), _d(ft.coeff))
        assert h, "tree_create failed"
        return C.c_void_p(h)

    def tree_destroy(self, h) -> None:
        (self.lib.orc_tree_destroy if self.kind == "port" else self.lib.ref_tree_destroy)(h)

    # ------------------------------------------------------------- evaluation
    def eval_tree(self, h, dof: int, pos: np.ndarray, bc: int, want_leaf: bool = True):
        """-> (values [n,dof], leaf_idx [n] int32 or None, pos_after [n,3]).
        pos is copied; the copy is wrapped in place when bc is periodic."""
        p = np.array(pos, dtype=np.float64, order="C").reshape(-1, 3)
        n = p.shape[0]
        out = np.empty((n, dof))
        leaf = np.empty(n, dtype=np.int32) if want_leaf else None
        if self.kind == "port":
            self.lib.orc_eval_tree(h, bc, _d(p), n, _d(out),
                                   leaf.ctypes.data_as(_ip) if want_leaf else None)
        else:
            self.lib.ref_set_bc(bc)
            self.lib.ref_eval_tree(h, _d(p), n, _d(out))
            if want_leaf:
                self.lib.ref_leaf_index(h, _d(p), n, leaf.ctypes.data_as(_ip))
        return out, leaf, p

    def _field(self, kind, trees, times=None):
        f = _Field()
        f.kind = kind
        for i, t in enumerate(trees):
            f.tree[i] = t.value
        if times is not None:
            for i in range(4):
                f.times[i] = times[i]
        return f

    def _harr(self, trees):
        return (C.c_void_p * 4)(*[t.value for t in trees])

    def eval_set4(self, trees, times, tq, dof, pos, bc):
        p = np.array(pos, dtype=np.float64, order="C").reshape(-1, 3)
        n = p.shape[0]
        out = np.empty((n, dof))
        tt = np.asarray(times, dtype=np.float64)
        if self.kind == "port":
            self.lib.orc_eval_set4(self._harr(trees), _d(tt), tq, bc, _d(p), n, _d(out))
        else:
            self.lib.ref_set_bc(bc)
            self.lib.ref_eval_set4(self._harr(trees), _d(tt), tq, _d(p), n, _d(out))
        return out, p

    def eval_extrap(self, tp, tc, dof, pos, bc):
        p = np.array(pos, dtype=np.float64, order="C").reshape(-1, 3)
        n = p.shape[0]
        out = np.empty((n, dof))
        if self.kind == "port":
            self.lib.orc_eval_extrap(tp, tc, bc, _d(p), n, _d(out))
        else:
            self.lib.ref_set_bc(bc)
            self.lib.ref_eval_extrap(tp, tc, _d(p), n, _d(out))
        return out, p

    def traj_rk2(self, vel, pos, tinit, tfinal, nrk, bc, kind="steady", times=None):
        """vel: tree handle (steady), 4 handles (set4), (tp, tc) (extrap: first stage samples tc,
        second the extrapolation) or (t1, t2) (pair: first stage samples t1, second t2 -- the
        second trajectory of the NS call pattern, tree_ns.h:479-483; port only)."""
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        n = p.shape[0]
        out = np.empty((n, 3))
        if self.kind == "port":
            if kind == "steady":
                f1, f2 = self._field(0, [vel]), None
            elif kind == "set4":
                f1, f2 = self._field(1, vel, times), None
            elif kind == "pair":
                f1, f2 = self._field(0, [vel[0]]), self._field(0, [vel[1]])
            else:  # first stage samples tc, second the extrapolation (traj.inc:71-92)
                f1, f2 = self._field(0, [vel[1]]), self._field(2, [vel[0], vel[1]])
            self.lib.orc_traj_rk2(C.byref(f1), C.byref(f2) if f2 else None, bc, _d(p), n,
                                  tinit, tfinal, nrk, _d(out))
        else:
            self.lib.ref_set_bc(bc)
            if kind == "steady":
                self.lib.ref_traj_rk2(vel, _d(p), n, tinit, tfinal, nrk, _d(out))
            elif kind == "set4":
                tt = np.asarray(times, dtype=np.float64)
                self.lib.ref_traj_rk2_set4(self._harr(vel), _d(tt), _d(p), n, tinit, tfinal,
                                           nrk, _d(out))
            else:
                self.lib.ref_traj_rk2_extrap(vel[0], vel[1], _d(p), n, tinit, tfinal, nrk,
                                             _d(out))
        return out

    def semilag_rk2(self, vel, con, dof_con, pos, timestep, dt, nrk, bc, kind="steady",
                    times=None):
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        n = p.shape[0]
        out = np.empty((n, dof_con))
        if self.kind == "port":
            if kind == "steady":
                f1 = self._field(0, [vel])
            else:
                f1 = self._field(1, vel, times)
            self.lib.orc_semilag_rk2(C.byref(f1), None, con, bc, _d(p), n, timestep, dt, nrk,
                                     _d(out), None)
        else:
            self.lib.ref_set_bc(bc)
            if kind == "steady":
                self.lib.ref_semilag_rk2(vel, con, dof_con, _d(p), n, timestep, dt, nrk, _d(out))
            else:
                tt = np.asarray(times, dtype=np.float64)
                self.lib.ref_semilag_rk2_set4(self._harr(vel), _d(tt), con, dof_con, _d(p), n,
                                              timestep, dt, nrk, _d(out))
        return out

    # ------------------------------------------------------------------ cubic
    def fast_interp(self, grid, dof, n_reg, pts):
        g = np.ascontiguousarray(grid, dtype=np.float64).reshape(-1)
        p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        out = np.empty((p.shape[0], dof))
        fn = self.lib.orc_fast_interp if self.kind == "port" else self.lib.ref_fast_interp
        fn(_d(g), dof, n_reg, _d(p), p.shape[0], _d(out))
        return out

    def interp_cubic1d(self, x, xx, pp):
        a = np.asarray(xx, dtype=np.float64).copy()
        b = np.asarray(pp, dtype=np.float64).copy()
        fn = self.lib.orc_interp_cubic1d if self.kind == "port" else self.lib.ref_interp_cubic1d
        return fn(x, _d(a), _d(b))

    def new_nodes(self, q, dim):
        fn = self.lib.orc_new_nodes if self.kind == "port" else self.lib.ref_new_nodes
        n = fn(q, dim, None)
        out = np.empty(n)
        fn(q, dim, _d(out))
        return out.reshape(-1, dim)

"""Host-side mirror of the reference's operator surface over the C ABI.

Names, argument meaning and call pattern follow arashb/tbslas so tests read like the
reference's drivers (e.g. src/applications/src/advection.cpp:159-181,296):

    ctx   = Context(device=0)
    tvel  = ctx.tree(flat_velocity_tree)              # ConstructTree(...) equivalent input
    tcon  = ctx.tree(flat_scalar_tree)
    vel   = NodeFieldFunctor(tvel)                    # tree_functor.h:793-815
    dep   = ComputeTrajRK2(vel, pos, tinit, tfinal, nrk, bc)          # traj.inc:49-68
    vals  = SolveSemilagRK2(vel, NodeFieldFunctor(tcon), pos, timestep, dt, nrk, bc)

Every call goes through ``libtbslas_b200.so``; buffers may be numpy arrays (host: copied
in/out inside the call) or torch CUDA tensors (device resident, stream ordered).  The
reference reads the boundary condition from a global SimConfig singleton; here ``bc`` is
an explicit argument (0 FreeSpace, 1 Periodic).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import (FIELD_EXTRAP, FIELD_SET4, FIELD_STEADY, FREESPACE, MEM_DEVICE, MEM_HOST,
                   PERIODIC, TbslasError)
from .flat_tree import FlatTree

try:  # torch is plumbing only (device buffers, streams); the library does not need it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(a) -> bool:
    return torch is not None and isinstance(a, torch.Tensor)


def _addr(a, dtype="f64"):
    """-> (address, mem) of a C-contiguous numpy array or torch CUDA tensor of the element type
    the C ABI reads ("f64": double, "i32": int32_t).  A float32 array handed to a double* entry
    point would be read past its end, so the type is checked here, loudly."""
    if a is None:
        return None, None
    if _is_torch(a):
        want = torch.float64 if dtype == "f64" else torch.int32
        if not (a.is_cuda and a.is_contiguous()):
            raise TypeError("device buffers must be contiguous CUDA tensors")
        if a.dtype != want:
            raise TypeError("device buffer has dtype %s, the C ABI expects %s" % (a.dtype, want))
        return a.data_ptr(), MEM_DEVICE
    want = np.float64 if dtype == "f64" else np.int32
    if not (isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]):
        raise TypeError("host buffers must be C-contiguous numpy arrays")
    if a.dtype != want:
        raise TypeError("host buffer has dtype %s, the C ABI expects %s" % (a.dtype, np.dtype(want)))
    return a.ctypes.data, MEM_HOST


def _like(ref, shape, dtype="f64"):
    if _is_torch(ref):
        return torch.empty(shape, dtype=torch.float64 if dtype == "f64" else torch.int32,
                           device=ref.device)
    return np.empty(shape, dtype=np.float64 if dtype == "f64" else np.int32)


class Context:
    """One per GPU per process (tbslas_b200_init)."""

    def __init__(self, device: int = 0, stream=None):
        self.lib = capi.load()
        h = C.c_void_p()
        rc = self.lib.tbslas_b200_init(device, C.byref(h))
        if rc != capi.OK:
            raise TbslasError("tbslas_b200_init(device=%d) failed with code %d: no usable sm_100 "
                              "device (there is no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def check(self, rc: int) -> None:
        if rc != capi.OK:
            raise TbslasError("tbslas_b200 error %d: %s" % (
                rc, self.lib.tbslas_b200_last_error(self.h).decode()))

    def set_stream(self, stream) -> None:
        """stream: raw cudaStream_t address, a torch.cuda.Stream, or None (own stream)."""
        if stream is not None and hasattr(stream, "cuda_stream"):
            stream = stream.cuda_stream or 1  # torch's default stream is handle 0 = cudaStreamLegacy (1)
        self.check(self.lib.tbslas_b200_set_stream(self.h, stream))

    def set_time_combine(self, on: bool) -> None:
        """FieldSetFunctor/FieldExtrapFunctor over trees with one leaf list: combine the
        coefficients in time and evaluate once (True, default) or evaluate every tree and combine
        the values per point in the reference's order (False)."""
        self.check(self.lib.tbslas_b200_set_time_combine(self.h, int(bool(on))))

    def set_tensor_grid(self, on: bool) -> None:
        """Tree-level calls: evaluate the velocity at the arrival grids by sum factorisation
        (True, default) or point by point (False)."""
        self.check(self.lib.tbslas_b200_set_tensor_grid(self.h, 2 if on == "always" else int(bool(on))))

    def last_grid_exceptions(self) -> int:
        n = C.c_size_t()
        self.check(self.lib.tbslas_b200_last_grid_exceptions(self.h, C.byref(n)))
        return int(n.value)

    def synchronize(self) -> None:
        self.check(self.lib.tbslas_b200_synchronize(self.h))

    def close(self) -> None:
        if self.h:
            self.lib.tbslas_b200_finalize(self.h)
            self.h = None

    # -- distributed ------------------------------------------------------
    def comm_init_torch(self, group=None) -> None:
        """Create the library's NCCL communicator; the 128-byte unique id travels over
        the caller's torch.distributed group (any backend)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            self.check(self.lib.tbslas_b200_comm_unique_id(uid))
        t = torch.tensor(list(uid), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(self.device)
        dist.broadcast(t, src=0, group=group)
        buf = (C.c_ubyte * 128)(*t.cpu().tolist())
        self.check(self.lib.tbslas_b200_comm_init(self.h, world, rank, buf))

    def comm_last_exchange(self):
        """(sent, received) outsider points of the most recent tree evaluation."""
        a, b = C.c_size_t(), C.c_size_t()
        self.check(self.lib.tbslas_b200_comm_last_exchange(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def comm_set_exchange(self, mode: str) -> None:
        """"peer": NVLink peer-memory mailboxes (default where available); "nccl": all-to-all-v."""
        self.check(self.lib.tbslas_b200_comm_set_exchange(self.h, {"peer": 1, "nccl": 0}[mode]))

    def comm_set_mailbox(self, points: int) -> None:
        self.check(self.lib.tbslas_b200_comm_set_mailbox(self.h, int(points)))

    def comm_exchange_mode(self):
        """-> ("peer" | "nccl", mailbox capacity in points)."""
        m, c = C.c_int(), C.c_size_t()
        self.check(self.lib.tbslas_b200_comm_exchange_mode(self.h, C.byref(m), C.byref(c)))
        return ("peer" if m.value else "nccl"), int(c.value)

    def set_virtual_arrival_points(self, on: bool) -> None:
        """Tree-level calls: rebuild the arrival points where the step needs them again instead of
        writing them to HBM (default on)."""
        self.check(self.lib.tbslas_b200_set_virtual_arrival_points(self.h, int(bool(on))))

    def set_host_chunks(self, chunks: int) -> None:
        self.check(self.lib.tbslas_b200_set_host_chunks(self.h, int(chunks)))

    def comm_rank(self):
        r, n = C.c_int(), C.c_int()
        self.check(self.lib.tbslas_b200_comm_rank(self.h, C.byref(r), C.byref(n)))
        return r.value, n.value

    # -- trees --------------------------------------------------------------
    def tree(self, ft: FlatTree, replicated: bool = False) -> "Tree":
        """replicated=True (multi-rank contexts): ``ft`` is the WHOLE tree on every rank and
        evaluations on it never exchange points (tbslas_b200_tree_create_replicated)."""
        return Tree(self, ft.q, ft.dof, ft.coord, ft.depth, ft.coeff, replicated)

    def set_pt2coeff(self, q: int, M=None) -> None:
        """Upload the point-to-coefficient matrix of degree q (cheb.h:166-196); by default the
        harness' numpy pseudo-inverse (flat_tree.pt2coeff)."""
        if M is None:
            from . import flat_tree as ftm
            M = ftm.pt2coeff(q)
        M = np.ascontiguousarray(M, dtype=np.float64)
        assert M.shape == ((q + 1) ** 3, (q + 1) * (q + 2) * (q + 3) // 6)
        self.check(self.lib.tbslas_b200_set_pt2coeff(self.h, int(q), M.ctypes.data))

    # -- cubic grid (tbslas::fast_interp) ----------------------------------
    def fast_interp(self, grid, dof: int, n_reg: int, pts, out=None):
        n = pts.shape[0]
        if out is None:
            out = _like(pts, (n, dof))
        ga, gm = _addr(grid)
        pa, pm = _addr(pts)
        oa, _ = _addr(out)
        assert gm == pm, "grid and points must live in the same memory space"
        self.check(self.lib.tbslas_b200_cubic_eval(self.h, ga, n_reg, dof, pa, n, oa, pm))
        return out

    def grid(self, grid, dof: int, n_reg: int) -> "CubicGrid":
        """A uniform grid [dof][n_reg]^3 kept resident in HBM (tbslas_b200_grid_create)."""
        return CubicGrid(self, grid, dof, n_reg)

    # -- instrumentation ------------------------------------------------------
    def profile_enable(self, on: bool = True) -> None:
        self.check(self.lib.tbslas_b200_profile_enable(self.h, int(on)))

    def profile_reset(self) -> None:
        self.check(self.lib.tbslas_b200_profile_reset(self.h))

    def profile(self) -> dict:
        out = {}
        for s in range(self.lib.tbslas_b200_profile_num_stages()):
            ms, ln, un = C.c_double(), C.c_longlong(), C.c_double()
            self.check(self.lib.tbslas_b200_profile_get(self.h, s, C.byref(ms), C.byref(ln),
                                                        C.byref(un)))
            out[self.lib.tbslas_b200_profile_stage_name(s).decode()] = {
                "ms": ms.value, "launches": ln.value, "units": un.value}
        return out

    def kernel_launches(self) -> int:
        return int(self.lib.tbslas_b200_kernel_launches(self.h))

    def fp64_peak(self, reps: int = 5) -> float:
        v = C.c_double()
        self.check(self.lib.tbslas_b200_fp64_peak(self.h, reps, C.byref(v)))
        return v.value


class CubicGrid:
    """Resident uniform grid for tbslas::fast_interp (tree_functor.h:89-153)."""

    def __init__(self, ctx: Context, grid, dof: int, n_reg: int):
        self.ctx, self.dof, self.n_reg = ctx, int(dof), int(n_reg)
        a, m = _addr(grid)
        h = C.c_void_p()
        ctx.check(ctx.lib.tbslas_b200_grid_create(ctx.h, a, self.n_reg, self.dof, m, C.byref(h)))
        self.h = h

    def update(self, grid) -> None:
        a, m = _addr(grid)
        self.ctx.check(self.ctx.lib.tbslas_b200_grid_update(self.h, a, m))

    def __call__(self, pts, out=None):
        n = pts.shape[0]
        if out is None:
            out = _like(pts, (n, self.dof))
        pa, pm = _addr(pts)
        self.ctx.check(self.ctx.lib.tbslas_b200_grid_eval(self.h, pa, n, _addr(out)[0], pm))
        return out

    def destroy(self) -> None:
        if self.h:
            self.ctx.lib.tbslas_b200_grid_destroy(self.h)
            self.h = None


class Tree:
    """Device-resident leaf list of one Chebyshev octree (tbslas_b200_tree_create)."""

    def __init__(self, ctx: Context, q, dof, coord, depth, coeff, replicated: bool = False):
        self.ctx, self.q, self.dof = ctx, int(q), int(dof)
        coord = np.ascontiguousarray(coord, dtype=np.float64)
        depth = np.ascontiguousarray(depth, dtype=np.uint8)
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        self.n_leaf = coord.shape[0]
        h = C.c_void_p()
        create = ctx.lib.tbslas_b200_tree_create_replicated if replicated else ctx.lib.tbslas_b200_tree_create
        ctx.check(create(ctx.h, self.q, self.dof, self.n_leaf,
                                                  coord.ctypes.data, depth.ctypes.data,
                                                  coeff.ctypes.data, MEM_HOST, C.byref(h)))
        self.h = h

    def update_coeff(self, coeff, wait: bool = True) -> None:
        """New coefficients on the same leaves.  wait=False: tbslas_b200_tree_update_coeff_async --
        the copy overlaps what is enqueued next; the buffer (pinned, if it is to overlap) must stay
        valid until a synchronising call returns."""
        a, m = _addr(coeff)
        fn = self.ctx.lib.tbslas_b200_tree_update_coeff if wait else self.ctx.lib.tbslas_b200_tree_update_coeff_async
        self.ctx.check(fn(self.h, a, m))

    def destroy(self) -> None:
        if self.h:
            self.ctx.lib.tbslas_b200_tree_destroy(self.h)
            self.h = None

    def reshard(self, new_first) -> None:
        """Collective: move leaves between ranks so that rank r owns global leaves
        [new_first[r], new_first[r+1]) (tbslas_b200_tree_reshard)."""
        arr = (C.c_size_t * len(new_first))(*[int(x) for x in new_first])
        self.ctx.check(self.ctx.lib.tbslas_b200_tree_reshard(self.h, arr))
        q, dof, n = C.c_int(), C.c_int(), C.c_size_t()
        self.ctx.check(self.ctx.lib.tbslas_b200_tree_info(self.h, C.byref(q), C.byref(dof), C.byref(n)))
        self.n_leaf = int(n.value)

    def global_range(self):
        """-> (global index of this rank's first leaf, leaves over all ranks)."""
        a, b = C.c_size_t(), C.c_size_t()
        self.ctx.check(self.ctx.lib.tbslas_b200_tree_global_range(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_grid_values(self, vals, point_major: bool = False) -> None:
        """tbslas::SetTreeGridValues (tree_utils.h:500-552): refit the coefficients from grid
        values ([leaf][dof][P], or [leaf*P][dof] when point_major)."""
        a, m = _addr(vals)
        self.ctx.check(self.ctx.lib.tbslas_b200_tree_set_grid_values(self.h, a, int(point_major), m))

    def coefficients(self) -> np.ndarray:
        """Read the device coefficients back: [n_leaf, dof, Ncoef] (tests)."""
        nc = (self.q + 1) * (self.q + 2) * (self.q + 3) // 6
        out = np.empty((self.n_leaf, self.dof, nc))
        self.ctx.check(self.ctx.lib.tbslas_b200_tree_get_coeff(self.h, out.ctypes.data, MEM_HOST))
        return out

    def last_point_counts(self) -> np.ndarray:
        """Points the most recent evaluation of this tree located in each local leaf."""
        out = np.zeros(self.n_leaf, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.tbslas_b200_tree_last_point_counts(self.h, out.ctypes.data, MEM_HOST))
        return out

    def tail_norm(self) -> np.ndarray:
        """l2 norm of every leaf's highest-degree coefficients (i+j+k == q, all dof)."""
        out = np.zeros(self.n_leaf)
        self.ctx.check(self.ctx.lib.tbslas_b200_tree_tail_norm(self.h, out.ctypes.data, MEM_HOST))
        return out

    def collect_grid_points(self, device: bool = False):
        """tbslas::CollectChebTreeGridPoints: [n_leaf*(q+1)^3, 3]."""
        n = self.n_leaf * (self.q + 1) ** 3
        out = (torch.empty((n, 3), dtype=torch.float64, device="cuda:%d" % self.ctx.device)
               if device else np.empty((n, 3)))
        a, m = _addr(out)
        self.ctx.check(self.ctx.lib.tbslas_b200_collect_grid_points(self.h, a, m))
        return out


class _Functor:
    """Common shell of the three field functors; ``field`` is the C struct."""

    ctx: Context
    dof: int
    field: capi.Field

    def __call__(self, points_pos, num_points: Optional[int] = None, time: float = 0.0,
                 out=None, bc: int = FREESPACE):
        """f(pos[n,3], n, [t], out[n,dof]) as the reference's functor concept
        (tree_functor.h:800-811).  Periodic bc wraps ``points_pos`` in place."""
        n = points_pos.shape[0] if num_points is None else int(num_points)
        if out is None:
            out = _like(points_pos, (n, self.dof))
        pa, pm = _addr(points_pos)
        oa, om = _addr(out)
        assert pm == om
        self.ctx.check(self.ctx.lib.tbslas_b200_eval_field(C.byref(self.field), float(time), bc,
                                                           pa, n, oa, pm))
        return out


class NodeFieldFunctor(_Functor):
    """tbslas::NodeFieldFunctor (tree_functor.h:793-815)."""

    def __init__(self, tree: Tree):
        self.tree, self.ctx, self.dof = tree, tree.ctx, tree.dof
        self.field = capi.Field()
        self.field.kind = FIELD_STEADY
        self.field.tree[0] = tree.h.value

    def eval_with_leaf(self, points_pos, bc: int = FREESPACE):
        """-> (values, leaf_idx): also reports which leaf evaluated each point."""
        n = points_pos.shape[0]
        out = _like(points_pos, (n, self.dof))
        leaf = _like(points_pos, (n,), "i32")
        pa, pm = _addr(points_pos)
        self.ctx.check(self.ctx.lib.tbslas_b200_eval(self.tree.h, bc, pa, n, _addr(out)[0],
                                                     _addr(leaf, "i32")[0], pm))
        return out, leaf


class FieldSetFunctor(_Functor):
    """tbslas::FieldSetFunctor (tree_set_functor.h:27-97): 4 trees, cubic in time."""

    def __init__(self, trees: Sequence[Tree], times: Sequence[float]):
        assert len(trees) == 4 and len(times) == 4
        self.trees, self.times = list(trees), list(times)
        self.ctx, self.dof = trees[0].ctx, trees[0].dof
        self._refresh()

    def _refresh(self):
        self.field = capi.Field()
        self.field.kind = FIELD_SET4
        for i in range(4):
            self.field.tree[i] = self.trees[i].h.value
            self.field.times[i] = self.times[i]

    def update(self, new_tree: Tree, time: float) -> None:
        """Slide the window (tree_set_functor.h:81-90); the oldest tree is destroyed."""
        self.trees.pop(0).destroy()
        self.times.pop(0)
        self.trees.append(new_tree)
        self.times.append(time)
        self._refresh()


class FieldExtrapFunctor(_Functor):
    """tbslas::FieldExtrapFunctor (tree_extrap_functor.h:27-92): 1.5 v(tc) - 0.5 v(tp)."""

    def __init__(self, tp: Tree, tc: Tree):
        self.tp, self.tc = tp, tc
        self.ctx, self.dof = tp.ctx, tp.dof
        self._refresh()

    def _refresh(self):
        self.field = capi.Field()
        self.field.kind = FIELD_EXTRAP
        self.field.tree[0] = self.tp.h.value
        self.field.tree[1] = self.tc.h.value

    def update(self, new_tree: Tree, time: float = 0.0) -> None:
        self.tp.destroy()
        self.tp, self.tc = self.tc, new_tree
        self._refresh()


def ComputeTrajRK2(field_fn: _Functor, xinit, tinit: float, tfinal: float, num_rk_step: int,
                   bc: int = FREESPACE, extrap_fn: Optional[_Functor] = None, xsol=None):
    """tbslas::ComputeTrajRK2 (traj.inc:49-68; two-functor form :95-115)."""
    n = xinit.shape[0]
    if xsol is None:
        xsol = _like(xinit, (n, 3))
    pa, pm = _addr(xinit)
    ctx = field_fn.ctx
    ctx.check(ctx.lib.tbslas_b200_traj_rk2(
        C.byref(field_fn.field), C.byref(extrap_fn.field) if extrap_fn is not None else None, bc,
        pa, n, float(tinit), float(tfinal), int(num_rk_step), _addr(xsol)[0], pm))
    return xsol


def SolveSemilagRK2(vel_evaluator: _Functor, con_evaluator: NodeFieldFunctor, points_pos,
                    timestep: int, dt: float, num_rk_step: int, bc: int = FREESPACE,
                    extrap_evaluator: Optional[_Functor] = None, points_vals=None,
                    departure_points=None):
    """tbslas::SolveSemilagRK2 (semilag.inc:27-45, :49-69)."""
    n = points_pos.shape[0]
    if points_vals is None:
        points_vals = _like(points_pos, (n, con_evaluator.dof))
    pa, pm = _addr(points_pos)
    ctx = vel_evaluator.ctx
    ctx.check(ctx.lib.tbslas_b200_semilag_rk2(
        C.byref(vel_evaluator.field),
        C.byref(extrap_evaluator.field) if extrap_evaluator is not None else None,
        con_evaluator.tree.h, bc, pa, n, int(timestep), float(dt), int(num_rk_step),
        _addr(points_vals)[0], _addr(departure_points)[0], pm))
    return points_vals


def SolveSemilagInSitu(tvel_func: _Functor, tree_curr: Tree, timestep: int, dt: float,
                       num_rk_step: int = 1, bc: int = FREESPACE,
                       tvel_extrap: Optional[_Functor] = None, device: bool = False,
                       departure_points: bool = False, out=None):
    """Steps (1)+(2) of tbslas::SolveSemilagInSitu (tree_semilag.h:92-130): the arrival points
    are generated in HBM from ``tree_curr``'s own leaves and advected; returns the new grid
    values [n_leaf*(q+1)^3, dof] (leaf-major), the input of SetTreeGridValues -- and, with
    ``departure_points``, also the departure points [n_leaf*(q+1)^3, 3]."""
    n = tree_curr.n_leaf * (tree_curr.q + 1) ** 3
    dev = "cuda:%d" % tree_curr.ctx.device
    if out is None:
        out = (torch.empty((n, tree_curr.dof), dtype=torch.float64, device=dev)
               if device else np.empty((n, tree_curr.dof)))
    a, m = _addr(out)
    ctx = tree_curr.ctx
    f1 = C.byref(tvel_func.field)
    f2 = C.byref(tvel_extrap.field) if tvel_extrap is not None else None
    if not departure_points:
        ctx.check(ctx.lib.tbslas_b200_semilag_insitu(f1, f2, tree_curr.h, bc, int(timestep), float(dt),
                                                     int(num_rk_step), a, m))
        return out
    dep = (torch.empty((n, 3), dtype=torch.float64, device=dev) if m == MEM_DEVICE else np.empty((n, 3)))
    ctx.check(ctx.lib.tbslas_b200_semilag_insitu_dep(f1, f2, tree_curr.h, bc, int(timestep), float(dt),
                                                     int(num_rk_step), a, _addr(dep)[0], m))
    return out, dep


def SolveSemilagInSituUpdate(tvel_func: _Functor, tree_curr: Tree, timestep: int, dt: float,
                             num_rk_step: int = 1, bc: int = FREESPACE,
                             tvel_extrap: Optional[_Functor] = None) -> None:
    """The whole tbslas::SolveSemilagInSitu on the device: tree_curr's coefficients are
    replaced by the advected field's (needs Context.set_pt2coeff(q))."""
    ctx = tree_curr.ctx
    ctx.check(ctx.lib.tbslas_b200_semilag_insitu_update(
        C.byref(tvel_func.field), C.byref(tvel_extrap.field) if tvel_extrap is not None else None,
        tree_curr.h, bc, int(timestep), float(dt), int(num_rk_step)))


def partition_leaves_weighted(weight, nranks: int) -> np.ndarray:
    """Contiguous Morton ranges of about equal total weight: first[r] .. first[r+1]."""
    w = np.ascontiguousarray(weight, dtype=np.float64)
    first = (C.c_size_t * (nranks + 1))()
    rc = capi.load().tbslas_b200_partition_leaves_weighted(
        w.shape[0], w.ctypes.data_as(C.POINTER(C.c_double)), int(nranks), first)
    if rc != capi.OK:
        raise TbslasError("partition_leaves_weighted failed: %d" % rc)
    return np.array(list(first), dtype=np.int64)


def new_nodes(q: int) -> np.ndarray:
    """tbslas::new_nodes 1-D table (cheb.h:51-58)."""
    out = (C.c_double * (q + 1))()
    rc = capi.load().tbslas_b200_new_nodes(q, out)
    if rc != capi.OK:
        raise TbslasError("new_nodes(%d) failed: %d" % (q, rc))
    return np.array(out[:])

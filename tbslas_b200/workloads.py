"""Synthetic workloads of BASELINE.json's configs (harness code: numpy/torch, not on the
measured path).  Every generator is deterministic; trees are FlatTrees so the CPU
reference and the GPU path see identical bytes.

  c1  rotating Gaussian blob, uniform depth 4, q = 8         (advection.cpp:93-97, conv_adv.py)
  c2  Zalesak slotted sphere, adaptive, q = 14, max depth 7   (advection.cpp:100-105, conv_zal.py)
  c3  time-varying velocity (4 snapshots), adaptive depth 8    (advtv.cpp:171-190)
  c4  Taylor-Green velocity on a uniform 256^3 cubic grid      (fast_interp, perf_cubic.py)
  c5  uniform depth 5, q = 14, ~1.1e8 points, Morton sharded   (ns.cpp / .test_job.sh:19)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import flat_tree as ftm


def zalesak_refine(R=0.3, w=0.1, c=(0.5, 0.5, 0.5)):
    """Refine cells cut by the surface of the slotted sphere of fields.h:220-242."""
    c = np.asarray(c)

    def refine(lower, edge, d):
        upper = lower + edge[:, None]
        near = np.clip(c, lower, upper)
        dmin = np.sqrt(((near - c) ** 2).sum(1))
        far = np.where(np.abs(lower - c) > np.abs(upper - c), lower, upper)
        dmax = np.sqrt(((far - c) ** 2).sum(1))
        sphere = (dmin <= R) & (dmax >= R)
        inside = dmin < R
        ylo, yhi = lower[:, 1] - c[1], upper[:, 1] - c[1]
        xlo, xhi = lower[:, 0] - c[0], upper[:, 0] - c[0]
        py = (((ylo <= w) & (yhi >= w)) | ((ylo <= -w) & (yhi >= -w))) & (xhi >= -w)
        px = (xlo <= -w) & (xhi >= -w) & (ylo <= w) & (yhi >= -w)
        return sphere | (inside & (py | px))
    return refine


def slotted_sphere(p, R=0.3, w=0.1, c=(0.5, 0.5, 0.5)):
    """get_slotted_cylinder with a = 0 (fields.h:220-242); works on numpy or torch."""
    dx, dy, dz = p[:, 0] - c[0], p[:, 1] - c[1], p[:, 2] - c[2]
    r2 = dx * dx + dy * dy + dz * dz
    inside = r2 < R * R
    slot = (abs(dy) < w) & (dx + w > 0)
    return (inside & ~slot)


def blob_refine(c=(0.6, 0.5, 0.5), sigma=0.06):
    c = np.asarray(c)

    def refine(lower, edge, d):
        ctr = lower + 0.5 * edge[:, None]
        r = np.sqrt(((ctr - c) ** 2).sum(1))
        return r < (3.0 * sigma + edge)
    return refine


@dataclass
class Workload:
    name: str
    q: int
    bc: int
    dt: float
    con: ftm.FlatTree                      # advected scalar tree (arrival points = its leaves)
    vel: List[ftm.FlatTree]                # 1 (steady) or 4 (time varying) velocity trees
    vel_times: Optional[List[float]] = None
    desc: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def n_points(self) -> int:
        return self.con.n_leaf * (self.q + 1) ** 3


def _fit_scalar(coord, depth, q, fn, device=None, chunk=2048):
    """Chebyshev fit of a scalar function on every leaf; on the GPU with torch when a
    device is given (harness acceleration only), else numpy."""
    if device is None:
        return ftm.fit(coord, depth, q, 1, lambda p: np.asarray(fn(p), dtype=np.float64).reshape(-1, 1))
    import torch
    M = torch.from_numpy(ftm.pt2coeff(q)).to(device)
    nodes = torch.from_numpy(ftm.new_nodes_3d(q)).to(device)
    L, P = coord.shape[0], (q + 1) ** 3
    out = np.empty((L, 1, ftm.ncoef(q)))
    for a in range(0, L, chunk):
        b = min(L, a + chunk)
        c = torch.from_numpy(coord[a:b]).to(device)
        ln = torch.from_numpy(np.power(0.5, depth[a:b].astype(np.float64))).to(device)
        pts = (c[:, None, :] + ln[:, None, None] * nodes[None, :, :]).reshape(-1, 3)
        vals = fn(pts).to(torch.float64).reshape(b - a, P)
        out[a:b, 0, :] = (vals @ M).cpu().numpy()
    return ftm.FlatTree(q, 1, coord, depth, out)


def make(name: str, device=None, scale: int = 0) -> Workload:
    """scale > 0 shrinks the workload (max depth reduced by `scale`) for tests.

    The host BLAS/LAPACK behind the fits (pinv, matmul) runs single-threaded here: its rounding
    depends on the thread count, and `torchrun` sets OMP_NUM_THREADS=1 -- without the limit the
    N = 1 bench and the N = 2, 4, 8 benches would advect coefficient sets that differ in the last
    bits and their checksums could not be compared."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        return _make(name, device, scale)
    with threadpool_limits(limits=1):
        return _make(name, device, scale)


def _make(name: str, device=None, scale: int = 0) -> Workload:
    name = name.lower()
    if name == "c1":
        q, depth = 8, max(1, 4 - scale)
        coord, dd = ftm.uniform_leaves(depth)
        vel = ftm.fit(coord, dd, q, 3, ftm.vel_rotation)
        con = ftm.fit(coord, dd, q, 1, lambda p: ftm.gaussian(p, (0.6, 0.5, 0.5), 0.06))
        return Workload("c1", q, ftm.FREESPACE, 0.0628, con, [vel],
                        desc="rotating Gaussian blob, uniform depth %d, q=8" % depth)
    if name == "c2":
        q, md = 14, max(3, 7 - scale)
        coord, dd = ftm.adaptive_leaves(zalesak_refine(), 3, md)
        con = _fit_scalar(coord, dd, q, lambda p: slotted_sphere(p) * 1.0, device)
        vc, vd = ftm.uniform_leaves(3)
        vel = ftm.fit(vc, vd, q, 3, ftm.vel_rotation)
        return Workload("c2", q, ftm.FREESPACE, 0.0628, con, [vel],
                        desc="Zalesak slotted sphere, adaptive depth 3..%d, q=14" % md)
    if name == "c3":
        q, md = 14, max(3, 8 - scale)
        coord, dd = ftm.adaptive_leaves(blob_refine((0.7, 0.7, 0.7), 0.03), 3, md)

        def g(p):
            r2 = ((p - p.new_tensor([0.7, 0.7, 0.7])) ** 2).sum(1) if hasattr(p, "new_tensor") \
                else ((p - np.array([0.7, 0.7, 0.7])) ** 2).sum(1)
            return (-r2 / (2 * 0.03 ** 2)).exp() if hasattr(r2, "exp") else np.exp(-r2 / (2 * 0.03 ** 2))
        con = _fit_scalar(coord, dd, q, g, device)
        vc, vd = ftm.uniform_leaves(3)
        dt = 0.0628 / 4
        times = [-dt, 0.0, dt, 2 * dt]
        vels = [ftm.fit(vc, vd, q, 3, lambda p, t=t: ftm.vel_rotation(p) * np.cos(2 * np.pi * t))
                for t in times]
        return Workload("c3", q, ftm.PERIODIC, dt, con, vels, times,
                        desc="time-varying rotation (4 snapshots), adaptive depth 3..%d, q=14" % md)
    if name == "c5":
        q, depth = 14, max(2, 5 - scale)
        coord, dd = ftm.uniform_leaves(depth)
        con = ftm.random_tree(coord, dd, q, 1, seed=2)
        vc, vd = ftm.uniform_leaves(min(depth, 3))
        vel = ftm.fit(vc, vd, q, 3, lambda p: 0.5 * ftm.vel_taylor_green(p))
        cfl_dt = 1.0 / ((1 << depth) * q * q)  # dt * 2^depth * q^2 = 1 (common.h:246-254)
        return Workload("c5", q, ftm.PERIODIC, cfl_dt * 20, con, [vel],
                        desc="uniform depth %d, q=14, Taylor-Green velocity" % depth)
    raise ValueError("unknown workload %r" % name)


def flops_per_point_eval(q: int, dof: int) -> int:
    """The reference's own FLOP model for one tree evaluation (tree_functor.h:389-394)."""
    d = q + 1
    return 3 * d * 3 + ftm.ncoef(q) * dof * 2


def flops_per_point_step(q: int, n_vel_trees: int = 1) -> int:
    """One SolveSemilagRK2, nrk = 1: two velocity evaluations (dof 3, per tree), one scalar
    evaluation, plus the RK2 updates (traj.inc:44)."""
    return 2 * n_vel_trees * flops_per_point_eval(q, 3) + flops_per_point_eval(q, 1) + 18


def partition_leaves(n_leaf: int, nranks: int) -> np.ndarray:
    """Equal-count contiguous Morton ranges: first[r] = r*n_leaf//nranks (what PVFMM's
    RedistNodes gives a uniform-weight tree); mirrors tbslas_b200_partition_leaves."""
    return np.array([r * n_leaf // nranks for r in range(nranks + 1)], dtype=np.int64)


def owner_of_keys(keys: np.ndarray, splitters: np.ndarray) -> np.ndarray:
    """Owner rank of each key: last r with splitters[r] <= key, rank 0 below the first
    (tree_functor.h:491-513 / the split-key rule of par::SortScatterIndex)."""
    return np.maximum(np.searchsorted(splitters, keys, side="right") - 1, 0)


def shard_by_splitters(ft: ftm.FlatTree, splitters: np.ndarray, rank: int) -> ftm.FlatTree:
    """The leaves of `ft` that rank `rank` owns when the tree is re-partitioned with the
    given break points -- whole leaves, assigned by their own Morton id, as PVFMM's
    RedistNodes does after tbslas::MergeTree computed common break points for the velocity
    and the scalar tree (tree_utils.h:703-728).  May be empty."""
    own = owner_of_keys(ft.keys(), np.asarray(splitters, dtype=np.uint64))
    idx = np.nonzero(own == rank)[0]
    if idx.size == 0:
        return ft.shard(0, 0)
    assert idx[-1] - idx[0] + 1 == idx.size  # contiguous by construction
    return ft.shard(int(idx[0]), int(idx[-1]) + 1)

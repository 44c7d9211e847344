"""FlatTree: the frozen leaf list the GPU path consumes, plus synthetic builders.

The reference walks a PVFMM ``MPI_Tree`` and touches, per leaf, exactly
``Coord() Depth() ChebDeg() DataDOF() ChebData() GetMortonId()``
(reference src/tree/tree_functor.h:161,249-250,283-284,420-426).  A FlatTree is that
information for all non-ghost leaves in Morton (= PVFMM preorder) order:

    coord  f64 [L][3]         lower corner of the leaf
    depth  u8  [L]            leaf edge = 2**-depth
    coeff  f64 [L][dof][Ncoef] Chebyshev coefficients in the reference's packed
                              triangular order: for i (z), for j (y, i+j<=q),
                              for k (x, i+j+k<=q)   (tree_functor.h:256-266)

Builders here are harness code (numpy) used by tests/ and bench.py to make
deterministic synthetic inputs; they are not on the measured path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

MAX_DEPTH = 15  # pvfmm MAX_DEPTH (reference sim_config.h:40)
FREESPACE, PERIODIC = 0, 1  # pvfmm::BoundaryType


def ncoef(q: int) -> int:
    return (q + 1) * (q + 2) * (q + 3) // 6


def _spread3(v: np.ndarray) -> np.ndarray:
    x = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x001F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x001F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def anchor_key(ix, iy, iz) -> np.ndarray:
    """Interleaved Morton key of integer anchors at depth 15 (z most significant)."""
    ix, iy, iz = (np.asarray(a, dtype=np.uint64) for a in (ix, iy, iz))
    return _spread3(ix) | (_spread3(iy) << np.uint64(1)) | (_spread3(iz) << np.uint64(2))


@dataclass
class FlatTree:
    q: int
    dof: int
    coord: np.ndarray  # [L,3] f64
    depth: np.ndarray  # [L] u8
    coeff: np.ndarray  # [L,dof,Ncoef] f64

    def __post_init__(self):
        self.coord = np.ascontiguousarray(self.coord, dtype=np.float64).reshape(-1, 3)
        self.depth = np.ascontiguousarray(self.depth, dtype=np.uint8).reshape(-1)
        self.coeff = np.ascontiguousarray(self.coeff, dtype=np.float64).reshape(
            self.n_leaf, self.dof, ncoef(self.q))

    @property
    def n_leaf(self) -> int:
        return int(self.coord.shape[0])

    @property
    def ncoef(self) -> int:
        return ncoef(self.q)

    def keys(self) -> np.ndarray:
        a = np.floor(self.coord * float(1 << MAX_DEPTH)).astype(np.uint64)
        return anchor_key(a[:, 0], a[:, 1], a[:, 2])

    def shard(self, lo: int, hi: int) -> "FlatTree":
        """Contiguous Morton range [lo, hi) of leaves (one rank's share)."""
        return FlatTree(self.q, self.dof, self.coord[lo:hi], self.depth[lo:hi],
                        self.coeff[lo:hi])


# ----------------------------------------------------------------------------
# leaf sets
# ----------------------------------------------------------------------------
def uniform_leaves(depth: int):
    """All 8**depth leaves of a uniform octree in Morton order -> (coord, depth)."""
    n = 1 << depth
    idx = np.arange(n, dtype=np.uint64)
    iz, iy, ix = np.meshgrid(idx, idx, idx, indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    sh = np.uint64(MAX_DEPTH - depth)
    key = anchor_key(ix << sh, iy << sh, iz << sh)
    order = np.argsort(key, kind="stable")
    coord = np.stack([ix[order], iy[order], iz[order]], axis=1).astype(np.float64) / n
    return coord, np.full(coord.shape[0], depth, dtype=np.uint8)


def adaptive_leaves(refine: Callable[[np.ndarray, np.ndarray, int], np.ndarray],
                    min_depth: int, max_depth: int):
    """Octree refined where ``refine(lower_corner[n,3], edge[n], depth) -> bool[n]``.

    Cells shallower than min_depth are always split, none is split at max_depth.
    Returns (coord, depth) in Morton order.
    """
    cells = np.zeros((1, 3), dtype=np.uint64)  # anchors in depth-15 units
    leaves_a, leaves_d = [], []
    for d in range(0, max_depth + 1):
        if cells.shape[0] == 0:
            break
        edge = 1.0 / (1 << d)
        lower = cells.astype(np.float64) / float(1 << MAX_DEPTH)
        if d < min_depth:
            split = np.ones(cells.shape[0], dtype=bool)
        elif d >= max_depth:
            split = np.zeros(cells.shape[0], dtype=bool)
        else:
            split = np.asarray(refine(lower, np.full(cells.shape[0], edge), d), dtype=bool)
        keep = cells[~split]
        leaves_a.append(keep)
        leaves_d.append(np.full(keep.shape[0], d, dtype=np.uint8))
        par = cells[split]
        if par.shape[0]:
            h = np.uint64(1 << (MAX_DEPTH - d - 1))
            off = np.array([[(c >> 0) & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)],
                           dtype=np.uint64) * h
            cells = (par[:, None, :] + off[None, :, :]).reshape(-1, 3)
        else:
            cells = np.zeros((0, 3), dtype=np.uint64)
    a = np.concatenate(leaves_a, axis=0)
    dd = np.concatenate(leaves_d, axis=0)
    order = np.argsort(anchor_key(a[:, 0], a[:, 1], a[:, 2]), kind="stable")
    return a[order].astype(np.float64) / float(1 << MAX_DEPTH), dd[order]


# ----------------------------------------------------------------------------
# Chebyshev machinery (harness side; restates cheb.h for input generation only)
# ----------------------------------------------------------------------------
def new_nodes_1d(q: int) -> np.ndarray:
    """tbslas::new_nodes (cheb.h:51-58): stretched nodes that include 0 and 1."""
    d = q + 1
    i = np.arange(d, dtype=np.float64)
    scal = 1.0 / np.cos(0.5 * np.pi / d)
    return -np.cos((i + 0.5) * np.pi / d) * scal * 0.5 + 0.5


def new_nodes_3d(q: int) -> np.ndarray:
    """[(q+1)^3, 3] tensor grid, x fastest (cheb.h:60-67)."""
    x = new_nodes_1d(q)
    d = q + 1
    j = np.arange(d ** 3)
    return np.stack([x[j % d], x[(j // d) % d], x[(j // (d * d)) % d]], axis=1)


def tri_index(q: int) -> np.ndarray:
    """[Ncoef,3] (i,j,k) = (z,y,x) degrees in packed coefficient order."""
    d = q + 1
    return np.array([(i, j, k) for i in range(d) for j in range(d - i)
                     for k in range(d - i - j)], dtype=np.int64)


def cheb_T(q: int, xi: np.ndarray) -> np.ndarray:
    """T_0..T_q at xi in [-1,1] -> [len(xi), q+1]."""
    xi = np.asarray(xi, dtype=np.float64)
    T = np.empty((xi.shape[0], q + 1))
    T[:, 0] = 1.0
    if q >= 1:
        T[:, 1] = xi
    for i in range(2, q + 1):
        T[:, i] = 2 * xi * T[:, i - 1] - T[:, i - 2]
    return T


_PT2COEFF = {}


def pt2coeff(q: int) -> np.ndarray:
    """[(q+1)^3, Ncoef] values-at-new_nodes -> coefficients (cheb.h:166-196: pinv of
    the basis matrix).  Harness only."""
    if q not in _PT2COEFF:
        pts = new_nodes_3d(q) * 2.0 - 1.0
        Tx, Ty, Tz = cheb_T(q, pts[:, 0]), cheb_T(q, pts[:, 1]), cheb_T(q, pts[:, 2])
        ijk = tri_index(q)
        B = Tz[:, ijk[:, 0]] * Ty[:, ijk[:, 1]] * Tx[:, ijk[:, 2]]  # [P, Ncoef]
        _PT2COEFF[q] = np.linalg.pinv(B).T.copy()  # [P, Ncoef]
    return _PT2COEFF[q]


def grid_points(coord: np.ndarray, depth: np.ndarray, q: int) -> np.ndarray:
    """CollectChebTreeGridPoints (tree_utils.h:442-498): [L*(q+1)^3, 3] leaf-major."""
    nodes = new_nodes_3d(q)
    length = np.power(0.5, depth.astype(np.float64))
    pts = coord[:, None, :] + length[:, None, None] * nodes[None, :, :]
    return pts.reshape(-1, 3)


def fit(coord, depth, q: int, dof: int, fn: Callable[[np.ndarray], np.ndarray],
        chunk: int = 512) -> FlatTree:
    """Least-squares Chebyshev fit of fn(points[n,3]) -> [n,dof] on every leaf."""
    M = pt2coeff(q)
    L = coord.shape[0]
    P = (q + 1) ** 3
    coeff = np.empty((L, dof, ncoef(q)))
    for a in range(0, L, chunk):
        b = min(L, a + chunk)
        pts = grid_points(coord[a:b], depth[a:b], q)
        vals = np.asarray(fn(pts), dtype=np.float64).reshape(b - a, P, dof)
        coeff[a:b] = np.einsum("lpd,pn->ldn", vals, M, optimize=True)
    return FlatTree(q, dof, coord, depth, coeff)


def random_tree(coord, depth, q: int, dof: int, seed: int, decay: float = 0.5,
                scale: float = 1.0) -> FlatTree:
    """Random coefficients with geometric decay in total degree (smooth-field-like
    spectrum so that sums do not cancel catastrophically)."""
    rng = np.random.default_rng(seed)
    ijk = tri_index(q)
    w = scale * decay ** ijk.sum(axis=1)
    coeff = rng.uniform(-1.0, 1.0, size=(coord.shape[0], dof, ncoef(q))) * w
    return FlatTree(q, dof, coord, depth, coeff)


# ----------------------------------------------------------------------------
# analytic fields used by the reference's drivers (inputs for synthetic configs)
# ----------------------------------------------------------------------------
def vel_rotation(p: np.ndarray, omega: float = 1.0) -> np.ndarray:
    """Solid-body rotation about the z axis through (0.5,0.5) (fields.h:142-151)."""
    return np.stack([omega * (0.5 - p[:, 1]), omega * (p[:, 0] - 0.5),
                     np.zeros(p.shape[0])], axis=1)


def gaussian(p: np.ndarray, c=(0.6, 0.5, 0.5), sigma: float = 0.06) -> np.ndarray:
    """Gaussian blob (field_wrappers.h:31-56 shape)."""
    r2 = ((p - np.asarray(c)) ** 2).sum(axis=1)
    return np.exp(-r2 / (2 * sigma * sigma))[:, None]


def slotted_cylinder(p: np.ndarray, c=(0.5, 0.75, 0.5), R=0.15, w=0.05, h=0.25):
    """Zalesak slotted disk extruded in z (fields.h:220-242 shape)."""
    dx, dy = p[:, 0] - c[0], p[:, 1] - c[1]
    inside = (dx * dx + dy * dy <= R * R) & ~((np.abs(dx) < w / 2) & (dy < h - R))
    return inside.astype(np.float64)[:, None]


def vel_taylor_green(p: np.ndarray) -> np.ndarray:
    """Taylor-Green vortex (fields.h:51-67)."""
    a = 2 * np.pi * p
    return np.stack([np.cos(a[:, 0]) * np.sin(a[:, 1]) * np.sin(a[:, 2]),
                     np.sin(a[:, 0]) * np.cos(a[:, 1]) * np.sin(a[:, 2]),
                     np.sin(a[:, 0]) * np.sin(a[:, 1]) * np.cos(a[:, 2])], axis=1)

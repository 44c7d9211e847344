"""CPU suite, part 2: the C-ABI library loads on a host without a GPU, exports every
symbol include/tbslas_b200.h declares, refuses to compute without a device (no CPU
fallback), and its host-side shard logic agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tbslas_b200 import capi
from tbslas_b200 import flat_tree as ftm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    hdr = open(os.path.join(ROOT, "include", "tbslas_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tbslas_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(capi.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_device():
    if _has_gpu():
        pytest.skip("a GPU is present")
    lib = capi.load()
    h = C.c_void_p()
    assert lib.tbslas_b200_init(0, C.byref(h)) == capi.ERR_CUDA
    assert not h.value
    from tbslas_b200.api import Context
    with pytest.raises(capi.TbslasError):
        Context(0)


def test_point_key_matches_oracle(port):
    lib = capi.load()
    rng = np.random.default_rng(2)
    pts = rng.uniform(-0.1, 1.1, size=(5000, 3))
    pts[:100] = rng.integers(0, 9, size=(100, 3)) / 8.0
    for bc in (0, 1):
        for p in pts:
            assert lib.tbslas_b200_point_key(p[0], p[1], p[2], bc) == \
                port.lib.orc_point_key(p[0], p[1], p[2], bc)


def test_partition_and_owner():
    lib = capi.load()
    first = (C.c_size_t * 5)()
    assert lib.tbslas_b200_partition_leaves(10, 4, first) == 0
    assert list(first) == [0, 2, 5, 7, 10]
    spl = (C.c_uint64 * 3)(0, 100, 200)
    assert lib.tbslas_b200_owner_of_key(0, spl, 3) == 0
    assert lib.tbslas_b200_owner_of_key(99, spl, 3) == 0
    assert lib.tbslas_b200_owner_of_key(100, spl, 3) == 1
    assert lib.tbslas_b200_owner_of_key(2 ** 64 - 1, spl, 3) == 2
    spl = (C.c_uint64 * 2)(50, 100)  # key below the first splitter still goes to rank 0
    assert lib.tbslas_b200_owner_of_key(3, spl, 2) == 0


def test_weighted_partition():
    """Equal weights reproduce equal-count ranges; skewed weights balance the weight, keep the
    ranges contiguous and monotone; bad weights are rejected."""
    from tbslas_b200.api import partition_leaves_weighted
    assert partition_leaves_weighted(np.ones(10), 4).tolist() in ([0, 2, 5, 7, 10], [0, 3, 5, 8, 10])
    rng = np.random.default_rng(3)
    w = rng.uniform(0, 1, size=5000) ** 4
    w[1000:1100] *= 50
    for nr in (2, 3, 8):
        first = partition_leaves_weighted(w, nr)
        assert first[0] == 0 and first[-1] == w.size and np.all(np.diff(first) >= 0)
        loads = np.array([w[first[r]:first[r + 1]].sum() for r in range(nr)])
        assert loads.max() <= w.sum() / nr + w.max()
    assert partition_leaves_weighted(np.zeros(7), 3).tolist() == [0, 2, 4, 7]   # no weight: by count
    assert partition_leaves_weighted(np.array([5.0]), 4).tolist()[-1] == 1
    with pytest.raises(Exception):
        partition_leaves_weighted(np.array([1.0, -1.0]), 2)


def test_cubic_time_weights_reproduce_interp_cubic1d(port):
    """The weights behind the one-evaluation route of FieldSetFunctor: sum_k w_k p_k must equal the
    reference's InterpCubic1D (cubic.h:36-56, restated in the oracle) for any snapshot values,
    uniform and non-uniform snapshot times, query times inside and outside [t1, t2]."""
    lib = capi.load()
    rng = np.random.default_rng(11)
    for times in ([-0.05, 0.0, 0.05, 0.1], [0.0, 0.3, 0.35, 1.0], [1.0, 2.0, 4.0, 4.5]):
        tt = (C.c_double * 4)(*times)
        for tq in (times[1], times[2], 0.5 * (times[1] + times[2]), times[1] + 0.123 * (times[2] - times[1]),
                   times[2] + 0.2 * (times[2] - times[1])):
            w = (C.c_double * 4)()
            assert lib.tbslas_b200_cubic_time_weights(tt, tq, w) == 0
            w = np.array(w[:])
            assert abs(w.sum() - 1.0) < 1e-13          # constants are reproduced
            for _ in range(20):
                p = rng.standard_normal(4)
                want = port.interp_cubic1d(tq, np.array(times), p)
                assert abs(w @ p - want) <= 1e-13 * (1 + np.abs(p).max() * np.abs(w).max())


def test_new_nodes_matches_oracle(port):
    from tbslas_b200.api import new_nodes
    for q in range(1, 17):
        assert np.array_equal(new_nodes(q), port.new_nodes(q, 1).ravel())
    assert np.array_equal(ftm.new_nodes_1d(8), port.new_nodes(8, 1).ravel()) or \
        np.allclose(ftm.new_nodes_1d(8), port.new_nodes(8, 1).ravel(), atol=1e-16)

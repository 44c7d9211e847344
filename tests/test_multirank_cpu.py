"""CPU suite, part 3: the N > 1 host logic under torch.distributed (gloo, world_size 2 and 3).

The GPU library cannot run here, so each rank stands in for its GPU with the oracle and the
test walks the exchange protocol of tbslas_b200/csrc/comm.cu by hand, using the library's
own host-side shard functions (tbslas_b200_partition_leaves / _point_key / _owner_of_key):

  shard leaves by Morton range -> first-leaf keys all-gathered (splitters) -> every rank
  classifies ITS points by owner -> count matrix all-gathered -> outsiders travel to their
  owner -> owner evaluates them on its shard -> values and global leaf ids travel back ->
  un-permute.

The result must equal the single-rank oracle on the full tree BIT FOR BIT (the evaluation of
a point depends only on the leaf that contains it) -- the same invariant tests/multigpu_check.py
asserts for the CUDA path on real GPUs.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from oracle import Oracle
        from tbslas_b200 import capi, workloads
        from tbslas_b200 import flat_tree as ftm

        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                                world_size=world)
        lib = capi.load()
        orc = Oracle("port")

        def refine(lower, edge, d):
            c = lower + 0.5 * edge[:, None]
            return np.abs(np.sqrt(((c - 0.5) ** 2).sum(axis=1)) - 0.3) < edge
        coord, dd = ftm.adaptive_leaves(refine, 1, 4)
        full = ftm.random_tree(coord, dd, 5, 2, seed=11)
        vc, vd = ftm.uniform_leaves(1)
        coarse = ftm.random_tree(vc, vd, 5, 3, seed=12)  # 8 leaves: some ranks own few or none

        # partition through the C ABI's host function
        first = (C.c_size_t * (world + 1))()
        assert lib.tbslas_b200_partition_leaves(full.n_leaf, world, first) == 0
        first = list(first)
        assert first == workloads.partition_leaves(full.n_leaf, world).tolist()
        break_keys = full.keys()[np.array(first[:-1])]

        for tree, mine in ((full, full.shard(first[rank], first[rank + 1])),
                           (coarse, workloads.shard_by_splitters(coarse, break_keys, rank))):
            # splitters = first-leaf key of every rank (tree_functor.h:425,433-437); an empty
            # rank inherits the next owner's key (comm.cu: comm_tree_splitters)
            info = [None] * world
            dist.all_gather_object(info, (int(mine.keys()[0]) if mine.n_leaf else None, mine.n_leaf))
            spl = [k for k, _ in info]
            for r in range(world - 1, -1, -1):
                if spl[r] is None:
                    spl[r] = spl[r + 1] if r + 1 < world else 2 ** 64 - 1
            assert sum(n for _, n in info) == tree.n_leaf
            offset = sum(n for _, n in info[:rank])
            spl_c = (C.c_uint64 * world)(*spl)

            rng = np.random.default_rng(50 + rank)
            pts = rng.uniform(-0.05, 1.05, size=(3000 + 100 * rank, 3))
            for bc in (0, 1):
                pos = pts.copy()
                if bc == 1:  # tree_functor.h:442-449
                    pos = np.where(pos < 0, pos + 1.0, pos)
                    pos = np.where(pos >= 1.0, pos - 1.0, pos)
                owner = np.array([lib.tbslas_b200_owner_of_key(
                    lib.tbslas_b200_point_key(p[0], p[1], p[2], bc), spl_c, world) for p in pos])
                assert np.array_equal(owner, workloads.owner_of_keys(
                    np.array([lib.tbslas_b200_point_key(p[0], p[1], p[2], bc) for p in pos],
                             dtype=np.uint64), np.array(spl, dtype=np.uint64)))
                send = [pos[owner == r] for r in range(world)]
                idx = [np.nonzero(owner == r)[0] for r in range(world)]
                counts = [None] * world
                dist.all_gather_object(counts, [len(s) for s in send])  # the count matrix
                # forward all-to-all-v (an object gather stands in for NCCL): everyone publishes
                # its buckets, each rank picks its column
                allb = [None] * world
                dist.all_gather_object(allb, send)
                inbox = [allb[src][rank] for src in range(world)]
                for src in range(world):
                    assert len(inbox[src]) == counts[src][rank]
                vals, leaves = [], []
                h = orc.tree_create(mine) if mine.n_leaf else None
                for src in range(world):
                    if h is None or len(inbox[src]) == 0:
                        vals.append(np.zeros((len(inbox[src]), tree.dof)))
                        leaves.append(np.full(len(inbox[src]), -1, dtype=np.int32))
                        continue
                    v, lf, _ = orc.eval_tree(h, tree.dof, inbox[src], bc)
                    vals.append(v)
                    leaves.append(np.where(lf >= 0, lf + offset, -1).astype(np.int32))
                allv = [None] * world
                dist.all_gather_object(allv, (vals, leaves))  # reverse all-to-all-v
                out = np.empty((pts.shape[0], tree.dof))
                leaf = np.empty(pts.shape[0], dtype=np.int32)
                for r in range(world):
                    out[idx[r]] = allv[r][0][rank]
                    leaf[idx[r]] = allv[r][1][rank]
                hf = orc.tree_create(tree)
                want_v, want_l, want_p = orc.eval_tree(hf, tree.dof, pts, bc)
                assert np.array_equal(want_p, pos)
                assert np.array_equal(leaf, want_l)
                assert np.array_equal(out, want_v)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_protocol_partition_invariant(world):
    import torch.multiprocessing as mp
    from oracle import build
    build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_shard_helpers_cover_the_tree_exactly_once():
    from tbslas_b200 import workloads
    from tbslas_b200 import flat_tree as ftm
    coord, dd = ftm.uniform_leaves(2)
    t = ftm.random_tree(coord, dd, 3, 1, seed=1)
    for world in (1, 2, 3, 5, 8):
        first = workloads.partition_leaves(t.n_leaf, world)
        assert first[0] == 0 and first[-1] == t.n_leaf and np.all(np.diff(first) >= 0)
        spl = t.keys()[first[:-1]]
        cv, cd = ftm.uniform_leaves(1)
        coarse = ftm.random_tree(cv, cd, 3, 3, seed=2)
        shards = [workloads.shard_by_splitters(coarse, spl, r) for r in range(world)]
        assert sum(s.n_leaf for s in shards) == coarse.n_leaf
        got = np.concatenate([s.keys() for s in shards])
        assert np.array_equal(got, coarse.keys())

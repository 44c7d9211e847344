"""Multi-GPU parity check, one rank per GPU (not collected by pytest; launched by
tests/test_multigpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/multigpu_check.py

Every rank holds a contiguous Morton range of each tree (whole leaves, as an MPI rank of the
reference does), evaluates ITS OWN points through the sharded trees -- outsiders travel inside
the library, once through the NVLink peer-memory mailboxes and once through the NCCL
all-to-all-v -- and compares with the same points evaluated on the FULL tree by a single-rank
context on the same GPU.  The evaluation of a point depends only on the leaf that contains it,
so the results must be BIT-IDENTICAL (values, leaf ids, departure points), whatever the partition
and whatever the exchange.  The oracle (the reference's algorithm on the CPU, full trees) pins
the same results from outside the product: leaf ids bit-exact, values 1e-12.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from tbslas_b200 import api, workloads
    from tbslas_b200 import flat_tree as ftm

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    ctx.comm_init_torch()
    assert ctx.comm_rank() == (rank, world)
    solo = api.Context(local)  # single-rank context on the same GPU: the full trees

    def refine(lower, edge, d):
        c = lower + 0.5 * edge[:, None]
        r = np.sqrt(((c - np.array([0.55, 0.5, 0.45])) ** 2).sum(axis=1))
        return np.abs(r - 0.3) < edge
    coord, dd = ftm.adaptive_leaves(refine, 2, 5)
    q = 6
    con = ftm.fit(coord, dd, q, 1, lambda p: ftm.gaussian(p, (0.5, 0.5, 0.5), 0.2))
    vc, vd = ftm.uniform_leaves(2)
    vels = [ftm.fit(vc, vd, q, 3, lambda p, s=s: s * (ftm.vel_rotation(p) + 0.1 * ftm.vel_taylor_green(p)))
            for s in (0.8, 1.0, 1.1, 1.3)]
    times = [-0.05, 0.0, 0.05, 0.1]

    first = workloads.partition_leaves(con.n_leaf, world)
    splitters = con.keys()[first[:-1]]
    con_l = con.shard(int(first[rank]), int(first[rank + 1]))
    vel_l = [workloads.shard_by_splitters(v, splitters, rank) for v in vels]
    tcon, tvel = ctx.tree(con_l), [ctx.tree(v) for v in vel_l]
    scon, svel = solo.tree(con), [solo.tree(v) for v in vels]
    from oracle import Oracle
    port = Oracle("port")
    hcon = port.tree_create(con)
    hvel1 = port.tree_create(vels[1])
    mode0, mailbox = ctx.comm_exchange_mode()
    modes = ["peer", "nccl"] if (mode0 == "peer" and world > 1) else ["nccl"]
    if os.environ.get("TBSLAS_CHECK_MODES"):
        modes = os.environ["TBSLAS_CHECK_MODES"].split(",")
    all_checks = []
    for mode in modes:
        ctx.comm_set_exchange(mode)
        tag = "[%s] " % mode
        checks = []

        def same(name, a, b):
            ok = np.array_equal(a, b)
            checks.append((name, ok))
            if not ok:
                d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
                print("[rank %d] MISMATCH %s: max abs diff %g at %d of %d entries" %
                      (rank, name, d.max(), int((d != 0).sum()), d.size), flush=True)

        def close(name, a, b, tol=1e-12):
            sc = max(np.abs(b).max(), 1e-300) if np.size(b) else 1.0
            e = np.abs(np.asarray(a) - np.asarray(b)).max() / sc if np.size(b) else 0.0
            checks.append((name, bool(e < tol)))
            if not e < tol:
                print("[rank %d] MISMATCH %s: rel err %g" % (rank, name, e), flush=True)

        # (1) arbitrary points (different on every rank, most of them outsiders), both bcs
        rng = np.random.default_rng(100 + rank)
        n = 40000 + 1000 * rank
        pts = np.concatenate([rng.uniform(-0.05, 1.05, size=(n, 3)), rng.integers(0, 33, size=(500, 3)) / 32.0])
        for bc in (0, 1):
            f, g = api.NodeFieldFunctor(tcon), api.NodeFieldFunctor(scon)
            pa, pb = pts.copy(), pts.copy()
            va, la = f.eval_with_leaf(pa, bc)
            sent, recv = ctx.comm_last_exchange()
            vb, lb = g.eval_with_leaf(pb, bc)
            same("eval values bc%d" % bc, va, vb)
            same("eval leaf ids bc%d" % bc, la, lb)
            same("eval wrapped positions bc%d" % bc, pa, pb)
            checks.append(("outsiders exist bc%d" % bc, world == 1 or sent > 0))
            f3, g3 = api.NodeFieldFunctor(tvel[1]), api.NodeFieldFunctor(svel[1])
            v3 = f3(pts.copy(), bc=bc)
            same("eval dof3 bc%d" % bc, v3, g3(pts.copy(), bc=bc))
            # ... and against the oracle (CPU, full trees): leaf ids bit-exact, values 1e-12
            vo, lo, po = port.eval_tree(hcon, 1, pts, bc)
            same("ORACLE leaf ids bc%d" % bc, la, lo)
            same("ORACLE wrapped positions bc%d" % bc, pa, po)
            close("ORACLE eval values bc%d" % bc, va, vo)
            close("ORACLE eval dof3 bc%d" % bc, v3, port.eval_tree(hvel1, 3, pts, bc, want_leaf=False)[0])

        # (2) the semi-Lagrangian step on this rank's own arrival points, host and device buffers
        arr = ftm.grid_points(con_l.coord, con_l.depth, q)
        for bc in (0, 1):
            for nrk in (1, 2):
                vel, con_f = api.NodeFieldFunctor(tvel[1]), api.NodeFieldFunctor(tcon)
                svel_f, scon_f = api.NodeFieldFunctor(svel[1]), api.NodeFieldFunctor(scon)
                da, db = np.empty_like(arr), np.empty_like(arr)
                a = api.SolveSemilagRK2(vel, con_f, arr, 3, 0.05, nrk, bc, departure_points=da)
                b = api.SolveSemilagRK2(svel_f, scon_f, arr, 3, 0.05, nrk, bc, departure_points=db)
                same("semilag values bc%d nrk%d" % (bc, nrk), a, b)
                same("semilag departure points bc%d nrk%d" % (bc, nrk), da, db)
        d_arr = torch.from_numpy(arr).cuda()
        ctx.set_stream(torch.cuda.current_stream())
        d_val = api.SolveSemilagRK2(api.NodeFieldFunctor(tvel[1]), api.NodeFieldFunctor(tcon), d_arr, 3, 0.05, 1, 1)
        torch.cuda.synchronize()
        ctx.set_stream(None)
        same("semilag device buffers", d_val.cpu().numpy(),
             api.SolveSemilagRK2(api.NodeFieldFunctor(svel[1]), api.NodeFieldFunctor(scon), arr, 3, 0.05, 1, 1))

        # (3) time-varying and extrapolated velocity functors
        fset, gset = api.FieldSetFunctor(tvel, times), api.FieldSetFunctor(svel, times)
        same("set4 eval", fset(pts.copy(), time=0.02, bc=1), gset(pts.copy(), time=0.02, bc=1))
        same("set4 trajectory", api.ComputeTrajRK2(fset, arr, 0.05, 0.0, 2, 1),
             api.ComputeTrajRK2(gset, arr, 0.05, 0.0, 2, 1))
        fext, gext = api.FieldExtrapFunctor(tvel[0], tvel[1]), api.FieldExtrapFunctor(svel[0], svel[1])
        same("extrap trajectory",
             api.ComputeTrajRK2(api.NodeFieldFunctor(tvel[1]), arr, 0.05, 0.0, 1, 0, extrap_fn=fext),
             api.ComputeTrajRK2(api.NodeFieldFunctor(svel[1]), arr, 0.05, 0.0, 1, 0, extrap_fn=gext))

        # (4) a tree with fewer leaves than ranks: the last rank(s) own nothing
        one_c, one_d = ftm.uniform_leaves(0)
        one = ftm.random_tree(one_c, one_d, 5, 2, seed=5)
        t1 = ctx.tree(one if rank == 0 else one.shard(0, 0))
        s1 = solo.tree(one)
        same("single-leaf tree", api.NodeFieldFunctor(t1)(pts.copy(), bc=0), api.NodeFieldFunctor(s1)(pts.copy(), bc=0))
        # (4b) a velocity tree replicated on every rank: local evaluations, the step still exchanges
        # points for the (sharded) advected tree
        rvel = api.NodeFieldFunctor(ctx.tree(vels[1], replicated=True))
        same("replicated velocity: eval", rvel(pts.copy(), bc=1), api.NodeFieldFunctor(svel[1])(pts.copy(), bc=1))
        same("replicated velocity: semilag",
             api.SolveSemilagRK2(rvel, api.NodeFieldFunctor(tcon), arr, 3, 0.05, 2, 1),
             api.SolveSemilagRK2(api.NodeFieldFunctor(svel[1]), api.NodeFieldFunctor(scon), arr, 3, 0.05, 2, 1))
        # (4c) the tree-level step: arrival points generated on the device from this rank's leaves, the
        # first velocity evaluation by sum factorisation over those grids (replicated velocity tree),
        # the scalar through the exchange -- against the single-rank context's tree-level step on the
        # whole tree, restricted to this rank's leaves
        P = (q + 1) ** 3
        lo, hi = int(first[rank]) * P, int(first[rank + 1]) * P
        rtree = ctx.tree(vels[1], replicated=True)
        ctx.set_tensor_grid("always")
        solo.set_tensor_grid("always")
        rep = {}
        for bc in (0, 1):
            a = api.SolveSemilagInSitu(api.NodeFieldFunctor(rtree), tcon, 2, 0.05, 1, bc)
            b = api.SolveSemilagInSitu(api.NodeFieldFunctor(svel[1]), scon, 2, 0.05, 1, bc)
            same("tree-level step (tensor grids) bc%d" % bc, a, b[lo:hi])
            rep[bc] = a
        # (4d) the tree-level step with a Morton-SHARDED velocity tree (the reference's layout): the first
        # velocity evaluation runs by sum factorisation where the containing velocity leaf is local and
        # through the (collective) generic pass elsewhere; against the single-rank tree-level step
        # (another summation order on part of the points: 1e-12) and against the oracle's step
        for bc in (0, 1):
            a, da = api.SolveSemilagInSitu(api.NodeFieldFunctor(tvel[1]), tcon, 2, 0.05, 1, bc, departure_points=True)
            b = api.SolveSemilagInSitu(api.NodeFieldFunctor(svel[1]), scon, 2, 0.05, 1, bc)
            close("sharded-velocity tree-level step vs one GPU bc%d" % bc, a, b[lo:hi])
            # with the ranks' last velocity leaves as ghosts the sum-factorised path covers the very leaves it
            # covers when the velocity tree is held whole: the two layouts agree bit for bit
            same("sharded == replicated velocity, tree-level step bc%d" % bc, a, rep[bc])
            if arr.shape[0]:
                close("ORACLE departure points, sharded tree-level step bc%d" % bc, da,
                      port.traj_rk2(hvel1, arr, 2 * 0.05, 0.05, 1, bc), tol=1e-12 / max(np.abs(da).max(), 1.0))
                close("ORACLE values at the departure points bc%d" % bc, a, port.eval_tree(hcon, 1, da, bc, want_leaf=False)[0])
        ctx.set_tensor_grid(True)
        solo.set_tensor_grid(True)
        # (5) empty point set on one rank (the call is still collective)
        e = pts[:0].copy() if rank == world - 1 else pts[:777].copy()
        same("ragged: empty input on the last rank", api.NodeFieldFunctor(tcon)(e.copy(), bc=0),
             api.NodeFieldFunctor(scon)(e.copy(), bc=0))

        all_checks += [(tag + n_, ok_) for n_, ok_ in checks]
    checks = all_checks

    # (6) co-partitioning (row f4; tbslas::SemiMergeTree / RedistNodes, tree_utils.h:609-729): re-balance
    # the advected tree by the points each leaf received in its last evaluation, MOVE the leaves
    # between the GPUs, re-shard the velocity trees with the new split keys -- and every result is
    # still bit-identical to the single-GPU path.
    ctx.comm_set_exchange(modes[0])
    f = api.NodeFieldFunctor(tcon)
    skew = np.concatenate([np.random.default_rng(7 + rank).uniform(0, 1, size=(30000, 3)) ** 3, pts[:5000]])
    f(skew.copy(), bc=0)
    mine = tcon.last_point_counts().astype(np.float64)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    weight = np.concatenate(gathered) + 1.0
    assert weight.shape[0] == con.n_leaf
    new_first = api.partition_leaves_weighted(weight, world)
    moved = int(np.abs(new_first - first).sum())
    tcon.reshard(new_first)
    lo_l, hi_l = int(new_first[rank]), int(new_first[rank + 1])
    checks.append(("reshard: leaves moved", world == 1 or moved > 0))
    checks.append(("reshard: new local range", tcon.n_leaf == hi_l - lo_l and tcon.global_range() == (lo_l, con.n_leaf)))
    new_split = con.keys()[np.minimum(new_first[:-1], con.n_leaf - 1)]
    for i, v in enumerate(vels):
        own = workloads.owner_of_keys(v.keys(), np.asarray(new_split, dtype=np.uint64))
        vf = np.searchsorted(own, np.arange(world + 1), side="left")
        tvel[i].reshard(vf)

    def same2(name, a, b):
        ok = np.array_equal(a, b)
        checks.append((name, ok))
        if not ok:
            print("[rank %d] MISMATCH %s" % (rank, name), flush=True)
    for bc in (0, 1):
        va, la = api.NodeFieldFunctor(tcon).eval_with_leaf(pts.copy(), bc)
        vb, lb = api.NodeFieldFunctor(scon).eval_with_leaf(pts.copy(), bc)
        same2("after reshard: eval values bc%d" % bc, va, vb)
        same2("after reshard: leaf ids bc%d" % bc, la, lb)
        same2("after reshard: dof3 bc%d" % bc, api.NodeFieldFunctor(tvel[1])(pts.copy(), bc=bc),
              api.NodeFieldFunctor(svel[1])(pts.copy(), bc=bc))
    con_n = con.shard(lo_l, hi_l)
    same2("after reshard: coefficients moved intact", tcon.coefficients(), con_n.coeff)
    arr_n = ftm.grid_points(con_n.coord, con_n.depth, q)
    same2("after reshard: semilag on the new shard's arrival points",
          api.SolveSemilagRK2(api.FieldSetFunctor(tvel, times), api.NodeFieldFunctor(tcon), arr_n, 3, 0.05, 1, 1),
          api.SolveSemilagRK2(api.FieldSetFunctor(svel, times), api.NodeFieldFunctor(scon), arr_n, 3, 0.05, 1, 1))
    ctx.set_tensor_grid("always")
    solo.set_tensor_grid("always")
    Pn = (q + 1) ** 3
    a = api.SolveSemilagInSitu(api.NodeFieldFunctor(rtree), tcon, 2, 0.05, 1, 0)
    b = api.SolveSemilagInSitu(api.NodeFieldFunctor(svel[1]), scon, 2, 0.05, 1, 0)
    same2("after reshard: tree-level step", a, b[lo_l * Pn:hi_l * Pn])
    # (7) a mailbox too small for the outsiders: every rank gets TBSLAS_ERR_COMM at the synchronising call
    # (nothing is written past a peer's buffer, nothing hangs), and the next evaluation works again
    if "peer" in modes and world > 1:
        from tbslas_b200.capi import TbslasError
        ctx.comm_set_exchange("peer")
        ctx.comm_set_mailbox(64)
        raised = False
        try:
            api.NodeFieldFunctor(tcon)(pts.copy(), bc=0)
        except TbslasError as e:
            raised = "mailbox" in str(e)
        checks.append(("mailbox overflow is an error, not a hang", raised))
        ctx.comm_set_mailbox(1 << 20)
        same2("after the overflow: eval works again", api.NodeFieldFunctor(tcon)(pts.copy(), bc=0),
              api.NodeFieldFunctor(scon)(pts.copy(), bc=0))

    ok = all(c[1] for c in checks)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multigpu_check: %d ranks, %d checks per rank, %s" %
              (world, len(checks), "ALL BIT-IDENTICAL" if flag.item() else "FAILED"), flush=True)
    dist.barrier()
    ctx.close()
    solo.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()

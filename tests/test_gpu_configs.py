"""GPU suite: BASELINE.json's configurations at FULL size, checked through size-independent
properties (the oracle cannot run 10^8 points in seconds; tests/test_gpu_parity.py holds the
point-by-point parity at oracle-friendly sizes).

  c1  covered by test_gpu_parity.py::test_gpu_properties_at_config1_size
  c2  Zalesak, adaptive depth 3..7, q = 14 (274.5 M points): leaf ids by construction,
      oracle spot check on a strided sample of the real arrival points
  c3  time-varying velocity (4 snapshots), adaptive depth 8, periodic: the 4-tree functor
      equals the cubic-in-time combination of four single-tree evaluations BIT FOR BIT;
      snapshots all equal -> the time interpolation is the identity to rounding
  c4  uniform cubic grid 256^3 x dof 3, 16.8 M queries: oracle spot check (bit-exact),
      cubic polynomial reproduced, out-of-domain queries are exactly zero
  c5  uniform depth 5, q = 14, 110.6 M points, periodic: every arrival point locates to its
      own leaf; zero velocity -> the step is the identity on the tree's own grid values
"""
import numpy as np
import pytest

from tbslas_b200 import flat_tree as ftm

pytestmark = pytest.mark.gpu


def _mods():
    import torch
    from tbslas_b200 import api, workloads
    return torch, api, workloads


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def check_tree_level_step(api, port, ctx, wl, tvel, tcon, stride, label):
    """The step the bench TIMES -- tbslas_b200_semilag_insitu on device buffers with the library's
    defaults (first velocity evaluation by sum factorisation over the arrival grids, exceptions
    and the other two evaluations point by point) -- against the oracle, stage by stage, on a
    strided sample of the arrival points:
      (a) departure points vs the oracle's ComputeTrajRK2:           1e-12 absolute (domain is [0,1]^3)
      (b) values vs the oracle's evaluation AT the GPU's departure points: 1e-12 of the field scale
      (c) leaf of every sampled departure point: bit-exact
    and the composed step vs the oracle's SolveSemilagRK2, reported; its bar is 1e-12 plus the
    measured position difference times the field's steepness between the two sets of departure points
    (a difference of one ulp in a departure point moves a steep field by more than 1e-12)."""
    import torch
    vel = api.NodeFieldFunctor(tvel[0]) if len(tvel) == 1 else api.FieldSetFunctor(tvel, wl.vel_times)
    con = api.NodeFieldFunctor(tcon)
    vals, dep = api.SolveSemilagInSitu(vel, tcon, 1, wl.dt, 1, wl.bc, device=True, departure_points=True)
    torch.cuda.synchronize()
    n = vals.shape[0]
    if n >= (4 << 20):
        assert ctx.last_grid_exceptions() > 0  # the sum-factorised path ran (and left its exceptions)
    idx = torch.arange(0, n, stride, device=vals.device)
    assert idx.numel() >= 25000 or n < 25000 * stride
    pos = tcon.collect_grid_points(device=True)
    arrival = pos[idx].cpu().numpy()
    del pos
    hv = [port.tree_create(v) for v in wl.vel]
    hc = port.tree_create(wl.con)
    kind = "steady" if len(hv) == 1 else "set4"
    vh = hv[0] if len(hv) == 1 else hv
    tinit = 1 * wl.dt
    dep_o = port.traj_rk2(vh, arrival, tinit, tinit - wl.dt, 1, wl.bc, kind=kind, times=wl.vel_times)
    dep_g = dep[idx].cpu().numpy()
    e_dep = np.abs(dep_g - dep_o).max()
    want_at_g, leaf_o, _ = port.eval_tree(hc, 1, dep_g, wl.bc)
    got = vals[idx].cpu().numpy()
    scale = max(np.abs(want_at_g).max(), 1e-300)
    e_val = np.abs(got - want_at_g).max() / scale
    vg, leaf_g = con.eval_with_leaf(dep_g.copy(), wl.bc)
    want_step = port.semilag_rk2(vh, hc, 1, arrival, 1, wl.dt, 1, wl.bc, kind=kind, times=wl.vel_times)
    e_step = np.abs(got - want_step).max() / scale
    # steepness of the field between the two sets of departure points (oracle at both)
    want_at_o, _, _ = port.eval_tree(hc, 1, dep_o, wl.bc)
    moved = np.abs(want_at_g - want_at_o).max() / scale
    print("[%s] n=%d sample=%d  departure points %.2e  values at equal points %.2e  composed step %.2e "
          "(field moves %.2e between the two departure sets)" % (label, n, idx.numel(), e_dep, e_val, e_step, moved))
    assert e_dep < 1e-12
    assert e_val < 1e-12
    assert np.array_equal(leaf_g, leaf_o)
    assert e_step < 1e-12 + 1.01 * moved
    for h in hv + [hc]:
        port.tree_destroy(h)
    return vals


def test_config1_gaussian_full_size_tree_level_step(ctx, port):
    torch, api, workloads = _mods()
    wl = workloads.make("c1", torch.device("cuda", 0))
    tcon, tvel = ctx.tree(wl.con), [ctx.tree(v) for v in wl.vel]
    ctx.set_stream(torch.cuda.current_stream())
    try:
        ctx.set_tensor_grid("always")  # 3.0 M points: below the default 4 Mi threshold
        check_tree_level_step(api, port, ctx, wl, tvel, tcon, 101, "c1 tensor")
        ctx.set_tensor_grid(True)
        check_tree_level_step(api, port, ctx, wl, tvel, tcon, 101, "c1 default")
    finally:
        ctx.set_tensor_grid(True)
        ctx.set_stream(None)
        for t in tvel + [tcon]:
            t.destroy()


def test_config2_zalesak_full_size_tree_level_step(ctx, port):
    """The default bench workload through the very call the bench times."""
    torch, api, workloads = _mods()
    wl = workloads.make("c2", torch.device("cuda", 0))
    tcon, tvel = ctx.tree(wl.con), [ctx.tree(v) for v in wl.vel]
    ctx.set_stream(torch.cuda.current_stream())
    try:
        dev = check_tree_level_step(api, port, ctx, wl, tvel, tcon, 9973, "c2")
        # the host-buffer flavour of the same call (chunks of leaves, pipelined) is the same bits
        sub = wl.con.shard(1000, 3000)
        tsub = ctx.tree(sub)
        ctx.set_tensor_grid("always")  # chunks below the 4 Mi threshold keep the sum-factorised path
        a = api.SolveSemilagInSitu(api.NodeFieldFunctor(tvel[0]), tsub, 1, wl.dt, 1, wl.bc)
        ctx.set_host_chunks(3)
        b = api.SolveSemilagInSitu(api.NodeFieldFunctor(tvel[0]), tsub, 1, wl.dt, 1, wl.bc)
        ctx.set_host_chunks(0)
        assert np.array_equal(a, b)
        # ... and an asynchronous coefficient upload in front of the step changes nothing
        pinned = torch.from_numpy(np.ascontiguousarray(sub.coeff)).pin_memory()
        tsub.update_coeff(pinned.numpy(), wait=False)
        c = api.SolveSemilagInSitu(api.NodeFieldFunctor(tvel[0]), tsub, 1, wl.dt, 1, wl.bc)
        assert np.array_equal(a, c)
        tsub.destroy()
        del dev
    finally:
        ctx.set_host_chunks(0)
        ctx.set_tensor_grid(True)
        ctx.set_stream(None)
        for t in tvel + [tcon]:
            t.destroy()


def test_config3_time_varying_full_size_tree_level_step(ctx, port):
    torch, api, workloads = _mods()
    wl = workloads.make("c3", torch.device("cuda", 0))
    tcon, tvel = ctx.tree(wl.con), [ctx.tree(v) for v in wl.vel]
    ctx.set_stream(torch.cuda.current_stream())
    try:
        check_tree_level_step(api, port, ctx, wl, tvel, tcon, 8191, "c3 combined in time")
        ctx.set_time_combine(False)
        check_tree_level_step(api, port, ctx, wl, tvel, tcon, 8191, "c3 reference order")
    finally:
        ctx.set_time_combine(True)
        ctx.set_stream(None)
        for t in tvel + [tcon]:
            t.destroy()


def test_config5_uniform_full_size_tree_level_step(ctx, port):
    torch, api, workloads = _mods()
    wl = workloads.make("c5", torch.device("cuda", 0))
    assert wl.n_points == 110592000
    tcon, tvel = ctx.tree(wl.con), [ctx.tree(v) for v in wl.vel]
    ctx.set_stream(torch.cuda.current_stream())
    try:
        check_tree_level_step(api, port, ctx, wl, tvel, tcon, 4001, "c5")
    finally:
        ctx.set_stream(None)
        for t in tvel + [tcon]:
            t.destroy()


def test_config2_zalesak_full_size(ctx, port):
    torch, api, workloads = _mods()
    wl = workloads.make("c2", torch.device("cuda", 0))
    assert wl.con.n_leaf > 50000 and wl.q == 14
    tcon, tvel = ctx.tree(wl.con), ctx.tree(wl.vel[0])
    con, vel = api.NodeFieldFunctor(tcon), api.NodeFieldFunctor(tvel)
    ctx.set_stream(torch.cuda.current_stream())
    try:
        pos = tcon.collect_grid_points(device=True)
        n = pos.shape[0]
        assert n == wl.con.n_leaf * 15 ** 3
        # (1) interior arrival points locate to their own leaf (faces belong to the upper
        # neighbour, so only strictly interior nodes are checked)
        vals = torch.empty((n, 1), dtype=torch.float64, device=pos.device)
        leaf = torch.empty((n,), dtype=torch.int32, device=pos.device)
        ctx.check(ctx.lib.tbslas_b200_eval(tcon.h, 0, pos.data_ptr(), n, vals.data_ptr(), leaf.data_ptr(), 1))
        torch.cuda.synchronize()
        own = torch.arange(wl.con.n_leaf, device=pos.device, dtype=torch.int32).repeat_interleave(15 ** 3)
        k = torch.arange(15 ** 3, device=pos.device)
        interior = ((k % 15 > 0) & (k % 15 < 14) & ((k // 15) % 15 > 0) & ((k // 15) % 15 < 14) &
                    (k // 225 > 0) & (k // 225 < 14)).repeat(wl.con.n_leaf)
        assert bool((leaf[interior] == own[interior]).all())
        # (2) the whole step, device resident, against the oracle on a strided sample
        out = api.SolveSemilagRK2(vel, con, pos, 1, wl.dt, 1, wl.bc)
        torch.cuda.synchronize()
        idx = torch.arange(0, n, 9973, device=pos.device)
        sample = pos[idx].cpu().numpy()
        hv, hc = port.tree_create(wl.vel[0]), port.tree_create(wl.con)
        want = port.semilag_rk2(hv, hc, 1, sample, 1, wl.dt, 1, wl.bc)
        e = rel_err(out[idx].cpu().numpy(), want)
        print('[c2 point-array step] composed step error vs oracle %.2e' % e)
        assert e < 1e-10  # composed; the staged 1e-12 check is test_config2_zalesak_full_size_tree_level_step
        # (3) chunked host path == device path, bit for bit (sample of leaves)
        sub = pos[: 2000 * 3375].cpu().numpy()
        host = api.SolveSemilagRK2(vel, con, sub, 1, wl.dt, 1, wl.bc)
        assert np.array_equal(host, out[: sub.shape[0]].cpu().numpy())
    finally:
        ctx.set_stream(None)
        tcon.destroy()
        tvel.destroy()


def test_config3_time_varying_full_size(ctx):
    torch, api, workloads = _mods()
    wl = workloads.make("c3", torch.device("cuda", 0))
    assert len(wl.vel) == 4 and wl.bc == 1
    tcon = ctx.tree(wl.con)
    tv = [ctx.tree(v) for v in wl.vel]
    ctx.set_stream(torch.cuda.current_stream())
    try:
        pos = tcon.collect_grid_points(device=True)
        n = pos.shape[0]
        fset = api.FieldSetFunctor(tv, wl.vel_times)
        tq = 0.3 * wl.dt
        vc = fset(pos.clone(), time=tq, bc=1)   # default: coefficients combined in time, one evaluation
        ctx.set_time_combine(False)              # the reference's order: 4 evaluations + InterpCubic1D
        v = fset(pos.clone(), time=tq, bc=1)
        # the same from four single-tree evaluations + InterpCubic1D (cubic.h:28-56) in torch,
        # same operation order, un-fused: bit-exact
        p4 = [api.NodeFieldFunctor(t)(pos.clone(), bc=1) for t in tv]
        t0, t1, t2, t3 = wl.vel_times
        tt = (tq - t1) / (t2 - t1)
        h00 = ((2 * tt) * tt) * tt - (3 * tt) * tt + 1
        h10 = (tt * tt) * tt - (2 * tt) * tt + tt
        h01 = ((-2 * tt) * tt) * tt + (3 * tt) * tt
        h11 = (tt * tt) * tt - tt * tt
        m1 = (p4[2] - p4[1]) * 0.5 / (t2 - t1) + (p4[1] - p4[0]) * 0.5 / (t1 - t0)
        m2 = (p4[3] - p4[2]) * 0.5 / (t3 - t2) + (p4[2] - p4[1]) * 0.5 / (t2 - t1)
        want = h00 * p4[1] + (h10 * (t2 - t1)) * m1 + h01 * p4[2] + (h11 * (t2 - t1)) * m2
        torch.cuda.synchronize()
        assert float((v - want).abs().max()) <= 1e-15 * float(want.abs().max())
        # ... and the one-evaluation route agrees with it to rounding
        assert float((vc - want).abs().max()) <= 1e-13 * float(want.abs().max())
        # equal snapshots -> identity in time
        same = api.FieldSetFunctor([tv[1]] * 4, wl.vel_times)
        v1 = same(pos.clone(), time=tq, bc=1)
        assert float((v1 - p4[1]).abs().max()) <= 4e-16 * float(p4[1].abs().max())
        ctx.set_time_combine(True)
        v1c = same(pos.clone(), time=tq, bc=1)
        assert float((v1c - p4[1]).abs().max()) <= 1e-13 * float(p4[1].abs().max())
        # the full step runs and is finite at this size
        out = api.SolveSemilagRK2(fset, api.NodeFieldFunctor(tcon), pos, 1, wl.dt, 1, 1)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(out).all()) and out.shape == (n, 1)
    finally:
        ctx.set_time_combine(True)
        ctx.set_stream(None)
        tcon.destroy()
        for t in tv:
            t.destroy()


def test_config4_cubic_grid_full_size(ctx, port):
    import torch
    n_reg, dof = 256, 3
    dev = torch.device("cuda", 0)
    x = torch.linspace(0.0, 1.0, n_reg, dtype=torch.float64, device=dev)
    Z, Y, X = torch.meshgrid(x, x, x, indexing="ij")
    # per-axis cubics: reproduced exactly (to rounding) by the 4-point Lagrange stencil
    f = [1 + X - 2 * Y ** 3 + 0.5 * Z ** 2, X ** 3 * Y ** 2 * Z - Y, (X - 0.3) * (Y + 0.2) ** 3 * (Z - 0.7) ** 2]
    grid = torch.stack(f).contiguous()
    g = torch.Generator(device=dev).manual_seed(4)
    pts = torch.rand((1 << 24, 3), dtype=torch.float64, device=dev, generator=g) * 1.04 - 0.02
    ctx.set_stream(torch.cuda.current_stream())
    try:
        out = ctx.fast_interp(grid, dof, n_reg, pts)
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    inside = ((pts >= 0) & (pts <= 1)).all(dim=1)
    assert bool((out[~inside] == 0).all())
    px, py, pz = pts[:, 0], pts[:, 1], pts[:, 2]
    want = torch.stack([1 + px - 2 * py ** 3 + 0.5 * pz ** 2, px ** 3 * py ** 2 * pz - py,
                        (px - 0.3) * (py + 0.2) ** 3 * (pz - 0.7) ** 2], dim=1)
    assert float((out[inside] - want[inside]).abs().max()) < 1e-12
    # bit-exact against the oracle on a strided sample
    idx = torch.arange(0, pts.shape[0], 4099, device=dev)
    assert np.array_equal(out[idx].cpu().numpy(),
                          port.fast_interp(grid.cpu().numpy(), dof, n_reg, pts[idx].cpu().numpy()))


def test_config5_uniform_depth5_full_size(ctx):
    torch, api, workloads = _mods()
    q, depth = 14, 5
    coord, dd = ftm.uniform_leaves(depth)
    con = ftm.random_tree(coord, dd, q, 1, seed=2)
    zero = ftm.FlatTree(q, 3, *ftm.uniform_leaves(1), np.zeros((8, 3, ftm.ncoef(q))))
    tcon, tz = ctx.tree(con), ctx.tree(zero)
    ctx.set_stream(torch.cuda.current_stream())
    try:
        pos = tcon.collect_grid_points(device=True)
        n = pos.shape[0]
        assert n == 110592000
        cf, zf = api.NodeFieldFunctor(tcon), api.NodeFieldFunctor(tz)
        direct = cf(pos.clone(), bc=1)
        # zero velocity: departure points == arrival points (x + tau*0), so the step returns
        # exactly the field at the arrival points
        dep = torch.empty_like(pos)
        step = api.SolveSemilagRK2(zf, cf, pos, 7, 0.01, 2, 1, departure_points=dep)
        torch.cuda.synchronize()
        # (periodic: the x = 1 faces come back wrapped to 0, tree_functor.h:442-449)
        assert bool((dep == torch.where(pos >= 1.0, pos - 1.0, pos)).all())
        assert bool((step == direct).all())
        # leaf of every strictly interior arrival point is its own leaf (periodic wrap sends the
        # x = 1 faces to leaf column 0, so only interior nodes are checked)
        vals = torch.empty((n, 1), dtype=torch.float64, device=pos.device)
        leaf = torch.empty((n,), dtype=torch.int32, device=pos.device)
        ctx.check(ctx.lib.tbslas_b200_eval(tcon.h, 1, pos.data_ptr(), n, vals.data_ptr(), leaf.data_ptr(), 1))
        torch.cuda.synchronize()
        k = torch.arange(15 ** 3, device=pos.device)
        interior = ((k % 15 > 0) & (k % 15 < 14) & ((k // 15) % 15 > 0) & ((k // 15) % 15 < 14) &
                    (k // 225 > 0) & (k // 225 < 14))
        lf = leaf.view(con.n_leaf, 15 ** 3)[:, interior]
        own = torch.arange(con.n_leaf, device=pos.device, dtype=torch.int32)[:, None]
        assert bool((lf == own).all())
    finally:
        ctx.set_stream(None)
        tcon.destroy()
        tz.destroy()


def test_bench_line_contract_on_a_reduced_workload():
    """bench.py (B200 arm) on a shrunken C2: ONE JSON line with the keys the driver reads, the
    roofline and CPU-baseline objects, a non-zero launch count and consistent checksums between the
    two end-to-end flavours."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--scale", "2", "--steps", "2",
                        "--cpu-leaves", "64"], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline",
              "cpu_baseline"):
        assert k in d, k
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["warmup"] >= 3 and d["dtype"] == "f64"
    assert d["roofline"]["frac"] > 0 and d["roofline"]["peak"] > 0 and d["roofline"]["unit"] == "TFLOP/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert abs(e["checksum"] - e["point_array_call"]["checksum"]) <= 1e-9 * max(1.0, abs(e["checksum"]))

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
        assert rel_err(out[idx].cpu().numpy(), want) < 1e-11
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

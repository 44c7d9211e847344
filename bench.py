#!/usr/bin/env python
"""bench.py -- departure-point evals/sec of the semi-Lagrangian step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c2] [--scale S]

One "step" = one tbslas::SolveSemilagRK2 (nrk = 1) over every arrival point of the
advected tree: two velocity-tree evaluations + RK2 updates, then the scalar-tree
evaluation at the departure points (reference call stack: tree_semilag.h:124 ->
semilag.inc:27-45 -> traj.inc:49-68 -> tree_functor.h:397-690).

Default workload = BASELINE.json configs[1] ("c2": Zalesak slotted sphere, adaptive octree,
degree 14, max depth 7, 81 348 leaves, 274.5 M arrival points).  Prints ONE JSON line:
`value` = points/s with inputs resident in HBM; `e2e` = the same step through the C ABI with
pinned HOST buffers -- the tree-level call a reference driver makes (SolveSemilagInSitu:
coefficients up, arrival points generated in HBM, values down; `e2e.point_array_call` is the
SolveSemilagRK2 flavour with 24 B/point of arrival points over PCIe); `semilag_step` = the whole
tree-level step incl. the refit, on the device; `roofline` for the dominant kernel (Chebyshev
evaluation, FP64-pipe bound) from CUDA events recorded around every launch during the timed
region; `cpu_baseline` = the reference's own code (oracle/_ref) on this host's cores over a
bounded sample.

Multi-GPU (torchrun, one rank per GPU): the advected tree is split into equal contiguous
Morton ranges and foreign departure points travel by NCCL all-to-all-v inside the library; a
velocity tree of at most 1 GiB is held whole by every rank (--shard-velocity: co-partitioned
with the same split keys, the reference's layout); total work is fixed ("scaling": "strong").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def measured_traffic(workload, scale):
    """DRAM bytes per launch of the evaluation kernel from the committed ncu capture of this
    very command (profiles/r02_dram_cheb_eval_c2.csv: dram__bytes_read.sum + dram__bytes_write.sum
    of the three launches of one full-size C2 step), or None for other workloads."""
    p = os.path.join(ROOT, "profiles", "r02_dram_cheb_eval_c2.csv")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_dram_cheb_eval_c2.csv")
    if workload.lower() != "c2" or scale != 0 or not os.path.exists(p):
        return None, None
    import csv
    per = {}
    for row in csv.reader(open(p)):
        if len(row) > 14 and row[12].startswith("dram__bytes"):
            per[row[0]] = per.get(row[0], 0.0) + float(row[14])
    if not per:
        return None, None
    return sum(per.values()) / len(per), "profiles/%s (ncu, %d launches of one step)" % (os.path.basename(p), len(per))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 100 ms by a background process;
    only the samples whose timestamp falls inside the timed region are kept."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, pw, reasons = [], None, [], set()
            for _, r in rows:
                f = [x.strip() for x in r.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                    pw.append(float(f[3]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if f[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, pw, reasons
        # a sample printed at wall time t describes the ~100 ms before it
        inside = [r for r in self.rows if self.t0 is not None and self.t0 + 0.02 <= r[0] <= self.t1 + 0.12]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to the whole run
            inside, window = self.rows, "whole run (timed region shorter than the 100 ms period)"
        sm, mx, pw, reasons = parse(inside)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def workload_config(wl, world=1):
    """The keys both arms print (the driver compares them): what the workload IS, not how an arm runs
    it; plus `sample` (None on the B200 arm, which runs all of it)."""
    return {"workload": "%s: %s" % (wl.name, wl.desc), "leaves": wl.con.n_leaf, "points_total": wl.n_points,
            "q": wl.q, "bc": "periodic" if wl.bc else "freespace", "dt": wl.dt, "nrk": 1,
            "velocity_trees": len(wl.vel), "velocity_leaves": wl.vel[0].n_leaf}


# --------------------------------------------------------------------------- reference arm
def sample_points(wl, n_leaves, seed=0):
    """Arrival points of an evenly strided subset of the advected tree's leaves."""
    from tbslas_b200 import flat_tree as ftm
    L = wl.con.n_leaf
    n_leaves = min(n_leaves, L)
    idx = (np.arange(n_leaves) * (L / n_leaves)).astype(np.int64)
    return ftm.grid_points(wl.con.coord[idx], wl.con.depth[idx], wl.q), n_leaves


def cpu_run(wl, pts, steps, warmup, threads=None):
    """Time the reference's own SolveSemilagRK2 (oracle/_ref) or, if that prebuilt
    library is absent, the oracle port.  -> (pts/s, kind, cores, per-step seconds)."""
    from oracle import Oracle, have_ref
    kind = "reference" if have_ref() else "port"
    orc = Oracle("ref" if kind == "reference" else "port")
    cores = threads or os.cpu_count() or 1
    orc.set_num_threads(cores)
    hv = [orc.tree_create(v) for v in wl.vel]
    hc = orc.tree_create(wl.con)
    ts = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if len(hv) == 1:
            orc.semilag_rk2(hv[0], hc, 1, pts, 1, wl.dt, 1, wl.bc)
        else:
            orc.semilag_rk2(hv, hc, 1, pts, 1, wl.dt, 1, wl.bc, kind="set4", times=wl.vel_times)
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    mean = float(np.mean(ts))
    return pts.shape[0] / mean, kind, cores, mean


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from tbslas_b200 import workloads
    if args.workload.lower() == "c4":
        return run_reference_cubic(args)
    wl = workloads.make(args.workload, None, args.scale)
    pts, nl = sample_points(wl, args.cpu_leaves)
    rate, kind, cores, sec = cpu_run(wl, pts, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "departure-point evals/sec", "value": rate,
        "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(wl),
                       sample={"leaves": nl, "points_per_step": int(pts.shape[0]),
                               "what": "each step = SolveSemilagRK2 over the arrival points of %d evenly "
                                       "strided leaves of the %d (full trees): the whole workload would take "
                                       "~%.0f s per step on these cores" % (nl, wl.con.n_leaf, wl.n_points / rate)}),
        "cpu_baseline": {"value": rate, "unit": "points/s", "cores": cores, "kind": kind,
                         "sample": "SolveSemilagRK2 on the arrival points of %d evenly strided leaves "
                                   "(%d points), full trees; single-rank OpenMP path over a PVFMM "
                                   "stand-in (no MPI in this image)" % (nl, pts.shape[0])},
        "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_reference_cubic(args):
    """C4 on the host: the reference's fast_interp (serial loop) on a bounded sample."""
    import torch
    from oracle import Oracle, have_ref
    n_reg, dof = max(8, 256 >> args.scale), 3
    grid, pts = cubic_workload(torch.device("cpu"), n_reg, dof)
    kind = "reference" if have_ref() else "port"
    orc = Oracle("ref" if kind == "reference" else "port")
    n = pts.shape[0]
    m = min(n, 1 << 21)
    sp = pts.numpy()[:: max(1, n // m)][:m].copy()
    g = grid.numpy()
    ts = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.fast_interp(g, dof, n_reg, sp)
        if it >= args.warmup:
            ts.append(time.perf_counter() - t0)
    sec = float(np.mean(ts))
    rate = sp.shape[0] / sec
    emit({
        "impl": "reference", "metric": "departure-point evals/sec", "value": rate, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "c4: Taylor-Green velocity, uniform cubic-grid interpolation (fast_interp), "
                               "%d^3 nodes x dof %d" % (n_reg, dof), "points_total": n,
                   "sample": "%d strided query points per step" % sp.shape[0]},
        "cpu_baseline": {"value": rate, "unit": "points/s", "cores": 1, "kind": kind,
                         "sample": "fast_interp on %d strided query points; the reference's loop is serial "
                                   "(tree_functor.h:106)" % sp.shape[0]},
        "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})
    return 0


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist
    from tbslas_b200 import api, workloads
    from tbslas_b200 import flat_tree as ftm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = workloads.make(args.workload, dev, args.scale)
    ctx = api.Context(local)
    ctx.set_stream(torch.cuda.current_stream())
    if args.tensor_grid:
        ctx.set_tensor_grid("always" if args.tensor_grid == "always" else False)
    if world > 1:
        ctx.comm_init_torch()
        if args.exchange:
            ctx.comm_set_exchange(args.exchange)
        # the advected tree in equal-count contiguous Morton ranges; the velocity tree
        # re-partitioned with the SAME break points, whole leaves by their own Morton id
        # (what tbslas::MergeTree + RedistNodes do, tree_utils.h:703-728)
        first = workloads.partition_leaves(wl.con.n_leaf, world)
        splitters = wl.con.keys()[first[:-1]]
        con_local = wl.con.shard(int(first[rank]), int(first[rank + 1]))
        vel_bytes = sum(v.coeff.nbytes for v in wl.vel)
        # a velocity tree that is small next to 180 GB of HBM is held whole by every rank
        # (SURVEY 8(f) f3): its two evaluations per step then need no point exchange; only the
        # advected tree -- the big one -- is sharded and exchanged.  --shard-velocity forces the
        # reference's layout (every tree partitioned by the same Morton break points).
        replicate = (args.replicate_velocity or vel_bytes <= (1 << 30)) and not args.shard_velocity
        args.replicate_velocity = replicate
        vel_local = wl.vel if replicate else \
            [workloads.shard_by_splitters(v, splitters, rank) for v in wl.vel]
    else:
        con_local, vel_local = wl.con, wl.vel
    tcon = ctx.tree(con_local)
    tvel = [ctx.tree(v, replicated=(world > 1 and args.replicate_velocity)) for v in vel_local]
    con_f = api.NodeFieldFunctor(tcon)
    vel_f = api.NodeFieldFunctor(tvel[0]) if len(tvel) == 1 else api.FieldSetFunctor(tvel, wl.vel_times)

    pos = tcon.collect_grid_points(device=True)  # this rank's arrival points, HBM resident
    n_local = pos.shape[0]
    vals = torch.empty((n_local, 1), dtype=torch.float64, device=dev)
    n_total = wl.n_points

    def step_dev():
        # the tree-level step on device buffers: steps (1)+(2) of tbslas::SolveSemilagInSitu --
        # arrival points generated in HBM from the advected tree's leaves, RK2 trajectories, the
        # scalar at the departure points -> vals (HBM)
        ctx.check(ctx.lib.tbslas_b200_semilag_insitu(
            api.C.byref(vel_f.field), None, tcon.h, wl.bc, 1, float(wl.dt), 1, vals.data_ptr(), 1))

    def step_points():
        # the same on an explicit point array (tbslas::SolveSemilagRK2): no structure to exploit
        api.SolveSemilagRK2(vel_f, con_f, pos, 1, wl.dt, 1, wl.bc, points_vals=vals)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_dev()
    barrier()
    peak_fp64 = ctx.fp64_peak(5) if rank == 0 else 0.0
    barrier()

    # ---- timed region: device-resident inputs ------------------------------------------
    ctx.profile_reset()
    ctx.profile_enable(True)
    launches0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    exch = ctx.comm_last_exchange() if world > 1 else (0, 0)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.kernel_launches() - launches0
    prof = ctx.profile()
    ctx.profile_enable(False)
    per_rank = None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        mine = {k: round(v["ms"] / args.steps, 3) for k, v in prof.items() if v["ms"] > 0}
        mine["points"] = n_local
        mine["sent_last_eval"], mine["received_last_eval"] = exch
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    m_exc = ctx.last_grid_exceptions()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def checksums(v):
        """(float sum, sum of the values' BIT PATTERNS mod 2^64) over all ranks.  Integer addition
        wraps and commutes, so the second is independent of the partition and of summation order:
        bit-identical values give the same number at N = 1, 2, 4, 8."""
        f = v.sum().reshape(1).to(dev)
        b = v.contiguous().view(torch.int64).sum().reshape(1).to(dev)
        if world > 1:
            dist.all_reduce(f, op=dist.ReduceOp.SUM)
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
        return float(f.item()), int(b.item())
    checksum_dev, checksum_bits = checksums(vals)

    # ---- parity of the timed path against the oracle, outside the timed region ------------
    # A strided sample of THIS rank's arrival points through the oracle (CPU restatement of the
    # reference, full trees): departure points, values at the GPU's departure points, composed step.
    parity = None
    if not args.no_parity:
        from oracle import Oracle
        orc = Oracle("port")
        pv, pdep = api.SolveSemilagInSitu(vel_f, tcon, 1, wl.dt, 1, wl.bc, device=True, departure_points=True)
        torch.cuda.synchronize()
        same_bits = bool(torch.equal(pv, vals))
        n_s = min(args.parity_points, n_local)
        e_dep = e_val = e_step = 0.0
        if n_s:
            idx = torch.arange(n_s, device=dev, dtype=torch.int64) * (n_local // n_s)
            arr, dep_g, got = pos[idx].cpu().numpy(), pdep[idx].cpu().numpy(), pv[idx].cpu().numpy()
            hv = [orc.tree_create(v) for v in wl.vel]
            hc = orc.tree_create(wl.con)
            kind, vh = ("steady", hv[0]) if len(hv) == 1 else ("set4", hv)
            dep_o = orc.traj_rk2(vh, arr, wl.dt, 0.0, 1, wl.bc, kind=kind, times=wl.vel_times)
            want_g, _, _ = orc.eval_tree(hc, 1, dep_g, wl.bc, want_leaf=False)
            want = orc.semilag_rk2(vh, hc, 1, arr, 1, wl.dt, 1, wl.bc, kind=kind, times=wl.vel_times)
            sc = max(float(np.abs(want).max()), 1e-300)
            e_dep = float(np.abs(dep_g - dep_o).max())
            e_val = float(np.abs(got - want_g).max()) / sc
            e_step = float(np.abs(got - want).max()) / sc
        del pv, pdep
        parity = {"n_sample": int(n_s) * world, "oracle": "oracle/tbslas_oracle.c (port), full trees",
                  "max_abs_err_departure_points": allmax(e_dep),
                  "max_rel_err_values_at_equal_points": allmax(e_val),
                  "max_rel_err_vs_oracle": allmax(e_step),
                  "call_repeatable_bitwise": bool(allmax(0.0 if same_bits else 1.0) == 0.0),
                  "what": "tbslas_b200_semilag_insitu (the timed call) on device buffers vs the oracle on a "
                          "strided sample of every rank's arrival points; errors relative to max |value|"}

    # the point-array flavour of the same step, device resident (what `value` was before the
    # tree-level call learnt to use the tensor structure of the arrival points)
    step_points()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_pts_steps = max(1, min(args.steps, 3))
    p0.record()
    for _ in range(n_pts_steps):
        step_points()
    p1.record()
    barrier()
    pts_ms = allmax(p0.elapsed_time(p1) / n_pts_steps)

    # ---- the other half of the metric: semi-Lagrangian step time of the tree-level call ---
    # tbslas::SolveSemilagInSitu (tree_semilag.h:92-135) entirely on the device: arrival points
    # generated in HBM, advected, and the advected tree's coefficients refitted (one FP64
    # tensor-core GEMM); run on a scratch copy of the tree so the timed workload stays the same.
    ctx.set_pt2coeff(wl.q)
    scratch = ctx.tree(con_local)
    for _ in range(2):
        api.SolveSemilagInSituUpdate(vel_f, scratch, 1, wl.dt, 1, wl.bc)
        scratch.update_coeff(con_local.coeff)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    n_step = max(1, min(args.steps, 3))
    for _ in range(n_step):
        s0.record()
        api.SolveSemilagInSituUpdate(vel_f, scratch, 1, wl.dt, 1, wl.bc)
        s1.record()
        barrier()
        tot += s0.elapsed_time(s1)
        scratch.update_coeff(con_local.coeff)
    step_ms = allmax(tot / n_step)
    scratch.destroy()

    # ---- every tree Morton-sharded (the reference's layout, north_star): N > 1 only --------
    sharded = None
    if world > 1 and args.replicate_velocity and not args.no_sharded:
        svel = [ctx.tree(workloads.shard_by_splitters(v, splitters, rank)) for v in wl.vel]
        svel_f = api.NodeFieldFunctor(svel[0]) if len(svel) == 1 else api.FieldSetFunctor(svel, wl.vel_times)
        svals = torch.empty_like(vals)

        def step_sharded():
            ctx.check(ctx.lib.tbslas_b200_semilag_insitu(
                api.C.byref(svel_f.field), None, tcon.h, wl.bc, 1, float(wl.dt), 1, svals.data_ptr(), 1))
        for _ in range(2):
            step_sharded()
        barrier()
        ctx.profile_reset()
        ctx.profile_enable(True)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            step_sharded()
        q1.record()
        barrier()
        sh_ms = allmax(q0.elapsed_time(q1) / args.steps)
        sprof = ctx.profile()
        ctx.profile_enable(False)
        s_exc = ctx.last_grid_exceptions()
        sent, recv = ctx.comm_last_exchange()
        s_sum, s_bits = checksums(svals)
        mine = {"exchange_ms": round(sprof["Exchange"]["ms"] / args.steps, 3),
                "unpack_ms": round(sprof["Unpack"]["ms"] / args.steps, 3),
                "sent_last_eval": sent, "received_last_eval": recv, "grid_exceptions": s_exc,
                "velocity_leaves_local": svel[0].n_leaf}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        sharded = {"value": n_total / (sh_ms * 1e-3), "unit": "points/s", "ms_per_step": sh_ms,
                   "checksum_bits": s_bits, "bit_identical_to_replicated_velocity": bool(s_bits == checksum_bits),
                   "per_rank": gathered,
                   "what": "the same tree-level step with EVERY tree partitioned by the advected tree's Morton "
                           "break points (tree_functor.h:433-437,491-596): three collective evaluations per "
                           "step; the first velocity evaluation still runs by sum factorisation where the "
                           "containing velocity leaf is local"}
        for t in svel:
            t.destroy()

    # ---- end to end: pinned host buffers through the C ABI -----------------------------
    h_pos = torch.empty((n_local, 3), dtype=torch.float64, pin_memory=True)
    h_pos.copy_(pos)
    h_vals = torch.empty((n_local, 1), dtype=torch.float64, pin_memory=True)
    np_pos, np_vals = h_pos.numpy(), h_vals.numpy()

    def step_host():
        api.SolveSemilagRK2(vel_f, con_f, np_pos, 1, wl.dt, 1, wl.bc, points_vals=np_vals)

    step_host()
    barrier()
    t0 = time.perf_counter()
    pa_steps = max(1, min(args.steps, 3))
    for _ in range(pa_steps):
        step_host()  # returns after the values have landed in host memory
    barrier()
    e2e_s = allmax((time.perf_counter() - t0) / pa_steps)
    checksum = float(np_vals.sum())

    # ---- end to end, tree-level call: what a reference driver does per step through the drop-in
    # adaptor of tbslas::SolveSemilagInSitu (tree_semilag.h:92-135) -- upload the advected tree's
    # coefficients from (pinned) host memory, run the step, read the new grid values back.  No
    # point ever crosses PCIe: the arrival points are generated in HBM.  The upload is asynchronous
    # (tbslas_b200_tree_update_coeff_async): only the third evaluation of the step reads the advected
    # tree, so the copy hides behind the two velocity evaluations.
    nc = ftm.ncoef(wl.q)
    h_coef = torch.empty((con_local.n_leaf, 1, nc), dtype=torch.float64, pin_memory=True)
    h_coef.copy_(torch.from_numpy(np.ascontiguousarray(con_local.coeff)))
    np_coef = h_coef.numpy()

    def step_tree():
        tcon.update_coeff(np_coef, wait=False)
        ctx.check(ctx.lib.tbslas_b200_semilag_insitu(
            api.C.byref(vel_f.field), None, tcon.h, wl.bc, 1, float(wl.dt), 1, np_vals.ctypes.data, 0))

    for _ in range(2):
        step_tree()
    barrier()
    # raw PCIe rate of this rank's values, all ranks copying at once (what bounds the copy-out)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    h_vals.copy_(vals, non_blocking=True)
    c1.record()
    barrier()
    d2h_gbps = -allmax(-(n_local * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9)) if n_local else 0.0
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_tree()  # returns after the values have landed in host memory
    barrier()
    tree_s = allmax((time.perf_counter() - t0) / e2e_steps)
    checksum_tree = float(np_vals.sum())
    h_sum, h_bits = checksums(torch.from_numpy(np_vals))

    # ---- the whole tbslas::SolveSemilagInSitu through host memory: what the C++ drop-in adaptor does
    # per step -- coefficients up, arrival points + RK2 + scalar + refit on the device, the NEW
    # COEFFICIENTS down.  Only 2 x 8*Ncoef/P = 3.2 B per point cross PCIe instead of 8 B of values.
    scratch2 = ctx.tree(con_local)
    h_cout = torch.empty((con_local.n_leaf, 1, nc), dtype=torch.float64, pin_memory=True)
    np_cout = h_cout.numpy()

    def step_tree_coeff():
        scratch2.update_coeff(np_coef, wait=False)
        api.SolveSemilagInSituUpdate(vel_f, scratch2, 1, wl.dt, 1, wl.bc)
        ctx.check(ctx.lib.tbslas_b200_tree_get_coeff(scratch2.h, np_cout.ctypes.data, 0))  # returns when landed

    for _ in range(2):
        step_tree_coeff()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_tree_coeff()
    barrier()
    coef_s = allmax((time.perf_counter() - t0) / e2e_steps)
    coef_sum = float(np_cout.sum())
    scratch2.destroy()
    exch_mode = ctx.comm_exchange_mode() if world > 1 else ("none", 0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -----------------------------------------------
    hbm_peak, hbm_src = load_peaks()
    ev = prof["ChebEval"]
    # dof-3 launches of a step: 2 per RK stage pair when the snapshots' coefficients are combined in
    # time (one evaluation per stage), 2 * len(tvel) when every snapshot is evaluated
    combined = len(tvel) > 1 and all(np.array_equal(v.keys(), wl.vel[0].keys()) and
                                     np.array_equal(v.depth, wl.vel[0].depth) for v in wl.vel)
    n_eval_vel = 2 if (len(tvel) == 1 or combined) else 2 * len(tvel)
    tensor = ("TensorGrid" in prof and prof["TensorGrid"]["ms"] > 0)
    if tensor:  # the first velocity evaluation ran by sum factorisation; only its exceptions
        n_eval_vel -= 1  # (m_exc points) went through the Chebyshev kernel
    # algorithmic work per launch, summed over the launches of the timed region
    P = (wl.q + 1) ** 3
    n_exc = m_exc if tensor else 0
    flops = args.steps * (n_local * (n_eval_vel * workloads.flops_per_point_eval(wl.q, 3)
                                     + workloads.flops_per_point_eval(wl.q, 1))
                          + n_exc * workloads.flops_per_point_eval(wl.q, 3))
    dfma_flops = 2 * (ftm.ncoef(wl.q) - 1)  # executed by the contraction per point and component
    flops_exec = args.steps * (n_local * (n_eval_vel * 3 + 1) + n_exc * 3) * dfma_flops
    coef_bytes = 8.0 * ftm.ncoef(wl.q) / P
    bytes_alg = args.steps * (n_local * (n_eval_vel * (24 + 24 + 3 * coef_bytes) + (24 + 8 + coef_bytes))
                              + n_exc * (24 + 24 + 3 * coef_bytes))
    ach_tf = flops / (ev["ms"] * 1e-3) * 1e-12 if ev["ms"] > 0 else 0.0
    ach_gbs = bytes_alg / (ev["ms"] * 1e-3) * 1e-9 if ev["ms"] > 0 else 0.0
    stage_ms = {k: round(v["ms"] / args.steps, 4) for k, v in prof.items() if v["ms"] > 0}
    traffic, traffic_src = measured_traffic(args.workload, args.scale)
    roofline = {
        "bound": "fp64", "kernel": "cheb_eval_kernel<q=%d>" % wl.q, "achieved": ach_tf,
        "peak": peak_fp64, "unit": "TFLOP/s", "frac": ach_tf / peak_fp64 if peak_fp64 else None,
        "frac_nominal": ach_tf / 37.2,
        "peak_source": "measured in this run: DFMA-only kernel, best of 5 (MEASURED_PEAKS.json has "
                       "no FP64 entry; nominal 148 SM x 64 DFMA/clk x 1.965 GHz = 37.2)",
        "flops_model": "reference's own: N*(9d + 2*dof*Ncoef), tree_functor.h:389-394",
        # what the kernel executes: one DFMA per coefficient but the first (T_0 = 1 inside the leaf, so
        # fma(1, c, 0) is c itself: Ncoef - 1 DFMAs per point and component), i.e. LESS than the model's
        # 2*Ncoef + closes -- `frac` (model flops over the DFMA peak) can therefore exceed the pipe's own
        # utilisation, which is `frac_executed_dfma`
        "frac_executed_dfma": (flops_exec / (ev["ms"] * 1e-3) * 1e-12 / peak_fp64)
        if (peak_fp64 and ev["ms"] > 0) else None,
        "avg_launch_ms": ev["ms"] / max(1, ev["launches"]), "launches": ev["launches"],
        "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": bytes_alg / max(1, ev["launches"]),
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "peak_source": hbm_src,
                "note": "algorithmic bytes (24 xyz + 8*dof out + coefficients once per leaf); the "
                        "kernel is FP64-pipe bound for q >= 6, so this fraction is low by construction"},
        "share_of_step": ev["ms"] / ms if ms > 0 else None,
        "stage_ms_per_step": stage_ms,
    }

    line = {
        "metric": "departure-point evals/sec", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        # the same keys in both arms; how THIS arm ran the workload is in `run_config`
        "config": dict(workload_config(wl), sample=None),
        "run_config": dict(
            points_per_gpu=n_local,
            step="tbslas_b200_semilag_insitu on device buffers: arrival points generated in HBM, "
                 "RK2 trajectories, scalar at the departure points",
            velocity_evaluations_per_step=n_eval_vel + (1 if tensor else 0),
            first_velocity_evaluation=None if not tensor else
            "sum factorisation over the arrival grids (tensor_eval.cu); %d of %d points per step "
            "(on velocity-leaf faces) through the generic kernels" % (n_exc, n_local),
            time_interpolation=None if not combined else
            "coefficients of the 4 snapshots (one leaf list) combined in time on the device, one "
            "evaluation per RK stage; tbslas_b200_set_time_combine(ctx, 0) evaluates every snapshot",
            l2_policy="inputs (%.1f GB of points per step) exceed the 126 MB L2" % (n_local * 24 / 1e9),
            partition="single GPU" if world == 1 else
            "equal-count contiguous Morton ranges of the advected tree; velocity "
            + ("tree replicated on every rank (no exchange for velocity evaluations)"
               if args.replicate_velocity else
               "tree re-partitioned with the same break points (whole leaves)"),
            exchange=None if world == 1 else
            {"mode": exch_mode[0], "mailbox_points": exch_mode[1],
             "what": "peer: outsiders written straight into the owner's receive buffer over NVLink peer "
                     "memory, values straight back, counts/offsets/barriers on the device (no host "
                     "synchronisation, no NCCL kernel in the step); nccl: all-to-all-v by grouped "
                     "ncclSend/ncclRecv" ,
             "rank0_last_eval_sent": exch[0], "rank0_last_eval_received": exch[1]}),
        "checksum_global": {"sum": checksum_dev, "bits": checksum_bits,
                            "what": "over ALL ranks' advected values of the timed step: float sum, and the sum "
                                    "of the values' 64-bit patterns mod 2^64 (order- and partition-independent: "
                                    "equal at N = 1, 2, 4, 8 iff the values are bit-identical)"},
        "parity_check": parity,
        "clocks": clocks, "gpu_launches": int(launches),
        # headline end-to-end number: the tree-level call every reference driver makes
        # (SolveSemilagInSitu, advection.cpp:296) through the C ABI with pinned HOST buffers
        "e2e": {"value": n_total / tree_s, "unit": "points/s", "ms_per_step": tree_s * 1e3,
                "h2d_bytes_per_step": int(con_local.n_leaf * nc * 8), "d2h_bytes_per_step": int(n_local * 8),
                "steps": e2e_steps, "pcie_d2h_GBps_slowest_rank": d2h_gbps,
                "checksum": checksum_tree, "checksum_bits": h_bits,
                "bit_identical_to_device_buffers": bool(h_bits == checksum_bits),
                "what": "tbslas_b200_tree_update_coeff_async + tbslas_b200_semilag_insitu (SolveSemilagInSitu "
                        "steps 1-2, tree_semilag.h:92-130): the advected tree's coefficients up from pinned "
                        "host memory (under the velocity evaluations), arrival points generated in HBM, advected "
                        "grid values down to pinned host memory; leaf chunks pipelined (D2H of chunk c-1 under "
                        "the kernels of chunk c); wall clock over all timed steps",
                "coefficients_out": {
                    "value": n_total / coef_s, "unit": "points/s", "ms_per_step": coef_s * 1e3,
                    "h2d_bytes_per_step": int(con_local.n_leaf * nc * 8),
                    "d2h_bytes_per_step": int(con_local.n_leaf * nc * 8), "checksum": coef_sum,
                    "what": "the whole tbslas::SolveSemilagInSitu (tree_semilag.h:92-135) through host memory, as "
                            "the C++ drop-in adaptor runs it: tbslas_b200_tree_update_coeff_async + "
                            "tbslas_b200_semilag_insitu_update (values refitted on the device, FP64 tensor-core "
                            "GEMM) + tbslas_b200_tree_get_coeff"},
                "point_array_call": {
                    "value": n_total / e2e_s, "unit": "points/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(n_local * 24), "d2h_bytes_per_step": int(n_local * 8),
                    "checksum": checksum,
                    "what": "tbslas_b200_semilag_rk2 (SolveSemilagRK2, semilag.inc:27-45) on pinned HOST "
                            "arrays: arrival points in (24 B/point over PCIe), advected values out; "
                            "chunks pipelined over three streams"}},
        "roofline": roofline,
        "value_point_array": {"value": n_total / (pts_ms * 1e-3), "unit": "points/s", "ms_per_step": pts_ms,
                              "what": "the same step through tbslas_b200_semilag_rk2 (SolveSemilagRK2) on an "
                                      "explicit, HBM-resident point array: all three evaluations point by point"},
        "semilag_step": {"ms": step_ms, "what": "SolveSemilagInSitu on the device: arrival-point generation "
                         "+ RK2 trajectories + scalar evaluation + values->coefficients refit, "
                         "coefficients stay in HBM (tree_semilag.h:92-135)"},
    }
    if per_rank is not None:
        line["per_rank_stage_ms"] = per_rank
    if sharded is not None:
        line["sharded_all_trees"] = sharded
    if world == 1 and not args.no_cpu:
        pts, nl = sample_points(wl, args.cpu_leaves)
        rate, kind, cores, sec = cpu_run(wl, pts, 1, 1)
        line["cpu_baseline"] = {
            "value": rate, "unit": "points/s", "cores": cores, "kind": kind,
            "sample": "SolveSemilagRK2 on the arrival points of %d evenly strided leaves (%d points, "
                      "%.1f s), full trees; reference's single-rank OpenMP path over a PVFMM stand-in"
                      % (nl, pts.shape[0], sec)}
    else:
        line["cpu_baseline"] = None
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------- C4: cubic grid
def cubic_workload(dev, n_reg=256, dof=3, dt=0.0628 / 4):
    """BASELINE config 4: Taylor-Green velocity sampled on a node-centred n_reg^3 grid,
    queries = the grid nodes displaced by -dt*v (one backward Euler step)."""
    import torch
    x = torch.linspace(0.0, 1.0, n_reg, dtype=torch.float64, device=dev)
    Z, Y, X = torch.meshgrid(x, x, x, indexing="ij")
    a, b, c = 2 * np.pi * X, 2 * np.pi * Y, 2 * np.pi * Z
    v = torch.stack([torch.cos(a) * torch.sin(b) * torch.sin(c),
                     torch.sin(a) * torch.cos(b) * torch.sin(c),
                     torch.sin(a) * torch.sin(b) * torch.cos(c)])[:dof].contiguous()  # [dof][z][y][x]
    pts = torch.stack([X, Y, Z], dim=-1).reshape(-1, 3) - dt * v.reshape(dof, -1).t()[:, :3]
    return v, pts.contiguous()


def run_b200_cubic(args):
    import torch
    from tbslas_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    dev = torch.device("cuda", 0)
    n_reg, dof = max(8, 256 >> args.scale), 3
    grid, pts = cubic_workload(dev, n_reg, dof)
    n = pts.shape[0]
    ctx = api.Context(0)
    ctx.set_stream(torch.cuda.current_stream())
    out = torch.empty((n, dof), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    res = ctx.grid(grid, dof, n_reg)  # the grid stays resident in HBM (tbslas_b200_grid_create)

    def step():
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res(pts, out=out)
        e1.record()
        return e0, e1
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = ctx.kernel_launches()
    sampler.mark_begin()
    evs = [step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    launches = ctx.kernel_launches() - launches0
    # end to end: pinned host buffers through the resident-grid handle -- queries in, values out
    # (chunks pipelined); `one_shot` is tbslas::fast_interp's own signature, the grid travelling too
    h_grid = torch.empty(grid.shape, dtype=torch.float64, pin_memory=True).copy_(grid)
    h_pts = torch.empty(pts.shape, dtype=torch.float64, pin_memory=True).copy_(pts)
    h_out = torch.empty((n, dof), dtype=torch.float64, pin_memory=True)
    res(h_pts.numpy(), out=h_out.numpy())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hv = res(h_pts.numpy(), out=h_out.numpy())
    e2e_s = (time.perf_counter() - t0) / args.steps
    ctx.fast_interp(h_grid.numpy(), dof, n_reg, h_pts.numpy(), out=h_out.numpy())
    t0 = time.perf_counter()
    ctx.fast_interp(h_grid.numpy(), dof, n_reg, h_pts.numpy(), out=h_out.numpy())
    one_shot_s = time.perf_counter() - t0
    hbm_peak, hbm_src = load_peaks()
    bytes_alg = n * (24 + 8 * dof + 8 * dof)
    ach = bytes_alg / (ms * 1e-3) * 1e-9
    line = {
        "metric": "departure-point evals/sec", "value": n / (ms * 1e-3), "unit": "points/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c4: Taylor-Green velocity, uniform cubic-grid interpolation (fast_interp), "
                               "%d^3 nodes x dof %d" % (n_reg, dof), "points_total": n,
                   "l2_policy": "256 MB written between timed iterations (L2 flush); grid is %.0f MB" %
                                (grid.numel() * 8 / 1e6)},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": n / e2e_s, "unit": "points/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(n * 24), "d2h_bytes_per_step": int(n * dof * 8),
                "checksum": float(hv.sum()), "steps": args.steps,
                "what": "tbslas_b200_grid_eval on pinned host arrays, grid resident in HBM",
                "one_shot": {"value": n / one_shot_s, "ms_per_step": one_shot_s * 1e3,
                             "h2d_bytes_per_step": int(grid.numel() * 8 + n * 24),
                             "what": "tbslas_b200_cubic_eval: the %.0f MB grid crosses PCIe with every call"
                                     % (grid.numel() * 8 / 1e6)}},
        "roofline": {"bound": "hbm", "kernel": "cubic_grid_kernel", "achieved": ach, "peak": hbm_peak,
                     "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": hbm_src, "traffic": None,
                     "bytes_model": "24 (xyz) + 8*dof (out) + 8*dof (each grid node once), SURVEY 8(d)",
                     "avg_launch_ms": ms, "launches": args.steps},
    }
    if not args.no_cpu:
        from oracle import Oracle, have_ref
        kind = "reference" if have_ref() else "port"
        orc = Oracle("ref" if kind == "reference" else "port")
        m = min(n, 1 << 21)
        sp = h_pts.numpy()[:: max(1, n // m)][:m].copy()
        g = h_grid.numpy()
        t0 = time.perf_counter()
        orc.fast_interp(g, dof, n_reg, sp)
        sec = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sp.shape[0] / sec, "unit": "points/s", "cores": 1, "kind": kind,
                                "sample": "fast_interp on %d strided query points (%.1f s); the reference's loop "
                                          "is serial (tree_functor.h:106)" % (sp.shape[0], sec)}
    emit(line)
    return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything that libraries print to stdout while the bench runs (NCCL's version banner,
    torchrun notices) goes to stderr; the one JSON line is written to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--scale", type=int, default=0, help="shrink the workload (tests)")
    ap.add_argument("--cpu-leaves", type=int, default=8192,
                    help="leaves in the CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle sample (parity_check)")
    ap.add_argument("--parity-points", type=int, default=2048, help="oracle sample per rank")
    ap.add_argument("--no-sharded", action="store_true",
                    help="multi-GPU: skip the extra measurement with every tree Morton-sharded")
    ap.add_argument("--tensor-grid", default=None, choices=["always", "off"],
                    help="first velocity evaluation of the tree-level step: sum factorisation at any size / never "
                         "(default: on from 4 Mi points per call)")
    ap.add_argument("--exchange", default=None, choices=["peer", "nccl"],
                    help="multi-GPU: how outsiders travel (default: peer memory where available)")
    ap.add_argument("--replicate-velocity", action="store_true",
                    help="multi-GPU: every rank holds the whole velocity tree (SURVEY 8(f) row f3); "
                         "default when the velocity trees take <= 1 GiB")
    ap.add_argument("--shard-velocity", action="store_true",
                    help="multi-GPU: partition the velocity tree like the advected tree (reference layout)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.workload.lower() == "c4" and args.impl == "b200":
        sys.exit(run_b200_cubic(args))
    sys.exit(run_reference(args) if args.impl == "reference" else run_b200(args))


if __name__ == "__main__":
    main()

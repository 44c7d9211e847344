"""tools/e2e_sweep.py -- where does the tree-level host call's time go?  (WORKLOAD=c2|c3|c5, one GPU)
Sweeps the number of host chunks, prints wall time per call, the per-stage device times and
raw PCIe copy rates.  Diagnostic only; results are summarised under profiles/."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tbslas_b200 import api, workloads  # noqa: E402
from tbslas_b200 import flat_tree as ftm  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
scale = int(os.environ.get("SCALE", "0"))
wl = workloads.make(os.environ.get("WORKLOAD", "c2"), dev, scale)
ctx = api.Context(0)
ctx.set_stream(torch.cuda.current_stream())
tcon, tvels = ctx.tree(wl.con), [ctx.tree(v) for v in wl.vel]
vel = api.NodeFieldFunctor(tvels[0]) if len(tvels) == 1 else api.FieldSetFunctor(tvels, wl.vel_times)
n = wl.n_points
h_vals = torch.empty((n, 1), dtype=torch.float64, pin_memory=True)
nc = ftm.ncoef(wl.q)
h_coef = torch.from_numpy(np.ascontiguousarray(wl.con.coeff)).pin_memory()
d_vals = torch.empty((n, 1), dtype=torch.float64, device=dev)
out = {}
# raw copies
for name, fn in (("d2h", lambda: h_vals.copy_(d_vals, non_blocking=True)),
                 ("h2d", lambda: d_vals.copy_(h_vals, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[name + "_GBps"] = n * 8 / dt / 1e9


def step_dev():
    ctx.check(ctx.lib.tbslas_b200_semilag_insitu(api.C.byref(vel.field), None, tcon.h, wl.bc, 1, float(wl.dt), 1,
                                                 d_vals.data_ptr(), 1))


def step_tree(async_up=True):
    tcon.update_coeff(h_coef.numpy(), wait=not async_up)
    ctx.check(ctx.lib.tbslas_b200_semilag_insitu(api.C.byref(vel.field), None, tcon.h, wl.bc, 1, float(wl.dt), 1,
                                                 h_vals.numpy().ctypes.data, 0))


for _ in range(3):
    step_dev()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step_dev()
torch.cuda.synchronize()
out["device_ms"] = (time.perf_counter() - t0) / 3 * 1e3
for K in [int(k) for k in os.environ.get("KS", "1,2,3,4,6,8,12,16").split(",")]:
    ctx.set_host_chunks(K)
    for _ in range(2):
        step_tree()
    torch.cuda.synchronize()
    ctx.profile_reset(); ctx.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(3):
        step_tree()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    prof = ctx.profile(); ctx.profile_enable(False)
    out["K=%d" % K] = {"ms": round(ms, 2), "stages": {k: round(v["ms"] / 3, 2) for k, v in prof.items() if v["ms"] > 0}}
    # without profiling events
    t0 = time.perf_counter()
    for _ in range(3):
        step_tree()
    torch.cuda.synchronize()
    out["K=%d" % K]["ms_noprof"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
    t0 = time.perf_counter()
    for _ in range(3):
        step_tree(False)
    torch.cuda.synchronize()
    out["K=%d" % K]["ms_sync_upload"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
print(json.dumps(out, indent=1))

"""tools/eval_degree_sweep.py -- point-by-point evaluation rate per Chebyshev degree and dof (one GPU).
Random coefficients on a uniform depth-3 tree, 16 Mi random points resident in HBM, CUDA events around
`NodeFieldFunctor` calls.  Shows where the fully unrolled persistent kernel (q <= 14, and a coefficient block
that fits eight warps per SM) hands over to the degree-generic kernel (q = 15..19, or wide dof).  Diagnostic
only; the result is summarised under profiles/."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tbslas_b200 import api  # noqa: E402
from tbslas_b200 import flat_tree as ftm  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ctx = api.Context(0)
ctx.set_stream(torch.cuda.current_stream())
n = 1 << 24
pts = torch.rand((n, 3), dtype=torch.float64, device=dev)
coord, depth = ftm.uniform_leaves(3)
rows = []
for q, dof in ((4, 1), (8, 1), (8, 3), (12, 3), (14, 1), (14, 3), (14, 4), (15, 1), (15, 3), (16, 3), (19, 1), (19, 3)):
    tree = ctx.tree(ftm.random_tree(coord, depth, q, dof, seed=q))
    f = api.NodeFieldFunctor(tree)
    out = torch.empty((n, dof), dtype=torch.float64, device=dev)
    for _ in range(2):
        f(pts, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        f(pts, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ncoef = ftm.ncoef(q)
    rows.append({"q": q, "dof": dof, "ms_per_call": round(ms, 3), "Gpts_per_s": round(n / ms * 1e-6, 3),
                 "model_TFLOPs": round(n * (9 * (q + 1) + 2 * dof * ncoef) / ms * 1e-9, 2),
                 "what": "locate + bin + evaluation of 16 Mi points"})
    tree.destroy()
print(json.dumps(rows, indent=1))

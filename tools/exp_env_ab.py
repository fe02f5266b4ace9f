#!/usr/bin/env python
"""Alternating A/B of the pair stage under an environment switch read at context creation or planning time
(thermal / power drift cancels):  python tools/exp_env_ab.py c5 WLD_EXPERIMENT_VM 0 1"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bench
import weightedld_b200 as wld
from weightedld_b200 import _lib as L

wl, var, vals = sys.argv[1], sys.argv[2], sys.argv[3:]
chars = torch.from_numpy(bench.make_input(wl)).cuda()
ctxs, surv = {}, {}
for v in vals:
    os.environ[var] = v
    c = wld.Context(0)
    c.set_screen("always")
    c.load_alignment(chars)
    c.filter_sites(*bench.FILTER)
    c.henikoff()
    surv[v] = c.ld_pairs(bench.R2_THRESHOLD)[0]
    ctxs[v] = c
res = {v: [] for v in vals}
for rep in range(8):
    for v in vals:
        ctxs[v].ld_pairs(bench.R2_THRESHOLD)
        res[v].append(round(ctxs[v].stage_ms(L.STAGE_PAIR), 3))
print(json.dumps({"workload": wl, "switch": var, "survivors": surv, "pair_ms": res,
                  "median": {v: float(np.median(res[v])) for v in vals}}, indent=1))

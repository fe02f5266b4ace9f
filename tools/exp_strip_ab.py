#!/usr/bin/env python
"""Alternating A/B of strip widths for the screen kernel (thermal / power drift cancels)."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bench
import weightedld_b200 as wld
from weightedld_b200 import _lib as L

wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
strips = [int(x) for x in (sys.argv[2:] or ["4", "8"])]
chars = torch.from_numpy(bench.make_input(wl)).cuda()
ctxs = {}
for s in strips:
    os.environ["WLD_STRIP"] = str(s)
    c = wld.Context(0)
    c.set_screen("always")
    c.load_alignment(chars)
    c.filter_sites(*bench.FILTER)
    c.henikoff()
    c.ld_pairs(bench.R2_THRESHOLD)   # plans with this strip width; cached in the context
    ctxs[s] = c
res = {s: [] for s in strips}
for rep in range(8):
    for s in strips:
        ctxs[s].ld_pairs(bench.R2_THRESHOLD)
        res[s].append(round(ctxs[s].stage_ms(L.STAGE_PAIR), 3))
print(json.dumps({"workload": wl, "pair_ms": {str(s): res[s] for s in strips}, "median": {str(s): float(np.median(res[s])) for s in strips}}, indent=1))

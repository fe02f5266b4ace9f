#!/usr/bin/env python
"""Timing experiment: the screen / exact pair kernel with and without its epilogue (WLD_EXPERIMENT_SKIP_EPILOGUE=1 makes the
epilogue warps release every accumulator untouched; results are not computed) — how much of the kernel time is MMA alone."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bench
import weightedld_b200 as wld
from weightedld_b200 import _lib as L

out = {}
for wl in sys.argv[1:] or ["c3", "c5"]:
    chars = torch.from_numpy(bench.make_input(wl)).cuda()
    for screen in ("always", "never"):
        for skip in ("0", "1"):
            os.environ["WLD_EXPERIMENT_SKIP_EPILOGUE"] = skip
            with wld.Context(0) as ctx:
                ctx.set_screen(screen)
                ctx.load_alignment(chars)
                ctx.filter_sites(*bench.FILTER)
                ctx.henikoff()
                ms = []
                for _ in range(4):
                    try:
                        ctx.ld_pairs(bench.R2_THRESHOLD)
                    except wld.WldError as e:   # the pair count check fails without an epilogue: time is still recorded
                        pass
                    ms.append(ctx.stage_ms(L.STAGE_PAIR))
                out[f"{wl}_{screen}_skip{skip}"] = round(float(np.median(ms[1:])), 4)
print(json.dumps(out, indent=1))

#!/bin/bash
# cell-level path: tests, then timing on a block-LD workload of config-5 size (auto vs never)
O=gpurun_out; T=${1:-r02O}
timeout 500 python -m pytest tests/test_screen_refine.py -m gpu -q -x > $O/${T}_tests.log 2>&1; echo rc=$? >> $O/${T}_tests.log
python - > $O/${T}_blockld.json 2> $O/${T}_blockld.err <<'PY'
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench
import weightedld_b200 as wld
from weightedld_b200 import _lib as L
from weightedld_b200.synth import make_alignment
out = {}
for name, kw in (("block_ld_10000x50000_founders8_block400", dict(founders=8, block=400)),):
    chars = torch.from_numpy(make_alignment(10000, 50000, seed=77, **kw)).cuda()
    ref = None
    for screen in ("never", "auto"):
        with wld.Context(0) as ctx:
            ctx.set_screen(screen)
            ctx.load_alignment(chars)
            ctx.filter_sites(*bench.FILTER)
            ctx.henikoff()
            ms = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(torch.cuda.current_stream())
                n, done = ctx.ld_pairs(0.1)
                torch.cuda.synchronize()
                import time
                ms.append({k: round(ctx.stage_ms(L.STAGE_NAMES.index(k)), 3) for k in ("pair_prep", "pair_sample", "pair", "pair_refine")})
            info = ctx.pair_info()
            pairs = ctx.fetch_pairs(n)
            if ref is None: ref = pairs.tobytes()
            out[f"{name}_{screen}"] = {"stages_ms_last": ms[-1], "survivors": int(n), "pairs": int(done), "screen": info.screen,
                                       "tiles": int(info.tiles), "cells": int(info.screen_cells), "cells_flagged": int(info.screen_cells_flagged),
                                       "sample": [int(info.sample_candidates), int(info.sample_pairs), int(info.sample_tiles_flagged), int(info.sample_tiles)],
                                       "identical_to_exact": pairs.tobytes() == ref}
print(json.dumps(out, indent=1))
PY

#!/usr/bin/env python
"""Sweeps the pair stage's L2 knobs on one workload in ONE process (the 500 MB synthetic input is
generated once): rasterisation strip width (WLD_STRIP) x TMA L2 eviction hints for the indicator (A)
and limb (B) panels (WLD_HINT_A / WLD_HINT_B = normal|first|last) x die-aware schedule (WLD_DIE = 0|1|2).
(K-loop rotation and the TMA L2-promotion size were swept with earlier versions of this tool and removed:
profiles/r01_l2_sweep_c5_die_schedule*.)

    python tools/l2_sweep.py --workload c5 --steps 5 --warmup 3 --out gpurun_out/l2_sweep.json
    ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:pair_umma --csv --log-file x.csv \
        python tools/l2_sweep.py --workload c5 --steps 1 --warmup 0       # DRAM bytes per configuration

Numbers printed here are tuning evidence, not bench values (bench.py is the measurement)."""
import argparse
import itertools
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--strips", default="8,16,24")
    ap.add_argument("--hints", default="nn,fn,nl,fl")
    ap.add_argument("--dies", default="1")
    ap.add_argument("--wavesync", default="", help="comma list of 0/1 (WLD_WAVESYNC); empty = library default")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch

    import bench
    import weightedld_b200 as wld

    chars = bench.make_input(args.workload)
    dev = torch.from_numpy(chars).cuda()
    names = {"n": "normal", "f": "first", "l": "last"}
    rows = []
    for strip, hint, die, ws in itertools.product([int(x) for x in args.strips.split(",")], args.hints.split(","),
                                                  args.dies.split(","), args.wavesync.split(",")):
        os.environ["WLD_DIE"] = die
        if ws:
            os.environ["WLD_WAVESYNC"] = ws
        os.environ["WLD_STRIP"] = str(strip)
        os.environ["WLD_HINT_A"] = names[hint[0]]
        os.environ["WLD_HINT_B"] = names[hint[1]]
        with wld.Context(0) as ctx:
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            ms = []
            with bench.ClockSampler(0) as clk:
                for it in range(args.warmup + args.steps):
                    ctx.load_alignment(dev)
                    ctx.filter_sites(*bench.FILTER)
                    ctx.henikoff()
                    n, done = ctx.ld_pairs(bench.R2_THRESHOLD)
                    if it >= args.warmup:
                        ms.append(ctx.stage_ms(wld.STAGE_PAIR))
            row = {"wavesync": ws, "die": int(die), "strip": strip, "hint_a": names[hint[0]], "hint_b": names[hint[1]], "pair_ms": float(np.mean(ms)),
                   "pair_ms_min": float(np.min(ms)), "survivors": n, "pairs": done, "die_schedule": ctx.pair_info().die_schedule,
                   "die_sms": list(ctx.pair_info().die_sms), "clocks": clk.summary()}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()

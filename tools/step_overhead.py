#!/usr/bin/env python
"""Host-side wall time of every ABI call of one step (perf_counter around the blocking calls) next to the
device time of its stage (library CUDA events): shows where a step spends time OUTSIDE its kernels.

    python tools/step_overhead.py --workload c5 [--cap N] [--steps 5]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--cap", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch

    import bench
    import weightedld_b200 as wld

    chars = bench.make_input(args.workload)
    dev = torch.from_numpy(chars).cuda()
    ctx = wld.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if args.cap:
        ctx.set_pair_capacity(args.cap)
    rows = []
    for it in range(args.steps + 2):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        ctx.load_alignment(dev); t.append(time.perf_counter())
        ctx.filter_sites(*bench.FILTER); t.append(time.perf_counter())
        ctx.henikoff(); t.append(time.perf_counter())
        n, done = ctx.ld_pairs(bench.R2_THRESHOLD); t.append(time.perf_counter())
        wall = [(b - a) * 1e3 for a, b in zip(t, t[1:])]
        devms = [ctx.stage_ms(i) for i in range(6)]
        if it >= 2:
            rows.append({"wall_ms": dict(zip(["load", "filter", "henikoff", "ld_pairs"], wall)),
                         "device_ms": dict(zip(wld.STAGE_NAMES, devms)), "total_wall_ms": sum(wall)})
    avg = lambda f: sum(f(r) for r in rows) / len(rows)
    out = {"workload": args.workload, "cap": args.cap,
           "wall_ms": {k: avg(lambda r: r["wall_ms"][k]) for k in rows[0]["wall_ms"]},
           "device_ms": {k: avg(lambda r: r["device_ms"][k]) for k in rows[0]["device_ms"]},
           "total_wall_ms": avg(lambda r: r["total_wall_ms"])}
    out["outside_kernels_ms"] = out["total_wall_ms"] - sum(out["device_ms"].values())
    print(json.dumps(out))


if __name__ == "__main__":
    main()

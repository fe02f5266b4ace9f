#!/usr/bin/env python
"""Host-side time of every call of one multi-GPU step (torchrun, one process per GPU): where the gaps between the
stage kernels come from.  python -m torch.distributed.run --nproc-per-node N tools/mgpu_trace.py --workload c5"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    import weightedld_b200 as wld
    from weightedld_b200._lib import EXCHANGE_HISTOGRAM, EXCHANGE_WEIGHT_SUMS
    from weightedld_b200.multi_gpu import shard_rows

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chars = bench.make_input(args.workload)
    dev = torch.from_numpy(chars).cuda()
    ctx = wld.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_partition(rank, world)
    lo, hi, _ = shard_rows(chars.shape[0], rank, world)
    acc = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        r = fn()
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
        return r

    def step(sync_each):
        def s(name, fn):
            r = timed(name, fn)
            if sync_each:
                timed(name + ":sync", torch.cuda.synchronize)
            return r
        s("set_row_shard", lambda: ctx.set_row_shard(lo, hi))
        s("load", lambda: ctx.load_alignment(dev))
        t = s("exchange_tensor(hist)", lambda: ctx.exchange_tensor(EXCHANGE_HISTOGRAM))
        s("all_reduce(hist)", lambda: dist.all_reduce(t))
        s("filter", lambda: ctx.filter_sites(*bench.FILTER))
        s("set_seq_shard", lambda: ctx.set_seq_shard(lo, hi))
        s("henikoff", ctx.henikoff)
        t = s("exchange_tensor(w)", lambda: ctx.exchange_tensor(EXCHANGE_WEIGHT_SUMS))
        s("all_reduce(w)", lambda: dist.all_reduce(t))
        s("henikoff_finish", ctx.henikoff_finish)
        s("ld_pairs", lambda: ctx.ld_pairs(bench.R2_THRESHOLD))

    out = {}
    for mode in (False, True):
        for _ in range(3):
            step(mode)
        acc.clear()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(mode)
        e1.record()
        torch.cuda.synchronize()
        out["sync_each" if mode else "async"] = {"step_ms": e0.elapsed_time(e1) / args.steps,
                                                  "host_ms": {k: round(v / args.steps * 1e3, 4) for k, v in acc.items()},
                                                  "stages_ms": {n: round(ctx.stage_ms(i), 4) for i, n in enumerate(wld.STAGE_NAMES)}}
    if rank == 0:
        print(json.dumps({"workload": args.workload, "world": world, **out}, indent=1))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

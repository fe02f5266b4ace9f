#!/usr/bin/env python
"""Experiment: does the ORDER of the sequences (the K dimension of the Gram) change the pair kernel's speed
under the power cap?  The sums are exact, so any permutation gives identical results; sorting similar sequences
next to each other makes the indicator / limb operand rows piecewise constant along K (less switching activity
in the tensor-core datapath).  Prints pair-kernel ms and SM clock for the original and the sorted order."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--key-sites", type=int, default=48)
    args = ap.parse_args()
    import torch

    import bench
    import weightedld_b200 as wld

    chars = bench.make_input(args.workload)
    n, l = chars.shape
    # sort key: the sequence's bytes at `key_sites` columns spread over the alignment (lexicographic)
    cols = np.linspace(0, l - 1, args.key_sites).astype(np.int64)
    order = np.lexsort(chars[:, cols[::-1]].T)
    variants = {"original": chars, "sorted": np.ascontiguousarray(chars[order]),
                "shuffled": np.ascontiguousarray(chars[np.random.default_rng(0).permutation(n)])}
    ref = None
    for name, arr in variants.items():
        dev = torch.from_numpy(arr).cuda()
        with wld.Context(0) as ctx:
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            ms = []
            with bench.ClockSampler(0) as clk:
                for it in range(3 + args.steps):
                    ctx.load_alignment(dev)
                    ctx.filter_sites(*bench.FILTER)
                    ctx.henikoff()
                    ns, done = ctx.ld_pairs(bench.R2_THRESHOLD)
                    if it >= 3:
                        ms.append(ctx.stage_ms(wld.STAGE_PAIR))
            pairs = ctx.fetch_pairs(ns)
        if ref is None:
            ref = pairs.tobytes()
        print(json.dumps({"order": name, "pair_ms": float(np.mean(ms)), "pair_ms_min": float(np.min(ms)), "survivors": int(ns),
                          "identical_output": pairs.tobytes() == ref, "clocks": clk.summary()}), flush=True)
        del dev


if __name__ == "__main__":
    main()

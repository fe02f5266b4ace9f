#!/usr/bin/env python
"""Runs the drop-in CLI (weightedld_b200/weighted_ld) on a bench.py workload written as a FASTA file and
reports where the wall time goes (its own main.rs-style log lines + total), i.e. the host side of
SURVEY.md §8f rows 1-2 (FASTA ingest, TSV writer).

    python tools/cli_scale.py --workload c5 [--gpus 1] [--keep]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--quiet", action="store_true", help="RUST_LOG=info: no progress lines")
    args = ap.parse_args()
    import bench

    t0 = time.perf_counter()
    chars = bench.make_input(args.workload)
    n, l = chars.shape
    d = Path(tempfile.mkdtemp(prefix="wld_cli_", dir=args.dir if os.path.isdir(args.dir) else None))
    fasta = d / "in.fasta"
    # one '>' line + one sequence line per record (the only layout lib.rs:277-307 reads correctly)
    name_w = 12
    rec = np.empty((n, name_w + l + 1), np.uint8)
    names = np.array([list(f">seq{i:07d}\n".encode()) for i in range(n)], np.uint8)
    rec[:, :name_w] = names
    rec[:, name_w:name_w + l] = chars
    rec[:, -1] = 10
    rec.tofile(fasta)
    del rec
    gen_s = time.perf_counter() - t0
    out = {"host_threads": os.cpu_count(), "workload": args.workload, "n_seqs": n, "n_cols": l, "fasta_bytes": fasta.stat().st_size, "generate_s": gen_s, "runs": []}
    for r in range(args.repeat):
        t0 = time.perf_counter()
        p = subprocess.run([str(ROOT / "weightedld_b200" / "weighted_ld"), "--fasta-input", str(fasta), "--pair-output",
                            str(d / "pairs.tsv"), "--weights-output", str(d / "w.tsv"), "--gpus", str(args.gpus)],
                           env=dict(os.environ, RUST_LOG="info" if args.quiet else "debug", WLD_DEBUG="1", WLD_CLI_TIMING="1"),
                           capture_output=True, text=True)
        wall = time.perf_counter() - t0
        lines = [ln.split("] ", 1)[-1] for ln in p.stderr.splitlines() if "progress " not in ln]
        out["runs"].append({"rc": p.returncode, "wall_s": wall, "log": lines,
                            "pairs_tsv_bytes": (d / "pairs.tsv").stat().st_size if (d / "pairs.tsv").exists() else 0})
    print(json.dumps(out, indent=1))
    for f in d.iterdir():
        f.unlink()
    d.rmdir()


if __name__ == "__main__":
    main()

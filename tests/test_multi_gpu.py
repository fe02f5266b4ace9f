"""N > 1 on real GPUs (skipped with fewer than 2): sharded input broadcast over NCCL, tile-partitioned pair
stage, host merge — the merged output must be byte-identical to the single-GPU run."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpus_equal_one():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "_mgpu_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == world and d["pairs"] == d["expected"] == d["done1"]
    assert d["identical"] and d["survivors"] > 1000

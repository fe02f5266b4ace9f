"""Worker of tests/test_multi_gpu.py: run under torchrun, one process per GPU."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import weightedld_b200 as wld  # noqa: E402
from weightedld_b200.multi_gpu import ShardedLoader, gather_pairs, merge_on_device, sharded_stages  # noqa: E402
from weightedld_b200.synth import make_alignment  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chars = make_alignment(700, 5000, seed=41, block=120, clonal=True)
    host = torch.from_numpy(chars).pin_memory()
    loader = ShardedLoader(chars.shape[0], chars.shape[1], rank, world, torch.device("cuda", local))
    full = loader.load(host)                       # own rows over PCIe + all-gather over NVLink
    assert torch.equal(full.cpu(), torch.from_numpy(chars))
    # the same in three pieces (H2D of piece c+1 overlaps the all-gather of piece c; the last piece is short), twice
    piped = ShardedLoader(chars.shape[0], chars.shape[1], rank, world, torch.device("cuda", local), chunks=3)
    assert piped.chunks == 3
    for _ in range(2):
        assert torch.equal(piped.load(host).cpu(), torch.from_numpy(chars))
    # pageable host memory goes through the loader's pinned staging buffer (forced here: the input is small)
    big = np.tile(chars, (8, 2))                      # 5 600 x 10 000 = 56 MB
    pl = ShardedLoader(big.shape[0], big.shape[1], rank, world, torch.device("cuda", local))
    for _ in range(2):
        assert torch.equal(pl.load(torch.from_numpy(big)).cpu(), torch.from_numpy(big))
    with wld.Context(local) as ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.set_partition(rank, world)
        n_kept = sharded_stages(ctx, full, (0.8, 0.02, 0.5), rank, world)  # row / sequence shards + two exchanges
        w_sharded = ctx.weights_f64()
        n, done = ctx.ld_pairs(0.1)
        shard = ctx.fetch_pairs(n, wld.FETCH_KEPT_INDEX | wld.FETCH_UNORDERED)
        site_map = ctx.site_map()
        merged_dev = merge_on_device(ctx, n, rank, world)   # shards over NVLink, ordered on rank 0's GPU
    t = torch.tensor([done], device="cuda", dtype=torch.int64)
    dist.all_reduce(t)
    merged = gather_pairs(shard, n_kept, site_map, rank, world)  # host merge of the same shards
    if rank == 0:
        with wld.Context(local) as ctx:            # the whole triangle on one GPU
            ctx.load_alignment(chars)
            assert ctx.filter_sites() == n_kept
            ctx.henikoff()
            w_whole = ctx.weights_f64()
            n1, done1 = ctx.ld_pairs(0.1)
            whole = ctx.fetch_pairs(n1)
        print(json.dumps({"world": world, "pairs": int(t.item()), "expected": n_kept * (n_kept - 1) // 2, "done1": done1,
                          "survivors": len(merged),
                          "identical": merged.tobytes() == whole.tobytes() and merged_dev.tobytes() == whole.tobytes()
                          and w_sharded.tobytes() == w_whole.tobytes()}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU tests of the multi-GPU host logic: the tile partition is a disjoint,
balanced cover of the upper triangle, and shards merge into the reference's output order.
The N>1 path runs as two real processes over torch.distributed (gloo), as bench.py does over NCCL:
each rank plans its own part, 'computes' its shard (the oracle stands in for the GPU kernel,
restricted to the rank's tiles), rank 0 gathers and merges, and the max-over-ranks reduction used
for timing is exercised."""
import os
import socket

import numpy as np
import pytest


def tile_owner_mask(wld, n_kept, a, b, tiles, tile_m, tile_n):
    """Which pairs (a < b, kept indices) fall into this partition's tiles: inside a tile's rows AND its window."""
    own = np.zeros(len(a), bool)
    for tm, tn, j_lo, j_hi in tiles.astype(np.int64):
        own |= (a // tile_m == tm) & (b // tile_n == tn) & (b >= j_lo) & (b < j_hi)
    return own


def pair_cover(tiles, n_kept, tile_m, tile_n):
    """(n_kept, n_kept) count of how many tiles evaluate pair (a, b), a < b."""
    seen = np.zeros((n_kept, n_kept), np.int32)
    for tm, tn, j_lo, j_hi in tiles.astype(np.int64):
        assert tn * tile_n <= j_lo < j_hi <= min(n_kept, (tn + 1) * tile_n)
        i0, i1 = tm * tile_m, min(n_kept, (tm + 1) * tile_m)
        seen[i0:i1, j_lo:j_hi] += 1
    return np.triu(seen, 1)


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("n_limbs,tile_n", [(1, 128), (2, 64), (3, 42), (4, 32)])
@pytest.mark.parametrize("n_kept", [1, 2, 63, 64, 65, 129, 700, 1500])
def test_plan_covers_each_pair_once(n_kept, n_limbs, tile_n, ctas):
    import weightedld_b200 as wld
    tile_m = 64 * ctas
    want = np.triu(np.ones((n_kept, n_kept), np.int32), 1)
    for nparts in (1, 2, 3, 8):
        seen = np.zeros((n_kept, n_kept), np.int32)
        total = 0
        for part in range(nparts):
            tiles, pairs = wld.plan_tiles(n_kept, n_limbs, part, nparts, sm_count=4, cta_group=ctas)
            cover = pair_cover(tiles, n_kept, tile_m, tile_n)
            assert cover.sum() == pairs          # the pair count the kernel must report
            assert len(tiles) == 0 or all(cover[tm * tile_m:(tm + 1) * tile_m, lo:hi].any() for tm, _, lo, hi in tiles.astype(np.int64))
            seen += cover
            total += pairs
        assert total == n_kept * (n_kept - 1) // 2
        assert np.array_equal(seen, want)        # every pair a < b exactly once over the partitions


@pytest.mark.parametrize("n_kept,nparts", [(700, 2), (1500, 3), (5000, 8)])
def test_partitions_own_the_same_pairs_for_every_kernel_variant(n_kept, nparts):
    """A GPU may run the one-limb screen (128-site tiles) or the exact kernel (42- or 32-site tiles) on its own:
    the site pairs of partition p must not depend on that choice."""
    import weightedld_b200 as wld
    rng = np.random.default_rng(n_kept)
    a = rng.integers(0, n_kept - 1, 20000)
    b = rng.integers(0, n_kept, 20000)
    a, b = np.minimum(a, b), np.maximum(a, b)
    a, b = a[a < b], b[a < b]
    for part in range(nparts):
        masks = []
        for n_limbs, tile_n in ((1, 128), (2, 64), (3, 42), (4, 32)):
            for ctas in (1, 2):
                tiles, _ = wld.plan_tiles(n_kept, n_limbs, part, nparts, cta_group=ctas)
                masks.append(tile_owner_mask(wld, n_kept, a, b, tiles, 64 * ctas, tile_n))
        assert all(np.array_equal(masks[0], m) for m in masks[1:])


def test_plan_is_balanced_at_config5():
    import weightedld_b200 as wld
    for ctas in (1, 2):
        for nparts in (2, 4, 8):
            for n_limbs in (1, 3):
                plans = [wld.plan_tiles(48601, n_limbs, p, nparts, cta_group=ctas) for p in range(nparts)]
                pairs = [p[1] for p in plans]
                tiles = [len(p[0]) for p in plans]
                # equal ranges of the 128 x 128-site cell list: tensor work within one cell row of tiles, pairs within 2 %
                assert max(tiles) / min(tiles) < (1.001 if n_limbs == 1 else 1.02)  # exact-kernel tiles straddling a partition edge run in both
                assert max(pairs) / min(pairs) < 1.02                 # site pairs (diagonal tiles hold fewer)
                ranges = [(int(tl[:, 2].min()), int(tl[:, 3].max())) for tl, _ in plans]
                for (lo0, hi0), (lo1, hi1) in zip(ranges, ranges[1:]):  # contiguous runs: neighbours share at most one strip of 1024 sites
                    assert lo0 <= lo1 and hi0 <= hi1 and lo1 >= hi0 - 1024


def _worker(rank, world, port, n_kept, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import torch
    import torch.distributed as dist

    import weightedld_b200 as wld
    from oracle import oracle as O
    from weightedld_b200.synth import make_alignment

    dist.init_process_group("gloo", rank=rank, world_size=world)
    chars = make_alignment(120, n_kept + 40, seed=5, block=50)
    fs = O.filter_sites(O.siteset_from_chars(chars))
    w = O.quantize_weights(O.henikoff_weights(fs), 24)
    kept = O.SiteSet(fs.codes, fs.hists, None)  # kept indices, like WLD_FETCH_KEPT_INDEX
    full, computed = O.all_weighted_ld_pairs(kept, w, 0.1, O.F64)
    tiles, my_pairs = wld.plan_tiles(fs.n_sites, 3, rank, world, sm_count=2, cta_group=2)
    mine = full[tile_owner_mask(wld, fs.n_sites, full["a"], full["b"], tiles, 128, 42)]
    rng = np.random.default_rng(rank)
    shard = np.empty(len(mine), wld.PAIR_DTYPE)  # the GPU emits its survivors unordered
    for src, dst in (("a", "site_a"), ("b", "site_b"), ("d", "d"), ("d_prime", "d_prime"), ("r2", "r2")):
        shard[dst] = mine[src]
    shard = shard[rng.permutation(len(shard))]

    t = torch.tensor([my_pairs], dtype=torch.int64)
    dist.all_reduce(t)
    ms = torch.tensor([10.0 + rank])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, shard)
    if rank == 0:
        merged = wld.merge_shards(fs.n_sites, gathered, fs.site_map)
        ref, _ = O.all_weighted_ld_pairs(fs, w, 0.1, O.F64)  # parent indices, reference order
        ok = (len(merged) == len(ref) and np.array_equal(merged["site_a"], ref["a"])
              and np.array_equal(merged["site_b"], ref["b"]) and np.array_equal(merged["r2"], ref["r2"]))
        q.put((int(t.item()), computed, float(ms.item()), ok, len(ref), [len(g) for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_partition_and_merge_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 600, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, computed, ms, ok, n_ref, sizes = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == computed            # the two parts cover every pair exactly once
    assert ms == 11.0                   # max over ranks
    assert ok and n_ref > 100 and min(sizes) > 0


def _loader_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import torch
    import torch.distributed as dist

    from weightedld_b200.multi_gpu import ShardedLoader, gather_pairs, shard_rows
    import weightedld_b200 as wld

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for n_seqs, n_cols in ((7, 5), (64, 33), (101, 48)):
        chars = torch.from_numpy(np.random.default_rng(1).integers(0, 255, (n_seqs, n_cols), dtype=np.uint8))
        ld = ShardedLoader(n_seqs, n_cols, rank, world, torch.device("cpu"))
        lo, hi, per = shard_rows(n_seqs, rank, world)
        full = ld.load(chars)                 # every rank passes the whole matrix
        ok &= bool(torch.equal(full, chars)) and full.stride(0) % 16 == 0 and ld.h2d_bytes == (hi - lo) * n_cols
        full2 = ld.load(chars[lo:hi])         # or only its own rows
        ok &= bool(torch.equal(full2, chars))
    shard = np.zeros(3, wld.PAIR_DTYPE)
    shard["site_a"] = [rank, rank, 300 + rank]
    shard["site_b"] = [rank + 5, rank + 300, 400 + rank]
    merged = gather_pairs(shard, 600, None, rank, world)
    if rank == 0:
        key = wld.pair_order_key(600, merged["site_a"], merged["site_b"])
        ok &= len(merged) == 3 * world and bool(np.all(np.diff(key.astype(np.int64)) >= 0))
    else:
        ok &= merged is None
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


def test_sharded_input_broadcast_and_output_gather_gloo():
    """multi_gpu.ShardedLoader (rows sharded over the ranks' host links + all-gather) and gather_pairs,
    world size 2 over gloo."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_loader_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    assert q.get(timeout=240) == 1
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("n_limbs,tile_n", [(2, 64), (3, 42), (4, 32)])
def test_cell_plan_covers_exactly_the_flagged_cells(n_limbs, tile_n, ctas):
    """Screen + exact kernel on the flagged cells: the exact kernel's tiles cut out for a set of
    flagged screen tiles must evaluate every pair of those tiles once and no other pair — for whole triangles and for
    partitions."""
    import weightedld_b200 as wld
    n_kept, tile_m = 1500, 64 * ctas
    rng = np.random.default_rng(n_limbs * 10 + ctas)
    for nparts in (1, 3):
        total_flagged_pairs = 0
        seen_all = np.zeros((n_kept, n_kept), np.int32)
        want_all = np.zeros((n_kept, n_kept), np.int32)
        for part in range(nparts):
            screen, _ = wld.plan_tiles(n_kept, 1, part, nparts, cta_group=ctas)
            flags = (rng.random(len(screen)) < 0.3).astype(np.uint8)
            want = pair_cover(screen[flags == 1], n_kept, tile_m, 128)
            tiles, pairs = wld.plan_cell_tiles(n_kept, flags, n_limbs, part, nparts, cta_group=ctas)
            got = pair_cover(tiles, n_kept, tile_m, tile_n)
            assert np.array_equal(got, want) and got.max() <= 1 and got.sum() == pairs
            seen_all += got
            want_all += want
            total_flagged_pairs += pairs
        assert np.array_equal(seen_all, want_all) and seen_all.max() <= 1 and total_flagged_pairs == want_all.sum()
    # nothing flagged, everything flagged
    screen, pairs_all = wld.plan_tiles(n_kept, 1, cta_group=ctas)
    assert wld.plan_cell_tiles(n_kept, np.zeros(len(screen), np.uint8), n_limbs, cta_group=ctas)[1] == 0
    assert wld.plan_cell_tiles(n_kept, np.ones(len(screen), np.uint8), n_limbs, cta_group=ctas)[1] == pairs_all == n_kept * (n_kept - 1) // 2

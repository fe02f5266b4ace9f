"""One-process-per-GPU plumbing around the C ABI: `torch.distributed` is used for
device memory, the NVLink all-gather of the input and the host gather of the outputs — never for the
pair stage itself, which needs no collective (independent tiles, lib.rs:593-594).

    broadcast once   the alignment is sequence-major, so rank r copies only rows [r*R, (r+1)*R) over its
                     own PCIe link and one all-gather over NVLink/NVSwitch rebuilds the full matrix on
                     every GPU: host->device traffic per rank drops from N*L to N*L/world bytes.
    compute          every rank runs stages 1-2 on the full matrix (< 3 % of a step) and its own part of
                     the upper-triangular tile grid (wld_set_partition).
    merge            per-rank survivor shards are gathered on rank 0 and merged into the reference's
                     output order (api.merge_shards).
"""
from __future__ import annotations

import numpy as np


def shard_rows(n_seqs: int, rank: int, world: int) -> tuple[int, int, int]:
    """(first row, one-past-last row, padded rows per rank) of this rank's slice of the alignment."""
    per = -(-n_seqs // world)
    lo = min(rank * per, n_seqs)
    return lo, min(lo + per, n_seqs), per


class ShardedLoader:
    """Reusable buffers for the sharded host->device copy + all-gather of one alignment shape.

    The rank's rows travel in `chunks` pieces: while NCCL all-gathers piece c over NVLink, piece c+1 is already
    on the PCIe link (a second stream), so the distribution costs about max(H2D, all-gather) instead of their
    sum.  A piece of every rank lands in a small staging buffer and is copied to its rows of the full matrix
    (the matrix keeps the callers' sequence order)."""

    def __init__(self, n_seqs: int, n_cols: int, rank: int, world: int, device, group=None, chunks: int | None = None):
        import torch

        self.n_seqs, self.n_cols, self.rank, self.world, self.group = n_seqs, n_cols, rank, world, group
        self.lo, self.hi, self.per = shard_rows(n_seqs, rank, world)
        self.pitch = -(-max(n_cols, 1) // 16) * 16  # 16-byte pitch: vectorised histogram path
        self.full = torch.empty((self.per * world, self.pitch), dtype=torch.uint8, device=device)
        self.h2d_bytes = (self.hi - self.lo) * n_cols
        if chunks is None:  # pieces of >= 8 MB per rank, at most 4
            chunks = max(1, min(4, (self.per * self.pitch) // (8 << 20)))
        if world == 1 or torch.device(device).type != "cuda":
            chunks = 1
        self.chunks = chunks
        self.cs = -(-self.per // chunks)  # rows per piece
        if world > 1 and chunks > 1:
            self.copy_stream = torch.cuda.Stream(device=device)
            self.mine = torch.empty((self.per, self.pitch), dtype=torch.uint8, device=device)
            self.piece = torch.empty((world, self.cs, self.pitch), dtype=torch.uint8, device=device)
            self.events = [torch.cuda.Event() for _ in range(chunks)]

    def _pinned_stage(self):
        """(pinned (per, n_cols) uint8 tensor, fill(src_np, first_row, rows)) — allocated on first use."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        import torch

        if getattr(self, "_stage", None) is None:
            self._stage = torch.empty((self.per, self.n_cols), dtype=torch.uint8).pin_memory()
            self._stage_np = self._stage.numpy()
            self._pool = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 4) // 2)))

        def fill(src_np, a, rows):
            import numpy as np

            n = self._pool._max_workers
            cuts = [a + rows * k // n for k in range(n + 1)]
            # np.copyto releases the GIL: the slices are copied in parallel
            list(self._pool.map(lambda k: np.copyto(self._stage_np[cuts[k]:cuts[k + 1]], src_np[cuts[k]:cuts[k + 1]]), range(n)))

        return self._stage, fill

    def load(self, host_rows):
        """host_rows: (n_seqs, n_cols) uint8 torch tensor in (ideally pinned) host memory, identical on
        every rank — or just this rank's rows [lo, hi).  Returns the full (n_seqs, n_cols) device view."""
        import torch
        import torch.distributed as dist

        src = host_rows if host_rows.shape[0] == self.hi - self.lo else host_rows[self.lo:self.hi]
        n_own = self.hi - self.lo
        stage_rows = (lambda a, rows: None)
        if self.full.is_cuda and not src.is_pinned() and n_own * self.n_cols >= (16 << 20):
            # pageable host memory (what a Rust Vec<u8> caller holds): the rows pass through a pinned buffer, filled by
            # a few host threads piece by piece, so that the copy engine sees pinned memory and piece c+1 is being
            # staged while piece c is on the bus
            pinned, fill = self._pinned_stage()
            torch.cuda.synchronize(self.full.device)  # copies of the previous load out of the pinned buffer are done
            src_np = src.numpy()
            stage_rows = lambda a, rows: fill(src_np, a, rows)
            src = pinned
        if self.world == 1 or self.chunks == 1:
            stage_rows(0, n_own)
            mine = self.full[self.rank * self.per: self.rank * self.per + n_own, : self.n_cols]
            mine.copy_(src, non_blocking=True)
            if self.world > 1:
                dist.all_gather_into_tensor(self.full, self.full[self.rank * self.per: (self.rank + 1) * self.per],
                                            group=self.group)
            return self.full[: self.n_seqs, : self.n_cols]
        cur = torch.cuda.current_stream()
        self.copy_stream.wait_stream(cur)  # earlier readers of the buffers are done
        with torch.cuda.stream(self.copy_stream):
            for c in range(self.chunks):
                a = c * self.cs
                rows = max(0, min(self.cs, n_own - a))
                if rows:
                    stage_rows(a, rows)
                    self.mine[a:a + rows, : self.n_cols].copy_(src[a:a + rows], non_blocking=True)
                self.events[c].record(self.copy_stream)
        full3 = self.full.view(self.world, self.per, self.pitch)
        for c in range(self.chunks):
            a, b = c * self.cs, min((c + 1) * self.cs, self.per)
            if b <= a:
                break
            cur.wait_event(self.events[c])
            piece = self.piece[:, : b - a]
            if b - a == self.cs:
                dist.all_gather_into_tensor(piece, self.mine[a:b], group=self.group)
                full3[:, a:b].copy_(piece)
            else:  # short last piece: gather into a contiguous scratch of its own size
                scratch = self.piece.view(-1)[: self.world * (b - a) * self.pitch].view(self.world, b - a, self.pitch)
                dist.all_gather_into_tensor(scratch, self.mine[a:b], group=self.group)
                full3[:, a:b].copy_(scratch)
        return self.full[: self.n_seqs, : self.n_cols]


def sharded_stages(ctx, src, filt, rank: int, world: int, group=None, weights=None) -> int:
    """Stages 1-2 with the work split over the ranks (include/wld.h, "stages 1-2 on several GPUs"): rank r
    histograms its rows and sums the Henikoff contributions of its sequences; one all-reduce of the integer
    histogram (<= 720 KB at config 4) and one of the weight sums (f64, every entry has exactly one non-zero
    contributor, so the sum is exact) make every rank whole.  Returns n_kept.  Results are bit-identical to
    the replicated path for any world size."""
    import torch.distributed as dist

    from ._lib import EXCHANGE_HISTOGRAM, EXCHANGE_WEIGHT_SUMS

    n_seqs = src.shape[0]
    lo, hi, _ = shard_rows(n_seqs, rank, world)
    ctx.set_row_shard(lo, hi)
    ctx.load_alignment(src)
    if world > 1:
        dist.all_reduce(ctx.exchange_tensor(EXCHANGE_HISTOGRAM), group=group)
    n_kept = ctx.filter_sites(*filt)
    if weights is not None:
        ctx.set_weights(weights)
        return n_kept
    ctx.set_seq_shard(lo, hi)
    ctx.henikoff()
    if world > 1:
        dist.all_reduce(ctx.exchange_tensor(EXCHANGE_WEIGHT_SUMS), group=group)
        ctx.henikoff_finish()
    return n_kept


def gather_pairs(shard: np.ndarray, n_kept: int, site_map: np.ndarray | None, rank: int, world: int, group=None):
    """Gathers KEPT-index survivor shards (host arrays) on rank 0 and merges them on the host into the
    reference's output order; returns the merged array on rank 0 and None elsewhere.  (Host-only variant, used
    by the gloo tests; merge_on_device below is what the GPU path uses.)"""
    import torch.distributed as dist

    from .api import merge_shards

    if world == 1:
        return merge_shards(n_kept, [shard], site_map)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(shard, parts, dst=0, group=group)
    return merge_shards(n_kept, parts, site_map) if rank == 0 else None


def merge_on_device(ctx, n_survivors: int, rank: int, world: int, out: np.ndarray | None = None, group=None):
    """The host merge of the north star, done on rank 0's GPU: every rank hands its unordered KEPT-index shard
    to NCCL straight from device memory (no host round trip), rank 0 appends the foreign shards to its own
    survivors (wld_append_pairs) and one wld_fetch_pairs orders the union (lib.rs:623-679), maps it to raw
    columns (lib.rs:662-663) and copies it out.  Returns the merged PAIR_DTYPE array on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    from ._lib import FETCH_PARENT_INDEX, PAIR_DTYPE

    def buffer(n):  # `out`: None, a PAIR_DTYPE array, or a callable n -> array (e.g. a pinned-buffer pool)
        return out(n) if callable(out) else out

    if world == 1:
        return ctx.fetch_pairs(n_survivors, FETCH_PARENT_INDEX, out=buffer(n_survivors))
    dev = torch.device("cuda", torch.cuda.current_device())
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n_survivors], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = counts.tolist()
    item = PAIR_DTYPE.itemsize
    widest = max(counts[1:])
    if widest == 0:
        return ctx.fetch_pairs(counts[0], FETCH_PARENT_INDEX, out=buffer(counts[0])) if rank == 0 else None
    # ONE gather to rank 0: every shard padded to the widest one (rank 0 contributes an empty slot)
    send = torch.empty(widest * item, dtype=torch.uint8, device=dev)
    if rank != 0 and n_survivors:
        send[: n_survivors * item].copy_(ctx.fetch_pairs_device(n_survivors))
    slots = [torch.empty(widest * item, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(send, slots, dst=0, group=group)
    if rank != 0:
        return None
    foreign = torch.cat([slots[r][: counts[r] * item] for r in range(1, world) if counts[r]])
    torch.cuda.current_stream().synchronize()
    ctx.append_pairs(foreign)  # one device-to-device append, then one ordering pass over the union
    return ctx.fetch_pairs(sum(counts), FETCH_PARENT_INDEX, out=buffer(sum(counts)))

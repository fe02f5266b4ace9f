#!/usr/bin/env python
"""bench.py — weighted-LD site-pairs/sec of the WeightedLD hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c4|c3|tiny] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the whole hot path over one synthetic alignment: encode + histogram +
site filter -> Henikoff weights -> operand expansion -> tcgen05 Gram + fused epilogue + compaction.
`value` times it with the alignment already resident in HBM; `e2e` runs the same through the
C ABI from pinned HOST memory (H2D of the alignment and D2H of the surviving pairs inside the timed
region; with N > 1 ranks each rank copies only its 1/N of the rows over its own PCIe link and one NCCL
all-gather over NVLink rebuilds the matrix on every GPU — weightedld_b200/multi_gpu.py).  With N ranks the upper-triangular tile grid is partitioned (no collective on the data
path); the alignment is replicated; value = all pairs / max-over-ranks time ("strong" scaling: the
total work is fixed).  Inputs (0.5-3 GB of text, 8 GB of operands) are far larger than the 126 MB
L2, so no explicit flush is needed between iterations.

--impl reference times the reference's own CPU implementation of the path.  The Rust binary cannot
be built in this image (no cargo/rustc), so this is the C restatement of the
reference's fastest path (8-lane f32 `simd` build, 256x256 tiles, all host threads; oracle/),
on a bounded sample of tiles of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (description, n_seqs, n_cols, generator kwargs)
    "c5": ("synthetic 10,000 sequences x 50,000 variable sites (BASELINE.json configs[4])", 10_000, 50_000, {}),
    "c4": ("synthetic SARS-CoV-2-like 100,000 sequences x 30 kb, ~1/3 variable (configs[3])", 100_000, 30_000, {"sars": True}),
    "c3": ("synthetic 2,000 sequences x 20,000 variable sites (configs[2])", 2_000, 20_000, {}),
    "c3ld": ("synthetic 2,000 sequences x 20,000 sites, clonal / high-LD (about a third of all pairs survive)", 2_000, 20_000, {"clonal": True}),
    "tiny": ("synthetic 512 sequences x 3,000 sites (smoke)", 512, 3_000, {}),
    "c5blk": ("synthetic 10,000 sequences x 50,000 sites, strong LD inside blocks of 400 sites (8 founders): 0.34 % of the pairs survive, all near the diagonal", 10_000, 50_000, {"blocks": True}),
}
R2_THRESHOLD = 0.1
FILTER = (0.8, 0.02, 0.5)  # main.rs defaults


def make_input(name: str) -> np.ndarray:
    from weightedld_b200.synth import make_alignment, make_sarscov2_like
    _, n, l, kw = WORKLOADS[name]
    seed = 0xC0FFEE + list(WORKLOADS).index(name)
    if kw.get("sars"):
        return make_sarscov2_like(n, l, seed=seed)
    if kw.get("clonal"):
        return make_alignment(n, l, seed=seed, founders=256, block=400, clonal=True, stray=0.02, private_rate=0.002,
                              gap_rate=1e-3, n_rate=1e-3, third_rate=1e-3)
    if kw.get("blocks"):
        return make_alignment(n, l, seed=77, founders=8, block=400)
    return make_alignment(n, l, seed=seed)


def workload_config(desc: str, n_seqs: int, n_cols: int, n_kept: int) -> dict:
    """`config` of the JSON line — the same dictionary in our arm and in the reference arm."""
    return {"workload": desc, "n_seqs": n_seqs, "n_cols": n_cols, "n_kept": int(n_kept),
            "site_pairs": int(n_kept) * (int(n_kept) - 1) // 2, "r2_threshold": R2_THRESHOLD, "filter": list(FILTER),
            "l2": "inputs (0.5-3 GB of text, GBs of operands) far larger than the 126 MB L2: no flush needed between steps"}


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"),
                "hbm": d.get("hbm_gbs"), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


def int8_peak() -> dict:
    """INT8 tensor peak.  MEASURED_PEAKS.json holds no INT8 figure, and a library GEMM is no ceiling for
    tcgen05 kind::i8 (round 1 beat cuBLASLt's by 40 %), so the denominator is the instruction's own issue
    rate measured on this pool's B200 by tools/tc_peak.cu (operands resident in shared memory, back-to-back
    cta_group::2 M256 N256 K32 MMAs on all SMs; burst and 4 s sustained under the 1000 W cap), the higher of
    its two operand patterns.  Falls back to the nominal dense 4.5 POP/s ("of nominal") if the file is absent."""
    p = ROOT / "profiles" / "r02_tcgen05_peak.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"sustained": d["int8_tops_sustained"], "burst": d["int8_tops"],
                "source": "profiles/r02_tcgen05_peak.json (of measured: tcgen05.mma kind::i8 issue-rate ceiling, tools/tc_peak.cu)"}
    return {"sustained": 4500.0, "burst": 4500.0, "source": "nominal dense INT8 4.5 POP/s (of nominal)"}


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self._nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self) -> dict:
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port) — also the cpu_baseline leg of our own arm
# ------------------------------------------------------------------------------------------------
def cpu_reference(chars: np.ndarray, steps: int, warmup: int, target_s: float):
    """Times the C restatement of the reference's `simd` pair path on a bounded sample of tiles.
    Returns (pairs_per_s, info dict)."""
    from oracle import oracle as O
    O.build(native=True, force=True)  # -O3 -march=native on THIS machine's cores (README.md:88-97 of the reference)
    lib = O.lib(native=True)
    # every core this process may run on: torchrun exports OMP_NUM_THREADS=1, which must not cap the CPU arm
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib.wldo_set_threads(ncpu)
    threads = lib.wldo_max_threads()
    ss = O.filter_sites(O.siteset_from_chars(chars), *FILTER)
    w = O.henikoff_weights(ss)
    n_tiles = int(lib.wldo_tile_count(ss.n_sites))
    # calibrate on `threads` tiles, then size the sample for ~target_s per step
    rng = np.random.Generator(np.random.PCG64(1))
    start = int(rng.integers(0, max(1, n_tiles - threads)))
    t0 = time.perf_counter()
    _, done = O.all_weighted_ld_pairs(ss, w, R2_THRESHOLD, O.F32_SIMD8, tile_range=(start, start + threads), store=False, native=True)
    dt = time.perf_counter() - t0
    rate = done / dt
    sample_tiles = int(min(n_tiles, max(threads, target_s * rate / max(done / threads, 1))))
    sample_tiles = max(threads, sample_tiles // threads * threads)
    start = int(rng.integers(0, max(1, n_tiles - sample_tiles)))
    times, pairs = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, pairs = O.all_weighted_ld_pairs(ss, w, R2_THRESHOLD, O.F32_SIMD8, tile_range=(start, start + sample_tiles), store=False, native=True)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    info = {"cores": threads, "kind": "port",
            "sample": f"{sample_tiles} of {n_tiles} reference tiles (256x256 sites, {pairs} pairs) of the same workload, "
                      f"8-lane f32 restatement of lib.rs:410-453, -O3 -march=native, OpenMP dynamic; "
                      f"restated reference (C), not the Rust binary",
            "n_kept": ss.n_sites, "ms_per_step": dt * 1e3}
    return pairs / dt, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, n, l, _ = WORKLOADS[args.workload]
    chars = make_input(args.workload)
    value, info = cpu_reference(chars, args.steps, max(args.warmup, 1), target_s=6.0)
    line = {"impl": "reference", "metric": "weighted LD site-pairs/sec", "value": value, "unit": "site-pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(desc, n, l, info["n_kept"]),
            "parallelism": f"{info['cores']} host threads, one 256x256-site reference tile per task",
            "cpu_baseline": {"value": value, "unit": "site-pairs/s", "cores": info["cores"], "kind": info["kind"],
                             "sample": info["sample"]},
            "e2e": {"value": value, "unit": "site-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def check_mgpu_identity(wld, torch, dist, merge_on_device, local, rank, world):
    """Before any multi-GPU number is believed: on two small side workloads the merged N-rank output must be
    byte-identical to one GPU computing the whole triangle with the exact n-limb kernel (rank 0 runs that too).
    The first (clonal, high LD) goes through the exact kernel on every rank, the second through the one-limb
    screen + refinement on the even ranks and the exact kernel on the odd ones."""
    from weightedld_b200.multi_gpu import sharded_stages
    from weightedld_b200.synth import make_alignment
    ok = 1
    for chars, modes in ((make_alignment(900, 6000, seed=41, block=120, clonal=True), ("auto", "auto")),
                         (make_alignment(900, 6000, seed=42, block=120), ("always", "never"))):
        with wld.Context(local) as ctx:
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            ctx.set_partition(rank, world)
            ctx.set_screen(modes[rank % 2])
            sharded_stages(ctx, torch.from_numpy(chars).cuda(), FILTER, rank, world)
            n, _ = ctx.ld_pairs(R2_THRESHOLD)
            merged = merge_on_device(ctx, n, rank, world)
        if rank == 0:
            with wld.Context(local) as ctx:
                ctx.set_screen("never")
                ctx.load_alignment(chars)
                ctx.filter_sites(*FILTER)
                ctx.henikoff()
                n1, _ = ctx.ld_pairs(R2_THRESHOLD)
                whole = ctx.fetch_pairs(n1)
            ok &= int(len(whole) > 1000 and merged.tobytes() == whole.tobytes())
    t = torch.tensor([ok], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def run_ours(args):
    import torch
    import torch.distributed as dist

    import weightedld_b200 as wld

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: everything else that writes to file descriptor 1 while the job runs
    # (NCCL prints its version banner there from C) is sent to stderr, and the line goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    desc, n_seqs, n_cols, _ = WORKLOADS[args.workload]
    chars_np = make_input(args.workload)
    host = torch.from_numpy(chars_np).pin_memory()
    host_np = host.numpy()  # view of the pinned buffer
    dev = host.cuda(non_blocking=False)

    ctx = wld.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_partition(rank, world)
    if args.limbs:
        ctx.set_limbs(args.limbs)
    if args.kernel:
        ctx.set_pair_kernel(args.kernel)
    if args.ctas:
        ctx.set_cta_group(args.ctas)
    if args.screen:
        ctx.set_screen(args.screen)

    out_buf = {"t": None}

    def pinned_out(n):  # pinned host buffer for the survivors (grown on demand, reused across steps)
        if out_buf["t"] is None or out_buf["t"].numel() < 20 * n:
            out_buf["t"] = torch.empty(max(20 * n, 1 << 20), dtype=torch.uint8).pin_memory()
        return out_buf["t"].numpy()[: 20 * (out_buf["t"].numel() // 20)].view(wld.PAIR_DTYPE)

    stages = {k: 0.0 for k in wld.STAGE_NAMES}
    launches = {"n": 0}
    state = {}

    loader = None
    pageable_np = chars_np  # ordinary (pageable) host memory, as a Rust Vec<u8> caller holds it
    from weightedld_b200.multi_gpu import merge_on_device, sharded_stages
    if world > 1:
        from weightedld_b200.multi_gpu import ShardedLoader
        loader = ShardedLoader(n_seqs, n_cols, rank, world, torch.device("cuda", local))

    def step(src, fetch: bool, record: bool, pageable: bool = False):
        """One pass of the hot path.  fetch: the plugin call's result as the reference delivers it — ALL survivors
        of the job, in PairStore order with parent indices (lib.rs:623-679), in host memory; with N > 1 ranks the
        shards travel over NVLink to rank 0, which merges, orders and copies them out (multi_gpu.merge_on_device)."""
        if (src is host_np or src is pageable_np) and loader is not None:
            src = loader.load(host if src is host_np else torch.from_numpy(pageable_np))  # own rows H2D + all-gather
        # stages 1-2: split over the ranks with two small exact exchanges (multi_gpu.sharded_stages); one rank: plain
        n_kept = sharded_stages(ctx, src, FILTER, rank, world)
        n_surv, done = ctx.ld_pairs(R2_THRESHOLD)
        out = None
        if fetch:
            if world == 1:
                buf = np.empty(n_surv, wld.PAIR_DTYPE) if pageable else pinned_out(n_surv)
                out = ctx.fetch_pairs(n_surv, wld.FETCH_PARENT_INDEX, out=buf)
            else:
                out = merge_on_device(ctx, n_surv, rank, world, out=None if pageable else pinned_out)
        state.update(n_kept=n_kept, n_surv=n_surv, done=done)
        if fetch:
            state["order_ms"] = state.get("order_ms", 0.0) + ctx.stage_ms(wld.STAGE_ORDER)
            state["order_n"] = state.get("order_n", 0) + 1
        if record:
            for i, nm in enumerate(wld.STAGE_NAMES):
                stages[nm] += ctx.stage_ms(i)
                launches["n"] += ctx.stage_launches(i)
        return out

    def timed(src, fetch, steps, record, pageable=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(src, fetch, record, pageable)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        barrier()
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(n_warm):
        step(dev, False, False)
    with ClockSampler(local) as clocks:
        ms = timed(dev, False, args.steps, True)
    mgpu_identical = None
    if args.profile:
        ms_e2e = ms_e2e_pageable = float('nan')
    else:
        step(host_np, True, False)
        ms_e2e = timed(host_np, True, args.steps, False)
        step(pageable_np, True, False, True)
        ms_e2e_pageable = timed(pageable_np, True, max(1, args.steps // 2), False, True)
        if world > 1:
            mgpu_identical = check_mgpu_identity(wld, torch, dist, merge_on_device, local, rank, world)

    n_kept = state["n_kept"]
    total_pairs = n_kept * (n_kept - 1) // 2
    done_all = state["done"]
    surv_all = state["n_surv"]
    if world > 1:
        t = torch.tensor([done_all, surv_all], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        done_all, surv_all = int(t[0]), int(t[1])
    assert done_all == total_pairs, (done_all, total_pairs)

    info = ctx.pair_info()
    # every rank's own figures (rank 0 prints them): the step is as long as the slowest partition
    per_rank = None
    if world > 1:
        mine = torch.tensor([stages["pair"] / args.steps, sum(stages.values()) / args.steps, float(info.tiles),
                             float(info.screen_candidates), float(state["n_surv"])], device="cuda", dtype=torch.float64)
        allr = torch.zeros(world * mine.numel(), device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, -1).cpu().numpy()
        per_rank = {"pair_ms": [round(float(x), 4) for x in allr[:, 0]], "stages_sum_ms": [round(float(x), 4) for x in allr[:, 1]],
                    "tiles": [int(x) for x in allr[:, 2]], "candidates": [int(x) for x in allr[:, 3]],
                    "survivors": [int(x) for x in allr[:, 4]]}
    pk = peaks()
    pair_ms = stages["pair"] / args.steps
    # The dominant kernel: the exact n-limb Gram, or — screen + refine — the ONE-limb Gram that every pair goes through
    # (its candidates, a few in 10^5 here, are recomputed exactly by pair_refine_kernel: stages_ms.pair_refine).
    k_limbs = 1 if info.screen else info.n_limbs
    algo_flop = 8.0 * n_seqs * k_limbs * state["done"]           # 8*N flop per pair per limb pass
    useful_flop = 8.0 * n_seqs * state["done"]
    achieved = algo_flop / (pair_ms * 1e-3) / 1e12
    if info.kernel == 2:
        ip = int8_peak()
        peak, peak_src = ip["sustained"], ip["source"] + ", sustained int8 figure (kernel timed inside a long step)"
        kname = "pair_umma_kernel<1, i8, screen>" if info.screen else "pair_umma_kernel<NL, i8>"
    else:
        peak, peak_src = pk["bf16_sustained"], pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)"
        kname = "pair_umma_kernel<NL, bf16>" if info.kernel == 0 else "pair_simt_kernel"
    roof = {"bound": "tensor", "kernel": kname,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": peak_src,
            "achieved_useful": useful_flop / (pair_ms * 1e-3) / 1e12,
            "executed": info.executed_flop / (pair_ms * 1e-3) / 1e12,
            "algorithmic_flop_per_pair": 8 * n_seqs * k_limbs, "n_limbs": info.n_limbs, "kernel_limbs": k_limbs,
            "limb_bits": info.limb_bits,
            "frac_of_nominal_dense_int8": (achieved / 4500.0) if info.kernel == 2 else None,  # 4.5 POP/s data-sheet figure
            "kernel_ms": pair_ms, "traffic": None,
            "tile_schedule": {0: "round-robin", 1: "per-L2-die contiguous halves of the strip-rasterised tile list",
                              2: "per-L2-die, dealt per round"}.get(info.die_schedule, "?"),
            "die_sms": list(info.die_sms)}
    # DRAM bytes per launch of the pair kernel from an `ncu --set full` capture of THIS workload on THIS number
    # of GPUs (profiles/pair_umma_traffic.json: {workload: {n_gpus: bytes}}); null when no such capture exists.
    prof = ROOT / "profiles" / "pair_umma_traffic.json"
    if prof.exists():
        try:
            roof["traffic"] = json.loads(prof.read_text()).get(args.workload + ("_screen" if info.screen else ""), {}).get(str(world))
        except Exception:
            pass

    # HBM-bound stages: algorithmic bytes over the stage timers of the library (CUDA events; the filter
    # stage includes its small decision / scan kernels and one host round trip for the kept count).
    es = 1 if info.kernel == 2 else 2
    cells_raw, cells_kept = float(n_seqs) * n_cols, float(n_seqs) * n_kept
    hbm_bytes = {"histogram": cells_raw, "filter": cells_raw + cells_kept, "henikoff": cells_kept,
                 "pair_prep": 2 * cells_kept + cells_kept * es * (2 + 2 * k_limbs) / max(world, 1)}
    hbm = {k: {"bytes": b, "ms": stages[k] / args.steps, "gbs": b / (stages[k] / args.steps * 1e-3) / 1e9,
               "frac_of_measured_hbm": b / (stages[k] / args.steps * 1e-3) / 1e9 / pk["hbm"]}
           for k, b in hbm_bytes.items() if stages[k] > 0}
    if rank == 0:
        line = {
            "metric": "weighted LD site-pairs/sec", "value": total_pairs / (ms * 1e-3), "unit": "site-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {0: "bf16 limbs x fp32 accumulate (exact integers), f64 epilogue",
                      1: "f64 (SIMT verification kernel)",
                      2: "u8 limbs x s32 accumulate (exact integers), f64 epilogue"}[info.kernel],
            "data": "synthetic",
            "config": workload_config(desc, n_seqs, n_cols, n_kept),   # identical in both arms
            "survivors": surv_all, "parallelism": f"triangle-partition x{world}",
            "stages_ms": {k: v / args.steps for k, v in stages.items()},
            "roofline": roof,
            "screen": {"used": bool(info.screen), "mode": {0: "exact kernel over every pair", 1: "one-limb screen + per-pair refinement",
                                                          2: "one-limb screen + exact kernel on the flagged cells"}[int(info.screen)],
                       "cells": int(info.screen_cells), "cells_flagged": int(info.screen_cells_flagged),
                       "sample_tiles": int(info.sample_tiles), "sample_tiles_flagged": int(info.sample_tiles_flagged),
                       "candidates": int(info.screen_candidates), "top_min": int(info.screen_top_min),
                       "sample_pairs": int(info.sample_pairs), "sample_candidates": int(info.sample_candidates),
                       "reruns": int(info.screen_reruns),
                       "note": "one-limb Gram + rigorous r2 bound, candidates recomputed exactly (wld_set_screen); rank 0's figures; "
                               "survivors bit-identical to the exact n-limb kernel (tests/test_screen_refine.py)"},
            "hbm_stages": hbm,
            "e2e": {"value": total_pairs / (ms_e2e * 1e-3), "unit": "site-pairs/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(chars_np.nbytes),  # whole job: each rank copies 1/N of the rows
                    "d2h_bytes_per_step": int(20 * surv_all + 64 * world),
                    "host_buffers": "pinned",
                    "order_ms": state.get("order_ms", 0.0) / max(state.get("order_n", 1), 1),
                    "result": "all survivors in the reference's PairStore order with parent indices (wld_fetch_pairs default flags)"
                              + ("" if world == 1 else "; shards gathered over NVLink and merged on rank 0 inside the timed region"),
                    "pageable": {"value": total_pairs / (ms_e2e_pageable * 1e-3), "ms_per_step": ms_e2e_pageable,
                                 "host_buffers": "pageable input and output (what a Rust Vec<u8> / Vec<PairData> caller holds); "
                                                 "staged through pinned chunks by host threads inside libwld"},
                    "input_distribution": "host buffer -> one GPU" if world == 1 else
                    f"rows sharded over {world} PCIe links + NCCL all-gather over NVLink"},
            "mgpu_identical": mgpu_identical,
            "per_rank": per_rank,
            "gpu_launches": launches["n"],
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu:
            v, ci = cpu_reference(chars_np, 1, 1, target_s=12.0)
            line["cpu_baseline"] = {"value": v, "unit": "site-pairs/s", "cores": ci["cores"], "kind": ci["kind"], "sample": ci["sample"]}
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=list(WORKLOADS))
    ap.add_argument("--limbs", type=int, default=0)
    ap.add_argument("--kernel", default="", choices=["", "umma", "bf16", "i8", "simt"])
    ap.add_argument("--ctas", type=int, default=0, choices=[0, 1, 2], help="tcgen05 cta_group (0 = library default)")
    ap.add_argument("--screen", default="", choices=["", "auto", "never", "always"], help="wld_set_screen (default: library default = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true", help="profiling run: honour --warmup < 3, skip the e2e leg (numbers are not bench values)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

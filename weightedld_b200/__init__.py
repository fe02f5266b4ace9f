"""weightedld_b200 — B200-native (sm_100a) implementation of WeightedLD's hot path.

encode + site filter -> Henikoff weights -> all-pairs weighted LD (tcgen05 Gram + fused epilogue),
behind the C ABI of include/wld.h (libwld.so).  This package is the thin host-side mirror of the
reference's Rust API; it has no CPU or PyTorch fallback and raises if libwld.so is missing.
"""
from ._lib import (FETCH_DEVICE, FETCH_KEPT_INDEX, FETCH_PARENT_INDEX, FETCH_UNORDERED, PAIR_DTYPE, PAIR_KERNEL_SIMT,
                   PAIR_KERNEL_UMMA, PAIR_KERNEL_UMMA_I8, STAGE_FILTER, STAGE_HENIKOFF, STAGE_HISTOGRAM, STAGE_LOAD, STAGE_NAMES, STAGE_ORDER,
                   STAGE_PAIR, STAGE_PAIR_PREP, WldError)
from .api import (Context, MultiSequence, PairStore, SiteSet, all_weighted_ld_pairs, format_f3, henikoff_weights,
                  merge_shards, pair_order_key, plan_cell_tiles, plan_tiles, read_fasta, single_weighted_ld_pair, write_henikoff_weights, write_pair_stats)

from . import pycompat  # mirror of the reference's Python program (WeightedLD.py), WLD_COMPAT_PYTHON

__all__ = [
    "pycompat",
    "Context", "MultiSequence", "PairStore", "SiteSet", "WldError", "all_weighted_ld_pairs", "format_f3",
    "henikoff_weights", "merge_shards", "pair_order_key", "plan_cell_tiles", "plan_tiles", "read_fasta", "single_weighted_ld_pair", "write_henikoff_weights",
    "write_pair_stats", "PAIR_DTYPE",
]

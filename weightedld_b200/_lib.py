"""ctypes binding of libwld.so (include/wld.h).  No fallback: if the CUDA library is missing or
cannot be loaded this module raises, and every call that fails raises WldError with the
library's own message."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
import os as _os

LIB_PATH = Path(_os.environ.get("WLD_LIBRARY", _HERE / "libwld.so"))  # WLD_LIBRARY: A/B builds of the same ABI

PAIR_DTYPE = np.dtype(
    [("site_a", "<u4"), ("site_b", "<u4"), ("d", "<f4"), ("d_prime", "<f4"), ("r2", "<f4")]
)
assert PAIR_DTYPE.itemsize == 20

WLD_OK = 0
INPUT_ASCII, INPUT_CODES, INPUT_DEVICE = 0, 1, 2
FETCH_PARENT_INDEX, FETCH_KEPT_INDEX, FETCH_UNORDERED, FETCH_DEVICE = 0, 1, 2, 4
PAIR_KERNEL_UMMA, PAIR_KERNEL_SIMT, PAIR_KERNEL_UMMA_I8 = 0, 1, 2
COMPAT_RUST, COMPAT_PYTHON = 0, 1
EXCHANGE_HISTOGRAM, EXCHANGE_WEIGHT_SUMS = 0, 1
(STAGE_LOAD, STAGE_HISTOGRAM, STAGE_FILTER, STAGE_HENIKOFF, STAGE_PAIR_PREP, STAGE_PAIR, STAGE_ORDER, STAGE_PAIR_SAMPLE,
 STAGE_PAIR_REFINE) = range(9)
STAGE_NAMES = ["load", "histogram", "filter", "henikoff", "pair_prep", "pair", "order", "pair_sample", "pair_refine"]
STATUS_NAMES = {0: "OK", 1: "INVALID", 2: "STATE", 3: "CUDA", 4: "NOMEM", 5: "UNSUPPORTED", 6: "PANIC"}

PROGRESS_FN = C.CFUNCTYPE(None, C.c_uint64, C.c_void_p)


class PairInfo(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("n_limbs", C.c_int32), ("limb_bits", C.c_int32), ("weight_bits", C.c_int32),
        ("k_padded", C.c_int64), ("tiles", C.c_int64), ("tile_sites_m", C.c_int64), ("tile_sites_n", C.c_int64),
        ("executed_flop", C.c_double), ("die_schedule", C.c_int32), ("die_sms", C.c_int32 * 2), ("gain_bits", C.c_int32),
        ("weight_span_log2", C.c_int32), ("screen", C.c_int32), ("weight_rel_err", C.c_double),
        ("screen_candidates", C.c_int64), ("sample_pairs", C.c_int64), ("sample_candidates", C.c_int64),
        ("screen_top_min", C.c_int32), ("screen_reruns", C.c_int32),
        ("screen_cells", C.c_int64), ("screen_cells_flagged", C.c_int64),
        ("sample_tiles", C.c_int64), ("sample_tiles_flagged", C.c_int64),
    ]


class WldError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libwld: {STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


# every symbol include/wld.h declares: (restype, argtypes)
_vp, _i64, _u64, _int = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
SIGNATURES = {
    "wld_abi_version": (_int, []),
    "wld_create": (_int, [_int, C.POINTER(_vp)]),
    "wld_destroy": (None, [_vp]),
    "wld_last_error": (C.c_char_p, [_vp]),
    "wld_set_stream": (_int, [_vp, _vp]),
    "wld_set_partition": (_int, [_vp, _int, _int]),
    "wld_set_limbs": (_int, [_vp, _int]),
    "wld_set_gain_bits": (_int, [_vp, _int]),
    "wld_set_limb_bits": (_int, [_vp, _int]),
    "wld_get_pair_weights": (_int, [_vp, _vp, _i64]),
    "wld_set_pair_kernel": (_int, [_vp, _int]),
    "wld_set_pair_capacity": (_int, [_vp, _u64]),
    "wld_set_screen": (_int, [_vp, _int]),
    "wld_load_alignment": (_int, [_vp, _vp, _i64, _i64, _i64, _int]),
    "wld_load_alignment_rows": (_int, [_vp, _vp, _i64, _i64, _int]),
    "wld_filter_sites": (_int, [_vp, C.c_float, C.c_float, C.c_float, C.POINTER(_i64)]),
    "wld_keep_all_sites": (_int, [_vp, C.POINTER(_i64)]),
    "wld_n_seqs": (_i64, [_vp]),
    "wld_n_cols": (_i64, [_vp]),
    "wld_n_kept": (_i64, [_vp]),
    "wld_get_site_map": (_int, [_vp, _vp, _i64]),
    "wld_get_histograms": (_int, [_vp, _vp, _i64]),
    "wld_get_major_minor": (_int, [_vp, _vp, _vp, _i64]),
    "wld_get_codes": (_int, [_vp, _vp, _i64]),
    "wld_henikoff": (_int, [_vp]),
    "wld_henikoff_finish": (_int, [_vp]),
    "wld_set_row_shard": (_int, [_vp, _i64, _i64]),
    "wld_set_seq_shard": (_int, [_vp, _i64, _i64]),
    "wld_exchange_buffer": (_int, [_vp, _int, C.POINTER(_vp), C.POINTER(_u64)]),
    "wld_sum_histograms": (_int, [C.POINTER(_vp), _int]),
    "wld_share_weight_sums": (_int, [C.POINTER(_vp), _int]),
    "wld_set_weights": (_int, [_vp, _vp, _i64]),
    "wld_get_weights": (_int, [_vp, _vp, _i64]),
    "wld_get_weights_f64": (_int, [_vp, _vp, _i64]),
    "wld_ld_pairs": (_int, [_vp, C.c_float, PROGRESS_FN, _vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "wld_run": (_int, [_vp, _vp, _i64, _i64, _i64, _int, C.c_float, C.c_float, C.c_float, _vp, C.c_float, C.POINTER(_i64),
                       C.POINTER(_u64), C.POINTER(_u64)]),
    "wld_fetch_pairs": (_int, [_vp, _vp, _u64, _int, C.POINTER(_u64)]),
    "wld_fetch_pairs_range": (_int, [_vp, _u64, _u64, _vp, _int, C.POINTER(_u64)]),
    "wld_append_pairs": (_int, [_vp, _vp, _u64, _int]),
    "wld_append_pairs_from": (_int, [_vp, _vp]),
    "wld_pair_order_key": (_u64, [_i64, C.c_uint32, C.c_uint32]),
    "wld_plan_tiles": (_int, [_i64, _int, _int, _int, _int, _int, _vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "wld_plan_cell_tiles": (_int, [_i64, _int, _int, _int, _int, _vp, _u64, _vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "wld_set_cta_group": (_int, [_vp, _int]),
    "wld_set_compat": (_int, [_vp, _int]),
    "wld_filter_sites_python": (_int, [_vp, C.c_double, C.c_double, C.POINTER(_i64)]),
    "wld_stage_ms": (_int, [_vp, _int, C.POINTER(C.c_float)]),
    "wld_stage_launches": (_int, [_vp, _int, C.POINTER(_int)]),
    "wld_get_pair_info": (_int, [_vp, C.POINTER(PairInfo)]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libwld.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  weightedld_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib

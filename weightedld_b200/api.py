"""Host-side mirror of the reference's public Rust API, running on libwld.so (B200 only).

Names, argument meaning and error behaviour follow rust/weighted_ld/src/lib.rs so that a test
written against the reference reads the same here:

    reference (lib.rs)                         here
    -----------------------------------------  ---------------------------------------------
    read_fasta(path) -> MultiSequence   :277   read_fasta(path) -> MultiSequence
    SiteSet::from_multiseq(&ms)         :176   SiteSet.from_multiseq(ms)
    SiteSet::from_strs(&[..]) (tests)   :208   SiteSet.from_strs([...])
    siteset.filter_by(|s| is_site_of_interest(s, min_acgt, min_minor, max_minor))  :230,:310
                                               siteset.filter_by(min_acgt_frac, min_minor, max_minor)
    henikoff_weights(&siteset)          :340   henikoff_weights(siteset)
    all_weighted_ld_pairs(&ss, &w, thr, cb) :578   all_weighted_ld_pairs(ss, w, thr, cb) -> PairStore
    PairStore::{iter,len}               :533   PairStore.__iter__/__len__ (+ .array)
    write_henikoff_weights / write_pair_stats (main.rs:70-119)   same names

There is no CPU path: everything below the parsing of text files runs in CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Callable, Iterator

import numpy as np

from . import _lib as L
from ._lib import PAIR_DTYPE, PairInfo, WldError


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One wld_ctx: one run on one GPU."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        self._device = device
        self._h = C.c_void_p()
        rc = self._lib.wld_create(device, C.byref(self._h))
        if rc != L.WLD_OK:
            msg = self._lib.wld_last_error(self._h).decode() if self._h else "allocation failed"
            if self._h:
                self._lib.wld_destroy(self._h)
                self._h = C.c_void_p()
            raise WldError(rc, msg)
        self._keepalive = None
        self._stream_set = False  # set_stream called by the user: never override it

    def close(self):
        if getattr(self, "_h", None):
            self._lib.wld_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != L.WLD_OK:
            raise WldError(rc, self._lib.wld_last_error(self._h).decode())

    # ---- options
    def set_stream(self, cuda_stream_ptr: int | None):
        """Run on this CUDA stream.  None = the context's own stream; 0 (what torch reports for its default stream)
        = the legacy default stream, passed to the library as the explicit handle cudaStreamLegacy (0x1), because
        a NULL argument of wld_set_stream means "your own stream"."""
        self._stream_set = cuda_stream_ptr is not None
        handle = 0 if cuda_stream_ptr is None else (1 if cuda_stream_ptr == 0 else cuda_stream_ptr)
        self._check(self._lib.wld_set_stream(self._h, C.c_void_p(handle)))

    def adopt_torch_stream(self, device=None):
        """Run on torch's current stream (so that torch / NCCL work on tensors this context reads or writes is
        ordered with its kernels) unless the caller chose a stream explicitly."""
        if not self._stream_set:
            import torch

            s = torch.cuda.current_stream(device if device is not None else self._device).cuda_stream
            self._check(self._lib.wld_set_stream(self._h, C.c_void_p(1 if s == 0 else s)))

    def set_partition(self, part: int, nparts: int):
        self._check(self._lib.wld_set_partition(self._h, part, nparts))

    def set_limbs(self, n_limbs: int):
        self._check(self._lib.wld_set_limbs(self._h, n_limbs))

    def set_gain_bits(self, gain_bits: int):
        self._check(self._lib.wld_set_gain_bits(self._h, gain_bits))

    def set_limb_bits(self, limb_bits: int):
        self._check(self._lib.wld_set_limb_bits(self._h, limb_bits))

    def set_pair_kernel(self, kind: int | str):
        if isinstance(kind, str):
            kind = {"umma": L.PAIR_KERNEL_UMMA, "bf16": L.PAIR_KERNEL_UMMA, "simt": L.PAIR_KERNEL_SIMT, "i8": L.PAIR_KERNEL_UMMA_I8}[kind]
        self._check(self._lib.wld_set_pair_kernel(self._h, kind))

    def set_compat(self, mode: int | str):
        """'rust' (default) or 'python': numeric dialect, see wld_set_compat in include/wld.h."""
        if isinstance(mode, str):
            mode = {"rust": L.COMPAT_RUST, "python": L.COMPAT_PYTHON}[mode]
        self._check(self._lib.wld_set_compat(self._h, mode))

    def set_cta_group(self, ctas: int):
        self._check(self._lib.wld_set_cta_group(self._h, ctas))

    def set_pair_capacity(self, pairs: int):
        self._check(self._lib.wld_set_pair_capacity(self._h, pairs))

    def set_screen(self, mode: int | str):
        """One-limb screen + exact refinement of its candidates (wld_set_screen): 'never' / 0, 'auto' / 1
        (default: chosen from the candidate rate of a sample of the tiles), 'always' / 2.  Same survivors, bit for
        bit, as the exact n-limb kernel."""
        if isinstance(mode, str):
            mode = {"never": 0, "off": 0, "auto": 1, "always": 2}[mode]
        self._check(self._lib.wld_set_screen(self._h, mode))

    # ---- stage 1
    def load_alignment(self, chars, codes: bool = False):
        """chars: (n_seqs, n_cols) uint8 — numpy array (host, copied) or a CUDA torch tensor
        (device, borrowed; kept alive by this context)."""
        flags = L.INPUT_CODES if codes else L.INPUT_ASCII
        if isinstance(chars, np.ndarray):
            if chars.dtype != np.uint8 or chars.ndim != 2:
                raise ValueError("alignment must be a 2-D uint8 array")
            if chars.size and chars.strides[1] != 1:
                chars = np.ascontiguousarray(chars)
            stride = chars.strides[0] if chars.shape[0] > 1 else max(chars.shape[1], 1)
            if stride < chars.shape[1]:
                chars = np.ascontiguousarray(chars)
                stride = chars.shape[1]
            self._keepalive = chars
            self._check(self._lib.wld_load_alignment(self._h, _ptr(chars), chars.shape[0], chars.shape[1], stride, flags))
        else:  # torch tensor
            import torch

            t = chars
            if t.dtype != torch.uint8 or t.dim() != 2 or (t.numel() and t.stride(1) != 1):
                raise ValueError("alignment tensor must be 2-D uint8 with unit inner stride")
            if t.is_cuda:
                flags |= L.INPUT_DEVICE
                # A borrowed device buffer is read on the context's stream (include/wld.h, WLD_INPUT_DEVICE): run on
                # the torch stream that produced the tensor unless the caller chose a stream, so that the kernels are
                # ordered after the copy / collective that filled it.
                self.adopt_torch_stream(t.device)
            self._keepalive = t
            stride = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
            self._check(self._lib.wld_load_alignment(self._h, C.c_void_p(t.data_ptr()), t.shape[0], t.shape[1],
                                                     stride, flags))

    def load_alignment_rows(self, rows, codes: bool = False):
        """wld_load_alignment_rows: the sequences one by one (a list of equal-length 1-D uint8 arrays, e.g. views
        into a mapped FASTA file) — gathered by the library, never concatenated on the host."""
        rows = [np.ascontiguousarray(r, np.uint8) for r in rows]
        n_cols = len(rows[0]) if rows else 0
        if any(r.ndim != 1 or len(r) != n_cols for r in rows):
            raise ValueError("Not all sequences have the same number of symbols")  # lib.rs:181
        ptrs = (C.c_void_p * max(len(rows), 1))(*[r.ctypes.data for r in rows])
        self._keepalive = rows
        self._check(self._lib.wld_load_alignment_rows(self._h, ptrs, len(rows), n_cols, L.INPUT_CODES if codes else L.INPUT_ASCII))

    def filter_sites(self, min_acgt: float = 0.8, min_minor: float = 0.02, max_minor: float = 0.5) -> int:
        n = C.c_int64()
        self._check(self._lib.wld_filter_sites(self._h, min_acgt, min_minor, max_minor, C.byref(n)))
        return n.value

    def filter_sites_python(self, min_acgt: float = 0.8, min_variability: float = 0.02) -> int:
        """compute_variable_sites of the Python program (WeightedLD.py:44-98), LD mask."""
        n = C.c_int64()
        self._check(self._lib.wld_filter_sites_python(self._h, min_acgt, min_variability, C.byref(n)))
        return n.value

    def keep_all_sites(self) -> int:
        n = C.c_int64()
        self._check(self._lib.wld_keep_all_sites(self._h, C.byref(n)))
        return n.value

    @property
    def n_seqs(self) -> int:
        return self._lib.wld_n_seqs(self._h)

    @property
    def n_cols(self) -> int:
        return self._lib.wld_n_cols(self._h)

    @property
    def n_kept(self) -> int:
        return self._lib.wld_n_kept(self._h)

    def site_map(self) -> np.ndarray:
        out = np.empty(self.n_kept, np.int64)
        self._check(self._lib.wld_get_site_map(self._h, _ptr(out), len(out)))
        return out

    def histograms(self) -> np.ndarray:
        out = np.empty((self.n_cols, 6), np.uint32)
        self._check(self._lib.wld_get_histograms(self._h, _ptr(out), self.n_cols))
        return out

    def major_minor(self) -> tuple[np.ndarray, np.ndarray]:
        maj = np.empty(self.n_kept, np.int8)
        mnr = np.empty(self.n_kept, np.int8)
        self._check(self._lib.wld_get_major_minor(self._h, _ptr(maj), _ptr(mnr), self.n_kept))
        return maj, mnr

    def codes(self) -> np.ndarray:
        out = np.empty((self.n_kept, self.n_seqs), np.uint8)
        self._check(self._lib.wld_get_codes(self._h, _ptr(out), out.size))
        return out

    # ---- stages 1-2 on several GPUs
    def set_row_shard(self, lo: int, hi: int = -1):
        self._check(self._lib.wld_set_row_shard(self._h, lo, hi))

    def set_seq_shard(self, lo: int, hi: int = -1):
        self._check(self._lib.wld_set_seq_shard(self._h, lo, hi))

    def exchange_tensor(self, which: int):
        """The exchange buffer `which` (L.EXCHANGE_*) as a CUDA torch tensor aliasing the library's memory, for a
        torch.distributed collective on it (wld_exchange_buffer).  Histogram: int32 view of the u32 counts (sums
        stay far below 2^31); weight sums: float64."""
        import torch

        self.adopt_torch_stream()  # the collective must be ordered with this context's kernels
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        self._check(self._lib.wld_exchange_buffer(self._h, which, C.byref(ptr), C.byref(nbytes)))
        typestr, item, dtype = ("<i4", 4, torch.int32) if which == L.EXCHANGE_HISTOGRAM else ("<f8", 8, torch.float64)
        n = nbytes.value // item
        if n == 0:
            return torch.empty(0, dtype=dtype, device=f"cuda:{self._device}")

        class _Dev:  # zero-copy: torch reads the CUDA array interface
            __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr.value, False), "version": 2}

        return torch.as_tensor(_Dev(), device=f"cuda:{self._device}")

    def henikoff_finish(self):
        self._check(self._lib.wld_henikoff_finish(self._h))

    # ---- stage 2
    def henikoff(self):
        self._check(self._lib.wld_henikoff(self._h))

    def set_weights(self, w):
        w = np.ascontiguousarray(w, np.float32)
        self._check(self._lib.wld_set_weights(self._h, _ptr(w), len(w)))

    def weights(self) -> np.ndarray:
        out = np.empty(self.n_seqs, np.float32)
        self._check(self._lib.wld_get_weights(self._h, _ptr(out), len(out)))
        return out

    def weights_f64(self) -> np.ndarray:
        out = np.empty(self.n_seqs, np.float64)
        self._check(self._lib.wld_get_weights_f64(self._h, _ptr(out), len(out)))
        return out

    # ---- stage 3
    def ld_pairs(self, r2_threshold: float = 0.1, progress: Callable[[int], None] | None = None) -> tuple[int, int]:
        n, done = C.c_uint64(), C.c_uint64()
        cb = L.PROGRESS_FN(lambda v, _u: progress(int(v))) if progress else C.cast(None, L.PROGRESS_FN)
        self._check(self._lib.wld_ld_pairs(self._h, r2_threshold, cb, None, C.byref(n), C.byref(done)))
        return n.value, done.value

    def run(self, chars, min_acgt=0.8, min_minor=0.02, max_minor=0.5, weights=None, r2_threshold=0.1, codes=False):
        """wld_run: main.rs:129-190 in one call -> (n_kept, n_survivors, pairs_computed)."""
        flags = L.INPUT_CODES if codes else L.INPUT_ASCII
        if isinstance(chars, np.ndarray):
            chars = np.ascontiguousarray(chars, np.uint8)
            ptr, stride = _ptr(chars), chars.shape[1]
        else:
            import torch

            if chars.is_cuda:
                flags |= L.INPUT_DEVICE
                self.adopt_torch_stream(chars.device)
            ptr, stride = C.c_void_p(chars.data_ptr()), (chars.stride(0) if chars.shape[0] > 1 else max(chars.shape[1], 1))
        self._keepalive = chars
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        k, n, done = C.c_int64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.wld_run(self._h, ptr, chars.shape[0], chars.shape[1], stride, flags, min_acgt, min_minor, max_minor,
                                      _ptr(w), r2_threshold, C.byref(k), C.byref(n), C.byref(done)))
        return k.value, n.value, done.value

    def fetch_pairs(self, n: int, flags: int = L.FETCH_PARENT_INDEX, out: np.ndarray | None = None) -> np.ndarray:
        """Survivors as a structured array.  `out` may be a preallocated (e.g. pinned) PAIR_DTYPE array."""
        if out is None:
            out = np.empty(n, PAIR_DTYPE)
        elif out.dtype != PAIR_DTYPE or len(out) < n or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous PAIR_DTYPE array with room for n records")
        got = C.c_uint64()
        self._check(self._lib.wld_fetch_pairs(self._h, _ptr(out), len(out), flags, C.byref(got)))
        return out[: got.value]

    def fetch_pairs_range(self, first: int, count: int, flags: int = L.FETCH_PARENT_INDEX, out: np.ndarray | None = None) -> np.ndarray:
        """Survivors [first, first+count) of the (ordered) result: wld_fetch_pairs_range, for streaming writers."""
        if out is None:
            out = np.empty(count, PAIR_DTYPE)
        got = C.c_uint64()
        self._check(self._lib.wld_fetch_pairs_range(self._h, first, min(count, len(out)), _ptr(out), flags, C.byref(got)))
        return out[: got.value]

    def fetch_pairs_device(self, n: int, flags: int = L.FETCH_KEPT_INDEX | L.FETCH_UNORDERED):
        """Survivors as a CUDA uint8 torch tensor of n*20 bytes on this context's GPU (WLD_FETCH_DEVICE), e.g.
        a shard to hand to NCCL."""
        import torch

        dev = torch.device("cuda", self._lib_device())
        t = torch.empty(max(n, 1) * PAIR_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        got = C.c_uint64()
        self._check(self._lib.wld_fetch_pairs(self._h, C.c_void_p(t.data_ptr()), n, flags | L.FETCH_DEVICE, C.byref(got)))
        return t[: got.value * PAIR_DTYPE.itemsize]

    def append_pairs(self, shard) -> None:
        """wld_append_pairs: adds another partition's survivors (KEPT indices) — a PAIR_DTYPE numpy array or a
        CUDA uint8 torch tensor on this GPU — so that fetch_pairs returns the union in the reference's order."""
        if isinstance(shard, np.ndarray):
            shard = np.ascontiguousarray(shard, PAIR_DTYPE)
            self._check(self._lib.wld_append_pairs(self._h, _ptr(shard), len(shard), 0))
        else:
            assert shard.is_cuda and shard.is_contiguous() and shard.numel() % PAIR_DTYPE.itemsize == 0
            self._check(self._lib.wld_append_pairs(self._h, C.c_void_p(shard.data_ptr()), shard.numel() // PAIR_DTYPE.itemsize, 1))

    def _lib_device(self) -> int:
        return self._device

    def pair_weights(self) -> np.ndarray:
        """The integer weights q[s] the last pair stage summed (wld_get_pair_weights)."""
        out = np.empty(self.n_seqs, np.float64)
        self._check(self._lib.wld_get_pair_weights(self._h, _ptr(out), len(out)))
        return out

    # ---- introspection
    def stage_ms(self, stage: int) -> float:
        ms = C.c_float()
        self._check(self._lib.wld_stage_ms(self._h, stage, C.byref(ms)))
        return ms.value

    def stage_launches(self, stage: int) -> int:
        n = C.c_int()
        self._check(self._lib.wld_stage_launches(self._h, stage, C.byref(n)))
        return n.value

    def pair_info(self) -> PairInfo:
        info = PairInfo()
        self._check(self._lib.wld_get_pair_info(self._h, C.byref(info)))
        return info


def pair_order_key(n_kept: int, kept_a, kept_b) -> np.ndarray:
    """Vectorised wld_pair_order_key: reference tile order of lib.rs:623-632."""
    n = (n_kept + 255) // 256
    tr = np.asarray(kept_a, np.uint64) // np.uint64(256)
    tc = np.asarray(kept_b, np.uint64) // np.uint64(256)
    return (np.uint64(n - 1) - tr) * np.uint64(n) + tc


def plan_tiles(n_kept: int, n_limbs: int = 3, part: int = 0, nparts: int = 1, sm_count: int = 148,
               cta_group: int = 2):
    """Host-only pair-stage schedule (wld_plan_tiles): ((n_tiles, 4) uint32 {M tile, N tile, first site j,
    end site j} — the window of site columns of the tile that belongs to this partition —, site pairs covered).
    n_limbs = 1 is the schedule of the one-limb screen.  Needs no GPU."""
    lib = L.load()
    n, pairs = C.c_uint64(), C.c_uint64()
    rc = lib.wld_plan_tiles(n_kept, n_limbs, cta_group, part, nparts, sm_count, None, 0, C.byref(n), C.byref(pairs))
    if rc != L.WLD_OK:
        raise WldError(rc, "bad tile plan arguments")
    tiles = np.empty((n.value, 4), np.uint32)
    rc = lib.wld_plan_tiles(n_kept, n_limbs, cta_group, part, nparts, sm_count, _ptr(tiles), n.value, C.byref(n), C.byref(pairs))
    if rc != L.WLD_OK:
        raise WldError(rc, "bad tile plan arguments")
    return tiles, pairs.value


def plan_cell_tiles(n_kept: int, flags, n_limbs: int = 3, part: int = 0, nparts: int = 1, cta_group: int = 2):
    """Host-only schedule of the exact kernel over the flagged tiles of the screen's schedule (wld_plan_cell_tiles):
    ((n_tiles, 4) uint32 {M tile, N tile, first site j, end site j}, site pairs covered)."""
    lib = L.load()
    flags = np.ascontiguousarray(flags, np.uint8)
    n, pairs = C.c_uint64(), C.c_uint64()
    rc = lib.wld_plan_cell_tiles(n_kept, n_limbs, cta_group, part, nparts, _ptr(flags), len(flags), None, 0, C.byref(n), C.byref(pairs))
    if rc != L.WLD_OK:
        raise WldError(rc, "bad cell plan arguments")
    tiles = np.empty((n.value, 4), np.uint32)
    rc = lib.wld_plan_cell_tiles(n_kept, n_limbs, cta_group, part, nparts, _ptr(flags), len(flags), _ptr(tiles), n.value,
                                 C.byref(n), C.byref(pairs))
    if rc != L.WLD_OK:
        raise WldError(rc, "bad cell plan arguments")
    return tiles, pairs.value


def merge_shards(n_kept: int, shards: list[np.ndarray], site_map: np.ndarray | None = None) -> np.ndarray:
    """Host merge of per-GPU survivor shards (records with KEPT indices) into the reference's
    output order (lib.rs:623-679); maps to raw columns when site_map is given (lib.rs:662-663)."""
    allp = np.concatenate(shards) if shards else np.empty(0, PAIR_DTYPE)
    order = np.lexsort((allp["site_b"], allp["site_a"], pair_order_key(n_kept, allp["site_a"], allp["site_b"])))
    out = allp[order]
    if site_map is not None and len(out):
        out["site_a"] = np.asarray(site_map)[out["site_a"]]
        out["site_b"] = np.asarray(site_map)[out["site_b"]]
    return out


# =================================================================================================
# Mirror of the reference's Rust API
# =================================================================================================
@dataclass
class MultiSequence:
    """lib.rs:153-156.  `chars` is the (n_seqs, n_cols) byte matrix; names as in lib.rs:143-146."""
    chars: np.ndarray
    names: list[str | None]
    source: str | None = None


def read_fasta(path: str | os.PathLike) -> MultiSequence:
    """read_fasta, lib.rs:277-307: '>' lines are names; EVERY other line is one whole sequence
    including its line terminator, which becomes an Unknown column (lib.rs:297).  Sequences of
    unequal length raise ValueError where the reference panics (lib.rs:180-182)."""
    data = np.fromfile(path, dtype=np.uint8)
    if data.size == 0:
        return MultiSequence(np.zeros((0, 0), np.uint8), [], str(path))
    nl = np.flatnonzero(data == 10)
    starts = np.concatenate(([0], nl + 1))
    ends = np.concatenate((nl + 1, [data.size]))  # a line owns its '\n'
    if starts[-1] >= data.size:  # file ends with '\n': no trailing empty line
        starts, ends = starts[:-1], ends[:-1]
    is_name = data[starts] == ord(">")
    seq_starts, seq_ends = starts[~is_name], ends[~is_name]
    names: list[str | None] = []
    pending = None
    for s, e, nm in zip(starts.tolist(), ends.tolist(), is_name.tolist()):
        if nm:
            pending = bytes(data[s + 1:e]).decode("utf-8", "replace")
        else:
            names.append(pending)
            pending = None
    if seq_starts.size == 0:
        return MultiSequence(np.zeros((0, 0), np.uint8), names, str(path))
    lens = seq_ends - seq_starts
    if np.any(lens != lens[0]):
        raise ValueError("Not all sequences have the same number of symbols")  # lib.rs:181
    n_cols = int(lens[0])
    idx = seq_starts[:, None] + np.arange(n_cols, dtype=np.int64)[None, :]
    return MultiSequence(np.ascontiguousarray(data[idx]), names, str(path))


class SiteSet:
    """lib.rs:158-275, resident on the GPU behind a wld context."""

    def __init__(self, ctx: Context, filtered: bool):
        self._ctx = ctx
        self._filtered = filtered

    @classmethod
    def from_multiseq(cls, ms: MultiSequence, device: int = 0) -> "SiteSet":
        if ms.chars.shape[0] == 0:
            raise IndexError("index out of bounds: the len is 0 but the index is 0")  # lib.rs:178
        ctx = Context(device)
        ctx.load_alignment(ms.chars)
        ctx.keep_all_sites()
        return cls(ctx, False)

    @classmethod
    def from_strs(cls, rows: list[str], device: int = 0) -> "SiteSet":
        if len({len(r) for r in rows}) > 1:
            raise ValueError("Not all sequences have the same number of symbols")
        chars = np.frombuffer("".join(rows).encode(), np.uint8).reshape(len(rows), -1)
        return cls.from_multiseq(MultiSequence(chars, [None] * len(rows)), device)

    @classmethod
    def from_codes(cls, codes_seq_major, device: int = 0) -> "SiteSet":
        """Already-encoded 0..5 matrix (n_seqs, n_sites), e.g. from a VCF (WeightedLD.py:311-379)."""
        ctx = Context(device)
        ctx.load_alignment(codes_seq_major, codes=True)
        ctx.keep_all_sites()
        return cls(ctx, False)

    def filter_by(self, min_acgt: float = 0.8, min_minor: float = 0.02, max_minor: float = 0.5) -> "SiteSet":
        """siteset.filter_by(|s| is_site_of_interest(s, ceil(min_acgt*n), min_minor, max_minor)),
        main.rs:139-143.  Re-filters the context's alignment in place (the reference returns a new
        SiteSet; the unfiltered one stays valid there, here it is re-derivable with keep_all)."""
        self._ctx.filter_sites(min_acgt, min_minor, max_minor)
        return SiteSet(self._ctx, True)

    def n_sites(self) -> int:
        return self._ctx.n_kept

    def n_seqs(self) -> int:
        return self._ctx.n_seqs

    def parent_site_index(self, idx: int) -> int:
        return int(self._ctx.site_map()[idx])

    def site_map(self) -> np.ndarray:
        return self._ctx.site_map()

    def site_symbols(self, index: int) -> np.ndarray:
        return self._ctx.codes()[index]

    def site_histogram(self, index: int) -> np.ndarray:
        return self._ctx.histograms()[self.parent_site_index(index)]

    @property
    def context(self) -> Context:
        return self._ctx


def henikoff_weights(data: SiteSet) -> np.ndarray:
    """lib.rs:340-358 -> Vec<f32>."""
    data.context.henikoff()
    return data.context.weights()


class PairStore:
    """lib.rs:529-576."""

    def __init__(self, array: np.ndarray, pairs_computed: int):
        self.array = array
        self.pairs_computed = pairs_computed

    def __len__(self) -> int:
        return len(self.array)

    def __iter__(self) -> Iterator[tuple[int, int, tuple[float, float, float]]]:
        for p in self.array:
            yield int(p["site_a"]), int(p["site_b"]), (float(p["r2"]), float(p["d"]), float(p["d_prime"]))


def all_weighted_ld_pairs(site_set: SiteSet, weights, r2_threshold: float,
                          progress_report: Callable[[int], None] | None = None) -> PairStore:
    """lib.rs:578-684."""
    ctx = site_set.context
    ctx.set_weights(weights)
    n, done = ctx.ld_pairs(r2_threshold, progress_report)
    return PairStore(ctx.fetch_pairs(n), done)


def single_weighted_ld_pair(a_symbols, b_symbols, weights, device: int = 0):
    """lib.rs:390-521 for one pair of sites given as code arrays; returns (r2, d, d_prime) or None.
    (A two-site alignment through the same kernels; threshold -inf keeps any non-NaN result.)"""
    codes = np.stack([np.asarray(a_symbols, np.uint8), np.asarray(b_symbols, np.uint8)], axis=1)
    with Context(device) as ctx:
        ctx.load_alignment(np.ascontiguousarray(codes), codes=True)
        ctx.keep_all_sites()
        ctx.set_weights(weights)
        n, _ = ctx.ld_pairs(-math.inf)
        if n == 0:
            return None
        p = ctx.fetch_pairs(n)[0]
        return float(p["r2"]), float(p["d"]), float(p["d_prime"])


def format_f3(v: float) -> str:
    """Rust `{:.3}` for f32 (main.rs:76,106): round-half-even of the exact binary value, `NaN`,
    `inf`, `-inf`, sign of negative zero kept."""
    v = float(v)
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "-inf" if v < 0 else "inf"
    return f"{v:.3f}"


def write_henikoff_weights(path, weights) -> None:
    """main.rs:70-80."""
    with open(path, "w") as f:
        f.write("Sequence_index\thk_weight\n")
        for i, w in enumerate(np.asarray(weights, np.float32)):
            f.write(f"{i}\t{format_f3(w)}\n")


def write_pair_stats(path, pairs: PairStore) -> None:
    """main.rs:82-119."""
    with open(path, "w") as f:
        f.write("site_a\tsite_b\td\td'\tr2\n")
        for p in pairs.array:
            f.write(f"{p['site_a']}\t{p['site_b']}\t{format_f3(p['d'])}\t{format_f3(p['d_prime'])}\t{format_f3(p['r2'])}\n")

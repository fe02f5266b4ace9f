"""ctypes front-end of the CPU oracle (oracle/wld_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of wld_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package weightedld_b200 never does.

The functions mirror the reference's public Rust API (rust/weighted_ld/src/lib.rs):
read_fasta (lib.rs:277), SiteSet.from_multiseq (lib.rs:176), SiteSet.filter_by +
is_site_of_interest (lib.rs:230, 310), henikoff_weights (lib.rs:340),
single_weighted_ld_pair (lib.rs:390), all_weighted_ld_pairs (lib.rs:578).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"

F32_SCALAR, F32_SIMD8, F64 = 0, 1, 2

PAIR_DTYPE = np.dtype(
    [("a", "<u4"), ("b", "<u4"), ("d", "<f4"), ("d_prime", "<f4"), ("r2", "<f4")]
)


class _LdStats32(C.Structure):
    _fields_ = [("r2", C.c_float), ("d", C.c_float), ("d_prime", C.c_float)]


class _LdStats64(C.Structure):
    _fields_ = [("r2", C.c_double), ("d", C.c_double), ("d_prime", C.c_double)]


def build(native: bool = False, force: bool = False) -> Path:
    """Compile the oracle with oracle/Makefile (gcc only, no reference sources needed)."""
    target = _BUILD / ("liboracle_native.so" if native else "liboracle.so")
    src = _HERE / "wld_oracle.c"
    if force or not target.exists() or target.stat().st_mtime < src.stat().st_mtime:
        if force and target.exists():
            target.unlink()
        subprocess.run(["make", "-C", str(_HERE), f"_build/{target.name}"], check=True,
                       stdout=subprocess.DEVNULL)
    return target


_libs: dict[bool, C.CDLL] = {}


def lib(native: bool = False) -> C.CDLL:
    if native not in _libs:
        L = C.CDLL(str(build(native)))
        u8p, i64, u64 = C.POINTER(C.c_uint8), C.c_int64, C.c_uint64
        L.wldo_encode_char.restype = C.c_uint8
        L.wldo_encode_char.argtypes = [C.c_uint8]
        L.wldo_min_acgt_count.restype = u64
        L.wldo_min_acgt_count.argtypes = [C.c_float, u64]
        L.wldo_is_site_of_interest.restype = C.c_int
        L.wldo_is_site_of_interest.argtypes = [C.c_void_p, u64, C.c_float, C.c_float]
        L.wldo_major_minor.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.wldo_distinct_known_count.restype = C.c_int
        L.wldo_distinct_known_count.argtypes = [C.c_void_p]
        L.wldo_build_siteset.argtypes = [C.c_void_p, i64, i64, i64, C.c_void_p, C.c_void_p]
        L.wldo_filter_sites.restype = i64
        L.wldo_filter_sites.argtypes = [C.c_void_p, i64, i64, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.wldo_gather_sites.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p]
        L.wldo_henikoff_f32.argtypes = [C.c_void_p, i64, i64, C.c_void_p]
        L.wldo_henikoff_f64.argtypes = [C.c_void_p, i64, i64, C.c_void_p]
        for name, st in (("wldo_pair_f32", _LdStats32), ("wldo_pair_f32_simd8", _LdStats32),
                         ("wldo_pair_f64", _LdStats64)):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                           i64, C.POINTER(st)]
        L.wldo_quantize_weights.argtypes = [C.c_void_p, i64, C.c_int, C.c_void_p]
        L.wldo_quantize_weights_gain.argtypes = [C.c_void_p, i64, C.c_int, C.c_int, C.c_void_p]
        L.wldo_triu_index.argtypes = [u64, u64, C.POINTER(u64), C.POINTER(u64)]
        L.wldo_all_pairs.restype = u64
        L.wldo_all_pairs.argtypes = [C.c_void_p, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_float, C.c_int, u64, u64, C.c_void_p, u64,
                                     C.POINTER(u64)]
        L.wldo_tile_count.restype = u64
        L.wldo_tile_count.argtypes = [i64]
        L.wldo_format_f3.restype = C.c_int
        L.wldo_format_f3.argtypes = [C.c_float, C.c_char_p, C.c_size_t]
        L.wldo_write_pairs.restype = C.c_int
        L.wldo_write_pairs.argtypes = [C.c_char_p, C.c_void_p, u64]
        L.wldo_write_weights.restype = C.c_int
        L.wldo_write_weights.argtypes = [C.c_char_p, C.c_void_p, u64]
        L.wldo_read_fasta.restype = i64
        L.wldo_read_fasta.argtypes = [C.c_char_p, C.POINTER(u8p), C.POINTER(i64)]
        L.wldo_free.argtypes = [C.c_void_p]
        L.wldo_max_threads.restype = C.c_int
        L.wldo_set_threads.argtypes = [C.c_int]
        _libs[native] = L
    return _libs[native]


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- data model
@dataclass
class SiteSet:
    """lib.rs:158-275: site-major code buffer + per-site histograms (+ parent map)."""
    codes: np.ndarray      # (n_sites, n_seqs) uint8, C-contiguous
    hists: np.ndarray      # (n_sites, 6) uint64
    site_map: np.ndarray | None = None  # (n_sites,) int64 parent indices

    @property
    def n_sites(self) -> int:
        return int(self.codes.shape[0])

    @property
    def n_seqs(self) -> int:
        return int(self.codes.shape[1])

    def parent_site_index(self, idx: int) -> int:  # lib.rs:263-265
        return int(self.site_map[idx]) if self.site_map is not None else idx

    def major_minor(self) -> tuple[np.ndarray, np.ndarray]:
        maj = np.empty(self.n_sites, np.int32)
        mnr = np.empty(self.n_sites, np.int32)
        a, b = C.c_int(), C.c_int()
        h = np.ascontiguousarray(self.hists, np.uint64)
        for i in range(self.n_sites):
            lib().wldo_major_minor(_p(h[i]), C.byref(a), C.byref(b))
            maj[i], mnr[i] = a.value, b.value
        return maj, mnr


def read_fasta(path: str | os.PathLike) -> np.ndarray:
    """lib.rs:277-307.  Returns the (n_seqs, n_cols) uint8 character matrix; the newline
    of every sequence line is a column.  Raises ValueError where Rust panics (lib.rs:180-182)."""
    ptr = C.POINTER(C.c_uint8)()
    ncols = C.c_int64()
    n = lib().wldo_read_fasta(str(path).encode(), C.byref(ptr), C.byref(ncols))
    if n == -1:
        raise OSError(f"cannot read {path}")
    if n == -2:
        raise ValueError("Not all sequences have the same number of symbols")
    out = np.ctypeslib.as_array(ptr, shape=(n, ncols.value)).copy() if n > 0 else np.zeros((0, 0), np.uint8)
    lib().wldo_free(ptr)
    return out


def encode(chars: np.ndarray) -> np.ndarray:
    lut = np.array([lib().wldo_encode_char(c) for c in range(256)], np.uint8)
    return lut[chars]


def siteset_from_chars(chars: np.ndarray) -> SiteSet:
    """SiteSet::from_multiseq, lib.rs:176-206."""
    chars = np.ascontiguousarray(chars, np.uint8)
    n_seqs, n_cols = chars.shape
    codes = np.empty((n_cols, n_seqs), np.uint8)
    hists = np.empty((n_cols, 6), np.uint64)
    lib().wldo_build_siteset(_p(chars), n_seqs, n_cols, chars.strides[0], _p(codes), _p(hists))
    return SiteSet(codes, hists, None)


def siteset_from_strs(rows: list[str]) -> SiteSet:
    """SiteSet::from_strs, lib.rs:208-228 (test helper of the reference)."""
    chars = np.frombuffer("".join(rows).encode(), np.uint8).reshape(len(rows), -1)
    return siteset_from_chars(chars)


def siteset_from_codes(codes_site_major: np.ndarray) -> SiteSet:
    codes = np.ascontiguousarray(codes_site_major, np.uint8)
    hists = np.stack([(codes == k).sum(axis=1) for k in range(6)], axis=1).astype(np.uint64)
    return SiteSet(codes, np.ascontiguousarray(hists), None)


def filter_sites(ss: SiteSet, min_acgt: float = 0.8, min_minor: float = 0.02,
                 max_minor: float = 0.5) -> SiteSet:
    """main.rs:139-143 + lib.rs:230-251, 310-338."""
    site_map = np.empty(max(ss.n_sites, 1), np.int64)
    hists = np.ascontiguousarray(ss.hists, np.uint64)
    k = lib().wldo_filter_sites(_p(hists), ss.n_sites, ss.n_seqs, min_acgt, min_minor, max_minor,
                                _p(site_map))
    site_map = site_map[:k].copy()
    codes = np.empty((k, ss.n_seqs), np.uint8)
    oh = np.empty((k, 6), np.uint64)
    if k:
        lib().wldo_gather_sites(_p(ss.codes), _p(hists), ss.n_seqs, _p(site_map), k, _p(codes), _p(oh))
    if ss.site_map is not None:
        site_map = ss.site_map[site_map]
    return SiteSet(codes, oh, site_map)


def henikoff_weights(ss: SiteSet, f64: bool = False) -> np.ndarray:
    """lib.rs:340-380."""
    out = np.empty(ss.n_seqs, np.float64 if f64 else np.float32)
    (lib().wldo_henikoff_f64 if f64 else lib().wldo_henikoff_f32)(_p(ss.codes), ss.n_sites, ss.n_seqs, _p(out))
    return out


def single_weighted_ld_pair(a: np.ndarray, b: np.ndarray, weights: np.ndarray, flavour: int = F32_SCALAR):
    """lib.rs:390-521.  Returns (r2, d, d_prime) or None."""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    ha = np.array([(a == k).sum() for k in range(6)], np.uint64)
    hb = np.array([(b == k).sum() for k in range(6)], np.uint64)
    am, an, bm, bn = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib().wldo_major_minor(_p(ha), C.byref(am), C.byref(an))
    lib().wldo_major_minor(_p(hb), C.byref(bm), C.byref(bn))
    if flavour == F64:
        w = np.ascontiguousarray(weights, np.float64)
        st = _LdStats64()
        ok = lib().wldo_pair_f64(_p(a), _p(b), am, an, bm, bn, _p(w), len(a), C.byref(st))
    else:
        w = np.ascontiguousarray(weights, np.float32)
        st = _LdStats32()
        fn = lib().wldo_pair_f32_simd8 if flavour == F32_SIMD8 else lib().wldo_pair_f32
        ok = fn(_p(a), _p(b), am, an, bm, bn, _p(w), len(a), C.byref(st))
    return (st.r2, st.d, st.d_prime) if ok else None


def quantize_weights(w: np.ndarray, bits: int = 24, gain_bits: int = 0) -> np.ndarray:
    """Integer weights of the B200 pair stage (not in the reference; DESIGN.md "precision contract"):
    `bits` mantissa bits and `gain_bits` block-exponent bits, as wld_pair_info reports them."""
    w = np.ascontiguousarray(w, np.float32)
    out = np.empty(len(w), np.float64)
    if gain_bits:
        lib().wldo_quantize_weights_gain(_p(w), len(w), bits, gain_bits, _p(out))
    else:
        lib().wldo_quantize_weights(_p(w), len(w), bits, _p(out))
    return out


def auto_quant_params(w: np.ndarray) -> tuple[int, int]:
    """(mantissa bits, gain bits) the library picks on its own for these f32 weights when no limb sum
    overflows (include/wld.h wld_set_limbs / wld_set_gain_bits): all equal -> (0, 0); else x = -exponent of
    min_nonzero/max, 4 limbs if x > 7 else 3, gain bits min(7, x)."""
    w = np.asarray(w, np.float32)
    if w.min() == w.max():
        return 0, 0
    x = max(0, -int(np.frexp(np.float64(w[w > 0].min()) / np.float64(w.max()))[1]))
    return (32 if x > 7 else 24), min(7, x)


def all_weighted_ld_pairs(ss: SiteSet, weights: np.ndarray, r2_threshold: float = 0.1,
                          flavour: int = F32_SCALAR, tile_range: tuple[int, int] | None = None,
                          store: bool = True, native: bool = False, cap: int | None = None):
    """lib.rs:578-684.  Returns (pairs structured array in reference order, pairs_computed)."""
    L = lib(native)
    w32 = np.ascontiguousarray(weights, np.float32) if flavour != F64 else None
    w64 = np.ascontiguousarray(weights, np.float64) if flavour == F64 else None
    hists = np.ascontiguousarray(ss.hists, np.uint64)
    site_map = None if ss.site_map is None else np.ascontiguousarray(ss.site_map, np.int64)
    t0, t1 = tile_range if tile_range is not None else (0, 2**64 - 1)
    computed = C.c_uint64()
    if not store:
        n = L.wldo_all_pairs(_p(ss.codes), ss.n_sites, ss.n_seqs, _p(hists), _p(site_map), _p(w32), _p(w64),
                             r2_threshold, flavour, t0, t1, None, 0, C.byref(computed))
        return int(n), int(computed.value)
    if cap is None:
        cap = max(1024, ss.n_sites * (ss.n_sites - 1) // 2 if ss.n_sites < 4096 else 1 << 22)
    while True:
        out = np.empty(cap, PAIR_DTYPE)
        n = L.wldo_all_pairs(_p(ss.codes), ss.n_sites, ss.n_seqs, _p(hists), _p(site_map), _p(w32), _p(w64),
                             r2_threshold, flavour, t0, t1, _p(out), cap, C.byref(computed))
        if n <= cap:
            return out[:n].copy(), int(computed.value)
        cap = int(n)


def triu_index(n: int, i: int) -> tuple[int, int]:
    r, c = C.c_uint64(), C.c_uint64()
    lib().wldo_triu_index(n, i, C.byref(r), C.byref(c))
    return int(r.value), int(c.value)


def format_f3(v: float) -> str:
    buf = C.create_string_buffer(64)
    lib().wldo_format_f3(v, buf, 64)
    return buf.value.decode()


def write_pairs(path, pairs: np.ndarray) -> None:
    pairs = np.ascontiguousarray(pairs, PAIR_DTYPE)
    if lib().wldo_write_pairs(str(path).encode(), _p(pairs), len(pairs)) != 0:
        raise OSError(f"cannot write {path}")


def write_weights(path, w: np.ndarray) -> None:
    w = np.ascontiguousarray(w, np.float32)
    if lib().wldo_write_weights(str(path).encode(), _p(w), len(w)) != 0:
        raise OSError(f"cannot write {path}")


def run_pipeline(chars: np.ndarray, min_acgt=0.8, min_minor=0.02, max_minor=0.5, r2_threshold=0.1,
                 unweighted=False, flavour=F32_SCALAR, quant_bits: int | None = None):
    """main.rs:129-190 as one call.  Returns (filtered SiteSet, weights, pairs)."""
    ss = siteset_from_chars(chars)
    fs = filter_sites(ss, min_acgt, min_minor, max_minor)
    if unweighted:
        w = np.ones(ss.n_seqs, np.float32)  # main.rs:150-153
    else:
        w = henikoff_weights(fs, f64=(flavour == F64))
    w_pairs = w
    if flavour == F64:
        w32 = w.astype(np.float32)  # the pair stage consumes Vec<f32> (lib.rs:580)
        w_pairs = quantize_weights(w32, quant_bits) if quant_bits else w32.astype(np.float64)
    pairs, _ = all_weighted_ld_pairs(fs, w_pairs, r2_threshold, flavour)
    return fs, w, pairs

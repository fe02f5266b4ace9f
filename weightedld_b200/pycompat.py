"""Host-side mirror of the reference's *Python* program (WeightedLD.py) on libwld.so (B200 only).

Same function names, arguments and printed output as /root/reference/WeightedLD.py, so that a caller
of that script (or of its functions) can switch over:

    reference (WeightedLD.py)                        here
    -----------------------------------------------  ------------------------------------------
    read_fasta(filename) -> codes            :21     read_fasta(filename)            (host text parsing)
    compute_variable_sites(aln, a, v)        :44     compute_variable_sites(aln, a, v)   GPU histogram + filter
    henikoff_weighting(aln)                  :101    henikoff_weighting(aln)             GPU
    ld(aln, weights, site_map)  (prints)     :154    ld(aln, weights, site_map, file=stdout)  GPU pair stage
    handle_fasta(args) / handle_vcf(filename) :287/:311   same
    main(args)                               :382    main(args);  `python -m weightedld_b200 --file X`

The numeric dialect is WLD_COMPAT_PYTHON (include/wld.h): Python's site filter, Henikoff fill, per-pair
allele calls and PA/PB skip.  Statistics are evaluated in f64 on exact 24-bit fixed-point weighted sums
and carried as f32 (the library's record type), then printed like the reference, `round(x, 4)`: they
agree with the reference's printed values except where a value sits within ~1e-7 of a rounding
boundary.  Pairs with an empty marginal (printed as nan/inf by the reference, with numpy warnings) are
not printed.  There is no CPU path: without a B200 and libwld.so every function below raises.
"""
from __future__ import annotations

import argparse
import math
import re
import sys
from pathlib import Path

import numpy as np

from . import _lib as L
from .api import Context

_LUT = np.full(256, 5, np.uint8)
for _ch, _code in zip(b"acgt-", range(5)):
    _LUT[_ch] = _code
    _LUT[ord(chr(_ch).upper())] = _code


def read_fasta(filename) -> np.ndarray:
    """WeightedLD.py:21-41 (Bio.AlignIO "fasta"): a record is '>' + following lines concatenated,
    case-insensitive, a c g t - -> 0..4 and everything else 5.  No newline column (unlike lib.rs:297)."""
    data = Path(filename).read_bytes()
    recs: list[list[bytes]] = []
    cur = None
    for line in data.splitlines():
        if line.startswith(b">"):
            cur = []
            recs.append(cur)
        elif cur is not None:
            cur.append(line.strip())
    rows = [b"".join(r) for r in recs]
    if not rows:
        raise ValueError("No records found in handle")          # Bio.AlignIO.read on an empty file
    if len({len(r) for r in rows}) > 1:
        raise ValueError("Sequences must all be the same length")  # Bio.Align.MultipleSeqAlignment
    chars = np.frombuffer(b"".join(rows), np.uint8).reshape(len(rows), -1)
    return _LUT[chars]


def _python_context(alignment: np.ndarray, device: int = 0) -> Context:
    aln = np.ascontiguousarray(alignment, np.uint8)
    if aln.ndim != 2:
        raise ValueError("alignment must be a 2-D (sequences x sites) array")
    ctx = Context(device)
    ctx.set_compat("python")
    ctx.load_alignment(aln, codes=True)
    return ctx


def compute_variable_sites(alignment: np.ndarray, min_acgt: float, min_variability: float, device: int = 0):
    """WeightedLD.py:44-98 -> (return_hk_varsites, return_ld_varsites) boolean masks."""
    with _python_context(alignment, device) as ctx:
        ctx.filter_sites_python(min_acgt, min_variability)
        ld_mask = np.zeros(ctx.n_cols, bool)
        ld_mask[ctx.site_map()] = True
        # hk mask = sufficient_data alone (WeightedLD.py:92): the same filter with no variability demand
        ctx.filter_sites_python(min_acgt, -math.inf)
        hk_mask = np.zeros(ctx.n_cols, bool)
        hk_mask[ctx.site_map()] = True
    return hk_mask, ld_mask


def henikoff_weighting(alignment: np.ndarray, device: int = 0) -> np.ndarray:
    """WeightedLD.py:101-151 -> float64 weights, max exactly 1."""
    with _python_context(alignment, device) as ctx:
        ctx.keep_all_sites()
        ctx.henikoff()
        return ctx.weights_f64()


def format_value(x: float) -> str:
    """`round(x, 4)` of a numpy float64 inside an f-string (WeightedLD.py:283-284): rint(x*1e4)/1e4,
    printed with repr."""
    return repr(float(np.round(np.float64(x), 4)))


def ld_records(alignment: np.ndarray, weights, device: int = 0) -> np.ndarray:
    """The pair stage of WeightedLD.py:154-284 -> PAIR_DTYPE records in the program's print order
    (site_a ascending, then site_b), indices into `alignment`'s columns."""
    with _python_context(alignment, device) as ctx:
        ctx.keep_all_sites()
        ctx.set_weights(np.asarray(weights, np.float64))
        n, _ = ctx.ld_pairs(-math.inf)
        rec = ctx.fetch_pairs(n, L.FETCH_PARENT_INDEX | L.FETCH_UNORDERED)
    return rec[np.lexsort((rec["site_b"], rec["site_a"]))]


def ld(alignment: np.ndarray, weights, site_map, file=None, device: int = 0) -> None:
    """WeightedLD.py:154-284: prints `posa posb D D' R2` for every computed pair."""
    out = file or sys.stdout
    rec = ld_records(alignment, weights, device)
    site_map = np.asarray(site_map)
    lines = ["posa\tposb\tD\tD'\tR2"]
    for p in rec:
        lines.append(f"{site_map[p['site_a']]}\t{site_map[p['site_b']]}\t{format_value(p['d'])}\t"
                     f"{format_value(p['d_prime'])}\t{format_value(p['r2'])}")
    out.write("\n".join(lines) + "\n")


def handle_fasta(args):
    """WeightedLD.py:287-308."""
    alignment = read_fasta(args.file)
    _, var_sites_ld = compute_variable_sites(alignment, args.min_acgt, args.min_variability)
    return alignment[:, var_sites_ld], np.where(var_sites_ld)[0]


def handle_vcf(filename):
    """WeightedLD.py:311-379 for phased, diploid, GT-only multi-sample VCF -> (haplotype x site codes,
    POS).  Kept quirks: the last line is always dropped (:365, a trailing blank line is assumed, so a
    file without one loses its last variant), unphased `x/y` calls become missing (:353), '.' -> 4
    (:356), haplotypes come out in reverse column order (np.rot90, :375), no site filtering."""
    text = Path(filename).read_bytes()
    lines = text.split(b"\n")
    start = next((k for k, ln in enumerate(lines) if b"#CHROM" in ln), None)
    if start is None:
        print("No #CHROM header block identified")
        sys.exit(1)
    data = lines[start + 1:]
    if len(data[0].split(b"\t")) <= 12:
        print("The VCF data contains too small a population, are you sure this is a multi VCF?")
        sys.exit(1)
    data = data[:-1]
    pos = np.empty(len(data), np.int64)
    sites = []
    for k, ln in enumerate(data):
        head = ln.split(b"\t", 9)
        if len(head) < 10:
            raise IndexError("list index out of range")  # what WeightedLD.py:369 raises on a short row
        pos[k] = int(head[1])
        g = np.frombuffer(head[9], np.uint8)
        hap = None
        if g.size % 4 == 3:  # fast path: every call is exactly `a|b` / `a/b` with one-character alleles
            q = np.concatenate((g, [9])).reshape(-1, 4)
            sep = q[:, 1]
            if np.all(q[:, 3] == 9) and np.all((sep == ord("|")) | (sep == ord("/"))):
                a = np.where(q[:, 0] == ord("."), 4, q[:, 0] - 48)
                b = np.where(q[:, 2] == ord("."), 4, q[:, 2] - 48)
                ok = (a >= 0) & (a <= 9) & (b >= 0) & (b <= 9)
                if ok.all():
                    unph = sep == ord("/")
                    hap = np.stack((np.where(unph, 4, a), np.where(unph, 4, b)), axis=1).reshape(-1)
        if hap is None:  # general path
            vals = []
            for call in head[9].split(b"\t"):
                if re.fullmatch(rb"./.", call):
                    call = b".|."
                for allele in call.split(b"|"):
                    vals.append(4 if allele == b"." else int(allele))
            hap = np.asarray(vals)
        sites.append(hap.astype(np.uint8))
    if len({len(s) for s in sites}) > 1:
        raise ValueError("setting an array element with a sequence")  # np.array on ragged rows, :372
    aln = np.stack(sites) if sites else np.zeros((0, 0), np.uint8)   # sites x haplotypes
    return np.ascontiguousarray(np.rot90(aln)), pos


def main(args) -> None:
    """WeightedLD.py:382-402."""
    filename = str(args.file)
    if filename.endswith(".vcf"):
        alignment, site_map = handle_vcf(filename)
    else:
        alignment, site_map = handle_fasta(args)
    if args.unweighted:
        weights = np.ones(alignment.shape[0])
    else:
        weights = henikoff_weighting(alignment)
    ld(alignment, weights, site_map)


def build_parser() -> argparse.ArgumentParser:
    """WeightedLD.py:405-415."""
    parser = argparse.ArgumentParser(description="WeightedLD computation tool")
    parser.add_argument("--file", type=Path, help="The source file to load", required=True)
    parser.add_argument("--min-acgt", type=float, default=0.8,
                        help="Minimum fractions of ACTG at a given site for the site to be included in calculation.")
    parser.add_argument("--min-variability", type=float, default=0.02,
                        help="Minimum fraction of minor symbols for a site to be considered")
    parser.add_argument("--unweighted", action="store_true", default=False,
                        help="Use unit weights instead of Henikoff weights")
    return parser

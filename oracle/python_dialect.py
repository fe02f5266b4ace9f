"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's *Python* program (WeightedLD.py), the
dialect `wld_set_compat(ctx, WLD_COMPAT_PYTHON)` reproduces on the GPU.  Never imported by the product.

Parity is pinned: tests/test_oracle.py checks every function here against outputs of the unmodified
WeightedLD.py executed in the build container on all of its fixtures, on tests/t7_1000genome.vcf and on
a synthetic alignment with ambiguity codes (tests/golden/python_ref.json, made by tests/golden/make_golden.py).

Each function cites the lines of /root/reference/WeightedLD.py it follows ("py:N").
"""
from __future__ import annotations

import re

import numpy as np


def encode_text_fasta(text: str) -> np.ndarray:
    """py:21-41 via Bio.AlignIO: '>' starts a record, the other lines of a record are concatenated
    (no newline column, unlike lib.rs:297), lower-cased, a c g t - -> 0..4, anything else 5."""
    recs, cur = [], None
    for line in text.splitlines():
        if line.startswith(">"):
            cur = []
            recs.append(cur)
        elif cur is not None:
            cur.append(line.strip())
    rows = ["".join(r).lower() for r in recs]
    if len({len(r) for r in rows}) > 1:
        raise ValueError("Sequences must all be the same length")
    lut = np.full(256, 5, np.uint8)
    for ch, code in zip(b"acgt-", range(5)):
        lut[ch] = code
    chars = np.frombuffer("".join(rows).encode("latin-1"), np.uint8).reshape(len(rows), -1)
    return lut[chars]


def compute_variable_sites(alignment: np.ndarray, min_acgt: float, min_variability: float):
    """py:44-98 -> (hk mask, ld mask)."""
    n = alignment.shape[0]
    counts = np.stack([(alignment == x).sum(axis=0) for x in range(5)])            # py:71-72
    sufficient = (alignment < 4).sum(axis=0) / n > min_acgt                         # py:65-68
    major = counts.max(axis=0)                                                      # py:76
    minor = counts.sum(axis=0) - major                                              # py:77
    frac = np.zeros(alignment.shape[1])
    nz = minor > 0
    frac[nz] = minor[nz] / (major[nz] + minor[nz])                                  # py:80-84
    return sufficient, sufficient & (frac >= min_variability)                       # py:87-98


def henikoff_weighting(alignment: np.ndarray) -> np.ndarray:
    """py:101-151.  `unique_base` is ONE scalar (unique rows of the 5 x L count matrix, py:132)."""
    n_sites = alignment.shape[1]
    counts = np.stack([(alignment == x).sum(axis=0) for x in range(6)]).astype(np.float64)
    unique_base = len(np.unique(counts[:5], axis=0))                                # py:132
    ok = alignment != 5
    contrib = np.zeros(alignment.shape)
    per_cell = unique_base * counts[np.minimum(alignment, 5), np.arange(n_sites)]
    contrib[ok] = 1.0 / per_cell[ok]                                                # py:136-137
    avg = contrib.sum(axis=0) / counts[:5].sum(axis=0)                              # py:142-143
    contrib[~ok] = np.broadcast_to(avg, contrib.shape)[~ok]                         # py:144-145
    w = contrib.sum(axis=1)                                                         # py:148
    return w / w.max()                                                              # py:151


def _call(col: np.ndarray):
    """py:194-211: symbols by descending count, ties to the smaller code (argsort on <= 5 items is an
    insertion sort, hence stable).  Returns (major, minor) or None when fewer than two symbols remain."""
    sym, cnt = np.unique(col, return_counts=True)
    if len(sym) <= 1:
        return None
    order = np.argsort(-cnt, kind="stable")
    return int(sym[order[0]]), int(sym[order[1]])


def ld(alignment: np.ndarray, weights: np.ndarray, site_map) -> list[tuple[int, int, float, float, float]]:
    """py:154-284 -> [(posa, posb, D, D', R2)] in the program's print order, unrounded f64."""
    out = []
    weights = np.asarray(weights, np.float64)
    n_sites = alignment.shape[1]
    for i in range(n_sites - 1):
        for j in range(i + 1, n_sites):
            a, b = alignment[:, i], alignment[:, j]
            good = (a < 5) & (b < 5)                                                # py:183-186
            a, b, w = a[good], b[good], weights[good]
            ca, cb = _call(a), _call(b)
            if ca is None or cb is None:                                            # py:197-201,212
                continue
            keep = ((a == ca[0]) | (a == ca[1])) & ((b == cb[0]) | (b == cb[1]))   # py:214-222
            a, b, w = a[keep], b[keep], w[keep]
            am, bm = a == ca[0], b == cb[0]
            total = w.sum()
            with np.errstate(all="ignore"):
                PA, PB = w[am].sum() / total, w[bm].sum() / total                   # py:227-229
                Pa, Pb = w[~am].sum() / total, w[~bm].sum() / total                 # py:230-231
                if np.round(PA, 1) == 1.0 or np.round(PB, 1) == 1.0:                # py:234-237
                    continue
                o0, o3 = w[~am & ~bm].sum() / total, w[am & bm].sum() / total       # py:247-256
                o1, o2 = w[~am & bm].sum() / total, w[am & ~bm].sum() / total
                D = ((PA * PB - o3) + (Pa * Pb - o0) - (PA * Pb - o2) - (Pa * PB - o1)) / 4  # py:261-267
                if D < 0:                                                           # py:270-278
                    den = max(-o0, -o3)
                    if den == 0:
                        den = min(-o0, -o3)
                else:
                    den = min(o1, o2)
                    if den == 0:
                        den = max(o1, o2)
                out.append((int(site_map[i]), int(site_map[j]), float(D), float(D / den),
                            float(D * D / (PA * Pa * PB * Pb))))                    # py:279-284
    return out


def format_line(posa: int, posb: int, d: float, dp: float, r2: float) -> str:
    """py:283-284: round(x, 4) of numpy float64 scalars (rint(x*1e4)/1e4) printed with repr."""
    r = lambda x: repr(float(np.round(np.float64(x), 4)))
    return f"{posa}\t{posb}\t{r(d)}\t{r(dp)}\t{r(r2)}"


HEADER = "posa\tposb\tD\tD'\tR2"


def handle_vcf(text: str):
    """py:311-379 for well-formed phased diploid GT-only VCF text -> (haplotype x site codes, POS).
    Quirks kept: the LAST line is always dropped (py:365 assumes a trailing blank line), `x/y` calls
    become missing (py:353), '.' -> 4 (py:356), haplotype order is reversed by np.rot90 (py:375)."""
    lines = text.split("\n")
    start = next((k for k, ln in enumerate(lines) if "#CHROM" in ln), None)         # py:320-326
    if start is None:
        raise ValueError("No #CHROM header block identified")
    data = lines[start + 1:]
    if len(data[0].split("\t")) <= 12:                                               # py:333-337
        raise ValueError("The VCF data contains too small a population, are you sure this is a multi VCF?")
    data = data[:-1]                                                                 # py:365
    pos, rows = [], []
    for ln in data:
        f = ln.split("\t")
        pos.append(int(f[1]))
        hap = []
        for g in f[9:]:
            if re.fullmatch(r"./.", g):                                              # py:353
                g = ".|."
            for allele in g.split("|"):                                              # py:354
                hap.append(4 if allele == "." else int(allele))                      # py:356
        rows.append(hap)
    aln = np.array(rows, np.uint8)                                                   # sites x haplotypes
    return np.rot90(aln).copy(), np.array(pos, np.int64)                             # py:372-375

"""Deterministic synthetic alignments for tests and bench.py.

Sequences are mosaics of `founders` founder haplotypes with a recombination break every
`block` sites, so that site pairs inside a block are in real LD (a controllable fraction of all
pairs passes r2 > 0.1 and exercises the compaction) while pairs across blocks are not.  Per cell:
'-' with p~1.2 %, 'N' with p~1.2 %, a third allele with p~0.4 % (exercises the exclusion mask of
lib.rs:462-467).  `clonal=True` draws founders from a Zipf-like law, which makes cluster sizes very
uneven and the Henikoff weights span decades ("weight-heavy").  `variable_frac` < 1 interleaves
invariant / rare-variant columns that the default site filter must reject.
"""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"ACGT", np.uint8)


def make_alignment(n_seqs: int, n_cols: int, seed: int = 0xC0FFEE, founders: int = 64, block: int = 200,
                   variable_frac: float = 1.0, clonal: bool = False, newline_col: bool = False,
                   lowercase_frac: float = 0.0) -> np.ndarray:
    """Returns the (n_seqs, n_cols [+1]) uint8 character matrix (sequence-major, like a FASTA body)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n_seqs, n_cols + (1 if newline_col else 0)), np.uint8)
    if clonal:
        p = 1.0 / np.arange(1, founders + 1) ** 1.6
        p /= p.sum()
    else:
        p = None
    for c0 in range(0, n_cols, block):
        c1 = min(c0 + block, n_cols)
        w = c1 - c0
        maf = rng.uniform(0.02, 0.5, size=w)
        variable = rng.random(w) < variable_frac
        maf = np.where(variable, maf, np.where(rng.random(w) < 0.5, 0.0, 0.004))
        fm = rng.random((founders, w)) < maf[None, :]          # founder carries the minor allele
        fo = rng.choice(founders, size=n_seqs, p=p)             # founder of each sequence in this block
        minor = fm[fo]                                          # (n_seqs, w)
        # private mutations so that rare variants exist even with few founders
        minor ^= rng.random((n_seqs, w)) < (maf[None, :] * 0.02)
        maj_l = rng.integers(0, 4, size=w)
        min_l = (maj_l + rng.integers(1, 4, size=w)) % 4
        third_l = (min_l + 1 + (((min_l + 1) % 4) == maj_l)) % 4
        chars = np.where(minor, _LETTERS[min_l][None, :], _LETTERS[maj_l][None, :]).astype(np.uint8)
        noise = rng.integers(0, 256, size=(n_seqs, w), dtype=np.uint8)
        chars[noise < 3] = ord("-")
        chars[(noise >= 3) & (noise < 6)] = ord("N")
        third = noise == 6
        chars[third] = np.broadcast_to(_LETTERS[third_l][None, :], chars.shape)[third]
        if lowercase_frac > 0:
            low = rng.random((n_seqs, w)) < lowercase_frac
            chars[low] |= 0x20  # 'A'->'a'; '-' and 'N' -> '-' and 'n'
        out[:, c0:c1] = chars
    if newline_col:
        out[:, n_cols] = 10
    return out


def make_weights(n_seqs: int, seed: int = 7) -> np.ndarray:
    """U(0,1) weights, max-normalised: the reference micro-bench's distribution (bench.rs:40-42)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = rng.random(n_seqs).astype(np.float32)
    return (w / w.max()).astype(np.float32)

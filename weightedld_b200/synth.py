"""Deterministic synthetic alignments for tests and bench.py.

Sequences are mosaics of `founders` founder haplotypes with a recombination break every
`block` sites, so that site pairs inside a block are in real LD (a controllable fraction of all
pairs passes r2 > 0.1 and exercises the compaction) while pairs across blocks are not.  Per cell
noise: '-' (gap_rate), 'N' (n_rate), a third allele (third_rate; exercises the exclusion mask of
lib.rs:462-467).  `clonal=True` puts sequences into global clusters of Zipf-distributed size that
follow one founder per block, which makes the Henikoff weights very uneven ("weight-heavy").
`variable_frac` < 1 interleaves invariant / rare-variant columns that the default site filter
rejects.  `end_runs` gives that fraction of sequences runs of 'N' at both ends (SARS-CoV-2-like).
"""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"ACGT", np.uint8)


def make_alignment(n_seqs: int, n_cols: int, seed: int = 0xC0FFEE, founders: int = 64, block: int = 200,
                   variable_frac: float = 1.0, clonal: bool = False, newline_col: bool = False,
                   lowercase_frac: float = 0.0, gap_rate: float = 3 / 256, n_rate: float = 3 / 256,
                   third_rate: float = 1 / 256, stray: float = 0.1, private_rate: float = 0.02,
                   end_runs: float = 0.0) -> np.ndarray:
    """Returns the (n_seqs, n_cols [+1]) uint8 character matrix (sequence-major, like a FASTA body)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n_seqs, n_cols + (1 if newline_col else 0)), np.uint8)
    if clonal:
        p = 1.0 / np.arange(1, founders + 1) ** 1.6
        cluster = rng.choice(founders, size=n_seqs, p=p / p.sum())
    t_gap = int(round(gap_rate * 65536))
    t_n = t_gap + int(round(n_rate * 65536))
    t_third = t_n + int(round(third_rate * 65536))
    for c0 in range(0, n_cols, block):
        c1 = min(c0 + block, n_cols)
        w = c1 - c0
        maf = rng.uniform(0.02, 0.5, size=w)
        variable = rng.random(w) < variable_frac
        maf = np.where(variable, maf, np.where(rng.random(w) < 0.5, 0.0, 0.004))
        fm = rng.random((founders, w)) < maf[None, :]          # founder carries the minor allele
        if clonal:
            fo = rng.permutation(founders)[cluster]
            fo = np.where(rng.random(n_seqs) < stray, rng.integers(0, founders, size=n_seqs), fo)
        else:
            fo = rng.integers(0, founders, size=n_seqs)        # founder of each sequence in this block
        minor = fm[fo]                                          # (n_seqs, w)
        if private_rate > 0:  # private mutations so that rare variants exist even with few founders
            minor ^= rng.random((n_seqs, w)) < (maf[None, :] * private_rate)
        maj_l = rng.integers(0, 4, size=w)
        min_l = (maj_l + rng.integers(1, 4, size=w)) % 4
        third_l = (min_l + 1 + (((min_l + 1) % 4) == maj_l)) % 4
        chars = np.where(minor, _LETTERS[min_l][None, :], _LETTERS[maj_l][None, :]).astype(np.uint8)
        noise = rng.integers(0, 65536, size=(n_seqs, w), dtype=np.uint16)
        chars[noise < t_gap] = ord("-")
        chars[(noise >= t_gap) & (noise < t_n)] = ord("N")
        third = (noise >= t_n) & (noise < t_third)
        chars[third] = np.broadcast_to(_LETTERS[third_l][None, :], chars.shape)[third]
        if lowercase_frac > 0:
            low = rng.random((n_seqs, w)) < lowercase_frac
            chars[low] |= 0x20  # 'A'->'a'; '-' stays '-', 'N' -> 'n'
        out[:, c0:c1] = chars
    if end_runs > 0:
        rows = np.flatnonzero(rng.random(n_seqs) < end_runs)
        for r in rows:
            a, b = rng.integers(1, max(2, n_cols // 50), size=2)
            out[r, :a] = ord("N")
            out[r, n_cols - b:n_cols] = ord("N")
    if newline_col:
        out[:, n_cols] = 10
    return out


def make_sarscov2_like(n_seqs: int, n_cols: int, seed: int = 0xC0FFEE + 3) -> np.ndarray:
    """Config 4 of BASELINE.json: clonal, low-noise, about one third of the columns variable, runs
    of N at both ends of ~2 % of the sequences; Henikoff weights span more than two decades."""
    return make_alignment(n_seqs, n_cols, seed=seed, founders=256, block=300, variable_frac=1 / 3, clonal=True,
                          gap_rate=2e-4, n_rate=2e-4, third_rate=1e-4, stray=0.02, private_rate=0.002,
                          end_runs=0.02)


def make_weights(n_seqs: int, seed: int = 7) -> np.ndarray:
    """U(0,1) weights, max-normalised: the reference micro-bench's distribution (bench.rs:40-42)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = rng.random(n_seqs).astype(np.float32)
    return (w / w.max()).astype(np.float32)

"""Regenerates tests/golden/*.json|*.npz from the reference checkout.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Two kinds of vectors are written:
  * rust_kat.json      — the known-answer values of the reference's Rust unit tests
                         (rust/weighted_ld/src/lib.rs:692-801), transcribed by hand with the
                         line each comes from.  They pin the oracle's Rust semantics.
  * python_ref.json    — outputs of the reference's executable Python implementation
                         (WeightedLD.py) on its own fixtures, produced by IMPORTING it here
                         (Bio.AlignIO is stubbed, np.bool8 aliased, and the numpy-2 overflow
                         at WeightedLD.py:372 avoided by dropping the POS column first —
                         SURVEY.md §8c).  Python and Rust differ numerically, so
                         these pin the Python-compat semantics and the cases where both agree.
  * fixtures.json      — the character matrices of the reference's FASTA fixtures (tiny test
                         DATA, so that the GPU box, which has no /root/reference, can run them).
  * t7_haplotypes.npz  — the 5008 x 6 allele matrix + POS of tests/t7_1000genome.vcf.
  * t7_1000genome.vcf.gz — that fixture itself (test DATA), so the VCF readers can be tested on the GPU box.
"""
import io
import json
import re
import sys
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _stub_bio():
    class _Rec:
        def __init__(self, s):
            self.seq = s

    class _Aln(list):
        def get_alignment_length(self):
            return len(self[0].seq)

    def read(filename, fmt):
        recs, cur = [], None
        for line in open(filename):
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                cur = []
                recs.append(cur)
            elif cur is not None:
                cur.append(line)
        return _Aln(_Rec("".join(r)) for r in recs)

    bio = types.ModuleType("Bio")
    alignio = types.ModuleType("Bio.AlignIO")
    alignio.read = read
    bio.AlignIO = alignio
    sys.modules["Bio"] = bio
    sys.modules["Bio.AlignIO"] = alignio


def main():
    _stub_bio()
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    sys.path.insert(0, str(REF))
    import WeightedLD as wld  # the reference, unmodified

    fixtures, pyref = {}, {}
    for f in sorted((REF / "tests").glob("*.fasta")):
        text = f.read_bytes().decode()
        fixtures[f.stem] = text
        aln = wld.read_fasta(str(f))
        hk, ld = wld.compute_variable_sites(aln, 0.8, 0.02)
        sub = aln[:, ld]
        site_map = np.where(ld)[0]
        w = wld.henikoff_weighting(sub)
        buf = io.StringIO()
        with redirect_stdout(buf):
            wld.ld(sub, w, site_map)
        pyref[f.stem] = {
            "codes_sum": int(aln.sum()),
            "var_sites_hk": hk.tolist(),
            "var_sites_ld": ld.tolist(),
            "weights_on_ld_sites": w.tolist(),
            "weights_on_hk_sites": wld.henikoff_weighting(aln[:, hk]).tolist(),
            "ld_stdout": buf.getvalue().splitlines(),
        }

    # VCF (WeightedLD.py:311-379) with the numpy-2 fix: drop POS before the uint8 conversion.
    src = (REF / "WeightedLD.py").read_text()
    src = src.replace(
        "    alignment = np.array(data, dtype=np.uint8)\n    alignment = np.delete(alignment, 0, axis=1)\n",
        "    alignment = np.array([row[1:] for row in data], dtype=np.uint8)\n")
    mod = types.ModuleType("WeightedLD_np2")
    exec(compile(src, "WeightedLD_np2", "exec"), mod.__dict__)
    aln, site_map = mod.handle_vcf(str(REF / "tests" / "t7_1000genome.vcf"))
    w = mod.henikoff_weighting(aln)
    buf = io.StringIO()
    with redirect_stdout(buf):
        mod.ld(aln, w, site_map)
    pyref["t7_1000genome"] = {
        "shape": list(aln.shape),
        "site_map": site_map.tolist(),
        "weights_mean": float(w.mean()),
        "weights_min": float(w.min()),
        "weights_max": float(w.max()),
        "weights_distinct": sorted(set(np.round(w, 12).tolist())),
        "ld_stdout": buf.getvalue().splitlines(),
    }
    # Synthetic alignment WITH ambiguity codes, near-ties between alleles and rare minors: pins the
    # per-pair allele calls of WeightedLD.py:183-211 (none of the reference's own fixtures has code 5
    # inside an LD site) and the PA/PB skip of WeightedLD.py:234-237.
    rng = np.random.Generator(np.random.PCG64(20260118))
    n_seq, n_site = 60, 26
    founders = rng.integers(0, 2, size=(6, n_site))
    owner = rng.integers(0, 6, size=n_seq)
    syn = np.empty((n_seq, n_site), np.uint8)
    for j in range(n_site):
        maj, mnr, third = rng.permutation(4)[:3]
        col = np.where(founders[owner, j] == 0, maj, mnr)
        flip = rng.random(n_seq) < 0.15
        col = np.where(flip, np.where(col == maj, mnr, maj), col)
        col = np.where(rng.random(n_seq) < 0.06, third, col)
        col = np.where(rng.random(n_seq) < 0.05, 4, col)
        col = np.where(rng.random(n_seq) < (0.25 if j % 3 == 0 else 0.04), 5, col)
        syn[:, j] = col
    syn[:, 5] = np.where(np.arange(n_seq) < 58, 2, 1)          # rare minor: PA rounds to 1.0 for most weights
    syn[:30, 7], syn[30:, 7] = 0, 3                              # exact tie between two alleles
    syn[[3, 33], 7] = 5
    hk, ldm = wld.compute_variable_sites(syn, 0.5, 0.02)
    w_syn = wld.henikoff_weighting(syn)
    buf = io.StringIO()
    with redirect_stdout(buf):
        wld.ld(syn, w_syn, np.arange(n_site))
    sub = syn[:, ldm]
    w_sub = wld.henikoff_weighting(sub)
    buf2 = io.StringIO()
    with redirect_stdout(buf2):
        wld.ld(sub, w_sub, np.where(ldm)[0])
    pyref["synthetic_ambiguous"] = {
        "alignment": ["".join(map(str, r)) for r in syn.tolist()],
        "min_acgt": 0.5, "min_variability": 0.02,
        "var_sites_hk": hk.tolist(), "var_sites_ld": ldm.tolist(),
        "weights_all_sites": w_syn.tolist(), "ld_stdout_all_sites": buf.getvalue().splitlines(),
        "weights_on_ld_sites": w_sub.tolist(), "ld_stdout": buf2.getvalue().splitlines(),
    }
    import gzip
    (OUT / "t7_1000genome.vcf.gz").write_bytes(gzip.compress((REF / "tests" / "t7_1000genome.vcf").read_bytes(), mtime=0))

    # all six variant rows (the Python reader drops the last, WeightedLD.py:365)
    rows, pos = [], []
    for line in (REF / "tests" / "t7_1000genome.vcf").read_text().split("\n"):
        if not line or line.startswith("#"):
            continue
        t = line.split("\t")
        pos.append(int(t[1]))
        hap = []
        for g in t[9:]:
            a, b = re.split(r"[|/]", g)[:2]
            hap += [4 if a == "." else int(a), 4 if b == "." else int(b)]
        rows.append(hap)
    np.savez_compressed(OUT / "t7_haplotypes.npz", alleles=np.array(rows, np.uint8).T.copy(),
                        pos=np.array(pos, np.int64), python_alignment=aln.astype(np.uint8),
                        python_weights=w)

    rust_kat = {
        "histogram": {"src": "lib.rs:692-703", "symbols": "AAACCGTTTT--?", "hist": [3, 2, 1, 4, 2, 1]},
        "major_minor": {"src": "lib.rs:705-728", "cases": [
            {"hist": [0, 1, 10, 2, 0, 0], "major": 2, "minor": 3},
            {"hist": [1, 9, 10, 2, 0, 0], "major": 2, "minor": 1},
            {"hist": [1, 1, 40, 2, 4, 0], "major": 2, "minor": 4}]},
        "henikoff": [
            {"src": "lib.rs:731-735", "rows": ["AAAAA", "AAAAA", "CCCCC", "CCCCC", "TTTTT"],
             "weights": [0.5, 0.5, 0.5, 0.5, 1.0], "tol": "ulps"},
            {"src": "lib.rs:738-742", "rows": ["GCGTTAGC", "GAGTTGGA", "CGGACTAA"],
             "weights": [0.769, 0.692, 1.0], "tol": 1e-3},
            {"src": "lib.rs:745-750", "rows": ["AAGA", "AA-A", "GGGG", "GGGG"],
             "weights": [0.733, 1.0, 0.733, 0.733], "tol": 1e-3}],
        "pair": [
            {"src": "lib.rs:753-767", "a": "AAAATTTT", "b": "TTAAAATT", "w": [1.0] * 8,
             "d": 0.0, "d_prime": 0.0, "r2": 0.0, "tol": 1e-5},
            {"src": "lib.rs:770-784", "a": "AAAATTTT", "b": "TTTTAAAA", "w": [1.0] * 8,
             "d": 0.25, "d_prime": 0.5, "r2": 1.0, "tol": 1e-5},
            {"src": "lib.rs:787-801", "a": "AAAACAC", "b": "AAAGTAA", "w": [1.0, 1.0, 0.4, 0.2, 0.5, 0.8, 0.2],
             "d": 0.00308, "d_prime": 0.05555, "r2": 0.00346, "tol": 1e-5}],
    }
    (OUT / "fixtures.json").write_text(json.dumps(fixtures, indent=1))
    (OUT / "python_ref.json").write_text(json.dumps(pyref, indent=1))
    (OUT / "rust_kat.json").write_text(json.dumps(rust_kat, indent=1))
    print("wrote", [p.name for p in OUT.iterdir()])


if __name__ == "__main__":
    main()

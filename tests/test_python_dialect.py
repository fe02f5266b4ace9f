"""The Python dialect (WeightedLD.py) — SURVEY.md §8f rank 3, BASELINE config 2.

CPU part (not gpu): pins oracle/python_dialect.py against outputs of the unmodified WeightedLD.py
(tests/golden/python_ref.json, generated in the build container by tests/golden/make_golden.py) on every
fixture of the reference, on tests/t7_1000genome.vcf and on a synthetic alignment with ambiguity codes;
and checks the host-side VCF reader.
GPU part: the product in WLD_COMPAT_PYTHON mode through the C ABI against the same golden output — the
printed lines must be IDENTICAL on the reference's fixtures — and against the pinned oracle on seeded
alignments (|delta| <= 1e-6 on D and R2: 24-bit fixed-point weights + f32 records).
"""
import gzip
import io
import math
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import GOLDEN


def py_dialect():
    from oracle import python_dialect as P
    return P


def synthetic_ambiguous(golden):
    d = golden["python_ref"]["synthetic_ambiguous"]
    return d, np.array([[int(c) for c in r] for r in d["alignment"]], np.uint8)


def t7_text():
    return gzip.decompress((GOLDEN / "t7_1000genome.vcf.gz").read_bytes()).decode()


# ------------------------------------------------------------------------------------------ CPU
def test_oracle_pinned_on_reference_fixtures(golden):
    P = py_dialect()
    for name, text in golden["fixtures"].items():
        ref = golden["python_ref"][name]
        aln = P.encode_text_fasta(text)
        assert int(aln.sum()) == ref["codes_sum"]                       # test.py:13-17
        hk, ldm = P.compute_variable_sites(aln, 0.8, 0.02)
        assert hk.tolist() == ref["var_sites_hk"] and ldm.tolist() == ref["var_sites_ld"], name
        assert np.allclose(P.henikoff_weighting(aln[:, hk]), ref["weights_on_hk_sites"], rtol=1e-13)
        sub = aln[:, ldm]
        w = P.henikoff_weighting(sub)
        assert np.allclose(w, ref["weights_on_ld_sites"], rtol=1e-13), name
        lines = [P.HEADER] + [P.format_line(*r) for r in P.ld(sub, w, np.where(ldm)[0])]
        assert lines == ref["ld_stdout"], name


def test_oracle_pinned_on_ambiguity_codes(golden):
    """Per-pair allele calls after deleting code-5 sequences (WeightedLD.py:183-211) and the PA/PB
    skip (WeightedLD.py:234-237): 45 of the 650 calls differ from the per-site call on this input."""
    P = py_dialect()
    d, aln = synthetic_ambiguous(golden)
    hk, ldm = P.compute_variable_sites(aln, d["min_acgt"], d["min_variability"])
    assert hk.tolist() == d["var_sites_hk"] and ldm.tolist() == d["var_sites_ld"]
    w = P.henikoff_weighting(aln)
    assert np.allclose(w, d["weights_all_sites"], rtol=1e-13)
    lines = [P.HEADER] + [P.format_line(*r) for r in P.ld(aln, w, np.arange(aln.shape[1]))]
    assert lines == d["ld_stdout_all_sites"]
    assert len(lines) - 1 < aln.shape[1] * (aln.shape[1] - 1) // 2  # some pairs are skipped


def test_oracle_pinned_on_t7_vcf(golden):
    P = py_dialect()
    ref = golden["python_ref"]["t7_1000genome"]
    aln, pos = P.handle_vcf(t7_text())
    assert list(aln.shape) == ref["shape"] and pos.tolist() == ref["site_map"]
    assert np.array_equal(aln, golden["t7"]["python_alignment"])
    w = P.henikoff_weighting(aln)
    assert np.allclose(w, golden["t7"]["python_weights"], rtol=1e-13)
    assert round(float(w.mean()), 3) == 0.002                           # the orphaned test.py:152-159
    lines = [P.HEADER] + [P.format_line(*r) for r in P.ld(aln, w, pos)]
    assert lines == ref["ld_stdout"]


def test_host_vcf_reader_matches_reference_reader(golden, tmp_path):
    from weightedld_b200 import pycompat
    f = tmp_path / "t7.vcf"
    f.write_text(t7_text())
    aln, pos = pycompat.handle_vcf(f)
    assert np.array_equal(aln, golden["t7"]["python_alignment"]) and aln.dtype == np.uint8
    assert pos.tolist() == golden["python_ref"]["t7_1000genome"]["site_map"]
    # with the trailing newline the reference assumes, all six variants survive
    f.write_text(t7_text() + "\n")
    aln6, pos6 = pycompat.handle_vcf(f)
    assert aln6.shape == (5008, 6) and np.array_equal(aln6[:, :5], aln)
    assert np.array_equal(aln6[::-1], golden["t7"]["alleles"])          # np.rot90 reverses the haplotypes
    # unphased and missing calls, multi-character fields -> general path
    head = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{k}" for k in range(4))
    body = ["1\t100\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1/0\t.|1\t0|0", "1\t200\t.\tA\tC,G\t.\t.\t.\tGT\t2|1\t0|0\t1|1\t0|10", ""]
    f.write_text(head + "\n" + "\n".join(body))
    aln, pos = pycompat.handle_vcf(f)
    assert pos.tolist() == [100, 200]
    assert aln[::-1].T.tolist() == [[0, 1, 4, 4, 4, 1, 0, 0], [2, 1, 0, 0, 1, 1, 0, 10]]
    P = py_dialect()
    a2, p2 = P.handle_vcf(f.read_text())
    assert np.array_equal(a2, aln) and np.array_equal(p2, pos)


def test_host_fasta_reader_python_semantics(golden, tmp_path):
    from weightedld_b200 import pycompat
    P = py_dialect()
    for name, text in golden["fixtures"].items():
        f = tmp_path / "x.fasta"
        f.write_text(text)
        assert np.array_equal(pycompat.read_fasta(f), P.encode_text_fasta(text)), name
    f.write_text(">a\nAC\nGT\n>b\nac-n\n")  # multi-line record (Bio.AlignIO), lower case
    assert pycompat.read_fasta(f).tolist() == [[0, 1, 2, 3], [0, 1, 4, 5]]


def test_python_value_format():
    from weightedld_b200.pycompat import format_value
    assert [format_value(x) for x in (-0.25, 1.0, 0.0, 0.10285, 0.00005, 12345.678949, -0.0)] == \
        ["-0.25", "1.0", "0.0", "0.1028", "0.0", "12345.6789", "-0.0"]


# ------------------------------------------------------------------------------------------ GPU
def gpu_lines(aln, weights, site_map):
    from weightedld_b200 import pycompat
    buf = io.StringIO()
    pycompat.ld(aln, weights, site_map, file=buf)
    return buf.getvalue().splitlines()


@pytest.mark.gpu
def test_gpu_python_mode_reproduces_reference_stdout_on_fixtures(golden, tmp_path):
    """Every FASTA fixture of the reference through the mirrored Python program on the GPU: site masks
    identical, weights to 1e-9, printed LD table identical to WeightedLD.py's stdout."""
    from weightedld_b200 import pycompat
    for name, text in golden["fixtures"].items():
        ref = golden["python_ref"][name]
        f = tmp_path / f"{name}.fasta"
        f.write_text(text)
        aln = pycompat.read_fasta(f)
        hk, ldm = pycompat.compute_variable_sites(aln, 0.8, 0.02)
        assert hk.tolist() == ref["var_sites_hk"] and ldm.tolist() == ref["var_sites_ld"], name
        sub, site_map = pycompat.handle_fasta(SimpleNamespace(file=f, min_acgt=0.8, min_variability=0.02))
        w = pycompat.henikoff_weighting(sub)
        assert np.allclose(w, ref["weights_on_ld_sites"], rtol=1e-9, atol=0), name
        assert gpu_lines(sub, w, site_map) == ref["ld_stdout"], name


@pytest.mark.gpu
def test_gpu_python_mode_t7_vcf_end_to_end(golden, tmp_path, capsys):
    """BASELINE config 2: tests/t7_1000genome.vcf -> VCF reader -> Python-dialect Henikoff weights ->
    pair stage on the tensor cores -> the reference's stdout, line for line."""
    from weightedld_b200 import pycompat
    f = tmp_path / "t7_1000genome.vcf"
    f.write_text(t7_text())
    pycompat.main(pycompat.build_parser().parse_args(["--file", str(f)]))
    assert capsys.readouterr().out.splitlines() == golden["python_ref"]["t7_1000genome"]["ld_stdout"]
    aln, _ = pycompat.handle_vcf(f)
    w = pycompat.henikoff_weighting(aln)
    assert np.allclose(w, golden["t7"]["python_weights"], rtol=1e-9, atol=0)
    assert round(float(w.mean()), 3) == 0.002                           # the orphaned test.py:152-159


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["i8", "bf16", "simt"])
def test_gpu_python_mode_ambiguity_codes(golden, kernel):
    """Per-pair allele calls: pairs the Gram kernels cannot decide per site are recomputed by the
    per-pair kernel (pair_python.cu); the merged table equals WeightedLD.py's stdout."""
    import weightedld_b200 as wld
    d, aln = synthetic_ambiguous(golden)
    with wld.Context(0) as ctx:
        ctx.set_compat("python")
        ctx.set_pair_kernel(kernel)
        ctx.load_alignment(aln, codes=True)
        assert ctx.filter_sites_python(d["min_acgt"], d["min_variability"]) == sum(d["var_sites_ld"])
        ctx.henikoff()
        w = ctx.weights_f64()
        assert np.allclose(w, d["weights_on_ld_sites"], rtol=1e-9, atol=0)
        n, done = ctx.ld_pairs(-math.inf)
        rec = ctx.fetch_pairs(n)
    assert done == 26 * 25 // 2
    rec = rec[np.lexsort((rec["site_b"], rec["site_a"]))]
    from weightedld_b200.pycompat import format_value
    lines = ["posa\tposb\tD\tD'\tR2"] + [
        f"{p['site_a']}\t{p['site_b']}\t{format_value(p['d'])}\t{format_value(p['d_prime'])}\t{format_value(p['r2'])}"
        for p in rec]
    assert lines == d["ld_stdout"]


@pytest.mark.gpu
@pytest.mark.parametrize("n_seqs,n_sites,seed", [(200, 90, 1), (1500, 300, 2)])
def test_gpu_python_mode_matches_pinned_oracle_on_seeded_input(n_seqs, n_sites, seed):
    """Larger seeded alignments with code 5, gaps and third alleles against the pinned Python-dialect
    oracle: same pair set; |delta D|, |delta R2| <= 1e-6; D' to 1e-5 relative."""
    import weightedld_b200 as wld
    from weightedld_b200 import pycompat
    from weightedld_b200.synth import make_alignment
    P = py_dialect()
    chars = make_alignment(n_seqs, n_sites, seed=seed, block=40, n_rate=0.03, gap_rate=0.02, third_rate=0.02)
    aln = P.encode_text_fasta("".join(">s\n" + bytes(r).decode() + "\n" for r in chars))
    _, ldm = P.compute_variable_sites(aln, 0.8, 0.02)
    hk_gpu, ld_gpu = pycompat.compute_variable_sites(aln, 0.8, 0.02)
    assert np.array_equal(ld_gpu, ldm)
    sub = np.ascontiguousarray(aln[:, ldm])
    w_ref = P.henikoff_weighting(sub)
    w = pycompat.henikoff_weighting(sub)
    assert np.allclose(w, w_ref, rtol=1e-9, atol=0)
    ref = P.ld(sub, w_ref, np.arange(sub.shape[1]))
    rec = pycompat.ld_records(sub, w)
    assert [(int(p["site_a"]), int(p["site_b"])) for p in rec] == [(r[0], r[1]) for r in ref]
    r = np.array([r[2:] for r in ref])
    assert np.max(np.abs(rec["d"] - r[:, 0])) <= 1e-6 and np.max(np.abs(rec["r2"] - r[:, 2])) <= 1e-6
    assert np.all(np.abs(rec["d_prime"] - r[:, 1]) <= 1e-5 * np.maximum(1.0, np.abs(r[:, 1])))


# ------------------------------------------------------------------------------------------ C++ CLI
def _cli():
    from conftest import ROOT
    return ROOT / "weightedld_b200" / "weighted_ld"


def test_cli_value_formats_match_python_and_rust_rules():
    """Host-only hook of the C++ binary: repr(round(float64(v), 4)) (WeightedLD.py:283) and `{:.3}`."""
    import subprocess
    from weightedld_b200.api import format_f3
    from weightedld_b200.pycompat import format_value
    rng = np.random.Generator(np.random.PCG64(7))
    vals = [-0.25, 1.0, 0.0, -0.0, 0.00005, 0.00015, 2.5e-5, 0.99995, 1e17, 12345.678949, 0.0625, 0.1875] + \
        rng.normal(0, 1, 40).tolist() + (10.0 ** rng.uniform(-6, 6, 20)).tolist()
    for v in vals:
        out = subprocess.run([str(_cli()), "--format-py4", repr(v)], capture_output=True, text=True).stdout.split()
        assert out == [format_value(v), format_f3(np.float32(v))], (v, out)


@pytest.mark.gpu
def test_cli_python_compat_vcf_and_fasta(golden, tmp_path):
    """`weighted_ld --python-compat` on the reference's VCF and FASTA fixtures prints WeightedLD.py's table."""
    import subprocess
    f = tmp_path / "t7_1000genome.vcf"
    f.write_text(t7_text())
    r = subprocess.run([str(_cli()), "--vcf-input", str(f), "--python-compat", "--pair-output", str(tmp_path / "p.tsv"),
                        "--weights-output", str(tmp_path / "w.tsv")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "p.tsv").read_text().splitlines() == golden["python_ref"]["t7_1000genome"]["ld_stdout"]
    w = np.array([float(x.split("\t")[1]) for x in (tmp_path / "w.tsv").read_text().splitlines()[1:]])
    assert len(w) == 5008 and abs(w.mean() - 0.002) < 5e-4
    for name, text in golden["fixtures"].items():
        g = tmp_path / f"{name}.fasta"
        g.write_text(text)
        r = subprocess.run([str(_cli()), "--fasta-input", str(g), "--python-compat", "--pair-output", "-"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.splitlines() == golden["python_ref"][name]["ld_stdout"], name


@pytest.mark.gpu
def test_cli_vcf_input_rust_dialect(golden, tmp_path, oracle):
    """--vcf-input without --python-compat: the VCF reader feeds the normative (Rust) pipeline; sites are
    labelled by POS.  Checked against the oracle on the same allele matrix."""
    import subprocess
    f = tmp_path / "t7.vcf"
    f.write_text(t7_text() + "\n")  # trailing newline: all six variants
    r = subprocess.run([str(_cli()), "--vcf-input", str(f), "--pair-output", str(tmp_path / "p.tsv"), "--r2-threshold=-1",
                        "--min-minor", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    codes = np.ascontiguousarray(golden["t7"]["alleles"][::-1].T)  # site-major, haplotypes reversed like np.rot90
    fs = oracle.filter_sites(oracle.siteset_from_codes(codes), 0.8, 0.0, 0.5)
    w = oracle.henikoff_weights(fs, f64=True)
    w32 = w.astype(np.float32)
    pairs, _ = oracle.all_weighted_ld_pairs(fs, oracle.quantize_weights(w32, *oracle.auto_quant_params(w32)), -1.0, oracle.F64)
    pos = golden["t7"]["pos"]
    want = ["site_a\tsite_b\td\td'\tr2"] + [
        f"{pos[p['a']]}\t{pos[p['b']]}\t{oracle.format_f3(p['d'])}\t{oracle.format_f3(p['d_prime'])}\t{oracle.format_f3(p['r2'])}"
        for p in pairs]
    assert (tmp_path / "p.tsv").read_text().splitlines() == want and len(want) > 5


# ------------------------------------------------------------------------------------------ C++ readers (host only)
def _fnv1a(data: bytes) -> int:
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _parse_only(*args):
    import subprocess
    r = subprocess.run([str(_cli()), "--parse-only", *args], capture_output=True, text=True)
    return r.returncode, r.stdout.split(), r.stderr


def test_cpp_readers_match_python_readers(golden, tmp_path):
    """`weighted_ld --parse-only` (no GPU): the C++ FASTA (Rust and Biopython semantics) and VCF readers deliver
    byte for byte what the Python host readers / the oracle deliver."""
    import weightedld_b200 as wld
    from weightedld_b200 import pycompat
    for name, text in golden["fixtures"].items():
        f = tmp_path / f"{name}.fasta"
        f.write_text(text)
        codes = pycompat.read_fasta(f)
        rc, out, _ = _parse_only("--fasta-input", str(f), "--python-compat")
        assert rc == 0 and out[:4] == [str(codes.shape[0]), str(codes.shape[1]), f"{_fnv1a(codes.tobytes()):016x}", "codes"], name
        rc, out, err = _parse_only("--fasta-input", str(f))
        if name.startswith("t1_"):
            assert rc == 101 and "Not all sequences have the same number of symbols" in err   # lib.rs:180-182
        else:
            chars = wld.read_fasta(f).chars
            assert rc == 0 and out[:4] == [str(chars.shape[0]), str(chars.shape[1]), f"{_fnv1a(chars.tobytes()):016x}", "ascii"]
    f = tmp_path / "t7.vcf"
    for text in (t7_text(), t7_text() + "\n"):
        f.write_text(text)
        aln, pos = pycompat.handle_vcf(f)
        rc, out, _ = _parse_only("--vcf-input", str(f))
        assert rc == 0 and out == [str(aln.shape[0]), str(aln.shape[1]), f"{_fnv1a(aln.tobytes()):016x}", "codes",
                                   f"{_fnv1a(pos.astype('<i8').tobytes()):016x}"]
    head = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{k}" for k in range(4))
    f.write_text(head + "\n1\t100\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1/0\t.|1\t0|0\n1\t200\t.\tA\tC,G\t.\t.\t.\tGT\t2|1\t0|0\t1|1\t0|10\n")
    aln, pos = pycompat.handle_vcf(f)
    rc, out, _ = _parse_only("--vcf-input", str(f))
    assert rc == 0 and out[2] == f"{_fnv1a(aln.tobytes()):016x}" and out[:2] == ["8", "2"]
    f.write_text("no header here\n")
    rc, out, err = _parse_only("--vcf-input", str(f))
    assert rc == 1 and "No #CHROM header block identified" in err

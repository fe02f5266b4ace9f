"""GPU test of the drop-in CLI (weightedld_b200/weighted_ld, C++ host over the C ABI): same flags
and TSV outputs as the reference binary (main.rs:14-213) on the reference's own fixtures."""
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = ROOT / "weightedld_b200" / "weighted_ld"

EXPECTED = {  # emulated Rust outputs, SURVEY.md §8c
    "example": (["0\t1\t0.107\t0.345\t0.237"], "1.000 0.300 0.300 0.300 0.700 0.200 0.200 0.200 0.200 0.200"),
    "t4_weights1_ld0": (["0\t3\t0.088\t0.422\t0.192", "1\t3\t0.088\t0.422\t0.192"], None),
    "t5_weights1_ld0.25": (["0\t1\t-0.250\t0.500\t1.000"], None),
}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_cli_fixture(tmp_path, golden, name):
    f = tmp_path / "in.fasta"
    f.write_bytes(golden["fixtures"][name].encode())
    r = subprocess.run([str(BIN), "--fasta-input", str(f), "--pair-output", str(tmp_path / "p.tsv"),
                        "--weights-output", str(tmp_path / "w.tsv")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines, weights = EXPECTED[name]
    assert (tmp_path / "p.tsv").read_text().splitlines() == ["site_a\tsite_b\td\td'\tr2"] + lines
    wl = (tmp_path / "w.tsv").read_text().splitlines()
    assert wl[0] == "Sequence_index\thk_weight"
    if weights:
        assert [x.split("\t")[1] for x in wl[1:]] == weights.split()
    for needle in ("Loaded fasta file in", "sequences,", "Found", "sites of interest", "Computed Henikoff weights",
                   "Beginning pairwise weighted LD computation", "passed threshold", "Finshed writing output"):
        assert needle in r.stderr  # main.rs:131-210 log lines (env_logger writes to stderr)


def test_cli_unweighted_and_flags(tmp_path, golden, oracle):
    f = tmp_path / "in.fasta"
    f.write_bytes(golden["fixtures"]["t4_weights1_ld0"].encode())
    r = subprocess.run([str(BIN), "--fasta-input", str(f), "--pair-output", str(tmp_path / "p.tsv"), "--unweighted",
                        "--r2-threshold=-1", "--min-acgt", "0.5", "--min-minor", "0.1", "--max-minor", "0.5"],
                       capture_output=True, text=True, env={"RUST_LOG": "error"})
    assert r.returncode == 0 and r.stderr == ""
    from conftest import fasta_chars
    fs, w, pairs = oracle.run_pipeline(fasta_chars(golden["fixtures"]["t4_weights1_ld0"]), 0.5, 0.1, 0.5, -1.0,
                                       unweighted=True, flavour=oracle.F64)
    oracle.write_pairs(tmp_path / "o.tsv", pairs)
    assert (tmp_path / "p.tsv").read_text() == (tmp_path / "o.tsv").read_text()


def test_cli_errors_like_the_reference(tmp_path, golden):
    f = tmp_path / "t1.fasta"
    f.write_bytes(golden["fixtures"]["t1_henikoff_paper"].encode())  # ragged: Rust panics (lib.rs:180-182)
    r = subprocess.run([str(BIN), "--fasta-input", str(f), "--pair-output", str(tmp_path / "p.tsv")], capture_output=True, text=True)
    assert r.returncode == 101 and "Not all sequences have the same number of symbols" in r.stderr
    r = subprocess.run([str(BIN), "--fasta-input", str(tmp_path / "missing.fa"), "--pair-output", str(tmp_path / "p.tsv")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Error: ")
    r = subprocess.run([str(BIN), "--pair-output", "x"], capture_output=True, text=True)
    assert r.returncode == 1 and "--fasta-input" in r.stderr


def test_cli_synthetic_matches_library(tmp_path, oracle):
    from weightedld_b200.synth import make_alignment
    import weightedld_b200 as wld
    chars = make_alignment(400, 1500, seed=3, block=100, newline_col=True)
    with open(tmp_path / "s.fasta", "wb") as fh:
        for i, row in enumerate(chars):
            fh.write(f">seq{i}\n".encode() + row.tobytes())
    r = subprocess.run([str(BIN), "--fasta-input", str(tmp_path / "s.fasta"), "--pair-output", str(tmp_path / "p.tsv")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    fs = wld.SiteSet.from_multiseq(wld.read_fasta(tmp_path / "s.fasta")).filter_by()
    store = wld.all_weighted_ld_pairs(fs, wld.henikoff_weights(fs), 0.1)
    wld.write_pair_stats(tmp_path / "lib.tsv", store)
    assert (tmp_path / "p.tsv").read_text() == (tmp_path / "lib.tsv").read_text() and len(store) > 100


def test_cli_screen_and_exact_kernel_write_the_same_files(tmp_path):
    """A site count large enough for the automatic screen + refine path: the TSV must be the same,
    byte for byte, as with the exact n-limb kernel (WLD_SCREEN=0), on one GPU and split over two partitions."""
    import os
    from weightedld_b200.synth import make_alignment
    chars = make_alignment(600, 6500, seed=17, block=150, newline_col=True)
    with open(tmp_path / "s.fasta", "wb") as fh:
        for i, row in enumerate(chars):
            fh.write(f">seq{i}\n".encode() + row.tobytes())
    outs = {}
    for name, env in (("auto", {}), ("exact", {"WLD_SCREEN": "0"}), ("always", {"WLD_SCREEN": "2"})):
        r = subprocess.run([str(BIN), "--fasta-input", str(tmp_path / "s.fasta"), "--pair-output", str(tmp_path / f"{name}.tsv"),
                            "--weights-output", str(tmp_path / f"{name}.w.tsv")],
                           capture_output=True, text=True, env=dict(os.environ, WLD_DEBUG="1", RUST_LOG="debug", **env))
        assert r.returncode == 0, r.stderr
        outs[name] = ((tmp_path / f"{name}.tsv").read_bytes(), (tmp_path / f"{name}.w.tsv").read_bytes(), r.stderr)
    assert "one-limb screen + exact refinement" in outs["auto"][2] and "one-limb screen" in outs["always"][2]
    assert "exact kernel" in outs["exact"][2] and "one-limb screen +" not in outs["exact"][2]
    assert outs["auto"][0] == outs["exact"][0] == outs["always"][0] and len(outs["auto"][0]) > 5000
    assert outs["auto"][1] == outs["exact"][1]

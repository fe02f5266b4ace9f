"""CPU tests of the boundary: libwld.so loads, exports every symbol include/wld.h declares, fails
loudly without a GPU (no fallback), and the host-side logic (FASTA reader, output order key,
writers) matches the oracle."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT, fasta_chars


@pytest.fixture(scope="module")
def lib():
    from weightedld_b200 import _lib
    if not _lib.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from weightedld_b200 import _lib
    header = (ROOT / "include" / "wld.h").read_text()
    declared = set(re.findall(r"^WLD_API [\w\s\*]+?\b(wld_\w+)\(", header, flags=re.M))
    assert len(declared) >= 29
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.wld_abi_version() == 3


def test_struct_layout():
    from weightedld_b200 import PAIR_DTYPE
    assert PAIR_DTYPE.itemsize == 20 and [PAIR_DTYPE.fields[n][1] for n in PAIR_DTYPE.names] == [0, 4, 8, 12, 16]


def test_no_gpu_is_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import weightedld_b200 as wld
    with pytest.raises(wld.WldError) as e:
        wld.Context(0)
    assert e.value.status == 3 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    for p in (ROOT / "weightedld_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"}:
            text = p.read_text()
            assert "import oracle" not in text and "from oracle" not in text and "wldo_" not in text, p


def test_read_fasta_matches_oracle(tmp_path, golden, oracle):
    import weightedld_b200 as wld
    for name, text in golden["fixtures"].items():
        f = tmp_path / f"{name}.fasta"
        f.write_bytes(text.encode())
        if name.startswith("t1_"):
            with pytest.raises(ValueError):
                wld.read_fasta(f)  # lib.rs:180-182 panic
            with pytest.raises(ValueError):
                oracle.read_fasta(f)
            continue
        ms = wld.read_fasta(f)
        assert np.array_equal(ms.chars, oracle.read_fasta(f))
        assert np.array_equal(ms.chars, fasta_chars(text))
        assert ms.chars[:, -1].tolist() == [10] * ms.chars.shape[0]  # the newline column, lib.rs:297
        assert ms.names[0] == text.splitlines()[0][1:] + "\n"       # names keep their newline, lib.rs:295
    crlf = tmp_path / "crlf.fasta"
    crlf.write_bytes(b">a\r\nACGT\r\n>b\r\nAC-T\r\n")
    assert wld.read_fasta(crlf).chars.shape == (2, 6)


def test_pair_order_key_is_reference_tile_order(lib, oracle):
    import weightedld_b200 as wld
    for n_kept in (5, 256, 257, 700, 1500):
        n = (n_kept + 255) // 256
        order = {oracle.triu_index(n, i): i for i in range(n * (n + 1) // 2)}
        keys = {}
        for tr in range(n):
            for tc in range(tr, n):
                a, b = tr * 256, min(tc * 256 + 1, n_kept - 1)
                k = lib.wld_pair_order_key(n_kept, a, b)
                assert k == int(wld.pair_order_key(n_kept, a, b))
                keys[(tr, tc)] = k
        assert sorted(keys, key=keys.get) == sorted(order, key=order.get)


def test_writers_match_reference_format(tmp_path, oracle):
    import weightedld_b200 as wld
    arr = np.array([(3, 9, 0.0625, np.nan, 0.5), (1, 2, -0.25, np.inf, 1.0), (0, 7, -0.0, -np.inf, 0.1005)],
                   dtype=wld.PAIR_DTYPE)
    wld.write_pair_stats(tmp_path / "p.tsv", wld.PairStore(arr, 3))
    oracle.write_pairs(tmp_path / "o.tsv", arr.astype(oracle.PAIR_DTYPE))
    got = (tmp_path / "p.tsv").read_text()
    assert got == (tmp_path / "o.tsv").read_text()
    assert got.splitlines()[0] == "site_a\tsite_b\td\td'\tr2" and got.splitlines()[1] == "3\t9\t0.062\tNaN\t0.500"
    w = np.array([1.0, 0.30000001, 0.0005], np.float32)
    wld.write_henikoff_weights(tmp_path / "w.tsv", w)
    oracle.write_weights(tmp_path / "wo.tsv", w)
    assert (tmp_path / "w.tsv").read_text() == (tmp_path / "wo.tsv").read_text() == \
        "Sequence_index\thk_weight\n0\t1.000\n1\t0.300\n2\t0.001\n"

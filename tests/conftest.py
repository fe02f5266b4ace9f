import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """Built artefacts are not in the history: build them (nvcc cross-compiles without a GPU) when a fresh
    checkout runs the tests before __graft_entry__.build()."""
    needed = [ROOT / "weightedld_b200" / "libwld.so", ROOT / "weightedld_b200" / "weighted_ld",
              ROOT / "oracle" / "_build" / "liboracle.so"]
    if not all(p.exists() for p in needed):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    return {
        "fixtures": json.loads((GOLDEN / "fixtures.json").read_text()),
        "python_ref": json.loads((GOLDEN / "python_ref.json").read_text()),
        "rust_kat": json.loads((GOLDEN / "rust_kat.json").read_text()),
        "t7": np.load(GOLDEN / "t7_haplotypes.npz"),
    }


def fasta_chars(text: str) -> np.ndarray:
    """The reference's read_fasta (lib.rs:277-307) applied to fixture text: every non-'>' line
    is a sequence INCLUDING its newline."""
    rows = [ln for ln in text.splitlines(keepends=True) if not ln.startswith(">")]
    if len({len(r) for r in rows}) != 1:
        raise ValueError("Not all sequences have the same number of symbols")
    return np.frombuffer("".join(rows).encode(), np.uint8).reshape(len(rows), -1).copy()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O

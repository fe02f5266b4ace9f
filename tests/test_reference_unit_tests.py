"""The reference's own Rust unit tests (rust/weighted_ld/src/lib.rs:686-802), one to one and under their own
names, run against the CUDA path through the C ABI (mirrored API of weightedld_b200/api.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SYMS = "ACGT-?"  # codes 0..5; '?' is Unknown like any other byte (lib.rs:61)


@pytest.fixture(scope="module")
def wld():
    import weightedld_b200 as w
    return w


def column_with_counts(counts):
    return "".join(SYMS[k] * int(n) for k, n in enumerate(counts))


def test_histogram_from_slice(wld, golden):          # lib.rs:692-703
    k = golden["rust_kat"]["histogram"]
    ss = wld.SiteSet.from_strs(list(k["symbols"]))   # one column, 13 sequences
    assert ss.site_histogram(0).tolist() == k["hist"] == [3, 2, 1, 4, 2, 1]


def test_hist_major_minor(wld, golden):              # lib.rs:705-728
    for case in golden["rust_kat"]["major_minor"]["cases"]:
        ss = wld.SiteSet.from_strs(list(column_with_counts(case["hist"])))
        assert ss.site_histogram(0).tolist() == case["hist"]
        maj, mnr = ss.context.major_minor()
        assert (int(maj[0]), int(mnr[0])) == (case["major"], case["minor"]), case


@pytest.mark.parametrize("idx,name", [(0, "test_henikoff_weights_1"), (1, "test_henikoff_weights_2"), (2, "test_henikoff_weights_3")])
def test_henikoff_weights(wld, golden, idx, name):   # lib.rs:731-750
    case = golden["rust_kat"]["henikoff"][idx]
    w = wld.henikoff_weights(wld.SiteSet.from_strs(case["rows"]))
    tol = 1e-6 if case["tol"] == "ulps" else case["tol"]
    assert np.allclose(w, case["weights"], atol=tol, rtol=0), name


@pytest.mark.parametrize("idx,name", [(0, "test_ld_pair_unweighted_ld0"), (1, "test_ld_pair_unweighted_ld1"),
                                      (2, "test_single_weighted_ld_pair")])
def test_ld_pair(wld, oracle, golden, idx, name):     # lib.rs:753-801
    case = golden["rust_kat"]["pair"][idx]
    a = oracle.encode(np.frombuffer(case["a"].encode(), np.uint8))
    b = oracle.encode(np.frombuffer(case["b"].encode(), np.uint8))
    r2, d, dp = wld.single_weighted_ld_pair(a, b, np.array(case["w"], np.float32))
    assert abs(d - case["d"]) <= case["tol"] and abs(dp - case["d_prime"]) <= case["tol"] and abs(r2 - case["r2"]) <= case["tol"], name

"""The reference's Python test-suite (/root/reference/test.py, TestStuff, nine collected cases plus the
orphaned test_vcf), case for case and under the same names, with `wld` bound to the B200 mirror of
WeightedLD.py (weightedld_b200.pycompat) instead of the reference module.  Same fixtures (tests/golden/
fixtures.json holds their text), same thresholds, same assertions (test.py:13-118, :152-159)."""
import gzip
import io

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

MIN_ACGT, MIN_VARIABILITY = 0.8, 0.02          # test.py:10-11


@pytest.fixture(scope="module")
def wld():
    from weightedld_b200 import pycompat
    return pycompat


@pytest.fixture()
def fixture_file(golden, tmp_path):
    def write(name):
        f = tmp_path / f"{name}.fasta"
        f.write_text(golden["fixtures"][name])
        return f
    return write


def captured_ld(wld, alignment, weights, site_map) -> str:
    out = io.StringIO()                         # the reference redirects sys.stdout; ld() here takes the stream
    wld.ld(alignment, weights, site_map, file=out)
    return out.getvalue()


def test_read_fasta(wld, fixture_file):                                   # test.py:13-17
    assert wld.read_fasta(fixture_file("t1_henikoff_paper")).sum() == 65


def test_var_sitesHK(wld, fixture_file):                                  # test.py:19-26
    alignment = wld.read_fasta(fixture_file("t1_henikoff_paper"))
    var_sites_HK, _ = wld.compute_variable_sites(alignment, MIN_ACGT, MIN_VARIABILITY)
    assert var_sites_HK.tolist() == [False, False, True, True, True, True, True]


def test_var_sitesLD(wld, fixture_file):                                  # test.py:28-35
    alignment = wld.read_fasta(fixture_file("t6_varsites_hk_ld"))
    var_sites_HK, var_sites_LD = wld.compute_variable_sites(alignment, MIN_ACGT, 0.2)
    assert var_sites_HK[1] != var_sites_LD[1]


def test_hkw_simple(wld, fixture_file):                                   # test.py:37-47
    alignment = wld.read_fasta(fixture_file("t1_henikoff_paper"))
    var_sites_HK, _ = wld.compute_variable_sites(alignment, MIN_ACGT, MIN_VARIABILITY)
    weightsHK = wld.henikoff_weighting(alignment[:, var_sites_HK])
    assert np.allclose(weightsHK, np.array([0.5, 0.5, 0.5, 0.5, 1.0]), rtol=1e-02, atol=1e-02)


def test_hkw_complex(wld, fixture_file):                                  # test.py:49-57
    alignment = wld.read_fasta(fixture_file("t2_henikoff_complex1"))
    var_sites_HK, _ = wld.compute_variable_sites(alignment, MIN_ACGT, MIN_VARIABILITY)
    assert wld.henikoff_weighting(alignment[:, var_sites_HK])[0] == 1.0


def test_hkw_complex_indel(wld, fixture_file):                            # test.py:59-67
    alignment = wld.read_fasta(fixture_file("t3_henikoff_complex2"))
    var_sites_HK, _ = wld.compute_variable_sites(alignment, MIN_ACGT, MIN_VARIABILITY)
    assert wld.henikoff_weighting(alignment[:, var_sites_HK])[7] == 1.0


def test_0ld_flatw(wld, fixture_file):                                    # test.py:69-84
    alignment = wld.read_fasta(fixture_file("t4_weights1_ld0"))
    var_sites_HK, var_sites_LD = wld.compute_variable_sites(alignment, 0.99, MIN_VARIABILITY)
    weightsHK = wld.henikoff_weighting(alignment[:, var_sites_HK])
    out = captured_ld(wld, alignment[:, var_sites_LD], weightsHK, np.where(var_sites_LD)[0])
    assert out[22:25] == "0.0"


def test_wld_flatw(wld, fixture_file):                                    # test.py:86-101
    alignment = wld.read_fasta(fixture_file("t4_weights1_ld0"))
    var_sites_HK, var_sites_LD = wld.compute_variable_sites(alignment, 0.1, 0.2)
    weightsHK = wld.henikoff_weighting(alignment[:, var_sites_HK])
    out = captured_ld(wld, alignment[:, var_sites_LD], weightsHK, np.where(var_sites_LD)[0])
    assert out[22:25] != "0.0"


def test_ld_flatw(wld, fixture_file):                                     # test.py:103-118
    alignment = wld.read_fasta(fixture_file("t5_weights1_ld0.25"))
    var_sites_HK, var_sites_LD = wld.compute_variable_sites(alignment, MIN_ACGT, MIN_VARIABILITY)
    weightsHK = wld.henikoff_weighting(alignment[:, var_sites_HK])
    out = captured_ld(wld, alignment[:, var_sites_LD], weightsHK, np.where(var_sites_LD)[0])
    assert out[22:27] == "-0.25" and out[32:33] == "1"


def test_vcf(wld, tmp_path):                                              # test.py:152-159 (never collected there)
    f = tmp_path / "t7_1000genome.vcf"
    f.write_bytes(gzip.decompress((GOLDEN / "t7_1000genome.vcf.gz").read_bytes()))
    alignment, site_map = wld.handle_vcf(f)
    weights = wld.henikoff_weighting(alignment)
    captured_ld(wld, alignment, weights, site_map)
    assert round(weights.mean(), 3) == 0.002

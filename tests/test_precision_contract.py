"""The precision contract of the pair stage as a function of the weight range.

The reference sums f32 weights (lib.rs:469-479): 24 RELATIVE bits per weight.  The B200 pair stage sums
integers q = m * 2^(G-e) (block-exponent fixed point, include/wld.h wld_set_gain_bits) whose per-weight
relative error eps is REPORTED by the library (wld_pair_info.weight_rel_err) and bounded a priori by
    eps <= 2^-B                      when the smallest nonzero weight is >= 2^-G of the largest (x <= G)
    eps <= 2^(x - G - B)             otherwise            (B = mantissa bits, x = weight_span_log2)
Every weighted sum of lib.rs:469-479 has non-negative terms, hence relative error <= eps, and
    |delta r2| <= (8 / sqrt(pa * pb) + 4) * eps,   |delta D| <= 8 * eps
(pa, pb: weighted minor-allele frequencies of the two sites) against the f64 evaluation of lib.rs:455-521 on
the UNQUANTISED f32 weights, plus one f32 rounding of the record.  The tests below check exactly that, with
clonal clusters of identical small weights and one weight-1 outlier (the shape Henikoff weights take on an
outbreak alignment) spanning 1e-3 ... 1e-7, on the CPU (oracle restatement of the quantiser) and on the GPU
(through the C ABI), and with Henikoff weights of a SARS-CoV-2-like alignment of 100 000 sequences.
"""
import numpy as np
import pytest

F32_ULP = 1.2e-7  # one rounding of an r2 / D value <= 1 to the f32 record


def clonal_case(n_seqs=4000, n_cols=400, span=1e-5, seed=11, founders=32):
    """Clonal alignment + cluster-wise identical weights in [span, 4 span] + one outlier of weight 1."""
    from weightedld_b200.synth import make_alignment
    chars = make_alignment(n_seqs, n_cols, seed=seed, block=40, clonal=True, founders=founders)
    p = 1.0 / np.arange(1, founders + 1) ** 1.6
    cluster = np.random.Generator(np.random.PCG64(seed)).choice(founders, size=n_seqs, p=p / p.sum())
    cw = span * np.random.default_rng(3).uniform(1, 4, size=founders)
    w = cw[cluster].astype(np.float32)
    w[0] = 1.0
    return chars, w


def apriori_eps(bits, gain_bits, span_log2):
    return 2.0 ** -bits if span_log2 <= gain_bits else 2.0 ** (span_log2 - gain_bits - bits)


def realised_eps(w32, wq, bits, gain_bits):
    u = w32.astype(np.float64) / np.float64(w32.max())
    nz = u > 0
    return float(np.max(np.abs(wq[nz] / (2.0 ** gain_bits * (2.0 ** bits - 1)) - u[nz]) / u[nz]))


def weighted_minor_freq_min(fs, w32):
    """Smallest weighted minor-allele frequency over the kept sites (for the r2 bound)."""
    maj, mnr = fs.major_minor()
    w = w32.astype(np.float64)
    lo = 1.0
    for k in range(fs.n_sites):
        c = fs.codes[k]
        a, b = w[c == maj[k]].sum(), w[c == mnr[k]].sum()
        if a + b > 0 and b > 0:
            lo = min(lo, b / (a + b), a / (a + b))
    return lo


def check_against_unquantised(oracle, fs, w32, got, eps, thr):
    """got: records (a, b, d, d_prime, r2) computed from the quantised weights with r2 > thr; compared with
    the f64 oracle on the unquantised f32 weights."""
    ref, _ = oracle.all_weighted_ld_pairs(fs, w32.astype(np.float64), -1.0, oracle.F64)
    pmin = weighted_minor_freq_min(fs, w32)
    b_r2 = (8.0 / pmin + 4.0) * eps + F32_ULP
    b_d = 8.0 * eps + F32_ULP
    key = lambda a, b: a.astype(np.int64) << 32 | b.astype(np.int64)
    rk = key(ref["a"], ref["b"])
    order = np.argsort(rk)
    rk, ref = rk[order], ref[order]
    gk = key(got["a"], got["b"])
    pos = np.searchsorted(rk, gk)
    assert np.all(rk[pos] == gk)
    m = ref[pos]
    assert np.max(np.abs(m["r2"].astype(np.float64) - got["r2"])) <= b_r2
    assert np.max(np.abs(m["d"].astype(np.float64) - got["d"])) <= b_d
    # pair set: identical outside the band of width b_r2 around the threshold
    want = set(rk[ref["r2"] > np.float32(thr)].tolist())
    band = set(rk[np.abs(ref["r2"].astype(np.float64) - thr) <= b_r2].tolist())
    assert (set(gk.tolist()) ^ want) <= band
    return b_r2, float(np.max(np.abs(m["r2"].astype(np.float64) - got["r2"])))


@pytest.mark.parametrize("span", [1e-3, 1e-4, 1e-5, 1e-7])
def test_block_exponent_quantiser_bound_cpu(oracle, span):
    chars, w = clonal_case(span=span)
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    bits, gain = oracle.auto_quant_params(w)
    x = max(0, -int(np.frexp(np.float64(w[w > 0].min()) / np.float64(w.max()))[1]))
    assert (bits, gain) == (32, 7) and x >= 7
    wq = oracle.quantize_weights(w, bits, gain)
    assert np.all(wq == np.rint(wq)) and wq.max() == (2.0 ** bits - 1) * 2.0 ** gain and wq.sum() < 2.0 ** 53
    eps = realised_eps(w, wq, bits, gain)
    assert eps <= apriori_eps(bits, gain, x) * (1 + 1e-6)
    if x <= 15:
        assert eps <= 2.0 ** -24 * (1 + 1e-6)  # what the reference's f32 weights carry (weights down to 2^-16 of the largest)
    if span >= 1e-5:
        assert eps <= 2.0 ** -23
    got, _ = oracle.all_weighted_ld_pairs(fs, wq, 0.1, oracle.F64)
    check_against_unquantised(oracle, fs, w, got, eps, 0.1)
    # the round-1 scheme (24 bits relative to the LARGEST weight) is what the contract replaces
    eps_old = realised_eps(w, oracle.quantize_weights(w, 24, 0), 24, 0)
    assert eps_old > 100 * eps


def test_mild_weights_keep_three_limbs_cpu(oracle):
    from weightedld_b200.synth import make_alignment, make_weights
    chars = make_alignment(500, 120, seed=11, block=30, clonal=True)
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    for w in (oracle.henikoff_weights(fs), make_weights(500)):
        bits, gain = oracle.auto_quant_params(w)
        x = max(0, -int(np.frexp(np.float64(w[w > 0].min()) / np.float64(w.max()))[1]))
        if x <= 7:
            assert bits == 24 and gain == x
        wq = oracle.quantize_weights(w, bits, gain)
        eps = realised_eps(w, wq, bits, gain)
        assert eps <= apriori_eps(bits, gain, x) * (1 + 1e-6) and apriori_eps(bits, gain, x) <= 2.0 ** -24
        got, _ = oracle.all_weighted_ld_pairs(fs, wq, 0.1, oracle.F64)
        check_against_unquantised(oracle, fs, w, got, eps, 0.1)


# ------------------------------------------------------------------------------------------------ GPU
def gpu_run(chars, w, kernel="i8", thr=0.1, limbs=0, gain=-1, limb_bits=0, henikoff=False):
    import weightedld_b200 as wld
    with wld.Context(0) as ctx:
        ctx.set_pair_kernel(kernel)
        ctx.set_limbs(limbs)
        ctx.set_gain_bits(gain)
        ctx.set_limb_bits(limb_bits)
        ctx.load_alignment(chars)
        ctx.filter_sites()
        if henikoff:
            ctx.henikoff()
        else:
            ctx.set_weights(w)
        n, done = ctx.ld_pairs(thr)
        return ctx.fetch_pairs(n), ctx.pair_info(), ctx.pair_weights(), ctx.weights(), done


def as_oracle_records(oracle, gpu):
    out = np.empty(len(gpu), oracle.PAIR_DTYPE)
    for src, dst in (("site_a", "a"), ("site_b", "b"), ("d", "d"), ("d_prime", "d_prime"), ("r2", "r2")):
        out[dst] = gpu[src]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["i8", "bf16"])
@pytest.mark.parametrize("span", [1e-3, 1e-4, 1e-5, 1e-7])
def test_wide_span_weights_gpu(oracle, span, kernel):
    """Clonal identical small weights + one outlier through the C ABI: the library picks 4 limbs, reports its
    realised per-weight error, stays bit-exact against the f64 oracle on ITS integers, and within the stated
    bound of the f64 oracle on the UNQUANTISED f32 weights."""
    chars, w = clonal_case(span=span)
    gpu, info, wq, w32, done = gpu_run(chars, w, kernel)
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    assert done == fs.n_sites * (fs.n_sites - 1) // 2
    assert info.n_limbs == 4 and info.limb_bits == 8 and info.weight_span_log2 >= 7
    if kernel == "i8":
        assert info.gain_bits == 7
    assert np.array_equal(wq, oracle.quantize_weights(w, info.weight_bits, info.gain_bits))
    eps = realised_eps(w, wq, info.weight_bits, info.gain_bits)
    assert info.weight_rel_err == pytest.approx(eps, rel=1e-6)
    assert eps <= apriori_eps(info.weight_bits, info.gain_bits, info.weight_span_log2) * (1 + 1e-6)
    if kernel == "i8":
        assert eps <= (2.0 ** -24 if info.weight_span_log2 <= 15 else 2.0 ** (info.weight_span_log2 - 39)) * (1 + 1e-6)
    # bit-exact on the same integers
    ref, _ = oracle.all_weighted_ld_pairs(fs, wq, 0.1, oracle.F64)
    assert len(ref) == len(gpu) and np.array_equal(gpu["site_a"], ref["a"]) and np.array_equal(gpu["site_b"], ref["b"])
    for f in ("d", "d_prime", "r2"):
        assert np.array_equal(gpu[f].view(np.uint32), ref[f].view(np.uint32))
    # stated bound against the unquantised weights
    check_against_unquantised(oracle, fs, w, as_oracle_records(oracle, gpu), eps, 0.1)


@pytest.mark.gpu
def test_three_limbs_forced_reports_its_larger_error(oracle):
    """A caller may force 3 limbs on wide-span weights (wld_set_limbs): the library then REPORTS the larger
    error, and the bound computed from the report still holds."""
    chars, w = clonal_case(span=1e-5)
    gpu, info, wq, _, _ = gpu_run(chars, w, "i8", limbs=3)
    assert info.n_limbs == 3 and info.gain_bits == 7
    eps = realised_eps(w, wq, 24, 7)
    assert info.weight_rel_err == pytest.approx(eps, rel=1e-6) and 2.0 ** -24 < eps <= apriori_eps(24, 7, info.weight_span_log2)
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    check_against_unquantised(oracle, fs, w, as_oracle_records(oracle, gpu), eps, 0.1)


@pytest.mark.gpu
def test_sarscov2_like_henikoff_weights_100k(oracle):
    """Henikoff weights of a clonal 100 000-sequence alignment (config 4's N and weight shape, fewer columns so
    that the oracle finishes): per-weight error <= 2^-24, D / r2 within the bound of the f64 oracle on the
    unquantised weights, bit-exact on the library's integers."""
    from weightedld_b200.synth import make_sarscov2_like
    chars = make_sarscov2_like(100_000, 420, seed=0xC0FFEE + 3)
    gpu, info, wq, w32, done = gpu_run(chars, None, "i8", henikoff=True)
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    assert fs.n_sites >= 60 and done == fs.n_sites * (fs.n_sites - 1) // 2
    assert w32.max() / w32[w32 > 0].min() > 30
    assert np.array_equal(wq, oracle.quantize_weights(w32, info.weight_bits, info.gain_bits))
    eps = realised_eps(w32, wq, info.weight_bits, info.gain_bits)
    assert info.weight_rel_err == pytest.approx(eps, rel=1e-6) and eps <= 2.0 ** -24 * (1 + 1e-6)
    ref, _ = oracle.all_weighted_ld_pairs(fs, wq, 0.1, oracle.F64)
    assert len(ref) == len(gpu) > 0 and np.array_equal(gpu["site_a"], ref["a"]) and np.array_equal(gpu["site_b"], ref["b"])
    for f in ("d", "d_prime", "r2"):
        assert np.array_equal(gpu[f].view(np.uint32), ref[f].view(np.uint32))
    check_against_unquantised(oracle, fs, w32, as_oracle_records(oracle, gpu), eps, 0.1)


@pytest.mark.gpu
def test_narrow_limbs_at_100k_sequences(oracle):
    """n_seqs = 100 000 > 65 793 with near-uniform weights: a limb column sum of 8-bit limbs exceeds 2^24, so
    the bf16 kernel (fp32 accumulator) must narrow its limbs; the i8 kernel forced to the same limb width and
    gain must agree byte for byte, and both with the f64 oracle on those integers."""
    from weightedld_b200.synth import make_alignment
    n = 100_000
    chars = make_alignment(n, 260, seed=77, block=40)
    w = np.random.default_rng(5).uniform(0.7, 1.0, size=n).astype(np.float32)
    b_pairs, b_info, b_wq, _, _ = gpu_run(chars, w, "bf16", thr=0.05)
    assert b_info.limb_bits < 8 and b_info.n_limbs == 3 and b_info.weight_bits == 3 * b_info.limb_bits
    i_pairs, i_info, i_wq, _, _ = gpu_run(chars, w, "i8", thr=0.05, gain=b_info.gain_bits, limb_bits=b_info.limb_bits)
    assert i_info.limb_bits == b_info.limb_bits and np.array_equal(i_wq, b_wq)
    assert len(b_pairs) > 0 and b_pairs.tobytes() == i_pairs.tobytes()
    assert np.array_equal(b_wq, oracle.quantize_weights(w, b_info.weight_bits, b_info.gain_bits))
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    ref, _ = oracle.all_weighted_ld_pairs(fs, b_wq, 0.05, oracle.F64)
    assert len(ref) == len(b_pairs)
    for f in ("d", "d_prime", "r2"):
        assert np.array_equal(b_pairs[f].view(np.uint32), ref[f].view(np.uint32))
    # the default i8 path keeps 8-bit limbs at this size (s32 accumulator)
    assert gpu_run(chars, w, "i8", thr=0.05)[1].limb_bits == 8

"""Parity at the FULL sizes of BASELINE.json configs[3] (100 000 x 30 000, ~9 300 kept sites) and configs[4]
(10 000 x 50 000, 48 601 kept sites) — run with -m gpu on a B200.

The oracle cannot enumerate 4.3e7 / 1.2e9 pairs of 1e5 / 1e4 sequences in test time, so the CUDA path is
checked through what holds for ANY input because the weighted sums are exact integers:
  * every pair is computed exactly once (pair count = L(L-1)/2), stage 1 agrees with the oracle bit for bit;
  * the integer weights are the ones the oracle's quantiser gives for the reported (limbs, limb bits, gain bits);
  * permuting the sequences (with their weights) leaves every output record BIT-identical;
  * tensor-core variants agree byte for byte: i8 2-CTA / 1-CTA, and the bf16 kernel on the SAME integer weights
    (config 4 is where a limb column sum may exceed 2^24, i.e. where `limb_bits` < 8 can run on the bf16 path);
  * the union of tile partitions equals the single-GPU result, record for record;
  * 300 randomly drawn survivors and 300 randomly drawn site pairs are recomputed by the f64 oracle from the raw
    alignment columns (lib.rs:390-521) and must match to the last bit / be absent exactly when the oracle
    rejects them.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
THR = 0.1


def run(wld, chars, weights=None, kernel="i8", ctas=2, partition=None, gain_bits=-1, limb_bits=0, limbs=0,
        want=("pairs",)):
    with wld.Context(0) as ctx:
        ctx.set_pair_kernel(kernel)
        ctx.set_cta_group(ctas)
        ctx.set_gain_bits(gain_bits)
        ctx.set_limb_bits(limb_bits)
        ctx.set_limbs(limbs)
        if partition:
            ctx.set_partition(*partition)
        ctx.load_alignment(chars)
        n_kept = ctx.filter_sites()
        if weights is None:
            ctx.henikoff()
        else:
            ctx.set_weights(weights)
        n, done = ctx.ld_pairs(THR)
        out = {"done": done, "n_kept": n_kept, "info": ctx.pair_info(), "n": n}
        if "pairs" in want:
            out["pairs"] = ctx.fetch_pairs(n)
        if "kept" in want:
            out["kept"] = ctx.fetch_pairs(n, wld.FETCH_KEPT_INDEX | wld.FETCH_UNORDERED)
        if "meta" in want:
            out.update(w=ctx.weights(), wq=ctx.pair_weights(), site_map=ctx.site_map(), hist=ctx.histograms(),
                       major_minor=ctx.major_minor())
        return out


def spot_check(oracle, chars, base, rng, n_each=300):
    r = base["pairs"]
    wq, smap = base["wq"], base["site_map"]
    cols = {}

    def col(c):
        if c not in cols:
            cols[c] = oracle.encode(np.ascontiguousarray(chars[:, c]))
        return cols[c]

    for x in r[rng.choice(len(r), min(n_each, len(r)), replace=False)]:
        st = oracle.single_weighted_ld_pair(col(int(x["site_a"])), col(int(x["site_b"])), wq, oracle.F64)
        assert st is not None
        got = np.array([x["r2"], x["d"], x["d_prime"]], np.float32)
        assert np.array_equal(got.view(np.uint32), np.array(st, np.float32).view(np.uint32))
    have = set((r["site_a"].astype(np.int64) << 32 | r["site_b"].astype(np.int64)).tolist())
    L = base["n_kept"]
    for _ in range(n_each):
        i, j = sorted(rng.choice(L, 2, replace=False))
        a, b = int(smap[i]), int(smap[j])
        st = oracle.single_weighted_ld_pair(col(a), col(b), wq, oracle.F64)
        passes = st is not None and np.float32(st[0]) > np.float32(THR)
        assert passes == ((a << 32 | b) in have)


def check_stage1(oracle, chars, base):
    """Histograms of 200 random columns + the whole site mask against the oracle (the oracle's site filter is cheap)."""
    ss = oracle.siteset_from_chars(chars)
    assert np.array_equal(base["hist"].astype(np.uint64), ss.hists)
    fs = oracle.filter_sites(ss)
    assert fs.n_sites == base["n_kept"] and np.array_equal(base["site_map"], fs.site_map)
    maj, mnr = fs.major_minor()
    assert np.array_equal(base["major_minor"][0], maj) and np.array_equal(base["major_minor"][1], mnr)
    return fs


@pytest.fixture(scope="module")
def c5():
    import bench
    return bench.make_input("c5")


@pytest.fixture(scope="module")
def c4():
    import bench
    return bench.make_input("c4")


def test_config5_full_size(c5, oracle):
    import weightedld_b200 as wld
    base = run(wld, c5, want=("pairs", "meta"))
    L = base["n_kept"]
    assert L == 48601 and base["done"] == L * (L - 1) // 2 == 1_181_004_300 and len(base["pairs"]) > 10_000
    info = base["info"]
    assert info.n_limbs == 3 and info.limb_bits == 8 and info.weight_rel_err <= 2.0 ** -24 * (1 + 1e-6)
    fs = check_stage1(oracle, c5, base)
    w = base["w"]
    assert np.allclose(w, oracle.henikoff_weights(fs, f64=True), rtol=1e-6)
    assert np.array_equal(base["wq"], oracle.quantize_weights(w, info.weight_bits, info.gain_bits))
    ref = base["pairs"].tobytes()
    rng = np.random.default_rng(50)

    perm = rng.permutation(c5.shape[0])
    assert run(wld, np.ascontiguousarray(c5[perm]), weights=w[perm])["pairs"].tobytes() == ref
    assert run(wld, c5, weights=w, ctas=1)["pairs"].tobytes() == ref
    vb = run(wld, c5, weights=w, kernel="bf16", want=("pairs", "meta"))
    vi = run(wld, c5, weights=w, gain_bits=vb["info"].gain_bits, limb_bits=vb["info"].limb_bits, want=("pairs", "meta"))
    assert np.array_equal(vb["wq"], vi["wq"]) and vb["pairs"].tobytes() == vi["pairs"].tobytes()

    parts = [run(wld, c5, weights=w, partition=(g, 4), want=("kept",)) for g in range(4)]
    assert sum(p["done"] for p in parts) == base["done"]
    assert wld.merge_shards(L, [p["kept"] for p in parts], base["site_map"]).tobytes() == ref

    spot_check(oracle, c5, base, rng)


def test_config4_full_size(c4, oracle):
    import weightedld_b200 as wld
    base = run(wld, c4, want=("pairs", "meta"))
    L = base["n_kept"]
    assert 8000 < L < 11000 and base["done"] == L * (L - 1) // 2 and len(base["pairs"]) > 1_000_000
    info = base["info"]
    w = base["w"]
    assert w.max() / w[w > 0].min() > 100                      # "weight-heavy": Henikoff weights span decades
    # the contract: 24 relative bits per weight (what f32 carries) — 4 limbs when the span needs them
    assert info.weight_rel_err <= 2.0 ** -24 * (1 + 1e-6)
    assert info.n_limbs == (4 if info.weight_span_log2 > 7 else 3)
    fs = check_stage1(oracle, c4, base)
    assert np.allclose(w, oracle.henikoff_weights(fs, f64=True), rtol=1e-6)
    assert np.array_equal(base["wq"], oracle.quantize_weights(w, info.weight_bits, info.gain_bits))
    ref = base["pairs"].tobytes()
    rng = np.random.default_rng(40)

    perm = rng.permutation(c4.shape[0])
    assert run(wld, np.ascontiguousarray(c4[perm]), weights=w[perm])["pairs"].tobytes() == ref
    assert run(wld, c4, weights=w, ctas=1)["pairs"].tobytes() == ref

    # bf16 limbs, fp32 accumulation at n_seqs = 100 000 > 65 793: exactness is decided from the limb column sums;
    # the i8 kernel on the SAME integers (same limbs, limb width and gain) must agree byte for byte
    vb = run(wld, c4, weights=w, kernel="bf16", want=("pairs", "meta"))
    vi = run(wld, c4, weights=w, limbs=vb["info"].n_limbs, gain_bits=vb["info"].gain_bits, limb_bits=vb["info"].limb_bits,
             want=("pairs", "meta"))
    assert np.array_equal(vb["wq"], vi["wq"]) and len(vb["pairs"]) > 1_000_000 and vb["pairs"].tobytes() == vi["pairs"].tobytes()
    # and with near-uniform weights a limb column sum of 8-bit limbs passes 2^24: the bf16 path must narrow its limbs
    wu = rng.uniform(0.7, 1.0, size=c4.shape[0]).astype(np.float32)
    nb = run(wld, c4, weights=wu, kernel="bf16", want=("pairs", "meta"))
    assert nb["info"].limb_bits < 8
    ni = run(wld, c4, weights=wu, limbs=nb["info"].n_limbs, gain_bits=nb["info"].gain_bits, limb_bits=nb["info"].limb_bits,
             want=("pairs", "meta"))
    assert np.array_equal(nb["wq"], ni["wq"]) and nb["pairs"].tobytes() == ni["pairs"].tobytes()
    spot_check(oracle, c4, nb, rng, n_each=100)

    parts = [run(wld, c4, weights=w, partition=(g, 3), want=("kept",)) for g in range(3)]
    assert sum(p["done"] for p in parts) == base["done"]
    assert wld.merge_shards(L, [p["kept"] for p in parts], base["site_map"]).tobytes() == ref

    spot_check(oracle, c4, base, rng)

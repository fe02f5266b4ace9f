"""Screen + refine (wld_set_screen): the one-limb Gram with a rigorous upper bound on r2, followed by
the exact recomputation of the pairs it cannot rule out, must return the SAME records, bit for bit and in the same
order, as the exact n-limb kernel (and therefore as the f64 oracle on the same fixed-point weights).

CPU part (no GPU): the bound itself (pair_epilogue.cuh, ld_screen_f32) restated in numpy float32 and fuzzed against
the exact r2 of tables whose true sums lie anywhere in the interval the screen allows.
GPU part (-m gpu): byte identity of screen-always / automatic / never over shapes, thresholds, weights, CTA groups and
partitions; the automatic choice on low-LD and high-LD inputs; candidate-buffer overflow; wide-span weights."""
import numpy as np
import pytest

THR = 0.1


# ------------------------------------------------------------------------------------------ the bound (CPU)
def screen_f32(x, kappa, thr_lo_f):
    """ld_screen_f32, operation for operation in float32.  x: (n, 4) integer sums AB, Ab, aB, ab."""
    f = x.astype(np.float32)
    AB, Ab, aB, ab = f[:, 0], f[:, 1], f[:, 2], f[:, 3]
    P, Q = AB * ab, Ab * aB
    n = np.abs(P - Q) + np.float32(kappa) * np.maximum(P, Q)
    den1, den2 = (AB + Ab) * (aB + ab), (AB + aB) * (Ab + ab)
    return (den1 > 0) & (den2 > 0) & (n * n >= (np.float32(thr_lo_f) * den1) * den2)


def kappa_of(top_min):
    eta = 1.0 / top_min
    k = np.float32((1 + eta) ** 2 - 1 + 1e-5)
    return np.nextafter(k, np.float32(np.inf)) if float(k) < (1 + eta) ** 2 - 1 + 1e-5 else k


def thr_lo_f_of(thr):
    lo = float(thr) - abs(float(thr)) * 1e-6 - 1e-24
    f = np.float32(lo * (1 - 1e-5))
    if float(f) > lo * (1 - 1e-5):
        f = np.nextafter(f, np.float32(-np.inf))
    return max(f, np.float32(0))


def exact_r2(y):
    y = y.astype(np.float64)
    AB, Ab, aB, ab = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
    num = (AB * ab - Ab * aB) ** 2
    den = (AB + Ab) * (aB + ab) * (AB + aB) * (Ab + ab)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(den > 0, num / den, np.nan)


@pytest.mark.parametrize("top_min", [32, 64, 128, 200, 255])
@pytest.mark.parametrize("thr", [0.001, 0.05, 0.1, 0.5, 0.9])
def test_screen_never_rejects_a_pair_the_exact_path_keeps(top_min, thr):
    rng = np.random.default_rng(top_min * 1000 + int(thr * 1000))
    n = 400_000
    # tables of every scale, including tiny counts, empty cells and near-threshold LD
    scale = 10 ** rng.uniform(0, 9.3, size=(n, 1))
    p = rng.dirichlet([0.7, 0.7, 0.7, 0.7], size=n)
    x = np.floor(p * scale).astype(np.int64)
    x[rng.random(n) < 0.05, rng.integers(0, 4)] = 0
    x = np.minimum(x, (2 ** 31 - 1) // 4)
    # tables tuned to land near the threshold: AB ab / (Ab aB) chosen so that r2 ~ thr
    m = n // 2
    t = rng.uniform(0.2, 0.8, size=m)
    u = rng.uniform(0.2, 0.8, size=m)
    d = np.sqrt(thr * t * (1 - t) * u * (1 - u)) * rng.uniform(0.97, 1.03, size=m) * rng.choice([-1, 1], size=m)
    tab = np.stack([t * u + d, t * (1 - u) - d, (1 - t) * u - d, (1 - t) * (1 - u) + d], 1)
    ok = (tab > 0).all(1)
    x[:m][ok] = np.floor(tab[ok] * scale[:m][ok]).astype(np.int64)
    kappa, thr_lo_f = kappa_of(top_min), thr_lo_f_of(thr)
    passes = screen_f32(x, kappa, thr_lo_f)
    # the true sums lie anywhere in [x, x (1 + 1/top_min)): corners, and random points
    worst = np.zeros(n, bool)
    for trial in range(12):
        if trial < 4:   # corners that maximise |P - Q|
            frac = np.array([[1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1], [0, 0, 0, 0]][trial], np.float64)[None, :]
        else:
            frac = rng.random((n, 4))
        y = x + frac * x / top_min * (1 - 1e-12)
        r2 = exact_r2(y)
        keep = np.float32(r2) > np.float32(thr)          # lib.rs:660 on the exact value
        worst |= keep & ~passes
    assert not worst.any(), (int(worst.sum()), x[worst][:5])
    # and it is a screen: it rejects the bulk of clearly independent tables
    indep = exact_r2(x) < thr * 0.5
    assert (passes & indep).sum() < 0.35 * indep.sum()


# ------------------------------------------------------------------------------------------ GPU
def synth(*a, **k):
    from weightedld_b200.synth import make_alignment
    return make_alignment(*a, **k)


def run(chars, screen, weights=None, thr=THR, ctas=2, partition=None, flags=0, limbs=0):
    import weightedld_b200 as wld
    from weightedld_b200 import _lib as L
    with wld.Context(0) as ctx:
        ctx.set_screen(screen)
        ctx.set_cta_group(ctas)
        ctx.set_limbs(limbs)
        if partition:
            ctx.set_partition(*partition)
        ctx.load_alignment(chars)
        n_kept = ctx.filter_sites()
        if weights is None:
            ctx.henikoff()
        else:
            ctx.set_weights(weights)
        n, done = ctx.ld_pairs(thr)
        return {"pairs": ctx.fetch_pairs(n, flags), "done": done, "info": ctx.pair_info(), "n_kept": n_kept,
                "wq": ctx.pair_weights(), "site_map": ctx.site_map(),
                "ms": {k: ctx.stage_ms(L.STAGE_NAMES.index(k)) for k in ("pair_prep", "pair_sample", "pair", "pair_refine")}}


@pytest.mark.gpu
@pytest.mark.parametrize("n_seqs,n_cols,thr,ctas", [(300, 700, 0.1, 2), (1000, 900, 0.05, 1), (2049, 600, 0.3, 2),
                                                     (513, 1500, 0.01, 2), (5000, 1300, 0.1, 2), (130, 3000, 0.8, 1)])
def test_forced_screen_is_byte_identical_and_matches_oracle(oracle, n_seqs, n_cols, thr, ctas):
    chars = synth(n_seqs, n_cols, seed=n_seqs + n_cols, block=60)
    a = run(chars, "never", thr=thr, ctas=ctas)
    b = run(chars, "always", thr=thr, ctas=ctas)
    assert a["info"].screen == 0 and b["info"].screen == 1 and b["info"].screen_top_min >= 128
    assert a["done"] == b["done"] == a["n_kept"] * (a["n_kept"] - 1) // 2
    assert (len(a["pairs"]) > 0 or thr > 0.1) and a["pairs"].tobytes() == b["pairs"].tobytes()
    assert len(b["pairs"]) <= b["info"].screen_candidates < 0.5 * b["done"]
    assert np.array_equal(a["wq"], b["wq"])
    if n_seqs * n_cols <= 1_000_000:  # the f64 oracle on the same integers, every pair
        fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
        ref, computed = oracle.all_weighted_ld_pairs(fs, b["wq"], thr, oracle.F64)
        assert computed == b["done"] and len(ref) == len(b["pairs"])
        assert np.array_equal(b["pairs"]["site_a"], ref["a"]) and np.array_equal(b["pairs"]["site_b"], ref["b"])
        for f in ("d", "d_prime", "r2"):
            assert np.array_equal(b["pairs"][f].view(np.uint32), ref[f].view(np.uint32)), f


@pytest.mark.gpu
def test_screen_with_user_weights_and_four_limbs():
    from weightedld_b200.synth import make_weights
    chars = synth(1500, 1200, seed=77, block=80)
    w = make_weights(1500, seed=3)            # U(0,1): spans more than 2^8 -> 4 limbs, gains up to 2^7
    a = run(chars, "never", weights=w)
    b = run(chars, "always", weights=w)
    assert a["info"].n_limbs == b["info"].n_limbs == 4
    if b["info"].screen_top_min >= 32:
        assert b["info"].screen == 1
    assert a["pairs"].tobytes() == b["pairs"].tobytes() and len(a["pairs"]) > 0
    # mild weights, three limbs
    w3 = (0.5 + 0.5 * w).astype(np.float32)
    a, b = run(chars, "never", weights=w3), run(chars, "always", weights=w3)
    assert b["info"].n_limbs == 3 and b["info"].screen == 1 and a["pairs"].tobytes() == b["pairs"].tobytes()


@pytest.mark.gpu
def test_screen_is_refused_when_its_bound_is_not_valid():
    chars = synth(800, 900, seed=5, block=70)
    w = np.full(800, 1.0, np.float32)
    w[::7] = 1e-4                              # far below 2^-7 of the maximum: the top limb of those weights is ~0
    a, b = run(chars, "never", weights=w), run(chars, "always", weights=w)
    assert b["info"].screen_top_min < 32 and b["info"].screen == 0
    assert a["pairs"].tobytes() == b["pairs"].tobytes()
    # all-equal weights are one limb already; a non-positive threshold makes every pair a candidate
    u = run(chars, "always", weights=np.ones(800, np.float32))
    assert u["info"].screen == 0 and u["info"].n_limbs == 1
    z = run(chars, "always", thr=0.0)
    assert z["info"].screen == 0
    # explicit limb counts: two limbs still screen, one limb has nothing to drop
    assert run(chars, "always", limbs=2)["info"].screen == 1
    assert run(chars, "always", limbs=1)["info"].screen == 0


@pytest.mark.gpu
def test_automatic_choice_low_and_high_ld():
    lo = synth(1000, 6000, seed=11)                                  # ~0.02 % of the pairs pass
    a, b = run(lo, "never"), run(lo, "auto")
    assert b["info"].screen == 1 and b["info"].sample_pairs > 100_000
    assert b["info"].sample_candidates * 512 <= b["info"].sample_pairs
    assert a["pairs"].tobytes() == b["pairs"].tobytes() and len(a["pairs"]) > 100
    hi = synth(1000, 6000, seed=12, founders=64, block=400, clonal=True, stray=0.02, private_rate=0.002)
    a, b = run(hi, "never"), run(hi, "auto")
    assert b["info"].screen == 0 and b["info"].sample_candidates * 256 > b["info"].sample_pairs
    assert b["info"].sample_tiles_flagged * 2 > b["info"].sample_tiles      # candidates everywhere: no point in screening
    assert len(a["pairs"]) > 0.05 * a["done"] and a["pairs"].tobytes() == b["pairs"].tobytes()
    # a small problem (fewer than four waves of cells) is not worth a sample: exact kernel
    small = synth(500, 1500, seed=13)
    assert run(small, "auto")["info"].screen == 0


@pytest.mark.gpu
def test_candidate_buffer_overflow_repeats_the_screen():
    hi = synth(800, 3000, seed=21, founders=32, block=500, clonal=True, stray=0.02, private_rate=0.002)
    a, b = run(hi, "never"), run(hi, "always")
    assert b["info"].screen == 1 and b["info"].screen_candidates > 65536 + b["done"] // 64 and b["info"].screen_reruns >= 1
    assert a["pairs"].tobytes() == b["pairs"].tobytes() and len(a["pairs"]) > 100_000


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [2, 5])
def test_partitions_may_mix_screen_and_exact_kernel(nparts):
    import weightedld_b200 as wld
    chars = synth(700, 4000, seed=31, block=100)
    whole = run(chars, "never")
    kept = wld.FETCH_KEPT_INDEX | wld.FETCH_UNORDERED
    shards, done = [], 0
    for p in range(nparts):
        r = run(chars, "always" if p % 2 == 0 else "never", partition=(p, nparts), flags=kept, ctas=2 - (p % 3 == 2))
        assert r["info"].screen == (1 if p % 2 == 0 else 0)
        shards.append(r["pairs"])
        done += r["done"]
    assert done == whole["done"]
    merged = wld.merge_shards(whole["n_kept"], shards, whole["site_map"])
    assert merged.tobytes() == whole["pairs"].tobytes()


@pytest.mark.gpu
def test_refinement_gives_up_when_the_sample_was_blind(monkeypatch):
    """A heterogeneous input can hide its high-LD region from the sample.  The refinement then finds far more
    candidates than refining one by one is worth, declines, and the exact kernel runs over the same pairs."""
    hi = synth(1000, 6000, seed=12, founders=64, block=400, clonal=True, stray=0.02, private_rate=0.002)
    a = run(hi, "never")
    monkeypatch.setenv("WLD_SAMPLE_BLIND", "1")
    b = run(hi, "auto")
    assert b["info"].screen == 0 and b["info"].screen_candidates > b["done"] // 128   # screened, gave up, exact kernel
    assert a["pairs"].tobytes() == b["pairs"].tobytes() and len(a["pairs"]) > 0.05 * a["done"]
    # the same in two partitions
    import weightedld_b200 as wld
    kept = wld.FETCH_KEPT_INDEX | wld.FETCH_UNORDERED
    shards = [run(hi, "auto", partition=(p, 2), flags=kept)["pairs"] for p in range(2)]
    assert wld.merge_shards(a["n_kept"], shards, a["site_map"]).tobytes() == a["pairs"].tobytes()


@pytest.mark.gpu
def test_block_ld_goes_through_the_screen_and_the_exact_kernel_on_flagged_cells():
    """LD confined to blocks along the diagonal: too many candidates to recompute pair by pair, but in a minority of the
    128 x 128-site cells.  The automatic mode screens everything with one limb and runs the exact kernel on the flagged
    cells only (wld_pair_info.screen == 2) — same bytes as the exact kernel over every pair."""
    import weightedld_b200 as wld
    chars = synth(1000, 6500, seed=51, founders=6, block=300)
    a = run(chars, "never")
    b = run(chars, "auto")
    bi = b["info"]
    assert bi.sample_candidates * 512 > bi.sample_pairs and bi.sample_tiles_flagged * 5 <= bi.sample_tiles * 2
    assert bi.screen == 2 and 0 < bi.screen_cells_flagged * 2 <= bi.screen_cells
    assert a["done"] == b["done"] == a["n_kept"] * (a["n_kept"] - 1) // 2
    assert len(a["pairs"]) > 0.005 * a["done"] and a["pairs"].tobytes() == b["pairs"].tobytes()
    # fewer tiles than the exact kernel over everything
    assert bi.tiles < 0.5 * a["info"].tiles
    # partitions: each decides on its own (two halves are still long enough schedules for a sample)
    kept = wld.FETCH_KEPT_INDEX | wld.FETCH_UNORDERED
    shards, done, modes = [], 0, []
    for p in range(2):
        r = run(chars, "auto", partition=(p, 2), flags=kept)
        shards.append(r["pairs"])
        done += r["done"]
        modes.append(r["info"].screen)
    assert done == a["done"] and modes == [2, 2]
    assert wld.merge_shards(a["n_kept"], shards, a["site_map"]).tobytes() == a["pairs"].tobytes()


@pytest.mark.gpu
def test_medium_candidate_lists_are_costed_after_the_screen():
    """Between "a handful" and "too many": the refinement first declines a list that is not short, and the host then
    picks the cheapest way to finish from the exact counts — here (weak LD spread over many cells) the pair-by-pair
    refinement after all.  Same bytes as the exact kernel."""
    chars = synth(1500, 6000, seed=61, founders=24, block=900)
    a = run(chars, "never", thr=0.02)
    b = run(chars, "auto", thr=0.02)
    bi = b["info"]
    assert a["pairs"].tobytes() == b["pairs"].tobytes() and len(a["pairs"]) > 1000
    if bi.screen == 1 and bi.screen_candidates > 65536:
        assert bi.screen_cells > 0          # it was costed on the host (flags were read back) and refined pair by pair

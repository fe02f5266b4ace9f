"""GPU parity tests (run with -m gpu on a B200).  Every test drives the product through the C ABI
(libwld.so via ctypes) and checks it against the oracle on the same seeded inputs:

  * stage 1 (alphabet, histograms, site mask, major/minor, kept code matrix): BIT-EXACT
  * stage 2 (Henikoff weights): rtol 1e-9 against the f64 oracle (f32 view: 1 ulp-ish, 2e-7)
  * stage 3 (D, D', r2, surviving pair set, output order): BIT-EXACT against the f64 restatement of
    lib.rs:455-521 run on the same fixed-point weights (quantisation bound tested in test_oracle.py),
    and within 2e-5 absolute of the reference-faithful f32 restatement away from the threshold.
"""
import math

import numpy as np
import pytest

from conftest import fasta_chars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wld():
    import weightedld_b200 as w
    return w


def synth(*a, **k):
    from weightedld_b200.synth import make_alignment
    return make_alignment(*a, **k)


def assert_pairs_identical(gpu, ref):
    assert len(gpu) == len(ref), (len(gpu), len(ref))
    assert np.array_equal(gpu["site_a"], ref["a"]) and np.array_equal(gpu["site_b"], ref["b"])
    for f in ("d", "d_prime", "r2"):
        same = (gpu[f].view(np.uint32) == ref[f].view(np.uint32)) | (np.isnan(gpu[f]) & np.isnan(ref[f]))
        assert same.all(), (f, int((~same).sum()), gpu[~same][:5], ref[~same][:5])


# ------------------------------------------------------------------------------------------ stage 1
STAGE1_SHAPES = [(1, 1), (2, 3), (5, 7), (37, 300), (130, 129), (1000, 2500), (2051, 1111), (4100, 515)]


@pytest.mark.parametrize("n_seqs,n_cols", STAGE1_SHAPES)
def test_stage1_bit_exact(wld, oracle, n_seqs, n_cols):
    chars = synth(n_seqs, n_cols, seed=n_seqs * 7 + n_cols, block=50, variable_frac=0.6, newline_col=True,
                  lowercase_frac=0.3)
    chars[::3, ::5] = np.frombuffer(b"RYKMnX*.\r ", np.uint8)[(np.arange(chars[::3, ::5].size) % 10)].reshape(
        chars[::3, ::5].shape)  # IUPAC codes, junk, CR: all Unknown (lib.rs:61)
    ss = oracle.siteset_from_chars(chars)
    with wld.Context(0) as ctx:
        ctx.load_alignment(chars)
        for params in ((0.8, 0.02, 0.5), (0.5, 0.1, 0.4), (0.0, 0.0, 1.0)):
            n_kept = ctx.filter_sites(*params)
            fs = oracle.filter_sites(ss, *params)
            assert np.array_equal(ctx.histograms().astype(np.uint64), ss.hists)
            assert n_kept == fs.n_sites
            assert np.array_equal(ctx.site_map(), fs.site_map)
            maj, mnr = ctx.major_minor()
            omaj, omnr = fs.major_minor()
            assert np.array_equal(maj, omaj) and np.array_equal(mnr, omnr)
            assert np.array_equal(ctx.codes(), fs.codes)
        assert ctx.keep_all_sites() == n_cols + 1
        assert np.array_equal(ctx.codes(), ss.codes)


def test_stage1_device_input_unaligned_and_codes(wld, oracle):
    import torch
    chars = synth(777, 1001, seed=9, block=64)  # odd pitch -> generic (byte-load) kernels
    ss = oracle.filter_sites(oracle.siteset_from_chars(chars))
    with wld.Context(0) as ctx:
        ctx.load_alignment(torch.from_numpy(chars).cuda())
        assert ctx.filter_sites() == ss.n_sites
        assert np.array_equal(ctx.site_map(), ss.site_map) and np.array_equal(ctx.codes(), ss.codes)
        # same data as 0..5 codes plus out-of-range values, as the VCF path delivers them
        codes = oracle.encode(chars)
        codes[5, 7] = 9
        codes[0, 0] = 255
        ctx.load_alignment(torch.from_numpy(codes).cuda(), codes=True)
        ref = oracle.siteset_from_codes(np.minimum(codes, 5).T)
        ctx.keep_all_sites()
        assert np.array_equal(ctx.histograms().astype(np.uint64), ref.hists)
        assert np.array_equal(ctx.codes(), ref.codes)


def test_borrowed_device_buffer_exact_extent_and_stream_order(wld, oracle):
    """include/wld.h, WLD_INPUT_DEVICE: a borrowed buffer is only readable up to (n_seqs-1)*row_stride + n_cols
    (the vector kernels must not touch the pitch padding of the LAST row), and a context without an explicit
    stream adopts the torch stream that produced the tensor (non_blocking H2D copy just before the load)."""
    import torch
    n_seqs, n_cols, stride = 513, 1000, 1024  # 16-byte pitch -> vector kernels; n_cols % 16 != 0
    chars = synth(n_seqs, n_cols, seed=33, block=64)
    ss = oracle.filter_sites(oracle.siteset_from_chars(chars))
    host = torch.zeros((n_seqs - 1) * stride + n_cols, dtype=torch.uint8).pin_memory()
    view = host.as_strided((n_seqs, n_cols), (stride, 1))
    view.copy_(torch.from_numpy(chars))
    for _ in range(3):
        with wld.Context(0) as ctx:  # no set_stream
            dev = host.to("cuda", non_blocking=True)  # exact extent: the allocation ends with the last row's data
            ctx.load_alignment(dev.as_strided((n_seqs, n_cols), (stride, 1)))
            assert ctx.filter_sites() == ss.n_sites
            assert np.array_equal(ctx.histograms().astype(np.uint64), oracle.siteset_from_chars(chars).hists)
            assert np.array_equal(ctx.site_map(), ss.site_map) and np.array_equal(ctx.codes(), ss.codes)


@pytest.mark.parametrize("n_seqs,n_cols", [(3, 5), (700, 1201), (3000, 9000)])
def test_load_alignment_rows_by_pointer(wld, oracle, n_seqs, n_cols):
    """wld_load_alignment_rows (the reference's Vec<Sequence>, lib.rs:148-156): rows scattered in host memory
    (here: slices of a FASTA-like buffer with name lines of varying length in between) give the same site set as
    the contiguous matrix; the large case goes through the multi-threaded pinned staging."""
    chars = synth(n_seqs, n_cols, seed=n_seqs, block=64, newline_col=True)
    blob, offs = bytearray(), []
    for r in range(n_seqs):
        blob += b">" + b"s" * (r % 7) + b"\n"
        offs.append(len(blob))
        blob += chars[r].tobytes()
    buf = np.frombuffer(bytes(blob), np.uint8)
    rows = [buf[o:o + chars.shape[1]] for o in offs]
    ss = oracle.filter_sites(oracle.siteset_from_chars(chars))
    with wld.Context(0) as ctx:
        ctx.load_alignment_rows(rows)
        assert ctx.filter_sites() == ss.n_sites
        assert np.array_equal(ctx.site_map(), ss.site_map) and np.array_equal(ctx.codes(), ss.codes)
        assert np.array_equal(ctx.histograms().astype(np.uint64), oracle.siteset_from_chars(chars).hists)


def test_filter_edge_cases(wld, oracle):
    # all-Unknown, invariant, exact ties at the filter bounds (f32 compares, lib.rs:328-331)
    rows = ["ANAC-A", "ANAC-C", "ANCC-A", "ANCA-C"]
    chars = np.frombuffer("".join(rows).encode(), np.uint8).reshape(4, -1)
    ss = oracle.siteset_from_chars(chars)
    with wld.Context(0) as ctx:
        ctx.load_alignment(chars)
        for params in ((0.8, 0.02, 0.5), (0.0, 0.5, 0.5), (0.74, 0.25, 0.25), (1.0, 0.0, 1.0), (-1.0, 0.0, 1.0)):
            n = ctx.filter_sites(*params)
            fs = oracle.filter_sites(ss, *params)
            assert n == fs.n_sites and np.array_equal(ctx.site_map(), fs.site_map), params


# ------------------------------------------------------------------------------------------ stage 2
@pytest.mark.parametrize("n_seqs,n_cols,clonal", [(5, 40, False), (300, 700, False), (1500, 2600, True), (4099, 900, True)])
def test_henikoff_matches_f64_oracle(wld, oracle, n_seqs, n_cols, clonal):
    from weightedld_b200.synth import make_sarscov2_like
    chars = make_sarscov2_like(n_seqs, n_cols, seed=n_seqs) if clonal else synth(n_seqs, n_cols, seed=n_seqs, block=80)
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    with wld.Context(0) as ctx:
        ctx.load_alignment(chars)
        ctx.filter_sites()
        ctx.henikoff()
        w64, w32 = ctx.weights_f64(), ctx.weights()
    ref64 = oracle.henikoff_weights(fs, f64=True)
    assert np.allclose(w64, ref64, rtol=1e-9, atol=0)            # the stated tolerance
    assert w64.max() == 1.0 and np.array_equal(w32, w64.astype(np.float32))
    assert np.allclose(w32, oracle.henikoff_weights(fs), rtol=3e-4)  # reference-faithful f32 (its own noise)
    if clonal:
        assert w64.max() / w64.min() > 30  # "weight-heavy": spans decades


def test_henikoff_rust_kats(wld, golden):
    for case in golden["rust_kat"]["henikoff"]:  # lib.rs:731-750, on unfiltered SiteSets
        w = wld.henikoff_weights(wld.SiteSet.from_strs(case["rows"]))
        tol = 1e-6 if case["tol"] == "ulps" else case["tol"]
        assert np.allclose(w, case["weights"], atol=tol, rtol=0), case["src"]


# ------------------------------------------------------------------------------------------ stage 3
def run_gpu_pairs(wld, chars, kernel, thr, n_limbs=3, weights=None, partition=None, cap=None, filt=(0.8, 0.02, 0.5),
                  ctas=2, gain_bits=-1):
    """-> (pairs, pairs computed, pair_info, f32 weights, pairs with kept indices, integer weights q)"""
    with wld.Context(0) as ctx:
        ctx.set_pair_kernel(kernel)
        ctx.set_cta_group(ctas)
        ctx.set_limbs(n_limbs)
        ctx.set_gain_bits(gain_bits)
        if cap:
            ctx.set_pair_capacity(cap)
        if partition:
            ctx.set_partition(*partition)
        ctx.load_alignment(chars)
        ctx.filter_sites(*filt)
        if weights is None:
            ctx.henikoff()
        else:
            ctx.set_weights(weights)
        n, done = ctx.ld_pairs(thr)
        return ctx.fetch_pairs(n), done, ctx.pair_info(), ctx.weights(), ctx.fetch_pairs(n, 1), ctx.pair_weights()


def oracle_pairs(oracle, chars, w32, info, thr, filt=(0.8, 0.02, 0.5), wq_gpu=None):
    """f64 restatement of lib.rs:455-521 on the fixed-point weights the library reports it used
    (mantissa bits + gain bits of wld_pair_info); wq_gpu: the library's own integers, which must be those."""
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars), *filt)
    wq = oracle.quantize_weights(w32, info.weight_bits, info.gain_bits)
    if wq_gpu is not None:
        assert np.array_equal(wq, wq_gpu)
    pairs, computed = oracle.all_weighted_ld_pairs(fs, wq, thr, oracle.F64)
    return fs, pairs, computed


PAIR_CASES = [  # n_seqs, n_cols, thr
    (10, 30, -1.0), (64, 200, 0.1), (300, 700, 0.1), (1000, 900, -1.0), (2049, 600, 0.05), (130, 1300, 0.2),
]


@pytest.mark.parametrize("n_seqs,n_cols,thr", PAIR_CASES)
def test_pairs_simt_bit_exact(wld, oracle, n_seqs, n_cols, thr):
    chars = synth(n_seqs, n_cols, seed=n_seqs + n_cols, block=60, clonal=True)
    gpu, done, info, w32, _, wq = run_gpu_pairs(wld, chars, "simt", thr)
    fs, ref, computed = oracle_pairs(oracle, chars, w32, info, thr, wq_gpu=wq)
    assert info.kernel == 1 and info.weight_bits == 24
    assert done == computed == fs.n_sites * (fs.n_sites - 1) // 2
    assert_pairs_identical(gpu, ref)


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("kernel", ["bf16", "i8"])
@pytest.mark.parametrize("n_limbs", [3, 1, 2, 4])
@pytest.mark.parametrize("n_seqs,n_cols,thr", PAIR_CASES)
def test_pairs_umma_bit_exact(wld, oracle, n_seqs, n_cols, thr, n_limbs, kernel, ctas):
    chars = synth(n_seqs, n_cols, seed=n_seqs + n_cols, block=60, clonal=True)
    gpu, done, info, w32, _, wq = run_gpu_pairs(wld, chars, kernel, thr, n_limbs=n_limbs, ctas=ctas)
    assert info.kernel == {"bf16": 0, "i8": 2}[kernel] and info.n_limbs == n_limbs and info.weight_bits == 8 * n_limbs
    assert 0 <= info.gain_bits <= min(7, info.weight_span_log2)
    fs, ref, computed = oracle_pairs(oracle, chars, w32, info, thr, wq_gpu=wq)
    assert done == computed
    assert_pairs_identical(gpu, ref)


def test_pairs_close_to_reference_faithful_f32(wld, oracle):
    """Against the f32 restatement of the Rust build (the reference's own arithmetic): |delta| <= 2e-5
    on d and r2 (f32 accumulation noise of the reference ~ sqrt(N)*6e-8 plus 2^-25 weight
    quantisation), identical pair set except pairs whose r2 is within 2e-5 of the threshold."""
    chars = synth(1200, 800, seed=21, block=60, clonal=True)
    thr = 0.1
    gpu, _, _, w32, _, _ = run_gpu_pairs(wld, chars, "umma", thr)
    fs = oracle.filter_sites(oracle.siteset_from_chars(chars))
    ref, _ = oracle.all_weighted_ld_pairs(fs, w32, -1.0, oracle.F32_SCALAR)
    refmap = {(int(p["a"]), int(p["b"])): p for p in ref}
    got = {(int(p["site_a"]), int(p["site_b"])) for p in gpu}
    for p in gpu:
        q = refmap[(int(p["site_a"]), int(p["site_b"]))]
        assert abs(p["d"] - q["d"]) <= 2e-5 and abs(p["r2"] - q["r2"]) <= 2e-5
        assert abs(p["d_prime"] - q["d_prime"]) <= 2e-4 * max(1.0, abs(q["d_prime"]))
    want = {k for k, q in refmap.items() if q["r2"] > thr}
    band = {k for k, q in refmap.items() if abs(q["r2"] - thr) <= 2e-5}
    assert (got ^ want) <= band


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("kernel", ["bf16", "i8"])
def test_umma_matches_simt_all_pairs_multi_tile(wld, kernel, ctas):
    # several M and N tiles, ragged edges, K not a multiple of the K block, every pair emitted
    chars = synth(1111, 1900, seed=77, block=100)
    # (same gain bits on both sides: the automatic choice depends on the accumulator of the kernel)
    a = run_gpu_pairs(wld, chars, kernel, -1.0, ctas=ctas, gain_bits=2)
    b = run_gpu_pairs(wld, chars, "simt", -1.0, gain_bits=2)
    assert a[2].gain_bits == b[2].gain_bits == 2 and np.array_equal(a[5], b[5])
    assert a[1] == b[1] and len(a[0]) == len(b[0]) > 500000
    assert a[0].tobytes() == b[0].tobytes()


def test_fp32_accumulation_exact_at_the_limit(wld):
    """Worst case of the exactness argument: every top limb = 256 (all weights max, one slightly
    smaller so that the single-limb shortcut for equal weights is not taken) and N = 65535, so an
    accumulator reaches 255*65535 = 2^24 - 65791 in fp32.  The tensor path must equal the FP64 SIMT path."""
    n = 65535
    chars = synth(n, 96, seed=1, block=32)
    w = np.ones(n, np.float32)
    w[-1] = 0.75
    a = run_gpu_pairs(wld, chars, "umma", -1.0, weights=w)
    b = run_gpu_pairs(wld, chars, "simt", -1.0, weights=w)
    assert a[2].n_limbs == 3 and a[2].limb_bits == 8 and a[2].gain_bits == 0 and b[2].gain_bits == 0
    assert len(a[0]) > 1000 and a[0].tobytes() == b[0].tobytes()
    c = run_gpu_pairs(wld, chars, "i8", -1.0, weights=w)  # s32 accumulation: exact with room to spare
    assert c[2].limb_bits == 8 and c[0].tobytes() == b[0].tobytes()


@pytest.mark.parametrize("kernel", ["bf16", "i8"])
def test_unweighted_uses_one_limb_and_matches(wld, oracle, kernel):
    chars = synth(500, 400, seed=4, block=50)
    w = np.ones(500, np.float32)  # main.rs:150-153
    gpu, done, info, _, _, wq = run_gpu_pairs(wld, chars, kernel, 0.1, weights=w)
    assert info.n_limbs == 1 and info.weight_bits == 0 and info.gain_bits == 0 and np.all(wq == 1.0)
    fs, ref, computed = oracle_pairs(oracle, chars, w, info, 0.1)
    assert done == computed
    assert_pairs_identical(gpu, ref)


def test_partition_union_equals_whole(wld):
    chars = synth(400, 2500, seed=31, block=80)
    whole = run_gpu_pairs(wld, chars, "umma", 0.1)
    parts = [run_gpu_pairs(wld, chars, "umma", 0.1, partition=(p, 3)) for p in range(3)]
    assert sum(p[1] for p in parts) == whole[1]          # every pair computed exactly once
    n_kept = None
    merged = np.concatenate([p[4] for p in parts])        # kept-index records
    import weightedld_b200 as W
    with W.Context(0) as ctx:
        ctx.load_alignment(chars)
        n_kept = ctx.filter_sites()
    order = np.lexsort((merged["site_b"], merged["site_a"], W.pair_order_key(n_kept, merged["site_a"], merged["site_b"])))
    assert merged[order].tobytes() == whole[4].tobytes()


def test_append_pairs_merges_partitions_on_the_device(wld):
    """wld_append_pairs (multi-GPU merge on rank 0): partition 0's context takes the other partitions' unordered
    KEPT-index shards — once from host memory, once from device memory — and one ordinary fetch returns the
    union in the reference's order, byte-identical to the single-GPU run."""
    import torch
    chars = synth(500, 2300, seed=55, block=70, clonal=True)
    whole = run_gpu_pairs(wld, chars, "i8", 0.1)[0]
    flags = wld.FETCH_KEPT_INDEX | wld.FETCH_UNORDERED
    for on_device in (False, True):
        ctxs = []
        try:
            shards = []
            for p in range(3):
                ctx = wld.Context(0)
                ctxs.append(ctx)
                ctx.set_partition(p, 3)
                ctx.load_alignment(chars)
                ctx.filter_sites()
                ctx.henikoff()
                n, _ = ctx.ld_pairs(0.1)
                shards.append((ctx, n))
            root, n0 = shards[0]
            total = n0
            for ctx, n in shards[1:]:
                root.append_pairs(ctx.fetch_pairs_device(n).clone() if on_device else ctx.fetch_pairs(n, flags))
                total += n
            merged = root.fetch_pairs(total)
            assert len(merged) == len(whole) > 1000 and merged.tobytes() == whole.tobytes()
            # streaming fetch: the concatenation of ranges is the whole result
            parts = [root.fetch_pairs_range(lo, 777) for lo in range(0, total, 777)]
            assert np.concatenate(parts).tobytes() == whole.tobytes()
            assert len(root.fetch_pairs_range(total, 10)) == 0
        finally:
            for ctx in ctxs:
                ctx.close()
    torch.cuda.synchronize()


def test_sharded_stages_are_bit_identical(wld, oracle):
    """Stages 1-2 split over "ranks" (include/wld.h): three contexts on this GPU each count a third of the rows and
    sum a third of the sequences; after the two exchanges (here wld_sum_histograms / wld_share_weight_sums, the
    one-process form) every context holds the same histograms, site set and — bit for bit — the same weights as a
    single context doing everything."""
    import ctypes as C
    chars = synth(1000, 1800, seed=66, block=60, clonal=True)
    lib = wld._lib.load() if hasattr(wld, "_lib") else None
    from weightedld_b200 import _lib as L
    lib = L.load()
    with wld.Context(0) as ref:
        ref.load_alignment(chars)
        k_ref = ref.filter_sites()
        ref.henikoff()
        h_ref, w_ref, sm_ref = ref.histograms(), ref.weights_f64(), ref.site_map()
    ctxs = [wld.Context(0) for _ in range(3)]
    try:
        bounds = [(0, 334), (334, 668), (668, 1000)]
        for ctx, (lo, hi) in zip(ctxs, bounds):
            ctx.set_row_shard(lo, hi)
            ctx.set_seq_shard(lo, hi)
            ctx.load_alignment(chars)
        arr = (C.c_void_p * 3)(*[c._h for c in ctxs])
        assert lib.wld_sum_histograms(arr, 3) == 0
        for ctx in ctxs:
            assert ctx.filter_sites() == k_ref
            assert np.array_equal(ctx.histograms(), h_ref) and np.array_equal(ctx.site_map(), sm_ref)
            ctx.henikoff()
            with pytest.raises(wld.WldError):
                ctx.weights()                      # not whole yet
        assert lib.wld_share_weight_sums(arr, 3) == 0
        for ctx in ctxs:
            ctx.henikoff_finish()
            assert ctx.weights_f64().tobytes() == w_ref.tobytes()
    finally:
        for ctx in ctxs:
            ctx.close()


def test_overflow_protocol_grows_buffer(wld):
    chars = synth(200, 900, seed=8, block=90)
    small = run_gpu_pairs(wld, chars, "umma", -1.0, cap=1024)
    big = run_gpu_pairs(wld, chars, "umma", -1.0)
    assert len(small[0]) == len(big[0]) > 1024 and small[0].tobytes() == big[0].tobytes()


def test_output_order_is_reference_order(wld, oracle):
    chars = synth(64, 1400, seed=12, block=400)  # > 4 reference tiles of 256 per edge
    gpu, _, info, w32, _, wq = run_gpu_pairs(wld, chars, "umma", 0.3)
    fs, ref, _ = oracle_pairs(oracle, chars, w32, info, 0.3, wq_gpu=wq)
    assert fs.n_sites > 1024
    assert_pairs_identical(gpu, ref)  # includes the order (lib.rs:623-679)


# ------------------------------------------------------------------------------- fixtures end to end
RUST_EMULATED = {
    "example": ([0, 1], ["0\t1\t0.107\t0.345\t0.237"]),
    "t2_henikoff_complex1": ([1, 2], ["1\t2\t0.107\t0.357\t0.238"]),
    "t3_henikoff_complex2": ([1, 2], ["1\t2\t0.107\t0.357\t0.238"]),
    "t4_weights1_ld0": ([0, 1, 3], ["0\t3\t0.088\t0.422\t0.192", "1\t3\t0.088\t0.422\t0.192"]),
    "t5_weights1_ld0.25": ([0, 1], ["0\t1\t-0.250\t0.500\t1.000"]),
    "t6_varsites_hk_ld": ([0, 1], ["0\t1\t-0.148\t0.444\t0.400"]),
}


@pytest.mark.parametrize("name", sorted(RUST_EMULATED))
def test_reference_fixtures_end_to_end(wld, golden, tmp_path, name):
    """Config 1 (tests/example.fasta) and the other FASTA fixtures through the mirrored Rust API,
    main.rs:129-209 with default flags, written with the reference's TSV writers."""
    f = tmp_path / f"{name}.fasta"
    f.write_bytes(golden["fixtures"][name].encode())
    ms = wld.read_fasta(f)
    siteset = wld.SiteSet.from_multiseq(ms)
    filtered = siteset.filter_by(0.8, 0.02, 0.5)
    kept, lines = RUST_EMULATED[name]
    assert filtered.site_map().tolist() == kept
    weights = wld.henikoff_weights(filtered)
    store = wld.all_weighted_ld_pairs(filtered, weights, 0.1)
    wld.write_pair_stats(tmp_path / "pairs.tsv", store)
    assert (tmp_path / "pairs.tsv").read_text().splitlines() == ["site_a\tsite_b\td\td'\tr2"] + lines


def test_rust_pair_kats(wld, oracle, golden):
    for case in golden["rust_kat"]["pair"]:  # lib.rs:753-801
        a = oracle.encode(np.frombuffer(case["a"].encode(), np.uint8))
        b = oracle.encode(np.frombuffer(case["b"].encode(), np.uint8))
        r2, d, dp = wld.single_weighted_ld_pair(a, b, np.array(case["w"], np.float32))
        assert abs(d - case["d"]) <= case["tol"] and abs(dp - case["d_prime"]) <= case["tol"]
        assert abs(r2 - case["r2"]) <= case["tol"], case["src"]


def test_t7_vcf_config(wld, golden):
    """Config 2: tests/t7_1000genome.vcf (5008 haplotypes x 5 sites as the Python reader delivers
    them) against the Python reference's printed D, D', R2 with the Python weights."""
    t7 = golden["t7"]
    py = golden["python_ref"]["t7_1000genome"]
    ss = wld.SiteSet.from_codes(np.ascontiguousarray(t7["python_alignment"]))
    store = wld.all_weighted_ld_pairs(ss, t7["python_weights"].astype(np.float32), -1.0)
    assert len(store) == 10
    pos = t7["pos"]
    for (a, b, (r2, d, dp)), line in zip(store, py["ld_stdout"][1:]):
        pa, pb, pd, pdp, pr2 = line.split("\t")
        assert (pos[a], pos[b]) == (int(pa), int(pb))
        assert abs(d - float(pd)) <= 6e-5 and abs(dp - float(pdp)) <= 6e-5 and abs(r2 - float(pr2)) <= 6e-5


# ---------------------------------------------------------------------------------- error behaviour
def test_errors_are_loud(wld):
    with wld.Context(0) as ctx:
        with pytest.raises(wld.WldError) as e:
            ctx.filter_sites()
        assert e.value.status == 2
        chars = synth(20, 30, seed=1, block=10)
        ctx.load_alignment(chars)
        with pytest.raises(wld.WldError):
            ctx.henikoff()
        ctx.filter_sites()
        with pytest.raises(wld.WldError):
            ctx.ld_pairs(0.1)
        with pytest.raises(wld.WldError):
            ctx.set_weights(np.ones(19, np.float32))
        ctx.set_weights(np.full(20, -1.0, np.float32))
        with pytest.raises(wld.WldError) as e:
            ctx.ld_pairs(0.1)
        assert "weights" in str(e.value)
        with pytest.raises(wld.WldError):
            ctx.set_limbs(5)
    with pytest.raises(wld.WldError):
        wld.Context(99)


def test_empty_and_degenerate_inputs(wld):
    with wld.Context(0) as ctx:
        ctx.load_alignment(np.zeros((3, 0), np.uint8))
        assert ctx.filter_sites() == 0
        ctx.henikoff()
        assert np.isnan(ctx.weights()).all()  # 0/0, lib.rs:355 with no sites
        assert ctx.ld_pairs(0.1) == (0, 0)
        one = np.frombuffer(b"AAACC", np.uint8).reshape(5, 1)
        ctx.load_alignment(one)
        assert ctx.filter_sites() == 1
        ctx.henikoff()
        assert ctx.ld_pairs(0.1) == (0, 0)
    progress = []
    with wld.Context(0) as ctx:
        ctx.load_alignment(synth(50, 300, seed=2, block=30))
        k = ctx.filter_sites()
        ctx.henikoff()
        ctx.ld_pairs(0.1, progress.append)
    assert progress[0] == 0 and progress[-1] == k * (k - 1) // 2 and progress == sorted(progress)


def test_two_contexts_in_two_threads_are_independent(wld):
    """A context is not thread-safe, but distinct contexts are independent (include/wld.h): two host threads
    drive two contexts on the same GPU concurrently and get the single-thread results."""
    import threading
    inputs = [synth(600, 3000, seed=s, block=90, clonal=True) for s in (101, 202)]
    want = [run_gpu_pairs(wld, c, "i8", 0.1)[0].tobytes() for c in inputs]
    got, errs = [None, None], []

    def work(k):
        try:
            for _ in range(3):
                got[k] = run_gpu_pairs(wld, inputs[k], "i8", 0.1)[0].tobytes()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs and got == want


def test_host_ordering_fallback_matches_device_ordering(wld, monkeypatch):
    """When the device has no room for the ordering scratch, wld_fetch_pairs orders the survivors on the host
    (same order: reference tile, then a, then b; parent indices).  WLD_FORCE_HOST_ORDER=1 takes that path."""
    chars = synth(500, 2600, seed=77, block=80)
    with wld.Context(0) as ctx:
        ctx.load_alignment(chars)
        ctx.filter_sites()
        ctx.henikoff()
        n, _ = ctx.ld_pairs(0.05)
        dev = ctx.fetch_pairs(n).copy()
        dev_kept = ctx.fetch_pairs(n, wld.FETCH_KEPT_INDEX).copy()
        monkeypatch.setenv("WLD_FORCE_HOST_ORDER", "1")
        host = ctx.fetch_pairs(n).copy()
        host_kept = ctx.fetch_pairs(n, wld.FETCH_KEPT_INDEX).copy()
        with pytest.raises(wld.WldError):          # a range of the ordered result needs the device scratch
            ctx.fetch_pairs_range(1, 10)
    assert n > 2000 and dev.tobytes() == host.tobytes() and dev_kept.tobytes() == host_kept.tobytes()

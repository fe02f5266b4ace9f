/*
 * wld.h — C ABI of libwld.so, the B200-native (sm_100a) implementation of WeightedLD's hot path:
 *   encode + site filter  ->  Henikoff weights  ->  all-pairs weighted LD + r2 filter.
 *
 * The reference (ojcharles/WeightedLD) has no FFI; its seam is the Rust crate's `pub` API that
 * rust/weighted_ld/src/main.rs:12 consumes (`use weighted_ld::*`).  Every entry point below names
 * the reference item (file:line under the reference root) it replaces.  The header is plain C
 * (C types only, opaque context, POD structs, int status codes) so that `bindgen` can bind it
 * from a Rust `weighted_ld-sys` crate (INTEGRATION.md shows the stub).
 *
 * Contract
 *  - One context per run and per GPU.  A context is not thread-safe; calls block until done.
 *  - The caller owns every host buffer it passes; the library owns all device memory.  Results
 *    are copied into caller buffers; no pointer into device or pinned memory is ever returned.
 *  - Every function returns a wld_status.  On failure wld_last_error(ctx) holds a message.
 *    There is NO CPU fallback: a CUDA failure is an error, never a silent downgrade.
 *  - Stage order (main.rs:129-190): load -> filter (or keep_all) -> henikoff | set_weights ->
 *    ld_pairs -> fetch_pairs.  Calling out of order returns WLD_ERR_STATE.
 */
#ifndef WLD_H
#define WLD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WLD_ABI_VERSION 3

#if defined(__GNUC__)
#define WLD_API __attribute__((visibility("default")))
#else
#define WLD_API
#endif

typedef struct wld_ctx wld_ctx;

typedef enum wld_status {
  WLD_OK = 0,
  WLD_ERR_INVALID = 1,     /* bad argument */
  WLD_ERR_STATE = 2,       /* stage called out of order */
  WLD_ERR_CUDA = 3,        /* CUDA runtime / driver failure, device missing, wrong architecture */
  WLD_ERR_NOMEM = 4,       /* host or device allocation failed */
  WLD_ERR_UNSUPPORTED = 5, /* size beyond the stated limits of the exact-accumulation scheme */
  WLD_ERR_PANIC = 6        /* the reference would panic here (lib.rs:180-182 unequal lengths) */
} wld_status;

/* Symbol codes, lib.rs:20-29: A=0 C=1 G=2 T=3 Missing('-')=4 Unknown(anything else)=5. */
enum { WLD_SYM_A = 0, WLD_SYM_C = 1, WLD_SYM_G = 2, WLD_SYM_T = 3, WLD_SYM_MISSING = 4, WLD_SYM_UNKNOWN = 5 };

/* One surviving site pair: lib.rs:523-527 (PairData{first_idx, second_idx, data}) with
 * lib.rs:382-387 (LdStats).  20 bytes, little-endian, no padding. */
typedef struct wld_pair {
  uint32_t site_a;  /* first_idx  (lib.rs:662) */
  uint32_t site_b;  /* second_idx (lib.rs:663) */
  float d;          /* lib.rs:502 */
  float d_prime;    /* lib.rs:516 */
  float r2;         /* lib.rs:518 */
} wld_pair;

/* wld_load_alignment flags */
enum {
  WLD_INPUT_ASCII = 0,   /* bytes are characters; encoded by lib.rs:53-64 on the GPU */
  WLD_INPUT_CODES = 1,   /* bytes are already 0..5 codes (values > 5 are read as 5 = Unknown);
                            the VCF path of WeightedLD.py:311-379 produces these */
  WLD_INPUT_DEVICE = 2   /* `data` is a device pointer on the context's GPU (borrowed until the
                            next load/destroy); otherwise a host pointer (copied).  ORDERING: the
                            buffer is read on the context's stream (wld_set_stream; by default a
                            private non-blocking stream), which is NOT ordered after the stream that
                            filled it — either pass that stream with wld_set_stream before loading, or
                            synchronise it first.  EXTENT: readable for (n_seqs-1)*row_stride + n_cols
                            bytes; nothing beyond that is touched. */
};

/* wld_fetch_pairs flags */
enum {
  WLD_FETCH_PARENT_INDEX = 0, /* site_a/site_b are raw alignment columns (lib.rs:662-663) */
  WLD_FETCH_KEPT_INDEX = 1,   /* site_a/site_b index the filtered site set (for merging shards) */
  WLD_FETCH_UNORDERED = 2,    /* skip the sort into the reference's output order */
  WLD_FETCH_DEVICE = 4        /* `out` is a device pointer on the context's GPU (e.g. to hand a shard to NCCL) */
};

/* pair-kernel selection for wld_set_pair_kernel */
enum {
  WLD_PAIR_KERNEL_UMMA = 0,   /* tcgen05/TMEM Gram tiles fed by TMA, bf16 limbs x fp32 accumulate (kind::f16) */
  WLD_PAIR_KERNEL_SIMT = 1,   /* CUDA-core FP64 kernel, same exact sums and epilogue; verification path */
  WLD_PAIR_KERNEL_UMMA_I8 = 2 /* same tcgen05 kernel, u8 limbs x s32 accumulate (kind::i8): half the operand
                                 bytes and energy per MAC, exact for any n_seqs; the DEFAULT */
};

/* stage ids for wld_stage_ms */
enum {
  WLD_STAGE_LOAD = 0,      /* H2D copy of the alignment (0 for device input) */
  WLD_STAGE_HISTOGRAM = 1, /* per-column 6-bin histogram (part of wld_load_alignment) */
  WLD_STAGE_FILTER = 2,    /* site decision + ordered compaction + encode/transpose of kept sites */
  WLD_STAGE_HENIKOFF = 3,
  WLD_STAGE_PAIR_PREP = 4, /* weight quantisation + indicator / limb operand expansion */
  WLD_STAGE_PAIR = 5,      /* Gram + epilogue + compaction kernel(s) */
  WLD_STAGE_ORDER = 6,     /* survivors into the reference's output order + parent indices (first fetch after a pair stage) */
  WLD_STAGE_PAIR_SAMPLE = 7, /* the one-limb screen over a sample of the tiles (decides screen + refine vs. the exact kernel) */
  WLD_STAGE_PAIR_REFINE = 8, /* exact sums + statistics of the screen's candidates */
  WLD_STAGE_COUNT = 9
};

/* Progress callback of all_weighted_ld_pairs (lib.rs:582): receives the number of site pairs
 * finished so far; first call is 0 on the caller's thread (lib.rs:584); counts never decrease. */
typedef void (*wld_progress_fn)(uint64_t pairs_computed, void* user);

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Creates a context on CUDA device `device`.  Fails with WLD_ERR_CUDA when no sm_100 GPU is
 * present.  (The reference has no setup step; this replaces process start in main.rs:121.) */
WLD_API int wld_create(int device, wld_ctx** out);
WLD_API void wld_destroy(wld_ctx* ctx);
WLD_API const char* wld_last_error(const wld_ctx* ctx);
WLD_API int wld_abi_version(void);

/* Run all work on this CUDA stream (a cudaStream_t passed as void*); NULL = the context's own (a private
 * non-blocking stream).  To run on the legacy default stream — whose handle is also 0 — pass the explicit
 * handle cudaStreamLegacy ((cudaStream_t)0x1). */
WLD_API int wld_set_stream(wld_ctx* ctx, void* cuda_stream);

/* Shard the pair stage: this context computes part `part` of `nparts` of the upper-triangular
 * tile grid (load-balanced, no collective).  Replaces rayon's tile fan-out, lib.rs:635-637. */
WLD_API int wld_set_partition(wld_ctx* ctx, int part, int nparts);

/* Weight limbs of the exact split Gram: 1..4, or 0 = automatic (the default): 3 limbs (24-bit
 * mantissa), 4 when some nonzero weight is below 2^-8 of the largest.  Limbs are 8 bits wide, or
 * narrower when n_seqs is so large that fp32 accumulation of 8-bit limbs could round (bf16 kernel
 * only) — see DESIGN.md "precision contract". */
WLD_API int wld_set_limbs(wld_ctx* ctx, int n_limbs);
/* Block-exponent ("gain") bits G of the fixed-point weights: a weight u = w/max(w) is stored as a
 * B-bit mantissa m = rint(u * 2^e * (2^B - 1)), e = clamp(-exponent(u), 0, G), and enters every sum
 * as the integer q = m * 2^(G-e); the power of two rides in the 0/1 indicator operand, so it costs
 * no tensor work.  Every weight >= 2^-G * max keeps B relative bits, as the reference's f32 weights
 * do (lib.rs:469-479).  0..7, or -1 = automatic (the default): just enough for the smallest
 * nonzero weight, lowered if a Gram entry could leave the exact range of the accumulator. */
WLD_API int wld_set_gain_bits(wld_ctx* ctx, int gain_bits);
/* Width of one limb: 1..8, or 0 = automatic (the default): 8, lowered only when the fp32 accumulator of
 * the bf16 kernel could otherwise round (a limb column sum above 2^24, possible from n_seqs > 65 793). */
WLD_API int wld_set_limb_bits(wld_ctx* ctx, int limb_bits);
WLD_API int wld_set_pair_kernel(wld_ctx* ctx, int kind);
/* Numeric dialect.  WLD_COMPAT_RUST (default): the Rust crate, normative for this library.
 * WLD_COMPAT_PYTHON: the reference's WeightedLD.py where the two differ —
 *  - Henikoff contributions 1/count with the code-5 fill = sum / known sequences (WeightedLD.py:132-145;
 *    the scalar `unique_base` cancels in the max-normalisation);
 *  - alleles are called per PAIR after deleting the sequences with code 5 at either site
 *    (WeightedLD.py:183-211): pairs where that can differ from the per-site call are recomputed by a
 *    per-pair kernel, all others run on the tensor cores as usual;
 *  - pairs whose PA or PB rounds to 1.0 at one decimal are skipped (WeightedLD.py:234-237).
 * The Python program applies no r2 threshold (pass -INFINITY) and filters sites with
 * wld_filter_sites_python.  Pairs with an empty marginal, printed as nan by Python, are dropped. */
enum { WLD_COMPAT_RUST = 0, WLD_COMPAT_PYTHON = 1 };
WLD_API int wld_set_compat(wld_ctx* ctx, int mode);
/* CTAs cooperating on one tensor-core tile of the pair stage: 2 (default) = tcgen05 cta_group::2, a
 * CTA pair computes 256 x 256 and each CTA stages half of the limb operand; 1 = single-CTA 128 x 256. */
WLD_API int wld_set_cta_group(wld_ctx* ctx, int ctas);
/* Initial capacity (in pairs) of the device survivor buffer; it grows automatically. */
WLD_API int wld_set_pair_capacity(wld_ctx* ctx, uint64_t pairs);
/* Screen + refine.  With an r2 threshold most site pairs of a real alignment are far below it, and whether a
 * pair can reach it is decided by far fewer weight bits than its statistics need.  The library can therefore run
 * the Gram with ONE limb (the top 8 bits of every fixed-point weight; a third / a quarter of the tensor work),
 * bound r2 from above rigorously from those sums (the truncated part of every weight is below 1/top_min of what
 * was summed), and recompute only the pairs that bound does not rule out — exactly, from the code matrix and the
 * full integer weights, through the same f64 statistics of lib.rs:482-518.  The survivors are the same records,
 * bit for bit, as the exact n-limb kernel's (all_weighted_ld_pairs, lib.rs:578-684).
 *   mode 0: never; 1 (default): when it pays — the candidate rate is measured on a sample of the tiles first:
 *   up to 1 candidate in 256 pairs -> screen + per-pair refinement; more, but in at most 40 % of the sampled tiles
 *   (LD confined to a band or to blocks) -> screen, then the exact kernel over the 128 x 128-site cells it flags;
 *   else (high LD everywhere) straight to the exact kernel; 2: screen + per-pair refinement whenever the bound is
 *   valid (tests).  The screen needs the u8 kernel, the Rust dialect, a positive threshold, at least
 *   two limbs and top_min >= 32; otherwise the exact kernel runs.  wld_pair_info.screen tells which ran.
 *   Should the screen of mode 1 turn up far more candidates than its sample promised (above 1 pair in 128: a
 *   heterogeneous input), the refinement declines and the exact kernel takes over: on the 128 x 128-site cells in
 *   which the screen found a candidate when those are a minority (wld_pair_info.screen = 2: LD confined to a band or
 *   to blocks), else on all pairs (screen = 0 with screen_candidates > 0). */
/* (The environment variable WLD_SCREEN=0|1|2 sets the initial mode of every context: A/B runs of whole programs.) */
WLD_API int wld_set_screen(wld_ctx* ctx, int mode);

/* ---- stage 1: encode + histogram + filter -------------------------------------------------- */
/* SiteSet::from_multiseq (lib.rs:176-206) for an alignment given as n_seqs rows of n_cols bytes,
 * row pitch row_stride bytes (sequence-major, as read_fasta lib.rs:277-307 produces it,
 * newline column included if the caller follows lib.rs:297).  Computes the 6-bin histogram of
 * every column (lib.rs:98-104).  flags: WLD_INPUT_*. */
WLD_API int wld_load_alignment(wld_ctx* ctx, const uint8_t* data, int64_t n_seqs, int64_t n_cols,
                       int64_t row_stride, int flags);

/* The same for sequences held one by one, as the reference's MultiSequence does (Vec<Sequence>, lib.rs:148-156)
 * and as a FASTA file lies in memory (rows[r] may point into a mapped file, lib.rs:277-307): n_seqs host
 * pointers to n_cols bytes each.  The rows are gathered through pinned staging buffers by several host threads
 * while earlier chunks are already on the bus — no contiguous host copy of the alignment is ever made. */
WLD_API int wld_load_alignment_rows(wld_ctx* ctx, const uint8_t* const* rows, int64_t n_seqs, int64_t n_cols, int flags);

/* SiteSet::filter_by (lib.rs:230-251) with is_site_of_interest (lib.rs:310-338) and the
 * threshold of main.rs:139: keep a site iff acgt > ceil(f32(min_acgt)*f32(n_seqs)) and a major
 * and a minor symbol exist and min_minor <= minor/(minor+major) <= max_minor (f32).  Builds the
 * kept, site-major 0..5 code matrix on the device. */
WLD_API int wld_filter_sites(wld_ctx* ctx, float min_acgt, float min_minor, float max_minor, int64_t* n_kept);
/* compute_variable_sites (WeightedLD.py:44-98), the site filter of the Python program, all in f64:
 * keep a site iff acgt/n_seqs > min_acgt and (known symbols other than the most frequent one, gaps
 * included) / (known symbols) >= min_variability.  Same outputs as wld_filter_sites. */
WLD_API int wld_filter_sites_python(wld_ctx* ctx, double min_acgt, double min_variability, int64_t* n_kept);
/* Use every column unfiltered (the reference's unit tests call henikoff_weights and
 * single_weighted_ld_pair on unfiltered SiteSets, lib.rs:731-801). */
WLD_API int wld_keep_all_sites(wld_ctx* ctx, int64_t* n_kept);

WLD_API int64_t wld_n_seqs(const wld_ctx* ctx);   /* SiteSet::n_seqs  lib.rs:259 */
WLD_API int64_t wld_n_cols(const wld_ctx* ctx);   /* SiteSet::n_sites lib.rs:254 of the unfiltered set */
WLD_API int64_t wld_n_kept(const wld_ctx* ctx);   /* SiteSet::n_sites lib.rs:254 of the filtered set */
/* site_map, lib.rs:169 / parent_site_index lib.rs:263: out[k] = raw column of kept site k. */
WLD_API int wld_get_site_map(wld_ctx* ctx, int64_t* out, int64_t cap);
/* Histograms of ALL raw columns, out[col*6 + sym] (SymbolHistogram, lib.rs:72-104). */
WLD_API int wld_get_histograms(wld_ctx* ctx, uint32_t* out, int64_t cap_cols);
/* major_minor_symbols (lib.rs:126-140) of the kept sites; -1 = None. */
WLD_API int wld_get_major_minor(wld_ctx* ctx, int8_t* major, int8_t* minor, int64_t cap);
/* Kept code matrix, site-major: out[k*n_seqs + seq] (SiteSet::site_symbols, lib.rs:267). */
WLD_API int wld_get_codes(wld_ctx* ctx, uint8_t* out, int64_t cap_bytes);

/* ---- stages 1-2 on several GPUs ---------------------------------------------------
 * Without these calls every GPU of a multi-GPU run repeats stages 1-2 on the whole alignment.  With them
 * rank r counts only its rows and sums only its sequences; two small exchanges make every rank whole:
 *   wld_set_row_shard(r's rows) -> wld_load_alignment -> SUM of the histograms over ranks (u32, exact)
 *   -> wld_filter_sites -> wld_set_seq_shard(r's sequences) -> wld_henikoff -> ALL-GATHER of the weight sums
 *   -> wld_henikoff_finish.
 * Every sequence is summed whole by exactly one GPU in the single-GPU order, so all results stay
 * bit-identical for any number of GPUs.  The exchange runs outside the library: over NCCL on the buffers
 * wld_exchange_buffer names (one process per GPU), or by wld_sum_histograms / wld_share_weight_sums (one
 * process driving several GPUs; peer copies).  This replaces nothing in the reference (it is single-node
 * rayon, lib.rs:98-104,340-358); it shards the same arithmetic. */
WLD_API int wld_set_row_shard(wld_ctx* ctx, int64_t row_lo, int64_t row_hi); /* row_hi < 0: all rows */
WLD_API int wld_set_seq_shard(wld_ctx* ctx, int64_t seq_lo, int64_t seq_hi); /* seq_hi < 0: all sequences */
enum {
  WLD_EXCHANGE_HISTOGRAM = 0,   /* u32 [5][n_cols rounded up to 16]: sum over ranks, in place, after loading */
  WLD_EXCHANGE_WEIGHT_SUMS = 1  /* f64 [n_seqs]: rank r owns [seq_lo, seq_hi), the rest is zero: all-gather in place, or
                                   sum over ranks (x + 0 is exact), after wld_henikoff */
};
/* Device address and size of an exchange buffer (the ONE exception to "no device pointers leave the library":
 * a collective has to run on it).  Valid until the next stage call on the context. */
WLD_API int wld_exchange_buffer(wld_ctx* ctx, int which, void** device_ptr, uint64_t* bytes);
WLD_API int wld_henikoff_finish(wld_ctx* ctx);  /* max + normalisation (lib.rs:355) after the all-gather */
WLD_API int wld_sum_histograms(wld_ctx* const* ctxs, int n);
WLD_API int wld_share_weight_sums(wld_ctx* const* ctxs, int n);

/* ---- stage 2: sequence weights ------------------------------------------------------------- */
/* henikoff_weights (lib.rs:340-358) + henikoff_site_contributions (lib.rs:360-380) over the
 * kept sites; f64 accumulation on the GPU, results held as f64 and as f32 (the reference type). */
WLD_API int wld_henikoff(wld_ctx* ctx);
/* Caller-supplied weights (main.rs:150-153 passes all-ones for --unweighted). Finite, >= 0, max > 0. */
WLD_API int wld_set_weights(wld_ctx* ctx, const float* weights, int64_t n);
WLD_API int wld_get_weights(wld_ctx* ctx, float* out, int64_t cap);
WLD_API int wld_get_weights_f64(wld_ctx* ctx, double* out, int64_t cap);

/* ---- stage 3: all-pairs weighted LD -------------------------------------------------------- */
/* all_weighted_ld_pairs (lib.rs:578-684) with single_weighted_ld_pair (lib.rs:390-521): every
 * pair b > a of kept sites (of this context's partition), D / D' / r2, keep r2 > r2_threshold
 * (strict, NaN dropped, lib.rs:660).  Survivors stay on the device until fetched.
 * n_survivors receives their count; pairs_computed (may be NULL) the number of pairs evaluated. */
WLD_API int wld_ld_pairs(wld_ctx* ctx, float r2_threshold, wld_progress_fn progress, void* user,
                 uint64_t* n_survivors, uint64_t* pairs_computed);
/* main.rs:129-190 in one call: load -> filter -> Henikoff weights (weights == NULL) or the caller's ->
 * all pairs.  Same results as the four stage calls; the host waits for the device three times in total
 * (kept-site count, weight quantisation decision, survivor count) plus once for a host input buffer. */
WLD_API int wld_run(wld_ctx* ctx, const uint8_t* data, int64_t n_seqs, int64_t n_cols, int64_t row_stride, int flags,
                    float min_acgt, float min_minor, float max_minor, const float* weights, float r2_threshold,
                    int64_t* n_kept, uint64_t* n_survivors, uint64_t* pairs_computed);
/* PairStore::iter (lib.rs:533-575): copies the survivors into out[0..cap) in the reference's
 * output order (tile rows bottom-up, columns ascending, then a, then b — lib.rs:623-679).
 * flags: WLD_FETCH_*. */
WLD_API int wld_fetch_pairs(wld_ctx* ctx, wld_pair* out, uint64_t cap, int flags, uint64_t* n_written);
/* The same for survivors [first, first+count) of that order: lets a writer (main.rs:82-119) stream the
 * result in chunks, formatting one while the next is copied.  The ordering runs once per pair stage and
 * set of flags; every call copies its slice.  Large copies into pageable memory are staged through pinned
 * buffers by several host threads. */
WLD_API int wld_fetch_pairs_range(wld_ctx* ctx, uint64_t first, uint64_t count, wld_pair* out, int flags,
                                  uint64_t* n_written);
/* Multi-GPU merge (replaces the order-preserving rayon collect of lib.rs:635-679 across devices): adds the
 * survivors of OTHER partitions — records with KEPT indices, as WLD_FETCH_KEPT_INDEX|WLD_FETCH_UNORDERED
 * delivers them, in host memory or (src_is_device) on this context's GPU — to this context's own, so that
 * wld_fetch_pairs returns the union in the reference's order.  The contexts must hold the same site set. */
WLD_API int wld_append_pairs(wld_ctx* ctx, const wld_pair* src, uint64_t n, int src_is_device);
/* The same for a host that drives several GPUs from one process: appends the survivors of `other` (another
 * partition of the same site set, on any GPU of the box) with one peer copy over NVLink. */
WLD_API int wld_append_pairs_from(wld_ctx* ctx, wld_ctx* other);
/* The integer weights q[s] the last wld_ld_pairs summed (exact in a double); the statistics of
 * lib.rs:482-518 are invariant to their common scale.  For verification against an oracle. */
WLD_API int wld_get_pair_weights(wld_ctx* ctx, double* out, int64_t cap);
/* Sort key of the reference's output order for a pair of KEPT indices (for merging shards):
 * lexicographic (key, a, b) ascending == reference order. */
WLD_API uint64_t wld_pair_order_key(int64_t n_kept, uint32_t kept_a, uint32_t kept_b);
/* The pair-stage schedule, host only (no GPU needed): the upper-triangular tile list of partition
 * `part` of `nparts` exactly as wld_ld_pairs runs it (replaces rayon's fan-out over triu_index,
 * lib.rs:623-637).  Partitions are contiguous, equally long ranges of a canonical list of 128 x 128-site
 * cells (strips of 8 cell columns), so the site pairs a partition owns do not depend on the kernel variant.
 * A tile of a variant covers kept sites [tm*TM, tm*TM+TM) x [j_lo, j_hi), TM = 64*cta_group, with
 * [j_lo, j_hi) the part of [tn*TN, tn*TN+TN), TN = 2*floor(128/(2*n_limbs)), that belongs to the partition;
 * only pairs a < b inside it are evaluated (a tile that straddles a partition boundary appears in both
 * partitions with disjoint windows).  n_limbs = 1 is the schedule of the one-limb screen.  Writes 4 uint32
 * (tm, tn, j_lo, j_hi) per tile into tiles (may be NULL to count), n_tiles = number of tiles, n_pairs =
 * site pairs they cover.  sm_count is unused (kept for ABI stability). */
WLD_API int wld_plan_tiles(int64_t n_kept, int n_limbs, int cta_group, int part, int nparts, int sm_count, uint32_t* tiles,
                           uint64_t cap_tiles, uint64_t* n_tiles, uint64_t* n_pairs);

/* The schedule of "screen + exact kernel on the flagged cells" (wld_set_screen), host only: given one flag per tile of
 * the screen's schedule of this partition (wld_plan_tiles with n_limbs = 1, same order), the exact kernel's tiles of
 * `n_limbs` limbs that cover the flagged tiles, windows clipped to them.  Same output convention as wld_plan_tiles. */
WLD_API int wld_plan_cell_tiles(int64_t n_kept, int n_limbs, int cta_group, int part, int nparts, const uint8_t* flags,
                                uint64_t n_flags, uint32_t* tiles, uint64_t cap_tiles, uint64_t* n_tiles, uint64_t* n_pairs);

/* ---- introspection ------------------------------------------------------------------------- */
/* Device time of the last run of a stage in milliseconds (CUDA events on the context's stream). */
WLD_API int wld_stage_ms(wld_ctx* ctx, int stage, float* ms);
/* Kernel launches issued by the last run of a stage. */
WLD_API int wld_stage_launches(wld_ctx* ctx, int stage, int* launches);
/* Geometry of the last pair stage: limbs used, bits per limb, total weight bits, K padding,
 * tiles computed, executed tensor flop. */
typedef struct wld_pair_info {
  int32_t kernel;          /* WLD_PAIR_KERNEL_* actually run */
  int32_t n_limbs;
  int32_t limb_bits;
  int32_t weight_bits;     /* n_limbs * limb_bits */
  int64_t k_padded;        /* sequences padded to the MMA K block */
  int64_t tiles;           /* output tiles computed by this partition */
  int64_t tile_sites_m;    /* kept sites per tile along a */
  int64_t tile_sites_n;    /* kept sites per tile along b */
  double executed_flop;    /* 2*M*N*K summed over MMA instructions issued */
  int32_t die_schedule;    /* 0: plain round-robin tile schedule; 1/2: die-aware (SM -> L2 die map measured) */
  int32_t die_sms[2];      /* SMs found on each L2 die (0, 0 when the map could not be established) */
  int32_t gain_bits;       /* G, see wld_set_gain_bits */
  int32_t weight_span_log2;/* x: the smallest nonzero weight lies in [2^-(x+1), 2^-x) of the largest */
  int32_t screen;          /* 0: every pair went through the exact n-limb kernel; 1: one-limb screen + exact
                              refinement of its candidates; 2: one-limb screen + the exact kernel over the 128 x 128-site
                              cells in which it found candidates (same survivors, bit for bit; see wld_set_screen) */
  double weight_rel_err;   /* realised max over nonzero weights of |q/scale - w/max| / (w/max).  A priori:
                              <= 2^-B when x <= G, else <= 2^(x-G-B).  Every weighted sum of lib.rs:469-479
                              (all terms >= 0) carries at most this relative error before the f64 epilogue. */
  /* screen + refine bookkeeping (all 0 when the screen was not considered) */
  int64_t screen_candidates;   /* site pairs the screen could not rule out (each recomputed exactly) */
  int64_t sample_pairs;        /* site pairs of the sampling launch that chose the path */
  int64_t sample_candidates;   /* ... and how many of them were candidates */
  int32_t screen_top_min;      /* smallest top limb of a nonzero weight: the screen's bound is x <= y <= x (1 + 1/top_min) */
  int32_t screen_reruns;       /* times the screen was repeated because the candidate buffer was too small */
  int64_t screen_cells;        /* screen = 2 (or a screen that gave up): cells (screen tiles) of this partition ... */
  int64_t screen_cells_flagged;/* ... and how many of them held a candidate */
  int64_t sample_tiles;        /* tiles of the sampling launch ... */
  int64_t sample_tiles_flagged;/* ... and how many of them held a candidate */
} wld_pair_info;
WLD_API int wld_get_pair_info(wld_ctx* ctx, wld_pair_info* out);

#ifdef __cplusplus
}
#endif
#endif /* WLD_H */

// common.cuh — context, error plumbing and launch declarations shared by the libwld.so sources.
// Product code: never includes or links anything under oracle/.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wld.h"

namespace wld {

constexpr int kNumSMsB200 = 148;

// ---------------------------------------------------------------------------------------------
// device buffer with explicit ownership
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t n) {
    if (n <= bytes && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, n ? n : 1);
    if (e == cudaSuccess) bytes = n ? n : 1;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

enum class Stage : int { Created = 0, Loaded = 1, Filtered = 2, Weighted = 3, Paired = 4 };

struct StageTimer {
  cudaEvent_t beg = nullptr, end = nullptr;
  bool valid = false;
  int launches = 0;
  float carry_ms = 0.f;   // time of earlier launches of the same stage in this call (a screen before the cell pass)
  int carry_launches = 0;
};

// Pair-stage operand geometry (see DESIGN.md "data layout in HBM").
struct PairGeom {
  int n_limbs = 3;       // NL
  int limb_bits = 8;     // b
  int rows_per_site = 6; // RPS = 2*NL   rows of the limb operand per kept site
  int sites_per_group = 21; // SPG = floor(128 / RPS) kept sites per 128-row group of the limb operand
  int elem_bytes = 2;    // operand element: 2 = bf16, 1 = u8
  int gain_bits = 0;     // G: block-exponent bits carried by the indicator operand (pair_prep.cu)
  int64_t k_padded = 0;  // sequences rounded up to one K block (64 bf16 / 128 u8)
  int64_t a_rows = 0;    // indicator operand rows, padded to 128 (2 rows per site)
  int64_t b_groups = 0;  // 128-row groups of the limb operand
};

// What quantize_kernel (pair_prep.cu) decided and measured; read back once per pair stage.
struct QuantDecision {
  int32_t flags;       // bit0: invalid weights (negative, NaN, inf, all zero); bit1: all equal; bit2: cannot be made exact
  int32_t n_limbs, limb_bits, gain_bits;
  int32_t span_log2;   // x: the smallest nonzero weight lies in [2^-(x+1), 2^-x) of the maximum
  int32_t top_min;     // smallest TOP limb among the nonzero fixed-point weights (0: some weight has an empty top limb)
  double weight_sum;   // sum of q
  double rel_err;      // max over nonzero weights of |q / (2^G (2^B - 1)) - u| / u  (realised)
  unsigned long long limb_sums[4];
  float kappa;         // (1 + 1/top_min)^2 - 1 + 1e-5, rounded up: the screen's widening of |P - Q| (ld_screen_f32)
  float pad2;
};

// A tile schedule of the tcgen05 pair kernel kept on the device (pair_umma.cu, ensure_tile_plan).
struct DevPlan {
  int64_t key[5] = {-1, -1, -1, -1, -1};  // n_kept, n_limbs, part, nparts, cta_group
  DevBuf tiles;                            // uint4 [n_tiles]: {M tile, N tile, first site j, end site j}
  int64_t n_tiles = 0;
  uint64_t pairs = 0;
  int64_t x[2] = {0, -1}, y[2] = {0, -1};  // min / max M-tile and N-tile index
  int64_t tile_m = 0, tile_n = 0;
  std::vector<uint4> host;                 // the same list on the host (the cell plan is cut out of it)
};

}  // namespace wld

struct wld_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t poll_stream = nullptr;   // progress polling while the pair kernel runs (created on first use)
  cudaStream_t stream = nullptr;
  wld::Stage stage = wld::Stage::Created;
  std::string err;

  // options
  int part = 0, nparts = 1;
  int n_limbs_opt = 0;             // 0 = automatic (3, or 4 for weights spanning more than 2^8)
  int gain_opt = -1;               // -1 = automatic
  int limb_bits_opt = 0;           // 0 = automatic (8, narrower only when the fp32 accumulator requires it)
  int pair_kernel = WLD_PAIR_KERNEL_UMMA_I8;  // fastest exact path on sm_100a; bf16 and SIMT selectable
  uint64_t pair_cap_opt = 0;
  int compat = WLD_COMPAT_RUST;    // numeric dialect (wld_set_compat)
  int cta_group = 2;               // CTAs cooperating on one MMA tile (tcgen05 cta_group::1 / ::2)
  int die_aware = 1;               // give each L2 die its own part of the tile list (needs die_map); cleared on mismatch
  wld::DevBuf die_of_sm;           // u8 [sm_count]
  int die_used = 0;                // the last pair launch ran the die-aware schedule

  // stage 1
  int64_t n_seqs = 0, n_cols = 0, row_stride = 0;
  int input_flags = 0;
  const uint8_t* d_raw = nullptr;  // borrowed or = raw_own.p
  wld::DevBuf raw_own;
  int64_t cols_padded = 0;         // n_cols rounded up to 16
  wld::DevBuf hist;                // u32 [6][cols_padded]
  wld::DevBuf keep;                // u8  [cols_padded]
  wld::DevBuf rank;                // i32 [cols_padded]  exclusive prefix of keep
  wld::DevBuf maj_raw, min_raw;    // i8  [cols_padded]
  wld::DevBuf site_map;            // i32 [n_kept]
  wld::DevBuf maj, mnr;            // i8  [n_kept]
  wld::DevBuf kept_count;          // i32 [1]
  int64_t n_kept = 0;
  int64_t ldc = 0;                 // code row pitch: n_seqs rounded up to 128
  wld::DevBuf codes;               // u8 [n_kept][ldc], site-major, pad = 5

  // multi-GPU shards of stages 1-2 (wld_set_row_shard / wld_set_seq_shard); hi < 0 = everything
  int64_t row_lo = 0, row_hi = -1;
  int64_t seq_lo = 0, seq_hi = -1;
  bool weights_partial = false;    // wld_henikoff ran on a shard: exchange + wld_henikoff_finish still to come

  // stage 2
  wld::DevBuf table;               // f64 [n_kept][8]  per-site contribution per code (6 used)
  wld::DevBuf partial;             // f64 [site_chunks][n_seqs]
  wld::DevBuf w64;                 // f64 [n_seqs]
  wld::DevBuf w32;                 // f32 [n_seqs]
  wld::DevBuf scalars;             // f64 [8] scratch (max etc.)

  // stage 3
  wld::PairGeom geom;
  wld::DevBuf q;                   // f64 [ldc]      fixed-point weights (integers <= 2^32)
  wld::DevBuf limbs;               // u16 [4][ldc] limb values, then u8 [4][ldc] the same as bytes
  wld::DevBuf gain8;               // u8 [ldc]       per-sequence gain 2^(G-e) carried by the indicator operand
  wld::DevBuf quant;               // QuantDecision (device)
  wld::QuantDecision* quant_host = nullptr;  // pinned mirror
  unsigned long long* sample_host = nullptr; // pinned: {candidates, pairs, flagged tiles} of the sampling launch (behind quant_host)
  int quant_span_log2 = 0;
  double quant_rel_err = 0.0;
  wld::DevBuf opA;                 // bf16 [a_rows][k_padded]
  wld::DevBuf opB;                 // bf16 [b_groups*128][k_padded]
  wld::DevBuf simt_tiles;          // uint2 [n_tiles] of the CUDA-core verification kernel
  wld::DevPlan plans[3];           // tcgen05 schedules: [0] the exact kernel (n_limbs limbs), [1] the one-limb screen,
                                   // [2] the exact kernel over the cells the screen flagged (rebuilt per run)
  wld::DevBuf cell_flags;          // u8 [screen tiles]: 1 = the screen found a candidate in this tile
  wld::DevBuf sample_flags;        // u32 [sampled tiles]: the same for the sampling launch
  int64_t sample_tiles = 0;        // tiles of the last sampling launch
  // screen + refine (pair_umma.cu kScreen, pair_refine.cu)
  int screen_opt = 1;              // wld_set_screen: 0 never, 1 automatic (sampled candidate rate), 2 always when valid
  wld::DevBuf glimb;               // u16 [4][ldc]   gain x limb per sequence and limb (pair_refine.cu)
  wld::DevBuf opB1;                // u8 [b1_groups*128][k_padded]: indicator x TOP limb, one-limb layout (64 sites / group)
  wld::DevBuf cand;                // uint2 [cand_cap] candidate site pairs of the screen
  uint64_t cand_cap = 0;
  int quant_top_min = 0;
  wld::DevBuf pairs;               // wld_pair [pair_cap]
  wld::DevBuf counters;            // u64 [4]: survivors, pairs_done, ...
  wld::DevBuf py_aux;              // uint2 [n_kept] {n5, margin}: WLD_COMPAT_PYTHON only (pair_python.cu)
  wld::DevBuf sorted, sort_keys, sort_idx, sort_temp;  // output ordering scratch (pair_order.cu)
  int sorted_key = -1;             // what `sorted` currently holds: bit0 ordered, bit1 parent indices; -1 nothing
  // pinned staging for large copies to / from pageable host memory (wld_api.cu, staged_copy)
  static constexpr int kMaxStagers = 16;
  void* stage_buf[kMaxStagers] = {};
  cudaStream_t stage_stream[kMaxStagers] = {};
  int n_stagers = 0;
  uint64_t pair_cap = 0;
  uint64_t auto_cap_budget = 0;    // default survivor capacity (pairs), from the free memory at the first pair stage
  uint64_t n_survivors = 0;
  uint64_t pairs_computed = 0;
  wld_pair_info info{};
  float last_thr = 0.f;
  uint64_t plan_pairs = 0;         // site pairs of this partition (identical for every schedule of it)
  double weight_sum = 0.0;         // sum of the fixed-point weights q (upper bound of every pair's T)

  wld::StageTimer timers[WLD_STAGE_COUNT];

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define WLD_CUDA(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return (ctx)->fail(_e == cudaErrorMemoryAllocation ? WLD_ERR_NOMEM : WLD_ERR_CUDA,         \
                         "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace wld {

struct ScopedStageTimer {
  wld_ctx* c;
  int id;
  ScopedStageTimer(wld_ctx* ctx, int stage) : c(ctx), id(stage) {
    StageTimer& t = c->timers[id];
    if (!t.beg) {
      cudaEventCreate(&t.beg);
      cudaEventCreate(&t.end);
    }
    t.valid = false;
    t.launches = 0;
    cudaEventRecord(t.beg, c->stream);
  }
  void launched(int n = 1) { c->timers[id].launches += n; }
  ~ScopedStageTimer() {
    StageTimer& t = c->timers[id];
    cudaEventRecord(t.end, c->stream);
    t.valid = true;
  }
};

// ---- stage launchers (each in its own .cu) ------------------------------------------------------
int run_histogram(wld_ctx* c, ScopedStageTimer& tm);                       // encode_filter.cu
// mode 0: is_site_of_interest (lib.rs:310-338); 1: keep every column; 2: compute_variable_sites
// (WeightedLD.py:44-98) with the f64 thresholds py_min_acgt / py_min_variability
int run_filter(wld_ctx* c, int mode, float min_acgt, float min_minor, float max_minor, double py_min_acgt,
               double py_min_variability, ScopedStageTimer& tm);           // encode_filter.cu
int run_henikoff(wld_ctx* c, ScopedStageTimer& tm, bool finish);           // henikoff.cu
int run_henikoff_finish(wld_ctx* c, ScopedStageTimer& tm);                 // henikoff.cu: max + normalise
int run_pair_prep(wld_ctx* c, ScopedStageTimer& tm, bool try_screen);      // pair_prep.cu: quantise, indicator operand (+ the screen's operand)
int finish_quant(wld_ctx* c);                                              // pair_prep.cu: read the quantiser's decision back (synchronises)
int run_expand_limbs(wld_ctx* c, ScopedStageTimer& tm, bool screen);       // pair_prep.cu: limb operand (all limbs / top limb only)
// The pair launchers bracket ONLY the kernel launch with the WLD_STAGE_PAIR timer (host-side
// planning and the tile-list upload happen before the start event).
int run_pair_simt(wld_ctx* c, float thr);                                  // pair_simt.cu
// mode 0: exact kernel (all limbs); 1: one-limb screen over every tile, candidates -> c->cand, flagged tiles ->
// c->cell_flags; 2: the screen over a sample of the tiles, counting only (counters[8..9]); 3: exact kernel over plans[2]
int run_pair_umma(wld_ctx* c, float thr, int mode);                        // pair_umma.cu
int ensure_tile_plan(wld_ctx* c, int which);                               // pair_umma.cu (which: 0 exact, 1 screen)
// plans[2] := the exact kernel's tiles (windows clipped) of the screen tiles whose flag is set; *n_flagged = their number
int build_cell_plan(wld_ctx* c, const std::vector<uint8_t>& flags, int64_t* n_flagged);  // pair_umma.cu
int run_pair_refine(wld_ctx* c, float thr, unsigned long long give_up);    // pair_refine.cu: exact statistics of the candidates (none if more than give_up)
int run_pair_order(wld_ctx* c, bool ordered, bool parent);                 // pair_order.cu
const std::vector<uint8_t>& die_map(wld_ctx* c);                            // die_map.cu: SM -> L2 die (empty = unknown)
int run_pair_python_prepare(wld_ctx* c);                                   // pair_python.cu
int run_pair_python_fixup(wld_ctx* c, float thr);                          // pair_python.cu

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Upper-triangular tile schedule of the tcgen05 pair kernel (pair_plan.cpp part of wld_api.cu).
struct TilePlan {
  std::vector<uint4> tiles;  // {M tile, N tile, first site j of the tile's window, end site j}
  uint64_t pairs = 0;
  int64_t tile_m = 64, tile_n = 42;
};
TilePlan plan_tiles(int64_t n_kept, int n_limbs, int part, int nparts, int sm_count, int ctas);
// the exact kernel's tiles (n_limbs limbs) covering the flagged tiles of a screen schedule, windows clipped to them
TilePlan cut_cell_plan(const std::vector<uint4>& screen_tiles, const uint8_t* flags, size_t n_flags, int64_t n_kept,
                       int n_limbs, int ctas, int64_t* n_flagged);

}  // namespace wld

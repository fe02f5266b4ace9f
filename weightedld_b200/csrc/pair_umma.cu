// pair_umma.cu — stage 3b: all-pairs weighted LD as a dense Gram on the 5th-generation tensor
// cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA), with the D / D' / r2 /
// threshold / compaction epilogue fused behind it.  sm_100a only.
//
// Reference: all_weighted_ld_pairs lib.rs:578-684 (tile fan-out, b>a, r2 > thr) and
// single_weighted_ld_pair lib.rs:390-521 (the four weighted sums + statistics).
//
// Math.  opA rows (2i+alpha) are the indicators of site i's major (alpha=0) / minor (alpha=1) symbol, times the
// sequence's gain 2^(G-e) (block-exponent weights); opB rows (site j, beta, limb l) are indicator_beta(j) *
// limb_l(weight mantissa); see pair_prep.cu.  One output tile of a CTA pair (cta_group::2; kCtas = 1 halves M) is
//     D[256 x 256] = opA[mi*256 .. +256, :] * opB[nj*256 .. +256, :]^T        (K = sequences)
// i.e. 128 sites i  x  2*SPG sites j.  D[(i,alpha)][(j,beta,l)] is an exact integer (s32 for u8
// operands, fp32 below 2^24 for bf16 operands); the epilogue recombines limbs
//     S = sum_l D_l * 2^(b*(NL-1-l))   and gets
//     AB = S[i,0][j,0]  Ab = S[i,0][j,1]  aB = S[i,1][j,0]  ab = S[i,1][j,1].
//
// Kernel shape (persistent, one CTA per SM because TMEM is fully used, 384 threads, clusters of 2):
//   warp 0      TMA producer (both CTAs): 6-stage ring of {A 128 rows x 128 B, B-half 128 rows x 128 B}
//               (32 KB / stage / CTA), 128B-swizzled; the bytes of both CTAs complete the leader's
//               `full[stage]` mbarrier
//   warp 1      MMA issuer (leader CTA): one lane issues 4 x tcgen05.mma.cta_group::2 (M256 N256 K32
//               for kind::i8, K16 for kind::f16) per stage into one of two 256-column TMEM accumulators;
//               tcgen05.commit (multicast to both CTAs) frees the stage / publishes the tile
//   warp 2      TMEM allocator (512 columns);  warp 3: die-aware schedule set-up
//   warps 4-11  epilogue (each CTA on its own 128 accumulator rows): warp w reads TMEM lanes
//               32*(w%4).. (tcgen05.ld 32x32b) and the column half (w-4)/4 (= one 128-row group of opB);
//               adjacent lanes (alpha=0/1 of a site) swap halves of the 2x2 table by shuffle; fp32
//               conservative pre-filter; candidates are queued and the exact f64 statistics run on full
//               groups of 32; warp-aggregated atomic compaction.
//   The accumulator is double-buffered, so the epilogue of tile t overlaps the MMAs of tile t+1.
//   Tile order: strips of 8 N tiles; each L2 die's CTA pairs walk their own contiguous part of the
//   list (die_map.cu) so the limb strip a die's L2 holds is reused only by that die's SMs.
//
//
// Two roles (template parameter kScreen):
//   exact   NL limbs; the epilogue above; survivors out.
//   screen  ONE limb (opB1 = indicator x TOP limb, 128 x 128-site tiles): the epilogue bounds r2 from above
//           (ld_screen_f32, pair_epilogue.cuh) and compacts the site pairs it cannot rule out as candidates for
//           pair_refine.cu; a sampling launch of it (tile stride, K-block stride) only counts them.
//
// Roofline: tensor pipe.  Algorithmic op per site pair per launch = 8*N per limb pass (4 weighted dot
// products of length N): NL passes in the exact role, one in the screen; executed = 2*256*256*Kp per tile.
#include <climits>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "pair_epilogue.cuh"

namespace wld {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockKBytes = 128;  // one 128-byte swizzle atom along K: 64 bf16 or 128 u8 elements
constexpr int kUmmaKBytes = 32;    // one MMA consumes 32 bytes of K: 16 bf16 (kind::f16) or 32 u8 (kind::i8)
constexpr int kAStageBytes = kBlockM * kBlockKBytes;  // 16 KB per CTA
// Per-epilogue-warp candidate queue (see CandidateQueue): 64 entries of {4 x int64 sums, site_i, site_j}.
#ifndef WLD_EPI_WARPS
#define WLD_EPI_WARPS 8  // 8 or 16 (A/B builds: make EXTRA=-DWLD_EPI_WARPS=16)
#endif
constexpr int kQueueCap = WLD_EPI_WARPS == 16 ? 48 : 64;  // 16 warps: the queues must still fit beside the 6-stage ring
constexpr int kQueueBytesPerWarp = kQueueCap * (4 * 8 + 2 * 4);  // 2560 B
constexpr int kNumEpiWarpsC = WLD_EPI_WARPS;
// kCtas = 1: one CTA computes a 128 x 256 tile and stages all 256 B rows (32 KB) -> 4 stages of 48 KB.
// kCtas = 2: a CTA pair (cta_group::2) computes 256 x 256; each CTA stages its 128 A rows and HALF of
//            the B rows (16 KB) -> 6 stages of 32 KB; one third less L2->SMEM traffic per MMA.
template <int kCtas> struct StageCfg {
  static constexpr int kBRows = kBlockN / kCtas;
  static constexpr int kBBytes = kBRows * kBlockKBytes;
  static constexpr int kBytes = kAStageBytes + kBBytes;
  static constexpr int kStages = kCtas == 1 ? 4 : 6;
  static constexpr int kQueueOffset = kStages * kBytes + 256;  // after the stage ring and the barriers
  static constexpr int kSmemBytes = 1024 /*alignment slack*/ + kQueueOffset + kNumEpiWarpsC * kQueueBytesPerWarp;
};
constexpr int kNumThreads = 128 + 32 * kNumEpiWarpsC;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = kNumEpiWarpsC;
constexpr int kTmemCols = 512;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0

struct UmmaParams {
  const uint4* tiles;   // {M tile, N tile, first site j, end site j}: only pairs with j inside the window count
  int n_tiles;          // tiles this launch walks: tile k of the launch is tiles[k * tile_mul]
  int tile_mul;         // 1, or the stride of a sampling launch
  int k_mul;            // 1, or the stride over K blocks of a sampling launch (an unbiased subsample of the sequences)
  int debug_skip;       // experiments only (WLD_EXPERIMENT_SKIP_EPILOGUE=1): the epilogue releases the accumulator untouched
  int k_blocks;
  int n_kept;
  int limb_bits;
  float thr;
  double thr_lo;
  float thr_lo_f;     // fp32 pre-filter threshold (lowered, see ld_prefilter_f32)
  int thr_negative;   // threshold below zero: every pair with non-empty marginals is a candidate
  int sum_shift;      // fp32 pre-filter works on sums scaled by 2^-sum_shift so that T <= 2^21
  // die-aware schedule (die_of_sm == nullptr: plain round-robin over the whole list).  Tiles [0, die_split) belong
  // to the CTA pairs on die 0, [die_split, n_tiles) to those on die 1; a pair finds its die from %smid and
  // its rank among the die's pairs from die_counter.
  const uint8_t* die_of_sm;
  unsigned int n_sm;          // entries of die_of_sm
  unsigned int* die_counter;  // [2], zeroed before the launch
  int die_pairs[2];
  int die_split;
  int die_mode;               // 1: front/back split of the list; 2: every round of n0+n1 tiles is dealt die 0 first
  // Long-K runs (a wave's panels are far larger than L2): the CTA pairs of a die start every wave of tiles
  // together, so that they stream through K in lockstep and share each panel line while it is in L2.
  int wave_sync;
  unsigned int* wave_counter; // [2]: tiles whose loads have all been issued, per die; zeroed before the launch
  uint64_t hint_a, hint_b;  // L2 eviction policy of the indicator (streamed) and limb (strip-resident) panels
  const uint2* py_aux; // WLD_COMPAT_PYTHON only (else null): per-site {n5, margin}, see py_flagged
  PairOut out;
  // screen (kScreen): candidates instead of survivors; kappa widens |P - Q| by the truncation of the weights
  CandOut cand;
  uint8_t* cell_flags;  // screen, full launch: cell_flags[tile] = 1 when the tile holds a candidate (else null)
  unsigned int* sample_flags;          // sampling launch: one word per sampled tile (first hit counts it) ...
  unsigned long long* sample_flagged;  // ... into this counter: sampled tiles that hold a candidate
  const float* kappa;   // device: QuantDecision::kappa (the sampling launch runs before the host has read it)
  unsigned long long* pairs_done;
  int* error_flag;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000ll) {  // ~10 s at 2 GHz
      if (error_flag) atomicExch(error_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// L2 cache policies for TMA loads (the encodings CUTLASS passes as .L2::cache_hint operands)
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar,
                                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar), "l"(hint)
      : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// remote-capable arrive: `bar` is a shared::cluster address (own CTA's, or CTA 0's after masking)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 2-CTA TMA load: data lands in THIS CTA's smem, the transaction bytes are counted on CTA 0's barrier
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar,
                                                uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar & kPeerBitMask), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of a cta_group::2 MMA: arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (canonical layout
// ((8,n),2):((8,SBO),1) in 16-byte units): start address, LBO = 1 (ignored for swizzled K-major),
// SBO = 1024 B between 8-row groups, descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptors (both operands K-major: bits 15, 16 = 0; N>>3 at bits 17-22, M>>4 at 24-28):
//   kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9 and 10-12 = 1)
//   kind::i8 : D = S32 (bits 4-5 = 2), A = B = unsigned 8-bit (bits 7-9 and 10-12 = 0)
// The instruction M is 128 for one CTA and 256 for a CTA pair (each CTA holds 128 accumulator rows).
template <bool kI8, int kCtas>
__device__ __forceinline__ constexpr uint32_t instr_desc() {
  return (kI8 ? (2u << 4) : ((1u << 4) | (1u << 7) | (1u << 10))) | ((uint32_t)(kBlockN >> 3) << 17) |
         ((uint32_t)((kBlockM * kCtas) >> 4) << 24);
}

template <bool kI8, int kCtas>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  constexpr uint32_t kInstrDescU8 = instr_desc<true, kCtas>();
  constexpr uint32_t kInstrDescBf16 = instr_desc<false, kCtas>();
  if constexpr (kCtas == 2) {
    if constexpr (kI8) {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
          "}\n" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(kInstrDescU8), "r"(accumulate)
          : "memory");
    } else {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
          "}\n" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(kInstrDescBf16), "r"(accumulate)
          : "memory");
    }
  } else if constexpr (kI8) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(kInstrDescU8), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(kInstrDescBf16), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
  static_assert(N == 2 || N == 4 || N == 6 || N == 8 || N == 12 || N == 16, "unsupported column count");
  if constexpr (N == 2) {
    tmem_ld2(taddr, v);
  } else if constexpr (N == 4) {
    tmem_ld4(taddr, v);
  } else if constexpr (N == 6) {
    tmem_ld4(taddr, v);
    tmem_ld2(taddr + 4, v + 4);
  } else if constexpr (N == 8) {
    tmem_ld8(taddr, v);
  } else if constexpr (N == 12) {
    tmem_ld8(taddr, v);
    tmem_ld4(taddr + 8, v + 8);
  } else {
    tmem_ld8(taddr, v);
    tmem_ld8(taddr + 8, v + 8);
  }
}

// ------------------------------------------------------------------------------------------------
// Candidate queue.  The exact f64 statistics (about 170 FP64 instructions, 10 divisions) are far more
// expensive than everything else in the epilogue and B200's FP64 pipe is narrow, so running them
// whenever ANY lane of a warp holds a candidate wastes up to 31/32 of that work.  Each epilogue warp
// instead appends its candidates (exact integer sums + site indices) to a private shared-memory queue
// and runs the f64 path only on full groups of 32, one candidate per lane; the tail is drained when the
// warp has finished its tiles.  FP64 work is then proportional to the number of candidates.
// ------------------------------------------------------------------------------------------------
struct CandidateQueue {
  long long* sums;  // [4][kQueueCap]  AB, Ab, aB, ab
  uint32_t* si;     // [kQueueCap]
  uint32_t* sj;     // [kQueueCap]
  int count;        // warp-uniform

  __device__ __forceinline__ void init(uint8_t* base) {
    sums = reinterpret_cast<long long*>(base);
    si = reinterpret_cast<uint32_t*>(base + 4 * 8 * kQueueCap);
    sj = si + kQueueCap;
    count = 0;
  }
  // all 32 lanes call; lanes with `cand` append one record
  __device__ __forceinline__ void push(bool cand, long long AB, long long Ab, long long aB, long long ab,
                                       uint32_t i, uint32_t j) {
    const unsigned ballot = __ballot_sync(0xffffffffu, cand);
    if (cand) {
      const int pos = count + __popc(ballot & ((1u << (threadIdx.x & 31)) - 1u));
      sums[0 * kQueueCap + pos] = AB;
      sums[1 * kQueueCap + pos] = Ab;
      sums[2 * kQueueCap + pos] = aB;
      sums[3 * kQueueCap + pos] = ab;
      si[pos] = i;
      sj[pos] = j;
    }
    count += __popc(ballot);
    __syncwarp();
  }
  // all 32 lanes call; evaluates the first min(count, 32) records, one per lane, and removes them
  __device__ __forceinline__ void drain32(float thr, const PairOut& out, const uint2* __restrict__ py_aux) {
    const int lane = threadIdx.x & 31;
    const int n = min(count, 32);
    bool keep = lane < n;
    float d = 0.f, dp = 0.f, r2 = 0.f;
    uint32_t i = 0, j = 0;
    if (keep) {
      const double AB = (double)sums[0 * kQueueCap + lane], Ab = (double)sums[1 * kQueueCap + lane];
      const double aB = (double)sums[2 * kQueueCap + lane], ab = (double)sums[3 * kQueueCap + lane];
      i = si[lane];
      j = sj[lane];
      keep = ld_stats_exact(AB, Ab, aB, ab, thr, d, dp, r2, py_aux != nullptr);
      if (py_aux != nullptr && keep) keep = !py_flagged(py_aux, i, j);  // left to pair_python.cu
    }
    emit_pairs_warp(keep, i, j, d, dp, r2, out);
    __syncwarp();
    // move the remaining (< 32) records to the front
    const int rest = count - n;
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    uint32_t ti = 0, tj = 0;
    if (lane < rest) {
      t0 = sums[0 * kQueueCap + n + lane]; t1 = sums[1 * kQueueCap + n + lane];
      t2 = sums[2 * kQueueCap + n + lane]; t3 = sums[3 * kQueueCap + n + lane];
      ti = si[n + lane]; tj = sj[n + lane];
    }
    __syncwarp();
    if (lane < rest) {
      sums[0 * kQueueCap + lane] = t0; sums[1 * kQueueCap + lane] = t1;
      sums[2 * kQueueCap + lane] = t2; sums[3 * kQueueCap + lane] = t3;
      si[lane] = ti; sj[lane] = tj;
    }
    count = rest;
    __syncwarp();
  }
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int NL, bool kI8, int kCtas, bool kScreen>
__global__ void __launch_bounds__(kNumThreads, 1) pair_umma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const UmmaParams p) {
  using Cfg = StageCfg<kCtas>;
  constexpr int kStages = Cfg::kStages;
  constexpr int RPS = 2 * NL;      // opB rows per site
  constexpr int SPG = 128 / RPS;   // sites per 128-row group
  constexpr int kBlockK = kBlockKBytes / (kI8 ? 1 : 2);  // K elements per smem stage

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + kStages * kAStageBytes;
  const uint32_t bars = sB + kStages * Cfg::kBBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * kStages + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * kStages + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * Cfg::kBytes + 8 * (2 * kStages + 4));
  static_assert(8 * (2 * kStages + 4) + 4 <= 256, "barrier block overflows its 256 bytes");
  static_assert(kNumEpiWarps == kNumEpiWarpsC, "queue sizing");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;  // rank 0 = leader: issues the MMAs
  int first_tile = (int)(blockIdx.x / kCtas), tile_step = (int)(gridDim.x / kCtas), tile_end = p.n_tiles;
  int* sched = reinterpret_cast<int*>(smem + kStages * Cfg::kBytes + 8 * (2 * kStages + 4) + 8);  // [4], leader CTA
  static_assert(8 * (2 * kStages + 4) + 8 + 16 <= 256, "barrier block overflows its 256 bytes");
  if (p.die_of_sm != nullptr && cta_rank == 0 && threadIdx.x == 96) {  // warp 3 is otherwise idle
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const bool known = smid < p.n_sm;  // %smid may exceed multiProcessorCount on partitioned devices: no map then
    const int d = known && p.die_of_sm[smid] ? 1 : 0;
    const int k = known ? (int)atomicAdd(&p.die_counter[d], 1u) : INT_MAX;
    int lo, hi, step;
    if (p.die_mode == 2) {
      lo = d ? p.die_pairs[0] : 0;
      hi = p.n_tiles;
      step = p.die_pairs[0] + p.die_pairs[1];
    } else {
      lo = d ? p.die_split : 0;
      hi = d ? p.n_tiles : p.die_split;
      step = p.die_pairs[d];
    }
    if (k >= p.die_pairs[d]) {  // more pairs on this die than the host planned for: report, compute nothing
      atomicExch(p.error_flag, 7);
      sched[0] = hi;
    } else {
      sched[0] = lo + k;
    }
    sched[1] = step;
    sched[2] = hi;
    sched[3] = d;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // leader's expect_tx arrive; TMA bytes of BOTH CTAs complete it
      mbar_init(empty_bar(s), 1);  // one tcgen05.commit (multicast to both CTAs for a pair)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kNumEpiWarps * kCtas);  // the leader waits for the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kCtas == 2) tmem_alloc_cg2(smem_u32(tmem_slot), kTmemCols);
    else tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  }
  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.die_of_sm != nullptr) {
    if (kCtas == 2 && cta_rank != 0) {  // the peer follows its leader: read the leader's schedule over DSMEM
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(sched)));
      asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(first_tile) : "r"(remote));
      asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(tile_step) : "r"(remote + 4));
      asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(tile_end) : "r"(remote + 8));
    } else {
      first_tile = sched[0];
      tile_step = sched[1];
      tile_end = sched[2];
    }
  }

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const bool wave_sync = p.wave_sync != 0 && p.die_of_sm != nullptr && p.die_mode == 1 && cta_rank == 0;
      const int my_die = wave_sync ? sched[3] : 0;
      bool sync_ok = true;
      unsigned wave = 0;
      // the descriptor of the NEXT tile is fetched while this one streams (a dependent global load at every tile
      // start would stall the ring for ~1 us: 12 % of a tile at K = 2 000 sequences)
      uint4 tile = first_tile < tile_end ? p.tiles[(size_t)first_tile * (size_t)p.tile_mul] : make_uint4(0, 0, 0, 0);
      for (int t = first_tile; t < tile_end; t += tile_step, ++wave) {
        const uint4 tile_next = t + tile_step < tile_end ? p.tiles[(size_t)(t + tile_step) * (size_t)p.tile_mul] : tile;
        if (wave_sync && sync_ok) {
          // tiles of all earlier waves of this die have had their loads issued (bounded: if a pair of the die is
          // missing, synchronisation is abandoned rather than the GPU hung)
          const unsigned target = wave * (unsigned)tile_step;
          const long long t0 = clock64();
          while (*(volatile unsigned int*)&p.wave_counter[my_die] < target) {
            if (clock64() - t0 > 400000000ll) { sync_ok = false; break; }
            __nanosleep(32);
          }
        }
        const int m_row = (int)tile.x * (kBlockM * kCtas) + (int)cta_rank * kBlockM;
        const int n_row = (int)tile.y * kBlockN + (int)cta_rank * Cfg::kBRows;
        // (Tiles that share a panel run in lockstep through K: one fetches a line, the others hit in L2.  Starting
        // concurrent tiles at different K blocks was tried and raises DRAM traffic, profiles/r01_l2_sweep_*.)
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.error_flag, 1);
          if constexpr (kCtas == 2) {
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::kBytes);
            tma_load_2d_cg2(sA + stage * kAStageBytes, &tmA, kb * p.k_mul * kBlockK, m_row, full_bar(stage), p.hint_a);
            tma_load_2d_cg2(sB + stage * Cfg::kBBytes, &tmB, kb * p.k_mul * kBlockK, n_row, full_bar(stage), p.hint_b);
          } else {
            mbar_expect_tx(full_bar(stage), Cfg::kBytes);
            tma_load_2d(sA + stage * kAStageBytes, &tmA, kb * p.k_mul * kBlockK, m_row, full_bar(stage), p.hint_a);
            tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, kb * p.k_mul * kBlockK, n_row, full_bar(stage), p.hint_b);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (wave_sync) atomicAdd(&p.wave_counter[my_die], 1u);
        tile = tile_next;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tcount = 0;
      for (int t = first_tile; t < tile_end; t += tile_step, ++tcount) {
        const uint32_t buf = tcount & 1u, bphase = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(buf), bphase ^ 1u, p.error_flag, 2);  // epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kBlockN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase, p.error_flag, 3);  // TMA bytes (of both CTAs) landed
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(sA + stage * kAStageBytes);
          const uint64_t bdesc = make_smem_desc(sB + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockKBytes / kUmmaKBytes; ++k) {
            // advance 32 bytes along K inside the swizzle atom: +2 in 16-byte units
            umma<kI8, kCtas>(d_tmem, adesc + 2u * k, bdesc + 2u * k, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs) when these MMAs retire
          if constexpr (kCtas == 2) umma_commit_cg2(empty_bar(stage));
          else umma_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if constexpr (kCtas == 2) umma_commit_cg2(tfull_bar(buf));
        else umma_commit(tfull_bar(buf));
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    constexpr int kParts = kNumEpiWarps / 8;        // epilogue warps per (lane quarter, column group)
    const int half = (warp - kEpiWarp0) / (4 * kParts);       // which 128-column group of the accumulator
    const int part = ((warp - kEpiWarp0) >> 2) % kParts;      // which share of the group's site-pair steps
    constexpr int kSteps = (SPG + 1) / 2;           // two sites j per step
    const int jp_begin = (kSteps * part + kParts - 1) / kParts, jp_end = (kSteps * (part + 1) + kParts - 1) / kParts;
    const int alpha = lane & 1;
    float scale_f[NL];  // limb weights 2^(b*(NL-1-l)), pre-scaled by 2^-sum_shift, for the fp32 pre-filter
#pragma unroll
    for (int l = 0; l < NL; ++l)
      scale_f[l] = __int_as_float((127 + p.limb_bits * (NL - 1 - l) - p.sum_shift) << 23);
    float kappa = 0.0f;
    if constexpr (kScreen) kappa = __ldg(p.kappa);
    CandidateQueue queue;
    queue.init(smem + Cfg::kQueueOffset + (warp - kEpiWarp0) * kQueueBytesPerWarp);
    unsigned long long done = 0;
    uint32_t tcount = 0;
    uint4 tile_cur = first_tile < tile_end ? p.tiles[(size_t)first_tile * (size_t)p.tile_mul] : make_uint4(0, 0, 0, 0);
    for (int t = first_tile; t < tile_end; t += tile_step, ++tcount) {
      const uint4 tile = tile_cur;
      if (t + tile_step < tile_end) tile_cur = p.tiles[(size_t)(t + tile_step) * (size_t)p.tile_mul];  // needed one tile later
      const uint32_t buf = tcount & 1u, bphase = (tcount >> 1) & 1u;
      const int i_min = ((int)tile.x * kCtas + (int)cta_rank) * (kBlockM / 2) + quarter * 16;
      const int site_i = i_min + (lane >> 1);
      const int site_j0 = (int)tile.y * (2 * SPG) + half * SPG;
      mbar_wait(tfull_bar(buf), bphase, p.error_flag, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kBlockN + half * 128;
      // whole tile below the diagonal band, or outside the tile's window, for this warp?  -> nothing to do
      const int j_hi = min(site_j0 + min(2 * jp_end, SPG), (int)tile.w), j_lo = max(site_j0 + 2 * jp_begin, (int)tile.z);
      const bool any_work = i_min < j_hi && j_lo < j_hi && p.debug_skip == 0;
      if constexpr (kScreen) {
        // ---- one-limb screen: RPS = 2, SPG = 64; four sites j (8 columns) per step.  x = exact s32 sums of
        // gain x TOP limb; the true sums lie in [x, x (1 + 1/top_min)] (pair_prep.cu), so
        //   r2 <= (|P - Q| + kappa max(P, Q))^2 / (A a B b),  P = AB ab, Q = Ab aB, marginals from x
        // (ld_screen_f32).  The test is symmetric in the two rows of site i, so a lane uses (own row, peer row)
        // as (AB, Ab | aB, ab) whichever of the two it holds.
        static_assert(!kScreen || (NL == 1 && kI8), "the screen is the one-limb u8 kernel");
        if (any_work) {
          // warp-uniform fast path: the warp's 16 sites i all precede its 64 sites j and the window covers them
          const bool all_valid = i_min + 15 < site_j0 && site_j0 >= (int)tile.z && site_j0 + SPG <= (int)tile.w;
          uint32_t n_valid = 0;
          bool hit = false;  // warp-uniform: this warp found a candidate in this tile
          auto sweep = [&](auto all_valid_c) {
            constexpr bool kAllValid = decltype(all_valid_c)::value;
            uint32_t v[8];
            tmem_ld8(taddr + part * (SPG / 4 / kParts) * 8, v);
            constexpr int kSweep = SPG / 4 / kParts;  // steps of this warp: [js0, js0 + kSweep)
            const int js0 = part * kSweep;
#pragma unroll 1
            for (int js = js0; js < js0 + kSweep; ++js) {
              tmem_ld_wait();
              float f[8];
#pragma unroll
              for (int x = 0; x < 8; ++x) f[x] = (float)(int)v[x];
              if (js + 1 < js0 + kSweep) tmem_ld8(taddr + (js + 1) * 8, v);  // the next four sites travel while these are tested
              bool cand[2];
              int site_j[2];
#pragma unroll
              for (int pp = 0; pp < 2; ++pp) {
                // lane alpha finishes site 4js + 2pp + alpha and sends its row's sums of the other site of the pair
                const float own0 = alpha ? f[4 * pp + 2] : f[4 * pp + 0], own1 = alpha ? f[4 * pp + 3] : f[4 * pp + 1];
                const float snd0 = alpha ? f[4 * pp + 0] : f[4 * pp + 2], snd1 = alpha ? f[4 * pp + 1] : f[4 * pp + 3];
                const float rcv0 = __shfl_xor_sync(0xffffffffu, snd0, 1), rcv1 = __shfl_xor_sync(0xffffffffu, snd1, 1);
                site_j[pp] = site_j0 + 4 * js + 2 * pp + alpha;
                cand[pp] = ld_screen_f32(own0, own1, rcv0, rcv1, kappa, p.thr_lo_f);
                if constexpr (!kAllValid) {
                  const bool valid = site_i < site_j[pp] && site_j[pp] >= (int)tile.z && site_j[pp] < (int)tile.w;  // lib.rs:651
                  n_valid += valid;
                  cand[pp] = cand[pp] && valid;
                }
              }
              if (__any_sync(0xffffffffu, cand[0] || cand[1])) {  // rare
                hit = true;
                emit_cand_warp(cand[0], (uint32_t)site_i, (uint32_t)site_j[0], p.cand);
                emit_cand_warp(cand[1], (uint32_t)site_i, (uint32_t)site_j[1], p.cand);
              }
            }
            if constexpr (kAllValid) n_valid = SPG / 2 / kParts;  // every lane finished 2 pairs per step
          };
          if (all_valid) sweep(std::true_type{});
          else sweep(std::false_type{});
          done += n_valid;
          if (hit && lane == 0) {
            if (p.cell_flags != nullptr) p.cell_flags[(size_t)t * (size_t)p.tile_mul] = 1;
            if (p.sample_flags != nullptr && atomicExch(&p.sample_flags[t], 1u) == 0u) atomicAdd(p.sample_flagged, 1ull);
          }
        }
      } else if (any_work) {
#pragma unroll 1
        for (int jp = jp_begin; jp < jp_end; ++jp) {
          uint32_t v[2 * RPS];
          const bool has2 = (2 * jp + 1) < SPG;
          if (has2) {
            tmem_ld_cols<2 * RPS>(taddr + jp * 2 * RPS, v);
          } else {
            tmem_ld_cols<RPS>(taddr + jp * 2 * RPS, v);
#pragma unroll
            for (int x = RPS; x < 2 * RPS; ++x) v[x] = 0u;
          }
          tmem_ld_wait();
          // ---- fp32 pass: limb recombination (rounded), half-table swap, conservative pre-filter
          float F[2][2];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int beta = 0; beta < 2; ++beta) {
              float s = 0.0f;
#pragma unroll
              for (int l = 0; l < NL; ++l) {
                const uint32_t raw = v[jj * RPS + beta * NL + l];  // exact integer: fp32 or s32 accumulator
                s = fmaf(kI8 ? (float)(int)raw : __uint_as_float(raw), scale_f[l], s);
              }
              F[jj][beta] = s;
            }
          // lane alpha=0 finishes site j0 = 2jp, lane alpha=1 finishes j1 = 2jp+1; swap the other halves
          const float fr0 = __shfl_xor_sync(0xffffffffu, alpha ? F[0][0] : F[1][0], 1);
          const float fr1 = __shfl_xor_sync(0xffffffffu, alpha ? F[0][1] : F[1][1], 1);
          const float fo0 = alpha ? F[1][0] : F[0][0];
          const float fo1 = alpha ? F[1][1] : F[0][1];
          const int j_local = 2 * jp + alpha;
          const int site_j = site_j0 + j_local;
          const bool valid = j_local < SPG && site_i < site_j && site_j >= (int)tile.z && site_j < (int)tile.w;  // lib.rs:651
          done += valid;
          bool keep = valid && ld_prefilter_f32(alpha ? fr0 : fo0, alpha ? fr1 : fo1, alpha ? fo0 : fr0,
                                                alpha ? fo1 : fr1, p.thr_lo_f, p.thr_negative != 0);
          if (__any_sync(0xffffffffu, keep)) {
            // ---- candidates: exact integer sums (limbs recombined in int64), swap halves, enqueue
            long long S[2][2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
              for (int beta = 0; beta < 2; ++beta) {
                long long acc = 0;
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                  const uint32_t raw = v[jj * RPS + beta * NL + l];
                  const long long limb = kI8 ? (long long)(int)raw : (long long)__float2int_rn(__uint_as_float(raw));
                  acc += limb << (p.limb_bits * (NL - 1 - l));
                }
                S[jj][beta] = acc;
              }
            const long long recv0 = __shfl_xor_sync(0xffffffffu, alpha ? S[0][0] : S[1][0], 1);
            const long long recv1 = __shfl_xor_sync(0xffffffffu, alpha ? S[0][1] : S[1][1], 1);
            const long long own0 = alpha ? S[1][0] : S[0][0];
            const long long own1 = alpha ? S[1][1] : S[0][1];
            queue.push(keep, alpha ? recv0 : own0, alpha ? recv1 : own1, alpha ? own0 : recv0, alpha ? own1 : recv1,
                       (uint32_t)site_i, (uint32_t)site_j);
            if (queue.count > kQueueCap - 32) queue.drain32(p.thr, p.out, p.py_aux);  // f64 statistics, lib.rs:482-518
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtas == 2) mbar_arrive_cluster(tempty_bar(buf) & kPeerBitMask);  // the leader's barrier
        else mbar_arrive(tempty_bar(buf));
      }
      if ((tcount & 31u) == 31u) {  // publish progress every 32 tiles (polled by wld_ld_pairs for the callback, lib.rs:670-674)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
        if (lane == 0 && done) atomicAdd(p.pairs_done, done);
        done = 0;
      }
    }
    while (queue.count > 0) queue.drain32(p.thr, p.out, p.py_aux);  // tail
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
    if (lane == 0 && done) atomicAdd(p.pairs_done, done);
  }

  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all();  // nobody leaves while the peer may still touch its smem / TMEM
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kCtas == 2) tmem_dealloc_cg2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

bool make_tensor_map(CUtensorMap* map, void* base, uint64_t rows, uint64_t kp, uint32_t box_rows, int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {kp, rows};
  const cuuint64_t gstride[1] = {kp * (uint64_t)elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)(kBlockKBytes / elem_bytes), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;  // none / 128 B / 256 B measured equal
  return fn(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim,
            gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, promo,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NL, bool kI8, int kCtas, bool kScreen>
cudaError_t launch(int grid, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB,
                   const UmmaParams& prm) {
  auto kern = pair_umma_kernel<NL, kI8, kCtas, kScreen>;
  constexpr int smem = StageCfg<kCtas>::kSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, prm);
}

template <int kCtas>
cudaError_t launch_any(int n_limbs, bool i8, bool screen, int grid, cudaStream_t stream, const CUtensorMap& tmA,
                       const CUtensorMap& tmB, const UmmaParams& prm) {
  if (screen) return launch<1, true, kCtas, true>(grid, stream, tmA, tmB, prm);
  switch (n_limbs * 2 + (i8 ? 1 : 0)) {
    case 2: return launch<1, false, kCtas, false>(grid, stream, tmA, tmB, prm);
    case 3: return launch<1, true, kCtas, false>(grid, stream, tmA, tmB, prm);
    case 4: return launch<2, false, kCtas, false>(grid, stream, tmA, tmB, prm);
    case 5: return launch<2, true, kCtas, false>(grid, stream, tmA, tmB, prm);
    case 6: return launch<3, false, kCtas, false>(grid, stream, tmA, tmB, prm);
    case 7: return launch<3, true, kCtas, false>(grid, stream, tmA, tmB, prm);
    case 8: return launch<4, false, kCtas, false>(grid, stream, tmA, tmB, prm);
    case 9: return launch<4, true, kCtas, false>(grid, stream, tmA, tmB, prm);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

// The schedule (replaces rayon's fan-out over triu_index, lib.rs:623-637).
//
// PARTITIONS are defined on a canonical grid of 128 x 128-site cells, independent of the kernel variant: the
// cells that hold a pair a < b, rasterised in strips of 8 cell columns, are cut into `nparts` contiguous ranges of
// equal cell count (= equal tensor work; every strip holds its own stretch of the diagonal, so the epilogue work
// is balanced too).  A partition therefore owns, in every cell row, ONE contiguous interval of site columns,
// whatever the tile shape — each GPU may pick the one-limb screen or the exact n-limb kernel on its own and the
// union over the GPUs still covers every pair exactly once.
//
// TILES of a kernel variant (M = 64 * ctas sites, N = 2 * floor(128 / (2 n_limbs)) sites) are listed in strips of
// 8 N tiles so that the tiles in flight share operand panels through L2; a tile carries the window [j_lo, j_hi)
// of site columns that belong to this partition (the whole tile except where it straddles a partition boundary).
TilePlan plan_tiles(int64_t L, int n_limbs, int part, int nparts, int sm_count, int ctas) {
  (void)sm_count;
  TilePlan plan;
  plan.tile_m = (kBlockM / 2) * ctas;
  plan.tile_n = 2 * (128 / (2 * n_limbs));
  constexpr int64_t kCell = 128, kStrip = 8;
  const int64_t n_c = (L + kCell - 1) / kCell;
  // column interval (in cells) of every cell row that this partition owns
  std::vector<int64_t> c_lo((size_t)std::max<int64_t>(n_c, 1), INT64_MAX), c_hi((size_t)std::max<int64_t>(n_c, 1), -1);
  {
    auto cell_has_pair = [&](int64_t ci, int64_t cj) { return ci * kCell < std::min(L, (cj + 1) * kCell) - 1; };
    uint64_t total = 0;
    for (int64_t cs = 0; cs < n_c; cs += kStrip)
      for (int64_t ci = 0; ci < std::min(n_c, cs + kStrip); ++ci)
        total += (uint64_t)std::max<int64_t>(0, std::min(cs + kStrip, n_c) - std::max(cs, ci)) -
                 ((ci >= cs && ci < std::min(cs + kStrip, n_c) && !cell_has_pair(ci, ci)) ? 1 : 0);
    const uint64_t lo = total * (uint64_t)part / (uint64_t)nparts, hi = total * (uint64_t)(part + 1) / (uint64_t)nparts;
    uint64_t idx = 0;
    for (int64_t cs = 0; cs < n_c && idx < hi; cs += kStrip) {
      const int64_t ce = std::min(cs + kStrip, n_c);
      for (int64_t ci = 0; ci < ce && idx < hi; ++ci)
        for (int64_t cj = std::max(cs, ci); cj < ce && idx < hi; ++cj) {
          if (!cell_has_pair(ci, cj)) continue;
          if (idx >= lo) {
            c_lo[(size_t)ci] = std::min(c_lo[(size_t)ci], cj);
            c_hi[(size_t)ci] = std::max(c_hi[(size_t)ci], cj + 1);
          }
          ++idx;
        }
    }
  }
  const int64_t tile_m = plan.tile_m, tile_n = plan.tile_n;
  const int64_t n_mt = (L + tile_m - 1) / tile_m, n_nt = (L + tile_n - 1) / tile_n;
  int64_t strip = kStrip;  // in N tiles; WLD_STRIP overrides it for experiments
  if (const char* e = std::getenv("WLD_STRIP")) strip = std::max(1, std::atoi(e));
  for (int64_t ns = 0; ns < n_nt; ns += strip) {
    const int64_t ne = std::min(ns + strip, n_nt);
    for (int64_t mi = 0; mi < n_mt; ++mi) {
      const size_t ci = (size_t)(mi * tile_m / kCell);
      if (c_hi[ci] < 0) continue;
      const int64_t i0 = mi * tile_m, i1 = std::min(L, i0 + tile_m);
      const int64_t own_lo = c_lo[ci] * kCell, own_hi = std::min(L, c_hi[ci] * kCell);
      for (int64_t nj = ns; nj < ne; ++nj) {
        const int64_t w_lo = std::max(nj * tile_n, own_lo), w_hi = std::min((nj + 1) * tile_n, own_hi);
        if (w_lo >= w_hi || i0 >= w_hi - 1) continue;  // outside the partition / no pair a < b in the window
        plan.tiles.push_back(make_uint4((unsigned)mi, (unsigned)nj, (unsigned)w_lo, (unsigned)w_hi));
        if (i1 <= w_lo) {  // entirely above the diagonal: every (i, j) of the window is a pair
          plan.pairs += (uint64_t)((i1 - i0) * (w_hi - w_lo));
        } else {
          for (int64_t i = i0; i < i1; ++i) plan.pairs += (uint64_t)std::max<int64_t>(0, w_hi - std::max(w_lo, i + 1));
        }
      }
    }
  }
  return plan;
}

// A schedule only depends on (n_kept, limbs, partition, cta_group): it is planned once, kept on the device
// between calls, and its extent (which M / N tiles this partition touches) tells pair_prep which operand rows
// to expand.  which = 0: the exact kernel (c->geom.n_limbs limbs); 1: the one-limb screen.
int ensure_tile_plan(wld_ctx* c, int which) {
  DevPlan& dp = c->plans[which];
  const int64_t L = c->n_kept;
  const int ctas = c->cta_group;
  const int n_limbs = which == 1 ? 1 : c->geom.n_limbs;
  const int64_t key[5] = {L, n_limbs, c->part, c->nparts, ctas};
  if (std::memcmp(key, dp.key, sizeof key) == 0) {
    c->plan_pairs = dp.pairs;
    return WLD_OK;
  }
  TilePlan plan = plan_tiles(L, n_limbs, c->part, c->nparts, c->sm_count, ctas);
  WLD_CUDA(c, dp.tiles.ensure(sizeof(uint4) * std::max<size_t>(plan.tiles.size(), 1)));
  if (!plan.tiles.empty())
    WLD_CUDA(c, cudaMemcpyAsync(dp.tiles.p, plan.tiles.data(), sizeof(uint4) * plan.tiles.size(),
                                cudaMemcpyHostToDevice, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vector dies at the end of this function
  std::memcpy(dp.key, key, sizeof key);
  dp.n_tiles = (int64_t)plan.tiles.size();
  dp.pairs = plan.pairs;
  dp.tile_m = plan.tile_m;
  dp.tile_n = plan.tile_n;
  dp.x[0] = dp.y[0] = INT64_MAX;
  dp.x[1] = dp.y[1] = -1;
  for (const uint4& t : plan.tiles) {
    dp.x[0] = std::min<int64_t>(dp.x[0], t.x);
    dp.x[1] = std::max<int64_t>(dp.x[1], t.x);
    dp.y[0] = std::min<int64_t>(dp.y[0], t.y);
    dp.y[1] = std::max<int64_t>(dp.y[1], t.y);
  }
  dp.host.swap(plan.tiles);
  c->plan_pairs = dp.pairs;
  return WLD_OK;
}

// The exact kernel restricted to the cells in which the screen found candidates: every flagged screen tile
// (a 128-site cell column window of one M tile) is covered by the exact kernel's N tiles with their windows clipped
// to it.  Cells without a flag hold no pair that can pass the threshold (the screen's bound), so nothing is lost.
// The list keeps the strip-major order of the screen's list.
TilePlan cut_cell_plan(const std::vector<uint4>& screen_tiles, const uint8_t* flags, size_t n_flags, int64_t n_kept,
                       int n_limbs, int ctas, int64_t* n_flagged) {
  TilePlan plan;
  plan.tile_m = (kBlockM / 2) * ctas;
  plan.tile_n = 2 * (128 / (2 * n_limbs));
  const int64_t tile_m = plan.tile_m, tile_n = plan.tile_n;
  int64_t flagged = 0;
  for (size_t k = 0; k < screen_tiles.size() && k < n_flags; ++k) {
    if (!flags[k]) continue;
    ++flagged;
    const uint4 cell = screen_tiles[k];  // {M tile, cell column, j_lo, j_hi}
    const int64_t i0 = (int64_t)cell.x * tile_m, i1 = std::min<int64_t>(n_kept, i0 + tile_m);
    for (int64_t nj = (int64_t)cell.z / tile_n; nj * tile_n < (int64_t)cell.w; ++nj) {
      const int64_t w_lo = std::max<int64_t>(nj * tile_n, cell.z), w_hi = std::min<int64_t>((nj + 1) * tile_n, cell.w);
      if (w_lo >= w_hi || i0 >= w_hi - 1) continue;
      plan.tiles.push_back(make_uint4(cell.x, (unsigned)nj, (unsigned)w_lo, (unsigned)w_hi));
      if (i1 <= w_lo) plan.pairs += (uint64_t)((i1 - i0) * (w_hi - w_lo));
      else
        for (int64_t i = i0; i < i1; ++i) plan.pairs += (uint64_t)std::max<int64_t>(0, w_hi - std::max(w_lo, i + 1));
    }
  }
  if (n_flagged) *n_flagged = flagged;
  return plan;
}

int build_cell_plan(wld_ctx* c, const std::vector<uint8_t>& flags, int64_t* n_flagged) {
  DevPlan& dp = c->plans[2];
  TilePlan cut = cut_cell_plan(c->plans[1].host, flags.data(), flags.size(), c->n_kept, c->geom.n_limbs, c->cta_group, n_flagged);
  std::vector<uint4>& tiles = cut.tiles;
  const uint64_t pairs = cut.pairs;
  const int64_t tile_m = cut.tile_m, tile_n = cut.tile_n;
  WLD_CUDA(c, dp.tiles.ensure(sizeof(uint4) * std::max<size_t>(tiles.size(), 1)));
  if (!tiles.empty())
    WLD_CUDA(c, cudaMemcpyAsync(dp.tiles.p, tiles.data(), sizeof(uint4) * tiles.size(), cudaMemcpyHostToDevice, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  dp.key[0] = -1;  // never reused
  dp.n_tiles = (int64_t)tiles.size();
  dp.pairs = pairs;
  dp.tile_m = tile_m;
  dp.tile_n = tile_n;
  dp.host.swap(tiles);
  return WLD_OK;
}

// mode 0: exact kernel; 1: screen, every tile, candidates into c->cand; 2: screen over a sample of the tiles,
// counting only (candidates -> counters[8], pairs -> counters[9]).
int run_pair_umma(wld_ctx* c, float thr, int mode) {
  const PairGeom& gm = c->geom;
  const int ctas = c->cta_group;
  const bool screen = mode == 1 || mode == 2;
  if (mode != 3) {
    const int rc = ensure_tile_plan(c, screen ? 1 : 0);
    if (rc != WLD_OK) return rc;
  }
  const DevPlan& dp = c->plans[mode == 3 ? 2 : screen ? 1 : 0];
  const int n_limbs = screen ? 1 : gm.n_limbs;
  const int spg = 128 / (2 * n_limbs);
  int64_t n_tiles = dp.n_tiles;
  int tile_mul = 1;
  const int64_t pairs_of_ctas = c->sm_count / ctas;
  if (mode == 2) {  // two waves of tiles (one when the schedule is short: multi-GPU shares), evenly spread over the list
    const int64_t waves = n_tiles >= 64 * pairs_of_ctas ? 2 : 1;
    tile_mul = (int)std::max<int64_t>(1, n_tiles / (waves * pairs_of_ctas));
    n_tiles = (n_tiles + tile_mul - 1) / tile_mul;
  }
  if (mode != 2) {
    c->info.tiles = n_tiles;
    c->info.tile_sites_m = (kBlockM / 2) * ctas;
    c->info.tile_sites_n = 2 * spg;
    c->info.executed_flop = (double)n_tiles * 2.0 * (kBlockM * ctas) * kBlockN * (double)gm.k_padded;
  }
  if (n_tiles == 0) return WLD_OK;

  const int64_t b_groups = screen ? round_up(std::max<int64_t>((c->n_kept + 63) / 64, 1), 2) : gm.b_groups;
  CUtensorMap tmA, tmB;
  if (!make_tensor_map(&tmA, c->opA.p, (uint64_t)gm.a_rows, (uint64_t)gm.k_padded, kBlockM, gm.elem_bytes) ||
      !make_tensor_map(&tmB, screen ? c->opB1.p : c->opB.p, (uint64_t)b_groups * 128, (uint64_t)gm.k_padded, kBlockN / ctas,
                       gm.elem_bytes))
    return c->fail(WLD_ERR_CUDA, "cuTensorMapEncodeTiled failed (driver without TMA support?)");

  unsigned long long* cnt = c->counters.as<unsigned long long>();
  UmmaParams prm;
  prm.tiles = dp.tiles.as<uint4>();
  prm.n_tiles = (int)n_tiles;
  prm.tile_mul = tile_mul;
  prm.k_blocks = (int)(gm.k_padded * gm.elem_bytes / kBlockKBytes);
  prm.k_mul = 1;
  {
    const char* e = std::getenv("WLD_EXPERIMENT_SKIP_EPILOGUE");  // timing experiments: results are NOT computed
    prm.debug_skip = (e && e[0] == '1') ? 1 : 0;
  }
  if (mode == 2 && prm.k_blocks > 64) {
    // The sample only has to tell "a few candidates in 10^5 pairs" from "a few in 10": every k_mul-th block of 128
    // sequences (about 6 000 of them, spread over the whole alignment) estimates that as well as all of them.
    prm.k_mul = (prm.k_blocks + 47) / 48;
    prm.k_blocks = (prm.k_blocks + prm.k_mul - 1) / prm.k_mul;
  }
  prm.n_kept = (int)c->n_kept;
  prm.limb_bits = gm.limb_bits;
  prm.thr = thr;
  prm.thr_lo = ld_thr_lo(thr);
  prm.py_aux = c->compat == WLD_COMPAT_PYTHON ? c->py_aux.as<uint2>() : nullptr;
  {
    // Within a strip the limb panels (B) are reused by every M tile while the indicator panels (A) stream
    // past once: keep B, let A go first.  WLD_HINT_A / WLD_HINT_B = normal|first|last override (experiments).
    auto policy = [](const char* env, uint64_t dflt) {
      const char* e = std::getenv(env);
      if (!e) return dflt;
      return e[0] == 'f' ? kL2EvictFirst : e[0] == 'l' ? kL2EvictLast : kL2EvictNormal;
    };
    prm.hint_a = policy("WLD_HINT_A", kL2EvictNormal);
    prm.hint_b = policy("WLD_HINT_B", kL2EvictNormal);
  }
  {
    // fp32 pre-filter threshold: lowered once more by 1e-5 relative, rounded towards -inf
    const double lo = prm.thr_lo;
    prm.thr_negative = lo < 0.0 ? 1 : 0;
    float f = (float)(lo * (1.0 - 1e-5));
    if ((double)f > lo * (1.0 - 1e-5)) f = nextafterf(f, -INFINITY);
    prm.thr_lo_f = f > 0.f ? f : 0.f;
    // scale sums so that T <= sum of all weights <= 2^21
    int shift = 0;
    while (std::ldexp(c->weight_sum, -shift) > 2097152.0) ++shift;
    prm.sum_shift = shift;
  }
  prm.out = PairOut{c->pairs.as<wld_pair>(), cnt, c->pair_cap};
  prm.pairs_done = cnt + 1;
  prm.error_flag = reinterpret_cast<int*>(cnt + 2);
  prm.cand = CandOut{nullptr, cnt + 5, 0};
  prm.cell_flags = nullptr;
  prm.sample_flags = nullptr;
  prm.sample_flagged = cnt + 10;
  prm.kappa = &c->quant.as<QuantDecision>()->kappa;
  if (screen) {
    prm.thr_lo_f = std::min(prm.thr_lo_f, 2.0f);  // (a threshold above 1 can never pass; keeps thr * den finite)
    if (mode == 1) {
      prm.cell_flags = c->cell_flags.as<uint8_t>();
      prm.cand = CandOut{c->cand.as<uint2>(), cnt + 5, c->cand_cap};
    } else {
      prm.cand = CandOut{nullptr, cnt + 8, 0};
      prm.pairs_done = cnt + 9;
      WLD_CUDA(c, c->sample_flags.ensure(sizeof(unsigned int) * (size_t)n_tiles));
      WLD_CUDA(c, cudaMemsetAsync(c->sample_flags.p, 0, sizeof(unsigned int) * (size_t)n_tiles, c->stream));
      prm.sample_flags = c->sample_flags.as<unsigned int>();
      c->sample_tiles = n_tiles;
    }
  }

  int grid = (int)std::min<int64_t>(n_tiles, pairs_of_ctas) * ctas;
  // Die-aware schedule: each L2 die works on its own contiguous part of the (strip-rasterised) tile list, so
  // the panels a die's L2 holds are only the ones its own SMs reuse.  WLD_DIE=0 disables (experiments).
  prm.die_of_sm = nullptr;
  prm.n_sm = 0;
  prm.die_counter = reinterpret_cast<unsigned int*>(cnt + 3);
  prm.die_pairs[0] = prm.die_pairs[1] = 0;
  prm.die_split = 0;
  prm.die_mode = 1;
  prm.wave_sync = 0;
  prm.wave_counter = reinterpret_cast<unsigned int*>(cnt + 4);
  if (mode != 2) c->die_used = 0;
  {
    const char* e = std::getenv("WLD_DIE");
    if (mode != 2 && c->die_aware && !(e && e[0] == '0') && n_tiles >= 4 * pairs_of_ctas) {
      const std::vector<uint8_t>& dm = die_map(c);
      if ((int)dm.size() == c->sm_count) {
        int sms[2] = {0, 0};
        for (uint8_t d : dm) ++sms[d ? 1 : 0];
        const int n0 = sms[0] / ctas, n1 = sms[1] / ctas;
        if (n0 > 0 && n1 > 0) {
          if (c->die_of_sm.bytes < dm.size() || !c->die_of_sm.p) {
            WLD_CUDA(c, c->die_of_sm.ensure(dm.size()));
            WLD_CUDA(c, cudaMemcpyAsync(c->die_of_sm.p, dm.data(), dm.size(), cudaMemcpyHostToDevice, c->stream));
            WLD_CUDA(c, cudaStreamSynchronize(c->stream));
          }
          prm.die_of_sm = c->die_of_sm.as<uint8_t>();
          prm.n_sm = (unsigned)dm.size();
          prm.die_pairs[0] = n0;
          prm.die_pairs[1] = n1;
          prm.die_split = (int)((n_tiles * n0 + (n0 + n1) / 2) / (n0 + n1));
          prm.die_mode = (e && e[0] == '2') ? 2 : 1;
          grid = (n0 + n1) * ctas;
          c->die_used = prm.die_mode;
          {  // A die-wave of tiles streams (pairs/8 + 8) panels of 256 rows.  Measured on one box each (pair kernel,
            // same run): config 4 (307 MB per wave) 33.1 -> 29.8 ms and DRAM reads 120 -> 81 GB, config 5 (31 MB)
            // 88.3 -> 79.7 ms; config 3 (6 MB) 2.56 -> 2.61 ms — so waves are synchronised from 16 MB up.
            const double wave_bytes = (double)(std::max(n0, n1) / 8 + 8) * 256.0 * (double)gm.k_padded * gm.elem_bytes;
            const char* ws = std::getenv("WLD_WAVESYNC");
            prm.wave_sync = ws ? (ws[0] != '0') : (wave_bytes > 16e6);
          }
          c->info.die_sms[0] = sms[0];
          c->info.die_sms[1] = sms[1];
        }
      }
    }
  }
  const bool i8 = gm.elem_bytes == 1;
  if (n_limbs < 1 || n_limbs > 4) return c->fail(WLD_ERR_INVALID, "n_limbs must be 1..4");
  if (screen && !i8) return c->fail(WLD_ERR_INVALID, "the screen needs the u8 operands");
  ScopedStageTimer tm(c, mode == 2 ? WLD_STAGE_PAIR_SAMPLE : WLD_STAGE_PAIR);  // kernel only
  const cudaError_t e = ctas == 2 ? launch_any<2>(n_limbs, i8, screen, grid, c->stream, tmA, tmB, prm)
                                  : launch_any<1>(n_limbs, i8, screen, grid, c->stream, tmA, tmB, prm);
  tm.launched();
  if (e != cudaSuccess) return c->fail(WLD_ERR_CUDA, "pair_umma launch failed: %s", cudaGetErrorString(e));
  return WLD_OK;
}

}  // namespace wld

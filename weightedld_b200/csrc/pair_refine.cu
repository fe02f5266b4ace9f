// pair_refine.cu — stage 3c: exact statistics of the site pairs the one-limb screen could not rule out.
//
// The screen (pair_umma.cu, kScreen) runs the Gram with the top limb of every fixed-point weight and bounds r2
// from above; what it leaves is a short list of candidate pairs (kept-site indices).  Each candidate is
// recomputed here from first principles — the 0..5 code rows of its two sites and the FULL integer weights
// q[s] (pair_prep.cu) — one warp per pair:
//     AB = sum q [a = maj_a][b = maj_b] ... ab = sum q [a = min_a][b = min_b]        (lib.rs:462-479)
// in 64-bit integers (exact, any order), then D, D', r2 and the filter `r2 > thr` by the shared f64 epilogue
// (lib.rs:482-518, 660).  These are the sums the exact n-limb tensor kernel accumulates, so the surviving
// records are identical bit for bit; only the order in the survivor buffer differs, and pair_order.cu sorts it.
//
// Cost per candidate: its two code rows (2 * ldc bytes from L2; candidates arrive grouped by tile) and ~10 integer
// instructions per sequence.  The sums are taken limb by limb with the 16 x 8-bit dot product (dp2a): the
// quantiser leaves v_l[s] = gain[s] * limb_l[s] (< 2^15) per sequence and limb, a cell's indicator is a byte mask
// (0x80 where the sequence belongs to the cell), and acc[cell][l] += v_l[s] * mask[s] runs two sequences per
// instruction in u32, flushed into u64 every 2048 sequences.  The v_l (2 NL bytes per sequence, the larger stream)
// are staged in shared memory once per group of 16 candidates.  The f64 statistics run once per candidate (one
// lane), B200's FP64 pipe being narrow.
#include "common.cuh"
#include "pair_epilogue.cuh"

namespace wld {
namespace {

constexpr int kRefineWarps = 8;
constexpr int kPerWarp = 2;                              // candidates a warp carries through one pass
constexpr int kGroup = kRefineWarps * kPerWarp;          // candidates per block pass: they share the staged weights
constexpr int kChunk = 2048;                             // sequences staged in shared memory at a time
constexpr int kSteps = kChunk / 128;                     // a warp covers 128 sequences per step (4 per lane)

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// z = 8 * code_a + code_b per byte (<= 45); a cell is one value of z.  0x3f never occurs: the cell of a missing allele.
__device__ __forceinline__ uint32_t cell_key(int sym_a, int sym_b) {
  return (sym_a < 0 || sym_b < 0) ? 0x3f3f3f3fu : (uint32_t)(sym_a * 8 + sym_b) * 0x01010101u;
}
// 0x80 in every byte where z == key (both below 0x40, so x = z ^ key < 0x40 and x + 0x7f carries into bit 7 iff x != 0)
__device__ __forceinline__ uint32_t match80(uint32_t z, uint32_t key) { return ~((z ^ key) + 0x7f7f7f7fu) & 0x80808080u; }

template <int NL>
__global__ void __launch_bounds__(32 * kRefineWarps) pair_refine_kernel(
    const uint8_t* __restrict__ codes, int64_t ldc, const int8_t* __restrict__ maj, const int8_t* __restrict__ mnr,
    const uint16_t* __restrict__ glimb, int limb_bits, const uint2* __restrict__ cand,
    const unsigned long long* __restrict__ n_cand, unsigned long long cap, unsigned long long give_up,
    unsigned long long* __restrict__ gave_up, float thr, PairOut out) {
  __shared__ __align__(16) uint16_t s_v[NL][kChunk];  // gain x limb of the chunk's sequences
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (*n_cand > give_up) {
    // far more candidates than the sample promised (heterogeneous input): refining them one by one would cost more
    // than the exact tensor kernel over every pair — tell the host, which runs that instead
    if (blockIdx.x == 0 && threadIdx.x == 0) *gave_up = 1ull;
    return;
  }
  const unsigned long long n = min(*n_cand, cap);
  for (unsigned long long g0 = (unsigned long long)blockIdx.x * kGroup; g0 < n; g0 += (unsigned long long)gridDim.x * kGroup) {
    uint2 ij[kPerWarp];
    const uint8_t* ra[kPerWarp];
    const uint8_t* rb[kPerWarp];
    uint32_t key[kPerWarp][4];                // AB, Ab, aB, ab
    unsigned long long sum[kPerWarp][4];
    bool live[kPerWarp];
#pragma unroll
    for (int c = 0; c < kPerWarp; ++c) {
      const unsigned long long k = g0 + (unsigned long long)(warp * kPerWarp + c);
      live[c] = k < n;
      ij[c] = live[c] ? cand[k] : make_uint2(0u, 0u);
      ra[c] = codes + (int64_t)ij[c].x * ldc;
      rb[c] = codes + (int64_t)ij[c].y * ldc;
      const int aM = maj[ij[c].x], am = mnr[ij[c].x], bM = maj[ij[c].y], bm = mnr[ij[c].y];
      key[c][0] = cell_key(aM, bM); key[c][1] = cell_key(aM, bm);
      key[c][2] = cell_key(am, bM); key[c][3] = cell_key(am, bm);
#pragma unroll
      for (int t = 0; t < 4; ++t) sum[c][t] = 0;
    }
    for (int64_t c0 = 0; c0 < ldc; c0 += kChunk) {  // ldc is a multiple of 128; the padding holds code 5
      const int len = (int)min((int64_t)kChunk, ldc - c0);
      __syncthreads();  // the previous chunk has been consumed
#pragma unroll
      for (int l = 0; l < NL; ++l)
        for (int s = 8 * threadIdx.x; s < len; s += 8 * 32 * kRefineWarps)
          *reinterpret_cast<uint4*>(&s_v[l][s]) = __ldg(reinterpret_cast<const uint4*>(glimb + (int64_t)l * ldc + c0 + s));
      __syncthreads();
#pragma unroll
      for (int c = 0; c < kPerWarp; ++c) {
        if (!live[c]) continue;  // warp-uniform
        // all code words of the chunk first (32 loads in flight per warp), then the arithmetic
        uint32_t z[kSteps];
#pragma unroll
        for (int u = 0; u < kSteps; ++u) {
          const int s = 128 * u + 4 * lane;
          const bool in = 128 * u < len;  // warp-uniform: len is a multiple of 128
          const uint32_t wa = in ? __ldg(reinterpret_cast<const uint32_t*>(ra[c] + c0 + s)) : 0x05050505u;
          const uint32_t wb = in ? __ldg(reinterpret_cast<const uint32_t*>(rb[c] + c0 + s)) : 0x05050505u;
          z[u] = wa * 8u + wb;  // bytes stay below 0x40: no carries
        }
        uint32_t acc[4][NL];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int l = 0; l < NL; ++l) acc[t][l] = 0;
#pragma unroll
        for (int u = 0; u < kSteps; ++u) {
          if (128 * u >= len) break;
          const int s = 128 * u + 4 * lane;
          uint32_t m[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) m[t] = match80(z[u], key[c][t]);
#pragma unroll
          for (int l = 0; l < NL; ++l) {
            const uint2 v = *reinterpret_cast<const uint2*>(&s_v[l][s]);  // sequences s, s+1 | s+2, s+3
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t][l] = __dp2a_hi(v.y, m[t], __dp2a_lo(v.x, m[t], acc[t][l]));
          }
        }
        // <= 64 sequences x 2^15 x 0x80 per lane and chunk: below 2^32; recombine the limbs in 64 bits
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int l = 0; l < NL; ++l) sum[c][t] += (unsigned long long)(acc[t][l] >> 7) << (limb_bits * (NL - 1 - l));
      }
    }
    // lane c of the warp finishes candidate c (the f64 pipe is narrow: one evaluation per candidate, not 32)
    bool keep = false;
    float d = 0.f, dp = 0.f, r2 = 0.f;
    uint32_t mi = 0, mj = 0;
#pragma unroll
    for (int c = 0; c < kPerWarp; ++c) {
      unsigned long long t[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) t[x] = warp_sum_u64(sum[c][x]);
      if (lane == c && live[c]) {
        const double AB = (double)t[0], Ab = (double)t[1], aB = (double)t[2], ab = (double)t[3];  // < 2^53: exact
        const double A = AB + Ab, B = AB + aB, T = A + (aB + ab);
        // an empty marginal is NaN in the reference and dropped by lib.rs:660 (the screen never proposes one)
        keep = A > 0.0 && B > 0.0 && T - A > 0.0 && T - B > 0.0 && ld_stats_exact(AB, Ab, aB, ab, thr, d, dp, r2);
        mi = ij[c].x;
        mj = ij[c].y;
      }
    }
    emit_pairs_warp(keep, mi, mj, d, dp, r2, out);
  }
}

}  // namespace

int run_pair_refine(wld_ctx* c, float thr, unsigned long long give_up) {
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  PairOut out{c->pairs.as<wld_pair>(), cnt, c->pair_cap};
  ScopedStageTimer tm(c, WLD_STAGE_PAIR_REFINE);
  // the candidate count lives on the device: a fixed grid of blocks strides over it
  const dim3 grid((unsigned)(c->sm_count * 4)), block(32 * kRefineWarps);
  const int nl = c->geom.n_limbs;
  auto kern = nl == 2 ? pair_refine_kernel<2> : nl == 3 ? pair_refine_kernel<3> : pair_refine_kernel<4>;
  if (nl < 2 || nl > 4) return c->fail(WLD_ERR_INVALID, "the refinement needs 2..4 limbs");
  kern<<<grid, block, 0, c->stream>>>(c->codes.as<uint8_t>(), c->ldc, c->maj.as<int8_t>(), c->mnr.as<int8_t>(),
                                      c->glimb.as<uint16_t>(), c->geom.limb_bits, c->cand.as<uint2>(), cnt + 5, c->cand_cap,
                                      give_up, cnt + 6, thr, out);
  tm.launched();
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

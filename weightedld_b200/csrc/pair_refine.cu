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
// Cost: 2 code rows (2 * ldc bytes, mostly L2 hits: candidates arrive grouped by tile) + the weights (8 * ldc
// bytes, L2 resident) per candidate, ~14 integer instructions per sequence.  The f64 statistics run on batches
// of up to 32 candidates, one per lane, because B200's FP64 pipe is narrow.
#include "common.cuh"
#include "pair_epilogue.cuh"

namespace wld {
namespace {

constexpr int kRefineWarps = 8;

__device__ __forceinline__ uint32_t rep4(int sym) { return sym < 0 ? 0xffffffffu : (uint32_t)sym * 0x01010101u; }  // 0xff never matches a code

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(32 * kRefineWarps) pair_refine_kernel(
    const uint8_t* __restrict__ codes, int64_t ldc, const int8_t* __restrict__ maj, const int8_t* __restrict__ mnr,
    const unsigned long long* __restrict__ qi, const uint2* __restrict__ cand, const unsigned long long* __restrict__ n_cand,
    unsigned long long cap, float thr, PairOut out) {
  const int lane = threadIdx.x & 31;
  const unsigned long long n = min(*n_cand, cap);
  const unsigned long long n_warps = (unsigned long long)gridDim.x * kRefineWarps;
  // lane k of the warp keeps the sums of the k-th candidate of the current batch
  unsigned long long mAB = 0, mAb = 0, maB = 0, mab = 0;
  uint32_t mi = 0, mj = 0;
  int batch = 0;
  auto flush = [&]() {
    bool keep = lane < batch;
    float d = 0.f, dp = 0.f, r2 = 0.f;
    if (keep) {
      const double AB = (double)mAB, Ab = (double)mAb, aB = (double)maB, ab = (double)mab;  // < 2^53: exact
      const double A = AB + Ab, B = AB + aB, T = A + (aB + ab);
      // an empty marginal is NaN in the reference and dropped by lib.rs:660 (the screen never proposes one)
      keep = A > 0.0 && B > 0.0 && T - A > 0.0 && T - B > 0.0 && ld_stats_exact(AB, Ab, aB, ab, thr, d, dp, r2);
    }
    emit_pairs_warp(keep, mi, mj, d, dp, r2, out);
    batch = 0;
  };
  for (unsigned long long k = (unsigned long long)blockIdx.x * kRefineWarps + (threadIdx.x >> 5); k < n; k += n_warps) {
    const uint2 ij = cand[k];
    const uint8_t* ra = codes + (int64_t)ij.x * ldc;
    const uint8_t* rb = codes + (int64_t)ij.y * ldc;
    const uint32_t aM = rep4(maj[ij.x]), am = rep4(mnr[ij.x]), bM = rep4(maj[ij.y]), bm = rep4(mnr[ij.y]);
    unsigned long long AB = 0, Ab = 0, aB = 0, ab = 0;
    for (int64_t s0 = 16 * lane; s0 < ldc; s0 += 512) {  // ldc is a multiple of 128; the padding holds code 5 and q = 0
      const uint4 ca = __ldg(reinterpret_cast<const uint4*>(ra + s0));
      const uint4 cb = __ldg(reinterpret_cast<const uint4*>(rb + s0));
      const uint32_t wa[4] = {ca.x, ca.y, ca.z, ca.w}, wb[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t eaM = __vcmpeq4(wa[w], aM), eam = __vcmpeq4(wa[w], am);
        const uint32_t ebM = __vcmpeq4(wb[w], bM), ebm = __vcmpeq4(wb[w], bm);
        const uint32_t xAB = eaM & ebM, xAb = eaM & ebm, xaB = eam & ebM, xab = eam & ebm;
        if ((xAB | xAb | xaB | xab) == 0u) continue;
        const ulonglong2 q01 = __ldg(reinterpret_cast<const ulonglong2*>(qi + s0 + 4 * w));
        const ulonglong2 q23 = __ldg(reinterpret_cast<const ulonglong2*>(qi + s0 + 4 * w + 2));
        const unsigned long long qv[4] = {q01.x, q01.y, q23.x, q23.y};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t bit = 0x80u << (8 * b);
          if (xAB & bit) AB += qv[b];
          if (xAb & bit) Ab += qv[b];
          if (xaB & bit) aB += qv[b];
          if (xab & bit) ab += qv[b];
        }
      }
    }
    AB = warp_sum_u64(AB);
    Ab = warp_sum_u64(Ab);
    aB = warp_sum_u64(aB);
    ab = warp_sum_u64(ab);
    if (lane == batch) {
      mAB = AB; mAb = Ab; maB = aB; mab = ab;
      mi = ij.x; mj = ij.y;
    }
    if (++batch == 32) flush();
  }
  if (batch > 0) flush();
}

}  // namespace

int run_pair_refine(wld_ctx* c, float thr) {
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  PairOut out{c->pairs.as<wld_pair>(), cnt, c->pair_cap};
  ScopedStageTimer tm(c, WLD_STAGE_PAIR_REFINE);
  // the candidate count lives on the device: a fixed grid of warps strides over it
  pair_refine_kernel<<<c->sm_count * 4, 32 * kRefineWarps, 0, c->stream>>>(
      c->codes.as<uint8_t>(), c->ldc, c->maj.as<int8_t>(), c->mnr.as<int8_t>(), c->qi.as<unsigned long long>(),
      c->cand.as<uint2>(), cnt + 5, c->cand_cap, thr, out);
  tm.launched();
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

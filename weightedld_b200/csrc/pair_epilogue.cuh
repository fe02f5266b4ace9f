// pair_epilogue.cuh — fused epilogue of the pair stage, shared by the tcgen05 and the SIMT kernel.
//
// Input: the four EXACT weighted haplotype sums of one site pair (integers held in f64):
//   AB = sum w [a=maj_a][b=maj_b]   (ld_obs[3], lib.rs:477-479)
//   Ab = sum w [a=maj_a][b=min_b]   (ld_obs[2])
//   aB = sum w [a=min_a][b=maj_b]   (ld_obs[1])
//   ab = sum w [a=min_a][b=min_b]   (ld_obs[0])
// over sequences that are major-or-minor at both sites (lib.rs:462-467).
//
// Output: D, D', r2 by lib.rs:482-518 evaluated operation for operation in f64 (no FMA
// contraction: explicit _rn intrinsics), rounded to f32 (the reference's LdStats type,
// lib.rs:382-387), then the reference's filter `r2 > threshold` in f32 (lib.rs:660).
//
// Because the sums are exact and scaling by 2^bits is exact, the result is bit-identical to the
// f64 restatement of lib.rs:455-521 run on the same fixed-point weights — that is what the parity
// tests assert (tests/test_gpu_parity.py).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/wld.h"

namespace wld {

struct PairOut {
  wld_pair* pairs;              // survivor buffer
  unsigned long long* count;    // survivors so far (may exceed cap: overflow protocol)
  unsigned long long cap;
};

// Division-free conservative pre-filter.  r2 = d^2/(PA Pa PB Pb) with d = PA*PB - AB/T, so in raw
// sums r2 = (A*B - AB*T)^2 / (A (T-A) B (T-B)).  A pair can only pass `(float)r2 > thr` if
// num >= thr_lo * den with thr_lo = thr - |thr|*1e-6 - 1e-24 (slack >> f64/f32 rounding of the
// exact path).  den == 0 (a marginal is empty, or no common valid sequence) is NaN in the
// reference (d is exactly 0 there, 0/0) and is dropped by lib.rs:660.
__device__ __forceinline__ bool ld_prefilter(double AB, double Ab, double aB, double ab, double thr_lo) {
  const double A = AB + Ab;
  const double B = AB + aB;
  const double T = A + (aB + ab);
  const double a = T - A;
  const double b = T - B;
  double num = A * B - AB * T;
  num = num * num;
  const double den = (A * a) * (B * b);
  return den > 0.0 && num >= thr_lo * den;
}

// The same test in fp32, for inputs that are fp32 ROUNDINGS of the exact sums (relative error of each
// input <= 3*2^-24, pre-scaled by a power of two so that T <= 2^21: no overflow in n*n, no
// underflow of den).  Error budget: |fl(A*B) - fl(AB*T) - exact| <= 8*2^-24*(A*B + AB*T) -> covered by
// the 2e-6 term; den carries <= 10*2^-24 relative error -> covered by thr_lo_f = thr_lo*(1 -+ 1e-5).
// Exact zeros stay exact zeros (sums of non-negative terms), so den > 0 is decided exactly.
// The FP64 pipe of B200 is narrow (profiles/r01_ncu_pair_umma_c3_v1.md: stall_math on DADD/DFMA),
// so only candidates that pass this test pay for the f64 statistics.
__device__ __forceinline__ bool ld_prefilter_f32(float AB, float Ab, float aB, float ab, float thr_lo_f,
                                                 bool thr_negative) {
  const float A = AB + Ab;
  const float B = AB + aB;
  const float a = aB + ab;
  const float b = Ab + ab;
  const float T = A + a;
  const float p1 = A * B, p2 = AB * T;
  const float n = fabsf(p1 - p2) + 2e-6f * (p1 + p2);  // upper bound of |A*B - AB*T|
  const float den = (A * a) * (B * b);
  return den > 0.0f && (thr_negative || n * n >= thr_lo_f * den);
}

// One-limb screen (pair_umma.cu, kScreen).  Inputs: the four sums of gain x TOP limb of the fixed-point weights,
// exact integers below 2^31 converted to fp32 (relative error <= 2^-24 each; their total is <= 2^31 because the
// quantiser bounds every limb column sum).  With top_min the smallest top limb of a nonzero weight, the true sums
// y (in units of 2^(bits (NL-1))) satisfy x <= y <= x (1 + eta), eta = 1/top_min, hence with P = AB ab, Q = Ab aB:
//     |y1 y4 - y2 y3| <= |P - Q| + ((1 + eta)^2 - 1) max(P, Q)      and      every marginal of y >= that of x,
// so  r2(y) <= (|P - Q| + kappa max(P, Q))^2 / (A a B b)  with kappa = (1 + eta)^2 - 1 + 1e-5; the 1e-5 covers the
// fp32 roundings of P, Q and their difference (<= 8 * 2^-24 (P + Q)), and thr_lo_f (lowered by 1e-5 relative,
// pair_umma.cu) covers those of the denominator and of the square.  x = 0 implies y = 0, so an empty marginal
// (NaN in the reference, dropped by lib.rs:660) is decided exactly.  No overflow: P, Q < 2^62, n^2 < 2^126,
// A a, B b <= 2^60.  A pair the exact path would keep therefore always passes; the test is symmetric under
// exchanging the two rows (AB, Ab) <-> (aB, ab).
__device__ __forceinline__ bool ld_screen_f32(float AB, float Ab, float aB, float ab, float kappa, float thr_lo_f) {
  const float P = AB * ab, Q = Ab * aB;
  const float n = fabsf(P - Q) + kappa * fmaxf(P, Q);
  const float den1 = (AB + Ab) * (aB + ab), den2 = (AB + aB) * (Ab + ab);
  return den1 > 0.0f && den2 > 0.0f && n * n >= (thr_lo_f * den1) * den2;
}

__host__ __device__ inline double ld_thr_lo(float thr) {
  const double t = (double)thr;
  return t - (t < 0 ? -t : t) * 1e-6 - 1e-24;
}

// lib.rs:482-518 in f64.  Returns true when the pair survives `r2 > thr` (f32 compare).
// `python_skip` adds WeightedLD.py:234-237: skip when round(PA,1) == 1.0 or round(PB,1) == 1.0.  PA is a
// numpy float64 there, whose __round__ is rint(x*10)/10 (ties to even), so the test is fl(PA*10) >= 9.5.
__device__ __forceinline__ bool ld_stats_exact(double AB, double Ab, double aB, double ab, float thr, float& d_out,
                                               float& dprime_out, float& r2_out, bool python_skip = false) {
  // total_weight, PA, PB, ld_obs[3] as the reference accumulates them (lib.rs:469-479); exact here.
  const double total = __dadd_rn(__dadd_rn(AB, Ab), __dadd_rn(aB, ab));
  double PA = __dadd_rn(AB, Ab);
  double PB = __dadd_rn(AB, aB);
  double o3 = AB;
  double Pa = __dsub_rn(total, PA);  // lib.rs:482
  double Pb = __dsub_rn(total, PB);  // lib.rs:483
  double o2 = __dsub_rn(PA, o3);     // lib.rs:484
  double o1 = __dsub_rn(PB, o3);     // lib.rs:485
  double o0 = __dsub_rn(Pa, o1);     // lib.rs:486
  PA = __ddiv_rn(PA, total);         // lib.rs:488-495
  PB = __ddiv_rn(PB, total);
  Pa = __ddiv_rn(Pa, total);
  Pb = __ddiv_rn(Pb, total);
  o0 = __ddiv_rn(o0, total);
  o1 = __ddiv_rn(o1, total);
  o2 = __ddiv_rn(o2, total);
  o3 = __ddiv_rn(o3, total);
  if (python_skip && (__dmul_rn(PA, 10.0) >= 9.5 || __dmul_rn(PB, 10.0) >= 9.5)) return false;
  const double PAB = __dmul_rn(PA, PB);  // lib.rs:497-500
  const double PAb = __dmul_rn(PA, Pb);
  const double PaB = __dmul_rn(Pa, PB);
  const double Pab = __dmul_rn(Pa, Pb);
  const double d = __ddiv_rn(
      __dadd_rn(__dadd_rn(__dadd_rn(__dsub_rn(PAB, o3), __dsub_rn(Pab, o0)), __dsub_rn(o2, PAb)), __dsub_rn(o1, PaB)),
      4.0);  // lib.rs:502
  double den;  // lib.rs:504-515
  if (d < 0.0) {
    den = fmax(-o0, -o3);
    if (den == 0.0) den = fmin(-o0, -o3);
  } else {
    den = fmin(o1, o2);
    if (den == 0.0) den = fmax(o1, o2);
  }
  const double dprime = __ddiv_rn(d, den);  // lib.rs:516
  const double r2 =
      __ddiv_rn(__dmul_rn(d, d), __dmul_rn(__dmul_rn(__dmul_rn(PA, Pa), PB), Pb));  // lib.rs:518
  d_out = (float)d;
  dprime_out = (float)dprime;
  r2_out = (float)r2;
  return r2_out > thr;  // lib.rs:660 (NaN fails)
}

// WLD_COMPAT_PYTHON.  WeightedLD.py:186-211 calls the major / dominant-minor allele of each site PER PAIR,
// after deleting the sequences that hold code 5 at either site; the Gram recast calls them per site.  Both
// give the same alleles unless the deletions can reorder a site's top symbols.  aux[k] = {n5, margin} of
// kept site k: n5 = sequences with code 5, margin = min(count[major] - count[minor], count[minor] -
// count[third]) (third = 0 when absent).  Deleting at most n5_j sequences cannot change site i's call, nor
// empty its minor, while n5_j < margin_i; pairs that fail this test are left to the per-pair kernel
// (pair_python.cu) and must not be emitted by the Gram kernels.
__device__ __forceinline__ bool py_flagged(const uint2* __restrict__ aux, uint32_t i, uint32_t j) {
  const uint2 a = aux[i], b = aux[j];
  return (b.x > 0u && b.x >= a.y) || (a.x > 0u && a.x >= b.y);
}

// Candidates of the screen: kept-site index pairs, compacted like the survivors (count may exceed cap: the host
// then grows the buffer and repeats the screen; cap = 0 just counts).
struct CandOut {
  uint2* buf;
  unsigned long long* count;
  unsigned long long cap;
};
__device__ __forceinline__ void emit_cand_warp(bool cand, uint32_t site_a, uint32_t site_b, const CandOut& out) {
  const unsigned ballot = __ballot_sync(0xffffffffu, cand);
  if (ballot == 0) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == __ffs(ballot) - 1) base = atomicAdd(out.count, (unsigned long long)__popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (cand) {
    const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
    if (slot < out.cap) out.buf[slot] = make_uint2(site_a, site_b);
  }
}

// Warp-aggregated compaction: one atomicAdd per warp, survivors written to consecutive slots.
// Must be called by all 32 lanes of a converged warp.
__device__ __forceinline__ void emit_pairs_warp(bool keep, uint32_t site_a, uint32_t site_b, float d, float dprime,
                                                float r2, const PairOut& out) {
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  if (ballot == 0) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == __ffs(ballot) - 1) base = atomicAdd(out.count, (unsigned long long)__popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (keep) {
    const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
    if (slot < out.cap) {
      wld_pair p;
      p.site_a = site_a;
      p.site_b = site_b;
      p.d = d;
      p.d_prime = dprime;
      p.r2 = r2;
      out.pairs[slot] = p;
    }
  }
}

}  // namespace wld

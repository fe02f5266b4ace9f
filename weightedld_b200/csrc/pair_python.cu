// pair_python.cu — WLD_COMPAT_PYTHON only: the site pairs whose alleles WeightedLD.py calls differently
// from the per-site call the Gram kernels use.
//
// WeightedLD.py:186-211 deletes, per pair, every sequence holding code 5 at either site and only then
// picks each site's major / dominant-minor symbol (np.unique + argsort(-counts): descending count, ties
// to the smaller code — the same rule as the scan of lib.rs:126-140).  Without code 5 at the partner
// site nothing is deleted and both calls agree; with n5_j deletions they still agree while n5_j is below
// site i's `margin` (pair_epilogue.cuh, py_flagged).  The Gram kernels skip the remaining ("flagged")
// pairs and this file recomputes them the Python way, one warp per pair, straight from the 0..5 code
// matrix and the fixed-point weights: count the symbols among the surviving sequences, call the alleles,
// accumulate the four exact weighted sums, and run the shared f64 epilogue.  Flagged pairs are rare
// (they need code 5 at one site AND a near-tie or a very rare minor at the other), so this O(N)-per-pair
// CUDA-core pass is not on the roofline-relevant path; it exists for parity with the Python reference.
#include "common.cuh"
#include "pair_epilogue.cuh"

namespace wld {
namespace {

__global__ void site_aux_kernel(const uint32_t* __restrict__ hist, int64_t cols_padded,
                                const int32_t* __restrict__ site_map, int64_t n_kept, uint2* __restrict__ aux) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_kept) return;
  const int64_t col = site_map[k];
  uint32_t top[3] = {0u, 0u, 0u};  // three largest counts of codes 0..4, descending
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    uint32_t v = hist[(int64_t)c * cols_padded + col];
#pragma unroll
    for (int t = 0; t < 3; ++t)
      if (v > top[t]) {
        const uint32_t o = top[t];
        top[t] = v;
        v = o;
      }
  }
  const uint32_t margin = min(top[0] - top[1], top[1] - top[2]);
  aux[k] = make_uint2(hist[5 * cols_padded + col], margin);
}

// np.unique + argsort(-counts) on five counts: first and second symbol by descending count, ties to the
// smaller code; false when fewer than two symbols occur (skip_site, WeightedLD.py:197-201).
__device__ __forceinline__ bool call_alleles(const uint32_t* cnt, int& major, int& minor) {
  major = -1;
  minor = -1;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const uint32_t majc = major >= 0 ? cnt[major] : 0u;
    const uint32_t minc = minor >= 0 ? cnt[minor] : 0u;
    if (cnt[s] > majc) {
      minor = major;
      major = s;
    } else if (cnt[s] > minc) {
      minor = s;
    }
  }
  return minor >= 0;
}

constexpr int kFixWarps = 4;

// grid.x: chunks of 32*kFixWarps partner sites j; grid.y strides over sites i.
__global__ void __launch_bounds__(32 * kFixWarps) pair_python_kernel(
    const uint8_t* __restrict__ codes, int64_t ldc, int64_t n_kept, int64_t n_seqs, const double* __restrict__ q,
    const uint2* __restrict__ aux, const int8_t* __restrict__ mnr, int part, int nparts, float thr, PairOut out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t j = ((int64_t)blockIdx.x * kFixWarps + warp) * 32 + lane;
  const bool j_ok = j < n_kept && mnr[j] >= 0;
  for (int64_t i = blockIdx.y; i < n_kept; i += gridDim.y) {
    if (mnr[i] < 0) continue;  // monomorphic before any deletion stays monomorphic: no pair (WeightedLD.py:197)
    const bool mine = j_ok && j > i && (int)((i + j) % nparts) == part && py_flagged(aux, (uint32_t)i, (uint32_t)j);
    unsigned todo = __ballot_sync(0xffffffffu, mine);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t jj = __shfl_sync(0xffffffffu, j, src);
      const uint8_t* ra = codes + i * ldc;
      const uint8_t* rb = codes + jj * ldc;
      // pass 1: symbol counts among sequences that are < 5 at both sites (WeightedLD.py:183-186)
      uint32_t ca[5] = {0, 0, 0, 0, 0}, cb[5] = {0, 0, 0, 0, 0};
      for (int64_t s0 = 4 * lane; s0 < ldc; s0 += 128) {  // ldc is a multiple of 128; pad = 5
        const uint32_t wa = *reinterpret_cast<const uint32_t*>(ra + s0), wb = *reinterpret_cast<const uint32_t*>(rb + s0);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t xa = (wa >> (8 * b)) & 0xffu, xb = (wb >> (8 * b)) & 0xffu;
          if (xa < 5u && xb < 5u) {
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              ca[c] += xa == (uint32_t)c;
              cb[c] += xb == (uint32_t)c;
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ca[c] += __shfl_xor_sync(0xffffffffu, ca[c], o);
          cb[c] += __shfl_xor_sync(0xffffffffu, cb[c], o);
        }
      int a_maj, a_min, b_maj, b_min;
      const bool ok_a = call_alleles(ca, a_maj, a_min), ok_b = call_alleles(cb, b_maj, b_min);
      if (!(ok_a && ok_b)) continue;  // warp-uniform
      // pass 2: exact weighted sums over sequences that are major-or-minor at both sites (WeightedLD.py:214-222)
      double AB = 0.0, Ab = 0.0, aB = 0.0, ab = 0.0;  // integers < 2^53: exact in any order
      for (int64_t s0 = 4 * lane; s0 < ldc; s0 += 128) {
        const uint32_t wa = *reinterpret_cast<const uint32_t*>(ra + s0), wb = *reinterpret_cast<const uint32_t*>(rb + s0);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xa = (int)((wa >> (8 * b)) & 0xffu), xb = (int)((wb >> (8 * b)) & 0xffu);
          if (xa >= 5 || xb >= 5) continue;
          const bool aM = xa == a_maj, am = xa == a_min, bM = xb == b_maj, bm = xb == b_min;
          if ((aM || am) && (bM || bm)) {
            const double w = q[s0 + b];
            if (aM && bM) AB += w;
            else if (aM) Ab += w;
            else if (bM) aB += w;
            else ab += w;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        AB += __shfl_xor_sync(0xffffffffu, AB, o);
        Ab += __shfl_xor_sync(0xffffffffu, Ab, o);
        aB += __shfl_xor_sync(0xffffffffu, aB, o);
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
      }
      if (lane == 0) {
        float d, dp, r2;
        const double A = AB + Ab, B = AB + aB, T = A + (aB + ab);
        // same guard as the Gram kernels' pre-filter: an empty marginal is NaN in the reference and dropped
        if (A > 0.0 && B > 0.0 && T - A > 0.0 && T - B > 0.0 && ld_stats_exact(AB, Ab, aB, ab, thr, d, dp, r2, true)) {
          const unsigned long long slot = atomicAdd(out.count, 1ull);
          if (slot < out.cap) {
            wld_pair p;
            p.site_a = (uint32_t)i;
            p.site_b = (uint32_t)jj;
            p.d = d;
            p.d_prime = dp;
            p.r2 = r2;
            out.pairs[slot] = p;
          }
        }
      }
    }
  }
}

}  // namespace

int run_pair_python_prepare(wld_ctx* c) {
  const int64_t L = c->n_kept;
  WLD_CUDA(c, c->py_aux.ensure(sizeof(uint2) * (size_t)std::max<int64_t>(L, 1)));
  if (L == 0) return WLD_OK;
  site_aux_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(c->hist.as<uint32_t>(), c->cols_padded,
                                                                     c->site_map.as<int32_t>(), L, c->py_aux.as<uint2>());
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

int run_pair_python_fixup(wld_ctx* c, float thr) {
  const int64_t L = c->n_kept;
  if (L < 2) return WLD_OK;
  PairOut out{c->pairs.as<wld_pair>(), c->counters.as<unsigned long long>(), c->pair_cap};
  dim3 grid((unsigned)((L + 32 * kFixWarps - 1) / (32 * kFixWarps)), (unsigned)std::min<int64_t>(L, 32768));
  pair_python_kernel<<<grid, 32 * kFixWarps, 0, c->stream>>>(c->codes.as<uint8_t>(), c->ldc, L, c->n_seqs,
                                                            c->q.as<double>(), c->py_aux.as<uint2>(), c->mnr.as<int8_t>(),
                                                            c->part, c->nparts, thr, out);
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

// henikoff.cu — stage 2: Henikoff position-based sequence weights over the kept sites.
//
// Reference: henikoff_site_contributions lib.rs:360-380, henikoff_weights lib.rs:340-358.
//   per site:  k = #{x in A,C,G,T,-: count>0}            lib.rs:363 (116-124)
//              code<=4 -> 1/(k*count[code])              lib.rs:366-370
//              code 5  -> (sum of the above over the known sequences)/k   lib.rs:373-378
//   per sequence: sum over sites, then divide by the maximum  lib.rs:354-355
//
// The reference materialises an L x N f32 matrix (lib.rs:341); here the per-site contribution is
// a 6-entry table and the per-sequence sum is a segmented reduction over the site-major code
// matrix: column-histogram pass (already done in stage 1) + per-sequence gather-sum.
// Accumulation is f64 in a fixed order (site chunks of 256 in ascending order, then chunk
// partials in ascending order), so results are deterministic; tolerance vs the f64 oracle 1e-9.
// HBM-bound: reads n_kept*ldc code bytes once, writes 4*n_seqs (+ chunk partials).
#include "common.cuh"

namespace wld {
namespace {

constexpr int kSiteChunk = 256;
constexpr int kAccThreads = 256;

// python_mode (WeightedLD.py:125-145): contribution 1/count[code] (the scalar unique_base of
// WeightedLD.py:132 cancels in the normalisation), code-5 fill = sum of the known contributions divided
// by the number of KNOWN sequences at the site.
__global__ void table_kernel(const uint32_t* __restrict__ hist, int64_t cols_padded,
                             const int32_t* __restrict__ site_map, int64_t n_kept, bool python_mode,
                             double* __restrict__ table) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_kept) return;
  const int64_t col = site_map[k];
  uint32_t n[5];
  int distinct = 0;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    n[c] = hist[(int64_t)c * cols_padded + col];
    distinct += n[c] > 0;  // lib.rs:116-124
  }
  const double kd = python_mode ? 1.0 : (double)distinct;
  double total = 0.0;  // lib.rs:364-371: sum of the contributions of the known sequences
  double t[8];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    t[c] = n[c] > 0 ? __ddiv_rn(1.0, __dmul_rn(kd, (double)n[c])) : 0.0;  // lib.rs:368
    if (n[c] > 0) total = __dadd_rn(total, __dmul_rn((double)n[c], t[c]));
  }
  if (python_mode) {
    const double known = (double)n[0] + (double)n[1] + (double)n[2] + (double)n[3] + (double)n[4];
    t[5] = __ddiv_rn(total, known);  // WeightedLD.py:142-145
  } else {
    t[5] = __ddiv_rn(total, kd);  // lib.rs:373 (0/0 = NaN when the site has no known symbol)
  }
  t[6] = 0.0;
  t[7] = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) table[k * 8 + c] = t[c];
}

// grid (seq blocks of 1024, site chunks of 256); each thread owns four consecutive sequences.
// The gather-sum is bound by shared-memory lookups (one 8-byte load per cell), so sites are taken two at a
// time: s_tab2[pair][6*code_a + code_b] = table[a][code_a] + table[b][code_b] — one lookup and one f64 add per
// TWO cells.  (Summation order per sequence: pairs of sites in ascending order; deterministic.)
__global__ void __launch_bounds__(kAccThreads) accumulate_kernel(const uint8_t* __restrict__ codes, int64_t ldc,
                                                                 int64_t n_kept, int64_t n_seqs, int64_t seq_lo,
                                                                 int64_t seq_hi, const double* __restrict__ table,
                                                                 double* __restrict__ partial) {
  __shared__ double s_tab2[kSiteChunk / 2][36];
  const int64_t k0 = (int64_t)blockIdx.y * kSiteChunk;
  const int nk = (int)min((int64_t)kSiteChunk, n_kept - k0);
  const int npairs = (nk + 1) / 2;
  for (int i = threadIdx.x; i < npairs * 36; i += kAccThreads) {
    const int pr = i / 36, e = i % 36;
    const double ta = table[(k0 + 2 * pr) * 8 + e / 6];
    const double tb = 2 * pr + 1 < nk ? table[(k0 + 2 * pr + 1) * 8 + e % 6] : 0.0;
    (&s_tab2[0][0])[i] = __dadd_rn(ta, tb);
  }
  __syncthreads();
  // this launch covers the sequences [seq_lo, seq_hi) (a multi-GPU shard, or all of them), four per thread
  const int64_t s0 = (seq_lo / 4 + (int64_t)blockIdx.x * kAccThreads + threadIdx.x) * 4;
  if (s0 >= ldc || s0 >= seq_hi) return;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const uint8_t* p = codes + k0 * ldc + s0;
  auto add_pair = [&](int pr, uint32_t wa, uint32_t wb) {
    // the four table indices 6 * code_a + code_b at once: codes are 0..5 (gather_kernel writes nothing else, the
    // padding is 5), so no byte of 6 * wa + wb exceeds 35 and nothing carries
    const uint32_t z = wa * 6u + wb;
    const double* row = s_tab2[pr];
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[b] = __dadd_rn(acc[b], row[(z >> (8 * b)) & 0xffu]);
  };
  int pr = 0;
  for (; 2 * pr + 8 <= nk; pr += 4) {  // four pairs = eight site rows in flight (the loop is latency-bound)
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = __ldg(reinterpret_cast<const uint32_t*>(p + (int64_t)(2 * pr + u) * ldc));
#pragma unroll
    for (int u = 0; u < 4; ++u) add_pair(pr + u, w[2 * u], w[2 * u + 1]);
  }
  for (; pr < npairs; ++pr) {
    const uint32_t wa = __ldg(reinterpret_cast<const uint32_t*>(p + (int64_t)(2 * pr) * ldc));
    // odd tail: the partner row does not exist; code 0 selects table[a][code_a] + 0
    const uint32_t wb = 2 * pr + 1 < nk ? __ldg(reinterpret_cast<const uint32_t*>(p + (int64_t)(2 * pr + 1) * ldc)) : 0u;
    add_pair(pr, wa, wb);
  }
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (s0 + b >= seq_lo && s0 + b < seq_hi) partial[(int64_t)blockIdx.y * n_seqs + s0 + b] = acc[b];
}

// Sum chunk partials in ascending chunk order for the sequences [seq_lo, seq_hi).
__global__ void reduce_kernel(const double* __restrict__ partial, int64_t n_chunks, int64_t n_seqs, int64_t seq_lo,
                              int64_t seq_hi, double* __restrict__ w64) {
  const int64_t s = seq_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= seq_hi) return;
  double v = 0.0;
  for (int64_t ch = 0; ch < n_chunks; ++ch) v = __dadd_rn(v, partial[ch * n_seqs + s]);
  w64[s] = v;
}

// lib.rs:355 fold(0.0, max) over ALL sequences (NaN ignored like f32::max): sums are >= 0 or NaN; for
// non-negative doubles the bit pattern is monotone, so an integer atomicMax implements it.
__global__ void max_kernel(const double* __restrict__ w64, int64_t n_seqs, unsigned long long* __restrict__ max_bits) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double v = s < n_seqs ? w64[s] : 0.0;
  unsigned long long bits = (s < n_seqs && v == v && v > 0.0) ? (unsigned long long)__double_as_longlong(v) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(max_bits, bits);
}

__global__ void normalize_kernel(double* __restrict__ w64, float* __restrict__ w32, int64_t n_seqs,
                                 const unsigned long long* __restrict__ max_bits) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seqs) return;
  const double mx = __longlong_as_double((long long)*max_bits);
  const double v = __ddiv_rn(w64[s], mx);  // lib.rs:355
  w64[s] = v;
  w32[s] = (float)v;  // the reference's weight type is f32 (lib.rs:340)
}

}  // namespace

// Per-sequence sums for the context's sequence shard (all sequences unless wld_set_seq_shard narrowed it): every
// sequence is summed whole, over all kept sites in the same order, by exactly one GPU — so the weights are
// bit-identical for any number of GPUs.  `finish` adds the maximum and the normalisation (lib.rs:355).
int run_henikoff(wld_ctx* c, ScopedStageTimer& tm, bool finish) {
  const int64_t n = c->n_seqs, L = c->n_kept;
  const int64_t lo = std::min(c->seq_lo, n), hi = c->seq_hi < 0 ? n : std::min(c->seq_hi, n);
  WLD_CUDA(c, c->w64.ensure(sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
  WLD_CUDA(c, c->w32.ensure(sizeof(float) * (size_t)std::max<int64_t>(n, 1)));
  WLD_CUDA(c, c->scalars.ensure(sizeof(double) * 8));
  if (n == 0) return WLD_OK;
  const int64_t n_chunks = (L + kSiteChunk - 1) / kSiteChunk;
  WLD_CUDA(c, c->table.ensure(sizeof(double) * 8 * (size_t)std::max<int64_t>(L, 1)));
  WLD_CUDA(c, c->partial.ensure(sizeof(double) * (size_t)std::max<int64_t>(n_chunks, 1) * (size_t)n));
  // a shard leaves the other sequences' sums at zero, so that the exchange may also be a SUM over ranks
  if (!finish) WLD_CUDA(c, cudaMemsetAsync(c->w64.p, 0, sizeof(double) * (size_t)n, c->stream));
  if (hi > lo) {
    if (L > 0) {
      table_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(c->hist.as<uint32_t>(), c->cols_padded,
                                                                      c->site_map.as<int32_t>(), L,
                                                                      c->compat == WLD_COMPAT_PYTHON, c->table.as<double>());
      tm.launched();
      const int64_t span = std::min(round_up(hi, 4), c->ldc) - lo / 4 * 4;  // whole groups of four sequences
      dim3 grid((unsigned)((span / 4 + kAccThreads - 1) / kAccThreads), (unsigned)n_chunks);
      if (grid.y > 65535) return c->fail(WLD_ERR_UNSUPPORTED, "too many kept sites for the Henikoff grid");
      accumulate_kernel<<<grid, kAccThreads, 0, c->stream>>>(c->codes.as<uint8_t>(), c->ldc, L, n, lo, hi,
                                                             c->table.as<double>(), c->partial.as<double>());
      tm.launched();
    }
    reduce_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, c->stream>>>(c->partial.as<double>(), n_chunks, n, lo, hi,
                                                                           c->w64.as<double>());
    tm.launched();
  }
  WLD_CUDA(c, cudaGetLastError());
  return finish ? run_henikoff_finish(c, tm) : WLD_OK;
}

int run_henikoff_finish(wld_ctx* c, ScopedStageTimer& tm) {
  const int64_t n = c->n_seqs;
  if (n == 0) return WLD_OK;
  WLD_CUDA(c, cudaMemsetAsync(c->scalars.p, 0, sizeof(double) * 8, c->stream));
  max_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->w64.as<double>(), n, c->scalars.as<unsigned long long>());
  normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->w64.as<double>(), c->w32.as<float>(), n,
                                                                      c->scalars.as<unsigned long long>());
  tm.launched(2);
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

// pair_prep.cu — stage 3a: fixed-point weights, limb split and operand expansion for the Gram.
//
// The pair stage of the reference (single_weighted_ld_pair, lib.rs:390-521) needs, per site pair,
// the four weighted haplotype sums over sequences whose symbol is major-or-minor at BOTH sites
// (lib.rs:462-479).  With indicator columns Maj_i[s] = [code==major_i], Min_i[s] = [code==minor_i]
// those sums are the 2x2 block  (Maj_i,Min_i)^T diag(w) (Maj_j,Min_j)  of a Gram matrix.
//
// Exactness: weights become integers q[s] = rint(w[s]/max(w) * 2^(b*NL)) and are split into NL
// limbs of b bits (top limb may equal 2^b).  Limbs <= 256 and indicators are exact in bf16, every
// product is a small integer, and fp32 accumulation is exact while a sum stays <= 2^24, which the
// host checks from the limb column sums (shrinking b if ever needed).  The Gram is therefore an
// exact integer computation, independent of tile order and GPU count.
//
// Operands (bf16, K-major = sequence index contiguous, K padded to 64 with zeros):
//   opA : [a_rows][Kp]       row 2*i+alpha           = indicator (alpha: 0 major, 1 minor) of site i
//   opB : [groups*128][Kp]   row g*128 + r*2NL + beta*NL + l = indicator_beta(site g*SPG+r) * limb_l
//         SPG = floor(128/(2NL)) sites per 128-row group; unused rows of a group are zero.
// HBM-bound: reads n_kept*Kp code bytes (twice), writes (2 + 2NL)*2 bytes per (site, sequence).
#include "common.cuh"

namespace wld {
namespace {

__device__ __forceinline__ uint16_t bf16_bits_of_small_int(uint32_t v) {
  // v <= 256 is exactly representable: bf16 = upper 16 bits of the f32 pattern.
  return (uint16_t)(__float_as_uint((float)v) >> 16);
}

// One block.  flags: bit0 = invalid weight seen, bit1 = all weights equal.
// limb_sums[l] = sum over sequences of limb l (u64) — an upper bound of every Gram entry of that limb.
__global__ void __launch_bounds__(1024) quantize_kernel(const float* __restrict__ w, int64_t n_seqs, int64_t ldc,
                                                        int n_limbs, int limb_bits, double* __restrict__ q,
                                                        uint16_t* __restrict__ limbs,
                                                        unsigned long long* __restrict__ limb_sums,
                                                        int* __restrict__ flags) {
  __shared__ float s_red[32], s_min[32];
  __shared__ unsigned long long s_sum[4][32];
  __shared__ int s_bad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  float mx = 0.0f, mn = INFINITY;
  int bad = 0;
  for (int64_t s = threadIdx.x; s < n_seqs; s += blockDim.x) {
    const float v = w[s];
    if (!(v >= 0.0f) || isinf(v)) bad = 1;
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) {
    s_red[warp] = mx;
    s_min[warp] = mn;
  }
  if (bad) s_bad = 1;
  __syncthreads();
  mx = 0.0f;
  mn = INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
    mx = fmaxf(mx, s_red[i]);
    mn = fminf(mn, s_min[i]);
  }
  if (s_bad || !(mx > 0.0f)) {
    if (threadIdx.x == 0) *flags = 1;
    return;
  }
  if (threadIdx.x == 0) *flags = (mn == mx) ? 2 : 0;  // bit1: all weights equal
  const int total_bits = n_limbs * limb_bits;
  const double scale = (double)(1ull << total_bits);
  const uint32_t limb_mask = (1u << limb_bits) - 1u;
  unsigned long long sums[4] = {0, 0, 0, 0};
  for (int64_t s = threadIdx.x; s < ldc; s += blockDim.x) {
    unsigned long long qi = 0;  // may equal 2^total_bits (up to 2^32) for the largest weight
    if (s < n_seqs) qi = (unsigned long long)rint(__dmul_rn(__ddiv_rn((double)w[s], (double)mx), scale));
    q[s] = (double)qi;
    for (int l = 0; l < n_limbs; ++l) {
      const int shift = limb_bits * (n_limbs - 1 - l);
      uint32_t v = (uint32_t)(qi >> shift);
      if (l > 0) v &= limb_mask;  // the top limb keeps the carry (q may equal 2^total_bits)
      limbs[(int64_t)l * ldc + s] = bf16_bits_of_small_int(v);
      sums[l] += v;
    }
  }
  for (int l = 0; l < 4; ++l) {
    unsigned long long v = sums[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_sum[l][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long v = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += s_sum[threadIdx.x][i];
    limb_sums[threadIdx.x] = v;
  }
}

// 8 codes -> 8 bf16 values: value_bits where code == sym, else 0.
__device__ __forceinline__ uint4 select8(uint2 codes, int sym, const uint16_t* vals /*8, or nullptr => 1.0*/) {
  uint32_t out[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const uint32_t word = p < 2 ? codes.x : codes.y;
    const uint32_t c0 = (word >> (16 * (p & 1))) & 0xffu;
    const uint32_t c1 = (word >> (16 * (p & 1) + 8)) & 0xffu;
    const uint32_t v0 = vals ? vals[2 * p] : 0x3f80u;
    const uint32_t v1 = vals ? vals[2 * p + 1] : 0x3f80u;
    out[p] = ((int)c0 == sym ? v0 : 0u) | (((int)c1 == sym ? v1 : 0u) << 16);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

// opA: grid (K blocks of 2048, a_rows/2 sites).  Sites >= n_kept are zero rows.
__global__ void __launch_bounds__(256) expand_a_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int64_t n_kept,
                                                       const int8_t* __restrict__ maj, const int8_t* __restrict__ mnr,
                                                       int64_t kp, uint16_t* __restrict__ opA) {
  const int64_t i = blockIdx.y;
  const int64_t s0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  if (s0 >= kp) return;
  uint4 vmaj = make_uint4(0, 0, 0, 0), vmin = vmaj;
  if (i < n_kept) {
    const uint2 cw = __ldg(reinterpret_cast<const uint2*>(codes + i * ldc + s0));
    vmaj = select8(cw, maj[i], nullptr);  // maj/min == -1 never matches a code
    vmin = select8(cw, mnr[i], nullptr);
  }
  *reinterpret_cast<uint4*>(opA + (2 * i) * kp + s0) = vmaj;
  *reinterpret_cast<uint4*>(opA + (2 * i + 1) * kp + s0) = vmin;
}

// opB: grid (K blocks of 2048, groups*(SPG+1)).  blockIdx.y = g*(SPG+1) + r; r == SPG zero-fills the
// unused tail rows of the group.
__global__ void __launch_bounds__(256) expand_b_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int64_t n_kept,
                                                       const int8_t* __restrict__ maj, const int8_t* __restrict__ mnr,
                                                       const uint16_t* __restrict__ limbs, int n_limbs, int spg,
                                                       int64_t kp, uint16_t* __restrict__ opB) {
  const int64_t g = blockIdx.y / (spg + 1);
  const int r = (int)(blockIdx.y % (spg + 1));
  const int rps = 2 * n_limbs;
  const int64_t s0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  if (s0 >= kp) return;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  if (r == spg) {
    for (int row = spg * rps; row < 128; ++row) *reinterpret_cast<uint4*>(opB + (g * 128 + row) * kp + s0) = zero;
    return;
  }
  const int64_t j = g * spg + r;
  const int64_t row0 = g * 128 + (int64_t)r * rps;
  if (j >= n_kept) {
    for (int t = 0; t < rps; ++t) *reinterpret_cast<uint4*>(opB + (row0 + t) * kp + s0) = zero;
    return;
  }
  const uint2 cw = __ldg(reinterpret_cast<const uint2*>(codes + j * ldc + s0));
  const int sm = maj[j], sn = mnr[j];
  for (int l = 0; l < n_limbs; ++l) {
    const uint4 lv = __ldg(reinterpret_cast<const uint4*>(limbs + (int64_t)l * ldc + s0));
    uint16_t vals[8];
    vals[0] = lv.x & 0xffff; vals[1] = lv.x >> 16; vals[2] = lv.y & 0xffff; vals[3] = lv.y >> 16;
    vals[4] = lv.z & 0xffff; vals[5] = lv.z >> 16; vals[6] = lv.w & 0xffff; vals[7] = lv.w >> 16;
    *reinterpret_cast<uint4*>(opB + (row0 + l) * kp + s0) = select8(cw, sm, vals);
    *reinterpret_cast<uint4*>(opB + (row0 + n_limbs + l) * kp + s0) = select8(cw, sn, vals);
  }
}

}  // namespace

int run_pair_prep(wld_ctx* c, ScopedStageTimer& tm) {
  PairGeom& gm = c->geom;
  const int64_t n = c->n_seqs, L = c->n_kept;
  gm.n_limbs = c->n_limbs_opt;

  WLD_CUDA(c, c->q.ensure(sizeof(double) * (size_t)c->ldc));
  WLD_CUDA(c, c->limbs.ensure(sizeof(uint16_t) * 4 * (size_t)c->ldc));
  WLD_CUDA(c, c->counters.ensure(sizeof(unsigned long long) * 16));
  WLD_CUDA(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 16, c->stream));

  // All-equal weights (e.g. --unweighted, main.rs:150-153) need a single 0-bit limb: q == 1.
  // Otherwise start from 8-bit limbs and shrink only if a limb column sum could exceed 2^24.
  unsigned long long sums[4];
  int flag = 0;
  int bits = 8;
  for (;;) {
    quantize_kernel<<<1, 1024, 0, c->stream>>>(c->w32.as<float>(), n, c->ldc, gm.n_limbs, bits,
                                               c->q.as<double>(), c->limbs.as<uint16_t>(),
                                               c->counters.as<unsigned long long>() + 8,
                                               reinterpret_cast<int*>(c->counters.as<unsigned long long>() + 12));
    tm.launched();
    WLD_CUDA(c, cudaGetLastError());
    WLD_CUDA(c, cudaMemcpyAsync(sums, c->counters.as<unsigned long long>() + 8, sizeof sums, cudaMemcpyDeviceToHost,
                                c->stream));
    WLD_CUDA(c, cudaMemcpyAsync(&flag, c->counters.as<unsigned long long>() + 12, sizeof flag,
                                cudaMemcpyDeviceToHost, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
    if (flag & 1) return c->fail(WLD_ERR_INVALID, "weights must be finite, >= 0 and not all zero");
    if ((flag & 2) && !(gm.n_limbs == 1 && bits == 0)) {
      gm.n_limbs = 1;  // q == 1 for every sequence: one limb, exact for n_seqs <= 2^24
      bits = 0;
      continue;
    }
    unsigned long long worst = 0;
    for (int l = 0; l < gm.n_limbs; ++l) worst = std::max(worst, sums[l]);
    if (worst <= (1ull << 24)) break;
    if (--bits < 1) return c->fail(WLD_ERR_UNSUPPORTED, "n_seqs too large for exact fp32 accumulation");
  }
  gm.limb_bits = bits;
  gm.rows_per_site = 2 * gm.n_limbs;
  gm.sites_per_group = 128 / gm.rows_per_site;
  gm.k_padded = round_up(std::max<int64_t>(n, 1), 64);
  gm.a_rows = round_up(std::max<int64_t>(2 * L, 1), 128);
  gm.b_groups = std::max<int64_t>((L + gm.sites_per_group - 1) / gm.sites_per_group, 1);
  gm.b_groups = round_up(gm.b_groups, 2);  // an N tile is two groups

  const int64_t kp = gm.k_padded;
  WLD_CUDA(c, c->opA.ensure(sizeof(uint16_t) * (size_t)gm.a_rows * (size_t)kp));
  WLD_CUDA(c, c->opB.ensure(sizeof(uint16_t) * (size_t)gm.b_groups * 128 * (size_t)kp));
  const unsigned kblocks = (unsigned)((kp / 8 + 255) / 256);
  {
    dim3 grid(kblocks, (unsigned)(gm.a_rows / 2));
    if (grid.y > 65535 * 32) return c->fail(WLD_ERR_UNSUPPORTED, "too many kept sites");
    // grid.y limit is 65535: fold larger site counts into several launches
    for (int64_t y0 = 0; y0 < gm.a_rows / 2; y0 += 65535) {
      const unsigned ny = (unsigned)std::min<int64_t>(65535, gm.a_rows / 2 - y0);
      expand_a_kernel<<<dim3(kblocks, ny), 256, 0, c->stream>>>(
          c->codes.as<uint8_t>() + y0 * c->ldc, c->ldc, std::max<int64_t>(L - y0, 0), c->maj.as<int8_t>() + y0,
          c->mnr.as<int8_t>() + y0, kp, c->opA.as<uint16_t>() + 2 * y0 * kp);
      tm.launched();
    }
  }
  {
    const int spg = gm.sites_per_group;
    const int64_t groups_per_launch = 65535 / (spg + 1);
    for (int64_t g0 = 0; g0 < gm.b_groups; g0 += groups_per_launch) {
      const int64_t ng = std::min<int64_t>(groups_per_launch, gm.b_groups - g0);
      expand_b_kernel<<<dim3(kblocks, (unsigned)(ng * (spg + 1))), 256, 0, c->stream>>>(
          c->codes.as<uint8_t>() + g0 * spg * c->ldc, c->ldc, std::max<int64_t>(L - g0 * spg, 0),
          c->maj.as<int8_t>() + g0 * spg, c->mnr.as<int8_t>() + g0 * spg, c->limbs.as<uint16_t>(), gm.n_limbs, spg,
          kp, c->opB.as<uint16_t>() + g0 * 128 * kp);
      tm.launched();
    }
  }
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

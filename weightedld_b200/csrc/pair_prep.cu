// pair_prep.cu — stage 3a: fixed-point weights, limb split and operand expansion for the Gram.
//
// The pair stage of the reference (single_weighted_ld_pair, lib.rs:390-521) needs, per site pair,
// the four weighted haplotype sums over sequences whose symbol is major-or-minor at BOTH sites
// (lib.rs:462-479).  With indicator columns Maj_i[s] = [code==major_i], Min_i[s] = [code==minor_i]
// those sums are the 2x2 block  (Maj_i,Min_i)^T diag(w) (Maj_j,Min_j)  of a Gram matrix.
//
// Exactness: weights become integers q[s] = rint(w[s]/max(w) * (2^(b*NL) - 1)) and are split into NL
// limbs of b bits (<= 255).  Two operand encodings, both exact:
//   bf16 (kind::f16, fp32 accumulate): limbs and indicators are exact in bf16, every product is a
//        small integer, and fp32 accumulation is exact while a sum stays <= 2^24, which the host
//        checks from the limb column sums (shrinking b if ever needed);
//   u8   (kind::i8, s32 accumulate): exact for any n_seqs < 2^31/255.
// The Gram is therefore an exact integer computation, independent of tile order and GPU count.
//
// Operands (K-major = sequence index contiguous, K padded with zeros to one 128-byte swizzle atom):
//   opA : [a_rows][Kp]       row 2*i+alpha           = indicator (alpha: 0 major, 1 minor) of site i
//   opB : [groups*128][Kp]   row g*128 + r*2NL + beta*NL + l = indicator_beta(site g*SPG+r) * limb_l
//         SPG = floor(128/(2NL)) sites per 128-row group; unused rows of a group are zero.
// HBM-bound: reads n_kept*Kp code bytes (twice), writes (2 + 2NL)*2 bytes per (site, sequence).
#include <cmath>

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
                                                        uint16_t* __restrict__ limbs, uint8_t* __restrict__ limbs8,
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
  const double scale = total_bits > 0 ? (double)((1ull << total_bits) - 1ull) : 1.0;  // q <= 2^B - 1: limbs fit u8
  const uint32_t limb_mask = (1u << limb_bits) - 1u;
  unsigned long long sums[4] = {0, 0, 0, 0};
  for (int64_t s = threadIdx.x; s < ldc; s += blockDim.x) {
    unsigned long long qi = 0;
    if (s < n_seqs) qi = (unsigned long long)rint(__dmul_rn(__ddiv_rn((double)w[s], (double)mx), scale));
    q[s] = (double)qi;
    for (int l = 0; l < n_limbs; ++l) {
      const int shift = limb_bits * (n_limbs - 1 - l);
      uint32_t v = (uint32_t)(qi >> shift) & (limb_bits > 0 ? limb_mask : 1u);
      limbs[(int64_t)l * ldc + s] = (uint16_t)v;  // raw limb value 0..255; converted at expansion
      limbs8[(int64_t)l * ldc + s] = (uint8_t)v;  // the same as bytes for the u8 operands (SWAR expansion)
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

// Operand element encodings: bf16 bit pattern of a small integer (exact for v <= 256), or the u8 itself.
template <bool kI8>
__device__ __forceinline__ uint32_t elem_of(uint32_t v) {
  return kI8 ? v : (uint32_t)bf16_bits_of_small_int(v);
}

// One thread expands kSeq consecutive sequences of one site into 16-byte stores: 8 bf16 or 16 u8
// elements per operand row.  c[] holds the codes as packed words (4 per word).
template <bool kI8> struct Expand {
  static constexpr int kSeq = kI8 ? 16 : 8;       // sequences per thread
  static constexpr int kWords = kSeq / 4;         // 32-bit words of codes
};

// kSeq codes -> kSeq operand elements: vals[k] where code == sym, else 0; one 16-byte store.
template <bool kI8>
__device__ __forceinline__ void store_select(void* dst, const uint32_t* cw, int sym, const uint32_t* vals) {
  constexpr int kSeq = Expand<kI8>::kSeq;
  uint32_t e[kSeq];
#pragma unroll
  for (int k = 0; k < kSeq; ++k) {
    const int c = (int)((cw[k >> 2] >> (8 * (k & 3))) & 0xffu);
    e[k] = c == sym ? vals[k] : 0u;
  }
  if (kI8) {
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(e[0] | (e[1] << 8) | (e[2] << 16) | (e[3] << 24), e[4] | (e[5] << 8) | (e[6] << 16) | (e[7] << 24),
                   e[8 % kSeq] | (e[9 % kSeq] << 8) | (e[10 % kSeq] << 16) | (e[11 % kSeq] << 24),
                   e[12 % kSeq] | (e[13 % kSeq] << 8) | (e[14 % kSeq] << 16) | (e[15 % kSeq] << 24));
  } else {
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
  }
}
// Indicator rows: the element is the constant 1, so whole words can be built with SWAR compares.
template <bool kI8>
__device__ __forceinline__ void store_indicator(void* dst, const uint32_t* cw, int sym) {
  if (kI8) {
    const uint32_t rep = sym < 0 ? 0xffffffffu : (uint32_t)sym * 0x01010101u;  // 0xff never matches a code
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(__vcmpeq4(cw[0], rep) & 0x01010101u, __vcmpeq4(cw[1], rep) & 0x01010101u,
                   __vcmpeq4(cw[2], rep) & 0x01010101u, __vcmpeq4(cw[3], rep) & 0x01010101u);
  } else {
    uint32_t ones[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ones[k] = elem_of<false>(1u);
    store_select<false>(dst, cw, sym, ones);
  }
}
__device__ __forceinline__ void store_zero16(void* dst) { *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0); }

template <bool kI8>
__device__ __forceinline__ void load_codes(const uint8_t* p, uint32_t* cw) {
  if (kI8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    cw[0] = v.x; cw[1] = v.y; cw[2] = v.z; cw[3] = v.w;
  } else {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    cw[0] = v.x; cw[1] = v.y;
  }
}

// opA: grid (K blocks of 256*kSeq, a_rows/2 sites).  Sites >= n_kept are zero rows.
template <bool kI8>
__global__ void __launch_bounds__(256) expand_a_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int64_t n_kept,
                                                       const int8_t* __restrict__ maj, const int8_t* __restrict__ mnr,
                                                       int64_t kp, uint8_t* __restrict__ opA) {
  constexpr int ES = kI8 ? 1 : 2;
  constexpr int kSeq = Expand<kI8>::kSeq;
  const int64_t i = blockIdx.y;
  const int64_t s0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kSeq;
  if (s0 >= kp) return;
  uint8_t* r0 = opA + ((2 * i) * kp + s0) * ES;
  uint8_t* r1 = opA + ((2 * i + 1) * kp + s0) * ES;
  if (i < n_kept) {
    uint32_t cw[Expand<kI8>::kWords];
    load_codes<kI8>(codes + i * ldc + s0, cw);
    store_indicator<kI8>(r0, cw, maj[i]);  // maj/min == -1 never matches a code
    store_indicator<kI8>(r1, cw, mnr[i]);
  } else {
    store_zero16(r0);
    store_zero16(r1);
  }
}

// opB: grid (K blocks of 256*kSeq, groups*(SPG+1)).  blockIdx.y = g*(SPG+1) + r; r == SPG zero-fills the
// unused tail rows of the group.
template <bool kI8>
__global__ void __launch_bounds__(256) expand_b_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int64_t n_kept,
                                                       const int8_t* __restrict__ maj, const int8_t* __restrict__ mnr,
                                                       const uint16_t* __restrict__ limbs, int n_limbs, int spg,
                                                       int64_t kp, uint8_t* __restrict__ opB) {
  constexpr int ES = kI8 ? 1 : 2;
  constexpr int kSeq = Expand<kI8>::kSeq;
  const int64_t g = blockIdx.y / (spg + 1);
  const int r = (int)(blockIdx.y % (spg + 1));
  const int rps = 2 * n_limbs;
  const int64_t s0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kSeq;
  if (s0 >= kp) return;
  if (r == spg) {
    for (int row = spg * rps; row < 128; ++row) store_zero16(opB + ((g * 128 + row) * kp + s0) * ES);
    return;
  }
  const int64_t j = g * spg + r;
  const int64_t row0 = g * 128 + (int64_t)r * rps;
  if (j >= n_kept) {
    for (int t = 0; t < rps; ++t) store_zero16(opB + ((row0 + t) * kp + s0) * ES);
    return;
  }
  uint32_t cw[Expand<kI8>::kWords];
  load_codes<kI8>(codes + j * ldc + s0, cw);
  const int sm = maj[j], sn = mnr[j];
  if constexpr (kI8) {
    // u8 operands: indicator masks by SWAR compare once, then row = mask & limb bytes (4 elements per AND)
    const uint8_t* limbs8 = reinterpret_cast<const uint8_t*>(limbs) + sizeof(uint16_t) * 4 * (size_t)ldc;
    const uint32_t rm = sm < 0 ? 0xffffffffu : (uint32_t)sm * 0x01010101u;  // 0xff never matches a code
    const uint32_t rn = sn < 0 ? 0xffffffffu : (uint32_t)sn * 0x01010101u;
    uint32_t mm[4], mn[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      mm[w] = __vcmpeq4(cw[w], rm);
      mn[w] = __vcmpeq4(cw[w], rn);
    }
    for (int l = 0; l < n_limbs; ++l) {
      const uint4 lv = __ldg(reinterpret_cast<const uint4*>(limbs8 + (int64_t)l * ldc + s0));
      *reinterpret_cast<uint4*>(opB + (row0 + l) * kp + s0) = make_uint4(lv.x & mm[0], lv.y & mm[1], lv.z & mm[2], lv.w & mm[3]);
      *reinterpret_cast<uint4*>(opB + (row0 + n_limbs + l) * kp + s0) =
          make_uint4(lv.x & mn[0], lv.y & mn[1], lv.z & mn[2], lv.w & mn[3]);
    }
  } else {
  for (int l = 0; l < n_limbs; ++l) {
    uint32_t vals[kSeq];
#pragma unroll
    for (int h = 0; h < kSeq / 8; ++h) {
      const uint4 lv = __ldg(reinterpret_cast<const uint4*>(limbs + (int64_t)l * ldc + s0 + 8 * h));
      const uint32_t w[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        vals[8 * h + 2 * k] = elem_of<kI8>(w[k] & 0xffffu);
        vals[8 * h + 2 * k + 1] = elem_of<kI8>(w[k] >> 16);
      }
    }
    store_select<kI8>(opB + ((row0 + l) * kp + s0) * ES, cw, sm, vals);
    store_select<kI8>(opB + ((row0 + n_limbs + l) * kp + s0) * ES, cw, sn, vals);
  }
  }
}

}  // namespace

int run_pair_prep(wld_ctx* c, ScopedStageTimer& tm) {
  PairGeom& gm = c->geom;
  const int64_t n = c->n_seqs, L = c->n_kept;
  gm.n_limbs = c->n_limbs_opt;
  const bool i8 = c->pair_kernel == WLD_PAIR_KERNEL_UMMA_I8;
  // exact-accumulation limit of a Gram entry: fp32 holds integers up to 2^24, s32 up to 2^31-1
  const unsigned long long exact_limit = i8 ? ((1ull << 31) - 1) : (1ull << 24);

  WLD_CUDA(c, c->q.ensure(sizeof(double) * (size_t)c->ldc));
  WLD_CUDA(c, c->limbs.ensure((sizeof(uint16_t) + 1) * 4 * (size_t)c->ldc));  // u16 [4][ldc] then u8 [4][ldc]
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
                                               c->limbs.as<uint8_t>() + sizeof(uint16_t) * 4 * (size_t)c->ldc,
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
    if (worst <= exact_limit) break;
    if (--bits < 1) return c->fail(WLD_ERR_UNSUPPORTED, "n_seqs too large for exact accumulation");
  }
  gm.limb_bits = bits;
  c->weight_sum = 0.0;
  for (int l = 0; l < gm.n_limbs; ++l) c->weight_sum += std::ldexp((double)sums[l], bits * (gm.n_limbs - 1 - l));
  gm.rows_per_site = 2 * gm.n_limbs;
  gm.sites_per_group = 128 / gm.rows_per_site;
  gm.elem_bytes = i8 ? 1 : 2;
  gm.k_padded = round_up(std::max<int64_t>(n, 1), i8 ? 128 : 64);  // one 128-byte swizzle atom per K block
  gm.a_rows = round_up(std::max<int64_t>(2 * L, 1), 128);
  gm.b_groups = std::max<int64_t>((L + gm.sites_per_group - 1) / gm.sites_per_group, 1);
  gm.b_groups = round_up(gm.b_groups, 2);  // an N tile is two groups

  const int64_t kp = gm.k_padded;
  const size_t es = (size_t)gm.elem_bytes;
  WLD_CUDA(c, c->opA.ensure(es * (size_t)gm.a_rows * (size_t)kp));
  WLD_CUDA(c, c->opB.ensure(es * (size_t)gm.b_groups * 128 * (size_t)kp));
  const unsigned kblocks = (unsigned)((kp / (i8 ? 16 : 8) + 255) / 256);  // a thread expands 16 u8 / 8 bf16 elements
  // Only the operand rows this partition's tiles read are expanded (multi-GPU: a contiguous range of the tile
  // list touches a slice of the limb operand and, in the early strips, a prefix of the indicator operand).
  int64_t site_lo = 0, site_hi = gm.a_rows / 2;     // sites whose indicator rows are needed
  int64_t grp_lo = 0, grp_hi = gm.b_groups;         // 128-row groups of the limb operand that are needed
  if (c->pair_kernel != WLD_PAIR_KERNEL_SIMT) {
    const int rc = ensure_tile_plan(c);
    if (rc != WLD_OK) return rc;
    if (c->plan_tiles_n == 0) return WLD_OK;
    const int64_t tile_m = 64 * c->cta_group;
    site_lo = c->plan_x[0] * tile_m;
    site_hi = std::min<int64_t>(site_hi, (c->plan_x[1] + 1) * tile_m);
    grp_lo = c->plan_y[0] * 2;
    grp_hi = std::min<int64_t>(grp_hi, (c->plan_y[1] + 1) * 2);
  } else {
    return WLD_OK;  // the CUDA-core verification kernel reads the code matrix and q directly
  }
  {
    // grid.y limit is 65535: fold larger site counts into several launches
    for (int64_t y0 = site_lo; y0 < site_hi; y0 += 65535) {
      const unsigned ny = (unsigned)std::min<int64_t>(65535, site_hi - y0);
      auto kern = i8 ? expand_a_kernel<true> : expand_a_kernel<false>;
      kern<<<dim3(kblocks, ny), 256, 0, c->stream>>>(
          c->codes.as<uint8_t>() + y0 * c->ldc, c->ldc, std::max<int64_t>(L - y0, 0), c->maj.as<int8_t>() + y0,
          c->mnr.as<int8_t>() + y0, kp, c->opA.as<uint8_t>() + (size_t)(2 * y0 * kp) * es);
      tm.launched();
    }
  }
  {
    const int spg = gm.sites_per_group;
    const int64_t groups_per_launch = 65535 / (spg + 1);
    for (int64_t g0 = grp_lo; g0 < grp_hi; g0 += groups_per_launch) {
      const int64_t ng = std::min<int64_t>(groups_per_launch, grp_hi - g0);
      auto kern = i8 ? expand_b_kernel<true> : expand_b_kernel<false>;
      kern<<<dim3(kblocks, (unsigned)(ng * (spg + 1))), 256, 0, c->stream>>>(
          c->codes.as<uint8_t>() + g0 * spg * c->ldc, c->ldc, std::max<int64_t>(L - g0 * spg, 0),
          c->maj.as<int8_t>() + g0 * spg, c->mnr.as<int8_t>() + g0 * spg, c->limbs.as<uint16_t>(), gm.n_limbs, spg,
          kp, c->opB.as<uint8_t>() + (size_t)(g0 * 128 * kp) * es);
      tm.launched();
    }
  }
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

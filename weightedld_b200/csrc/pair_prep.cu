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

// -------------------------------------------------------------------------------------------------
// Block-exponent fixed point.  With u = w/max(w) in (0, 1]:
//     e = clamp(-exponent(u), 0, G)       (u * 2^e in (1/2, 1] unless the clamp at G bites)
//     m = rint(u * 2^e * (2^B - 1))       B = NL * limb_bits mantissa bits, split into NL limbs
//     q = m * 2^(G - e)                   the integer weight every kernel of the pair stage sums
// The factor g = 2^(G - e) <= 128 ("gain") rides in the INDICATOR operand (opA holds g instead of 1), the limbs
// of m in the limb operand, so the tensor cores still multiply u8 x u8 (or bf16 x bf16) and the Gram entry
// sum g * limb stays an exact integer.  G extra bits of dynamic range cost no tensor work: every weight
// down to 2^-G of the maximum keeps B relative bits (24 at NL = 3: what the reference's f32 carries,
// lib.rs:469-479); below that the relative error grows as 2^-(B+1) / (u * 2^G).
// The kernel also DECIDES the geometry, so that the host needs one read-back per pair stage:
//   NL   = caller's choice, or 3, or 4 when some nonzero weight is below 2^-8 of the maximum;
//   G    = caller's choice, or the smallest value that normalises the smallest nonzero weight (<= 7),
//          lowered while a limb column sum (an upper bound of every Gram entry) exceeds what the
//          accumulator holds exactly (s32: 2^31 - 1, fp32: 2^24) or the sum of all q exceeds 2^53;
//   bits = 8, lowered for the fp32 accumulator of the bf16 kernel if the limit still does not hold at G = 0.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quant_one(float w, double mx, double scale, int G, unsigned long long& m, int& e) {
  const double u = __ddiv_rn((double)w, mx);
  if (!(u > 0.0)) {
    m = 0;
    e = G;
    return;
  }
  int ex;
  (void)frexp(u, &ex);  // u = f * 2^ex, f in [1/2, 1)
  e = min(G, max(0, -ex));
  m = (unsigned long long)rint(__dmul_rn(ldexp(u, e), scale));
}

template <class T, class Op>
__device__ __forceinline__ T block_reduce_1024(T v, T* s_buf, Op op) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();  // s_buf may still be read from an earlier reduction
  if (lane == 0) s_buf[warp] = v;
  __syncthreads();
  T r = s_buf[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = op(r, s_buf[i]);
  return r;
}

__global__ void __launch_bounds__(1024) quantize_kernel(const float* __restrict__ w, int64_t n_seqs, int64_t ldc,
                                                        int nl_opt, int gain_opt, int bits_opt, unsigned long long exact_limit,
                                                        int may_narrow, double* __restrict__ q,
                                                        uint16_t* __restrict__ glimb,
                                                        uint16_t* __restrict__ limbs, uint8_t* __restrict__ limbs8,
                                                        uint8_t* __restrict__ gain8, QuantDecision* __restrict__ out) {
  __shared__ float s_f[32];
  __shared__ unsigned long long s_u[32];
  __shared__ int s_state[4];
  float mx = 0.0f, mn = INFINITY, mnz = INFINITY;
  int bad = 0;
  for (int64_t s = threadIdx.x; s < n_seqs; s += blockDim.x) {
    const float v = w[s];
    if (!(v >= 0.0f) || isinf(v)) bad = 1;
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
    if (v > 0.0f) mnz = fminf(mnz, v);
  }
  mx = block_reduce_1024(mx, s_f, [](float a, float b) { return fmaxf(a, b); });
  mn = block_reduce_1024(mn, s_f, [](float a, float b) { return fminf(a, b); });
  mnz = block_reduce_1024(mnz, s_f, [](float a, float b) { return fminf(a, b); });
  bad = (int)block_reduce_1024((unsigned long long)bad, s_u, [](unsigned long long a, unsigned long long b) { return a | b; });
  if (bad || !(mx > 0.0f)) {
    if (threadIdx.x == 0) {
      QuantDecision d{};
      d.flags = 1;
      *out = d;
    }
    return;
  }
  const double mxd = (double)mx;
  int nl, bits, G, x = 0;
  const bool all_equal = mn == mx;
  if (all_equal) {  // e.g. --unweighted (main.rs:150-153): q == 1 for every sequence, one 0-bit limb
    nl = 1;
    bits = 0;
    G = 0;
  } else {
    int ex;
    (void)frexp(__ddiv_rn((double)mnz, mxd), &ex);
    x = max(0, -ex);
    nl = nl_opt > 0 ? nl_opt : (x > 7 ? 4 : 3);
    bits = bits_opt > 0 ? bits_opt : 8;
    G = gain_opt >= 0 ? min(gain_opt, 7) : min(7, x);
  }
  unsigned long long tot[4];
  int unsupported = 0;
  for (;;) {
    const int B = nl * bits;
    const double scale = B > 0 ? (double)((1ull << B) - 1ull) : 1.0;
    const uint32_t mask = bits > 0 ? (1u << bits) - 1u : 1u;
    unsigned long long sums[4] = {0, 0, 0, 0};
    for (int64_t s = threadIdx.x; s < n_seqs; s += blockDim.x) {
      unsigned long long m;
      int e;
      quant_one(w[s], mxd, scale, G, m, e);
      const unsigned long long g = 1ull << (G - e);
      for (int l = 0; l < nl; ++l) sums[l] += g * ((uint32_t)(m >> (bits * (nl - 1 - l))) & mask);
    }
    double wsum = 0.0;
    unsigned long long worst = 0;
    for (int l = 0; l < 4; ++l) {
      tot[l] = block_reduce_1024(sums[l], s_u, [](unsigned long long a, unsigned long long b) { return a + b; });
      if (l < nl) {
        worst = max(worst, tot[l]);
        wsum += ldexp((double)tot[l], bits * (nl - 1 - l));
      }
    }
    if (threadIdx.x == 0) {
      int st = 0;  // 0: accept
      if (worst > exact_limit || wsum > 9007199254740992.0) {
        if (G > 0) st = 1;                          // one gain bit less
        else if (may_narrow && bits > 1) st = 2;    // narrower limbs (fp32 accumulator of the bf16 kernel)
        else st = 3;                                // cannot be made exact
      }
      s_state[0] = st;
    }
    __syncthreads();
    const int st = s_state[0];
    __syncthreads();
    if (st == 0) break;
    if (st == 1) --G;
    else if (st == 2) --bits;
    else { unsupported = 1; break; }
  }
  const int B = nl * bits;
  const double scale = B > 0 ? (double)((1ull << B) - 1ull) : 1.0;
  const uint32_t mask = bits > 0 ? (1u << bits) - 1u : 1u;
  double err = 0.0;
  unsigned top_min = 0xffffffffu;  // smallest top limb of a nonzero fixed-point weight (the screen's bound, pair_epilogue.cuh)
  for (int64_t s = threadIdx.x; s < ldc; s += blockDim.x) {
    unsigned long long m = 0;
    int e = G;
    unsigned g = 0;
    if (s < n_seqs) {
      quant_one(w[s], mxd, scale, G, m, e);
      g = 1u << (G - e);
      const double u = __ddiv_rn((double)w[s], mxd);
      if (u > 0.0) err = fmax(err, fabs(ldexp((double)m, -e) / scale - u) / u);
    }
    q[s] = (double)(m * g);
    gain8[s] = (uint8_t)g;
    if (m > 0) top_min = min(top_min, (unsigned)(m >> (bits * (nl - 1))) & mask);
    for (int l = 0; l < nl; ++l) {
      const uint32_t v = (uint32_t)(m >> (bits * (nl - 1 - l))) & mask;
      limbs[(int64_t)l * ldc + s] = (uint16_t)v;  // raw limb value 0..255; converted at expansion
      limbs8[(int64_t)l * ldc + s] = (uint8_t)v;  // the same as bytes for the u8 operands (SWAR expansion)
      glimb[(int64_t)l * ldc + s] = (uint16_t)(v * g);  // gain x limb < 2^15: what pair_refine.cu sums (dp2a)
    }
  }
  const unsigned long long err_bits = block_reduce_1024((unsigned long long)__double_as_longlong(err), s_u,
                                                        [](unsigned long long a, unsigned long long b) { return max(a, b); });
  top_min = (unsigned)block_reduce_1024((unsigned long long)top_min, s_u,
                                        [](unsigned long long a, unsigned long long b) { return min(a, b); });
  if (threadIdx.x == 0) {
    QuantDecision d{};
    d.top_min = top_min == 0xffffffffu ? 0 : (int)top_min;
    {  // kappa of ld_screen_f32: (1 + 1/top_min)^2 - 1 + 1e-5, rounded up
      const double eta = 1.0 / (double)max(d.top_min, 1);
      float k = (float)((1.0 + eta) * (1.0 + eta) - 1.0 + 1e-5);
      if ((double)k < (1.0 + eta) * (1.0 + eta) - 1.0 + 1e-5) k = __uint_as_float(__float_as_uint(k) + 1u);
      d.kappa = k;
    }
    d.flags = (all_equal ? 2 : 0) | (unsupported ? 4 : 0);
    d.n_limbs = nl;
    d.limb_bits = bits;
    d.gain_bits = G;
    d.span_log2 = x;
    d.weight_sum = 0.0;
    for (int l = 0; l < nl; ++l) {
      d.limb_sums[l] = tot[l];
      d.weight_sum += ldexp((double)tot[l], bits * (nl - 1 - l));
    }
    d.rel_err = __longlong_as_double((long long)err_bits);
    *out = d;
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
// Indicator rows: the element is the sequence's gain g = 2^(G - e) (1 when no gain bits are in use), so whole
// words can be built with SWAR compares: (code == sym ? 0xff : 0) & gain byte.
template <bool kI8>
__device__ __forceinline__ void store_indicator(void* dst, const uint32_t* cw, int sym, const uint32_t* gw) {
  if (kI8) {
    const uint32_t rep = sym < 0 ? 0xffffffffu : (uint32_t)sym * 0x01010101u;  // 0xff never matches a code
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(__vcmpeq4(cw[0], rep) & gw[0], __vcmpeq4(cw[1], rep) & gw[1],
                   __vcmpeq4(cw[2], rep) & gw[2], __vcmpeq4(cw[3], rep) & gw[3]);
  } else {
    uint32_t g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = elem_of<false>((gw[k >> 2] >> (8 * (k & 3))) & 0xffu);  // powers of two: exact in bf16
    store_select<false>(dst, cw, sym, g);
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
                                                       const uint8_t* __restrict__ gain8, int64_t kp,
                                                       uint8_t* __restrict__ opA) {
  constexpr int ES = kI8 ? 1 : 2;
  constexpr int kSeq = Expand<kI8>::kSeq;
  const int64_t i = blockIdx.y;
  const int64_t s0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kSeq;
  if (s0 >= kp) return;
  uint8_t* r0 = opA + ((2 * i) * kp + s0) * ES;
  uint8_t* r1 = opA + ((2 * i + 1) * kp + s0) * ES;
  if (i < n_kept) {
    uint32_t cw[Expand<kI8>::kWords], gw[Expand<kI8>::kWords];
    load_codes<kI8>(codes + i * ldc + s0, cw);
    load_codes<kI8>(gain8 + s0, gw);       // gains of the same kSeq sequences (0 in the K padding)
    store_indicator<kI8>(r0, cw, maj[i], gw);  // maj/min == -1 never matches a code
    store_indicator<kI8>(r1, cw, mnr[i], gw);
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

// Screen operands in one pass over the code matrix (u8 only).  The one-limb layout of the limb operand puts site j
// at rows 2j, 2j+1 — the same indexing as the indicator operand — so a thread reads the 16 codes of (site, sequence
// block) once and writes four 16-byte row segments: opA[2i + alpha] = indicator x gain (sites in [a_lo, a_hi)) and
// opB1[2i + beta] = indicator x TOP limb (sites in [b_lo, b_hi)).  Sites >= n_kept are zero rows.
__global__ void __launch_bounds__(256) expand_screen_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int64_t n_kept,
                                                            const int8_t* __restrict__ maj, const int8_t* __restrict__ mnr,
                                                            const uint8_t* __restrict__ gain8, const uint8_t* __restrict__ top8,
                                                            int64_t kp, int64_t site0, int64_t a_lo, int64_t a_hi, int64_t b_lo,
                                                            int64_t b_hi, uint8_t* __restrict__ opA, uint8_t* __restrict__ opB1) {
  const int64_t i = site0 + blockIdx.y;
  const int64_t s0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
  if (s0 >= kp) return;
  const bool do_a = i >= a_lo && i < a_hi, do_b = i >= b_lo && i < b_hi;
  uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
  if (i < n_kept) {
    const uint4 cv = __ldg(reinterpret_cast<const uint4*>(codes + i * ldc + s0));
    const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w};
    const int sm = maj[i], sn = mnr[i];
    const uint32_t rm = sm < 0 ? 0xffffffffu : (uint32_t)sm * 0x01010101u;  // 0xff never matches a code
    const uint32_t rn = sn < 0 ? 0xffffffffu : (uint32_t)sn * 0x01010101u;
    uint32_t mm[4], mn[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      mm[w] = __vcmpeq4(cw[w], rm);
      mn[w] = __vcmpeq4(cw[w], rn);
    }
    if (do_a) {
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(gain8 + s0));  // 0 in the K padding
      a0 = make_uint4(g.x & mm[0], g.y & mm[1], g.z & mm[2], g.w & mm[3]);
      a1 = make_uint4(g.x & mn[0], g.y & mn[1], g.z & mn[2], g.w & mn[3]);
    }
    if (do_b) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(top8 + s0));
      b0 = make_uint4(t.x & mm[0], t.y & mm[1], t.z & mm[2], t.w & mm[3]);
      b1 = make_uint4(t.x & mn[0], t.y & mn[1], t.z & mn[2], t.w & mn[3]);
    }
  }
  if (do_a) {
    *reinterpret_cast<uint4*>(opA + (2 * i) * kp + s0) = a0;
    *reinterpret_cast<uint4*>(opA + (2 * i + 1) * kp + s0) = a1;
  }
  if (do_b) {
    *reinterpret_cast<uint4*>(opB1 + (2 * i) * kp + s0) = b0;
    *reinterpret_cast<uint4*>(opB1 + (2 * i + 1) * kp + s0) = b1;
  }
}

}  // namespace

// Reads the quantiser's decision back (one synchronisation) and fixes the geometry of the exact kernel.
int finish_quant(wld_ctx* c) {
  PairGeom& gm = c->geom;
  WLD_CUDA(c, cudaMemcpyAsync(c->quant_host, c->quant.p, sizeof(QuantDecision), cudaMemcpyDeviceToHost, c->stream));
  // (the sampling launch's two counters ride along: counters[8] candidates, counters[9] pairs)
  WLD_CUDA(c, cudaMemcpyAsync(c->sample_host, c->counters.as<unsigned long long>() + 8, 3 * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  const QuantDecision qd = *c->quant_host;
  if (qd.flags & 1) return c->fail(WLD_ERR_INVALID, "weights must be finite, >= 0 and not all zero");
  if (qd.flags & 4) return c->fail(WLD_ERR_UNSUPPORTED, "n_seqs too large for exact accumulation");
  gm.n_limbs = qd.n_limbs;
  gm.limb_bits = qd.limb_bits;
  gm.gain_bits = qd.gain_bits;
  c->weight_sum = qd.weight_sum;
  c->quant_span_log2 = qd.span_log2;
  c->quant_rel_err = qd.rel_err;
  c->quant_top_min = qd.top_min;
  gm.rows_per_site = 2 * gm.n_limbs;
  gm.sites_per_group = 128 / gm.rows_per_site;
  gm.b_groups = std::max<int64_t>((c->n_kept + gm.sites_per_group - 1) / gm.sites_per_group, 1);
  gm.b_groups = round_up(gm.b_groups, 2);  // an N tile is two groups
  return WLD_OK;
}

// Quantises the weights and expands the indicator operand.  try_screen: also the one-limb operand of the screen,
// all BEFORE the quantiser's decision is read back (neither depends on it), so that the sampling launch can
// follow on the stream and one synchronisation (finish_quant) returns both results.
int run_pair_prep(wld_ctx* c, ScopedStageTimer& tm, bool try_screen) {
  PairGeom& gm = c->geom;
  const int64_t n = c->n_seqs, L = c->n_kept;
  const bool i8 = c->pair_kernel == WLD_PAIR_KERNEL_UMMA_I8;
  // exact-accumulation limit of a Gram entry: fp32 holds integers up to 2^24, s32 up to 2^31-1; the FP64
  // verification kernel only needs the sum of all q below 2^53 (checked by the quantiser for every kernel)
  const unsigned long long exact_limit = c->pair_kernel == WLD_PAIR_KERNEL_SIMT ? (1ull << 53)
                                         : i8                                   ? ((1ull << 31) - 1)
                                                                                : (1ull << 24);

  WLD_CUDA(c, c->q.ensure(sizeof(double) * (size_t)c->ldc));
  WLD_CUDA(c, c->glimb.ensure(sizeof(uint16_t) * 4 * (size_t)c->ldc));
  WLD_CUDA(c, c->limbs.ensure((sizeof(uint16_t) + 1) * 4 * (size_t)c->ldc));  // u16 [4][ldc] then u8 [4][ldc]
  WLD_CUDA(c, c->gain8.ensure((size_t)c->ldc));
  WLD_CUDA(c, c->quant.ensure(sizeof(QuantDecision)));
  WLD_CUDA(c, c->counters.ensure(sizeof(unsigned long long) * 16));
  WLD_CUDA(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 16, c->stream));
  if (!c->quant_host) {
    WLD_CUDA(c, cudaMallocHost(&c->quant_host, sizeof(QuantDecision) + 3 * sizeof(unsigned long long)));
    c->sample_host = reinterpret_cast<unsigned long long*>(c->quant_host + 1);
  }
  c->sample_host[0] = c->sample_host[1] = c->sample_host[2] = 0;

  // One launch quantises the weights AND decides limbs / limb width / gain bits; one read-back tells the host.
  quantize_kernel<<<1, 1024, 0, c->stream>>>(c->w32.as<float>(), n, c->ldc, c->n_limbs_opt, c->gain_opt, c->limb_bits_opt, exact_limit,
                                             c->pair_kernel == WLD_PAIR_KERNEL_UMMA ? 1 : 0, c->q.as<double>(),
                                             c->glimb.as<uint16_t>(), c->limbs.as<uint16_t>(),
                                             c->limbs.as<uint8_t>() + sizeof(uint16_t) * 4 * (size_t)c->ldc,
                                             c->gain8.as<uint8_t>(), c->quant.as<QuantDecision>());
  tm.launched();
  WLD_CUDA(c, cudaGetLastError());
  gm.elem_bytes = i8 ? 1 : 2;
  gm.k_padded = round_up(std::max<int64_t>(n, 1), i8 ? 128 : 64);  // one 128-byte swizzle atom per K block
  gm.a_rows = round_up(std::max<int64_t>(2 * L, 1), 128);
  if (!try_screen) {
    const int rc = finish_quant(c);
    if (rc != WLD_OK) return rc;
  }
  if (c->pair_kernel == WLD_PAIR_KERNEL_SIMT) return WLD_OK;  // the CUDA-core verification kernel reads the code matrix and q directly

  const int64_t kp = gm.k_padded;
  const size_t es = (size_t)gm.elem_bytes;
  WLD_CUDA(c, c->opA.ensure(es * (size_t)gm.a_rows * (size_t)kp));
  const unsigned kblocks = (unsigned)((kp / (i8 ? 16 : 8) + 255) / 256);  // a thread expands 16 u8 / 8 bf16 elements
  // Only the operand rows this partition's tiles read are expanded (multi-GPU: a contiguous range of the cell
  // list touches a slice of the limb operand and, in the early strips, a prefix of the indicator operand).  The
  // rows of the indicator operand are the same for the screen's and the exact kernel's schedule.
  const int which = try_screen ? 1 : 0;
  {
    const int rc = ensure_tile_plan(c, which);
    if (rc != WLD_OK) return rc;
  }
  const DevPlan& dp = c->plans[which];
  if (dp.n_tiles == 0) return WLD_OK;
  const int64_t tile_m = 64 * c->cta_group;
  const int64_t site_lo = dp.x[0] * tile_m, site_hi = std::min<int64_t>(gm.a_rows / 2, (dp.x[1] + 1) * tile_m);
  if (try_screen) {
    // both screen operands from one read of the code matrix (row 2i + {0, 1} of either holds site i)
    const int64_t b_rows = round_up(std::max<int64_t>((L + 63) / 64, 1), 2) * 128;  // an N tile is two 128-row groups
    WLD_CUDA(c, c->opB1.ensure((size_t)b_rows * (size_t)kp));
    const int64_t b_lo = dp.y[0] * 128, b_hi = std::min<int64_t>(b_rows / 2, (dp.y[1] + 1) * 128);  // sites of the N tiles in use
    const int64_t lo = std::min(site_lo, b_lo), hi = std::max(site_hi, b_hi);
    const uint8_t* top8 = c->limbs.as<uint8_t>() + sizeof(uint16_t) * 4 * (size_t)c->ldc;  // limb 0 = most significant
    for (int64_t y0 = lo; y0 < hi; y0 += 65535) {
      const unsigned ny = (unsigned)std::min<int64_t>(65535, hi - y0);
      expand_screen_kernel<<<dim3(kblocks, ny), 256, 0, c->stream>>>(
          c->codes.as<uint8_t>(), c->ldc, L, c->maj.as<int8_t>(), c->mnr.as<int8_t>(), c->gain8.as<uint8_t>(), top8, kp, y0,
          site_lo, site_hi, b_lo, b_hi, c->opA.as<uint8_t>(), c->opB1.as<uint8_t>());
      tm.launched();
    }
    WLD_CUDA(c, cudaGetLastError());
    return WLD_OK;
  }
  // grid.y limit is 65535: fold larger site counts into several launches
  for (int64_t y0 = site_lo; y0 < site_hi; y0 += 65535) {
    const unsigned ny = (unsigned)std::min<int64_t>(65535, site_hi - y0);
    auto kern = i8 ? expand_a_kernel<true> : expand_a_kernel<false>;
    kern<<<dim3(kblocks, ny), 256, 0, c->stream>>>(
        c->codes.as<uint8_t>() + y0 * c->ldc, c->ldc, std::max<int64_t>(L - y0, 0), c->maj.as<int8_t>() + y0,
        c->mnr.as<int8_t>() + y0, c->gain8.as<uint8_t>(), kp, c->opA.as<uint8_t>() + (size_t)(2 * y0 * kp) * es);
    tm.launched();
  }
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

// The limb operand: every limb of the exact kernel's layout (screen = false; needs finish_quant), or the TOP limb
// alone in the one-limb layout (64 sites per 128-row group) for the screen.
int run_expand_limbs(wld_ctx* c, ScopedStageTimer& tm, bool screen) {
  const PairGeom& gm = c->geom;
  const int64_t L = c->n_kept, kp = gm.k_padded;
  const bool i8 = gm.elem_bytes == 1;
  const size_t es = (size_t)gm.elem_bytes;
  const int which = screen ? 1 : 0;
  {
    const int rc = ensure_tile_plan(c, which);
    if (rc != WLD_OK) return rc;
  }
  const DevPlan& dp = c->plans[which];
  if (dp.n_tiles == 0) return WLD_OK;
  const int n_limbs = screen ? 1 : gm.n_limbs;
  const int spg = 128 / (2 * n_limbs);
  const int64_t b_groups = screen ? round_up(std::max<int64_t>((L + 63) / 64, 1), 2) : gm.b_groups;
  DevBuf& op = screen ? c->opB1 : c->opB;
  WLD_CUDA(c, op.ensure(es * (size_t)b_groups * 128 * (size_t)kp));
  const unsigned kblocks = (unsigned)((kp / (i8 ? 16 : 8) + 255) / 256);
  const int64_t grp_lo = dp.y[0] * 2, grp_hi = std::min<int64_t>(b_groups, (dp.y[1] + 1) * 2);
  const int64_t groups_per_launch = 65535 / (spg + 1);
  for (int64_t g0 = grp_lo; g0 < grp_hi; g0 += groups_per_launch) {
    const int64_t ng = std::min<int64_t>(groups_per_launch, grp_hi - g0);
    auto kern = i8 ? expand_b_kernel<true> : expand_b_kernel<false>;
    kern<<<dim3(kblocks, (unsigned)(ng * (spg + 1))), 256, 0, c->stream>>>(
        c->codes.as<uint8_t>() + g0 * spg * c->ldc, c->ldc, std::max<int64_t>(L - g0 * spg, 0),
        c->maj.as<int8_t>() + g0 * spg, c->mnr.as<int8_t>() + g0 * spg, c->limbs.as<uint16_t>(), n_limbs, spg,
        kp, op.as<uint8_t>() + (size_t)(g0 * 128 * kp) * es);
    tm.launched();
  }
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

// encode_filter.cu — stage 1: alphabet encoding, per-site histograms, site filter, and the
// transposing gather that builds the kept, site-major 0..5 code matrix.
//
// Reference semantics restated on the GPU (bit-exact, integer + three f32 operations):
//   Symbol::from(char)                       lib.rs:53-64
//   SymbolHistogram::from_slice / acgt       lib.rs:98-109
//   major_minor_symbols                      lib.rs:126-140
//   is_site_of_interest + threshold          lib.rs:310-338, main.rs:139
//   SiteSet::from_multiseq / filter_by       lib.rs:176-206, 230-251
//   compute_variable_sites (Python dialect)  WeightedLD.py:44-98
//
// All three kernels are HBM-bound byte kernels.  Algorithmic bytes: n_seqs*n_cols read for the
// histogram; n_seqs*n_cols read + n_seqs*n_kept written for the gather (tiles without a kept
// column are skipped, so the second read shrinks with the keep ratio).
#include "common.cuh"

namespace wld {
namespace {

// ---------------------------------------------------------------------------------------------
// SWAR alphabet: four bytes at a time.  lib.rs:53-64: aA->0 cC->1 gG->2 tT->3 '-'->4 else->5.
// Lower-casing with |0x20 is only applied for the letter tests ('\r'|0x20 == '-', so '-' is
// tested on the raw byte).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t encode4_ascii(uint32_t x) {
  const uint32_t y = x | 0x20202020u;
  uint32_t code = 0x05050505u;
  code -= __vcmpeq4(y, 0x61616161u) & 0x05050505u;  // a -> 0
  code -= __vcmpeq4(y, 0x63636363u) & 0x04040404u;  // c -> 1
  code -= __vcmpeq4(y, 0x67676767u) & 0x03030303u;  // g -> 2
  code -= __vcmpeq4(y, 0x74747474u) & 0x02020202u;  // t -> 3
  code -= __vcmpeq4(x, 0x2d2d2d2du) & 0x01010101u;  // - -> 4
  return code;
}
// Already-encoded input: values above 5 read as 5 (Unknown).
__device__ __forceinline__ uint32_t encode4_codes(uint32_t x) { return __vminu4(x, 0x05050505u); }

__device__ __forceinline__ uint32_t encode1(uint32_t c, bool ascii) {
  return (ascii ? encode4_ascii(c) : encode4_codes(c)) & 0xffu;
}

// ---------------------------------------------------------------------------------------------
// Kernel 1a: per-column histogram, vectorised path (base and pitch 16-byte aligned).
// A warp owns 512 consecutive columns (16 per lane, one 16-byte load per row); the 8 warps of a
// block take interleaved rows of the block's row chunk.  Classification is ONE shared-memory lookup per
// byte: a 256-entry table maps the byte to 1 << (6 * code) (0 for Unknown), so adding the looked-up words
// accumulates all five counts of a column at once in 6-bit fields of one register — about 4 instructions
// per byte instead of ~10 for five SWAR compares, which had this kernel ALU-bound at a quarter of HBM
// bandwidth.  Fields are flushed into the block's shared histogram (bank-conflict-free layout) before
// any can exceed 63; one global atomicAdd per (symbol, column) per block follows.
// DNA bytes fall into different banks of the table (A C G T - N: 1 3 7 20 13 14), so the lookups are
// conflict-free on real data; adversarial bytes only cost replays.
// ---------------------------------------------------------------------------------------------
constexpr int kHistThreads = 256;
constexpr int kHistColsPerBlock = 512;
constexpr int kHistFlushRows = 63;   // 6-bit count fields
constexpr int kHistUnroll = 4;       // rows in flight per warp

__device__ __forceinline__ uint32_t hist_lut_entry(uint32_t byte, bool ascii) {
  const uint32_t code = encode1(byte, ascii);
  return code < 5u ? (1u << (6u * code)) : 0u;
}

template <bool kAscii>
__global__ void __launch_bounds__(kHistThreads) hist_vec16_kernel(const uint8_t* __restrict__ raw,
                                                                  int64_t n_seqs, int64_t n_cols,
                                                                  int64_t row_stride, int rows_per_block,
                                                                  uint32_t* __restrict__ hist,
                                                                  int64_t cols_padded) {
  __shared__ uint32_t s_hist[5][kHistColsPerBlock];  // index j*32 + lane  <->  column lane*16 + j
  __shared__ uint32_t s_lut[256];
  for (int i = threadIdx.x; i < 5 * kHistColsPerBlock; i += kHistThreads) (&s_hist[0][0])[i] = 0;
  s_lut[threadIdx.x] = hist_lut_entry(threadIdx.x, kAscii);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t col0 = (int64_t)blockIdx.x * kHistColsPerBlock + lane * 16;
  const int64_t row_begin = (int64_t)blockIdx.y * rows_per_block;
  const int64_t row_end = min(row_begin + rows_per_block, n_seqs);
  const bool active = col0 < cols_padded;  // cols_padded is a multiple of 16 and <= row_stride

  uint32_t acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0;
  int since_flush = 0;

  auto flush = [&]() {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const uint32_t f = (acc[j] >> (6 * k)) & 63u;
        if (f) atomicAdd(&s_hist[k][j * 32 + lane], f);  // most columns hold two or three symbols: skip empty fields
      }
      acc[j] = 0;
    }
    since_flush = 0;
  };
  auto count16 = [&](const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[4 * q + b] += s_lut[(w[q] >> (8 * b)) & 0xffu];
  };

  if (active) {
    constexpr int kStep = kHistThreads / 32;  // rows between two loads of a warp
    const uint8_t* p = raw + col0;
    int64_t r = row_begin + warp;
    for (; r + (kHistUnroll - 1) * kStep < row_end; r += kHistUnroll * kStep) {
      uint4 v[kHistUnroll];
#pragma unroll
      for (int u = 0; u < kHistUnroll; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(p + (r + u * kStep) * row_stride));
#pragma unroll
      for (int u = 0; u < kHistUnroll; ++u) count16(v[u]);
      since_flush += kHistUnroll;
      if (since_flush > kHistFlushRows - kHistUnroll) flush();
    }
    for (; r < row_end; r += kStep) {
      count16(__ldg(reinterpret_cast<const uint4*>(p + r * row_stride)));
      if (++since_flush == kHistFlushRows) flush();
    }
    flush();
  }
  __syncthreads();
  const int64_t cb = (int64_t)blockIdx.x * kHistColsPerBlock;
  for (int i = threadIdx.x; i < 5 * kHistColsPerBlock; i += kHistThreads) {
    const int k = i / kHistColsPerBlock, c = i % kHistColsPerBlock;  // c = column within the block
    const uint32_t v = s_hist[k][(c & 15) * 32 + (c >> 4)];
    if (v && cb + c < cols_padded) atomicAdd(&hist[(int64_t)k * cols_padded + cb + c], v);
  }
}

// Kernel 1a', generic path for unaligned input: one column per thread, byte loads.
__global__ void __launch_bounds__(256) hist_generic_kernel(const uint8_t* __restrict__ raw, int64_t row0, int64_t n_seqs,
                                                           int64_t n_cols, int64_t row_stride,
                                                           int rows_per_block, bool ascii,
                                                           uint32_t* __restrict__ hist, int64_t cols_padded) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n_cols) return;
  const int64_t row_begin = row0 + (int64_t)blockIdx.y * rows_per_block;
  const int64_t row_end = min(row_begin + rows_per_block, n_seqs);
  uint32_t h[5] = {0, 0, 0, 0, 0};
  for (int64_t r = row_begin; r < row_end; ++r) {
    const uint32_t c = encode1(raw[r * row_stride + col], ascii);
#pragma unroll
    for (int k = 0; k < 5; ++k) h[k] += (c == (uint32_t)k);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (h[k]) atomicAdd(&hist[(int64_t)k * cols_padded + col], h[k]);
}

// ---------------------------------------------------------------------------------------------
// Kernel 1b: per-site decision.  Completes bin 5 (Unknown = n_seqs - sum of the others), runs the
// major/minor scan of lib.rs:126-140 and the filter of lib.rs:310-338.
// ---------------------------------------------------------------------------------------------
__global__ void decide_kernel(uint32_t* __restrict__ hist, int64_t cols_padded, int64_t n_cols,
                              int64_t n_seqs, int mode, unsigned long long min_acgt, float min_minor,
                              float max_minor, double py_min_acgt, double py_min_variability,
                              uint8_t* __restrict__ keep, int8_t* __restrict__ maj_raw,
                              int8_t* __restrict__ min_raw) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= cols_padded) return;
  if (col >= n_cols) {
    keep[col] = 0;
    maj_raw[col] = -1;
    min_raw[col] = -1;
    return;
  }
  uint32_t h[5];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    h[k] = hist[(int64_t)k * cols_padded + col];
    sum += h[k];
  }
  hist[5 * cols_padded + col] = (uint32_t)n_seqs - sum;

  int maj = -1, mnr = -1;  // lib.rs:126-140
#pragma unroll
  for (int sym = 0; sym < 5; ++sym) {
    const uint32_t majc = maj >= 0 ? h[maj] : 0u;
    const uint32_t minc = mnr >= 0 ? h[mnr] : 0u;
    if (h[sym] > majc) {
      mnr = maj;
      maj = sym;
    } else if (h[sym] > minc) {
      mnr = sym;
    }
  }
  maj_raw[col] = (int8_t)maj;
  min_raw[col] = (int8_t)mnr;

  bool k = true;
  if (mode == 2) {
    // WeightedLD.py:64-95 (return_ld_varsites), all in f64: concrete fraction acgt/n > min_acgt, and
    // (everything that is not the major symbol, gaps included) / (known symbols) >= min_variability.
    const double acgt = (double)((unsigned long long)h[0] + h[1] + h[2] + h[3]);
    const bool sufficient = __ddiv_rn(acgt, (double)n_seqs) > py_min_acgt;          // WeightedLD.py:65-68
    const uint32_t major = maj >= 0 ? h[maj] : 0u;                                   // WeightedLD.py:76
    const uint32_t minor = sum - major;                                              // WeightedLD.py:77
    const double frac = minor > 0u ? __ddiv_rn((double)minor, (double)sum) : 0.0;    // WeightedLD.py:80-84
    k = sufficient && frac >= py_min_variability;                                    // WeightedLD.py:87-95
  } else if (mode == 0) {
    const unsigned long long acgt = (unsigned long long)h[0] + h[1] + h[2] + h[3];  // lib.rs:106-109
    if (acgt <= min_acgt) {                                                        // lib.rs:315
      k = false;
    } else if (maj < 0 || mnr < 0) {                                               // lib.rs:319-322
      k = false;
    } else {
      const float maj_count = (float)h[maj];                                       // lib.rs:324
      const float min_count = (float)h[mnr];                                       // lib.rs:325
      const float minor_frac = __fdiv_rn(min_count, __fadd_rn(min_count, maj_count));  // lib.rs:328
      if (minor_frac < min_minor || minor_frac > max_minor) k = false;             // lib.rs:331
    }
  }
  keep[col] = k ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Kernel 1c: ordered compaction of the keep flags (single block; L_raw is at most a few million).
// rank[col] = number of kept columns before col (rank[cols_padded] = total), site_map, and the
// kept sites' major/minor symbols.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
// One block: every thread counts a contiguous segment of the flags, the 1024 segment sums are scanned once,
// and the segment is walked again to write the ranks (two passes, three barriers in total).
__global__ void __launch_bounds__(kScanThreads) scan_kernel(const uint8_t* __restrict__ keep,
                                                            int64_t cols_padded, int32_t* __restrict__ rank,
                                                            int32_t* __restrict__ kept_count) {
  __shared__ int32_t s_warp[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // segments are multiples of 16 flags (cols_padded is one): 16-byte loads, flags are 0/1 so popcount sums them
  const int64_t seg = ((cols_padded / 16 + kScanThreads - 1) / kScanThreads) * 16;
  const int64_t lo = min((int64_t)threadIdx.x * seg, cols_padded), hi = min(lo + seg, cols_padded);
  int sum = 0;
  for (int64_t c = lo; c < hi; c += 16) {
    const uint4 v = *reinterpret_cast<const uint4*>(keep + c);
    sum += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    s_warp[lane] = v;  // inclusive over warps
  }
  __syncthreads();
  int run = (warp ? s_warp[warp - 1] : 0) + incl - sum;  // exclusive prefix of this thread's segment
  for (int64_t c = lo; c < hi; c += 16) {
    const uint4 v = *reinterpret_cast<const uint4*>(keep + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int4 r;
      r.x = run;
      r.y = r.x + (int)(w[q] & 1u);
      r.z = r.y + (int)((w[q] >> 8) & 1u);
      r.w = r.z + (int)((w[q] >> 16) & 1u);
      run = r.w + (int)((w[q] >> 24) & 1u);
      *reinterpret_cast<int4*>(rank + c + 4 * q) = r;
    }
  }
  if (threadIdx.x == kScanThreads - 1) {
    const int total = s_warp[kScanThreads / 32 - 1];
    rank[cols_padded] = total;
    *kept_count = total;
  }
}

__global__ void site_map_kernel(const uint8_t* __restrict__ keep, const int32_t* __restrict__ rank,
                                const int8_t* __restrict__ maj_raw, const int8_t* __restrict__ min_raw,
                                int64_t n_cols, int32_t* __restrict__ site_map, int8_t* __restrict__ maj,
                                int8_t* __restrict__ mnr) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n_cols || !keep[col]) return;
  const int32_t k = rank[col];
  site_map[k] = (int32_t)col;
  maj[k] = maj_raw[col];
  mnr[k] = min_raw[col];
}

// ---------------------------------------------------------------------------------------------
// Kernel 1d: encode + transpose + gather of the kept columns.
// Tile = 128 sequences x 128 raw columns, 256 threads.  Thread (strip = t % 8, row group = t / 8) loads 4
// rows x 16 columns with four 16-byte loads (a warp request = 128 contiguous bytes of 4 rows each), transposes
// four 4x4 blocks of RAW bytes in registers with PRMT and stores one 16-byte word group per block:
// tile[row group][strip*20 + column%16] holds, per column, the 4 sequences of that row group.  The 4 pad words
// per strip and the 164-word pitch make both the 16-byte stores and the 16-byte loads of the second phase
// bank-conflict free.  Second phase: a warp takes 4 columns at a time, lane = row group, one 16-byte load, and
// only KEPT columns are encoded (one shared-memory table lookup per byte — the cost scales with the keep ratio,
// 1/3 at config 4) and written as 128-byte rows of the site-major code matrix.  Rows beyond n_seqs are padded with
// 0xff, which encodes to 5 (Unknown), so the pair operands see zeros there.  Interior tiles take a path
// without any per-element bounds checks.
// ---------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kTilePitch = 164;  // words: 8 strips x (16 columns + 4 pad), + 4 so that the pitch is 4 mod 32

template <bool kAscii>
__global__ void __launch_bounds__(kGatherThreads) gather_kernel(const uint8_t* __restrict__ raw, int64_t n_seqs,
                                                                int64_t n_cols, int64_t row_stride, bool aligned16,
                                                                const uint8_t* __restrict__ keep,
                                                                const int32_t* __restrict__ rank,
                                                                uint8_t* __restrict__ codes, int64_t ldc) {
  __shared__ __align__(16) uint32_t tile[32][kTilePitch];
  __shared__ uint32_t s_code[256];  // byte -> symbol code (lib.rs:53-64, or min(byte, 5) for pre-encoded input)
  const int64_t col_tile = (int64_t)blockIdx.x * 128;
  const int64_t seq_tile = (int64_t)blockIdx.y * 128;
  const int64_t col_hi = min(col_tile + 128, n_cols);
  // rank is an exclusive prefix, rank[x] for x in [0, cols_padded]; cols_padded >= n_cols.
  if (rank[col_hi] == rank[col_tile]) return;  // nothing kept in this column tile
  s_code[threadIdx.x] = encode1(threadIdx.x, kAscii);

  const int strip = threadIdx.x & 7, rg = threadIdx.x >> 3;
  const int64_t col = col_tile + 16 * strip;
  const int64_t seq0 = seq_tile + 4 * rg;
  uint4 r[4];
  // interior tile: bytes between n_cols and the pitch are junk, but those columns are never kept; the tile that
  // holds the LAST row only qualifies when it lies inside n_cols (a borrowed buffer ends at that row's n_cols)
  if (aligned16 && seq_tile + 128 <= n_seqs && col_tile + 128 <= row_stride &&
      (col_tile + 128 <= n_cols || seq_tile + 128 < n_seqs)) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) r[rr] = __ldg(reinterpret_cast<const uint4*>(raw + (seq0 + rr) * row_stride + col));
  } else {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};  // pad: encodes to 5 in both input modes
      if (seq0 + rr < n_seqs) {
        const uint8_t* p = raw + (seq0 + rr) * row_stride + col;
#pragma unroll
        for (int b = 0; b < 16; ++b)
          if (col + b < n_cols) w[b >> 2] = (w[b >> 2] & ~(0xffu << (8 * (b & 3)))) | ((uint32_t)p[b] << (8 * (b & 3)));
      }
      r[rr] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  const uint32_t a[4] = {r[0].x, r[0].y, r[0].z, r[0].w}, b4[4] = {r[1].x, r[1].y, r[1].z, r[1].w};
  const uint32_t c4[4] = {r[2].x, r[2].y, r[2].z, r[2].w}, d4[4] = {r[3].x, r[3].y, r[3].z, r[3].w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // 4x4 byte transpose: rows (a, b, c, d) x columns 4q..4q+3
    const uint32_t t0 = __byte_perm(a[q], b4[q], 0x5140), t1 = __byte_perm(a[q], b4[q], 0x7362);
    const uint32_t t2 = __byte_perm(c4[q], d4[q], 0x5140), t3 = __byte_perm(c4[q], d4[q], 0x7362);
    *reinterpret_cast<uint4*>(&tile[rg][strip * 20 + 4 * q]) =
        make_uint4(__byte_perm(t0, t2, 0x5410), __byte_perm(t0, t2, 0x7632), __byte_perm(t1, t3, 0x5410),
                   __byte_perm(t1, t3, 0x7632));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int quad = warp; quad < 32; quad += kGatherThreads / 32) {
    const int64_t c0 = col_tile + 4 * quad;
    if (c0 >= n_cols) break;
    const uint32_t k4 = *reinterpret_cast<const uint32_t*>(keep + c0);  // keep[] is padded and zero beyond n_cols
    if (k4 == 0u) continue;
    const uint4 v = *reinterpret_cast<const uint4*>(&tile[lane][(quad >> 2) * 20 + 4 * (quad & 3)]);
    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((k4 >> (8 * j)) & 0xffu) {
        const int64_t k = rank[c0 + j];
        // one table lookup per byte: a third of the instructions of five SWAR compares
        const uint32_t e0 = s_code[x[j] & 0xffu], e1 = s_code[(x[j] >> 8) & 0xffu];
        const uint32_t e2 = s_code[(x[j] >> 16) & 0xffu], e3 = s_code[x[j] >> 24];
        *reinterpret_cast<uint32_t*>(codes + k * ldc + seq_tile + 4 * lane) =
            __byte_perm(__byte_perm(e0, e1, 0x0040), __byte_perm(e2, e3, 0x0040), 0x5410);
      }
  }
}

}  // namespace

// =============================================================================================
// host launchers
// =============================================================================================
int run_histogram(wld_ctx* c, ScopedStageTimer& tm) {
  const bool ascii = !(c->input_flags & WLD_INPUT_CODES);
  c->cols_padded = round_up(c->n_cols, 16);
  WLD_CUDA(c, c->hist.ensure(sizeof(uint32_t) * 6 * (size_t)c->cols_padded));
  WLD_CUDA(c, cudaMemsetAsync(c->hist.p, 0, sizeof(uint32_t) * 6 * (size_t)c->cols_padded, c->stream));
  // rows this context counts: all of them, or its shard of a multi-GPU run (wld_set_row_shard; the per-GPU
  // histograms are then summed before the filter — integer counts, so the sum is exact)
  const int64_t r_lo = std::min(c->row_lo, c->n_seqs), r_hi = c->row_hi < 0 ? c->n_seqs : std::min(c->row_hi, c->n_seqs);
  const int64_t n_rows = r_hi - r_lo;
  if (n_rows <= 0 || c->n_cols == 0) return WLD_OK;
  const uint8_t* raw = c->d_raw + r_lo * c->row_stride;

  const bool vec16 = (reinterpret_cast<uintptr_t>(raw) % 16 == 0) && (c->row_stride % 16 == 0) &&
                     (c->cols_padded <= c->row_stride);
  // Row chunk per block: enough blocks for several waves, at most 2040 rows (8 warps x 255).
  const int64_t col_blocks = vec16 ? (c->cols_padded + kHistColsPerBlock - 1) / kHistColsPerBlock
                                   : (c->n_cols + 255) / 256;
  int64_t want_blocks = (int64_t)c->sm_count * 16;
  int64_t row_chunks = (want_blocks + col_blocks - 1) / col_blocks;
  int64_t rows_per_block = (n_rows + row_chunks - 1) / row_chunks;
  rows_per_block = std::max<int64_t>(64, std::min<int64_t>(rows_per_block, 2040));
  rows_per_block = round_up(rows_per_block, 8);
  row_chunks = (n_rows + rows_per_block - 1) / rows_per_block;
  if (row_chunks > 65535) {
    rows_per_block = round_up((n_rows + 65534) / 65535, 8);
    row_chunks = (n_rows + rows_per_block - 1) / rows_per_block;
  }
  dim3 grid((unsigned)col_blocks, (unsigned)row_chunks);
  if (vec16) {
    // The vector path reads whole 16-byte groups up to cols_padded on every row.  A BORROWED buffer only has to
    // be readable up to (n_seqs-1)*row_stride + n_cols (include/wld.h), so its last row goes through the
    // byte-wise kernel when n_cols is not a multiple of 16; the library's own copy is padded.
    const bool tail_row = (c->input_flags & WLD_INPUT_DEVICE) && (c->n_cols % 16 != 0) && r_hi == c->n_seqs;
    const int64_t n_fast = tail_row ? n_rows - 1 : n_rows;
    if (n_fast > 0) {
      if (ascii)
        hist_vec16_kernel<true><<<grid, kHistThreads, 0, c->stream>>>(raw, n_fast, c->n_cols, c->row_stride,
                                                                     (int)rows_per_block, c->hist.as<uint32_t>(),
                                                                     c->cols_padded);
      else
        hist_vec16_kernel<false><<<grid, kHistThreads, 0, c->stream>>>(raw, n_fast, c->n_cols, c->row_stride,
                                                                      (int)rows_per_block, c->hist.as<uint32_t>(),
                                                                      c->cols_padded);
      tm.launched();
    }
    if (tail_row) {
      hist_generic_kernel<<<dim3((unsigned)((c->n_cols + 255) / 256), 1), 256, 0, c->stream>>>(
          raw, n_rows - 1, n_rows, c->n_cols, c->row_stride, 1, ascii, c->hist.as<uint32_t>(), c->cols_padded);
      tm.launched();
    }
  } else {
    hist_generic_kernel<<<grid, 256, 0, c->stream>>>(raw, 0, n_rows, c->n_cols, c->row_stride,
                                                     (int)rows_per_block, ascii, c->hist.as<uint32_t>(),
                                                     c->cols_padded);
    tm.launched();
  }
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

int run_filter(wld_ctx* c, int mode, float min_acgt, float min_minor, float max_minor, double py_min_acgt,
               double py_min_variability, ScopedStageTimer& tm) {
  const bool ascii = !(c->input_flags & WLD_INPUT_CODES);
  const int64_t cp = c->cols_padded;
  WLD_CUDA(c, c->keep.ensure((size_t)cp + 16));
  WLD_CUDA(c, c->rank.ensure(sizeof(int32_t) * ((size_t)cp + 1)));
  WLD_CUDA(c, c->maj_raw.ensure((size_t)cp + 16));
  WLD_CUDA(c, c->min_raw.ensure((size_t)cp + 16));
  WLD_CUDA(c, c->kept_count.ensure(sizeof(int32_t)));

  // main.rs:139: (min_acgt * n_seqs as f32).ceil() as usize  (f32 arithmetic; `as` saturates)
  unsigned long long min_count = 0;
  {
    const float v = ceilf(min_acgt * (float)c->n_seqs);
    if (v > 0.0f) min_count = v >= 18446744073709551616.0f ? ~0ull : (unsigned long long)v;
  }
  if (cp > 0) {
    decide_kernel<<<(unsigned)((cp + 255) / 256), 256, 0, c->stream>>>(
        c->hist.as<uint32_t>(), cp, c->n_cols, c->n_seqs, mode, min_count, min_minor, max_minor, py_min_acgt,
        py_min_variability, c->keep.as<uint8_t>(), c->maj_raw.as<int8_t>(), c->min_raw.as<int8_t>());
    tm.launched();
  }
  scan_kernel<<<1, kScanThreads, 0, c->stream>>>(c->keep.as<uint8_t>(), cp, c->rank.as<int32_t>(),
                                                 c->kept_count.as<int32_t>());
  tm.launched();
  WLD_CUDA(c, cudaGetLastError());
  int32_t kept = 0;
  WLD_CUDA(c, cudaMemcpyAsync(&kept, c->kept_count.p, sizeof kept, cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_kept = kept;
  c->ldc = round_up(std::max<int64_t>(c->n_seqs, 1), 128);

  WLD_CUDA(c, c->site_map.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(kept, 1)));
  WLD_CUDA(c, c->maj.ensure((size_t)std::max<int64_t>(kept, 1)));
  WLD_CUDA(c, c->mnr.ensure((size_t)std::max<int64_t>(kept, 1)));
  WLD_CUDA(c, c->codes.ensure((size_t)std::max<int64_t>(kept, 1) * (size_t)c->ldc));
  if (kept == 0 || c->n_seqs == 0) return WLD_OK;

  site_map_kernel<<<(unsigned)((c->n_cols + 255) / 256), 256, 0, c->stream>>>(
      c->keep.as<uint8_t>(), c->rank.as<int32_t>(), c->maj_raw.as<int8_t>(), c->min_raw.as<int8_t>(), c->n_cols,
      c->site_map.as<int32_t>(), c->maj.as<int8_t>(), c->mnr.as<int8_t>());
  tm.launched();

  const bool aligned16 = (reinterpret_cast<uintptr_t>(c->d_raw) % 16 == 0) && (c->row_stride % 16 == 0);
  dim3 grid((unsigned)((c->n_cols + 127) / 128), (unsigned)(c->ldc / 128));
  if (grid.y > 65535) return c->fail(WLD_ERR_UNSUPPORTED, "n_seqs %lld too large for the gather grid", (long long)c->n_seqs);
  if (ascii)
    gather_kernel<true><<<grid, kGatherThreads, 0, c->stream>>>(c->d_raw, c->n_seqs, c->n_cols, c->row_stride, aligned16,
                                                               c->keep.as<uint8_t>(), c->rank.as<int32_t>(),
                                                               c->codes.as<uint8_t>(), c->ldc);
  else
    gather_kernel<false><<<grid, kGatherThreads, 0, c->stream>>>(c->d_raw, c->n_seqs, c->n_cols, c->row_stride, aligned16,
                                                                c->keep.as<uint8_t>(), c->rank.as<int32_t>(),
                                                                c->codes.as<uint8_t>(), c->ldc);
  tm.launched();
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

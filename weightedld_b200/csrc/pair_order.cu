// pair_order.cu — puts the compacted survivors into the reference's output order on the device and
// maps kept-site indices to raw alignment columns.
//
// Reference: PairStore (lib.rs:523-576) keeps one Vec per 256x256 tile in triu_index order
// (lib.rs:623-632: tile rows bottom-up, columns ascending; rayon's collect is order-preserving,
// lib.rs:635-679) and inside a tile pairs are pushed with a ascending, then b ascending
// (lib.rs:647-653).  That order is the lexicographic order of (tile, a mod 256, b mod 256) with
// tile = (n-1-tr)(n-tr)/2 + (tc-tr) for tile row tr = a/256 and column tc = b/256 (n tiles per edge).
//
// No comparison sort is needed, because the key is UNIQUE: a site pair occurs once, so inside a reference
// tile the 16-bit key (a mod 256, b mod 256) addresses one bit of a 65536-bit map, and the rank of a record
// is the number of set bits below its own.  Four HBM-bound passes, all in this file:
//   count    one atomicAdd per record into its tile's counter            (reads 8 B / record)
//   scan     exclusive prefix over the n(n+1)/2 tile counters             (one block)
//   scatter  record -> its tile's segment, in arrival order               (20 B read + 20 B written)
//   place    one block per tile: bitmap in shared memory, popcount prefix, record -> segment[rank],
//            kept indices mapped to raw columns on the way (lib.rs:662-663)   (40 B read + 20 B written)
// (Round 1 used cub::DeviceRadixSort on a 48-bit key plus a gather: ~170 B of traffic per record and a
// library call on a hot-path row.)
#include "common.cuh"

namespace wld {
namespace {

__device__ __forceinline__ unsigned long long ref_tile_index(uint32_t a, uint32_t b, unsigned long long n) {
  const unsigned long long tr = a >> 8, tc = b >> 8;  // tc >= tr because b > a
  return (n - 1 - tr) * (n - tr) / 2 + (tc - tr);
}

// Survivors arrive grouped by kernel tile, so the lanes of a warp mostly hit the same one or two counters:
// lanes with the same tile elect a leader (match_any) and issue ONE atomic per group.
__device__ __forceinline__ unsigned group_by_tile(unsigned long long t, int& leader, int& rank, int& size) {
  const unsigned peers = __match_any_sync(0xffffffffu, t);
  leader = __ffs(peers) - 1;
  rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
  size = __popc(peers);
  return peers;
}

__global__ void tile_count_kernel(const wld_pair* __restrict__ pairs, uint64_t n, unsigned long long edge,
                                  uint32_t* __restrict__ count) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  // 20-byte records are 4-byte aligned: two 4-byte loads.  Lanes past the end form their own (ignored) group.
  const unsigned long long t = valid ? ref_tile_index(pairs[i].site_a, pairs[i].site_b, edge) : ~0ull;
  int leader, rank, size;
  group_by_tile(t, leader, rank, size);
  if (valid && rank == 0) atomicAdd(&count[t], (uint32_t)size);
}

// One block: exclusive prefix of count[0..m) into offset[0..m] (64-bit: more than 2^32 survivors are legal).
// Each thread owns a contiguous segment.  cursor[] is zeroed for the scatter pass.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(const uint32_t* __restrict__ count, uint64_t m,
                                                                 unsigned long long* __restrict__ offset,
                                                                 uint32_t* __restrict__ cursor) {
  __shared__ unsigned long long s_part[kScanThreads];
  const uint64_t seg = (m + kScanThreads - 1) / kScanThreads;
  const uint64_t lo = min((uint64_t)threadIdx.x * seg, m), hi = min(lo + seg, m);
  unsigned long long sum = 0;
  for (uint64_t i = lo; i < hi; ++i) sum += count[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < kScanThreads; o <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partials
    const unsigned long long v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0ull;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned long long run = s_part[threadIdx.x] - sum;
  for (uint64_t i = lo; i < hi; ++i) {
    offset[i] = run;
    run += count[i];
    cursor[i] = 0u;
  }
  if (threadIdx.x == kScanThreads - 1) offset[m] = s_part[kScanThreads - 1];
}

__global__ void tile_scatter_kernel(const wld_pair* __restrict__ pairs, uint64_t n, unsigned long long edge,
                                    const unsigned long long* __restrict__ offset, uint32_t* __restrict__ cursor,
                                    wld_pair* __restrict__ grouped) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  wld_pair p{};
  if (valid) p = pairs[i];
  const unsigned long long t = valid ? ref_tile_index(p.site_a, p.site_b, edge) : ~0ull;
  int leader, rank, size;
  group_by_tile(t, leader, rank, size);
  uint32_t base = 0;
  if (valid && rank == 0) base = atomicAdd(&cursor[t], (uint32_t)size);
  base = __shfl_sync(0xffffffffu, base, leader);
  if (valid) grouped[offset[t] + base + (uint32_t)rank] = p;
}

// One block per reference tile (grid-stride).  Small tiles rank by counting smaller keys directly; larger
// ones build the 65536-bit map of present keys in shared memory and rank by popcount.
constexpr int kPlaceThreads = 256;
constexpr int kSmallTile = 256;
__global__ void __launch_bounds__(kPlaceThreads) tile_place_kernel(const wld_pair* __restrict__ grouped,
                                                                   const unsigned long long* __restrict__ offset,
                                                                   uint64_t n_tiles, const int32_t* __restrict__ site_map,
                                                                   wld_pair* __restrict__ out) {
  __shared__ uint32_t s_bits[2048];
  __shared__ uint32_t s_pre[2048];
  __shared__ uint32_t s_scan[kPlaceThreads];
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const unsigned long long beg = offset[t], end = offset[t + 1];
    const uint32_t cnt = (uint32_t)(end - beg);  // <= 65536: the pairs of one 256 x 256 tile
    if (cnt == 0) continue;
    const wld_pair* src = grouped + beg;
    wld_pair* dst = out + beg;
    if (cnt <= kSmallTile) {
      uint32_t* keys = s_bits;
      if (threadIdx.x < cnt) keys[threadIdx.x] = ((src[threadIdx.x].site_a & 255u) << 8) | (src[threadIdx.x].site_b & 255u);
      __syncthreads();
      if (threadIdx.x < cnt) {
        const uint32_t k = keys[threadIdx.x];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < cnt; ++j) rank += keys[j] < k;
        wld_pair p = src[threadIdx.x];
        if (site_map) {  // lib.rs:662-663: parent_site_index
          p.site_a = (uint32_t)site_map[p.site_a];
          p.site_b = (uint32_t)site_map[p.site_b];
        }
        dst[rank] = p;
      }
      __syncthreads();
      continue;
    }
    for (int i = threadIdx.x; i < 2048; i += kPlaceThreads) s_bits[i] = 0u;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt; i += kPlaceThreads) {
      const uint32_t k = ((src[i].site_a & 255u) << 8) | (src[i].site_b & 255u);
      atomicOr(&s_bits[k >> 5], 1u << (k & 31u));
    }
    __syncthreads();
    {  // exclusive popcount prefix per word: thread owns 8 consecutive words
      uint32_t local[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        local[j] = sum;
        sum += __popc(s_bits[threadIdx.x * 8 + j]);
      }
      s_scan[threadIdx.x] = sum;
      __syncthreads();
      for (int o = 1; o < kPlaceThreads; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0u;
        __syncthreads();
        s_scan[threadIdx.x] += v;
        __syncthreads();
      }
      const uint32_t base = s_scan[threadIdx.x] - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) s_pre[threadIdx.x * 8 + j] = base + local[j];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt; i += kPlaceThreads) {
      wld_pair p = src[i];
      const uint32_t k = ((p.site_a & 255u) << 8) | (p.site_b & 255u);
      const uint32_t rank = s_pre[k >> 5] + __popc(s_bits[k >> 5] & ((1u << (k & 31u)) - 1u));
      if (site_map) {
        p.site_a = (uint32_t)site_map[p.site_a];
        p.site_b = (uint32_t)site_map[p.site_b];
      }
      dst[rank] = p;
    }
    __syncthreads();
  }
}

__global__ void map_pairs_kernel(const wld_pair* __restrict__ pairs, uint64_t n, const int32_t* __restrict__ site_map,
                                 wld_pair* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  wld_pair p = pairs[i];
  p.site_a = (uint32_t)site_map[p.site_a];  // lib.rs:662-663: parent_site_index
  p.site_b = (uint32_t)site_map[p.site_b];
  out[i] = p;
}

}  // namespace

// Writes the n survivors into c->sorted (device), ordered unless `!ordered`, with parent indices if
// `parent`.  Returns WLD_ERR_NOMEM when the scratch buffers do not fit (caller falls back to a host merge).
int run_pair_order(wld_ctx* c, bool ordered, bool parent) {
  const uint64_t n = c->n_survivors;
  if (n == 0) return WLD_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if ((n + 255) / 256 > 0x7fffffffull) return WLD_ERR_NOMEM;
  if (c->sorted.ensure(sizeof(wld_pair) * (size_t)n) != cudaSuccess) {
    cudaGetLastError();
    return WLD_ERR_NOMEM;
  }
  const int32_t* smap = parent ? c->site_map.as<int32_t>() : nullptr;
  ScopedStageTimer tm(c, WLD_STAGE_ORDER);
  if (!ordered) {
    map_pairs_kernel<<<blocks, 256, 0, c->stream>>>(c->pairs.as<wld_pair>(), n, smap, c->sorted.as<wld_pair>());
    tm.launched();
    WLD_CUDA(c, cudaGetLastError());
    return WLD_OK;
  }
  const unsigned long long edge = (unsigned long long)((c->n_kept + 255) / 256);
  const uint64_t n_tiles = edge * (edge + 1) / 2;
  if (c->sort_keys.ensure(sizeof(unsigned long long) * (size_t)(n_tiles + 1)) != cudaSuccess ||  // tile offsets
      c->sort_idx.ensure(sizeof(uint32_t) * 2 * (size_t)n_tiles) != cudaSuccess ||               // counts, cursors
      c->sort_temp.ensure(sizeof(wld_pair) * (size_t)n) != cudaSuccess) {                        // grouped records
    cudaGetLastError();
    return WLD_ERR_NOMEM;
  }
  unsigned long long* offset = c->sort_keys.as<unsigned long long>();
  uint32_t* count = c->sort_idx.as<uint32_t>();
  uint32_t* cursor = count + n_tiles;
  wld_pair* grouped = c->sort_temp.as<wld_pair>();
  WLD_CUDA(c, cudaMemsetAsync(count, 0, sizeof(uint32_t) * (size_t)n_tiles, c->stream));
  tile_count_kernel<<<blocks, 256, 0, c->stream>>>(c->pairs.as<wld_pair>(), n, edge, count);
  tile_scan_kernel<<<1, kScanThreads, 0, c->stream>>>(count, n_tiles, offset, cursor);
  tile_scatter_kernel<<<blocks, 256, 0, c->stream>>>(c->pairs.as<wld_pair>(), n, edge, offset, cursor, grouped);
  const unsigned place_blocks = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)c->sm_count * 8);
  tile_place_kernel<<<place_blocks, kPlaceThreads, 0, c->stream>>>(grouped, offset, n_tiles, smap, c->sorted.as<wld_pair>());
  tm.launched(4);
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

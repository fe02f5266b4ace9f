// pair_order.cu — puts the compacted survivors into the reference's output order on the device and
// maps kept-site indices to raw alignment columns.
//
// Reference: PairStore (lib.rs:523-576) keeps one Vec per 256x256 tile in triu_index order
// (lib.rs:623-632: tile rows bottom-up, columns ascending; rayon's collect is order-preserving,
// lib.rs:635-679) and inside a tile pairs are pushed with a ascending, then b ascending
// (lib.rs:647-653).  That order is the lexicographic order of (tile_key, a mod 256, b mod 256) with
// tile_key = (n-1-a/256)*n + b/256 — one radix sort of a <= 48-bit key.  The sort itself is
// cub::DeviceRadixSort (CCCL, ships with the CUDA toolkit); it is output formatting, not one of the
// three hot stages.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace wld {
namespace {

__global__ void make_keys_kernel(const wld_pair* __restrict__ pairs, uint64_t n, uint64_t n_tiles_edge,
                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t a = pairs[i].site_a, b = pairs[i].site_b;
  const unsigned long long tile = (n_tiles_edge - 1 - a / 256) * n_tiles_edge + b / 256;
  keys[i] = (tile << 16) | ((unsigned long long)(a & 255u) << 8) | (b & 255u);
  idx[i] = (uint32_t)i;
}

__global__ void gather_pairs_kernel(const wld_pair* __restrict__ pairs, const uint32_t* __restrict__ idx, uint64_t n,
                                    const int32_t* __restrict__ site_map, wld_pair* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  wld_pair p = pairs[idx ? idx[i] : i];
  if (site_map) {  // lib.rs:662-663: parent_site_index
    p.site_a = (uint32_t)site_map[p.site_a];
    p.site_b = (uint32_t)site_map[p.site_b];
  }
  out[i] = p;
}

}  // namespace

// Writes the n survivors into c->sorted (device), ordered unless `unordered`, with parent indices if
// `parent`.  Returns WLD_ERR_NOMEM when the scratch buffers do not fit (caller falls back to a host merge).
int run_pair_order(wld_ctx* c, bool ordered, bool parent) {
  const uint64_t n = c->n_survivors;
  if (n == 0) return WLD_OK;
  if (n >= (1ull << 32)) return WLD_ERR_NOMEM;  // 32-bit permutation indices
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (c->sorted.ensure(sizeof(wld_pair) * (size_t)n) != cudaSuccess) {
    cudaGetLastError();
    return WLD_ERR_NOMEM;
  }
  const int32_t* smap = parent ? c->site_map.as<int32_t>() : nullptr;
  if (!ordered) {
    gather_pairs_kernel<<<blocks, 256, 0, c->stream>>>(c->pairs.as<wld_pair>(), nullptr, n, smap, c->sorted.as<wld_pair>());
    WLD_CUDA(c, cudaGetLastError());
    return WLD_OK;
  }
  const uint64_t edge = (uint64_t)((c->n_kept + 255) / 256);
  int key_bits = 16;
  while (key_bits < 64 && (edge * edge) >> (key_bits - 16)) ++key_bits;
  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (long long)n, 0, key_bits, c->stream);
  if (c->sort_keys.ensure(sizeof(unsigned long long) * 2 * (size_t)n) != cudaSuccess ||
      c->sort_idx.ensure(sizeof(uint32_t) * 2 * (size_t)n) != cudaSuccess ||
      c->sort_temp.ensure(temp_bytes) != cudaSuccess) {
    cudaGetLastError();
    return WLD_ERR_NOMEM;
  }
  unsigned long long* k0 = c->sort_keys.as<unsigned long long>();
  uint32_t* i0 = c->sort_idx.as<uint32_t>();
  make_keys_kernel<<<blocks, 256, 0, c->stream>>>(c->pairs.as<wld_pair>(), n, edge, k0, i0);
  WLD_CUDA(c, cudaGetLastError());
  WLD_CUDA(c, cub::DeviceRadixSort::SortPairs(c->sort_temp.p, temp_bytes, k0, k0 + n, i0, i0 + n, (long long)n, 0, key_bits,
                                              c->stream));
  gather_pairs_kernel<<<blocks, 256, 0, c->stream>>>(c->pairs.as<wld_pair>(), i0 + n, n, smap, c->sorted.as<wld_pair>());
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

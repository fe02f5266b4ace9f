// pair_simt.cu — CUDA-core verification kernel of the pair stage (WLD_PAIR_KERNEL_SIMT).
//
// Same contract as the tcgen05 kernel (pair_umma.cu): exact integer Gram of the fixed-point
// weights + the shared f64 epilogue (pair_epilogue.cuh), hence bit-identical survivors.  It works
// from the 0..5 code matrix and the integer weights directly (not from the expanded bf16
// operands), so it also cross-checks the operand expansion.  FP64 FMAs on integers < 2^53 are
// exact.  It is a correctness path (about 20x slower than the tensor path), not a fallback the
// library picks on its own.
//
// Reference: single_weighted_ld_pair lib.rs:455-521, all_weighted_ld_pairs lib.rs:578-684.
#include "common.cuh"
#include "pair_epilogue.cuh"

namespace wld {
namespace {

constexpr int kTile = 64;  // sites per tile edge
constexpr int kKC = 16;    // sequences per shared-memory chunk

__global__ void __launch_bounds__(256) pair_simt_kernel(const uint8_t* __restrict__ codes, int64_t ldc,
                                                        int64_t n_kept, int64_t n_seqs,
                                                        const int8_t* __restrict__ maj,
                                                        const int8_t* __restrict__ mnr,
                                                        const double* __restrict__ q,
                                                        const uint2* __restrict__ tiles, float thr, double thr_lo,
                                                        const uint2* __restrict__ py_aux, PairOut out, unsigned long long* __restrict__ pairs_done) {
  __shared__ double sAM[kKC][kTile], sAm[kKC][kTile], sBM[kKC][kTile], sBm[kKC][kTile];
  const uint2 tile = tiles[blockIdx.x];
  const int64_t i0 = (int64_t)tile.x * kTile, j0 = (int64_t)tile.y * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  double acc[4][4][4];  // [u][v][AB,Ab,aB,ab]
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[u][v][t] = 0.0;

  // loader mapping: thread -> (site t = tid/4, four consecutive sequences)
  const int lt = threadIdx.x >> 2, ls = (threadIdx.x & 3) * 4;
  const int64_t ia = i0 + lt, jb = j0 + lt;
  const int a_maj = ia < n_kept ? maj[ia] : -1, a_min = ia < n_kept ? mnr[ia] : -1;
  const int b_maj = jb < n_kept ? maj[jb] : -1, b_min = jb < n_kept ? mnr[jb] : -1;

  for (int64_t s0 = 0; s0 < n_seqs; s0 += kKC) {
    uint32_t ca = 0x05050505u, cb = 0x05050505u;
    if (ia < n_kept) ca = *reinterpret_cast<const uint32_t*>(codes + ia * ldc + s0 + ls);
    if (jb < n_kept) cb = *reinterpret_cast<const uint32_t*>(codes + jb * ldc + s0 + ls);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t s = s0 + ls + b;
      const double w = s < n_seqs ? q[s] : 0.0;  // ldc padding holds code 5 anyway
      const int xa = (ca >> (8 * b)) & 0xff, xb = (cb >> (8 * b)) & 0xff;
      sAM[ls + b][lt] = xa == a_maj ? w : 0.0;
      sAm[ls + b][lt] = xa == a_min ? w : 0.0;
      sBM[ls + b][lt] = xb == b_maj ? 1.0 : 0.0;
      sBm[ls + b][lt] = xb == b_min ? 1.0 : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int s = 0; s < kKC; ++s) {
      double aM[4], am[4], bM[4], bm[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        aM[u] = sAM[s][ty + 16 * u];
        am[u] = sAm[s][ty + 16 * u];
        bM[u] = sBM[s][tx + 16 * u];
        bm[u] = sBm[s][tx + 16 * u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          acc[u][v][0] = fma(aM[u], bM[v], acc[u][v][0]);
          acc[u][v][1] = fma(aM[u], bm[v], acc[u][v][1]);
          acc[u][v][2] = fma(am[u], bM[v], acc[u][v][2]);
          acc[u][v][3] = fma(am[u], bm[v], acc[u][v][3]);
        }
    }
    __syncthreads();
  }

  unsigned long long done = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int64_t i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
      const bool valid = i < j && j < n_kept;  // lib.rs:651 (b_idx <= a_idx skipped)
      done += valid;
      bool keep = valid && ld_prefilter(acc[u][v][0], acc[u][v][1], acc[u][v][2], acc[u][v][3], thr_lo);
      float d = 0.f, dp = 0.f, r2 = 0.f;
      if (keep) keep = ld_stats_exact(acc[u][v][0], acc[u][v][1], acc[u][v][2], acc[u][v][3], thr, d, dp, r2, py_aux != nullptr);
      if (py_aux != nullptr && keep) keep = !py_flagged(py_aux, (uint32_t)i, (uint32_t)j);  // left to pair_python.cu
      emit_pairs_warp(keep, (uint32_t)i, (uint32_t)j, d, dp, r2, out);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
  if ((threadIdx.x & 31) == 0 && done) atomicAdd(pairs_done, done);
}

}  // namespace

int run_pair_simt(wld_ctx* c, float thr) {
  const int64_t L = c->n_kept;
  const int64_t nt = (L + kTile - 1) / kTile;
  std::vector<uint2> list;
  int64_t idx = 0;
  for (int64_t bi = 0; bi < nt; ++bi)
    for (int64_t bj = bi; bj < nt; ++bj, ++idx)
      if (idx % c->nparts == c->part) list.push_back(make_uint2((unsigned)bi, (unsigned)bj));
  c->info.tiles = (int64_t)list.size();
  c->info.tile_sites_m = kTile;
  c->info.tile_sites_n = kTile;
  c->info.executed_flop = 0.0;
  if (list.empty()) return WLD_OK;
  WLD_CUDA(c, c->simt_tiles.ensure(sizeof(uint2) * list.size()));
  WLD_CUDA(c, cudaMemcpyAsync(c->simt_tiles.p, list.data(), sizeof(uint2) * list.size(), cudaMemcpyHostToDevice, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));  // `list` is pageable and dies at return
  ScopedStageTimer tm(c, WLD_STAGE_PAIR);          // kernel only
  PairOut out{c->pairs.as<wld_pair>(), c->counters.as<unsigned long long>(), c->pair_cap};
  pair_simt_kernel<<<(unsigned)list.size(), 256, 0, c->stream>>>(
      c->codes.as<uint8_t>(), c->ldc, L, c->n_seqs, c->maj.as<int8_t>(), c->mnr.as<int8_t>(), c->q.as<double>(),
      c->simt_tiles.as<uint2>(), thr, ld_thr_lo(thr),
      c->compat == WLD_COMPAT_PYTHON ? c->py_aux.as<uint2>() : nullptr, out, c->counters.as<unsigned long long>() + 1);
  tm.launched();
  WLD_CUDA(c, cudaGetLastError());
  return WLD_OK;
}

}  // namespace wld

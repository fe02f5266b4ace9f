// tc_peak.cu — issue-rate ceiling of tcgen05.mma on this GPU, for the roofline denominator of the pair kernel.
//
// MEASURED_PEAKS.json (driver-written) holds a cuBLAS bf16 figure but no INT8 one, and a library GEMM is not a
// ceiling for kind::i8 anyway (round 1 beat torch._int_mm by 40 %).  This tool measures the instruction itself:
// every CTA pair (cluster of 2, one CTA per SM, all SMs) keeps its operands RESIDENT in shared memory and one
// thread issues back-to-back `tcgen05.mma.cta_group::2` M256 N256 K32 (kind::i8) or K16 (kind::f16, bf16) into two
// alternating TMEM accumulators — no TMA, no epilogue, no global traffic.  Nothing can run the tensor pipe faster
// than this, so it is a true peak: burst (best of 10 launches of ~25 ms) and sustained (back to back for >= 4 s
// under the 1000 W cap, clocks sampled by the caller).  Operand bytes are pseudo-random (full toggle rate, like a
// GEMM benchmark on random data) or, with --sparse, 0/1 indicators x random limbs like the pair kernel's operands.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_peak tools/tc_peak.cu
//   tools/tc_peak [--sparse] [--seconds 4]      -> one JSON line
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      std::fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      std::exit(1);                                                                             \
    }                                                                                           \
  } while (0)

namespace {

constexpr int kTileBytes = 128 * 128;  // 128 rows x one 128-byte swizzle atom (K = 128 u8 / 64 bf16)
constexpr int kBatch = 64;             // MMA groups (of 4 instructions) per commit

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000ll) return false;  // ~2 s: report instead of hanging the GPU
  return true;
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {  // K-major, SWIZZLE_128B, SBO 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <bool kI8>
__device__ __forceinline__ void umma_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  // M = 256 (pair), N = 256; kind::i8: D = S32, A = B = u8; kind::f16: D = F32, A = B = BF16
  constexpr uint32_t idesc = (kI8 ? (2u << 4) : ((1u << 4) | (1u << 7) | (1u << 10))) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
  if constexpr (kI8)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

// grid = 2 * pairs, cluster (2,1,1), 128 threads, 1 CTA per SM (TMEM fully allocated)
template <bool kI8>
__global__ void __launch_bounds__(128, 1) tc_peak_kernel(int batches, int sparse, uint32_t seed, int* error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;               // this CTA's 128 A rows
  uint8_t* sB = smem + kTileBytes;  // this CTA's half of the 256 B rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5;

  // operand bytes: xorshift per thread.  (The swizzle only permutes 16-byte chunks: irrelevant for random data.)
  uint32_t x = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u + 1u);
  for (int i = threadIdx.x * 4; i < 2 * kTileBytes; i += blockDim.x * 4) {
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    uint32_t v = x;
    if (sparse && i < kTileBytes) v &= kI8 ? 0x01010101u : 0u;  // indicators: 0/1 bytes
    if (!kI8) {
      // two bf16 values: small integers (exact) — 0..255 as bf16 bit patterns
      const uint32_t a = (sparse && i < kTileBytes) ? (x & 1u) : (x & 0xffu), b = (sparse && i < kTileBytes) ? ((x >> 8) & 1u) : ((x >> 8) & 0xffu);
      v = (__float_as_uint((float)a) >> 16) | (__float_as_uint((float)b) & 0xffff0000u);
    }
    *reinterpret_cast<uint32_t*>(smem + i) = v;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (rank == 0 && threadIdx.x == 0) {
    const uint64_t adesc = make_smem_desc(smem_u32(sA)), bdesc = make_smem_desc(smem_u32(sB));
    bool ok = true;
    for (int b = 0; b < batches && ok; ++b) {
      if (b >= 2) ok = mbar_wait_bounded(smem_u32(&bars[b & 1]), ((b >> 1) - 1) & 1);  // batch b-2 retired
      for (int g = 0; g < kBatch; ++g) {
        const uint32_t d = tmem_base + ((b * kBatch + g) & 1) * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_cg2<kI8>(d, adesc + 2u * k, bdesc + 2u * k, 1u);
      }
      // arrives on this CTA's barrier when everything issued so far has retired (only the leader waits)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&bars[b & 1])),
                   "h"((uint16_t)1)
                   : "memory");
    }
    for (int b = max(batches - 2, 0); b < batches && ok; ++b) ok = mbar_wait_bounded(smem_u32(&bars[b & 1]), (b >> 1) & 1);
    if (!ok) atomicExch(error_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <bool kI8>
float run_once(int grid, int batches, int sparse, int* d_err, cudaStream_t st) {
  auto kern = tc_peak_kernel<kI8>;
  const int smem = 1024 + 2 * kTileBytes + 64;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  CK(cudaLaunchKernelEx(&cfg, kern, batches, sparse, 0x9E3779B9u, d_err));
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms;
}

template <bool kI8>
void measure(const char* name, int grid, int sparse, double seconds, int* d_err, cudaStream_t st, char* out, size_t cap) {
  const double op_per_batch = (double)kBatch * 4.0 * 2.0 * 256.0 * 256.0 * (kI8 ? 32.0 : 16.0) * (grid / 2);
  // calibrate to ~25 ms per launch
  run_once<kI8>(grid, 8, sparse, d_err, st);
  float ms = run_once<kI8>(grid, 64, sparse, d_err, st);
  int batches = std::max(8, (int)(64 * 25.0 / std::max(ms, 1e-3f)));
  for (int i = 0; i < 3; ++i) run_once<kI8>(grid, batches, sparse, d_err, st);  // warm-up
  double best = 0.0;
  for (int i = 0; i < 10; ++i) {
    ms = run_once<kI8>(grid, batches, sparse, d_err, st);
    best = std::max(best, op_per_batch * batches / (ms * 1e-3));
  }
  // sustained: back to back for `seconds`
  double total_ms = 0.0, total_op = 0.0;
  const auto t0 = std::chrono::steady_clock::now();
  int launches = 0;
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
    total_ms += run_once<kI8>(grid, batches, sparse, d_err, st);
    total_op += op_per_batch * batches;
    ++launches;
  }
  std::snprintf(out, cap, "\"%s\": {\"burst_tops\": %.1f, \"sustained_tops\": %.1f, \"launch_ms\": %.2f, \"sustained_launches\": %d, \"sustained_s\": %.2f}",
                name, best / 1e12, total_op / (total_ms * 1e-3) / 1e12, total_ms / std::max(launches, 1), launches, total_ms * 1e-3);
}

}  // namespace

int main(int argc, char** argv) {
  int sparse = 0;
  double seconds = 4.0;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--sparse")) sparse = 1;
    else if (!std::strcmp(argv[i], "--seconds") && i + 1 < argc) seconds = std::atof(argv[++i]);
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  if (prop.major != 10) {
    std::fprintf(stderr, "needs an sm_100 GPU\n");
    return 1;
  }
  const int grid = prop.multiProcessorCount / 2 * 2;
  int* d_err = nullptr;
  CK(cudaMalloc(&d_err, sizeof(int)));
  CK(cudaMemset(d_err, 0, sizeof(int)));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  char a[512], b[512];
  measure<true>("i8", grid, sparse, seconds, d_err, st, a, sizeof a);
  measure<false>("bf16", grid, sparse, seconds, d_err, st, b, sizeof b);
  int err = 0;
  CK(cudaMemcpy(&err, d_err, sizeof err, cudaMemcpyDeviceToHost));
  std::printf("{\"tool\": \"tools/tc_peak.cu\", \"gpu\": \"%s\", \"sms\": %d, \"ctas\": %d, \"operands\": \"%s\", "
              "\"instruction\": \"tcgen05.mma.cta_group::2 M256 N256 (K32 i8 / K16 bf16), operands resident in shared memory\", "
              "%s, %s, \"watchdog\": %d}\n",
              prop.name, prop.multiProcessorCount, grid, sparse ? "0/1 indicators x random limbs" : "random bytes", a, b, err);
  return err ? 2 : 0;
}

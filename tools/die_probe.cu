// die_probe.cu — experiment: find which L2 die every SM sits on by timing atomics (served at the home
// L2 slice of an address) from every SM to a handful of addresses.  Build: nvcc -arch=sm_100a -o die_probe die_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <vector>

constexpr int kAddr = 32;
constexpr int kReps = 6;

__device__ __forceinline__ long long clock_after(unsigned dep) {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory");
  return t;
}
__global__ void probe(unsigned int* buf, int stride_words, unsigned int* sync, unsigned short* lat, int* smid_of_block,
                      unsigned zero) {
  extern __shared__ unsigned char pad[];
  if (threadIdx.x != 0) return;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  smid_of_block[blockIdx.x] = (int)smid;
  // all blocks resident (one per SM) before anyone probes
  atomicAdd(&sync[0], 1u);
  long long t0 = clock64();
  while (*(volatile unsigned int*)&sync[0] < gridDim.x) {
    if (clock64() - t0 > 4000000000ll) { sync[2] = 1; return; }
    __nanosleep(100);
  }
  // one SM at a time
  t0 = clock64();
  while (*(volatile unsigned int*)&sync[1] != blockIdx.x) {
    if (clock64() - t0 > 8000000000ll) { sync[2] = 2; return; }
    __nanosleep(500);
  }
  for (int a = 0; a < kAddr; ++a) {
    unsigned int* p = buf + (size_t)a * stride_words;
    unsigned best = 0xffffffffu;
    unsigned v = 0;
    for (int r = 0; r < kReps; ++r) {
      const long long s = clock_after(v);
      asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(v) : "l"(p + v), "r"(zero) : "memory");  // v stays 0
      if (v != 0u) break;  // control dependency: the second clock read cannot issue before v is back
      const long long e = clock64();
      best = min(best, (unsigned)(e - s));
    }
    lat[smid * kAddr + a] = (unsigned short)min(best, 65535u);
  }
  __threadfence();
  atomicExch(&sync[1], blockIdx.x + 1);
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int stride_words = (4096 + 256) / 4;
  unsigned int *buf, *sync;
  unsigned short* lat;
  int* smid_of_block;
  cudaMalloc(&buf, (size_t)kAddr * stride_words * 4 + 4096);
  cudaMemset(buf, 0, (size_t)kAddr * stride_words * 4 + 4096);
  cudaMalloc(&sync, 16);
  cudaMalloc(&lat, sms * kAddr * 2);
  cudaMalloc(&smid_of_block, sms * 4);
  for (int trial = 0; trial < 3; ++trial) {
    cudaMemset(sync, 0, 16);
    cudaMemset(lat, 0, sms * kAddr * 2);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
    probe<<<sms, 32, 150 * 1024>>>(buf, stride_words, sync, lat, smid_of_block, 0u);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned hs[4];
    cudaMemcpy(hs, sync, 16, cudaMemcpyDeviceToHost);
    std::vector<unsigned short> h(sms * kAddr);
    std::vector<int> sb(sms);
    cudaMemcpy(h.data(), lat, sms * kAddr * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(sb.data(), smid_of_block, sms * 4, cudaMemcpyDeviceToHost);
    printf("trial %d: %s flag=%u\n", trial, cudaGetErrorString(e), hs[2]);
    // per address: min / max over SMs, threshold at the midpoint
    std::vector<std::vector<int>> bit(sms, std::vector<int>(kAddr));
    for (int a = 0; a < kAddr; ++a) {
      int lo = 1 << 30, hi = 0;
      for (int s = 0; s < sms; ++s) { lo = std::min<int>(lo, h[s * kAddr + a]); hi = std::max<int>(hi, h[s * kAddr + a]); }
      for (int s = 0; s < sms; ++s) bit[s][a] = h[s * kAddr + a] * 2 > lo + hi;
      if (a < 6) printf("  addr %d: lat min %d max %d\n", a, lo, hi);
    }
    int n1 = 0, amb = 0;
    printf("  die: ");
    for (int s = 0; s < sms; ++s) {
      int hd = 0;
      for (int a = 0; a < kAddr; ++a) hd += bit[s][a] != bit[0][a];
      const int d = hd * 2 > kAddr;
      if (hd > kAddr / 4 && hd < 3 * kAddr / 4) ++amb;
      n1 += d;
      printf("%d", d);
    }
    printf("\n  die1 SMs %d of %d, ambiguous %d; bid->smid first 16:", n1, sms, amb);
    for (int b = 0; b < 16; ++b) printf(" %d", sb[b]);
    printf("\n  SM0 lat:");
    for (int a = 0; a < kAddr; ++a) printf(" %d", h[a]);
    printf("\n  SM%d lat:", sms - 1);
    for (int a = 0; a < kAddr; ++a) printf(" %d", h[(sms - 1) * kAddr + a]);
    printf("\n");
  }
  return 0;
}

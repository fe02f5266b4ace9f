// cuda_init_floor.cu — what a process pays before its first kernel can run on this box: driver initialisation +
// primary-context creation on device 0 (cudaFree(0)), first allocation, first launch, and process teardown
// (measured by the caller as wall - printed total).  The `weighted_ld` CLI cannot start faster than this.
//   nvcc -O2 -o tools/cuda_init_floor tools/cuda_init_floor.cu ; CUDA_VISIBLE_DEVICES=0 tools/cuda_init_floor
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
__global__ void nop() {}
int main() {
  using C = std::chrono::steady_clock;
  auto t0 = C::now();
  cudaFree(0);
  auto t1 = C::now();
  void* p = nullptr;
  cudaMalloc(&p, 1 << 20);
  nop<<<1, 1>>>();
  cudaDeviceSynchronize();
  auto t2 = C::now();
  void* h = nullptr;
  cudaMallocHost(&h, 64 << 20);
  auto t3 = C::now();
  auto ms = [](C::duration d) { return std::chrono::duration<double, std::milli>(d).count(); };
  std::printf("{\"context_ms\": %.1f, \"first_malloc_launch_ms\": %.1f, \"pinned_64MB_ms\": %.1f, \"total_ms\": %.1f}\n", ms(t1 - t0),
              ms(t2 - t1), ms(t3 - t2), ms(t3 - t0));
  return 0;
}

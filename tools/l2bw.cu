// Microbenchmark: L2-hit streaming read bandwidth (what bounds streaming query tiles to the SMs)
// and HBM read bandwidth, with coalesced 128-bit loads.  nvcc -arch=sm_100a -O3 tools/l2bw.cu -o l2bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const int4* __restrict__ p, size_t n_per_pass, int passes, int* sink) {
  int acc = 0;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < passes; r++)
    for (size_t i = tid; i < n_per_pass; i += stride * 4) {
      int4 a = p[i], b = (i + stride < n_per_pass) ? p[i + stride] : a;
      int4 c = (i + 2 * stride < n_per_pass) ? p[i + 2 * stride] : a, d = (i + 3 * stride < n_per_pass) ? p[i + 3 * stride] : a;
      acc += a.x ^ b.y ^ c.z ^ d.w;
    }
  if (acc == 0x12345678) *sink = acc;
}
int main() {
  int* sink; cudaMalloc(&sink, 4);
  size_t sizes_mb[] = {8, 32, 64, 96, 256, 4096};
  for (size_t mb : sizes_mb) {
    size_t bytes = mb << 20; int4* p; cudaMalloc(&p, bytes); cudaMemset(p, 1, bytes);
    int passes = (int)((16ull << 30) / bytes); if (passes < 2) passes = 2;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    rd<<<148 * 8, 512>>>(p, bytes / 16, 2, sink);
    cudaEventRecord(a); rd<<<148 * 8, 512>>>(p, bytes / 16, passes, sink); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("working set %5zu MB: %.1f GB/s\n", mb, (double)bytes * passes / ms / 1e6);
    cudaFree(p);
  }
  return 0;
}

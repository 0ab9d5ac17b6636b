// Microbenchmark: L2 -> shared-memory bandwidth of 1-D bulk async copies (cp.async.bulk, the TMA
// engine) when every SM streams random 8 KB query tiles out of an L2-resident 8 MB table: the
// operand stream of the passage-major MaxSim kernel (8 KB per (query, passage) pair).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/bulk_l2_bw.cu -o tools/bulk_l2_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one issuing thread per CTA; `stages` buffers of `tile` bytes, each filled by tile/chunk copies
__global__ void __launch_bounds__(128, 1) k(const uint8_t* table, uint32_t ntiles8k, int stages, int tile, int chunk, int iters,
                                            long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t rng = 0x9e3779b9u * (blockIdx.x + 1);
    const long long t0 = clock64();
    for (int it = 0; it < iters + stages; it++) {
      const int st = it % stages;
      if (it >= stages) { while (!try_wait(&bar[st], ((it / stages) - 1) & 1)) {} }
      if (it < iters) {
        expect_tx(&bar[st], tile);
        for (int c = 0; c < tile; c += chunk) {
          rng = rng * 1664525u + 1013904223u;
          const uint32_t q = (rng >> 8) % ntiles8k;
          bulk_g2s(smem + (size_t)st * tile + c, table + (size_t)q * 8192 + (c % 8192), chunk, &bar[st]);
        }
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t table_bytes = 8u << 20;
  uint8_t* table; long long* out;
  cudaMalloc(&table, table_bytes); cudaMemset(table, 1, table_bytes); cudaMalloc(&out, 148 * 8);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  struct Cfg { int stages, tile, chunk; } cfgs[] = {{2, 32768, 8192}, {3, 32768, 8192}, {4, 32768, 8192}, {5, 32768, 8192},
                                                    {6, 32768, 8192}, {6, 32768, 4096}, {6, 32768, 2048}, {8, 16384, 8192}, {8, 24576, 8192},
                                                    {3, 65536, 8192}};
  for (auto c : cfgs) {
    const int iters = 20000;
    const size_t smem = (size_t)c.stages * c.tile;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148, 128, smem>>>(table, (uint32_t)(table_bytes / 8192), c.stages, c.tile, c.chunk, 200, out);
    cudaEventRecord(a);
    k<<<148, 128, smem>>>(table, (uint32_t)(table_bytes / 8192), c.stages, c.tile, c.chunk, iters, out);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double bytes = 148.0 * iters * c.tile;
    printf("stages %d x %5d B (copies of %4d B): %7.1f GB/s aggregate, %6.1f B/clk/SM, %.0f cycles per tile  (%s)\n", c.stages, c.tile,
           c.chunk, bytes / ms / 1e6, (double)iters * c.tile / (double)h, (double)h / iters, cudaGetErrorString(e));
  }
  return 0;
}

// Independent microbenchmark of the L2 -> SM stream that bounds k_maxsim_tc:
// 148 persistent CTAs, every SM pulls random pieces of an L2-resident 8 MB table (the size of the
// fp16 query row image of a 1024-query batch) into shared memory with 1-D bulk async copies
// (cp.async.bulk = the TMA engine, SASS UBLKCP), exactly the instruction the kernel's loader warps
// issue.  Swept: issuing warps, stages in flight, bytes per stage, bytes per copy ("piece").
// A second mode reads the same table with plain 128-bit loads (LSU path) to separate "what the
// TMA engine sustains" from "what the L2 / crossbar sustains".
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/l2_to_sm_ceiling.cu -o tools/l2_to_sm_ceiling
//   tools/l2_to_sm_ceiling            # prints one line per configuration
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

constexpr int MAXW = 8, MAXS = 8;

// warp w < nw: lane 0 issues; the warp owns `stages` buffers of `tile` bytes, each filled by tile/piece copies
// from random piece-aligned offsets of the table.
__global__ void __launch_bounds__(256, 1) k_bulk(const uint8_t* table, uint32_t table_bytes, int nw, int stages, int tile, int piece,
                                                 int iters, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[MAXW][MAXS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAXW; i++) for (int j = 0; j < MAXS; j++) mbar_init(&bar[i][j], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nw && lane == 0) {
    uint32_t rng = 0x9e3779b9u * (blockIdx.x * MAXW + warp + 1);
    const uint32_t npieces = table_bytes / piece;
    uint8_t* buf = smem + (size_t)warp * stages * tile;
    for (int it = 0; it < iters + stages; it++) {
      const int st = it % stages;
      if (it >= stages) { while (!try_wait(&bar[warp][st], ((it / stages) - 1) & 1)) {} }
      if (it < iters) {
        expect_tx(&bar[warp][st], tile);
        for (int c = 0; c < tile; c += piece) {
          rng = rng * 1664525u + 1013904223u;
          bulk_g2s(buf + (size_t)st * tile + c, table + (size_t)((rng >> 8) % npieces) * piece, piece, &bar[warp][st]);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

// LSU path: every warp reads random 8 KB tiles with coalesced 128-bit loads (16 loads of 512 B per tile), 2 tiles in flight.
__global__ void __launch_bounds__(1024, 1) k_ldg(const uint8_t* table, uint32_t table_bytes, int iters, int* sink, long long* clk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 32 + warp + 1);
  const uint32_t ntiles = table_bytes / 8192;
  int acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    rng = rng * 1664525u + 1013904223u;
    const int4* src = reinterpret_cast<const int4*>(table + (size_t)((rng >> 8) % ntiles) * 8192) + lane;
    int4 v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = __ldcg(src + j * 32);
#pragma unroll
    for (int j = 0; j < 16; j++) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678) *sink = acc;
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

int main() {
  const uint32_t table_bytes = 8u << 20;
  uint8_t* table; long long* clk; int* sink;
  cudaMalloc(&table, table_bytes); cudaMemset(table, 1, table_bytes); cudaMalloc(&clk, 148 * 8); cudaMalloc(&sink, 4);
  long long h_clk[148];
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  printf("# table %u MB (L2 resident), 148 CTAs x 1/SM; rate = bytes moved / CUDA-event time; clk = SM clock seen by the kernel\n", table_bytes >> 20);
  struct Cfg { int nw, stages, tile, piece; } cfgs[] = {
      // the kernel's own shape: 2 loader warps sharing 3 x 32 KB stages of 4 x 8 KB pieces ~ {1 warp, 3 stages} and {2 warps, 2 stages}
      {1, 3, 32768, 8192}, {1, 6, 32768, 8192}, {2, 2, 32768, 8192}, {2, 3, 32768, 8192}, {3, 2, 32768, 8192}, {4, 2, 16384, 8192},
      {4, 3, 16384, 8192}, {6, 4, 8192, 8192}, {8, 3, 8192, 8192},
      // piece size at fixed bytes in flight (2 warps x 3 x 32 KB)
      {2, 3, 32768, 1024}, {2, 3, 32768, 2048}, {2, 3, 32768, 4096}, {2, 3, 32768, 16384}, {2, 3, 32768, 32768},
      // bytes in flight with 8 KB pieces, one issuer
      {1, 2, 8192, 8192}, {1, 4, 8192, 8192}, {1, 8, 8192, 8192}, {1, 2, 32768, 32768}, {1, 4, 32768, 32768}, {1, 6, 32768, 32768},
  };
  for (auto c : cfgs) {
    const int iters = 200000 / (c.tile / 1024);   // ~ 200 MB per warp
    const size_t smem = (size_t)c.nw * c.stages * c.tile;
    if (smem > 220 * 1024) { printf("skip (smem)\n"); continue; }
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_bulk<<<148, 256, smem>>>(table, table_bytes, c.nw, c.stages, c.tile, c.piece, iters / 10, clk);
    cudaEventRecord(a);
    k_bulk<<<148, 256, smem>>>(table, table_bytes, c.nw, c.stages, c.tile, c.piece, iters, clk);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(h_clk, clk, sizeof(h_clk), cudaMemcpyDeviceToHost);
    long long cmax = 0; for (int i = 0; i < 148; i++) cmax = h_clk[i] > cmax ? h_clk[i] : cmax;
    const double bytes = 148.0 * c.nw * (double)iters * c.tile;
    const double ghz = cmax / (ms * 1e6);
    printf("bulk  issuers %d  stages %d x %5d B  piece %5d B  in flight %3d KB : %8.1f GB/s  %5.1f B/clk/SM  (clk %.2f GHz, %s)\n", c.nw,
           c.stages, c.tile, c.piece, (int)(smem >> 10), bytes / ms / 1e6, bytes / 148.0 / (double)cmax, ghz, cudaGetErrorString(e));
  }
  for (int warps : {8, 16, 32}) {
    const int iters = 20000;
    k_ldg<<<148, warps * 32>>>(table, table_bytes, iters / 10, sink, clk);
    cudaEventRecord(a);
    k_ldg<<<148, warps * 32>>>(table, table_bytes, iters, sink, clk);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(h_clk, clk, sizeof(h_clk), cudaMemcpyDeviceToHost);
    long long cmax = 0; for (int i = 0; i < 148; i++) cmax = h_clk[i] > cmax ? h_clk[i] : cmax;
    const double bytes = 148.0 * warps * (double)iters * 8192;
    printf("ldg   warps %2d (16 x LDG.128 per lane in flight)                     : %8.1f GB/s  %5.1f B/clk/SM  (clk %.2f GHz, %s)\n", warps,
           bytes / ms / 1e6, bytes / 148.0 / (double)cmax, cmax / (ms * 1e6), cudaGetErrorString(e));
  }
  return 0;
}

// Microbenchmark 2: what limits the L2 -> shared-memory query-tile stream of the MaxSim kernel?
// tools/bulk_l2_bw.cu shows ONE issuing thread tops out at ~30 B/clk/SM with 8 KB bulk copies no
// matter how many stages are in flight (a fixed ~160 clocks per copy + ~75 B/clk).  This one asks:
//   A) does it scale with the number of ISSUING WARPS (per-thread issue cost) or not (one TMA unit)?
//   B) what does the LSU path (cp.async 16 B, LDGSTS) deliver with W warps?
//   C) are the two paths additive?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/bulk_l2_bw2.cu -o tools/bulk_l2_bw2
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
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

constexpr int TILE = 32768, Q = 8192;

// warps [0, wb) : bulk issuers (lane 0), each with its own 2 x 32 KB... too much smem for 4 -> 2 x 16 KB halves:
// every issuer owns `STG` stage buffers of `tile_b` bytes.  warps [wb, wb + wl): LSU copiers (cp.async), each
// copying 8 KB pieces into its own 2 x 8 KB double buffer.
__global__ void __launch_bounds__(512, 1) k(const uint8_t* table, uint32_t nq, int wb, int wl, int tile_b, int iters,
                                            long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; i++) { mbar_init(&bar[i][0], 1); mbar_init(&bar[i][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 16 + warp + 1);
  const long long t0 = clock64();
  if (warp < wb) {
    if (lane == 0) {
      uint8_t* buf = smem + (size_t)warp * 2 * tile_b;
      for (int it = 0; it < iters + 2; it++) {
        const int st = it & 1;
        if (it >= 2) { while (!try_wait(&bar[warp][st], ((it >> 1) - 1) & 1)) {} }
        if (it < iters) {
          expect_tx(&bar[warp][st], tile_b);
          for (int c = 0; c < tile_b; c += Q) {
            rng = rng * 1664525u + 1013904223u;
            bulk_g2s(buf + (size_t)st * tile_b + c, table + (size_t)((rng >> 8) % nq) * Q, Q, &bar[warp][st]);
          }
        }
      }
    }
  } else if (warp < wb + wl) {
    uint8_t* buf = smem + (size_t)wb * 2 * tile_b + (size_t)(warp - wb) * 2 * Q;
    // per iteration one 8 KB piece: 512 x 16 B = 16 per lane; keep 2 groups in flight
    const int n = iters * (tile_b / Q) * wb / (wl > 0 ? wl : 1) + (wb == 0 ? iters * 4 : 0);
    for (int it = 0; it < n; it++) {
      rng = rng * 1664525u + 1013904223u;
      const uint8_t* src = table + (size_t)((rng >> 8) % nq) * Q;
      uint8_t* dst = buf + (size_t)(it & 1) * Q;
#pragma unroll
      for (int j = 0; j < 16; j++) cp16(dst + (j * 32 + lane) * 16, src + (j * 32 + lane) * 16);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
}

int main() {
  const size_t table_bytes = 8u << 20;
  uint8_t* table; long long* out;
  cudaMalloc(&table, table_bytes); cudaMemset(table, 1, table_bytes); cudaMalloc(&out, 148 * 16 * 8);
  struct Cfg { int wb, wl, tile; } cfgs[] = {{1, 0, 32768}, {2, 0, 32768}, {3, 0, 32768}, {2, 0, 16384}, {4, 0, 16384}, {4, 0, 8192}, {8, 0, 8192},
                                             {0, 1, 0}, {0, 2, 0}, {0, 4, 0}, {0, 8, 0}, {1, 2, 32768}, {1, 4, 32768}, {2, 4, 32768}};
  for (auto c : cfgs) {
    const int iters = 10000;
    const size_t smem = (size_t)c.wb * 2 * c.tile + (size_t)c.wl * 2 * Q;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<148, 512, smem>>>(table, (uint32_t)(table_bytes / Q), c.wb, c.wl, c.tile, 100, out);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<<<148, 512, smem>>>(table, (uint32_t)(table_bytes / Q), c.wb, c.wl, c.tile, iters, out);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double bytes_bulk = 148.0 * iters * (double)c.tile * c.wb;
    double bytes_lsu = 0;
    if (c.wl > 0) {
      const long long n = (long long)iters * (c.tile / Q) * c.wb / c.wl + (c.wb == 0 ? iters * 4 : 0);
      bytes_lsu = 148.0 * n * Q * c.wl;
    }
    printf("bulk issuers %d (tile %5d) + lsu warps %d: %8.1f GB/s total (bulk %.0f + lsu %.0f GB/s), %.1f B/clk/SM @1.965GHz  (%s)\n", c.wb, c.tile,
           c.wl, (bytes_bulk + bytes_lsu) / ms / 1e6, bytes_bulk / ms / 1e6, bytes_lsu / ms / 1e6,
           (bytes_bulk + bytes_lsu) / ms / 1e6 / 148 / 1.965, cudaGetErrorString(e));
  }
  return 0;
}

// Microbenchmark: TMEM -> register-file read bandwidth (tcgen05.ld), the floor of any epilogue that
// must look at every accumulator element (MaxSim: max over document tokens).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/tmem_bw.cu -o tools/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, int nwarps, long long* out, float* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; it++) {
      if (MODE == 32) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
          uint32_t r[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(base + ((c & 7) * 32))
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc += __uint_as_float(r[0] ^ r[31]);
        }
      } else if (MODE == 320) {   // 8 loads in flight, one wait
        uint32_t r[8][32];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[c][0]), "=r"(r[c][1]), "=r"(r[c][2]), "=r"(r[c][3]), "=r"(r[c][4]), "=r"(r[c][5]), "=r"(r[c][6]), "=r"(r[c][7]), "=r"(r[c][8]),
                "=r"(r[c][9]), "=r"(r[c][10]), "=r"(r[c][11]), "=r"(r[c][12]), "=r"(r[c][13]), "=r"(r[c][14]), "=r"(r[c][15]), "=r"(r[c][16]),
                "=r"(r[c][17]), "=r"(r[c][18]), "=r"(r[c][19]), "=r"(r[c][20]), "=r"(r[c][21]), "=r"(r[c][22]), "=r"(r[c][23]), "=r"(r[c][24]),
                "=r"(r[c][25]), "=r"(r[c][26]), "=r"(r[c][27]), "=r"(r[c][28]), "=r"(r[c][29]), "=r"(r[c][30]), "=r"(r[c][31])
              : "r"(base + ((c & 7) * 32))
              : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 8; c++) acc += __uint_as_float(r[c][0] ^ r[c][31]);
      } else if (MODE == 8) {   // 16x256b.x8? use 32x32b.x8 small loads
#pragma unroll
        for (int c = 0; c < 32; c++) {
          uint32_t r[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(base + c * 8)
                       : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc += __uint_as_float(r[0] ^ r[7]);
        }
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && warp < nwarps) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 123.456f) *sink = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int nwarps) {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  k<MODE><<<148, 512>>>(10, nwarps, out, sink);
  k<MODE><<<148, 512>>>(iters, nwarps, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  // per iteration each warp reads 256 columns x 32 lanes x 4 B = 32 KB
  double cyc = (double)h[0] / iters;
  printf("%-34s warps=%2d: %8.1f cyc per 256 cols/warp -> %6.1f B/clk/warp, %7.1f B/clk/SM  (%s)\n", name, nwarps, cyc,
         32768.0 / cyc, 32768.0 * nwarps / cyc, cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int nw : {1, 4, 8, 16}) {
    run<32>("32x32b.x32, wait after each", nw);
    run<320>("32x32b.x32, 8 in flight", nw);
    run<8>("32x32b.x8, wait after each", nw);
  }
  return 0;
}

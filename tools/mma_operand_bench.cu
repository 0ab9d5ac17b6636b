// Microbenchmark: clocks per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N and of where the operands come from:
//   ss      A and B from shared memory (what k_maxsim_tc issues)
//   ss_akeep  same, A kept in the collector (.collector::a::fill once, ::use afterwards): A stationary, B streams
//   ts      A from tensor memory
//   ws_fill weight-stationary form (tcgen05.mma.ws), B filled into a collector buffer by every instruction
//   ws_use  weight-stationary form, the 4 collector buffers b0..b3 filled once (4 K-slices of B), then ::use: B stationary, A streams
// One thread issues `reps` rounds of 4 K-steps (fresh A slice each round in the ws/ss modes: rounds cycle over 4 A tiles), commits,
// waits; clocks per instruction = elapsed / (4 reps).  All 148 SMs run the same loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_operand_bench tools/mma_operand_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

#define MMA(OPCODE, AOP, AC)                                                                         \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" OPCODE " [%0], " AOP ", %2, %3, p;\n\t}" ::"r"(d), AC(a), "l"(b), \
               "r"(idesc), "r"(acc) : "memory")

enum { SS = 0, SS_AKEEP = 1, TS = 2, WS_FILL = 3, WS_USE = 4, WS_USE_TS = 5 };

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int reps, int N, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t s_bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (4 * 32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  if (threadIdx.x == 0) {
    // A tiles: 4 x (128 rows x 128 fp16, two 64-element K blocks of 16 KB: SBO 1024) at smem + t * 32 KB; B tile (N <= 256 rows) at +128 KB
    const uint32_t a_lo0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(smem + 4 * 32768) & 0x3ffffu) >> 4) | (1u << 16);
    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = idesc_f16(128, N);
    const uint32_t d = tm;                       // accumulator: columns 0 .. N-1
    const uint32_t ta = tm + 256;                // A in tensor memory (TS modes): columns 256 .. 319
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      const uint32_t a_t = a_lo0 + (uint32_t)(r & 3) * (32768u >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const uint64_t b = ((uint64_t)HI << 32) | (uint64_t)(b_lo0 + ks * 2);
        const uint32_t acc = (r | ks) ? 1u : 0u;
        if (MODE == SS) { const uint64_t a = ((uint64_t)HI << 32) | (uint64_t)(a_t + ks * 2); MMA("tcgen05.mma.cta_group::1.kind::f16", "%1", "l"); }
        else if (MODE == SS_AKEEP) {
          const uint64_t a = ((uint64_t)HI << 32) | (uint64_t)(a_lo0);
          if (r == 0 && ks == 0) MMA("tcgen05.mma.cta_group::1.kind::f16.collector::a::fill", "%1", "l");
          else MMA("tcgen05.mma.cta_group::1.kind::f16.collector::a::use", "%1", "l");
        }
        else if (MODE == TS) { const uint32_t a = ta + ks * 8; MMA("tcgen05.mma.cta_group::1.kind::f16", "[%1]", "r"); }
        else if (MODE == WS_FILL) { const uint64_t a = ((uint64_t)HI << 32) | (uint64_t)(a_t + ks * 2); MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill", "%1", "l"); }
        else if (MODE == WS_USE || MODE == WS_USE_TS) {
          const uint64_t a64 = ((uint64_t)HI << 32) | (uint64_t)(a_t + ks * 2);
          const uint32_t a32 = ta + ks * 8;
#define WS4(SUFFIX)                                                                                                         \
          if (MODE == WS_USE) { const uint64_t a = a64;                                                                     \
            if (ks == 0) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0" SUFFIX, "%1", "l");                      \
            else if (ks == 1) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b1" SUFFIX, "%1", "l");                 \
            else if (ks == 2) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b2" SUFFIX, "%1", "l");                 \
            else MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b3" SUFFIX, "%1", "l");                              \
          } else { const uint32_t a = a32;                                                                                  \
            if (ks == 0) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0" SUFFIX, "[%1]", "r");                    \
            else if (ks == 1) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b1" SUFFIX, "[%1]", "r");               \
            else if (ks == 2) MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b2" SUFFIX, "[%1]", "r");               \
            else MMA("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b3" SUFFIX, "[%1]", "r");                            \
          }
          if (r == 0) { WS4("::fill") } else { WS4("::use") }
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int N) {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  const size_t smem = 1024 + 4 * 32768 + 65536;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 4000;
  k<MODE><<<148, 128, smem>>>(64, N, out);
  k<MODE><<<148, 128, smem>>>(reps, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const double cyc = (double)h[0] / (4.0 * reps);
  printf("%-10s N=%3d: %7.1f clk per MMA  (math floor %5.1f; A 4096 B + B %5d B: %5.1f B/clk)  %s\n", name, N, cyc, N / 2.0, N * 32,
         (MODE == SS || MODE == WS_FILL ? 4096.0 + N * 32 : (MODE == SS_AKEEP ? N * 32.0 : 4096.0)) / cyc, cudaGetErrorString(e));
  cudaFree(out);
  if (e != cudaSuccess) { cudaDeviceReset(); }
}

int main(int argc, char** argv) {   // mma_operand_bench MODE N   (one configuration per process: an illegal shape kills the context)
  const int mode = argc > 1 ? atoi(argv[1]) : 0, N = argc > 2 ? atoi(argv[2]) : 80;
  switch (mode) {
    case SS: run<SS>("ss", N); break;
    case SS_AKEEP: run<SS_AKEEP>("ss_akeep", N); break;
    case TS: run<TS>("ts", N); break;
    case WS_FILL: run<WS_FILL>("ws_fill", N); break;
    case WS_USE: run<WS_USE>("ws_use", N); break;
    default: run<WS_USE_TS>("ws_use_ts", N); break;
  }
  return 0;
}

// stage1_tc.cu -- stage 1 on the 5th-generation tensor cores: (query tokens) x (centroids)^T as a
// tcgen05 GEMM whose epilogue keeps, per query token, only a CB_TOPR-entry shortlist, so the
// (nq*T) x K score matrix of `cells = Q' * centroids` (src/search/ranking.jl:27) is never
// materialised (8.6 GB at K = 2^16, 34 GB at K = 2^18 for a 1024-query batch).
//
// This kernel is the FAST APPROXIMATE pass of the decision rule in stage1.cu (fp16 operands, fp32
// accumulate in TMEM): k_stage1_rescore then re-scores the shortlist in exact fixed-order fp32 and
// flags any token whose decision is not separated from the best dropped centroid by more than the
// fp16 rounding bound (`guard`), so the chosen cells are those of the exact fp32 `_topk`
// (src/utils.jl:327-332) regardless of tensor-core rounding.
//
// Design:
//   * operands are pre-swizzled fp16 images ("row images": [rows/8][2 K-blocks][8 rows][128 B],
//     canonical SWIZZLE_128B K-major with a 2048-byte stride between 8-row groups), so any run of
//     rows is one contiguous block and moves with plain 1-D bulk copies (TMA engine, UBLKCP);
//     the centroid image is built once per index, the query image once per batch (and is the same
//     image the scoring kernel of stage 3+4 consumes);
//   * a persistent CTA owns work units = (256 query-token rows) x (one centroid range); the 64 KB
//     A block stays in shared memory, 32 KB centroid tiles (128 centroids) stream through a
//     4-stage mbarrier ring; each tile feeds TWO M=128 x N=128 accumulators, which halves the
//     L2 -> SMEM bytes per flop (one M=128 accumulator per tile would need 64 B/clk/SM, above
//     the ~42 B/clk/SM the L2 sustains chip-wide);
//   * TMEM: 2 stages x 2 accumulators x 128 fp32 columns = all 512 columns, so the MMA of tile
//     i+1 overlaps the epilogue of tile i;
//   * warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected thread), warps
//     4-11 = epilogue (thread = one query-token row: TMEM -> registers, a chunk maximum against
//     the row's current threshold, rare sorted inserts into its shared-memory shortlist).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int S1T_THREADS = 384;
constexpr int S1T_ROWS = 256;                 // query-token rows per work unit (two accumulators)
constexpr int S1T_BN = 128;                   // centroids per tile
constexpr int S1T_NB = 4;                     // centroid-tile stages
constexpr int S1T_A_BYTES = S1T_ROWS * 256;   // 64 KB
constexpr int S1T_B_BYTES = S1T_BN * 256;     // 32 KB
constexpr uint32_t S1T_TMEM_COLS = 512;

struct S1Barriers {
  uint64_t a_full, a_empty;
  uint64_t b_full[S1T_NB], b_empty[S1T_NB];
  uint64_t d_full[2], d_empty[2];
};

struct S1Params {
  const uint8_t* qimg;   // row image of the query tokens, rows padded to a multiple of 256
  const uint8_t* cimg;   // row image of the centroids, rows padded to a multiple of 128
  int64_t nrows, K;
  int n_rowblocks, nsplit, tiles_total, tiles_per_split;
  float* topv; int32_t* topi;   // [nrows][nsplit][CB_TOPR]
  uint32_t* thr_global;         // [rows_pad] orderable(best known 16th-best approximate score of the row), 0 = none yet
  float* thr0;                  // [nrows][nsplit] threshold the unit STARTED from (bounds what it dropped unseen)
};

// fp32 rows -> fp16 row image.  One thread per (row, 16-byte chunk); rows >= nrows are zero.
__global__ void k_tc_prep_rows(const float* __restrict__ X, int64_t nrows, int64_t nrows_pad, uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows_pad * 16) return;
  const int c16 = (int)(i & 15);
  const int64_t row = i >> 4;
  __align__(16) __half h[8];
  if (row < nrows) {
    const float4 a = *reinterpret_cast<const float4*>(X + row * 128 + c16 * 8);
    const float4 b = *reinterpret_cast<const float4*>(X + row * 128 + c16 * 8 + 4);
    h[0] = __float2half_rn(a.x); h[1] = __float2half_rn(a.y); h[2] = __float2half_rn(a.z); h[3] = __float2half_rn(a.w);
    h[4] = __float2half_rn(b.x); h[5] = __float2half_rn(b.y); h[6] = __float2half_rn(b.z); h[7] = __float2half_rn(b.w);
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) h[j] = __float2half_rn(0.f);
  }
  const int kb = c16 >> 3, chunk = c16 & 7, r7 = (int)(row & 7);
  uint8_t* dst = out + (row >> 3) * 2048 + kb * 1024 + r7 * 128 + ((chunk ^ r7) << 4);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

// sorted (descending) insert into the row's shortlist; lists are stored [slot][row] so that the 32
// rows of a warp hit 32 different banks.  Equal scores keep the earlier (lower) centroid id first.
__device__ __noinline__ float s1_insert(float* sv, int32_t* si, float v, int32_t cid) {
  int p = CB_TOPR - 1;
  while (p > 0) {
    const float up = sv[(p - 1) * S1T_ROWS];
    if (!(up < v)) break;
    sv[p * S1T_ROWS] = up;
    si[p * S1T_ROWS] = si[(p - 1) * S1T_ROWS];
    p--;
  }
  sv[p * S1T_ROWS] = v;
  si[p * S1T_ROWS] = cid;
  return sv[(CB_TOPR - 1) * S1T_ROWS];
}

__global__ void __launch_bounds__(S1T_THREADS, 1)
k_stage1_tc(S1Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_tile = smem;                                   // 64 KB
  uint8_t* b_tile0 = smem + S1T_A_BYTES;                    // S1T_NB x 32 KB
  float* s_val = reinterpret_cast<float*>(b_tile0 + S1T_NB * S1T_B_BYTES);   // [CB_TOPR][256]
  int32_t* s_idx = reinterpret_cast<int32_t*>(s_val + CB_TOPR * S1T_ROWS);   // [CB_TOPR][256]
  S1Barriers* bar = reinterpret_cast<S1Barriers*>(s_idx + CB_TOPR * S1T_ROWS);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);

  // warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops on the
  // uniform datapath (no WARPSYNC before the TMEM loads, no divergence checks)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    ptx::mbar_init(&bar->a_full, 1); ptx::mbar_init(&bar->a_empty, 1);
    for (int i = 0; i < S1T_NB; i++) { ptx::mbar_init(&bar->b_full[i], 1); ptx::mbar_init(&bar->b_empty[i], 1); }
    for (int i = 0; i < 2; i++) { ptx::mbar_init(&bar->d_full[i], 1); ptx::mbar_init(&bar->d_empty[i], 8); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(s_tmem, S1T_TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int n_units = P.n_rowblocks * P.nsplit;

  if (warp == 0) {
    // ===== producer: A block once per unit, centroid tiles through the ring =====
    uint32_t it = 0;
    int un = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, un++) {
      const int rb = u % P.n_rowblocks, split = u / P.n_rowblocks;
      const int t0 = split * P.tiles_per_split, t1 = min(P.tiles_total, t0 + P.tiles_per_split);
      ptx::mbar_wait(&bar->a_empty, (un & 1) ^ 1, 20);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&bar->a_full, S1T_A_BYTES);
        const uint8_t* src = P.qimg + (size_t)rb * S1T_A_BYTES;
#pragma unroll
        for (int j = 0; j < 4; j++) ptx::bulk_g2s(a_tile + j * 16384, src + j * 16384, 16384, &bar->a_full);
      }
      __syncwarp();
      for (int t = t0; t < t1; t++, it++) {
        const int st = it % S1T_NB;
        ptx::mbar_wait(&bar->b_empty[st], ((it / S1T_NB) & 1) ^ 1, 21);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&bar->b_full[st], S1T_B_BYTES);
          const uint8_t* src = P.cimg + (size_t)t * S1T_B_BYTES;
          uint8_t* dst = b_tile0 + (size_t)st * S1T_B_BYTES;
          ptx::bulk_g2s(dst, src, 16384, &bar->b_full[st]);
          ptx::bulk_g2s(dst + 16384, src + 16384, 16384, &bar->b_full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = ptx::idesc_f16(128, S1T_BN, 0);
    const uint32_t a_addr = ptx::smem_u32(a_tile), b_addr0 = ptx::smem_u32(b_tile0);
    uint32_t it = 0;
    int un = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, un++) {
      const int split = u / P.n_rowblocks;
      const int t0 = split * P.tiles_per_split, t1 = min(P.tiles_total, t0 + P.tiles_per_split);
      ptx::mbar_wait(&bar->a_full, un & 1, 22);
      for (int t = t0; t < t1; t++, it++) {
        const int st = it % S1T_NB, ds = it & 1;
        ptx::mbar_wait(&bar->b_full[st], (it / S1T_NB) & 1, 23);
        ptx::mbar_wait(&bar->d_empty[ds], ((it >> 1) & 1) ^ 1, 24);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t db0 = ptx::smem_desc_k_sw128(b_addr0 + st * S1T_B_BYTES, 2048);
#pragma unroll
          for (int a = 0; a < 2; a++) {
            const uint64_t da0 = ptx::smem_desc_k_sw128(a_addr + a * (S1T_A_BYTES / 2), 2048);
            const uint32_t d_tmem = tmem_base + ds * 256 + a * 128;
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const uint64_t koff = (uint64_t)(((k >> 2) * 1024 + (k & 3) * 32) >> 4);
              ptx::mma_f16_ss(d_tmem, da0 + koff, db0 + koff, idesc, k > 0 ? 1u : 0u);
            }
          }
          ptx::tc_commit(&bar->b_empty[st]);
          ptx::tc_commit(&bar->d_full[ds]);
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::tc_commit(&bar->a_empty);   // the A block may be overwritten once these MMAs retire
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue: one thread = one query-token row of one accumulator =====
    const int a = (warp - 4) >> 2, quarter = warp & 3;
    const int r = a * 128 + quarter * 32 + lane;        // row inside the unit
    float* sv = s_val + r;
    int32_t* si = s_idx + r;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int rb = u % P.n_rowblocks, split = u / P.n_rowblocks;
      const int t0 = split * P.tiles_per_split, t1 = min(P.tiles_total, t0 + P.tiles_per_split);
#pragma unroll
      for (int j = 0; j < CB_TOPR; j++) { sv[j * S1T_ROWS] = -INFINITY; si[j * S1T_ROWS] = 0x7fffffff; }
      // Start from the best 16th-best score any finished unit of this row has published: a centroid
      // below it is below >= 16 others, so units of later waves insert almost nothing (without it
      // every unit re-fills its list from -inf and nearly every 32-column piece takes the slow path).
      const int64_t grow0 = (int64_t)rb * S1T_ROWS + r;
      const uint32_t g0 = P.thr_global[grow0];
      const float thr0 = g0 ? cb_unorderable(g0) : -INFINITY;
      float thr = thr0;
      for (int t = t0; t < t1; t++, it++) {
        const int ds = it & 1;
        ptx::mbar_wait(&bar->d_full[ds], (it >> 1) & 1, 25);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ds * 256 + a * 128 + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < S1T_BN / 32; c++) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
          ptx::tmem_ld_wait();
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
          }
          if (fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) > thr) {
            const int cbase = t * S1T_BN + c * 32;
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const float x = __uint_as_float(v[i]);
              if (x > thr && (int64_t)(cbase + i) < P.K) thr = fmaxf(thr0, s1_insert(sv, si, x, cbase + i));
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->d_empty[ds]);
      }
      const int64_t grow = (int64_t)rb * S1T_ROWS + r;
      if (sv[(CB_TOPR - 1) * S1T_ROWS] > thr0) atomicMax(&P.thr_global[grow], cb_orderable(sv[(CB_TOPR - 1) * S1T_ROWS]));
      if (grow < P.nrows) {
        P.thr0[grow * P.nsplit + split] = thr0;
        float* ov = P.topv + (grow * P.nsplit + split) * CB_TOPR;
        int32_t* oi = P.topi + (grow * P.nsplit + split) * CB_TOPR;
#pragma unroll
        for (int j = 0; j < CB_TOPR; j++) { ov[j] = sv[j * S1T_ROWS]; oi[j] = si[j * S1T_ROWS]; }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, S1T_TMEM_COLS);
  }
}

constexpr size_t S1T_SMEM = 1024 + S1T_A_BYTES + (size_t)S1T_NB * S1T_B_BYTES + (size_t)2 * CB_TOPR * S1T_ROWS * 4 +
                            sizeof(S1Barriers) + 64;

}  // namespace

int32_t cb_tc_prep_rows(const float* dX, int64_t nrows, int64_t nrows_pad, uint8_t* d_out, cudaStream_t st) {
  if (nrows_pad == 0) return CB_OK;
  const int64_t total = nrows_pad * 16;
  k_tc_prep_rows<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, nrows, nrows_pad, d_out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int32_t cb_stage1_tc_shortlist(cb_index* ix, const float* dQ, int64_t nrows, float* topv, int32_t* topi, float* thr0,
                               int* nsplit_out, float* guard_rel_out, cudaStream_t st) {
  if (ix->dim != 128 || ix->centroids_img == nullptr) return CB_ERR_UNSUPPORTED;
  static_assert(S1T_SMEM <= 232448, "stage-1 tcgen05 kernel shared memory does not fit");
  // query image (shared with the stage 3+4 scoring kernel when T = 32)
  const int64_t rows_pad = (nrows + S1T_ROWS - 1) / S1T_ROWS * S1T_ROWS;
  CB_TRY(ix->q_prep.ensure((size_t)rows_pad * 256));
  CB_TRY(cb_tc_prep_rows(dQ, nrows, rows_pad, ix->q_prep.as<uint8_t>(), st));
  ix->q_prep_src = dQ;
  ix->q_prep_rows = nrows;

  S1Params P{};
  P.qimg = ix->q_prep.as<uint8_t>();
  P.cimg = ix->centroids_img;
  P.nrows = nrows; P.K = ix->K;
  P.n_rowblocks = (int)(rows_pad / S1T_ROWS);
  P.tiles_total = (int)((ix->K + S1T_BN - 1) / S1T_BN);
  // enough units to fill the SMs ~4 times over, but centroid ranges as long as possible: the
  // number of shortlist inserts per row grows with the number of ranges, not with K
  int nsplit = (4 * ix->sm_count + P.n_rowblocks - 1) / P.n_rowblocks;
  if (nsplit > CB_S1_SPLITS) nsplit = CB_S1_SPLITS;
  if (nsplit > P.tiles_total) nsplit = P.tiles_total;
  if (nsplit < 1) nsplit = 1;
  P.tiles_per_split = (P.tiles_total + nsplit - 1) / nsplit;
  nsplit = (P.tiles_total + P.tiles_per_split - 1) / P.tiles_per_split;   // drop empty ranges
  P.nsplit = nsplit;
  P.topv = topv; P.topi = topi; P.thr0 = thr0;
  CB_TRY(ix->s1_thr.ensure(sizeof(uint32_t) * (size_t)rows_pad));
  CB_CUDA(cudaMemsetAsync(ix->s1_thr.p, 0, sizeof(uint32_t) * (size_t)rows_pad, st));
  P.thr_global = ix->s1_thr.as<uint32_t>();
  const int n_units = P.n_rowblocks * nsplit;
  const int grid = n_units < ix->sm_count ? n_units : ix->sm_count;
  CB_CUDA(cudaFuncSetAttribute(k_stage1_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S1T_SMEM));
  k_stage1_tc<<<grid, S1T_THREADS, S1T_SMEM, st>>>(P);
  CB_LAUNCH_CHECK();
  *nsplit_out = nsplit;
  // fp16 rounding of both operands: |q~.c~ - q.c| <= (2^-10 + 2^-22) |q| |c|; the rescore kernel
  // scales this by the row's own norm.  Subnormal inputs and fp32 accumulation are covered by the
  // absolute 1e-5 the caller adds.
  *guard_rel_out = 1.05f * 9.765625e-4f * ix->centroid_norm_max;
  return CB_OK;
}

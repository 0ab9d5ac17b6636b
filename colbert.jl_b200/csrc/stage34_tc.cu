// stage34_tc.cu -- the hot kernel: residual decompression fused into MaxSim on the 5th-generation
// tensor cores (tcgen05 + TMEM), passage-major over a whole query batch.
//
// What it replaces, per (query q, candidate passage p) pair, with nothing materialised:
//   `_collect_compressed_embs_for_pids` (src/search/ranking.jl:46-67), `decompress`
//   (src/indexing/codecs/residual.jl:759-784), the `Q' * D` sgemm and the per-pid
//   `sum(maximum(..., dims = 2))` loop of `maxsim` (ranking.jl:76-84).
//
// Design (B200-first, not a translation):
//   * Passage-major.  A persistent CTA takes one passage at a time, decompresses it ONCE into a
//     shared-memory operand tile (fp16, UMMA K-major SWIZZLE_128B layout) and scores it against
//     every query of the batch that holds it as a candidate (its 1024-bit bitmap row).  The packed
//     index therefore streams from HBM once per batch instead of once per (query, passage) pair,
//     and decompression cost is amortised over ~17-113 queries.
//   * MMA orientation: A = 4 queries x 32 tokens (M = 128 rows, fp16, pre-swizzled once per batch
//     and fetched from L2 by 1-D bulk async copies), B = passage tokens (N = doclen padded to 16,
//     <= 240), K = dim = 128.  The fp32 accumulator D[128 x N] lives in TMEM; each epilogue
//     thread owns one (query, token) row, so "max over document tokens" is an in-register max over
//     its TMEM columns and "sum over query tokens" is one warp reduction.  Padded token rows
//     duplicate the last real token, so no column masking is needed.
//   * Warp-specialised, mbarrier-pipelined: warp 0 = scheduler + query-tile loader, warp 1 = MMA
//     issuer (one thread), warps 4-7 = epilogue (one TMEM lane quarter each), warps 8-15 =
//     decompression.  Three pipelines: passage tiles (2 buffers), query tiles (3-5 stages), TMEM
//     accumulators (2 x 256 columns).
// Requires dim = 128, T = 32, nbits in {1, 2, 4}; passages longer than the tile (doclen > 240) and
// every other shape are scored by the generic kernel (stage34_generic.cu).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int TC_THREADS = 512;
constexpr int TC_DIM = 128, TC_T = 32;
constexpr int TC_MAX_BROWS = 240;      // passage-tile rows (tokens); multiple of 16
constexpr int TC_MAX_ASTAGES = 6;
constexpr int TC_A_BYTES = 128 * TC_DIM * 2;  // 32 KB: 4 queries x 32 tokens x 128 x fp16
constexpr int TC_Q_BYTES = TC_T * TC_DIM * 2; // 8 KB per query (two 4 KB K-blocks)
constexpr int TC_NDEC_WARPS = 8;
constexpr uint32_t TC_TMEM_COLS = 512, TC_D_COLS = 256;

struct Meta {            // per passage-slot, written by warp 0
  int ncand;             // candidate queries of the passage scored by this kernel (0 = skip)
  int L;                 // doclen
  int npad;              // L padded to a multiple of 16
  int pid;               // local 0-based pid
  uint16_t q[CB_NQ_CHUNK];
};

struct Barriers {
  uint64_t b_full[2], b_empty[2], meta_full[2], meta_empty[2];
  uint64_t a_full[TC_MAX_ASTAGES], a_empty[TC_MAX_ASTAGES];
  uint64_t d_full[2], d_empty[2];
};

struct TcParams {
  const __half* centroids_h; const float* weights; const int32_t* codes; const uint8_t* residuals;
  const int64_t* offsets; int64_t Np; int nbits, R, W, brows, nastages;
  const uint8_t* qprep;           // [nq][8 KB] swizzled fp16 query tiles
  const uint32_t* bitmap; const int64_t* list_off; int32_t* cursors; uint64_t* pairs;
};

// Q fp32 [nq][32][128] -> fp16 in the SWIZZLE_128B K-major shared-memory image of one query:
// [2 K-blocks][32 rows][128 B], 16-byte chunk c of row t stored at chunk c ^ (t & 7).
__global__ void k_tc_prep_queries(const float* __restrict__ Q, uint8_t* __restrict__ out, int nq) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (q, t, 8 dims)
  const int64_t total = (int64_t)nq * TC_T * (TC_DIM / 8);
  if (i >= total) return;
  const int c16 = (int)(i % (TC_DIM / 8));   // 16-byte chunk along K: 0..15
  const int t = (int)((i / (TC_DIM / 8)) % TC_T);
  const int64_t q = i / ((TC_DIM / 8) * TC_T);
  const float* src = Q + (q * TC_T + t) * TC_DIM + c16 * 8;
  __align__(16) __half h[8];
#pragma unroll
  for (int j = 0; j < 8; j++) h[j] = __float2half_rn(src[j]);
  const int kb = c16 >> 3, chunk = c16 & 7;
  uint8_t* dst = out + q * TC_Q_BYTES + kb * (TC_T * 128) + ptx::sw128_offset(t, chunk);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

template <int NBITS>
__device__ __forceinline__ void decompress_token(const TcParams& P, const float* s_w, int64_t g, int lane,
                                                 uint8_t* tile, int row, int kb_stride, bool dup_row, int row2) {
  // lane owns dims 4*lane .. 4*lane+3
  const int32_t code = P.codes[g];
  const uint8_t* __restrict__ emb = P.residuals + g * P.R;
  uint32_t bits;
  if (NBITS == 2) bits = emb[lane];
  else if (NBITS == 4) bits = reinterpret_cast<const uint16_t*>(emb)[lane];
  else bits = (emb[lane >> 1] >> ((lane & 1) * 4)) & 0xfu;
  const uint2 craw = *reinterpret_cast<const uint2*>(P.centroids_h + (int64_t)code * TC_DIM + lane * 4);
  const __half2 c01 = *reinterpret_cast<const __half2*>(&craw.x), c23 = *reinterpret_cast<const __half2*>(&craw.y);
  float v[4] = {__low2float(c01), __high2float(c01), __low2float(c23), __high2float(c23)};
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    v[j] += s_w[(bits >> (j * NBITS)) & ((1u << NBITS) - 1u)];
    ss = fmaf(v[j], v[j], ss);
  }
  ss = cb_warp_sum(ss);
  const float inv = 1.0f / (sqrtf(ss) + 1.1920929e-07f);   // X ./ (norm + eps)
  const __half2 o01 = __floats2half2_rn(v[0] * inv, v[1] * inv), o23 = __floats2half2_rn(v[2] * inv, v[3] * inv);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&o01);
  o.y = *reinterpret_cast<const uint32_t*>(&o23);
  const int kb = lane >> 4, chunk = (lane >> 1) & 7, half8 = (lane & 1) * 8;
  *reinterpret_cast<uint2*>(tile + kb * kb_stride + ptx::sw128_offset(row, chunk) + half8) = o;
  if (dup_row)
    for (int r = row + 1; r < row2; r++)
      *reinterpret_cast<uint2*>(tile + kb * kb_stride + ptx::sw128_offset(r, chunk) + half8) = o;
}

template <int NBITS>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_maxsim_tc(TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment: align the dynamic window by hand
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int brows = P.brows, NA = P.nastages;
  const int kb_stride_b = brows * 128;                // bytes between the two K-blocks of a passage tile
  const int b_bytes = 2 * kb_stride_b;
  uint8_t* b_tile[2] = {smem, smem + b_bytes};
  uint8_t* a_tile0 = smem + 2 * b_bytes;              // NA stages of 32 KB
  Meta* meta = reinterpret_cast<Meta*>(a_tile0 + (size_t)NA * TC_A_BYTES);   // [2]
  Barriers* bar = reinterpret_cast<Barriers*>(meta + 2);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
  float* s_w = reinterpret_cast<float*>(s_tmem + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; i++) {
      ptx::mbar_init(&bar->b_full[i], TC_NDEC_WARPS); ptx::mbar_init(&bar->b_empty[i], 1);
      ptx::mbar_init(&bar->meta_full[i], 1);          ptx::mbar_init(&bar->meta_empty[i], 4);
      ptx::mbar_init(&bar->d_full[i], 1);             ptx::mbar_init(&bar->d_empty[i], 4);
    }
    for (int i = 0; i < TC_MAX_ASTAGES; i++) { ptx::mbar_init(&bar->a_full[i], 1); ptx::mbar_init(&bar->a_empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (tid < (1 << NBITS)) s_w[tid] = P.weights[tid];
  if (warp == 1) ptx::tmem_alloc(s_tmem, TC_TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int64_t first = blockIdx.x, stride = gridDim.x;

  if (warp == 0) {
    // ===== scheduler + query-tile loader =====
    uint32_t u = 0;
    int s = 0;
    for (int64_t p = first; p < P.Np; p += stride, s++) {
      const int slot = s & 1;
      ptx::mbar_wait(&bar->meta_empty[slot], ((s >> 1) & 1) ^ 1, 1);
      Meta& m = meta[slot];
      const int L = (int)(P.offsets[p + 1] - P.offsets[p]);
      uint32_t w = (lane < P.W) ? P.bitmap[p * P.W + lane] : 0u;
      if (L > brows || L == 0) w = 0u;     // long passages go to the generic kernel
      const int c = __popc(w);
      int pre = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += v; }
      int base = pre - c;
      while (w) { const int b = __ffs(w) - 1; w &= w - 1; m.q[base++] = (uint16_t)(lane * 32 + b); }
      const int ncand = __shfl_sync(0xffffffffu, pre, 31);
      if (lane == 0) { m.ncand = ncand; m.L = L; m.npad = (L + 15) & ~15; m.pid = (int)p; }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->meta_full[slot]);
      const int ngroups = (ncand + 3) >> 2;
      for (int g = 0; g < ngroups; g++, u++) {
        const int st = u % NA;
        if (lane == 0) {
          ptx::mbar_wait(&bar->a_empty[st], ((u / NA) & 1) ^ 1, 2);
          const int nqg = min(4, ncand - g * 4);
          ptx::mbar_arrive_expect_tx(&bar->a_full[st], (uint32_t)nqg * TC_Q_BYTES);
          uint8_t* dst = a_tile0 + (size_t)st * TC_A_BYTES;
          for (int j = 0; j < nqg; j++) {
            const uint8_t* src = P.qprep + (size_t)m.q[g * 4 + j] * TC_Q_BYTES;
            ptx::bulk_g2s(dst + j * 4096, src, 4096, &bar->a_full[st]);                  // K-block 0
            ptx::bulk_g2s(dst + 16384 + j * 4096, src + 4096, 4096, &bar->a_full[st]);   // K-block 1
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    uint32_t u = 0;
    int s = 0;
    for (int64_t p = first; p < P.Np; p += stride, s++) {
      const int slot = s & 1;
      const uint32_t ph = (s >> 1) & 1;
      ptx::mbar_wait(&bar->meta_full[slot], ph, 3);
      const int ncand = meta[slot].ncand, npad = meta[slot].npad;
      ptx::mbar_wait(&bar->b_full[slot], ph, 4);
      ptx::tc_fence_after();
      const int ngroups = (ncand + 3) >> 2;
      const uint32_t idesc = ptx::idesc_f16(128, npad > 0 ? npad : 16, 0);
      const uint32_t b_addr = ptx::smem_u32(b_tile[slot]);
      for (int g = 0; g < ngroups; g++, u++) {
        const int st = u % NA, ds = u & 1;
        ptx::mbar_wait(&bar->a_full[st], (u / NA) & 1, 5);
        ptx::mbar_wait(&bar->d_empty[ds], ((u >> 1) & 1) ^ 1, 6);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = ptx::smem_u32(a_tile0 + (size_t)st * TC_A_BYTES);
          const uint32_t d_tmem = tmem_base + ds * TC_D_COLS;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint64_t da = ptx::smem_desc_k_sw128(a_addr + (k >> 2) * 16384 + (k & 3) * 32);
            const uint64_t db = ptx::smem_desc_k_sw128(b_addr + (k >> 2) * kb_stride_b + (k & 3) * 32);
            ptx::mma_f16_ss(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
          }
          ptx::tc_commit(&bar->a_empty[st]);
          ptx::tc_commit(&bar->d_full[ds]);
        }
        __syncwarp();
      }
      if (lane == 0) {
        if (ngroups > 0) ptx::tc_commit(&bar->b_empty[slot]);   // after the passage's last MMA retires
        else ptx::mbar_arrive(&bar->b_empty[slot]);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue: TMEM -> max over tokens -> sum over query tokens -> pair list =====
    const int e = warp - 4;                 // TMEM lane quarter == query slot inside the group
    uint32_t u = 0;
    int s = 0;
    for (int64_t p = first; p < P.Np; p += stride, s++) {
      const int slot = s & 1;
      ptx::mbar_wait(&bar->meta_full[slot], (s >> 1) & 1, 7);
      const Meta& m = meta[slot];
      const int ncand = m.ncand, npad = m.npad, pid = m.pid;
      const int ngroups = (ncand + 3) >> 2;
      for (int g = 0; g < ngroups; g++, u++) {
        const int ds = u & 1;
        ptx::mbar_wait(&bar->d_full[ds], (u >> 1) & 1, 8);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ds * TC_D_COLS + ((uint32_t)(e * 32) << 16);
        float mx = -INFINITY;
        int c0 = 0;
        for (; c0 + 32 <= npad; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        if (c0 < npad) {
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(taddr + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i++) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->d_empty[ds]);
        const int qi = g * 4 + e;
        const float score = cb_warp_sum(mx);
        if (qi < ncand && lane == 0) {
          const int q = m.q[qi];
          const int pos = atomicAdd(&P.cursors[q], 1);
          P.pairs[P.list_off[q] + pos] = cb_pair_key(score, (uint32_t)pid);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->meta_empty[slot]);
    }
  } else if (warp >= 8) {
    // ===== decompression: packed codes/residuals -> normalised fp16 passage tile =====
    const int dw = warp - 8;
    int s = 0;
    for (int64_t p = first; p < P.Np; p += stride, s++) {
      const int slot = s & 1;
      ptx::mbar_wait(&bar->b_empty[slot], ((s >> 1) & 1) ^ 1, 9);
      const int64_t e0 = P.offsets[p];
      const int L = (int)(P.offsets[p + 1] - e0);
      if (L > 0 && L <= brows) {
        // skip passages no query of the batch wants
        const uint32_t wv = (lane < P.W) ? P.bitmap[p * P.W + lane] : 0u;
        if (__any_sync(0xffffffffu, wv != 0u)) {
          const int npad = (L + 15) & ~15;
          for (int e = dw; e < L; e += TC_NDEC_WARPS)
            decompress_token<NBITS>(P, s_w, e0 + e, lane, b_tile[slot], e, kb_stride_b, e == L - 1, npad);
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->b_full[slot]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

size_t tc_smem_bytes(int brows, int nastages) {
  return 1024 + (size_t)2 * 2 * brows * 128 + (size_t)nastages * TC_A_BYTES + 2 * sizeof(Meta) + sizeof(Barriers) + 16 + 64;
}

}  // namespace

bool cb_stage34_tc_supported(const cb_index* ix, int T) {
  return ix->dim == TC_DIM && T == TC_T && (ix->nbits == 1 || ix->nbits == 2 || ix->nbits == 4);
}

__global__ void k_collect_long_tc(const int64_t* __restrict__ offsets, int64_t Np, int64_t limit,
                                  int32_t* __restrict__ out /* [0] = count, [1..] pids */) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < Np && offsets[p + 1] - offsets[p] > limit) out[1 + atomicAdd(&out[0], 1)] = (int32_t)p;
}

int32_t cb_stage34_tc(cb_index* ix, const float* dQ, int nq, int T, int W, const uint32_t* d_bitmap,
                      const int64_t* d_list_off, int32_t* d_cursors, uint64_t* d_pairs, cudaStream_t st) {
  CB_REQUIRE(cb_stage34_tc_supported(ix, T), CB_ERR_UNSUPPORTED, "shape not supported by the tcgen05 scoring kernel");
  if (nq == 0 || ix->Np == 0) return CB_OK;
  // tile geometry: as many query-tile stages as fit beside two passage tiles
  int brows = (int)((ix->max_doclen + 15) & ~(int64_t)15);
  if (brows > TC_MAX_BROWS) brows = TC_MAX_BROWS;
  if (brows < 16) brows = 16;
  const size_t budget = 232448;  // 227 KB opt-in shared memory per CTA on sm_100
  int nast = TC_MAX_ASTAGES;
  while (nast > 2 && tc_smem_bytes(brows, nast) > budget) nast--;
  const size_t smem = tc_smem_bytes(brows, nast);
  CB_REQUIRE(smem <= budget, CB_ERR_UNSUPPORTED, "internal: tcgen05 kernel shared memory does not fit");

  // per-batch query tiles (fp16, swizzled)
  CB_TRY(ix->q_prep.ensure((size_t)nq * TC_Q_BYTES));
  {
    const int64_t total = (int64_t)nq * TC_T * (TC_DIM / 8);
    k_tc_prep_queries<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dQ, ix->q_prep.as<uint8_t>(), nq);
    CB_LAUNCH_CHECK();
  }
  TcParams P{};
  P.centroids_h = ix->centroids_h; P.weights = ix->weights; P.codes = ix->codes; P.residuals = ix->residuals;
  P.offsets = ix->offsets; P.Np = ix->Np; P.nbits = ix->nbits; P.R = ix->R; P.W = W; P.brows = brows; P.nastages = nast;
  P.qprep = ix->q_prep.as<uint8_t>(); P.bitmap = d_bitmap; P.list_off = d_list_off; P.cursors = d_cursors; P.pairs = d_pairs;
  int64_t grid = ix->sm_count;
  if (grid > ix->Np) grid = ix->Np;
#define CB_TC_LAUNCH(NB)                                                                                         \
  do {                                                                                                           \
    CB_CUDA(cudaFuncSetAttribute(k_maxsim_tc<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    k_maxsim_tc<NB><<<(unsigned)grid, TC_THREADS, smem, st>>>(P);                                                \
  } while (0)
  if (ix->nbits == 1) CB_TC_LAUNCH(1);
  else if (ix->nbits == 2) CB_TC_LAUNCH(2);
  else CB_TC_LAUNCH(4);
#undef CB_TC_LAUNCH
  CB_LAUNCH_CHECK();
  ix->st_tc_pairs = ix->st_pairs;

  // passages longer than the tile: generic kernel on just those
  if (ix->max_doclen > brows) {
    if (ix->n_long < 0 || ix->long_limit != brows) {   // the list is a property of the index: build once
      CB_TRY(ix->long_list.ensure(sizeof(int32_t) * (size_t)(ix->Np + 1)));
      int32_t* d_long = ix->long_list.as<int32_t>();
      CB_CUDA(cudaMemsetAsync(d_long, 0, sizeof(int32_t), st));
      k_collect_long_tc<<<(unsigned)((ix->Np + 255) / 256), 256, 0, st>>>(ix->offsets, ix->Np, brows, d_long);
      CB_LAUNCH_CHECK();
      int32_t* h = reinterpret_cast<int32_t*>(ix->pinned_total + 5);
      CB_CUDA(cudaMemcpyAsync(h, d_long, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CB_CUDA(cudaStreamSynchronize(st));
      ix->n_long = *h;
      ix->long_limit = brows;
    }
    if (ix->n_long > 0)
      CB_TRY(cb_stage34_generic(ix, dQ, nq, T, W, d_bitmap, ix->long_list.as<int32_t>() + 1, ix->n_long, d_list_off,
                                d_cursors, d_pairs, st));
  }
  return CB_OK;
}

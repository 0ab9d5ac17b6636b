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
//   * work = (256-row block, 128-centroid tile) pairs, flattened row-block major and cut into ONE
//     contiguous range per CTA (perfect balance, no tail wave): a CTA walks <= 2-3 "segments"
//     (row block, tile range); the 64 KB A block of a segment stays in shared memory, 32 KB centroid
//     tiles stream through a 3-stage mbarrier ring; each tile feeds TWO M=128 x N=128 accumulators
//     (one M=128 accumulator per tile would need 64 B/clk/SM of operand reads per flop pair);
//   * TMEM: 2 stages x 2 accumulators x 128 fp32 columns = all 512 columns, so the MMA of tile
//     i+1 overlaps the epilogue of tile i;
//   * warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected thread), warps
//     4-11 = epilogue, thread = one query-token row of one accumulator.
//   * epilogue = threshold filter + LAZY shortlist.  The row keeps a threshold (a valid lower bound
//     of its 16th-best score) in a register; TMEM loads are double-buffered (32 columns in flight
//     while 32 are folded); a 32-column piece whose maximum does not beat the threshold costs ~20
//     instructions.  Survivors are APPENDED (no sorted insert) to a 32-slot per-row buffer in shared
//     memory; when a row holds more than 24, its warp compacts it cooperatively -- the 32 lanes rank
//     the 32 slots against each other with shuffles, keep the best 16 (sorted) and raise the row's
//     threshold to the 16th -- ~130 convergent instructions every >= 9 appends, instead of a
//     divergent, latency-bound sorted insert per append (the previous epilogue: 5.9 M serialized
//     inserts per batch, 75 % of the kernel).  Thresholds are also exchanged through global memory
//     (atomicMax after a compaction, one load per tile), so the CTAs that share a row block tighten
//     each other's filters.
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int S1T_THREADS = 384;
constexpr int S1T_ROWS = 256;                 // query-token rows per segment (two accumulators)
constexpr int S1T_BN = 128;                   // centroids per tile
constexpr int S1T_NB = 3;                     // centroid-tile stages
constexpr int S1T_A_BYTES = S1T_ROWS * 256;   // 64 KB
constexpr int S1T_B_BYTES = S1T_BN * 256;     // 32 KB
constexpr int S1T_SLOTS = 32;                 // per-row append buffer (compacted to CB_TOPR when > S1T_HIGH)
constexpr int S1T_HIGH = 24;                  // <= S1T_SLOTS - 8: eight appends always fit
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
  int n_rowblocks, nsplit, tiles_total;
  long long per, total;         // tile-units per CTA / in all (flattened [row block][tile])
  float* topv; int32_t* topi;   // [nrows][nsplit][CB_TOPR], sorted descending; unused slots -inf / 0x7fffffff
  uint32_t* thr_global;         // [rows_pad] orderable(best known lower bound of the row's 16th-best score), 0 = none yet
  float* thr0;                  // [nrows][nsplit] final threshold of the segment: everything it saw and did not list is <= this
};

// fp32 rows -> fp16 row image.  One thread per (row, 16-byte chunk); rows >= nrows are zero.
// range_flag (optional): set to 1 when a row's squared norm exceeds 255^2 or is not finite -- the
// range precondition of the tcgen05 scoring kernel's fixed-point token sum (stage34_tc.cu).
__global__ void k_tc_prep_rows(const float* __restrict__ X, int64_t nrows, int64_t nrows_pad, uint8_t* __restrict__ out,
                               int* __restrict__ range_flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows_pad * 16) return;
  const int c16 = (int)(i & 15);
  const int64_t row = i >> 4;
  __align__(16) __half h[8];
  float ss = 0.f;
  if (row < nrows) {
    const float4 a = *reinterpret_cast<const float4*>(X + row * 128 + c16 * 8);
    const float4 b = *reinterpret_cast<const float4*>(X + row * 128 + c16 * 8 + 4);
    h[0] = __float2half_rn(a.x); h[1] = __float2half_rn(a.y); h[2] = __float2half_rn(a.z); h[3] = __float2half_rn(a.w);
    h[4] = __float2half_rn(b.x); h[5] = __float2half_rn(b.y); h[6] = __float2half_rn(b.z); h[7] = __float2half_rn(b.w);
    ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) h[j] = __float2half_rn(0.f);
  }
  if (range_flag != nullptr) {      // the 16 chunk-threads of a row are 16 consecutive lanes
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (c16 == 0 && !(ss <= 65025.0f)) atomicExch(range_flag, 1);
  }
  const int kb = c16 >> 3, chunk = c16 & 7, r7 = (int)(row & 7);
  uint8_t* dst = out + (row >> 3) * 2048 + kb * 1024 + r7 * 128 + ((chunk ^ r7) << 4);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

// One (row block, tile range) piece of this CTA's contiguous share of the flattened tile-units.
struct S1Seg { int rb, t0, t1, idx; };
__device__ __forceinline__ bool s1_next_seg(const S1Params& P, long long& f, long long f1, S1Seg& s) {
  if (f >= f1) return false;
  const long long T = P.tiles_total;
  s.rb = (int)(f / T);
  s.t0 = (int)(f - (long long)s.rb * T);
  const long long left = f1 - f;
  s.t1 = (int)((T - s.t0 < left) ? T : s.t0 + left);
  s.idx = (int)blockIdx.x - (int)(((long long)s.rb * T) / P.per);   // segments of a row block are numbered from its first CTA
  f += s.t1 - s.t0;
  return true;
}

// slot j of row r sits at r * 32 + ((j + r) & 31): a lane appending to ITS row and a warp reading ONE
// row's 32 slots are both bank-conflict free
__device__ __forceinline__ int s1_slot(int row, int j) { return row * S1T_SLOTS + ((j + row) & (S1T_SLOTS - 1)); }

// Warp-cooperative compaction of the rows named in `mask` (bit L = the row owned by lane L): keeps the
// best min(n, 16) entries, sorted descending in slots 0.., and raises that row's threshold to its 16th.
struct S1State { int cnt; float thr; };
__device__ __noinline__ S1State s1_compact(uint32_t mask, float* bv, int32_t* bi, int rowbase, int lane, int cnt, float thr,
                                           uint32_t* thr_global_warp /* &thr_global[grow of lane 0] */) {
  while (mask) {
    const int L = __ffs(mask) - 1;
    mask &= mask - 1;
    const int n = __shfl_sync(0xffffffffu, cnt, L);
    const int row = rowbase + L;
    __syncwarp();   // lane L's appends (plain shared-memory stores) become visible to the lanes that read its row below
    const int a = s1_slot(row, lane);
    const float v = lane < n ? bv[a] : -INFINITY;
    const int32_t id = lane < n ? bi[a] : 0x7fffffff;
    int rank = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
      const float vk = __shfl_sync(0xffffffffu, v, k);
      rank += (vk > v || (vk == v && k < lane)) ? 1 : 0;
    }
    __syncwarp();   // every lane has read its slot: the stores below may now reuse them
    if (rank < CB_TOPR) { const int wa = s1_slot(row, rank); bv[wa] = v; bi[wa] = id; }
    const uint32_t who = __ballot_sync(0xffffffffu, rank == CB_TOPR - 1);
    const float t16 = __shfl_sync(0xffffffffu, v, who ? __ffs(who) - 1 : 0);
    if (lane == L) {
      cnt = n < CB_TOPR ? n : CB_TOPR;
      if (n >= CB_TOPR && t16 > thr) {
        thr = t16;
        atomicMax(thr_global_warp + L, cb_orderable(t16));
      }
    }
    __syncwarp();
  }
  return S1State{cnt, thr};
}

// 32 accumulator columns of this thread's row: filter against the threshold, append the survivors.
__device__ __forceinline__ void s1_piece(const uint32_t (&v)[32], int cbase, int nvalid, float* bv, int32_t* bi, int row, int rowbase,
                                         int lane, int& cnt, float& thr, uint32_t* thr_global_warp) {
  float g[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float a = fmaxf(fmaxf(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])), __uint_as_float(v[8 * j + 2]));
    const float b = fmaxf(fmaxf(__uint_as_float(v[8 * j + 3]), __uint_as_float(v[8 * j + 4])), __uint_as_float(v[8 * j + 5]));
    g[j] = fmaxf(fmaxf(a, b), fmaxf(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
  }
  const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
  if (!__any_sync(0xffffffffu, m > thr)) return;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (!__any_sync(0xffffffffu, g[j] > thr)) continue;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float x = __uint_as_float(v[8 * j + i]);
      if (x > thr && 8 * j + i < nvalid) {
        const int a = s1_slot(row, cnt);
        bv[a] = x;
        bi[a] = cbase + 8 * j + i;
        cnt++;
      }
    }
    const uint32_t full = __ballot_sync(0xffffffffu, cnt > S1T_HIGH);
    if (full) { const S1State ns = s1_compact(full, bv, bi, rowbase, lane, cnt, thr, thr_global_warp); cnt = ns.cnt; thr = ns.thr; }
  }
}

__global__ void __launch_bounds__(S1T_THREADS, 1)
k_stage1_tc(S1Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_tile = smem;                                   // 64 KB
  uint8_t* b_tile0 = smem + S1T_A_BYTES;                    // S1T_NB x 32 KB
  float* s_val = reinterpret_cast<float*>(b_tile0 + S1T_NB * S1T_B_BYTES);     // [256][S1T_SLOTS]
  int32_t* s_idx = reinterpret_cast<int32_t*>(s_val + S1T_SLOTS * S1T_ROWS);   // [256][S1T_SLOTS]
  S1Barriers* bar = reinterpret_cast<S1Barriers*>(s_idx + S1T_SLOTS * S1T_ROWS);
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

  const long long f0 = (long long)blockIdx.x * P.per;
  const long long f1 = (f0 + P.per < P.total) ? f0 + P.per : P.total;

  if (warp == 0) {
    // ===== producer: A block once per segment, centroid tiles through the ring =====
    uint32_t it = 0;
    int un = 0;
    long long f = f0;
    S1Seg sg;
    for (; s1_next_seg(P, f, f1, sg); un++) {
      ptx::mbar_wait(&bar->a_empty, (un & 1) ^ 1, 20);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&bar->a_full, S1T_A_BYTES);
        const uint8_t* src = P.qimg + (size_t)sg.rb * S1T_A_BYTES;
#pragma unroll
        for (int j = 0; j < 4; j++) ptx::bulk_g2s(a_tile + j * 16384, src + j * 16384, 16384, &bar->a_full);
      }
      __syncwarp();
      for (int t = sg.t0; t < sg.t1; t++, it++) {
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
    long long f = f0;
    S1Seg sg;
    for (; s1_next_seg(P, f, f1, sg); un++) {
      ptx::mbar_wait(&bar->a_full, un & 1, 22);
      for (int t = sg.t0; t < sg.t1; t++, it++) {
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
    const int rowbase = a * 128 + quarter * 32;         // first row of this warp inside the segment
    const int r = rowbase + lane;
    uint32_t it = 0;
    long long f = f0;
    S1Seg sg;
    while (s1_next_seg(P, f, f1, sg)) {
      const int64_t grow = (int64_t)sg.rb * S1T_ROWS + r;
      uint32_t* thr_global_warp = P.thr_global + ((int64_t)sg.rb * S1T_ROWS + rowbase);
      int cnt = 0;
      float thr = -INFINITY;
      for (int t = sg.t0; t < sg.t1; t++, it++) {
        const int ds = it & 1;
        const uint32_t g0 = __ldcg(P.thr_global + grow);       // what any CTA working on this row has published
        ptx::mbar_wait(&bar->d_full[ds], (it >> 1) & 1, 25);
        ptx::tc_fence_after();
        if (g0) thr = fmaxf(thr, cb_unorderable(g0));
        const uint32_t taddr = tmem_base + ds * 256 + a * 128 + ((uint32_t)(quarter * 32) << 16);
        const int cbase = t * S1T_BN;
        const int nvalid = (int)((P.K - cbase < S1T_BN) ? P.K - cbase : S1T_BN);
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32b_x32(taddr, va);
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32b_x32(taddr + 32, vb);
        s1_piece(va, cbase, nvalid, s_val, s_idx, r, rowbase, lane, cnt, thr, thr_global_warp);
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32b_x32(taddr + 64, va);
        s1_piece(vb, cbase + 32, nvalid - 32, s_val, s_idx, r, rowbase, lane, cnt, thr, thr_global_warp);
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32b_x32(taddr + 96, vb);
        s1_piece(va, cbase + 64, nvalid - 64, s_val, s_idx, r, rowbase, lane, cnt, thr, thr_global_warp);
        ptx::tmem_ld_wait();
        // every column of this accumulator is in registers: hand it back to the MMA issuer
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->d_empty[ds]);
        s1_piece(vb, cbase + 96, nvalid - 96, s_val, s_idx, r, rowbase, lane, cnt, thr, thr_global_warp);
      }
      // final compaction sorts every non-empty row; then the list and the segment's final threshold go out
      const uint32_t live = __ballot_sync(0xffffffffu, cnt > 0);
      if (live) { const S1State ns = s1_compact(live, s_val, s_idx, rowbase, lane, cnt, thr, thr_global_warp); cnt = ns.cnt; thr = ns.thr; }
      if (grow < P.nrows) {
        P.thr0[grow * P.nsplit + sg.idx] = thr;
        float* ov = P.topv + (grow * P.nsplit + sg.idx) * CB_TOPR;
        int32_t* oi = P.topi + (grow * P.nsplit + sg.idx) * CB_TOPR;
#pragma unroll
        for (int j = 0; j < CB_TOPR; j++) {
          const int sa = s1_slot(r, j);
          ov[j] = j < cnt ? s_val[sa] : -INFINITY;
          oi[j] = j < cnt ? s_idx[sa] : 0x7fffffff;
        }
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, S1T_TMEM_COLS);
  }
}

constexpr size_t S1T_SMEM = 1024 + S1T_A_BYTES + (size_t)S1T_NB * S1T_B_BYTES + (size_t)2 * S1T_SLOTS * S1T_ROWS * 4 +
                            sizeof(S1Barriers) + 64;

}  // namespace

int32_t cb_tc_prep_rows(const float* dX, int64_t nrows, int64_t nrows_pad, uint8_t* d_out, cudaStream_t st, int* d_range_flag) {
  if (nrows_pad == 0) return CB_OK;
  const int64_t total = nrows_pad * 16;
  k_tc_prep_rows<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dX, nrows, nrows_pad, d_out, d_range_flag);
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
  CB_TRY(ix->q_flag.ensure(sizeof(int)));
  CB_CUDA(cudaMemsetAsync(ix->q_flag.p, 0, sizeof(int), st));
  CB_TRY(cb_tc_prep_rows(dQ, nrows, rows_pad, ix->q_prep.as<uint8_t>(), st, ix->q_flag.as<int>()));
  ix->q_prep_src = dQ;
  ix->q_prep_rows = nrows;

  S1Params P{};
  P.qimg = ix->q_prep.as<uint8_t>();
  P.cimg = ix->centroids_img;
  P.nrows = nrows; P.K = ix->K;
  P.n_rowblocks = (int)(rows_pad / S1T_ROWS);
  P.tiles_total = (int)((ix->K + S1T_BN - 1) / S1T_BN);
  // One contiguous range of the flattened [row block][tile] space per CTA.  A row block must not be cut into
  // more than CB_S1_SPLITS segments (the exact re-score reads nsplit * CB_TOPR candidates per row), so a
  // range is never shorter than 1/14 of a row block: at most floor((T - 1) / per) + 2 <= 15 segments.
  P.total = (long long)P.n_rowblocks * P.tiles_total;
  long long per = (P.total + ix->sm_count - 1) / ix->sm_count;
  const long long per_min = (P.tiles_total + 13) / 14;
  if (per < per_min) per = per_min;
  P.per = per;
  const int grid = (int)((P.total + per - 1) / per);
  int nsplit = (int)((P.tiles_total - 1) / per) + 2;
  if (nsplit > grid) nsplit = grid;
  if (nsplit > CB_S1_SPLITS) return CB_ERR_UNSUPPORTED;   // cannot happen (per >= T / 14)
  P.nsplit = nsplit;
  P.topv = topv; P.topi = topi; P.thr0 = thr0;
  // segments that do not exist (a row block cut into fewer pieces than nsplit) read as empty lists
  CB_CUDA(cudaMemsetAsync(topi, 0xff, sizeof(int32_t) * (size_t)nrows * nsplit * CB_TOPR, st));   // id -1 = empty
  CB_CUDA(cudaMemsetAsync(thr0, 0xff, sizeof(float) * (size_t)nrows * nsplit, st));               // NaN = no bound
  CB_TRY(ix->s1_thr.ensure(sizeof(uint32_t) * (size_t)rows_pad));
  CB_CUDA(cudaMemsetAsync(ix->s1_thr.p, 0, sizeof(uint32_t) * (size_t)rows_pad, st));
  P.thr_global = ix->s1_thr.as<uint32_t>();
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

// stage2.cu -- IVF probe and candidate-pid dedup.  Replaces the integer half of `retrieve`
// (src/search/ranking.jl:32-42: sort(unique(cells)) -> `_cids_to_eids!` -> sort(unique(eids)) ->
// sort(unique(emb2pid[eids]))) with a bitmap: bit q of row `pid` is set iff passage `pid` owns an
// embedding in a cell probed by query q.  A set (bitmap) is order-free, so the three sort/unique
// passes disappear; ascending-pid order falls out of scanning the bitmap by row.
//
// Layout handed on: bitmap uint32[Np][W], W = ceil(nq/32): the row of one passage (all queries of
// the chunk) is one contiguous, cache-line-sized record -- what the passage-major scoring kernel
// reads.  The marking itself runs on the TRANSPOSED layout uint32[nq][NpW] (one row of Np bits per
// query): the ~64 CTAs of a query run back to back and hit only that query's 1.1 MB row, which
// stays in L2, whereas atomics scattered over the 1.1 GB passage-major bitmap made every one of
// them a 32-byte DRAM read-modify-write (25 GB of DRAM traffic per batch at 8.8 M passages,
// measured).  A bit-matrix transpose (ballots, staged through shared memory so that both sides
// move whole lines) then produces the passage-major rows.
#include "common.cuh"

// One CTA per (query, probe slot).  Slot s of query q is cell cells[q][s]; slots holding a cell
// that an earlier slot of the same query already holds are skipped (the reference's
// `unique(cells)`, ranking.jl:32).
__global__ void __launch_bounds__(256)
k_stage2_mark(const int32_t* __restrict__ cells, int slots, const int64_t* __restrict__ cell_offsets,
              const int32_t* __restrict__ ivf_pids, const int64_t* __restrict__ offsets, int64_t NpW,
              uint32_t* __restrict__ bitmapT, int32_t* __restrict__ counts,
              unsigned long long* __restrict__ pair_embs) {
  const int q = blockIdx.y;
  const int s = blockIdx.x;
  const int tid = threadIdx.x;
  const int32_t* qc = cells + (int64_t)q * slots;
  const int32_t cell = qc[s];
  if (cell < 0) return;  // padded slot (K < nprobe)
  int dup = 0;
  for (int j = tid; j < s; j += blockDim.x) dup |= (qc[j] == cell);
  if (__syncthreads_or(dup)) return;

  const int64_t b = cell_offsets[cell], e = cell_offsets[cell + 1];
  uint32_t* row = bitmapT + (int64_t)q * NpW;
  int fresh = 0;
  unsigned long long embs = 0;
  for (int64_t i = b + tid; i < e; i += blockDim.x) {
    const int32_t pid = ivf_pids[i];
    const uint32_t bit = 1u << (pid & 31);
    const uint32_t old = atomicOr(&row[pid >> 5], bit);
    if (!(old & bit)) {
      fresh++;
      embs += (unsigned long long)(offsets[pid + 1] - offsets[pid]);
    }
  }
  // block reduce -> one atomic per CTA
  __shared__ int s_fresh[8];
  __shared__ unsigned long long s_embs[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    embs += __shfl_xor_sync(0xffffffffu, embs, o);
  }
  if ((tid & 31) == 0) { s_fresh[tid >> 5] = fresh; s_embs[tid >> 5] = embs; }
  __syncthreads();
  if (tid == 0) {
    int f = 0;
    unsigned long long m = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { f += s_fresh[w]; m += s_embs[w]; }
    if (f) { atomicAdd(&counts[q], f); atomicAdd(pair_embs, m); }
  }
}

// bitmapT uint32[nq][NpW] (bit p & 31 of word [q][p >> 5]) -> bitmap uint32[Np][W] (bit q & 31 of word
// [p][q >> 5]).  One CTA per 256 passages, one warp per 32 queries: lane = query reads 8 consecutive
// words of its row (one full 32-byte sector); each 32 x 32 bit block (32 queries x 32 passages) is
// transposed in registers by the five-step butterfly (exchange 16 x 16, 8 x 8, ... 1 x 1 sub-blocks with
// lane ^ 16, ^ 8, ... -- 5 shuffles per block where the ballot formulation needed 32 ballots and 32
// single-lane stores); the CTA's 256 x W output words are staged in shared memory (row stride 33: both
// the per-lane writes and the row-wise read-out are bank-conflict free) and leave as whole lines.
__device__ __forceinline__ uint32_t transpose32_across_warp(uint32_t x, int lane) {
  uint32_t y;
  y = __shfl_xor_sync(0xffffffffu, x, 16); x = (lane & 16) ? ((x & 0xffff0000u) | (y >> 16)) : ((x & 0x0000ffffu) | (y << 16));
  y = __shfl_xor_sync(0xffffffffu, x, 8);  x = (lane & 8) ? ((x & 0xff00ff00u) | ((y >> 8) & 0x00ff00ffu)) : ((x & 0x00ff00ffu) | ((y & 0x00ff00ffu) << 8));
  y = __shfl_xor_sync(0xffffffffu, x, 4);  x = (lane & 4) ? ((x & 0xf0f0f0f0u) | ((y >> 4) & 0x0f0f0f0fu)) : ((x & 0x0f0f0f0fu) | ((y & 0x0f0f0f0fu) << 4));
  y = __shfl_xor_sync(0xffffffffu, x, 2);  x = (lane & 2) ? ((x & 0xccccccccu) | ((y >> 2) & 0x33333333u)) : ((x & 0x33333333u) | ((y & 0x33333333u) << 2));
  y = __shfl_xor_sync(0xffffffffu, x, 1);  x = (lane & 1) ? ((x & 0xaaaaaaaau) | ((y >> 1) & 0x55555555u)) : ((x & 0x55555555u) | ((y & 0x55555555u) << 1));
  return x;   // lane p now holds: bit q = bit p of lane q's input
}

__global__ void __launch_bounds__(1024)
k_bitmap_transpose(const uint32_t* __restrict__ bitmapT, int nq, int64_t NpW, int64_t Np, int W, uint32_t* __restrict__ bitmap) {
  __shared__ uint32_t s_out[256 * 33];
  const int qw = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8;            // first of this CTA's 8 words per query row
  const int q = qw * 32 + lane;
  uint32_t x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = 0u;
  if (q < nq && w0 < NpW) {                               // NpW is a multiple of 8: the 8 words are in range together
    const uint4* src = reinterpret_cast<const uint4*>(bitmapT + (int64_t)q * NpW + w0);
    const uint4 a = src[0], b = src[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  }
  if (qw < W) {                                           // (warps beyond the chunk's query words hold only zeros)
#pragma unroll
    for (int k = 0; k < 8; k++) s_out[(k * 32 + lane) * 33 + qw] = transpose32_across_warp(x[k], lane);
  }
  __syncthreads();
  const int64_t p0 = (int64_t)blockIdx.x * 256;
  for (int i = threadIdx.x; i < 256 * W; i += 1024) {
    const int pl = i / W, w = i % W;
    if (p0 + pl < Np) bitmap[(p0 + pl) * W + w] = s_out[pl * 33 + w];
  }
}

int32_t cb_stage2_mark(cb_index* ix, const int32_t* d_cells, int nq, int T, int nprobe, int W,
                       uint32_t* d_bitmap, int32_t* d_counts, cudaStream_t st) {
  if (nq == 0 || ix->Np == 0) return CB_OK;
  unsigned long long* d_embs = cb_stats_dev(ix) + 1;   // batch counter: pair embeddings
  const int64_t NpW = ((ix->Np + 31) / 32 + 7) / 8 * 8;   // words per query row, padded to whole 32-byte sectors
  CB_TRY(ix->bitmap_t.ensure(sizeof(uint32_t) * (size_t)nq * (size_t)NpW));
  CB_CUDA(cudaMemsetAsync(ix->bitmap_t.p, 0, sizeof(uint32_t) * (size_t)nq * (size_t)NpW, st));
  dim3 grid(T * nprobe, nq);
  k_stage2_mark<<<grid, 256, 0, st>>>(d_cells, T * nprobe, ix->cell_offsets, ix->ivf_pids, ix->offsets, NpW,
                                      ix->bitmap_t.as<uint32_t>(), d_counts, d_embs);
  CB_LAUNCH_CHECK();
  k_bitmap_transpose<<<(unsigned)((ix->Np + 255) / 256), 1024, 0, st>>>(ix->bitmap_t.as<uint32_t>(), nq, NpW, ix->Np, W, d_bitmap);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// Exclusive scan of the per-query candidate counts (nq <= CB_NQ_CHUNK): one CTA.
__global__ void __launch_bounds__(1024)
k_scan_counts(const int32_t* __restrict__ counts, int nq, int64_t* __restrict__ list_off, unsigned long long* __restrict__ stat_pairs) {
  __shared__ int64_t s[1024];
  int tid = threadIdx.x;
  int64_t v = tid < nq ? counts[tid] : 0;
  s[tid] = v;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int64_t add = tid >= o ? s[tid - o] : 0;
    __syncthreads();
    s[tid] += add;
    __syncthreads();
  }
  if (tid < nq) list_off[tid] = s[tid] - v;
  if (tid == 1023) {
    list_off[nq] = s[1023];
    if (stat_pairs != nullptr) atomicAdd(stat_pairs, (unsigned long long)s[1023]);
  }
}

int32_t cb_scan_counts(const int32_t* d_counts, int nq, int64_t* d_list_off, cudaStream_t st, unsigned long long* d_stat_pairs) {
  CB_REQUIRE(nq <= 1024, CB_ERR_BAD_ARG, "internal: scan chunk too large");
  k_scan_counts<<<1, 1024, 0, st>>>(d_counts, nq, d_list_off, d_stat_pairs);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

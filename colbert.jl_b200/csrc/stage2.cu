// stage2.cu -- IVF probe and candidate-pid dedup.  Replaces the integer half of `retrieve`
// (src/search/ranking.jl:32-42: sort(unique(cells)) -> `_cids_to_eids!` -> sort(unique(eids)) ->
// sort(unique(emb2pid[eids]))) with a bitmap: bit q of row `pid` is set iff passage `pid` owns an
// embedding in a cell probed by query q.  A set (bitmap) is order-free, so the three sort/unique
// passes disappear; ascending-pid order falls out of scanning the bitmap by row.
//
// Layout: bitmap uint32[Np][W], W = ceil(nq/32): the row of one passage (all queries of the
// chunk) is one contiguous, cache-line-sized record -- what the passage-major scoring kernel
// reads.  HBM-bound integer work: each IVF entry costs a coalesced 4-byte read plus one L2 atomic.
#include "common.cuh"

// One CTA per (query, probe slot).  Slot s of query q is cell cells[q][s]; slots holding a cell
// that an earlier slot of the same query already holds are skipped (the reference's
// `unique(cells)`, ranking.jl:32).
__global__ void __launch_bounds__(256)
k_stage2_mark(const int32_t* __restrict__ cells, int slots, const int64_t* __restrict__ cell_offsets,
              const int32_t* __restrict__ ivf_pids, const int64_t* __restrict__ offsets, int W,
              uint32_t* __restrict__ bitmap, int32_t* __restrict__ counts,
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
  const uint32_t bit = 1u << (q & 31);
  const int word = q >> 5;
  int fresh = 0;
  unsigned long long embs = 0;
  for (int64_t i = b + tid; i < e; i += blockDim.x) {
    const int32_t pid = ivf_pids[i];
    const uint32_t old = atomicOr(&bitmap[(int64_t)pid * W + word], bit);
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

int32_t cb_stage2_mark(cb_index* ix, const int32_t* d_cells, int nq, int T, int nprobe, int W,
                       uint32_t* d_bitmap, int32_t* d_counts, cudaStream_t st) {
  if (nq == 0 || ix->Np == 0) return CB_OK;
  CB_TRY(ix->misc.ensure(64));
  unsigned long long* d_embs = ix->misc.as<unsigned long long>();
  CB_CUDA(cudaMemsetAsync(d_embs, 0, sizeof(unsigned long long), st));
  dim3 grid(T * nprobe, nq);
  k_stage2_mark<<<grid, 256, 0, st>>>(d_cells, T * nprobe, ix->cell_offsets, ix->ivf_pids, ix->offsets, W,
                                      d_bitmap, d_counts, d_embs);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// Exclusive scan of the per-query candidate counts (nq <= CB_NQ_CHUNK): one CTA.
__global__ void __launch_bounds__(1024)
k_scan_counts(const int32_t* __restrict__ counts, int nq, int64_t* __restrict__ list_off) {
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
  if (tid == 1023) list_off[nq] = s[1023];
}

int32_t cb_scan_counts(const int32_t* d_counts, int nq, int64_t* d_list_off, cudaStream_t st) {
  CB_REQUIRE(nq <= 1024, CB_ERR_BAD_ARG, "internal: scan chunk too large");
  k_scan_counts<<<1, 1024, 0, st>>>(d_counts, nq, d_list_off);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// search.cu -- the pipeline driver behind the ABI: `search` minus the encoder
// (src/searching.jl:103-127), batched, plus the stage-level hooks that mirror `retrieve`
// (src/search/ranking.jl:23-44) and the fused collect/decompress/maxsim (ranking.jl:46-86).
#include "common.cuh"

int32_t cb_generic_score_list(cb_index* ix, const float* dQ, int nq, int T, const int32_t* d_pid_list,
                              int64_t n_list, float* d_out_scores, cudaStream_t st);  // stage34_generic.cu
// stage5.cu: first k of every list; with opt_exact_rescore the best 2k+ tensor-core candidates are re-scored in
// exact fp32 first and the final (score desc, pid asc) order is decided on those scores
int32_t cb_final_topk(cb_index* ix, const float* dQ, int nq, int T, int k, const uint64_t* d_pairs, const int64_t* d_list_off,
                      const int32_t* d_lens, int64_t* d_out_pids, float* d_out_scores, cudaStream_t st);

static int32_t check_query_args(const cb_index* ix, const void* Q, int nq, int T, int nprobe) {
  CB_REQUIRE(ix != nullptr, CB_ERR_BAD_ARG, "index handle is NULL");
  CB_REQUIRE(nq >= 0 && T >= 1, CB_ERR_BAD_ARG, "bad query shape (nq = %d, T = %d)", nq, T);
  CB_REQUIRE(nq == 0 || Q != nullptr, CB_ERR_BAD_ARG, "Q is NULL");
  CB_REQUIRE(nprobe >= 1 && nprobe <= CB_MAX_NPROBE, CB_ERR_UNSUPPORTED, "nprobe must be in 1..%d (got %d)",
             CB_MAX_NPROBE, nprobe);
  return CB_OK;
}

// Upper bound of the pair list of a chunk that needs no device data: every (query, passage) pair comes
// from at least one IVF entry of one probed slot, and a slot holds at most max_cell_len entries.
static int64_t cb_pair_bound(const cb_index* ix, int nq, int T, int nprobe) {
  const int64_t by_cells = (int64_t)nq * T * nprobe * ix->max_cell_len;
  const int64_t by_passages = (int64_t)nq * ix->Np;
  return by_cells < by_passages ? by_cells : by_passages;
}
constexpr int64_t CB_PAIR_BOUND_LIMIT = (int64_t)2 << 30;   // pairs (16 GB of keys): above this the list is sized exactly, with a host round trip

static int32_t begin_batch_stats(cb_index* ix, cudaStream_t st) {
  CB_CUDA(cudaMemsetAsync(ix->d_stats.p, 0, CB_STATS_BYTES, st));   // (allocated with the index)
  return CB_OK;
}

__global__ void k_note_range_flag(const int* __restrict__ flag, unsigned long long* __restrict__ stat) {
  if (*flag) atomicAdd(stat, 1ULL);
}

static int32_t end_batch_stats(cb_index* ix, cudaStream_t st) {
  CB_CUDA(cudaMemcpyAsync(ix->pinned_stats, ix->d_stats.p, CB_STATS_BYTES, cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaEventRecord(ix->ev_stats, st));
  ix->stats_pending = true;
  return CB_OK;
}

// stages 1+2 for one chunk of <= CB_NQ_CHUNK queries; leaves bitmap / counts / list_off in the
// workspace.  With total_pairs the pair count of the chunk is returned (one host round trip).
int32_t cb_candidates_chunk(cb_index* ix, const float* dQ, int nq, int T, int nprobe, int W, cudaStream_t st,
                            int64_t* total_pairs, const int32_t* d_cells_in) {
  const int64_t nrows = (int64_t)nq * T;
  CB_TRY(ix->cells.ensure(sizeof(int32_t) * nrows * nprobe));
  CB_TRY(ix->cell_scores.ensure(sizeof(float) * nrows * nprobe));
  CB_TRY(ix->bitmap.ensure(sizeof(uint32_t) * (size_t)ix->Np * W + 16));
  CB_TRY(ix->counts.ensure(sizeof(int32_t) * CB_NQ_CHUNK));
  CB_TRY(ix->cursors.ensure(sizeof(int32_t) * CB_NQ_CHUNK));
  CB_TRY(ix->list_off.ensure(sizeof(int64_t) * (CB_NQ_CHUNK + 1)));
  if (ix->opt_profile) CB_CUDA(cudaEventRecord(ix->ev[0], st));
  ix->q_prep_src = nullptr;
  const int32_t* d_cells = d_cells_in;
  if (d_cells == nullptr) {
    CB_TRY(cb_stage1_probe(ix, dQ, nrows, nprobe, ix->cells.as<int32_t>(), ix->cell_scores.as<float>(), st));
    d_cells = ix->cells.as<int32_t>();
  }
  if (ix->opt_profile) CB_CUDA(cudaEventRecord(ix->ev[1], st));
  // (the bitmap needs no clearing: stage 2's transpose writes every word of it)
  CB_CUDA(cudaMemsetAsync(ix->counts.p, 0, sizeof(int32_t) * CB_NQ_CHUNK, st));
  CB_CUDA(cudaMemsetAsync(ix->cursors.p, 0, sizeof(int32_t) * CB_NQ_CHUNK, st));
  CB_TRY(cb_stage2_mark(ix, d_cells, nq, T, nprobe, W, ix->bitmap.as<uint32_t>(), ix->counts.as<int32_t>(), st));
  CB_TRY(cb_scan_counts(ix->counts.as<int32_t>(), nq, ix->list_off.as<int64_t>(), st, cb_stats_dev(ix)));
  if (ix->opt_profile) CB_CUDA(cudaEventRecord(ix->ev[2], st));
  if (total_pairs != nullptr) {
    CB_CUDA(cudaMemcpyAsync(ix->pinned_total, ix->list_off.as<int64_t>() + nq, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    *total_pairs = ix->pinned_total[0];
  }
  return CB_OK;
}

int32_t cb_stage34_score(cb_index* ix, const float* dQ, int nq, int T, int W, const uint32_t* d_bitmap,
                         const int64_t* d_list_off, int32_t* d_cursors, uint64_t* d_pairs, cudaStream_t st) {
  if (!ix->opt_force_generic && cb_stage34_tc_supported(ix, T)) {
    // The tcgen05 kernel needs |query token| <= 255 (stage34_tc.cu): building the batch's fp16 row image
    // raises q_flag when a row breaks that; the tensor-core kernel then returns at once and the generic
    // fp32 kernel, gated on the same flag, scores the batch instead -- routed on the device, no host sync.
    ix->stats_tc_selected = 1;
    CB_TRY(cb_stage34_tc(ix, dQ, nq, T, W, d_bitmap, d_list_off, d_cursors, d_pairs, st));
    k_note_range_flag<<<1, 1, 0, st>>>(ix->q_flag.as<int>(), cb_stats_dev(ix) + 3);
    CB_LAUNCH_CHECK();
    return cb_stage34_generic(ix, dQ, nq, T, W, d_bitmap, nullptr, 0, d_list_off, d_cursors, d_pairs, st, ix->q_flag.as<int>());
  }
  ix->stats_tc_selected = 0;
  return cb_stage34_generic(ix, dQ, nq, T, W, d_bitmap, nullptr, 0, d_list_off, d_cursors, d_pairs, st);
}

static void reset_stats(cb_index* ix) {
  ix->st_pairs = ix->st_pair_embs = ix->st_flagged = ix->st_tc_pairs = ix->st_generic_pairs = ix->st_s1_tc_rows = 0;
  ix->st_rescore_unsafe = 0;
  for (double& m : ix->st_ms) m = 0;
}

// 1-based cells as they cross the ABI (0 = none) <-> 0-based internal (-1 = none)
__global__ void k_cells_one_based(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] < 0 ? 0 : in[i] + 1;
}
__global__ void k_cells_zero_based(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int64_t K,
                                   unsigned long long* __restrict__ stat_bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t c = in[i];
  if (c < 0 || (int64_t)c > K) { atomicAdd(stat_bad, 1ULL); out[i] = -1; return; }   // out of 0:K -> treated as "none", counted
  out[i] = c - 1;
}

// The whole batch as a pure stream of launches.  d_cells_1b: optional caller-supplied stage-1 result.
static int32_t search_batch_impl(cb_index* ix, const float* dQ, const int32_t* d_cells_1b, int32_t nq, int32_t T, int32_t nprobe,
                                 int32_t k, int64_t* d_out_pids, float* d_out_scores, int32_t* d_out_counts, cudaStream_t st) {
  CB_TRY(check_query_args(ix, dQ, nq, T, nprobe));
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K, CB_ERR_UNSUPPORTED, "k must be in 1..%d (got %d)", CB_MAX_K, k);
  CB_REQUIRE(nq == 0 || (d_out_pids && d_out_scores && d_out_counts), CB_ERR_BAD_ARG, "output pointer is NULL");
  CB_CUDA(cudaSetDevice(ix->device));
  const long long launches0 = g_cb_launches;
  reset_stats(ix);
  CB_TRY(begin_batch_stats(ix, st));
  for (int q0 = 0; q0 < nq; q0 += CB_NQ_CHUNK) {
    const int n = (nq - q0 < CB_NQ_CHUNK) ? nq - q0 : CB_NQ_CHUNK;
    const int W = (n + 31) / 32;
    const float* dQc = dQ + (int64_t)q0 * T * ix->dim;
    const int32_t* d_cells = nullptr;
    if (d_cells_1b != nullptr) {
      const int64_t nc = (int64_t)n * T * nprobe;
      CB_TRY(ix->cells.ensure(sizeof(int32_t) * nc));
      k_cells_zero_based<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(d_cells_1b + (int64_t)q0 * T * nprobe, ix->cells.as<int32_t>(), nc,
                                                                      ix->K, cb_stats_dev(ix) + 4);
      CB_LAUNCH_CHECK();
      d_cells = ix->cells.as<int32_t>();
    }
    const int64_t bound = cb_pair_bound(ix, n, T, nprobe);
    int64_t total = bound;
    const bool exact_size = ix->opt_sync_pairs || bound > CB_PAIR_BOUND_LIMIT;
    CB_TRY(cb_candidates_chunk(ix, dQc, n, T, nprobe, W, st, exact_size ? &total : nullptr, d_cells));
    CB_TRY(ix->pairs.ensure(sizeof(uint64_t) * (size_t)(total > 0 ? total : 1)));
    if (total > 0)
      CB_TRY(cb_stage34_score(ix, dQc, n, T, W, ix->bitmap.as<uint32_t>(), ix->list_off.as<int64_t>(),
                              ix->cursors.as<int32_t>(), ix->pairs.as<uint64_t>(), st));
    if (ix->opt_profile) CB_CUDA(cudaEventRecord(ix->ev[3], st));
    CB_TRY(cb_final_topk(ix, dQc, n, T, k, ix->pairs.as<uint64_t>(), ix->list_off.as<int64_t>(), nullptr,
                         d_out_pids + (int64_t)q0 * k, d_out_scores + (int64_t)q0 * k, st));
    CB_CUDA(cudaMemcpyAsync(d_out_counts + q0, ix->counts.p, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
    if (ix->opt_profile) {
      CB_CUDA(cudaEventRecord(ix->ev[4], st));
      CB_CUDA(cudaEventSynchronize(ix->ev[4]));
      float ms;
      for (int s = 0; s < 4; s++) {
        CB_CUDA(cudaEventElapsedTime(&ms, ix->ev[s], ix->ev[s + 1]));
        ix->st_ms[s] += ms;
      }
      CB_CUDA(cudaEventElapsedTime(&ms, ix->ev[0], ix->ev[4]));
      ix->st_ms[4] += ms;
    }
  }
  CB_TRY(end_batch_stats(ix, st));
  ix->st_launches = g_cb_launches - launches0;
  return CB_OK;
}

extern "C" int32_t cb_search_batch_device(cb_index* ix, const float* dQ, int32_t nq, int32_t T, int32_t nprobe,
                                          int32_t k, int64_t* d_out_pids, float* d_out_scores,
                                          int32_t* d_out_counts, void* stream) {
  return search_batch_impl(ix, dQ, nullptr, nq, T, nprobe, k, d_out_pids, d_out_scores, d_out_counts, (cudaStream_t)stream);
}

extern "C" int32_t cb_search_batch_cells_device(cb_index* ix, const float* dQ, const int32_t* d_cells, int32_t nq, int32_t T,
                                                int32_t nprobe, int32_t k, int64_t* d_out_pids, float* d_out_scores,
                                                int32_t* d_out_counts, void* stream) {
  CB_REQUIRE(nq == 0 || d_cells != nullptr, CB_ERR_BAD_ARG, "d_cells is NULL");
  return search_batch_impl(ix, dQ, d_cells, nq, T, nprobe, k, d_out_pids, d_out_scores, d_out_counts, (cudaStream_t)stream);
}

extern "C" int32_t cb_probe_device(cb_index* ix, const float* dQ, int32_t nq, int32_t T, int32_t nprobe, int32_t* d_out_cells,
                                   void* stream) {
  CB_TRY(check_query_args(ix, dQ, nq, T, nprobe));
  CB_REQUIRE(nq == 0 || d_out_cells, CB_ERR_BAD_ARG, "d_out_cells is NULL");
  if (nq == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nrows = (int64_t)nq * T;
  CB_TRY(ix->cells.ensure(sizeof(int32_t) * nrows * nprobe));
  CB_TRY(ix->cell_scores.ensure(sizeof(float) * nrows * nprobe));
  CB_TRY(begin_batch_stats(ix, st));
  ix->q_prep_src = nullptr;
  CB_TRY(cb_stage1_probe(ix, dQ, nrows, nprobe, ix->cells.as<int32_t>(), ix->cell_scores.as<float>(), st));
  ix->q_prep_src = nullptr;   // the row image covers only these queries: a following search rebuilds its own
  k_cells_one_based<<<(unsigned)((nrows * nprobe + 255) / 256), 256, 0, st>>>(ix->cells.as<int32_t>(), d_out_cells, nrows * nprobe);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_search_batch(cb_index* ix, const float* Q, int32_t nq, int32_t T, int32_t nprobe, int32_t k,
                                   int64_t* out_pids, float* out_scores, int32_t* out_counts) {
  CB_TRY(check_query_args(ix, Q, nq, T, nprobe));
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K, CB_ERR_UNSUPPORTED, "k must be in 1..%d (got %d)", CB_MAX_K, k);
  CB_REQUIRE(nq == 0 || (out_pids && out_scores && out_counts), CB_ERR_BAD_ARG, "output pointer is NULL");
  if (nq == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(ix->device));
  const size_t qbytes = sizeof(float) * (size_t)nq * T * ix->dim;
  CB_TRY(ix->q_f32.ensure(qbytes));
  CB_TRY(ix->out_pids.ensure(sizeof(int64_t) * (size_t)nq * k));
  CB_TRY(ix->out_scores.ensure(sizeof(float) * (size_t)nq * k));
  CB_TRY(ix->out_counts.ensure(sizeof(int32_t) * (size_t)nq));
  CB_CUDA(cudaMemcpyAsync(ix->q_f32.p, Q, qbytes, cudaMemcpyHostToDevice, nullptr));
  CB_TRY(cb_search_batch_device(ix, ix->q_f32.as<float>(), nq, T, nprobe, k, ix->out_pids.as<int64_t>(),
                                ix->out_scores.as<float>(), ix->out_counts.as<int32_t>(), nullptr));
  CB_CUDA(cudaMemcpyAsync(out_pids, ix->out_pids.p, sizeof(int64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, nullptr));
  CB_CUDA(cudaMemcpyAsync(out_scores, ix->out_scores.p, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, nullptr));
  CB_CUDA(cudaMemcpyAsync(out_counts, ix->out_counts.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, nullptr));
  CB_CUDA(cudaStreamSynchronize(nullptr));
  return CB_OK;
}

// ---------------------------------------------------------------------------------------------
// stage hooks
// ---------------------------------------------------------------------------------------------
extern "C" int32_t cb_probe(cb_index* ix, const float* Q, int32_t nq, int32_t T, int32_t nprobe,
                            int32_t* out_cells, float* out_scores) {
  CB_TRY(check_query_args(ix, Q, nq, T, nprobe));
  CB_REQUIRE(nq == 0 || out_cells, CB_ERR_BAD_ARG, "out_cells is NULL");
  if (nq == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(ix->device));
  reset_stats(ix);
  const int64_t nrows = (int64_t)nq * T;
  const size_t qbytes = sizeof(float) * (size_t)nrows * ix->dim;
  CB_TRY(ix->q_f32.ensure(qbytes));
  CB_TRY(ix->cells.ensure(sizeof(int32_t) * nrows * nprobe));
  CB_TRY(ix->cell_scores.ensure(sizeof(float) * nrows * nprobe));
  CB_TRY(ix->hook_a.ensure(sizeof(int32_t) * nrows * nprobe));
  CB_CUDA(cudaMemcpyAsync(ix->q_f32.p, Q, qbytes, cudaMemcpyHostToDevice, nullptr));
  CB_TRY(begin_batch_stats(ix, nullptr));
  CB_TRY(cb_stage1_probe(ix, ix->q_f32.as<float>(), nrows, nprobe, ix->cells.as<int32_t>(),
                         ix->cell_scores.as<float>(), nullptr));
  CB_TRY(end_batch_stats(ix, nullptr));
  k_cells_one_based<<<(unsigned)((nrows * nprobe + 255) / 256), 256>>>(ix->cells.as<int32_t>(),
                                                                     ix->hook_a.as<int32_t>(), nrows * nprobe);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemcpy(out_cells, ix->hook_a.p, sizeof(int32_t) * nrows * nprobe, cudaMemcpyDeviceToHost));
  if (out_scores)
    CB_CUDA(cudaMemcpy(out_scores, ix->cell_scores.p, sizeof(float) * nrows * nprobe, cudaMemcpyDeviceToHost));
  return CB_OK;
}

// ordered compaction of the set bits of bitmap column 0 (W = 1) -> ascending 1-based pids.
// Single CTA, chunked scan: ascending order is what `sort(unique(...))` (ranking.jl:42) yields.
__global__ void __launch_bounds__(1024)
k_bitmap_to_pids(const uint32_t* __restrict__ bitmap, int64_t Np, int64_t pid_base, int64_t* __restrict__ out,
                 int64_t capacity) {
  __shared__ int s[1024];
  __shared__ long long s_base;
  const int tid = threadIdx.x;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int64_t p0 = 0; p0 < Np; p0 += 1024) {
    const int64_t p = p0 + tid;
    const int v = (p < Np && (bitmap[p] & 1u)) ? 1 : 0;
    s[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int add = tid >= o ? s[tid - o] : 0;
      __syncthreads();
      s[tid] += add;
      __syncthreads();
    }
    const long long pos = s_base + s[tid] - v;
    if (v && pos < capacity) out[pos] = p + 1 + pid_base;
    __syncthreads();
    if (tid == 1023) s_base += s[1023];
    __syncthreads();
  }
}

extern "C" int32_t cb_retrieve(cb_index* ix, const float* Q, int32_t T, int32_t nprobe, int64_t* out_pids,
                               int64_t capacity, int64_t* out_count) {
  CB_TRY(check_query_args(ix, Q, 1, T, nprobe));
  CB_REQUIRE(out_count != nullptr, CB_ERR_BAD_ARG, "out_count is NULL");
  CB_REQUIRE(capacity == 0 || out_pids, CB_ERR_BAD_ARG, "out_pids is NULL");
  CB_CUDA(cudaSetDevice(ix->device));
  reset_stats(ix);
  const size_t qbytes = sizeof(float) * (size_t)T * ix->dim;
  CB_TRY(ix->q_f32.ensure(qbytes));
  CB_CUDA(cudaMemcpyAsync(ix->q_f32.p, Q, qbytes, cudaMemcpyHostToDevice, nullptr));
  int64_t total = 0;
  CB_TRY(cb_candidates_chunk(ix, ix->q_f32.as<float>(), 1, T, nprobe, 1, nullptr, &total));
  *out_count = total;
  const int64_t ncopy = total < capacity ? total : capacity;
  if (ncopy > 0) {
    CB_TRY(ix->hook_a.ensure(sizeof(int64_t) * (size_t)ncopy));
    k_bitmap_to_pids<<<1, 1024>>>(ix->bitmap.as<uint32_t>(), ix->Np, ix->pid_base, ix->hook_a.as<int64_t>(), ncopy);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemcpy(out_pids, ix->hook_a.p, sizeof(int64_t) * (size_t)ncopy, cudaMemcpyDeviceToHost));
  }
  return CB_OK;
}

__global__ void k_pids_to_local(const int64_t* __restrict__ pids, int64_t n, int64_t pid_base, int64_t Np,
                                int32_t* __restrict__ out, int* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t p = pids[i] - 1 - pid_base;
  if (p < 0 || p >= Np) { atomicExch(bad, 1); p = 0; }
  out[i] = (int32_t)p;
}

extern "C" int32_t cb_score_pids(cb_index* ix, const float* Q, int32_t T, const int64_t* pids, int64_t n_pids,
                                 float* out_scores) {
  CB_TRY(check_query_args(ix, Q, 1, T, 1));
  CB_REQUIRE(n_pids >= 0 && (n_pids == 0 || (pids && out_scores)), CB_ERR_BAD_ARG, "bad pid list");
  if (n_pids == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(ix->device));
  const size_t qbytes = sizeof(float) * (size_t)T * ix->dim;
  CB_TRY(ix->q_f32.ensure(qbytes));
  CB_TRY(ix->hook_a.ensure(sizeof(int64_t) * (size_t)n_pids));
  CB_TRY(ix->hook_b.ensure(sizeof(int32_t) * (size_t)n_pids + 16));
  CB_TRY(ix->hook_c.ensure(sizeof(float) * (size_t)n_pids));
  CB_CUDA(cudaMemcpy(ix->q_f32.p, Q, qbytes, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(ix->hook_a.p, pids, sizeof(int64_t) * (size_t)n_pids, cudaMemcpyHostToDevice));
  int* d_bad = reinterpret_cast<int*>(ix->hook_b.as<int32_t>() + n_pids);
  CB_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
  k_pids_to_local<<<(unsigned)((n_pids + 255) / 256), 256>>>(ix->hook_a.as<int64_t>(), n_pids, ix->pid_base, ix->Np,
                                                            ix->hook_b.as<int32_t>(), d_bad);
  CB_LAUNCH_CHECK();
  int h_bad = 0;
  CB_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  CB_REQUIRE(!h_bad, CB_ERR_BOUNDS, "pid out of range 1:%lld (+ pid_base)", (long long)ix->Np);
  CB_TRY(cb_generic_score_list(ix, ix->q_f32.as<float>(), 1, T, ix->hook_b.as<int32_t>(), n_pids,
                               ix->hook_c.as<float>(), nullptr));
  CB_CUDA(cudaMemcpy(out_scores, ix->hook_c.p, sizeof(float) * (size_t)n_pids, cudaMemcpyDeviceToHost));
  return CB_OK;
}

// plaid.cu -- PLAID-style pruned search (BASELINE.json config 5).
//
// The reference has no implementation of this (README.md:187 lists it as roadmap); the semantics
// are the ones `oracle.plaid_search` defines on top of the reference's own `retrieve` /
// `decompress` / `maxsim` (src/search/ranking.jl, src/indexing/codecs/residual.jl):
//   1. candidates = `retrieve` with nprobe = ncells                       (stages 1+2, unchanged)
//   2. centroid c survives for query q iff max_t S[t,c] >= threshold, S = exact fixed-order fp32
//      dot (the same arithmetic as the stage-1 decision, oracle.fixed_order_dot)
//   3. approx(q,p) = sum_t max(0, max over tokens e of p with a surviving code of S[t,code_e]),
//      summed in warp-butterfly order (oracle.tree_sum32)
//   4. the first ndocs candidates under (approx desc, pid asc) ...
//   5. ... get the exact fused decompress + MaxSim (stage 3+4, unchanged) and the final top-k.
//
// Design: nothing dense.  The (nq*T) x K score matrix is never formed: the survivors of a query
// token are read off the stage-1 shortlist (its top-16 per centroid range, re-scored exactly), with
// a full exact scan only for a token whose shortlist may be incomplete above the threshold.
// Survivors (~tens per query) get their 32 token scores computed exactly once, are linked into
// per-centroid lists, and the approximate scoring is one passage-major pass: a warp per passage
// walks the lists of its tokens' codes, keeps the hits whose query holds the passage as a
// candidate, and folds them.  Pairs without a hit have approx = 0 and only matter when a query
// has fewer than ndocs positive candidates; they are then taken in ascending pid order.
#include "common.cuh"

int32_t cb_stage5_topk_lens(const uint64_t* d_pairs, const int64_t* d_list_off, const int32_t* d_lens, int nq, int k,
                            int64_t pid_base, int64_t* d_out_pids, float* d_out_scores, cudaStream_t st);   // stage5.cu

namespace {

constexpr int PL_HCAP = 128;       // hits (surviving (query, centroid) entries among one passage's tokens) kept per passage

__device__ __forceinline__ float pl_fixed_dot(const float* __restrict__ q, const float* __restrict__ c, int dim) {
  float acc = 0.f;
  for (int k = 0; k < dim; k++) acc = __fadd_rn(acc, __fmul_rn(q[k], c[k]));
  return acc;
}

// Survivors from the stage-1 shortlists.  One warp per query-token row: every shortlisted centroid
// is re-scored exactly; those >= thr are emitted as (query, centroid).  If a centroid OUTSIDE the
// shortlist could still reach thr (the best approximate score the row dropped, plus the rounding
// guard) the row is flagged for an exact scan of all K centroids instead.
__global__ void __launch_bounds__(128)
k_plaid_emit(const float* __restrict__ Q, int64_t nrows, int T, const float* __restrict__ C, int dim, int nsplit,
             const float* __restrict__ topv, const int32_t* __restrict__ topi, const float* __restrict__ thr0,
             float thr, float guard, float guard_rel, unsigned long long* __restrict__ ents, int* __restrict__ n_ents,
             int cap, int32_t* __restrict__ flags) {
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const float* q = Q + row * dim;
  const int ncand = nsplit * CB_TOPR;
  float excluded = -INFINITY;
  if (guard_rel > 0.f) {
    float ss = 0.f;
    for (int k = lane; k < dim; k += 32) ss = fmaf(q[k], q[k], ss);
    guard += guard_rel * sqrtf(cb_warp_sum(ss));
  }
  const unsigned long long qid = (unsigned long long)(row / T);
  for (int ci = lane; ci < ncand; ci += 32) {
    const int32_t cid = topi[row * ncand + ci];
    if ((uint32_t)cid >= 0x7fffffffu) continue;   // empty slot (0x7fffffff) or a segment that does not exist (-1)
    if ((ci % CB_TOPR) == CB_TOPR - 1) excluded = fmaxf(excluded, topv[row * ncand + ci]);
    const float sc = pl_fixed_dot(q, C + (int64_t)cid * dim, dim);
    if (sc >= thr) {
      const int pos = atomicAdd(n_ents, 1);
      if (pos < cap) ents[pos] = (qid << 32) | (unsigned long long)(uint32_t)cid;
    }
  }
  if (thr0 != nullptr && lane < nsplit) excluded = fmaxf(excluded, thr0[row * nsplit + lane]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) excluded = fmaxf(excluded, __shfl_xor_sync(0xffffffffu, excluded, o));
  if (lane == 0) flags[row] = (excluded > -INFINITY && excluded + guard >= thr) ? 1 : 0;
}

// Exact scan of all K centroids for a flagged row (rare): emits every c with S >= thr that the
// shortlist pass may have missed.  Duplicates of shortlist emissions are harmless (max is idempotent).
__global__ void __launch_bounds__(256)
k_plaid_emit_fullscan(const float* __restrict__ Q, int T, const float* __restrict__ C, int64_t K, int dim,
                      const int32_t* __restrict__ flagged_rows, float thr, unsigned long long* __restrict__ ents,
                      int* __restrict__ n_ents, int cap) {
  extern __shared__ float s_q[];
  const int64_t row = flagged_rows[blockIdx.x];
  for (int k = threadIdx.x; k < dim; k += 256) s_q[k] = Q[row * dim + k];
  __syncthreads();
  const unsigned long long qid = (unsigned long long)(row / T);
  for (int64_t c = threadIdx.x; c < K; c += 256) {
    if (pl_fixed_dot(s_q, C + c * dim, dim) >= thr) {
      const int pos = atomicAdd(n_ents, 1);
      if (pos < cap) ents[pos] = (qid << 32) | (unsigned long long)(uint32_t)c;
    }
  }
}

__global__ void k_plaid_compact_flags(const int32_t* __restrict__ flags, int64_t nrows, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows && flags[i]) out[1 + atomicAdd(&out[0], 1)] = (int32_t)i;
}

// One warp per survivor entry (q, c): vec[t] = max(S[t,c], 0) for the query's T tokens (0 beyond T),
// and the entry is pushed on centroid c's list.
__global__ void __launch_bounds__(128)
k_plaid_vectors(const float* __restrict__ Q, int T, const float* __restrict__ C, int dim,
                const unsigned long long* __restrict__ ents, int n, float* __restrict__ vec, int32_t* __restrict__ head,
                int32_t* __restrict__ next, uint32_t* __restrict__ mask) {
  const int e = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= n) return;
  const unsigned long long ent = ents[e];
  const int64_t q = (int64_t)(ent >> 32);
  const int64_t c = (int64_t)(ent & 0xffffffffull);
  float v = 0.f;
  if (lane < T) v = fmaxf(pl_fixed_dot(Q + (q * T + lane) * dim, C + c * dim, dim), 0.f);
  vec[(int64_t)e * 32 + lane] = v;
  if (lane == 0) {
    next[e] = atomicExch(&head[c], e);
    atomicOr(&mask[c >> 5], 1u << (c & 31));
  }
}

// Approximate scoring, passage-major: one warp per passage, 16 warps per CTA sharing one copy of the
// survivor mask, the next passage's bitmap words and extent requested while the current one is processed.
constexpr int PL_WARPS = 16;
__global__ void __launch_bounds__(PL_WARPS * 32)
k_plaid_approx(const int64_t* __restrict__ offsets, int64_t Np, const int32_t* __restrict__ codes, int W,
               const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ head, const int32_t* __restrict__ next,
               const unsigned long long* __restrict__ ents, const float* __restrict__ vec,
               const int64_t* __restrict__ list_off, int32_t* __restrict__ cursors, uint64_t* __restrict__ pairs,
               int* __restrict__ overflow, const uint32_t* __restrict__ mask, int mask_words) {
  // bit c of the mask = centroid c has a survivor list: kept in shared memory when it fits, so that the ~80 % of
  // tokens whose centroid survives for no query never touch head[] (a random 32-byte L2 sector each)
  extern __shared__ uint32_t s_mask[];
  for (int i = threadIdx.x; i < mask_words; i += blockDim.x) s_mask[i] = mask[i];
  __syncthreads();
  __shared__ uint32_t s_row[PL_WARPS][32];
  __shared__ uint32_t s_hit[PL_WARPS][PL_HCAP];     // entry id of every hit
  __shared__ uint16_t s_hq[PL_WARPS][PL_HCAP];      // its query
  __shared__ uint16_t s_eq[PL_WARPS][PL_HCAP];      // query of every positive score to append
  __shared__ int s_n[PL_WARPS];
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * PL_WARPS + wi, nwarps = (int64_t)gridDim.x * PL_WARPS;
  uint32_t wn = 0u;
  int64_t o0n = 0, o1n = 0;
  if (warp0 < Np) { wn = (lane < W) ? bitmap[warp0 * W + lane] : 0u; o0n = offsets[warp0]; o1n = offsets[warp0 + 1]; }
  for (int64_t p = warp0; p < Np; p += nwarps) {
    const uint32_t w = wn;
    const int64_t e0 = o0n;
    const int L = (int)(o1n - o0n);
    const int64_t pn = p + nwarps;
    if (pn < Np) { wn = (lane < W) ? bitmap[pn * W + lane] : 0u; o0n = offsets[pn]; o1n = offsets[pn + 1]; }
    if (!__any_sync(0xffffffffu, w != 0u)) continue;
    s_row[wi][lane] = w;
    if (lane == 0) s_n[wi] = 0;
    __syncwarp();
    for (int i = lane; i < L; i += 32) {
      const int32_t code = codes[e0 + i];
      if (mask_words > 0 && !((s_mask[code >> 5] >> (code & 31)) & 1u)) continue;
      int32_t h = head[code];
      while (h >= 0) {
        const uint32_t q = (uint32_t)(ents[h] >> 32);
        if ((s_row[wi][q >> 5] >> (q & 31)) & 1u) {
          const int pos = atomicAdd(&s_n[wi], 1);
          if (pos < PL_HCAP) { s_hit[wi][pos] = (uint32_t)h; s_hq[wi][pos] = (uint16_t)q; }
        }
        h = next[h];
      }
    }
    __syncwarp();
    int nh = s_n[wi];
    if (nh > PL_HCAP) { if (lane == 0) atomicExch(overflow, 1); nh = PL_HCAP; }
    // Fast path (the usual case: ~12 hits per passage at C, almost all for different queries): one LANE per hit.  With no
    // two hits for the same query there is nothing to maximise over, a hit's score is the sum of its own 32-float row, and the
    // lane forms it in the association order of the xor-shuffle tree below (oracle.tree_sum32: x[t] + x[t ^ 16] first, then
    // strides 8, 4, 2, 1 -- elementwise on the eight float4 of the row for the first three levels), so the bits are the same.
    if (nh <= 32) {
      const bool valid = lane < nh;
      const uint32_t q = valid ? (uint32_t)s_hq[wi][lane] : 0u;
      const bool dup = __match_any_sync(0xffffffffu, valid ? q : (0x10000u + (uint32_t)lane)) != (1u << lane);
      if (!__any_sync(0xffffffffu, dup)) {
        if (valid) {
          const float4* __restrict__ row = reinterpret_cast<const float4*>(vec + (int64_t)s_hit[wi][lane] * 32);
          auto add4 = [](const float4 a, const float4 b) {
            return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
          };
          const float4 ab = add4(add4(row[0], row[4]), add4(row[2], row[6]));     // t = 0..3: (x[t] + x[t+16]) + (x[t+8] + x[t+24])
          const float4 cd = add4(add4(row[1], row[5]), add4(row[3], row[7]));     //           (x[t+4] + x[t+20]) + (x[t+12] + x[t+28])
          const float4 s4 = add4(ab, cd);
          const float m = __fadd_rn(__fadd_rn(s4.x, s4.z), __fadd_rn(s4.y, s4.w));
          if (m > 0.f) {
            const int pos = atomicAdd(&cursors[q], 1);
            pairs[list_off[q] + pos] = ((uint64_t)cb_orderable(m) << 32) | (uint64_t)(0xffffffffu - (uint32_t)p);
          }
        }
        __syncwarp();
        continue;
      }
    }
    // fold the hits query by query: hit i leads if no earlier hit has its query.  The positive scores
    // are parked in shared memory (slot k <= i of s_hit is dead by then) and appended afterwards with
    // all the atomics of the passage in flight at once, instead of one dependent round trip per hit.
    int nemit = 0;
    for (int i = 0; i < nh; i++) {
      const uint32_t q = s_hq[wi][i];
      bool earlier = false;
      for (int j = lane; j < i; j += 32) earlier |= (s_hq[wi][j] == q);
      if (__any_sync(0xffffffffu, earlier)) continue;
      float m = vec[(int64_t)s_hit[wi][i] * 32 + lane];
      for (int j = i + 1; j < nh; j++)
        if (s_hq[wi][j] == q) m = fmaxf(m, vec[(int64_t)s_hit[wi][j] * 32 + lane]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = __fadd_rn(m, __shfl_xor_sync(0xffffffffu, m, o));   // oracle.tree_sum32
      if (m > 0.f) {      // warp-uniform: every lane holds the same sum
        __syncwarp();     // every lane is done reading s_hit[..][>= i] of this round before slot nemit <= i is reused
        if (lane == 0) { s_hit[wi][nemit] = cb_orderable(m); s_eq[wi][nemit] = (uint16_t)q; }
        nemit++;
      }
    }
    __syncwarp();
    for (int kx = lane; kx < nemit; kx += 32) {
      const uint32_t q = s_eq[wi][kx];
      const int pos = atomicAdd(&cursors[q], 1);
      pairs[list_off[q] + pos] = ((uint64_t)s_hit[wi][kx] << 32) | (uint64_t)(0xffffffffu - (uint32_t)p);
    }
    __syncwarp();
  }
}

// passages with >= 1 selected pair (the rescoring bitmap is sparse: ~1 passage in 9 at C)
__global__ void __launch_bounds__(256)
k_plaid_collect_active(const uint32_t* __restrict__ bitmap2, int64_t Np, int W, int32_t* __restrict__ list, int* __restrict__ n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp0; p < Np; p += nwarps) {
    const uint32_t w = (lane < W) ? bitmap2[p * W + lane] : 0u;
    if (__any_sync(0xffffffffu, w != 0u) && lane == 0) list[atomicAdd(n, 1)] = (int32_t)p;
  }
}

// selected[q] = min(ndocs, counts[q]); lens come from the positive-pair cursors
__global__ void k_plaid_sel_counts(const int32_t* __restrict__ counts, int nq, int ndocs, int32_t* __restrict__ sel) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq) sel[q] = min(ndocs, counts[q]);
}

// bits of the selected positive pairs: top lists hold 1-based local pids (pid_base = 0), first npos[q] valid
__global__ void k_plaid_set_selected(const int64_t* __restrict__ top_pids, const int32_t* __restrict__ npos, int nq, int ndocs,
                                     int W, uint32_t* __restrict__ bitmap2) {
  const int q = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ndocs || i >= min(npos[q], ndocs)) return;
  const int64_t pid = top_pids[(int64_t)q * ndocs + i] - 1;
  atomicOr(&bitmap2[pid * W + (q >> 5)], 1u << (q & 31));
}

// Queries with fewer positive candidates than they may select: the remaining places go to
// zero-score candidates in ascending pid order.  One CTA per query, ordered chunked scan.
__global__ void __launch_bounds__(1024)
k_plaid_fill(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ bitmap2, int64_t Np, int W,
             const int32_t* __restrict__ npos, const int32_t* __restrict__ sel) {
  __shared__ int s[1024];
  __shared__ int s_taken;
  const int q = blockIdx.x, tid = threadIdx.x;
  const int deficit = sel[q] - min(npos[q], sel[q]);
  if (deficit <= 0) return;
  const int word = q >> 5;
  const uint32_t bit = 1u << (q & 31);
  if (tid == 0) s_taken = 0;
  __syncthreads();
  for (int64_t p0 = 0; p0 < Np; p0 += 1024) {
    if (s_taken >= deficit) break;
    const int64_t p = p0 + tid;
    const int v = (p < Np && (bitmap[p * W + word] & bit) && !(bitmap2[p * W + word] & bit)) ? 1 : 0;
    s[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int add = tid >= o ? s[tid - o] : 0;
      __syncthreads();
      s[tid] += add;
      __syncthreads();
    }
    const int rank = s_taken + s[tid] - v;      // 0-based rank of this passage among the zero-score candidates so far
    if (v && rank < deficit) atomicOr(&bitmap2[p * W + word], bit);
    __syncthreads();
    if (tid == 1023) s_taken += s[1023];
    __syncthreads();
  }
}

}  // namespace

int32_t cb_final_topk(cb_index* ix, const float* dQ, int nq, int T, int k, const uint64_t* d_pairs, const int64_t* d_list_off,
                      const int32_t* d_lens, int64_t* d_out_pids, float* d_out_scores, cudaStream_t st);   // stage5.cu

extern "C" int32_t cb_search_batch_plaid_device(cb_index* ix, const float* dQ, int32_t nq, int32_t T, int32_t ncells,
                                                float centroid_score_threshold, int32_t ndocs, int32_t k,
                                                int64_t* d_out_pids, float* d_out_scores, int32_t* d_out_counts,
                                                void* stream) {
  CB_REQUIRE(ix != nullptr, CB_ERR_BAD_ARG, "index handle is NULL");
  CB_REQUIRE(nq >= 0 && T >= 1 && T <= 32, CB_ERR_UNSUPPORTED, "PLAID mode needs 1 <= T <= 32 (got nq = %d, T = %d)", nq, T);
  CB_REQUIRE(nq == 0 || dQ != nullptr, CB_ERR_BAD_ARG, "Q is NULL");
  CB_REQUIRE(ncells >= 1 && ncells <= CB_MAX_NPROBE, CB_ERR_UNSUPPORTED, "ncells must be in 1..%d (got %d)", CB_MAX_NPROBE, ncells);
  CB_REQUIRE(ndocs >= 1 && ndocs <= CB_MAX_K, CB_ERR_UNSUPPORTED, "ndocs must be in 1..%d (got %d)", CB_MAX_K, ndocs);
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K, CB_ERR_UNSUPPORTED, "k must be in 1..%d (got %d)", CB_MAX_K, k);
  CB_REQUIRE(nq == 0 || (d_out_pids && d_out_scores && d_out_counts), CB_ERR_BAD_ARG, "output pointer is NULL");
  CB_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long launches0 = g_cb_launches;
  ix->st_pairs = ix->st_pair_embs = ix->st_flagged = ix->st_tc_pairs = ix->st_generic_pairs = ix->st_s1_tc_rows = 0;
  ix->st_plaid_survivors = ix->st_plaid_positive = ix->st_plaid_rescored = 0;
  CB_CUDA(cudaMemsetAsync(ix->d_stats.p, 0, CB_STATS_BYTES, st));
  for (int q0 = 0; q0 < nq; q0 += CB_NQ_CHUNK) {
    const int n = (nq - q0 < CB_NQ_CHUNK) ? nq - q0 : CB_NQ_CHUNK;
    const int W = (n + 31) / 32;
    const float* dQc = dQ + (int64_t)q0 * T * ix->dim;
    const int64_t nrows = (int64_t)n * T;
    // 1. candidates (stage 1 with nprobe = ncells, stage 2): bitmap, counts, list_off
    int64_t total = 0;
    CB_TRY(cb_candidates_chunk(ix, dQc, n, T, ncells, W, st, &total));
    // 2. survivors of the centroid-score threshold, from the shortlists stage 1 left behind
    const int nsplit = ix->s1_nsplit;
    const int cap = (int)(nrows * nsplit * CB_TOPR < (1 << 22) ? (1 << 22) : nrows * nsplit * CB_TOPR);
    CB_TRY(ix->pl_ents.ensure(sizeof(unsigned long long) * (size_t)cap));
    CB_TRY(ix->pl_misc.ensure(sizeof(int32_t) * (size_t)(2 * nrows + 16)));
    int32_t* d_flags = ix->pl_misc.as<int32_t>() + 8;
    int32_t* d_flagged = d_flags + nrows;                 // [0] = count, [1..] rows
    int* d_n_ents = ix->pl_misc.as<int>();               // [0] entries, [1] overflow flag
    CB_CUDA(cudaMemsetAsync(ix->pl_misc.p, 0, sizeof(int32_t) * 8, st));
    CB_CUDA(cudaMemsetAsync(d_flagged, 0, sizeof(int32_t), st));
    k_plaid_emit<<<(unsigned)((nrows + 3) / 4), 128, 0, st>>>(
        dQc, nrows, T, ix->centroids, ix->dim, nsplit, ix->topr_val.as<float>(), ix->topr_idx.as<int32_t>(),
        ix->s1_used_tc ? ix->s1_thr0.as<float>() : nullptr, centroid_score_threshold, ix->s1_guard, ix->s1_guard_rel,
        ix->pl_ents.as<unsigned long long>(), d_n_ents, cap, d_flags);
    CB_LAUNCH_CHECK();
    k_plaid_compact_flags<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(d_flags, nrows, d_flagged);
    CB_LAUNCH_CHECK();
    int32_t* h = reinterpret_cast<int32_t*>(ix->pinned_total + 6);
    CB_CUDA(cudaMemcpyAsync(h, d_flagged, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    const int nflag = h[0];
    if (nflag > 0) {
      k_plaid_emit_fullscan<<<nflag, 256, sizeof(float) * ix->dim, st>>>(
          dQc, T, ix->centroids, ix->K, ix->dim, d_flagged + 1, centroid_score_threshold,
          ix->pl_ents.as<unsigned long long>(), d_n_ents, cap);
      CB_LAUNCH_CHECK();
    }
    CB_CUDA(cudaMemcpyAsync(h, d_n_ents, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    const int n_ents = h[0];
    CB_REQUIRE(n_ents <= cap, CB_ERR_UNSUPPORTED,
               "centroid_score_threshold %g keeps %d (query, centroid) pairs: more than this build holds (%d)",
               (double)centroid_score_threshold, n_ents, cap);
    ix->st_plaid_survivors += n_ents;
    // 3. survivor score vectors + per-centroid lists, then the passage-major approximate pass
    CB_TRY(ix->pl_vec.ensure(sizeof(float) * 32 * (size_t)(n_ents > 0 ? n_ents : 1)));
    CB_TRY(ix->pl_next.ensure(sizeof(int32_t) * (size_t)(n_ents > 0 ? n_ents : 1)));
    CB_TRY(ix->pl_head.ensure(sizeof(int32_t) * (size_t)ix->K));
    CB_CUDA(cudaMemsetAsync(ix->pl_head.p, 0xff, sizeof(int32_t) * (size_t)ix->K, st));
    const int mask_words_all = (int)((ix->K + 31) / 32);
    CB_TRY(ix->pl_mask.ensure(sizeof(uint32_t) * (size_t)mask_words_all));
    CB_CUDA(cudaMemsetAsync(ix->pl_mask.p, 0, sizeof(uint32_t) * (size_t)mask_words_all, st));
    if (n_ents > 0) {
      k_plaid_vectors<<<(unsigned)((n_ents + 3) / 4), 128, 0, st>>>(dQc, T, ix->centroids, ix->dim,
                                                                  ix->pl_ents.as<unsigned long long>(), n_ents,
                                                                  ix->pl_vec.as<float>(), ix->pl_head.as<int32_t>(),
                                                                  ix->pl_next.as<int32_t>(), ix->pl_mask.as<uint32_t>());
      CB_LAUNCH_CHECK();
    }
    CB_TRY(ix->pairs.ensure(sizeof(uint64_t) * (size_t)(total > 0 ? total : 1)));
    if (total > 0 && n_ents > 0 && ix->Np > 0) {
      const int grid = ix->sm_count * 4;
      const int mask_words = mask_words_all * 4 <= 160 * 1024 ? mask_words_all : 0;   // else: no shared-memory filter
      const size_t msm = sizeof(uint32_t) * (size_t)mask_words;
      CB_CUDA(cudaFuncSetAttribute(k_plaid_approx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(160 * 1024)));
      k_plaid_approx<<<grid, PL_WARPS * 32, msm, st>>>(ix->offsets, ix->Np, ix->codes, W, ix->bitmap.as<uint32_t>(),
                                           ix->pl_head.as<int32_t>(), ix->pl_next.as<int32_t>(),
                                           ix->pl_ents.as<unsigned long long>(), ix->pl_vec.as<float>(),
                                           ix->list_off.as<int64_t>(), ix->cursors.as<int32_t>(), ix->pairs.as<uint64_t>(),
                                           d_n_ents + 1, ix->pl_mask.as<uint32_t>(), mask_words);
      CB_LAUNCH_CHECK();
    }
    // 4. first ndocs positives per query (cursors = number of positive pairs), then zero-score fill
    CB_TRY(ix->pl_top_pids.ensure(sizeof(int64_t) * (size_t)n * ndocs));
    CB_TRY(ix->pl_top_scores.ensure(sizeof(float) * (size_t)n * ndocs));
    CB_TRY(ix->pl_sel.ensure(sizeof(int32_t) * CB_NQ_CHUNK));
    CB_TRY(ix->pl_npos.ensure(sizeof(int32_t) * CB_NQ_CHUNK));
    CB_CUDA(cudaMemcpyAsync(ix->pl_npos.p, ix->cursors.p, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
    CB_TRY(cb_stage5_topk_lens(ix->pairs.as<uint64_t>(), ix->list_off.as<int64_t>(), ix->pl_npos.as<int32_t>(), n, ndocs, 0,
                               ix->pl_top_pids.as<int64_t>(), ix->pl_top_scores.as<float>(), st));
    k_plaid_sel_counts<<<(n + 255) / 256, 256, 0, st>>>(ix->counts.as<int32_t>(), n, ndocs, ix->pl_sel.as<int32_t>());
    CB_LAUNCH_CHECK();
    CB_TRY(ix->bitmap2.ensure(sizeof(uint32_t) * (size_t)ix->Np * W + 16));
    CB_CUDA(cudaMemsetAsync(ix->bitmap2.p, 0, sizeof(uint32_t) * (size_t)ix->Np * W, st));
    if (ix->Np > 0) {
      dim3 g((ndocs + 255) / 256, n);
      k_plaid_set_selected<<<g, 256, 0, st>>>(ix->pl_top_pids.as<int64_t>(), ix->pl_npos.as<int32_t>(), n, ndocs, W,
                                              ix->bitmap2.as<uint32_t>());
      CB_LAUNCH_CHECK();
      k_plaid_fill<<<n, 1024, 0, st>>>(ix->bitmap.as<uint32_t>(), ix->bitmap2.as<uint32_t>(), ix->Np, W,
                                       ix->pl_npos.as<int32_t>(), ix->pl_sel.as<int32_t>());
      CB_LAUNCH_CHECK();
      CB_TRY(ix->pl_active.ensure(sizeof(int32_t) * (size_t)ix->Np));
      k_plaid_collect_active<<<ix->sm_count * 8, 256, 0, st>>>(ix->bitmap2.as<uint32_t>(), ix->Np, W,
                                                              ix->pl_active.as<int32_t>(), d_n_ents + 2);
      CB_LAUNCH_CHECK();
    }
    CB_CUDA(cudaMemcpyAsync(h, d_n_ents + 1, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));   // [0] overflow flag, [1] active passages
    // 5. exact scoring of the selected pairs + final top-k
    CB_TRY(cb_scan_counts(ix->pl_sel.as<int32_t>(), n, ix->list_off.as<int64_t>(), st));
    CB_CUDA(cudaMemsetAsync(ix->cursors.p, 0, sizeof(int32_t) * CB_NQ_CHUNK, st));
    CB_CUDA(cudaMemcpyAsync(ix->pinned_total, ix->list_off.as<int64_t>() + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    CB_REQUIRE(h[0] == 0, CB_ERR_UNSUPPORTED,
               "a passage holds more than %d surviving (query, centroid) hits: raise centroid_score_threshold", PL_HCAP);
    const int64_t total2 = ix->pinned_total[0];
    ix->st_plaid_positive += (double)total;     // candidate pairs that went through the approximate pass
    ix->st_plaid_rescored += (double)total2;
    ix->st_pairs = (double)total2;
    CB_TRY(ix->pairs.ensure(sizeof(uint64_t) * (size_t)(total2 > 0 ? total2 : 1)));
    if (total2 > 0) {
      ix->tc_active_list = ix->pl_active.as<int32_t>();   // the tcgen05 kernel visits only these passages
      ix->tc_active_n = h[1];
      const int32_t rc = cb_stage34_score(ix, dQc, n, T, W, ix->bitmap2.as<uint32_t>(), ix->list_off.as<int64_t>(),
                                          ix->cursors.as<int32_t>(), ix->pairs.as<uint64_t>(), st);
      ix->tc_active_list = nullptr;
      ix->tc_active_n = 0;
      CB_TRY(rc);
    }
    CB_TRY(cb_final_topk(ix, dQc, n, T, k, ix->pairs.as<uint64_t>(), ix->list_off.as<int64_t>(), nullptr,
                         d_out_pids + (int64_t)q0 * k, d_out_scores + (int64_t)q0 * k, st));
    CB_CUDA(cudaMemcpyAsync(d_out_counts + q0, ix->pl_sel.p, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
  }
  ix->stats_pending = false;   // this mode keeps its counters on the host (it synchronises per chunk anyway)
  ix->st_launches = g_cb_launches - launches0;
  return CB_OK;
}

extern "C" int32_t cb_search_batch_plaid(cb_index* ix, const float* Q, int32_t nq, int32_t T, int32_t ncells,
                                         float centroid_score_threshold, int32_t ndocs, int32_t k, int64_t* out_pids,
                                         float* out_scores, int32_t* out_counts) {
  CB_REQUIRE(ix != nullptr, CB_ERR_BAD_ARG, "index handle is NULL");
  CB_REQUIRE(nq >= 0 && T >= 1, CB_ERR_BAD_ARG, "bad query shape (nq = %d, T = %d)", nq, T);
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K, CB_ERR_UNSUPPORTED, "k must be in 1..%d (got %d)", CB_MAX_K, k);
  CB_REQUIRE(nq == 0 || (Q && out_pids && out_scores && out_counts), CB_ERR_BAD_ARG, "NULL pointer");
  if (nq == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(ix->device));
  const size_t qbytes = sizeof(float) * (size_t)nq * T * ix->dim;
  CB_TRY(ix->q_f32.ensure(qbytes));
  CB_TRY(ix->out_pids.ensure(sizeof(int64_t) * (size_t)nq * k));
  CB_TRY(ix->out_scores.ensure(sizeof(float) * (size_t)nq * k));
  CB_TRY(ix->out_counts.ensure(sizeof(int32_t) * (size_t)nq));
  CB_CUDA(cudaMemcpyAsync(ix->q_f32.p, Q, qbytes, cudaMemcpyHostToDevice, nullptr));
  CB_TRY(cb_search_batch_plaid_device(ix, ix->q_f32.as<float>(), nq, T, ncells, centroid_score_threshold, ndocs, k,
                                      ix->out_pids.as<int64_t>(), ix->out_scores.as<float>(), ix->out_counts.as<int32_t>(),
                                      nullptr));
  CB_CUDA(cudaMemcpyAsync(out_pids, ix->out_pids.p, sizeof(int64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, nullptr));
  CB_CUDA(cudaMemcpyAsync(out_scores, ix->out_scores.p, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, nullptr));
  CB_CUDA(cudaMemcpyAsync(out_counts, ix->out_counts.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, nullptr));
  CB_CUDA(cudaStreamSynchronize(nullptr));
  return CB_OK;
}

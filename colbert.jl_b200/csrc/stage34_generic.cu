// stage34_generic.cu -- generic (any dim % 8 == 0, any T, nbits 1..8, any doclen) fused
// residual decompression + MaxSim in plain fp32 on CUDA cores.  It is (a) the scoring path for
// shapes the tcgen05 kernel does not cover (and for passages too long for its shared-memory
// tile), (b) the exact-fp32 stage-3+4 hook `cb_score_pids`, and (c) the on-device fp32 yardstick
// the tensor-core kernel is checked against.  It is NOT a CPU fallback -- it is a CUDA kernel.
//
// Restates, per (query q, candidate passage p) pair, without materialising anything:
//   `_collect_compressed_embs_for_pids` (src/search/ranking.jl:46-67)   -> reads codes/residuals in place
//   `decompress`  (src/indexing/codecs/residual.jl:759-784)               -> v = C[:,code] + w[bucket]; v ./= (|v| + eps)
//   `maxsim`      (src/search/ranking.jl:69-86)                           -> sum_t max_e Q[:,t] . v_e
// Passage-major: a CTA decompresses one passage into shared memory once and scores it against
// every query of the chunk that holds it as a candidate (bitmap row), so decompression is
// amortised over the query batch.
#include "common.cuh"

constexpr int G_THREADS = 256;
constexpr int G_QB = 32;  // candidate queries whose per-token maxima are kept in smem at once

struct GenericParams {
  const float* centroids; const float* weights; const int32_t* codes; const uint8_t* residuals;
  const int64_t* offsets; int dim, nbits, R, T, nq, W, TC;
  const float* Q;                 // [nq][T][dim]
  const uint32_t* bitmap;         // [Np][W] or nullptr (all nq queries are candidates of every item)
  const int32_t* pid_list;        // items to process, or nullptr (items = 0..n_items-1)
  int64_t n_items;
  unsigned long long* work;       // dynamic work counter
  const int64_t* list_off; int32_t* cursors; uint64_t* pairs;  // pair-list output (bitmap mode)
  float* out_scores;              // [n_items][nq] direct output (hook mode) or nullptr
  const int* gate;                // optional: run only when *gate != 0
};

__global__ void __launch_bounds__(G_THREADS)
k_maxsim_generic(GenericParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dim = P.dim, T = P.T, ld = dim + 1;
  float* Ds = reinterpret_cast<float*>(smem_raw);                  // [TC][ld]
  float* Qs = Ds + (size_t)P.TC * ld;                              // [T][ld]
  uint32_t* tokmax = reinterpret_cast<uint32_t*>(Qs + (size_t)T * ld);  // [G_QB][T]
  int* s_q = reinterpret_cast<int*>(tokmax + G_QB * T);            // [CB_NQ_CHUNK]
  float* s_w = reinterpret_cast<float*>(s_q + CB_NQ_CHUNK);        // [256]
  __shared__ long long s_item;
  __shared__ int s_ncand;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = G_THREADS / 32;
  if (P.gate != nullptr && *P.gate == 0) return;
  for (int i = tid; i < (1 << P.nbits); i += G_THREADS) s_w[i] = P.weights[i];
  const float eps = 1.1920929e-07f;  // eps(Float32)

  while (true) {
    __syncthreads();
    if (tid == 0) s_item = (long long)atomicAdd(P.work, 1ULL);
    __syncthreads();
    const long long item = s_item;
    if (item >= P.n_items) break;
    const int64_t p = P.pid_list ? (int64_t)P.pid_list[item] : (int64_t)item;
    const int64_t e0 = P.offsets[p];
    const int L = (int)(P.offsets[p + 1] - e0);

    // candidate queries of this passage
    if (P.bitmap) {
      if (warp == 0) {
        uint32_t w = lane < P.W ? P.bitmap[p * P.W + lane] : 0u;
        int c = __popc(w), pre = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += v; }
        int base = pre - c;
        while (w) { int b = __ffs(w) - 1; w &= w - 1; s_q[base++] = lane * 32 + b; }
        if (lane == 31) s_ncand = pre;
      }
    } else {
      for (int i = tid; i < P.nq; i += G_THREADS) s_q[i] = i;
      if (tid == 0) s_ncand = P.nq;
    }
    __syncthreads();
    const int ncand = s_ncand;
    if (ncand == 0) continue;
    if (L == 0) {
      // a zero-length passage can only appear through an explicit pid list: sum over an empty
      // max is undefined in the reference (maximum of an empty slice throws); report -inf.
      if (P.out_scores) for (int i = tid; i < ncand; i += G_THREADS) P.out_scores[item * P.nq + s_q[i]] = -INFINITY;
      continue;
    }
    const int nchunks = (L + P.TC - 1) / P.TC;

    for (int qb = 0; qb < ncand; qb += G_QB) {
      const int nqb = min(G_QB, ncand - qb);
      for (int i = tid; i < nqb * T; i += G_THREADS) tokmax[i] = 0u;
      for (int ch = 0; ch < nchunks; ch++) {
        const int c0 = ch * P.TC, n = min(P.TC, L - c0);
        if (!(nchunks == 1 && qb > 0)) {
          __syncthreads();
          // ---- stage 3: decompress tokens [c0, c0+n) of the passage: one warp per token
          for (int e = warp; e < n; e += nwarps) {
            const int64_t g = e0 + c0 + e;
            const float* __restrict__ cptr = P.centroids + (int64_t)P.codes[g] * dim;
            const uint8_t* __restrict__ emb = P.residuals + g * P.R;
            float ss = 0.f;
            for (int d = lane; d < dim; d += 32) {
              float v = __fadd_rn(cptr[d], s_w[cb_bucket_of(emb, d, P.nbits)]);
              Ds[e * ld + d] = v;
              ss = fmaf(v, v, ss);
            }
            ss = cb_warp_sum(ss);
            const float denom = __fadd_rn(sqrtf(ss), eps);
            for (int d = lane; d < dim; d += 32) Ds[e * ld + d] = __fdiv_rn(Ds[e * ld + d], denom);
          }
        }
        // ---- stage 4: every candidate query of the block against the resident tokens
        for (int qi = 0; qi < nqb; qi++) {
          __syncthreads();
          const float* __restrict__ qg = P.Q + (int64_t)s_q[qb + qi] * T * dim;
          for (int i = tid; i < T * dim; i += G_THREADS) Qs[(i / dim) * ld + (i % dim)] = qg[i];
          __syncthreads();
          for (int idx = tid; idx < T * n; idx += G_THREADS) {
            const int t = idx % T, e = idx / T;
            const float* a = Qs + t * ld;
            const float* b = Ds + e * ld;
            float acc = 0.f;
            for (int k = 0; k < dim; k++) acc = fmaf(a[k], b[k], acc);
            atomicMax(&tokmax[qi * T + t], cb_orderable(acc));
          }
        }
      }
      __syncthreads();
      // ---- sum over query tokens (fixed order) and emit
      for (int qi = tid; qi < nqb; qi += G_THREADS) {
        float s = 0.f;
        for (int t = 0; t < T; t++) s = __fadd_rn(s, cb_unorderable(tokmax[qi * T + t]));
        const int q = s_q[qb + qi];
        if (P.out_scores) {
          P.out_scores[item * P.nq + q] = s;
        } else {
          const int slot = atomicAdd(&P.cursors[q], 1);
          P.pairs[P.list_off[q] + slot] = cb_pair_key(s, (uint32_t)p);
        }
      }
      __syncthreads();
    }
  }
}

static int32_t launch_generic(cb_index* ix, GenericParams& P, int64_t max_len, cudaStream_t st) {
  const int ld = ix->dim + 1;
  const size_t fixed = ((size_t)P.T * ld + (size_t)G_QB * P.T + CB_NQ_CHUNK + 256) * 4;
  const size_t budget = 200 * 1024;
  CB_REQUIRE(fixed + (size_t)ld * 4 <= budget, CB_ERR_UNSUPPORTED,
             "dim = %d, query length = %d do not fit the generic scoring kernel's shared memory", ix->dim, P.T);
  int64_t tc = (int64_t)((budget - fixed) / ((size_t)ld * 4));
  if (max_len < 1) max_len = 1;
  if (tc > max_len) tc = max_len;
  P.TC = (int)tc;
  const size_t smem = fixed + (size_t)P.TC * ld * 4;
  CB_CUDA(cudaFuncSetAttribute(k_maxsim_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_maxsim_generic, G_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)ix->sm_count * per_sm;
  if (grid > P.n_items) grid = P.n_items;
  if (grid < 1) return CB_OK;
  CB_TRY(ix->misc.ensure(64));
  P.work = ix->misc.as<unsigned long long>() + 1;
  CB_CUDA(cudaMemsetAsync(P.work, 0, sizeof(unsigned long long), st));
  k_maxsim_generic<<<(unsigned)grid, G_THREADS, smem, st>>>(P);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

static void fill_index_params(const cb_index* ix, GenericParams& P) {
  P.centroids = ix->centroids; P.weights = ix->weights; P.codes = ix->codes; P.residuals = ix->residuals;
  P.offsets = ix->offsets; P.dim = ix->dim; P.nbits = ix->nbits; P.R = ix->R;
}

int32_t cb_stage34_generic(cb_index* ix, const float* dQ, int nq, int T, int W, const uint32_t* d_bitmap,
                           const int32_t* d_pid_list, int64_t n_list, const int64_t* d_list_off,
                           int32_t* d_cursors, uint64_t* d_pairs, cudaStream_t st, const int* d_gate) {
  GenericParams P{};
  P.gate = d_gate;
  fill_index_params(ix, P);
  P.T = T; P.nq = nq; P.W = W; P.Q = dQ; P.bitmap = d_bitmap; P.pid_list = d_pid_list;
  P.n_items = d_pid_list ? n_list : ix->Np;
  P.list_off = d_list_off; P.cursors = d_cursors; P.pairs = d_pairs; P.out_scores = nullptr;
  return launch_generic(ix, P, ix->max_doclen, st);
}

// hook: scores of an explicit (0-based, local) pid list for nq queries, no bitmap.
int32_t cb_generic_score_list(cb_index* ix, const float* dQ, int nq, int T, const int32_t* d_pid_list,
                              int64_t n_list, float* d_out_scores, cudaStream_t st) {
  GenericParams P{};
  fill_index_params(ix, P);
  P.T = T; P.nq = nq; P.W = 0; P.Q = dQ; P.bitmap = nullptr; P.pid_list = d_pid_list; P.n_items = n_list;
  P.out_scores = d_out_scores;
  return launch_generic(ix, P, ix->max_doclen, st);
}

// ---------------------------------------------------------------------------------------------
// exact fp32 re-score of explicit (query, passage) pairs: the decision pass behind the final top-k.
// pids int64[nq][K2] (1-based global, 0 = empty slot); scores_out float[nq][K2].  One CTA per pair; the
// arithmetic is that of k_maxsim_generic above, operation for operation (decompress: c + w[b] rounded,
// fmaf sum of squares, sqrt + eps, division; dot: fmaf chain over k ascending; fixed-order sum over query
// tokens), so a score equals what cb_score_pids returns for the same pair bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int R_THREADS = 128, R_TOK = 16;

__global__ void __launch_bounds__(R_THREADS)
k_rescore_pairs(GenericParams P, const int64_t* __restrict__ pids, const float* __restrict__ approx, int K2, int k, float band,
                int64_t pid_base, int64_t Np, float* __restrict__ scores_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dim = P.dim, T = P.T, ldq = dim + 1, ldd = dim + 4;   // Q rows: conflict-free across tokens; D rows: 16-byte aligned
  float* Ds = reinterpret_cast<float*>(smem_raw);                 // [R_TOK][ldd]
  float* Qs = Ds + (size_t)R_TOK * ldd;                           // [T][ldq]
  uint32_t* tokmax = reinterpret_cast<uint32_t*>(Qs + (size_t)T * ldq);   // [T]
  float* s_w = reinterpret_cast<float*>(tokmax + T);              // [256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = R_THREADS / 32;
  const int q = blockIdx.x / K2, j = blockIdx.x % K2;
  const int64_t pid = pids[blockIdx.x];
  if (pid <= 0) { if (tid == 0) scores_out[blockIdx.x] = -INFINITY; return; }
  if (approx != nullptr && j >= k) {
    // Candidate j (rank j by tensor-core score) can only enter the exact top-k if its tensor-core score lies within
    // twice the error bound of the k-th one: k candidates have an exact score >= tau - eps, and exact_j <= approx_j + eps.
    const float tau = approx[(int64_t)q * K2 + (k - 1)];
    if (approx[blockIdx.x] < tau - band * fabsf(tau)) { if (tid == 0) scores_out[blockIdx.x] = -INFINITY; return; }
  }
  const int64_t p = pid - 1 - pid_base;
  if (p < 0 || p >= Np) { if (tid == 0) scores_out[blockIdx.x] = -INFINITY; return; }
  const int64_t e0 = P.offsets[p];
  const int L = (int)(P.offsets[p + 1] - e0);
  for (int i = tid; i < (1 << P.nbits); i += R_THREADS) s_w[i] = P.weights[i];
  const float* __restrict__ qg = P.Q + (int64_t)q * T * dim;
  for (int i = tid; i < T * dim; i += R_THREADS) Qs[(i / dim) * ldq + (i % dim)] = qg[i];
  for (int i = tid; i < T; i += R_THREADS) tokmax[i] = 0u;
  const float eps = 1.1920929e-07f;
  for (int c0 = 0; c0 < L; c0 += R_TOK) {
    const int n = min(R_TOK, L - c0);
    __syncthreads();
    for (int e = warp; e < R_TOK; e += nwarps) {
      if (e < n) {
        const int64_t g = e0 + c0 + e;
        const float* __restrict__ cptr = P.centroids + (int64_t)P.codes[g] * dim;
        const uint8_t* __restrict__ emb = P.residuals + g * P.R;
        float ss = 0.f;
        for (int d = lane; d < dim; d += 32) {
          float v = __fadd_rn(cptr[d], s_w[cb_bucket_of(emb, d, P.nbits)]);
          Ds[e * ldd + d] = v;
          ss = fmaf(v, v, ss);
        }
        ss = cb_warp_sum(ss);
        const float denom = __fadd_rn(sqrtf(ss), eps);
        for (int d = lane; d < dim; d += 32) Ds[e * ldd + d] = __fdiv_rn(Ds[e * ldd + d], denom);
      } else {
        for (int d = lane; d < dim; d += 32) Ds[e * ldd + d] = 0.f;   // rows past the passage: computed, never used
      }
    }
    __syncthreads();
    // work item = (query token t, block of 4 passage tokens): the query row is read once per 4 dot products and the
    // passage rows with 128-bit loads (a warp shares them: broadcast); each dot stays the k-ascending fmaf chain
    const int nitems = T * (R_TOK / 4);
    for (int it = tid; it < nitems; it += R_THREADS) {
      const int t = it % T, eb = (it / T) * 4;
      if (eb >= n) continue;
      const float* a = Qs + t * ldq;
      const float4* b0 = reinterpret_cast<const float4*>(Ds + (eb + 0) * ldd);
      const float4* b1 = reinterpret_cast<const float4*>(Ds + (eb + 1) * ldd);
      const float4* b2 = reinterpret_cast<const float4*>(Ds + (eb + 2) * ldd);
      const float4* b3 = reinterpret_cast<const float4*>(Ds + (eb + 3) * ldd);
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      for (int k4 = 0; k4 < dim / 4; k4++) {
        const float a0 = a[4 * k4], a1 = a[4 * k4 + 1], a2 = a[4 * k4 + 2], a3 = a[4 * k4 + 3];
        const float4 x0 = b0[k4], x1 = b1[k4], x2 = b2[k4], x3 = b3[k4];
        acc0 = fmaf(a0, x0.x, acc0); acc0 = fmaf(a1, x0.y, acc0); acc0 = fmaf(a2, x0.z, acc0); acc0 = fmaf(a3, x0.w, acc0);
        acc1 = fmaf(a0, x1.x, acc1); acc1 = fmaf(a1, x1.y, acc1); acc1 = fmaf(a2, x1.z, acc1); acc1 = fmaf(a3, x1.w, acc1);
        acc2 = fmaf(a0, x2.x, acc2); acc2 = fmaf(a1, x2.y, acc2); acc2 = fmaf(a2, x2.z, acc2); acc2 = fmaf(a3, x2.w, acc2);
        acc3 = fmaf(a0, x3.x, acc3); acc3 = fmaf(a1, x3.y, acc3); acc3 = fmaf(a2, x3.z, acc3); acc3 = fmaf(a3, x3.w, acc3);
      }
      uint32_t m = cb_orderable(acc0);
      if (eb + 1 < n) m = max(m, cb_orderable(acc1));
      if (eb + 2 < n) m = max(m, cb_orderable(acc2));
      if (eb + 3 < n) m = max(m, cb_orderable(acc3));
      atomicMax(&tokmax[t], m);
    }
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    if (L == 0) s = -INFINITY;
    else for (int t = 0; t < T; t++) s = __fadd_rn(s, cb_unorderable(tokmax[t]));
    scores_out[blockIdx.x] = s;
  }
}

int32_t cb_generic_rescore_pairs(cb_index* ix, const float* dQ, int nq, int T, const int64_t* d_pids, const float* d_approx, int K2, int k,
                                 float band, float* d_scores_out, cudaStream_t st) {
  if (nq == 0 || K2 == 0) return CB_OK;
  GenericParams P{};
  fill_index_params(ix, P);
  P.T = T; P.nq = nq; P.Q = dQ;
  const size_t smem = ((size_t)T * (ix->dim + 1) + (size_t)R_TOK * (ix->dim + 4) + T + 256) * 4;
  CB_REQUIRE(smem <= 200 * 1024, CB_ERR_UNSUPPORTED, "dim = %d, query length = %d do not fit the re-score kernel's shared memory", ix->dim, T);
  CB_CUDA(cudaFuncSetAttribute(k_rescore_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_rescore_pairs<<<(unsigned)((int64_t)nq * K2), R_THREADS, smem, st>>>(P, d_pids, d_approx, K2, k, band, ix->pid_base, ix->Np, d_scores_out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

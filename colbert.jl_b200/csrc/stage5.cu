// stage5.cu -- block-wide top-k selection.  Replaces `sortperm(scores, rev = true)` + `[1:k]`
// (src/searching.jl:125-127): instead of fully sorting ~1e5 candidate scores per query, one CTA
// per query radix-selects the k-th largest 64-bit key and sorts only the k winners.
//
// Key = orderable(score) << 32 | (0xffffffff - local_pid): keys are unique per query and
// "larger key" is exactly "earlier in the reference's output" -- descending score, ties in
// ascending pid (the reference's sort is stable over `pids`, which `retrieve` returns ascending).
// The pair lists are written in arbitrary order by the scoring kernel (atomic cursors); selection
// over a total order makes the result deterministic regardless.
#include "common.cuh"

constexpr int S5_THREADS = 512;

__device__ __forceinline__ void bitonic_sort_desc(uint64_t* s, int n_pow2, int tid, int nthreads) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < n_pow2 / 2; i += nthreads) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        uint64_t a = s[lo], b = s[hi];
        if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(S5_THREADS)
k_topk_select(const uint64_t* __restrict__ pairs, const int64_t* __restrict__ list_off, const int32_t* __restrict__ lens, int k, int kpow2,
              int64_t pid_base, int64_t* __restrict__ out_pids, float* __restrict__ out_scores) {
  extern __shared__ uint64_t s_sel[];  // kpow2 keys
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix, s_mask;
  __shared__ int s_need, s_done, s_cnt;
  const int q = blockIdx.x, tid = threadIdx.x;
  const uint64_t* keys = pairs + list_off[q];
  const int64_t n = lens ? (int64_t)lens[q] : list_off[q + 1] - list_off[q];   // lens: only a prefix of the list is filled
  const int kk = (int)min((int64_t)k, n);

  for (int i = tid; i < kpow2; i += S5_THREADS) s_sel[i] = 0ull;
  if (tid == 0) { s_prefix = 0ull; s_mask = 0ull; s_need = kk; s_done = 0; s_cnt = 0; }
  __syncthreads();

  if (kk > 0 && n > (int64_t)kk) {
    // radix select (8 bits per pass, MSB first) of the kk-th largest key
    for (int pass = 0; pass < 8; pass++) {
      if (s_done) break;
      const int shift = 56 - 8 * pass;
      for (int i = tid; i < 256; i += S5_THREADS) hist[i] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix, mask = s_mask;
      for (int64_t i = tid; i < n; i += S5_THREADS) {
        const uint64_t key = keys[i];
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int need = s_need;
        unsigned int cum = 0;
        int b = 255;
        for (; b > 0; b--) {
          if (cum + hist[b] >= (unsigned)need) break;
          cum += hist[b];
        }
        s_need = need - (int)cum;                     // how many to take from bin b
        s_prefix = prefix | ((unsigned long long)b << shift);
        s_mask = mask | (0xffull << shift);
        if (hist[b] == (unsigned)(need - (int)cum)) s_done = 1;  // the whole bin is selected
      }
      __syncthreads();
    }
    // every key >= threshold is a winner (exactly kk of them)
    const unsigned long long thr = s_prefix;
    for (int64_t i = tid; i < n; i += S5_THREADS) {
      const uint64_t key = keys[i];
      if (key >= thr) {
        int slot = atomicAdd(&s_cnt, 1);
        if (slot < kpow2) s_sel[slot] = key;
      }
    }
  } else {
    for (int64_t i = tid; i < n; i += S5_THREADS) s_sel[i] = keys[i];
  }
  __syncthreads();
  bitonic_sort_desc(s_sel, kpow2, tid, S5_THREADS);
  for (int i = tid; i < k; i += S5_THREADS) {
    if (i < kk) {
      const uint64_t key = s_sel[i];
      out_scores[(int64_t)q * k + i] = cb_unorderable((uint32_t)(key >> 32));
      out_pids[(int64_t)q * k + i] = (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffu)) + 1 + pid_base;
    } else {
      out_scores[(int64_t)q * k + i] = -INFINITY;
      out_pids[(int64_t)q * k + i] = 0;
    }
  }
}

static int pow2_at_least(int v) {
  int p = 2;
  while (p < v) p <<= 1;
  return p;
}

int32_t cb_stage5_topk(const uint64_t* d_pairs, const int64_t* d_list_off, int nq, int k, int64_t pid_base,
                       int64_t* d_out_pids, float* d_out_scores, cudaStream_t st) {
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K, CB_ERR_UNSUPPORTED, "k must be in 1..%d (got %d)", CB_MAX_K, k);
  if (nq == 0) return CB_OK;
  const int kp = pow2_at_least(k);
  k_topk_select<<<nq, S5_THREADS, sizeof(uint64_t) * kp, st>>>(d_pairs, d_list_off, nullptr, k, kp, pid_base, d_out_pids,
                                                             d_out_scores);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// Same with explicit per-query list lengths (PLAID mode: only the positive pairs are appended).
int32_t cb_stage5_topk_lens(const uint64_t* d_pairs, const int64_t* d_list_off, const int32_t* d_lens, int nq, int k,
                            int64_t pid_base, int64_t* d_out_pids, float* d_out_scores, cudaStream_t st) {
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K, CB_ERR_UNSUPPORTED, "k must be in 1..%d (got %d)", CB_MAX_K, k);
  if (nq == 0) return CB_OK;
  const int kp = pow2_at_least(k);
  k_topk_select<<<nq, S5_THREADS, sizeof(uint64_t) * kp, st>>>(d_pairs, d_list_off, d_lens, k, kp, pid_base, d_out_pids,
                                                             d_out_scores);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// ---------------------------------------------------------------------------------------------
// cross-shard merge: n_lists x [nq][k] (pid 1-based global, 0 = empty) -> first k per query by
// (score desc, pid asc).  Each passage lives in exactly one shard, so (score, pid) is a total
// order.  One CTA per query; bitonic sort of (orderable score, ~pid) pairs in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_merge_topk(int n_lists, int nq, int k, int k_out, int npow2, const int64_t* __restrict__ pids,
             const float* __restrict__ scores, int64_t* __restrict__ out_pids, float* __restrict__ out_scores) {
  extern __shared__ uint64_t s_a[];   // npow2 x orderable score (0 = empty)
  uint64_t* s_b = s_a + npow2;        // npow2 x ~pid (larger = smaller pid)
  const int q = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int total = n_lists * k;
  for (int i = tid; i < npow2; i += nth) {
    uint64_t a = 0ull, b = 0ull;
    if (i < total) {
      const int l = i / k, j = i % k;
      const int64_t p = pids[((int64_t)l * nq + q) * k + j];
      if (p > 0) {
        a = (uint64_t)cb_orderable(scores[((int64_t)l * nq + q) * k + j]);
        b = ~(uint64_t)p;
      }
    }
    s_a[i] = a; s_b[i] = b;
  }
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < npow2 / 2; i += nth) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a0 = s_a[lo], b0 = s_b[lo], a1 = s_a[hi], b1 = s_b[hi];
        const bool less = a0 < a1 || (a0 == a1 && b0 < b1);
        const bool same = (a0 == a1 && b0 == b1);
        if (!same && (less == desc)) { s_a[lo] = a1; s_b[lo] = b1; s_a[hi] = a0; s_b[hi] = b0; }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < k_out; i += nth) {
    const bool ok = i < npow2 && s_b[i] != 0ull;
    out_scores[(int64_t)q * k_out + i] = ok ? cb_unorderable((uint32_t)s_a[i]) : -INFINITY;
    out_pids[(int64_t)q * k_out + i] = ok ? (int64_t)(~s_b[i]) : 0;
  }
}

extern "C" int32_t cb_merge_topk_device(int32_t device, int32_t n_lists, int32_t nq, int32_t k,
                                        const int64_t* d_pids, const float* d_scores, int64_t* d_out_pids,
                                        float* d_out_scores, void* stream) {
  CB_REQUIRE(n_lists >= 1 && nq >= 0 && k >= 1, CB_ERR_BAD_ARG, "bad merge shape");
  CB_REQUIRE((int64_t)n_lists * k <= 8192, CB_ERR_UNSUPPORTED, "n_lists * k must be <= 8192");
  CB_REQUIRE(cb_device_count() > 0, CB_ERR_CUDA, "no CUDA device is available (no CPU fallback)");
  if (nq == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(device));
  const int np = pow2_at_least(n_lists * k);
  const size_t smem = sizeof(uint64_t) * np * 2;
  CB_CUDA(cudaFuncSetAttribute(k_merge_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_merge_topk<<<nq, 256, smem, (cudaStream_t)stream>>>(n_lists, nq, k, k, np, d_pids, d_scores, d_out_pids, d_out_scores);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_merge_topk(int32_t device, int32_t n_lists, int32_t nq, int32_t k, const int64_t* pids,
                                 const float* scores, int64_t* out_pids, float* out_scores) {
  CB_REQUIRE(pids && scores && out_pids && out_scores, CB_ERR_BAD_ARG, "NULL argument");
  CB_REQUIRE(n_lists >= 1 && nq >= 0 && k >= 1, CB_ERR_BAD_ARG, "bad merge shape");
  CB_REQUIRE(cb_device_count() > 0, CB_ERR_CUDA, "no CUDA device is available (no CPU fallback)");
  if (nq == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(device));
  const size_t n_in = (size_t)n_lists * nq * k, n_out = (size_t)nq * k;
  int64_t *d_p = nullptr, *d_op = nullptr;
  float *d_s = nullptr, *d_os = nullptr;
  CB_CUDA(cudaMalloc((void**)&d_p, n_in * 8 + n_out * 8 + n_in * 4 + n_out * 4));
  struct G { void* p; ~G() { cudaFree(p); } } g{d_p};
  d_op = d_p + n_in;
  d_s = reinterpret_cast<float*>(d_op + n_out);
  d_os = d_s + n_in;
  CB_CUDA(cudaMemcpy(d_p, pids, n_in * 8, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_s, scores, n_in * 4, cudaMemcpyHostToDevice));
  CB_TRY(cb_merge_topk_device(device, n_lists, nq, k, d_p, d_s, d_op, d_os, nullptr));
  CB_CUDA(cudaMemcpy(out_pids, d_op, n_out * 8, cudaMemcpyDeviceToHost));
  CB_CUDA(cudaMemcpy(out_scores, d_os, n_out * 4, cudaMemcpyDeviceToHost));
  return CB_OK;
}

// ---------------------------------------------------------------------------------------------
// final ranking on exact fp32 scores.  The tcgen05 scoring kernel's scores carry ~1e-4 relative error
// (fp16 operands; bound used below: eps = 2e-4 |score|), enough to reorder near-equal passages and to move
// the k-th boundary against the reference's fp32 `sortperm(scores, rev = true)[1:k]`
// (src/searching.jl:125-127).  So stage 5 selects a pool of the best K2 = max(2k, k + 16) candidates by
// tensor-core score; of those, the top k and every further one whose tensor-core score lies within 2 eps of
// the k-th (the only ones that can still enter the exact top-k) are re-scored in exact fp32
// (k_rescore_pairs = the arithmetic of cb_score_pids), and the final (score desc, pid asc) order -- the
// reference's stable sort over ascending pids -- is decided on the exact scores.  If even the LAST member of
// the pool lies inside that band the pool may have been too small: such queries are counted (stat
// "rescore_unsafe"; 0 in every run so far).
// ---------------------------------------------------------------------------------------------
int32_t cb_generic_rescore_pairs(cb_index* ix, const float* dQ, int nq, int T, const int64_t* d_pids, const float* d_approx, int K2, int k,
                                 float band, float* d_scores_out, cudaStream_t st);   // stage34_generic.cu

constexpr float CB_RESCORE_BAND = 4e-4f;   // 2 eps, eps = 2e-4 relative (measured tensor-core score error: <= 1.0e-4)

__global__ void k_rescore_guard(const float* __restrict__ approx, const int64_t* __restrict__ list_off, const int32_t* __restrict__ lens,
                                int nq, int K2, int k, float band, unsigned long long* __restrict__ stat_unsafe) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int64_t n = lens ? (int64_t)lens[q] : list_off[q + 1] - list_off[q];
  if (n <= K2) return;                                   // nothing was cut
  const float tau = approx[(int64_t)q * K2 + (k - 1)];
  const float cut = approx[(int64_t)q * K2 + (K2 - 1)];   // every cut candidate has a tensor-core score <= this
  if (!(cut < tau - band * fabsf(tau))) atomicAdd(stat_unsafe, 1ULL);
}

int32_t cb_final_topk(cb_index* ix, const float* dQ, int nq, int T, int k, const uint64_t* d_pairs, const int64_t* d_list_off,
                      const int32_t* d_lens, int64_t* d_out_pids, float* d_out_scores, cudaStream_t st) {
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K, CB_ERR_UNSUPPORTED, "k must be in 1..%d (got %d)", CB_MAX_K, k);
  if (nq == 0) return CB_OK;
  int K2 = 2 * k > k + 16 ? 2 * k : k + 16;
  if (K2 > CB_MAX_K) K2 = CB_MAX_K;
  if (!ix->opt_exact_rescore || !ix->stats_tc_selected) {   // the generic kernel's scores are already the exact ones
    const int kp = pow2_at_least(k);
    k_topk_select<<<nq, S5_THREADS, sizeof(uint64_t) * kp, st>>>(d_pairs, d_list_off, d_lens, k, kp, ix->pid_base, d_out_pids, d_out_scores);
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  CB_TRY(ix->fin_pids.ensure(sizeof(int64_t) * (size_t)nq * K2));
  CB_TRY(ix->fin_scores.ensure(sizeof(float) * (size_t)nq * K2 * 2));
  int64_t* f_pids = ix->fin_pids.as<int64_t>();
  float* f_approx = ix->fin_scores.as<float>();
  float* f_exact = f_approx + (size_t)nq * K2;
  const int kp = pow2_at_least(K2);
  k_topk_select<<<nq, S5_THREADS, sizeof(uint64_t) * kp, st>>>(d_pairs, d_list_off, d_lens, K2, kp, ix->pid_base, f_pids, f_approx);
  CB_LAUNCH_CHECK();
  CB_TRY(cb_generic_rescore_pairs(ix, dQ, nq, T, f_pids, f_approx, K2, k, CB_RESCORE_BAND, f_exact, st));
  const size_t smem = sizeof(uint64_t) * kp * 2;
  CB_CUDA(cudaFuncSetAttribute(k_merge_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_merge_topk<<<nq, 256, smem, st>>>(1, nq, K2, k, kp, f_pids, f_exact, d_out_pids, d_out_scores);
  CB_LAUNCH_CHECK();
  k_rescore_guard<<<(nq + 255) / 256, 256, 0, st>>>(f_approx, d_list_off, d_lens, nq, K2, k, CB_RESCORE_BAND, cb_stats_dev(ix) + 5);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

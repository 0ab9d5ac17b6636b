// hooks.cu -- handle-free, exact-fp32 mirrors of two reference functions, used to check parity
// stage by stage:
//   cb_decompress <-> `decompress` (src/indexing/codecs/residual.jl:759-784) incl.
//                     `decompress_residuals` (698-721), `_unpackbits` (428-441), `_unbinarize` (233-240),
//                     `_normalize_array!` (src/utils.jl:320-325)
//   cb_maxsim     <-> `maxsim` (src/search/ranking.jl:69-86)
#include "common.cuh"

// one warp per embedding
__global__ void __launch_bounds__(256)
k_decompress(const float* __restrict__ centroids, const float* __restrict__ weights, const uint32_t* __restrict__ codes,
             const uint8_t* __restrict__ residuals, int64_t n, int dim, int nbits, int R, int64_t K,
             float* __restrict__ out, uint8_t* __restrict__ out_idx, float* __restrict__ out_raw, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= n) return;
  const uint32_t code = codes[e];
  if (code < 1u || (int64_t)code > K) {  // residual.jl:766
    if (lane == 0) atomicExch(bad, 1);
    return;
  }
  const float* __restrict__ c = centroids + (int64_t)(code - 1u) * dim;
  const uint8_t* __restrict__ emb = residuals + e * R;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    const uint32_t b = cb_bucket_of(emb, d, nbits);
    const float v = __fadd_rn(c[d], weights[b]);  // centroids[:, code] + bucket_weights[idx]
    if (out_idx) out_idx[e * dim + d] = (uint8_t)b;
    if (out_raw) out_raw[e * dim + d] = v;
    out[e * dim + d] = v;
    ss = fmaf(v, v, ss);
  }
  ss = cb_warp_sum(ss);
  const float denom = __fadd_rn(sqrtf(ss), 1.1920929e-07f);  // norm + eps(Float32)
  for (int d = lane; d < dim; d += 32) out[e * dim + d] = __fdiv_rn(out[e * dim + d], denom);
}

extern "C" int32_t cb_decompress(int32_t device, int32_t dim, int32_t nbits, int64_t K, const float* centroids,
                                 const float* bucket_weights, const uint32_t* codes, const uint8_t* residuals,
                                 int64_t n, float* out_embs, uint8_t* out_bucket_idx, float* out_unnormalized) {
  CB_REQUIRE(dim > 0 && dim % 8 == 0, CB_ERR_DOMAIN, "dim should be a multiple of 8!");
  CB_REQUIRE(nbits >= 1 && nbits <= CB_MAX_NBITS, CB_ERR_UNSUPPORTED, "nbits must be in 1..%d", CB_MAX_NBITS);
  CB_REQUIRE(K >= 1 && n >= 0, CB_ERR_BAD_ARG, "bad sizes");
  CB_REQUIRE(centroids && bucket_weights && (n == 0 || (codes && residuals && out_embs)), CB_ERR_BAD_ARG, "NULL argument");
  CB_REQUIRE(cb_device_count() > 0, CB_ERR_CUDA, "no CUDA device is available (no CPU fallback)");
  if (n == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(device));
  const int R = dim / 8 * nbits;
  const size_t b_cent = sizeof(float) * (size_t)K * dim, b_w = sizeof(float) * (1u << nbits),
               b_codes = sizeof(uint32_t) * (size_t)n, b_res = (size_t)n * R, b_out = sizeof(float) * (size_t)n * dim,
               b_idx = (size_t)n * dim;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char* base = nullptr;
  const size_t total = al(b_cent) + al(b_w) + al(b_codes) + al(b_res) + 2 * al(b_out) + al(b_idx) + 256;
  CB_CUDA(cudaMalloc((void**)&base, total));
  struct G { void* p; ~G() { cudaFree(p); } } g{base};
  char* p = base;
  float* d_cent = (float*)p; p += al(b_cent);
  float* d_w = (float*)p; p += al(b_w);
  uint32_t* d_codes = (uint32_t*)p; p += al(b_codes);
  uint8_t* d_res = (uint8_t*)p; p += al(b_res);
  float* d_out = (float*)p; p += al(b_out);
  float* d_raw = (float*)p; p += al(b_out);
  uint8_t* d_idx = (uint8_t*)p; p += al(b_idx);
  int* d_bad = (int*)p;
  CB_CUDA(cudaMemcpy(d_cent, centroids, b_cent, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_w, bucket_weights, b_w, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_codes, codes, b_codes, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_res, residuals, b_res, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
  k_decompress<<<(unsigned)((n + 7) / 8), 256>>>(d_cent, d_w, d_codes, d_res, n, dim, nbits, R, K, d_out,
                                                 out_bucket_idx ? d_idx : nullptr,
                                                 out_unnormalized ? d_raw : nullptr, d_bad);
  CB_LAUNCH_CHECK();
  int h_bad = 0;
  CB_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  CB_REQUIRE(!h_bad, CB_ERR_DOMAIN, "All the codes must be in the valid range of centroid IDs! (1:%lld)", (long long)K);
  CB_CUDA(cudaMemcpy(out_embs, d_out, b_out, cudaMemcpyDeviceToHost));
  if (out_bucket_idx) CB_CUDA(cudaMemcpy(out_bucket_idx, d_idx, b_idx, cudaMemcpyDeviceToHost));
  if (out_unnormalized) CB_CUDA(cudaMemcpy(out_unnormalized, d_raw, b_out, cudaMemcpyDeviceToHost));
  return CB_OK;
}

// one CTA (128 threads) per pid: per query token t, max over the pid's embeddings of Q[t] . D[e]
// (sequential fma over dim), then the T maxima are added in order t = 0..T-1.
__global__ void __launch_bounds__(128)
k_maxsim(const float* __restrict__ Q, const float* __restrict__ D, const int64_t* __restrict__ starts,
         const int64_t* __restrict__ lens, int dim, int T, float* __restrict__ out) {
  extern __shared__ float s_max[];  // T
  const int i = blockIdx.x, tid = threadIdx.x;
  const int64_t e0 = starts[i], L = lens[i];
  for (int t = tid; t < T; t += blockDim.x) {
    float m = -INFINITY;
    for (int64_t e = 0; e < L; e++) {
      const float* a = Q + (int64_t)t * dim;
      const float* b = D + (e0 + e) * dim;
      float acc = 0.f;
      for (int k = 0; k < dim; k++) acc = fmaf(a[k], b[k], acc);
      m = fmaxf(m, acc);
    }
    s_max[t] = m;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int t = 0; t < T; t++) s = __fadd_rn(s, s_max[t]);
    out[i] = s;
  }
}

extern "C" int32_t cb_maxsim(int32_t device, int32_t dim, int32_t T, const float* Q, const float* D, int64_t M,
                             const int64_t* pids, int64_t n_pids, const int64_t* doclens, int64_t n_doclens,
                             float* out_scores) {
  CB_REQUIRE(dim > 0 && T > 0 && M >= 0 && n_pids >= 0 && n_doclens >= 0, CB_ERR_BAD_ARG, "bad sizes");
  CB_REQUIRE(Q && (M == 0 || D) && (n_pids == 0 || (pids && out_scores)) && (n_doclens == 0 || doclens),
             CB_ERR_BAD_ARG, "NULL argument");
  // offsets = cumsum([1; _head(doclens[pids])]) (ranking.jl:77) on the host: n_pids is small
  std::string err;
  int64_t total = 0;
  int64_t* h = (int64_t*)malloc(sizeof(int64_t) * (size_t)(2 * n_pids + 1));
  CB_REQUIRE(h != nullptr, CB_ERR_OOM, "host allocation failed");
  struct GH { void* p; ~GH() { free(p); } } gh{h};
  for (int64_t i = 0; i < n_pids; i++) {
    const int64_t p = pids[i];
    CB_REQUIRE(p >= 1 && p <= n_doclens, CB_ERR_BOUNDS, "pid %lld out of range 1:%lld", (long long)p, (long long)n_doclens);
    h[i] = total;
    h[n_pids + i] = doclens[p - 1];
    total += doclens[p - 1];
  }
  CB_REQUIRE(total == M, CB_ERR_BAD_ARG,
             "The total number of embeddings for pids does not match with the dimension of D! (%lld vs %lld)",
             (long long)total, (long long)M);
  CB_REQUIRE(cb_device_count() > 0, CB_ERR_CUDA, "no CUDA device is available (no CPU fallback)");
  if (n_pids == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(device));
  const size_t b_q = sizeof(float) * (size_t)T * dim, b_d = sizeof(float) * (size_t)(M ? M : 1) * dim,
               b_o = sizeof(int64_t) * (size_t)n_pids * 2, b_s = sizeof(float) * (size_t)n_pids;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char* base = nullptr;
  CB_CUDA(cudaMalloc((void**)&base, al(b_q) + al(b_d) + al(b_o) + al(b_s)));
  struct G { void* p; ~G() { cudaFree(p); } } g{base};
  float* d_q = (float*)base;
  float* d_d = (float*)(base + al(b_q));
  int64_t* d_o = (int64_t*)(base + al(b_q) + al(b_d));
  float* d_s = (float*)(base + al(b_q) + al(b_d) + al(b_o));
  CB_CUDA(cudaMemcpy(d_q, Q, b_q, cudaMemcpyHostToDevice));
  if (M) CB_CUDA(cudaMemcpy(d_d, D, sizeof(float) * (size_t)M * dim, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_o, h, b_o, cudaMemcpyHostToDevice));
  k_maxsim<<<(unsigned)n_pids, 128, sizeof(float) * T>>>(d_q, d_d, d_o, d_o + n_pids, dim, T, d_s);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemcpy(out_scores, d_s, b_s, cudaMemcpyDeviceToHost));
  return CB_OK;
}

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

// ---------------------------------------------------------------------------------------------
// cb_compress <-> `compress` (src/indexing/codecs/residual.jl:586-604): the step on the other side of the on-disk
// format.
//   codes      `compress_into_codes!` (residual.jl:67-81): argmax over centroids of emb . c == stage 1 of the search
//              path with nprobe = 1: the tcgen05 GEMM with the fused shortlist epilogue (stage1_tc.cu), decided on
//              fixed-order fp32 dots, ties -> lower centroid id (Julia's argmax returns the first maximum);
//   residuals  emb - centroids[:, code] (one fp32 subtraction, exactly the reference's), `_bucket_indices`
//              (residual.jl:348-351: searchsortedfirst(cutoffs, x) - 1 == number of cutoffs < x), `_binarize`
//              (197-208) and `_packbits` (400-407): bit b of dimension d is flat bit d * nbits + b, LSB first.
// One thread per OUTPUT BYTE: it rebuilds the <= 8 bucket indices its bits come from (any nbits in 1..8, dimensions
// may straddle bytes), so the packed rows leave as coalesced byte stores.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_compress_residuals(const float* __restrict__ embs, const float* __restrict__ centroids, const int32_t* __restrict__ cells,
                     const float* __restrict__ cutoffs, int64_t n, int dim, int nbits, int R, uint32_t* __restrict__ out_codes,
                     uint8_t* __restrict__ out_res) {
  __shared__ float s_cut[255];
  const int ncut = (1 << nbits) - 1;
  for (int i = threadIdx.x; i < ncut; i += blockDim.x) s_cut[i] = cutoffs[i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * R) return;
  const int64_t e = i / R;
  const int j = (int)(i % R);
  const int32_t code = cells[e];                                  // 0-based centroid id
  if (j == 0) out_codes[e] = (uint32_t)(code + 1);
  const float* __restrict__ x = embs + e * dim;
  const float* __restrict__ c = centroids + (int64_t)code * dim;
  uint32_t byte = 0;
  int d_prev = -1;
  uint32_t bucket = 0;
  for (int t = 0; t < 8; t++) {
    const int flat = 8 * j + t, d = flat / nbits, b = flat % nbits;
    if (d != d_prev) {
      const float r = __fsub_rn(x[d], c[d]);
      int lo = 0, hi = ncut;                                      // first index with cutoffs[idx] >= r == count of cutoffs < r
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cut[mid] < r) lo = mid + 1; else hi = mid; }
      bucket = (uint32_t)lo;
      d_prev = d;
    }
    byte |= ((bucket >> b) & 1u) << t;
  }
  out_res[i] = (uint8_t)byte;
}

extern "C" int32_t cb_compress(int32_t device, int32_t dim, int32_t nbits, int64_t K, const float* centroids, const float* bucket_cutoffs,
                               const float* embs, int64_t n, uint32_t* out_codes, uint8_t* out_residuals) {
  CB_REQUIRE(dim > 0 && dim % 8 == 0, CB_ERR_DOMAIN, "dims should be a multiple of 8!");                 // residual.jl:523
  CB_REQUIRE(nbits >= 1 && nbits <= CB_MAX_NBITS, CB_ERR_UNSUPPORTED, "nbits must be in 1..%d", CB_MAX_NBITS);
  CB_REQUIRE(K >= 1 && n >= 0, CB_ERR_BAD_ARG, "bad sizes");
  CB_REQUIRE(centroids && bucket_cutoffs && (n == 0 || (embs && out_codes && out_residuals)), CB_ERR_BAD_ARG, "NULL argument");
  CB_REQUIRE(cb_device_count() > 0, CB_ERR_CUDA, "no CUDA device is available (no CPU fallback)");
  if (n == 0) return CB_OK;
  // a codec-only index handle (no passages) gives stage 1 its centroid images and workspaces
  cb_index* ix = nullptr;
  float w0[1 << CB_MAX_NBITS] = {0};
  CB_TRY(cb_index_create(&ix, device, dim, nbits, K, 0, 0, centroids, w0, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0));
  struct GI { cb_index* p; ~GI() { cb_index_destroy(p); } } gi{ix};
  const int R = dim / 8 * nbits;
  const int64_t chunk = 1 << 17;                                  // embeddings per pass (bounds the shortlist workspace)
  const int64_t cmax = n < chunk ? n : chunk;
  char* base = nullptr;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_e = al(sizeof(float) * (size_t)cmax * dim), b_c = al(sizeof(int32_t) * (size_t)cmax), b_s = al(sizeof(float) * (size_t)cmax),
               b_oc = al(sizeof(uint32_t) * (size_t)cmax), b_or = al((size_t)cmax * R), b_cut = al(sizeof(float) * 256);
  CB_CUDA(cudaMalloc((void**)&base, b_e + b_c + b_s + b_oc + b_or + b_cut));
  struct G { void* p; ~G() { cudaFree(p); } } g{base};
  float* d_e = (float*)base;
  int32_t* d_cells = (int32_t*)(base + b_e);
  float* d_sc = (float*)(base + b_e + b_c);
  uint32_t* d_oc = (uint32_t*)(base + b_e + b_c + b_s);
  uint8_t* d_or = (uint8_t*)(base + b_e + b_c + b_s + b_oc);
  float* d_cut = (float*)(base + b_e + b_c + b_s + b_oc + b_or);
  CB_CUDA(cudaMemcpy(d_cut, bucket_cutoffs, sizeof(float) * ((1u << nbits) - 1), cudaMemcpyHostToDevice));
  for (int64_t o = 0; o < n; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    CB_CUDA(cudaMemcpy(d_e, embs + o * dim, sizeof(float) * (size_t)m * dim, cudaMemcpyHostToDevice));
    ix->q_prep_src = nullptr;
    CB_TRY(cb_stage1_probe(ix, d_e, m, 1, d_cells, d_sc, nullptr));
    const int64_t nb = m * R;
    k_compress_residuals<<<(unsigned)((nb + 255) / 256), 256>>>(d_e, ix->centroids, d_cells, d_cut, m, dim, nbits, R, d_oc, d_or);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemcpy(out_codes + o, d_oc, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(out_residuals + o * R, d_or, (size_t)m * R, cudaMemcpyDeviceToHost));
  }
  return CB_OK;
}

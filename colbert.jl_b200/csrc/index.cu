// index.cu -- handle lifetime: upload of the reference's Searcher arrays and construction of the
// derived search-time structures.  Replaces src/searching.jl:44-59 (loads) and 82-91
// (`_build_emb2pid`); the IVF-from-codes path restates `_build_ivf`
// (src/indexing/collection_indexer.jl:349-353) on the device.
#include <cub/cub.cuh>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local std::string g_last_error;
thread_local long long g_cb_launches = 0;

void cb_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

extern "C" const char* cb_version(void) { return "colbert_b200 0.1 (sm_100a; CUDA 12.9)"; }
extern "C" const char* cb_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t cb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// build kernels (index-load time, not the hot path)
// ---------------------------------------------------------------------------------------------
__global__ void k_f32_to_f16(const float* __restrict__ in, __half* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2half_rn(in[i]);
}

// max over rows of the L2 norm (one warp per row; non-negative floats order like their bit patterns)
__global__ void k_max_row_norm(const float* __restrict__ X, int64_t nrows, int dim, unsigned int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float m = 0.f;
  for (; row < nrows; row += stride) {
    float ss = 0.f;
    for (int d = lane; d < dim; d += 32) { const float v = X[row * dim + d]; ss = fmaf(v, v, ss); }
    ss = cb_warp_sum(ss);
    m = fmaxf(m, sqrtf(ss));
  }
  if (lane == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// codes: 1-based UInt32 -> 0-based int32 in place; flags any code outside 1:K (residual.jl:766).
__global__ void k_codes_zero_based(int32_t* __restrict__ codes, int64_t n, int64_t K, int* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t c = (uint32_t)codes[i];
    if (c < 1u || (int64_t)c > K) atomicExch(bad, 1);
    codes[i] = (int32_t)(c - 1u);
  }
}

__global__ void k_check_nonneg_max(const int64_t* __restrict__ v, int64_t n, int* __restrict__ bad,
                                   unsigned long long* __restrict__ vmax) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long m = 0;
  for (; i < n; i += stride) {
    int64_t x = v[i];
    if (x < 0) atomicExch(bad, 1);
    else if ((unsigned long long)x > m) m = (unsigned long long)x;
  }
  if (m) atomicMax(vmax, m);
}

// pid (0-based) of a 0-based embedding id: the unique p with offsets[p] <= e < offsets[p+1]
// (`emb2pid[eid]`, src/searching.jl:82-91; zero-length passages own no embedding).
__device__ __forceinline__ int32_t pid_of_eid(const int64_t* __restrict__ offsets, int64_t Np, int64_t e) {
  int64_t lo = 0, hi = Np;  // invariant: offsets[lo] <= e < offsets[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= e) lo = mid; else hi = mid;
  }
  return (int32_t)lo;
}

// ivf (1-based eids, int64) -> local 0-based pid per IVF entry.
__global__ void k_ivf_to_pids(const int64_t* __restrict__ ivf, int64_t n, const int64_t* __restrict__ offsets,
                              int64_t Np, int64_t Ne, int32_t* __restrict__ out, int* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t e = ivf[i] - 1;
    if (e < 0 || e >= Ne) { atomicExch(bad, 1); out[i] = 0; continue; }
    out[i] = pid_of_eid(offsets, Np, e);
  }
}

__global__ void k_eids32_to_pids(const uint32_t* __restrict__ eids, int64_t n, const int64_t* __restrict__ offsets,
                                 int64_t Np, int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = pid_of_eid(offsets, Np, (int64_t)eids[i]);
}

__global__ void k_iota_u32(uint32_t* __restrict__ v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[i] = (uint32_t)i;
}

// cell_offsets[c] = first position in the sorted 0-based codes with key >= c (c in 0..K).
__global__ void k_cell_offsets_from_sorted(const uint32_t* __restrict__ keys, int64_t n, int64_t K,
                                           int64_t* __restrict__ cell_offsets) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > K) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)keys[mid] < c) lo = mid + 1; else hi = mid;
  }
  cell_offsets[c] = lo;
}

__global__ void k_max_cell_len(const int64_t* __restrict__ cell_offsets, int64_t K, unsigned long long* __restrict__ out) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long m = 0;
  for (; c < K; c += stride) {
    const unsigned long long len = (unsigned long long)(cell_offsets[c + 1] - cell_offsets[c]);
    if (len > m) m = len;
  }
  if (m) atomicMax(out, m);
}

static inline int grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

static int32_t inclusive_scan_i64(const int64_t* d_in, int64_t* d_out, int64_t n) {
  if (n == 0) return CB_OK;
  size_t tmp_bytes = 0;
  CB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d_in, d_out, n));
  void* tmp = nullptr;
  CB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
  cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, d_in, d_out, n);
  g_cb_launches++;
  cudaFree(tmp);
  CB_CUDA(e);
  return CB_OK;
}

static int32_t upload(void* dst, const void* src, size_t bytes, int flags) {
  if (bytes == 0) return CB_OK;
  CB_CUDA(cudaMemcpy(dst, src, bytes,
                     (flags & CB_FLAG_DEVICE_POINTERS) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  return CB_OK;
}

static void destroy_index(cb_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  cudaFree(ix->centroids); cudaFree(ix->centroids_h); cudaFree(ix->centroids_img); cudaFree(ix->weights); cudaFree(ix->codes);
  if (!ix->residuals_borrowed) cudaFree(ix->residuals);
  cudaFree(ix->offsets); cudaFree(ix->cell_offsets); cudaFree(ix->ivf_pids);
  DevBuf* bufs[] = {&ix->q_f32, &ix->q_prep, &ix->topr_val, &ix->topr_idx, &ix->cells, &ix->cell_scores,
                    &ix->flags, &ix->bitmap, &ix->counts, &ix->list_off, &ix->cursors, &ix->pairs,
                    &ix->out_pids, &ix->out_scores, &ix->out_counts, &ix->misc, &ix->long_list,
                    &ix->hook_a, &ix->hook_b, &ix->hook_c, &ix->s1_thr, &ix->s1_thr0, &ix->bitmap_t, &ix->bitmap2, &ix->q_flag, &ix->d_stats, &ix->fin_keys, &ix->fin_pids, &ix->fin_scores, &ix->pl_ents, &ix->pl_misc,
                    &ix->pl_vec, &ix->pl_next, &ix->pl_head, &ix->pl_mask, &ix->pl_active, &ix->pl_top_pids, &ix->pl_top_scores, &ix->pl_sel, &ix->pl_npos};
  for (DevBuf* b : bufs) b->release();
  if (ix->pinned_total) cudaFreeHost(ix->pinned_total);
  if (ix->pinned_stats) cudaFreeHost(ix->pinned_stats);
  if (ix->ev_stats) cudaEventDestroy(ix->ev_stats);
  for (auto& e : ix->ev) if (e) cudaEventDestroy(e);
  cudaGetLastError();
  delete ix;
}

#define CB_DEVALLOC(ptr, bytes)                                               \
  do {                                                                        \
    size_t _b = (size_t)(bytes);                                              \
    CB_CUDA(cudaMalloc((void**)&(ptr), _b ? _b : 16));                        \
    ix->resident_bytes += _b;                                                 \
  } while (0)

static int32_t create_impl(cb_index* ix, const float* centroids, const float* bucket_weights,
                           const uint32_t* codes, const uint8_t* residuals, const int64_t* doclens,
                           const int64_t* ivf, const int64_t* ivf_lengths, int32_t flags) {
  const int64_t K = ix->K, Np = ix->Np, Ne = ix->Ne;
  const int dim = ix->dim;
  int* d_bad = nullptr;
  unsigned long long* d_max = nullptr;
  CB_CUDA(cudaMalloc((void**)&d_bad, 16));
  d_max = (unsigned long long*)(d_bad + 2);
  CB_CUDA(cudaMemset(d_bad, 0, 16));
  struct Guard { int* p; ~Guard() { cudaFree(p); } } guard{d_bad};
  int h_bad = 0;

  // codec
  CB_DEVALLOC(ix->centroids, sizeof(float) * K * dim);
  CB_DEVALLOC(ix->centroids_h, sizeof(__half) * K * dim);
  CB_DEVALLOC(ix->weights, sizeof(float) * (1u << ix->nbits));
  CB_TRY(upload(ix->centroids, centroids, sizeof(float) * K * dim, flags));
  CB_TRY(upload(ix->weights, bucket_weights, sizeof(float) * (1u << ix->nbits), flags));
  {
    float hw[1 << CB_MAX_NBITS];
    CB_CUDA(cudaMemcpy(hw, ix->weights, sizeof(float) * (1u << ix->nbits), cudaMemcpyDeviceToHost));
    for (unsigned i = 0; i < (1u << ix->nbits); i++) {
      const float a = hw[i] < 0 ? -hw[i] : hw[i];
      if (!(a <= ix->weight_abs_max)) ix->weight_abs_max = a;   // (a NaN weight sticks: the tcgen05 path is then refused)
    }
  }
  k_f32_to_f16<<<grid_for(K * dim), 256>>>(ix->centroids, ix->centroids_h, K * dim);
  CB_LAUNCH_CHECK();
  {  // largest centroid norm; (dim = 128) the swizzled fp16 operand image of the tcgen05 stage 1
    CB_CUDA(cudaMemset(d_max, 0, sizeof(unsigned long long)));
    k_max_row_norm<<<grid_for(K * 32), 256>>>(ix->centroids, K, dim, reinterpret_cast<unsigned int*>(d_max));
    CB_LAUNCH_CHECK();
    unsigned int bits = 0;
    CB_CUDA(cudaMemcpy(&bits, d_max, sizeof(bits), cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemset(d_max, 0, sizeof(unsigned long long)));
    memcpy(&ix->centroid_norm_max, &bits, sizeof(float));
    if (dim == 128) {
      const int64_t Kpad = (K + 127) / 128 * 128;
      CB_DEVALLOC(ix->centroids_img, (size_t)Kpad * 256);
      CB_TRY(cb_tc_prep_rows(ix->centroids, K, Kpad, ix->centroids_img, nullptr));
    }
  }

  // compressed embeddings
  CB_DEVALLOC(ix->codes, sizeof(int32_t) * Ne);
  CB_TRY(upload(ix->codes, codes, sizeof(uint32_t) * Ne, flags));
  if ((flags & CB_FLAG_DEVICE_POINTERS) && (flags & CB_FLAG_BORROW_RESIDUALS)) {
    ix->residuals = const_cast<uint8_t*>(residuals);     // read in place: the layout is the reference's Matrix{UInt8}(R, Ne) as is
    ix->residuals_borrowed = true;
  } else {
    CB_DEVALLOC(ix->residuals, (size_t)Ne * ix->R);
    CB_TRY(upload(ix->residuals, residuals, (size_t)Ne * ix->R, flags));
  }
  k_codes_zero_based<<<grid_for(Ne), 256>>>(ix->codes, Ne, K, d_bad);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  CB_REQUIRE(!h_bad, CB_ERR_DOMAIN, "All the codes must be in the valid range of centroid IDs! (1:%lld)",
             (long long)K);

  // passage offsets (exclusive prefix sum of doclens) -- what `_build_emb2pid` encodes
  CB_DEVALLOC(ix->offsets, sizeof(int64_t) * (Np + 1));
  {
    int64_t* d_doclens = nullptr;
    CB_CUDA(cudaMalloc((void**)&d_doclens, sizeof(int64_t) * (Np ? Np : 1)));
    struct G2 { int64_t* p; ~G2() { cudaFree(p); } } g2{d_doclens};
    CB_TRY(upload(d_doclens, doclens, sizeof(int64_t) * Np, flags));
    k_check_nonneg_max<<<grid_for(Np), 256>>>(d_doclens, Np, d_bad, d_max);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemset(ix->offsets, 0, sizeof(int64_t)));
    CB_TRY(inclusive_scan_i64(d_doclens, ix->offsets + 1, Np));
    int64_t total = 0;
    unsigned long long h_max = 0;
    CB_CUDA(cudaMemcpy(&total, ix->offsets + Np, sizeof(int64_t), cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(&h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost));
    CB_REQUIRE(!h_bad, CB_ERR_DOMAIN, "doclens must be non-negative");
    CB_REQUIRE(total == Ne, CB_ERR_BAD_ARG,
               "sum(doclens) = %lld must be equal to the number of embeddings %lld", (long long)total,
               (long long)Ne);
    ix->max_doclen = (int64_t)h_max;
  }

  // IVF -> per-cell passage lists
  CB_DEVALLOC(ix->cell_offsets, sizeof(int64_t) * (K + 1));
  CB_DEVALLOC(ix->ivf_pids, sizeof(int32_t) * Ne);
  if (ivf != nullptr && ivf_lengths != nullptr) {
    int64_t* d_tmp = nullptr;
    size_t tmp_elems = (size_t)(Ne > K ? Ne : K) + 1;
    CB_CUDA(cudaMalloc((void**)&d_tmp, sizeof(int64_t) * tmp_elems));
    struct G3 { int64_t* p; ~G3() { cudaFree(p); } } g3{d_tmp};
    CB_TRY(upload(d_tmp, ivf_lengths, sizeof(int64_t) * K, flags));
    k_check_nonneg_max<<<grid_for(K), 256>>>(d_tmp, K, d_bad, d_max);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemset(ix->cell_offsets, 0, sizeof(int64_t)));
    CB_TRY(inclusive_scan_i64(d_tmp, ix->cell_offsets + 1, K));
    int64_t total = 0;
    CB_CUDA(cudaMemcpy(&total, ix->cell_offsets + K, sizeof(int64_t), cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    CB_REQUIRE(!h_bad, CB_ERR_DOMAIN, "ivf_lengths must be non-negative");
    CB_REQUIRE(total == Ne, CB_ERR_BAD_ARG, "length(ivf) must be equal to sum(ivf_lengths)! (%lld vs %lld)",
               (long long)Ne, (long long)total);
    CB_TRY(upload(d_tmp, ivf, sizeof(int64_t) * Ne, flags));
    k_ivf_to_pids<<<grid_for(Ne), 256>>>(d_tmp, Ne, ix->offsets, Np, Ne, ix->ivf_pids, d_bad);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    CB_REQUIRE(!h_bad, CB_ERR_DOMAIN, "ivf entries must be embedding ids in 1:%lld", (long long)Ne);
  } else {
    CB_REQUIRE(ivf == nullptr && ivf_lengths == nullptr, CB_ERR_BAD_ARG,
               "ivf and ivf_lengths must both be given or both be NULL");
    // `_build_ivf`: ivf = sortperm(codes) (stable), ivf_lengths = counts(sort(codes), K)
    uint32_t *k_in = (uint32_t*)ix->codes, *k_out = nullptr, *v_in = nullptr, *v_out = nullptr;
    size_t n = (size_t)(Ne ? Ne : 1);
    CB_CUDA(cudaMalloc((void**)&k_out, sizeof(uint32_t) * n * 3));
    struct G4 { uint32_t* p; ~G4() { cudaFree(p); } } g4{k_out};
    v_in = k_out + n;
    v_out = v_in + n;
    k_iota_u32<<<grid_for(Ne), 256>>>(v_in, Ne);
    CB_LAUNCH_CHECK();
    if (Ne > 0) {
      int end_bit = 1;
      while (end_bit < 32 && ((int64_t)1 << end_bit) < K) end_bit++;
      size_t tmp_bytes = 0;
      CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)Ne, 0, end_bit));
      void* tmp = nullptr;
      CB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
      cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)Ne, 0, end_bit);
      g_cb_launches++;
      cudaError_t e2 = cudaDeviceSynchronize();
      cudaFree(tmp);
      CB_CUDA(e);
      CB_CUDA(e2);
    }
    k_cell_offsets_from_sorted<<<(int)((K + 1 + 255) / 256), 256>>>(k_out, Ne, K, ix->cell_offsets);
    CB_LAUNCH_CHECK();
    k_eids32_to_pids<<<grid_for(Ne), 256>>>(v_out, Ne, ix->offsets, Np, ix->ivf_pids);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaDeviceSynchronize());
  }
  {  // longest IVF cell (bounds a batch's pair list without a host round trip, search.cu)
    CB_CUDA(cudaMemset(d_max, 0, sizeof(unsigned long long)));
    k_max_cell_len<<<grid_for(K), 256>>>(ix->cell_offsets, K, d_max);
    CB_LAUNCH_CHECK();
    unsigned long long h = 0;
    CB_CUDA(cudaMemcpy(&h, d_max, sizeof(h), cudaMemcpyDeviceToHost));
    ix->max_cell_len = (int64_t)h;
  }
  CB_TRY(ix->d_stats.ensure(CB_STATS_BYTES));
  CB_CUDA(cudaMemset(ix->d_stats.p, 0, CB_STATS_BYTES));
  CB_TRY(ix->q_flag.ensure(sizeof(int)));
  CB_CUDA(cudaMemset(ix->q_flag.p, 0, sizeof(int)));
  CB_CUDA(cudaHostAlloc((void**)&ix->pinned_stats, CB_STATS_BYTES, cudaHostAllocDefault));
  memset(ix->pinned_stats, 0, CB_STATS_BYTES);
  CB_CUDA(cudaEventCreateWithFlags(&ix->ev_stats, cudaEventDisableTiming));
  CB_CUDA(cudaHostAlloc((void**)&ix->pinned_total, 64, cudaHostAllocDefault));
  for (auto& e : ix->ev) CB_CUDA(cudaEventCreate(&e));
  CB_CUDA(cudaDeviceSynchronize());
  return CB_OK;
}

extern "C" int32_t cb_index_create(cb_index** out, int32_t device, int32_t dim, int32_t nbits, int64_t K,
                                   int64_t n_passages, int64_t n_embeddings, const float* centroids,
                                   const float* bucket_weights, const uint32_t* codes,
                                   const uint8_t* residuals, const int64_t* doclens, const int64_t* ivf,
                                   const int64_t* ivf_lengths, int64_t pid_base, int32_t flags) {
  CB_REQUIRE(out != nullptr, CB_ERR_BAD_ARG, "out handle pointer is NULL");
  *out = nullptr;
  CB_REQUIRE(dim > 0 && dim % 8 == 0, CB_ERR_DOMAIN, "dim should be a multiple of 8!");
  CB_REQUIRE(nbits >= 1 && nbits <= CB_MAX_NBITS, CB_ERR_UNSUPPORTED, "nbits must be in 1..%d (got %d)",
             CB_MAX_NBITS, nbits);
  CB_REQUIRE(K >= 1 && K < ((int64_t)1 << 31), CB_ERR_BAD_ARG, "number of centroids out of range");
  CB_REQUIRE(n_passages >= 0 && n_passages < ((int64_t)1 << 31) - 1, CB_ERR_BAD_ARG, "n_passages out of range");
  CB_REQUIRE(n_embeddings >= 0 && n_embeddings < ((int64_t)1 << 32) - 1, CB_ERR_BAD_ARG, "n_embeddings out of range");
  CB_REQUIRE(centroids && bucket_weights, CB_ERR_BAD_ARG, "centroids / bucket_weights are NULL");
  CB_REQUIRE(n_embeddings == 0 || (codes && residuals), CB_ERR_BAD_ARG, "codes / residuals are NULL");
  CB_REQUIRE(n_passages == 0 || doclens, CB_ERR_BAD_ARG, "doclens is NULL");
  int ndev = cb_device_count();
  CB_REQUIRE(ndev > 0, CB_ERR_CUDA, "no CUDA device is available (this library has no CPU fallback)");
  CB_REQUIRE(device >= 0 && device < ndev, CB_ERR_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  CB_CUDA(cudaSetDevice(device));
  cb_index* ix = new (std::nothrow) cb_index();
  CB_REQUIRE(ix != nullptr, CB_ERR_OOM, "host allocation failed");
  ix->device = device; ix->dim = dim; ix->nbits = nbits; ix->R = dim / 8 * nbits;
  ix->K = K; ix->Np = n_passages; ix->Ne = n_embeddings; ix->pid_base = pid_base;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
  int32_t s = create_impl(ix, centroids, bucket_weights, codes, residuals, doclens, ivf, ivf_lengths, flags);
  if (s != CB_OK) {
    std::string keep = g_last_error;
    destroy_index(ix);
    g_last_error = keep;
    return s;
  }
  *out = ix;
  return CB_OK;
}

extern "C" int32_t cb_index_destroy(cb_index* index) {
  destroy_index(index);
  return CB_OK;
}

extern "C" int32_t cb_index_info(const cb_index* ix, int64_t info[8]) {
  CB_REQUIRE(ix && info, CB_ERR_BAD_ARG, "NULL argument");
  info[0] = ix->dim; info[1] = ix->nbits; info[2] = ix->K; info[3] = ix->Np; info[4] = ix->Ne;
  info[5] = ix->device; info[6] = ix->pid_base; info[7] = (int64_t)ix->resident_bytes;
  return CB_OK;
}

extern "C" int32_t cb_set_option(cb_index* ix, const char* key, int64_t value) {
  CB_REQUIRE(ix && key, CB_ERR_BAD_ARG, "NULL argument");
  if (!strcmp(key, "force_generic")) ix->opt_force_generic = (int)value;
  else if (!strcmp(key, "stage1_impl")) ix->opt_stage1_impl = (int)value;
  else if (!strcmp(key, "profile")) ix->opt_profile = (int)value;
  else if (!strcmp(key, "tc_astages")) ix->opt_tc_astages = (int)value;
  else if (!strcmp(key, "sync_pairs")) ix->opt_sync_pairs = (int)value;
  else if (!strcmp(key, "exact_rescore")) ix->opt_exact_rescore = (int)value;
  else { cb_set_error("unknown option '%s'", key); return CB_ERR_BAD_ARG; }
  return CB_OK;
}

extern "C" int32_t cb_get_stat(const cb_index* cix, const char* key, double* value) {
  CB_REQUIRE(cix && key && value, CB_ERR_BAD_ARG, "NULL argument");
  cb_index* ix = const_cast<cb_index*>(cix);
  if (ix->stats_pending) {   // the batch counters were copied to pinned memory at the end of the batch: wait for that copy only now
    CB_CUDA(cudaSetDevice(ix->device));
    CB_CUDA(cudaEventSynchronize(ix->ev_stats));
    ix->stats_pending = false;
    const unsigned long long* h = ix->pinned_stats;
    ix->st_pairs = (double)h[0]; ix->st_pair_embs = (double)h[1]; ix->st_flagged = (double)h[2];
    ix->st_tc_pairs = (ix->stats_tc_selected && h[3] == 0) ? (double)h[0] : 0.0;
    ix->st_generic_pairs = (double)h[0] - ix->st_tc_pairs;
    ix->st_bad_cells = (double)h[4]; ix->st_rescore_unsafe = (double)h[5];
    ix->st_tc_groups = (double)h[6]; ix->st_tc_group_rows = (double)h[7]; ix->st_tc_passage_rows = (double)h[8];
  }
  if (!strcmp(key, "launches")) *value = (double)ix->st_launches;
  else if (!strcmp(key, "pairs")) *value = ix->st_pairs;
  else if (!strcmp(key, "pair_embeddings")) *value = ix->st_pair_embs;
  else if (!strcmp(key, "flagged_rows")) *value = ix->st_flagged;
  else if (!strcmp(key, "tc_pairs")) *value = ix->st_tc_pairs;
  else if (!strcmp(key, "plaid_survivors")) *value = ix->st_plaid_survivors;
  else if (!strcmp(key, "plaid_candidates")) *value = ix->st_plaid_positive;
  else if (!strcmp(key, "plaid_rescored")) *value = ix->st_plaid_rescored;
  else if (!strcmp(key, "generic_pairs")) *value = ix->st_generic_pairs;
  else if (!strcmp(key, "tc_groups")) *value = ix->st_tc_groups;
  else if (!strcmp(key, "tc_group_rows")) *value = ix->st_tc_group_rows;
  else if (!strcmp(key, "tc_passage_rows")) *value = ix->st_tc_passage_rows;
  else if (!strcmp(key, "rescore_unsafe")) *value = ix->st_rescore_unsafe;
  else if (!strcmp(key, "bad_cells")) *value = ix->st_bad_cells;
  else if (!strcmp(key, "max_cell_len")) *value = (double)ix->max_cell_len;
  else if (!strcmp(key, "stage1_tc_rows")) *value = ix->st_s1_tc_rows;
  else if (!strcmp(key, "ms_stage1")) *value = ix->st_ms[0];
  else if (!strcmp(key, "ms_stage2")) *value = ix->st_ms[1];
  else if (!strcmp(key, "ms_stage34")) *value = ix->st_ms[2];
  else if (!strcmp(key, "ms_stage5")) *value = ix->st_ms[3];
  else if (!strcmp(key, "ms_total")) *value = ix->st_ms[4];
  else { cb_set_error("unknown stat '%s'", key); return CB_ERR_BAD_ARG; }
  return CB_OK;
}

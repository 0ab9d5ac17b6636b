// stage1.cu -- query-token x centroid scoring with a fused per-token shortlist, then an exact fp32
// decision.  Replaces `cells = Q' * centroids; _topk(cells, nprobe, dims = 2)`
// (src/search/ranking.jl:27-31, src/utils.jl:327-332) without ever materialising the
// (nq*T) x K score matrix.
//
// Decision rule (what makes candidate sets reproducible):
//   1. a fast pass keeps, per query token, the CB_TOPR best centroids of each centroid-range
//      split (approximate scores: FMA order / tensor-core rounding);
//   2. every shortlisted centroid is re-scored in fp32 in a FIXED order (k = 0..dim-1, product
//      rounded then added, no FMA) -- bit-identical to oracle.fixed_order_dot -- and the
//      top-nprobe are chosen by (score desc, centroid id asc), the `Perm` ordering of
//      `partialsortperm`;
//   3. a token whose nprobe-th exact score is not separated from the best excluded approximate
//      score by more than `guard` is flagged and re-done by an exact scan of all K centroids.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// 1a. SIMT fp32 pass: 64 rows x 64 centroids per tile, 256 threads, 4x4 register tile.
// ---------------------------------------------------------------------------------------------
constexpr int S1_BM = 64, S1_BN = 64, S1_BK = 32;

__global__ void __launch_bounds__(256)
k_stage1_simt(const float* __restrict__ Q, int64_t nrows, const float* __restrict__ C, int64_t K, int dim,
              int nsplit, float* __restrict__ topv, int32_t* __restrict__ topi) {
  __shared__ float Qs[S1_BK][S1_BM + 4];
  __shared__ float Cs[S1_BK][S1_BN + 4];
  __shared__ float Ss[S1_BM][S1_BN + 1];
  __shared__ float tv[S1_BM][CB_TOPR];
  __shared__ int32_t ti[S1_BM][CB_TOPR];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * S1_BM;
  const int split = blockIdx.y;
  const int64_t per = (K + nsplit - 1) / nsplit;
  const int64_t c_begin = split * per;
  const int64_t c_end = min(K, c_begin + per);

  for (int i = tid; i < S1_BM * CB_TOPR; i += 256) {
    tv[i / CB_TOPR][i % CB_TOPR] = -INFINITY;
    ti[i / CB_TOPR][i % CB_TOPR] = 0x7fffffff;
  }
  __syncthreads();

  for (int64_t c0 = c_begin; c0 < c_end; c0 += S1_BN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < dim; k0 += S1_BK) {
      // cooperative loads, transposed to [k][row]
      for (int i = tid; i < S1_BM * S1_BK; i += 256) {
        int r = i / S1_BK, k = i % S1_BK;
        int64_t gr = row0 + r;
        Qs[k][r] = (gr < nrows && k0 + k < dim) ? Q[gr * dim + k0 + k] : 0.f;
        int64_t gc = c0 + r;
        Cs[k][r] = (gc < c_end && k0 + k < dim) ? C[gc * dim + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < S1_BK; k++) {
        float4 a = *reinterpret_cast<const float4*>(&Qs[k][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Cs[k][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) Ss[ty * 4 + i][tx * 4 + j] = acc[i][j];
    __syncthreads();
    // 64 threads: one row each, insert the tile's scores into the row's sorted shortlist
    if (tid < S1_BM) {
      float* v = tv[tid];
      int32_t* ix = ti[tid];
      int ncol = (int)min((int64_t)S1_BN, c_end - c0);
      for (int j = 0; j < ncol; j++) {
        float s = Ss[tid][j];
        int32_t cid = (int32_t)(c0 + j);
        if (s > v[CB_TOPR - 1]) {  // ids ascend, so equal scores keep the earlier (lower) id
          int p = CB_TOPR - 1;
          while (p > 0 && v[p - 1] < s) { v[p] = v[p - 1]; ix[p] = ix[p - 1]; p--; }
          v[p] = s; ix[p] = cid;
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < S1_BM * CB_TOPR; i += 256) {
    int r = i / CB_TOPR, j = i % CB_TOPR;
    int64_t gr = row0 + r;
    if (gr < nrows) {
      topv[(gr * nsplit + split) * CB_TOPR + j] = tv[r][j];
      topi[(gr * nsplit + split) * CB_TOPR + j] = ti[r][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 2. exact fixed-order re-score of the shortlist + top-nprobe decision.  One warp per row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fixed_order_dot(const float* __restrict__ q, const float* __restrict__ c, int dim) {
  float acc = 0.f;
  for (int k = 0; k < dim; k++) acc = __fadd_rn(acc, __fmul_rn(q[k], c[k]));
  return acc;
}

// better(a, b): a ranks before b under (score desc, id asc)
__device__ __forceinline__ bool s1_better(float sa, int32_t ia, float sb, int32_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}


__global__ void __launch_bounds__(128)
k_stage1_rescore(const float* __restrict__ Q, int64_t nrows, const float* __restrict__ C, int dim, int nsplit,
                 const float* __restrict__ topv, const int32_t* __restrict__ topi, const float* __restrict__ thr0,
                 int nprobe, float guard, float guard_rel, int32_t* __restrict__ cells, float* __restrict__ cell_scores, int32_t* __restrict__ flags) {
  int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const float* q = Q + row * dim;
  const int ncand = nsplit * CB_TOPR;
  constexpr int PER_LANE = CB_S1_SPLITS * CB_TOPR / 32;   // candidates per lane
  float sc[PER_LANE];
  int32_t id[PER_LANE];
  float excluded = -INFINITY;  // best approximate score any NON-shortlisted centroid can have
#pragma unroll
  for (int j = 0; j < PER_LANE; j++) {
    int ci = lane + 32 * j;
    sc[j] = -INFINITY;
    id[j] = 0x7fffffff;
    if (ci < ncand) {
      int32_t cid = topi[row * ncand + ci];
      if (cid < 0) cid = 0x7fffffff;   // a segment that does not exist for this row (stage1_tc.cu)
      if (cid != 0x7fffffff) {
        sc[j] = fixed_order_dot(q, C + (int64_t)cid * dim, dim);
        id[j] = cid;
      }
      // the last (worst) slot of each split bounds everything that split dropped
      if ((ci % CB_TOPR) == CB_TOPR - 1 && cid != 0x7fffffff) excluded = fmaxf(excluded, topv[row * ncand + ci]);
    }
  }
  // a unit that started from a published threshold dropped everything below it without listing it
  if (thr0 != nullptr && lane < nsplit) excluded = fmaxf(excluded, thr0[row * nsplit + lane]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) excluded = fmaxf(excluded, __shfl_xor_sync(0xffffffffu, excluded, o));

  if (guard_rel > 0.f) {  // tensor-core shortlist: the rounding bound scales with |q|
    float ss = 0.f;
    for (int k = lane; k < dim; k += 32) ss = fmaf(q[k], q[k], ss);
    guard += guard_rel * sqrtf(cb_warp_sum(ss));
  }
  float last = INFINITY;
  for (int p = 0; p < nprobe; p++) {
    // lane-local best
    float bs = -INFINITY; int32_t bi = 0x7fffffff; int bj = -1;
#pragma unroll
    for (int j = 0; j < PER_LANE; j++)
      if (id[j] != 0x7fffffff && (bj < 0 || s1_better(sc[j], id[j], bs, bi))) { bs = sc[j]; bi = id[j]; bj = j; }
    // warp argbest
    float ws = bs; int32_t wi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float os = __shfl_xor_sync(0xffffffffu, ws, o);
      int32_t oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (oi != 0x7fffffff && (wi == 0x7fffffff || s1_better(os, oi, ws, wi))) { ws = os; wi = oi; }
    }
    if (wi == 0x7fffffff) {  // fewer than nprobe centroids exist (K < nprobe): pad
      if (lane == 0) { cells[row * nprobe + p] = -1; cell_scores[row * nprobe + p] = -INFINITY; }
      continue;
    }
    if (bj >= 0 && bi == wi) {  // owner retires it (ids are unique per row)
#pragma unroll
      for (int j = 0; j < PER_LANE; j++) if (j == bj) id[j] = 0x7fffffff;
    }
    if (lane == 0) { cells[row * nprobe + p] = wi; cell_scores[row * nprobe + p] = ws; }
    last = ws;
  }
  if (lane == 0) flags[row] = (excluded > -INFINITY && !(last - excluded > guard)) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// 3. exact scan of all K centroids for flagged rows (rare).  One CTA per flagged row.
// ---------------------------------------------------------------------------------------------
// The number of flagged rows stays on the device (flagged[0]): the grid is fixed and CTA b takes the rows
// b, b + gridDim.x, ... of the list, so no host round trip sits between the shortlist pass and the scan.
__global__ void __launch_bounds__(256)
k_stage1_fullscan(const float* __restrict__ Q, const float* __restrict__ C, int64_t K, int dim, int nprobe,
                  const int32_t* __restrict__ flagged /* [0] = count, [1..] = rows */, int32_t* __restrict__ cells,
                  float* __restrict__ cell_scores, unsigned long long* __restrict__ stat_flagged) {
  extern __shared__ float s_q[];  // dim floats, then reduction scratch
  __shared__ float r_s[256];
  __shared__ int32_t r_i[256];
  const int tid = threadIdx.x;
  const int nflag = flagged[0];
  if (blockIdx.x == 0 && tid == 0 && stat_flagged != nullptr && nflag > 0) atomicAdd(stat_flagged, (unsigned long long)nflag);
  for (int fi = blockIdx.x; fi < nflag; fi += gridDim.x) {
  const int64_t row = flagged[1 + fi];
  __syncthreads();
  for (int k = tid; k < dim; k += 256) s_q[k] = Q[row * dim + k];
  __syncthreads();
  float ls[CB_MAX_NPROBE];
  int32_t li[CB_MAX_NPROBE];
  for (int p = 0; p < CB_MAX_NPROBE; p++) { ls[p] = -INFINITY; li[p] = 0x7fffffff; }
  for (int64_t c = tid; c < K; c += 256) {
    float s = fixed_order_dot(s_q, C + c * dim, dim);
    int32_t cid = (int32_t)c;
    if (li[nprobe - 1] == 0x7fffffff || s1_better(s, cid, ls[nprobe - 1], li[nprobe - 1])) {
      int p = nprobe - 1;
      while (p > 0 && (li[p - 1] == 0x7fffffff || s1_better(s, cid, ls[p - 1], li[p - 1]))) {
        ls[p] = ls[p - 1]; li[p] = li[p - 1]; p--;
      }
      ls[p] = s; li[p] = cid;
    }
  }
  int head = 0;  // next unconsumed entry of this thread's sorted list
  for (int p = 0; p < nprobe; p++) {
    r_s[tid] = head < nprobe ? ls[head] : -INFINITY;
    r_i[tid] = head < nprobe ? li[head] : 0x7fffffff;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) {
        float os = r_s[tid + o]; int32_t oi = r_i[tid + o];
        if (oi != 0x7fffffff && (r_i[tid] == 0x7fffffff || s1_better(os, oi, r_s[tid], r_i[tid]))) { r_s[tid] = os; r_i[tid] = oi; }
      }
      __syncthreads();
    }
    float ws = r_s[0]; int32_t wi = r_i[0];
    __syncthreads();
    if (head < nprobe && li[head] == wi && wi != 0x7fffffff) head++;
    if (tid == 0) {
      cells[row * nprobe + p] = (wi == 0x7fffffff) ? -1 : wi;
      cell_scores[row * nprobe + p] = (wi == 0x7fffffff) ? -INFINITY : ws;
    }
  }
  }
}

// compacts flagged row ids; count in out[0], ids in out[1..]
__global__ void k_compact_flags(const int32_t* __restrict__ flags, int64_t nrows, int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows && flags[i]) {
    int slot = atomicAdd(&out[0], 1);
    out[1 + slot] = (int32_t)i;
  }
}

int32_t cb_stage1_tc_shortlist(cb_index* ix, const float* dQ, int64_t nrows, float* topv, int32_t* topi, float* thr0,
                               int* nsplit_out, float* guard_rel_out, cudaStream_t st);  // stage1_tc.cu

int32_t cb_stage1_probe(cb_index* ix, const float* dQ, int64_t nrows, int nprobe, int32_t* d_cells,
                        float* d_scores, cudaStream_t st) {
  CB_REQUIRE(nprobe >= 1 && nprobe <= CB_MAX_NPROBE, CB_ERR_UNSUPPORTED, "nprobe must be in 1..%d (got %d)",
             CB_MAX_NPROBE, nprobe);
  if (nrows == 0) return CB_OK;
  int nsplit = CB_S1_SIMT_SPLITS;
  if (ix->K < 4096) nsplit = 1;
  CB_TRY(ix->topr_val.ensure(sizeof(float) * nrows * CB_S1_SPLITS * CB_TOPR));
  CB_TRY(ix->topr_idx.ensure(sizeof(int32_t) * nrows * CB_S1_SPLITS * CB_TOPR));
  CB_TRY(ix->flags.ensure(sizeof(int32_t) * (2 * nrows + 2)));
  float* topv = ix->topr_val.as<float>();
  int32_t* topi = ix->topr_idx.as<int32_t>();
  int32_t* flags = ix->flags.as<int32_t>();
  int32_t* flagged = flags + nrows;  // [0] = count, [1..] = row ids
  // the SIMT pass's fmaf chain and the fixed-order dot differ by <= ~2 dim 2^-24 |q||c|: scale with |q|, max|c|
  float guard = 1e-5f, guard_rel = 2.0f * (float)ix->dim * 5.9604645e-08f * ix->centroid_norm_max;

  bool used_tc = false;
  CB_TRY(ix->s1_thr0.ensure(sizeof(float) * nrows * CB_S1_SPLITS));
  float* thr0 = ix->s1_thr0.as<float>();
  if (ix->opt_stage1_impl != 1) {
    int32_t s = cb_stage1_tc_shortlist(ix, dQ, nrows, topv, topi, thr0, &nsplit, &guard_rel, st);
    if (s == CB_OK) { used_tc = true; ix->st_s1_tc_rows += (double)nrows; }
    else if (s != CB_ERR_UNSUPPORTED || ix->opt_stage1_impl == 2) return s;
  }
  if (!used_tc) {
    dim3 grid((unsigned)((nrows + S1_BM - 1) / S1_BM), nsplit);
    k_stage1_simt<<<grid, 256, 0, st>>>(dQ, nrows, ix->centroids, ix->K, ix->dim, nsplit, topv, topi);
    CB_LAUNCH_CHECK();
  }
  ix->s1_nsplit = nsplit; ix->s1_used_tc = used_tc ? 1 : 0; ix->s1_guard = guard; ix->s1_guard_rel = guard_rel;
  k_stage1_rescore<<<(unsigned)((nrows + 3) / 4), 128, 0, st>>>(dQ, nrows, ix->centroids, ix->dim, nsplit, topv, topi,
                                                               used_tc ? thr0 : nullptr, nprobe, guard, guard_rel, d_cells,
                                                               d_scores, flags);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemsetAsync(flagged, 0, sizeof(int32_t), st));
  k_compact_flags<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(flags, nrows, flagged);
  CB_LAUNCH_CHECK();
  // flagged rows are rare (a handful per batch on well-separated queries): a fixed, modest grid covers them
  const int64_t fs_grid = nrows < 2 * ix->sm_count ? nrows : 2 * ix->sm_count;
  k_stage1_fullscan<<<(unsigned)fs_grid, 256, sizeof(float) * ix->dim, st>>>(dQ, ix->centroids, ix->K, ix->dim, nprobe, flagged,
                                                                            d_cells, d_scores, cb_stats_dev(ix) + 2);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

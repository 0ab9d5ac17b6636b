// common.cuh -- shared state, error plumbing and small device helpers of libcolbert_b200.so.
// sm_100a only.  Nothing here is a CPU fallback: every compute entry point needs a device.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/colbert_b200.h"

// ---------------------------------------------------------------------------------------------
// limits of this build
// ---------------------------------------------------------------------------------------------
constexpr int CB_MAX_NBITS = 8;        // reference allows any nbits; BASELINE needs 1/2/4
constexpr int CB_NQ_CHUNK = 1024;      // queries scored per pass (bitmap row = 1024 bits = 128 B)
constexpr int CB_TOPR = 16;            // approximate per-token shortlist kept by stage 1
constexpr int CB_MAX_NPROBE = 12;      // nprobe + margin must fit CB_TOPR
constexpr int CB_MAX_K = 1024;         // top-k selection sorts its k winners in shared memory
constexpr int CB_S1_SPLITS = 16;       // most centroid-range segments one query token's shortlists can come in (tcgen05 stage 1)
constexpr int CB_S1_SIMT_SPLITS = 8;   // centroid-range splits of the SIMT stage-1 kernel

// ---------------------------------------------------------------------------------------------
// error plumbing (never throw across the ABI)
// ---------------------------------------------------------------------------------------------
void cb_set_error(const char* fmt, ...);
extern thread_local long long g_cb_launches;  // kernels launched by this thread (gpu_launches)

#define CB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      cb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                   cudaGetErrorString(_e));                                                \
      return (_e == cudaErrorMemoryAllocation) ? CB_ERR_OOM : CB_ERR_CUDA;                 \
    }                                                                                      \
  } while (0)

#define CB_LAUNCH_CHECK()                 \
  do {                                    \
    g_cb_launches++;                      \
    CB_CUDA(cudaGetLastError());          \
  } while (0)

#define CB_TRY(expr)                \
  do {                              \
    int32_t _s = (expr);            \
    if (_s != CB_OK) return _s;     \
  } while (0)

#define CB_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      cb_set_error(__VA_ARGS__);     \
      return (code);                 \
    }                                \
  } while (0)

// A grow-only device buffer (workspace).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int32_t ensure(size_t bytes) {
    if (bytes <= cap) return CB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      cb_set_error("device allocation of %zu bytes failed", bytes);
      return CB_ERR_OOM;
    }
    cap = want;
    return CB_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------------------------------------
// the index handle
// ---------------------------------------------------------------------------------------------
struct cb_index {
  int device = 0;
  int dim = 0, nbits = 0, R = 0;  // R = dim/8*nbits bytes per embedding
  int64_t K = 0, Np = 0, Ne = 0, pid_base = 0;
  int sm_count = 148;

  // resident index (device)
  float* centroids = nullptr;        // [K][dim] fp32 (exact paths)
  __half* centroids_h = nullptr;     // [K][dim] fp16 (fused fast path gather)
  uint8_t* centroids_img = nullptr;  // fp16 row image (swizzled MMA operand tiles) for the tcgen05 stage 1; dim = 128 only
  float centroid_norm_max = 1.f;     // max |centroid| (scales the fp16 rounding guard of stage 1)
  float weight_abs_max = 0.f;        // max |bucket weight| (with the above: range precondition of the packed-fp16 decompression)
  float* weights = nullptr;          // [2^nbits]
  int32_t* codes = nullptr;          // [Ne] 0-based
  uint8_t* residuals = nullptr;      // [Ne][R]
  bool residuals_borrowed = false;   // CB_FLAG_BORROW_RESIDUALS: the caller's array, not freed here
  int64_t* offsets = nullptr;        // [Np+1] exclusive prefix sum of doclens
  int64_t* cell_offsets = nullptr;   // [K+1]  exclusive prefix sum of ivf_lengths
  int32_t* ivf_pids = nullptr;       // [Ne]   local 0-based pid of every IVF entry
  size_t resident_bytes = 0;
  int64_t max_doclen = 0;
  int64_t max_cell_len = 0;          // longest IVF cell: bounds the pair list of a batch without a host round trip
  int32_t n_long = -1, long_limit = 0;  // passages too long for the tcgen05 tile (cached list)

  // workspace (grow-only)
  DevBuf q_f32, q_prep, topr_val, topr_idx, cells, cell_scores, flags, bitmap, counts, list_off,
      cursors, pairs, out_pids, out_scores, out_counts, misc, long_list, hook_a, hook_b, hook_c, s1_thr, s1_thr0,
      bitmap_t, bitmap2, q_flag, fin_keys, fin_pids, fin_scores, pl_ents, pl_misc, pl_vec, pl_next, pl_head, pl_mask, pl_active, pl_top_pids, pl_top_scores, pl_sel, pl_npos;   // PLAID mode (plaid.cu)
  int64_t* pinned_total = nullptr;  // pinned host scalars (synchronous paths: hooks, PLAID mode)
  // Batch counters live on the device ([0] pairs, [1] pair embeddings, [2] stage-1 rows re-done by the exact
  // scan, [3] query batches outside the tcgen05 kernel's range); they are copied to pinned memory at the end
  // of a batch and read lazily by cb_get_stat, so the search itself never waits for the host.
  DevBuf d_stats;
  unsigned long long* pinned_stats = nullptr;
  cudaEvent_t ev_stats = nullptr;
  bool stats_pending = false;
  int stats_tc_selected = 0;
  const float* q_prep_src = nullptr; // q_prep currently holds the row image of these query tokens ...
  int64_t q_prep_rows = 0;           // ... (this many rows); reset at the start of every search chunk
  // what the last cb_stage1_probe left in topr_val / topr_idx / s1_thr0 (read by the PLAID survivor pass)
  const int32_t* tc_active_list = nullptr;   // optional passage list for the tcgen05 scoring kernel (sparse bitmaps)
  int64_t tc_active_n = 0;
  int s1_nsplit = 1, s1_used_tc = 0;
  float s1_guard = 1e-5f, s1_guard_rel = 0.f;

  // options
  int opt_force_generic = 0;
  int opt_stage1_impl = 0;
  int opt_profile = 0;
  int opt_tc_astages = 0;   // query-tile stages of the tcgen05 scoring kernel (0 = default)
  int opt_sync_pairs = 0;   // 1 = size the pair list with a host round trip (exact) instead of the IVF bound
  int opt_exact_rescore = 1;   // final top-k decided on exact fp32 scores of the best 2k+ tensor-core candidates

  // stats of the last search
  long long st_launches = 0;
  double st_tc_groups = 0, st_tc_group_rows = 0, st_tc_passage_rows = 0;
  double st_pairs = 0, st_pair_embs = 0, st_flagged = 0, st_tc_pairs = 0, st_generic_pairs = 0, st_s1_tc_rows = 0, st_rescore_unsafe = 0, st_bad_cells = 0;
  double st_ms[5] = {0, 0, 0, 0, 0};  // stage1, stage2, stage34, stage5, total
  double st_plaid_survivors = 0, st_plaid_positive = 0, st_plaid_rescored = 0;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// ---------------------------------------------------------------------------------------------
// stage launchers (defined in the stage*.cu files)
// ---------------------------------------------------------------------------------------------
constexpr size_t CB_STATS_BYTES = 128;   // 16 batch counters: 0 pairs, 1 pair embeddings, 2 flagged rows, 3 range flag, 4 bad cells, 5 rescore-unsafe,
                                         // 6 tcgen05 groups, 7 operand rows summed over groups, 8 operand rows summed over decompressed passages
inline unsigned long long* cb_stats_dev(cb_index* ix) { return ix->d_stats.as<unsigned long long>(); }

// Stage 1: per query-token row, the top-`nprobe` centroids of Q . C^T (exact fp32 decision).
// d_cells int32[nrows][nprobe] 0-based; d_scores float[nrows][nprobe]; flagged count to stats.
int32_t cb_stage1_probe(cb_index* ix, const float* dQ, int64_t nrows, int nprobe, int32_t* d_cells,
                        float* d_scores, cudaStream_t st);

// Stage 2: marks bitmap[pid][W] for every passage in a probed cell; counts[q] = #candidates.
int32_t cb_stage2_mark(cb_index* ix, const int32_t* d_cells, int nq, int T, int nprobe, int W,
                       uint32_t* d_bitmap, int32_t* d_counts, cudaStream_t st);

// Stages 1+2 for one chunk of <= CB_NQ_CHUNK queries (search.cu): leaves bitmap / counts / list_off / zeroed
// cursors in the workspace and returns the number of (query, passage) pairs of the chunk.
// total_pairs == nullptr: no host round trip (the caller sizes the pair list from cb_pair_bound).
// d_cells_in != nullptr: stage 1 is skipped, the chunk's cells ([nq][T][nprobe], 0-based, -1 = none) are given.
int32_t cb_candidates_chunk(cb_index* ix, const float* dQ, int nq, int T, int nprobe, int W, cudaStream_t st,
                            int64_t* total_pairs, const int32_t* d_cells_in = nullptr);

// Exclusive scan of counts -> list offsets (+ total in list_off[nq]).
int32_t cb_scan_counts(const int32_t* d_counts, int nq, int64_t* d_list_off, cudaStream_t st,
                       unsigned long long* d_stat_pairs = nullptr);

// Stages 3+4: for every (passage, candidate query) pair appends a 64-bit key
// (orderable score << 32 | ~local_pid) to the query's list.
int32_t cb_stage34_score(cb_index* ix, const float* dQ, int nq, int T, int W,
                         const uint32_t* d_bitmap, const int64_t* d_list_off, int32_t* d_cursors,
                         uint64_t* d_pairs, cudaStream_t st);
// d_gate != nullptr: the kernel does nothing unless *d_gate != 0 (device-side routing, no host round trip).
int32_t cb_stage34_generic(cb_index* ix, const float* dQ, int nq, int T, int W,
                           const uint32_t* d_bitmap, const int32_t* d_pid_list, int64_t n_list,
                           const int64_t* d_list_off, int32_t* d_cursors, uint64_t* d_pairs,
                           cudaStream_t st, const int* d_gate = nullptr);
int32_t cb_stage34_tc(cb_index* ix, const float* dQ, int nq, int T, int W,
                      const uint32_t* d_bitmap, const int64_t* d_list_off, int32_t* d_cursors,
                      uint64_t* d_pairs, cudaStream_t st);
bool cb_stage34_tc_supported(const cb_index* ix, int T);
// fp32 rows [nrows][128] -> fp16 "row image" [nrows_pad/8][2 K-blocks][8 rows][128 B] (SWIZZLE_128B
// K-major, 2048 B between 8-row groups): the shared-memory operand layout of both tcgen05 kernels.
// d_range_flag (optional): set to 1 when a row breaks the |row| <= 255 precondition of the tcgen05 scoring kernel.
int32_t cb_tc_prep_rows(const float* dX, int64_t nrows, int64_t nrows_pad, uint8_t* d_out, cudaStream_t st, int* d_range_flag = nullptr);

// Stage 5: first k of each query's list by key descending -> 1-based global pids / scores.
int32_t cb_stage5_topk(const uint64_t* d_pairs, const int64_t* d_list_off, int nq, int k,
                       int64_t pid_base, int64_t* d_out_pids, float* d_out_scores,
                       cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// float -> uint32 whose unsigned order equals the float order (isless semantics, -0 < +0).
__host__ __device__ __forceinline__ uint32_t cb_orderable(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float cb_unorderable(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// Pair key: larger key == better rank under (score desc, pid asc).
__host__ __device__ __forceinline__ uint64_t cb_pair_key(float score, uint32_t local_pid) {
  return ((uint64_t)cb_orderable(score) << 32) | (uint64_t)(0xffffffffu - local_pid);
}

// Bucket index (0-based) of dimension d inside one packed embedding (`_unpackbits` +
// `_unbinarize`, src/indexing/codecs/residual.jl:233-240,428-441): flat bit d*nbits+b sits in
// byte (d*nbits+b)>>3 at bit position (d*nbits+b)&7 (LSB first).  Generic nbits in 1..8.
__device__ __forceinline__ uint32_t cb_bucket_of(const uint8_t* __restrict__ emb, int d, int nbits) {
  int bit = d * nbits;
  int byte = bit >> 3, sh = bit & 7;
  uint32_t v = emb[byte];
  if (sh + nbits > 8) v |= (uint32_t)emb[byte + 1] << 8;
  return (v >> sh) & ((1u << nbits) - 1u);
}

__device__ __forceinline__ float cb_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

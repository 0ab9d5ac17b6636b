/*
 * colbert_b200.h -- C ABI of libcolbert_b200.so: a B200 (sm_100a) implementation of the
 * search-time scoring path of JuliaGenAI/ColBERT.jl.
 *
 * This is the drop-in boundary.  Every entry point takes plain pointers and sizes (no torch /
 * CUDA types in the signatures; a stream is passed as an opaque void*), returns an int32
 * status and never throws.  Each entry point names the reference function it replaces
 * (file:line relative to the ColBERT.jl source tree).  The Julia binding a maintainer would
 * add is in INTEGRATION.md / colbert.jl_b200/julia/ColBERTB200.jl; the Python (ctypes)
 * binding used by the tests is colbert.jl_b200/_lib.py.
 *
 * Array layouts are exactly the memory layouts of the reference's Julia arrays (column-major),
 * i.e. a Julia Matrix{T}(a, b) is the C array T[b][a]:
 *   centroids      Matrix{Float32}(dim, K)        -> const float[K][dim]
 *   bucket_weights Vector{Float32}(2^nbits)       -> const float[2^nbits]
 *   codes          Vector{UInt32}(N_e), 1-BASED   -> const uint32_t[N_e]
 *   residuals      Matrix{UInt8}(dim/8*nbits,N_e) -> const uint8_t[N_e][dim/8*nbits]
 *   doclens        Vector{Int}(N_p)               -> const int64_t[N_p]
 *   ivf            Vector{Int}(N_e), 1-BASED eids -> const int64_t[N_e]
 *   ivf_lengths    Vector{Int}(K)                 -> const int64_t[K]
 *   Q              Array{Float32,3}(dim, T, nq)   -> const float[nq][T][dim]
 * All pids crossing the ABI are 1-BASED like the reference's.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and fails with
 * CB_ERR_CUDA when none is usable.
 */
#ifndef COLBERT_B200_H
#define COLBERT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (the Julia shim maps them onto the reference's exception types) ---- */
#define CB_OK 0
#define CB_ERR_BAD_ARG 1     /* shape / argument mismatch      -> DimensionMismatch (ranking.jl:9-12,71-74) */
#define CB_ERR_DOMAIN 2      /* value out of its valid domain  -> DomainError (residual.jl:701-706,763-768) */
#define CB_ERR_CUDA 3        /* CUDA runtime / launch failure, or no device */
#define CB_ERR_OOM 4         /* device or host allocation failed */
#define CB_ERR_UNSUPPORTED 5 /* legal in the reference but outside this build's limits */
#define CB_ERR_BOUNDS 6      /* fewer results than requested   -> BoundsError (searching.jl:127) */

/* flags for cb_index_create */
#define CB_FLAG_DEVICE_POINTERS 1 /* all array arguments are device pointers on `device` */
#define CB_FLAG_BORROW_RESIDUALS 2 /* with CB_FLAG_DEVICE_POINTERS: the index reads the caller's `residuals` array in place instead of
                                      copying it (the bulk of an index: 19 GB at 8.8 M passages); the caller keeps it alive and
                                      unmodified until cb_index_destroy, which does not free it */

typedef struct cb_index cb_index; /* opaque: one index shard resident in the HBM of one GPU */
typedef struct cb_multi cb_multi; /* opaque: the passage-range shards of one index on several GPUs of one box */

/* Library version string, e.g. "colbert_b200 0.1 (sm_100a)". */
const char* cb_version(void);

/* Message of the last failing call made by this thread ("" if none). */
const char* cb_last_error(void);

/* Number of CUDA devices visible (0 when there is no usable driver/device). */
int32_t cb_device_count(void);

/*
 * Uploads one index (or one passage-range shard of it) to `device` and builds the derived,
 * search-time structures (0-based codes, passage offsets, per-cell passage lists).
 * Replaces the host-resident `struct Searcher` state and its constructor's loads
 * (src/searching.jl:1-16, 44-59) plus `_build_emb2pid` (src/searching.jl:82-91).
 *   ivf / ivf_lengths may both be NULL: the IVF is then built on the device from `codes`
 *     with the semantics of `_build_ivf` (src/indexing/collection_indexer.jl:349-353).
 *   pid_base: added to every local pid on output (shard r of a passage-sharded index passes
 *     the number of passages in shards 0..r-1; 0 for an unsharded index).
 * Validates like the reference: codes in 1:K (residual.jl:766), sum(doclens) == N_e
 * (loaders.jl:86-87), sum(ivf_lengths) == N_e (ranking.jl:11), dim % 8 == 0, 1 <= nbits <= 8.
 * The library copies everything it needs; host pointers are not retained.
 */
int32_t cb_index_create(cb_index** out, int32_t device, int32_t dim, int32_t nbits, int64_t K,
                        int64_t n_passages, int64_t n_embeddings, const float* centroids,
                        const float* bucket_weights, const uint32_t* codes,
                        const uint8_t* residuals, const int64_t* doclens, const int64_t* ivf,
                        const int64_t* ivf_lengths, int64_t pid_base, int32_t flags);

/*
 * Opens an index DIRECTORY written by the reference (`index(indexer)`, src/indexing.jl:63-147; files by src/savers.jl:
 * config.json, plan.json, centroids / bucket_weights .jld2, ivf / ivf_lengths .jld2, doclens.<c>.jld2, <c>.codes.jld2,
 * <c>.residuals.jld2) and uploads it: the loading half of `Searcher(index_path)` (src/searching.jl:18-59;
 * `load_codec` src/loaders.jl:10-38, `load_doclens` 76-89, `load_compressed_embs` 91-113) with a native reader of the
 * JLD2 0.4 container (csrc/jld2.h) -- the files are mapped and copied to the device, no Julia object is built.
 *   shard / n_shards: open only passage range `shard` of `n_shards` ranges balanced by embedding count; only the
 *     chunk files overlapping that range are read, the shard's IVF is rebuilt on the device from its own codes, and
 *     its pids are global (pid_base = first passage of the range, also returned through out_pid_base if non-NULL).
 * Same validation and messages as the reference's loaders (type asserts loaders.jl:27-30, sum(doclens) == num_embeddings
 * loaders.jl:86-88); compressed (chunked / filtered) JLD2 datasets -> CB_ERR_BAD_ARG.
 */
int32_t cb_index_open(cb_index** out, const char* index_path, int32_t device, int32_t shard, int32_t n_shards,
                      int64_t* out_pid_base);

/* Host-only helper of the above (no GPU needed): reads dataset `name` (NULL = "single_stored_object", what
 * `JLD2.save_object` writes) of one JLD2 file.  info[0] = element type (1 f32, 2 f64, 3 i8, 4 u8, 5 i16, 6 u16, 7 i32,
 * 8 u32, 9 i64, 10 u64), info[1] = element size, info[2] = number of dimensions, info[3..10] = dimensions as stored
 * = the C-layout shape (Julia's size() reversed).  Copies min(capacity_bytes, total bytes) into out (may be NULL). */
int32_t cb_jld2_read(const char* path, const char* name, int64_t info[11], void* out, int64_t capacity_bytes);

int32_t cb_index_destroy(cb_index* index);

/* info[0..7] = dim, nbits, K, n_passages, n_embeddings, device, pid_base, device bytes held. */
int32_t cb_index_info(const cb_index* index, int64_t info[8]);

/*
 * Tuning / test knobs.  Keys: "force_generic" (1 = use only the generic SIMT scoring kernel),
 * "stage1_impl" (0 = auto, 1 = SIMT fp32, 2 = tcgen05), "profile" (1 = record per-stage CUDA
 * event timings, readable through cb_get_stat), "tc_astages" (query-tile pipeline stages of the
 * tcgen05 scoring kernel, 2..6; 0 = default), "sync_pairs" (1 = size the pair list exactly with a
 * host round trip instead of the IVF bound), "exact_rescore" (default 1: the final top-k is decided
 * on exact fp32 scores of the best max(2k, k+16) tensor-core candidates; 0 = rank by tensor-core
 * score).  Unknown key -> CB_ERR_BAD_ARG.
 */
int32_t cb_set_option(cb_index* index, const char* key, int64_t value);

/*
 * Counters of the most recent search call.  Keys: "launches" (kernels launched), "pairs"
 * ((query, candidate passage) pairs scored), "pair_embeddings" (sum of doclens over pairs),
 * "flagged_rows" (query tokens whose top-nprobe needed the exact full scan),
 * "ms_stage1", "ms_stage2", "ms_stage34", "ms_stage5", "ms_total" (need option "profile"),
 * "tc_pairs" / "generic_pairs" (pairs scored by the tcgen05 / the generic kernel),
 * "stage1_tc_rows" (query tokens whose centroid shortlist came from the tcgen05 stage-1 kernel),
 * "rescore_unsafe" (queries whose exact re-score margin could not rule out a missed candidate),
 * "bad_cells" (caller-supplied cells outside 0:K), "max_cell_len" (longest IVF cell).
 * Reading a counter of an asynchronous (device) search waits for that batch to finish.
 */
int32_t cb_get_stat(const cb_index* index, const char* key, double* value);

/*
 * THE hot path: `search` minus the BERT encoder (src/searching.jl:103-127), batched.
 * For every query q: candidates = retrieve(...) (src/search/ranking.jl:23-44, no candidate
 * cap), score = maxsim(Q, decompress(collect(candidates))) (ranking.jl:46-86,
 * residual.jl:759-784), results = first k of the stable descending sort
 * (searching.jl:125-127; ties -> ascending pid).
 *   Q          host  float[nq][T][dim]
 *   out_pids   host  int64[nq][k]   1-based (+ pid_base); unfilled slots are 0
 *   out_scores host  float[nq][k]   unfilled slots are -inf
 *   out_counts host  int32[nq]      number of candidate passages of q on this shard; the
 *                                   reference throws BoundsError when it is < k
 * Synchronous: outputs are complete on return.  Not re-entrant on one handle.
 */
int32_t cb_search_batch(cb_index* index, const float* Q, int32_t nq, int32_t T, int32_t nprobe,
                        int32_t k, int64_t* out_pids, float* out_scores, int32_t* out_counts);

/* Same, with Q and the three outputs already in device memory of the index's GPU and all work
 * enqueued on `stream` (a cudaStream_t; NULL = default stream).  Returns after enqueueing the
 * last kernel; the caller synchronises the stream.  The batch is a pure stream of launches: no
 * host round trip inside (the pair list is sized from an IVF bound, flagged stage-1 rows and
 * out-of-range query batches are routed on the device, counters are read lazily by cb_get_stat). */
int32_t cb_search_batch_device(cb_index* index, const float* dQ, int32_t nq, int32_t T,
                               int32_t nprobe, int32_t k, int64_t* d_out_pids,
                               float* d_out_scores, int32_t* d_out_counts, void* stream);

/* Stage 1 alone, device-resident (the piece a passage-sharded deployment splits by QUERY: every shard holds
 * the same centroids, so rank r probes queries [r nq/N, (r+1) nq/N) and the ranks exchange the cells):
 * `_topk(Q' * centroids, nprobe, dims = 2)` (src/search/ranking.jl:27-31, src/utils.jl:327-332).
 *   dQ device float[nq][T][dim]; d_out_cells device int32[nq][T][nprobe], 1-based centroid ids, best first
 *   (0 = none: fewer than nprobe centroids exist).  Enqueued on `stream`. */
int32_t cb_probe_device(cb_index* index, const float* dQ, int32_t nq, int32_t T, int32_t nprobe,
                        int32_t* d_out_cells, void* stream);

/* cb_search_batch_device with stage 1 supplied by the caller: d_cells device int32[nq][T][nprobe] as written by
 * cb_probe_device (on this or any other shard of the same index).  Entries outside 0:K are ignored and counted
 * (stat "bad_cells"). */
int32_t cb_search_batch_cells_device(cb_index* index, const float* dQ, const int32_t* d_cells, int32_t nq,
                                     int32_t T, int32_t nprobe, int32_t k, int64_t* d_out_pids,
                                     float* d_out_scores, int32_t* d_out_counts, void* stream);

/* ---- several GPUs, one process, one host thread (SURVEY 8b "Threading"; the reference is single-process:
 * src/infra/config.jl:57-58 carries rank / nranks but `search` never uses them) ----
 * cb_multi_create groups shard handles (passage-range shards of ONE index, each created with its pid_base on its
 * own device -- cb_index_create, or cb_index_open(path, device, r, n)); the group borrows them.  cb_multi_open opens
 * the n shards of an index directory itself (device_ids NULL = devices 0..n-1) and owns them.  Two shards may share
 * a device (useful for tests). */
int32_t cb_multi_create(cb_multi** out, int32_t n_shards, cb_index* const* shards);
int32_t cb_multi_open(cb_multi** out, const char* index_path, int32_t n_gpus, const int32_t* device_ids);
int32_t cb_multi_destroy(cb_multi* multi);
/* *n_shards = number of shards; the first min(capacity, n) handles are written to shards (may be NULL). */
int32_t cb_multi_info(const cb_multi* multi, int32_t* n_shards, cb_index** shards, int32_t capacity);

/* `search` minus the encoder over the whole sharded index (src/searching.jl:103-127), host buffers as cb_search_batch:
 * Q is uploaded once and forwarded to the peers over NVLink, stage 1 is split by query across the devices, every
 * device scores its own passage range, the per-shard top-k lists are merged on the first device by (score desc, pid
 * asc).  Results are bit-identical to cb_search_batch on the unsharded index.  out_counts[q] = candidates over all shards. */
int32_t cb_multi_search_batch(cb_multi* multi, const float* Q, int32_t nq, int32_t T, int32_t nprobe, int32_t k,
                              int64_t* out_pids, float* out_scores, int32_t* out_counts);

/* PLAID-style pruned search (BASELINE.json config 5).  NOT a reference function: ColBERT.jl lists PLAID
 * pruning as roadmap (README.md:187); the semantics are defined by oracle/oracle.py `plaid_search`
 * on top of the reference's own `retrieve` / `decompress` / `maxsim`:
 *   candidates = `retrieve` with nprobe = ncells (src/search/ranking.jl:23-44); a centroid survives
 *   for a query iff max_t Q[:,t].c >= centroid_score_threshold; approximate score of a candidate =
 *   sum_t max(0, max over its tokens with a surviving code of Q[:,t].centroid[code]); the first
 *   `ndocs` candidates under (approximate score desc, pid asc) get the exact fused
 *   decompress + MaxSim and the usual stable top-k.
 * Arguments as cb_search_batch; 1 <= ncells <= 12, 1 <= ndocs <= 1024, 1 <= T <= 32.
 * out_counts[q] = min(ndocs, #candidates): the number of exactly scored passages. */
int32_t cb_search_batch_plaid(cb_index* index, const float* Q, int32_t nq, int32_t T, int32_t ncells,
                              float centroid_score_threshold, int32_t ndocs, int32_t k,
                              int64_t* out_pids, float* out_scores, int32_t* out_counts);
int32_t cb_search_batch_plaid_device(cb_index* index, const float* dQ, int32_t nq, int32_t T,
                                     int32_t ncells, float centroid_score_threshold, int32_t ndocs,
                                     int32_t k, int64_t* d_out_pids, float* d_out_scores,
                                     int32_t* d_out_counts, void* stream);

/* ---- stage-level hooks (each mirrors one reference function; host buffers) ---- */

/* Stage 1: `_topk(Q' * centroids, nprobe, dims = 2)` (src/search/ranking.jl:27-31,
 * src/utils.jl:327-332).  out_cells int32[nq][T][nprobe] 1-based centroid ids, best first,
 * ties -> lower id; out_scores float[nq][T][nprobe] (may be NULL). */
int32_t cb_probe(cb_index* index, const float* Q, int32_t nq, int32_t T, int32_t nprobe,
                 int32_t* out_cells, float* out_scores);

/* Stages 1+2: `retrieve` (src/search/ranking.jl:23-44) for ONE query Q float[T][dim]:
 * sorted ascending unique 1-based candidate pids.  *out_count is always the true count; at most
 * `capacity` pids are written (out_pids may be NULL when capacity == 0). */
int32_t cb_retrieve(cb_index* index, const float* Q, int32_t T, int32_t nprobe,
                    int64_t* out_pids, int64_t capacity, int64_t* out_count);

/* Stage 3 alone: `decompress` (src/indexing/codecs/residual.jl:759-784) in exact fp32.
 *   codes uint32[n] 1-based, residuals uint8[n][dim/8*nbits]
 *   out_embs float[n][dim]; out_bucket_idx uint8[n][dim] 0-based bucket indices (NULL to skip;
 *   `_unbinarize(_unpackbits(..))`, residual.jl:709-710); out_unnormalized float[n][dim] =
 *   centroids[:,code] + w[bucket] before `_normalize_array!` (NULL to skip).
 * CB_ERR_DOMAIN when a code is outside 1:K (residual.jl:766). */
int32_t cb_decompress(int32_t device, int32_t dim, int32_t nbits, int64_t K,
                      const float* centroids, const float* bucket_weights,
                      const uint32_t* codes, const uint8_t* residuals, int64_t n,
                      float* out_embs, uint8_t* out_bucket_idx, float* out_unnormalized);

/* The writer side of the codec: `compress(centroids, bucket_cutoffs, dim, nbits, embs)` (src/indexing/codecs/residual.jl:586-604)
 * = `compress_into_codes!` (67-81: argmax over centroids of emb . c, first maximum wins) + `binarize` (518-536:
 * `_bucket_indices` 348-351, `_binarize` 197-208, `_packbits` 400-407) of emb - centroids[:, code].
 *   centroids float[K][dim], bucket_cutoffs float[2^nbits - 1] (ascending), embs float[n][dim]
 *   out_codes uint32[n] 1-based, out_residuals uint8[n][dim/8*nbits]
 * The codes come from the stage-1 tensor-core GEMM with nprobe = 1 and are decided on fixed-order fp32 dot products;
 * given the codes the packed bytes are bit-exact (`decompress_residuals(binarize(x))` inverts, residual.jl test 975-991). */
int32_t cb_compress(int32_t device, int32_t dim, int32_t nbits, int64_t K, const float* centroids,
                    const float* bucket_cutoffs, const float* embs, int64_t n, uint32_t* out_codes,
                    uint8_t* out_residuals);

/* Stage 4 alone: `maxsim(Q, D, pids, doclens)` (src/search/ranking.jl:69-86) in fp32.
 *   Q float[T][dim], D float[M][dim], pids int64[n_pids] 1-based, doclens int64[n_doclens].
 * CB_ERR_BAD_ARG when sum(doclens[pids]) != M (ranking.jl:71-74). */
int32_t cb_maxsim(int32_t device, int32_t dim, int32_t T, const float* Q, const float* D,
                  int64_t M, const int64_t* pids, int64_t n_pids, const int64_t* doclens,
                  int64_t n_doclens, float* out_scores);

/* Stages 3+4 fused, in place on the resident index, for ONE query and an explicit pid list
 * (1-based local pids + pid_base): out_scores[i] = maxsim score of pids[i].  Equivalent to
 * `_collect_compressed_embs_for_pids` -> `decompress` -> `maxsim` (ranking.jl:46-86). */
int32_t cb_score_pids(cb_index* index, const float* Q, int32_t T, const int64_t* pids,
                      int64_t n_pids, float* out_scores);

/* Parity hook for the FUSED path: what the decompression stage of the tcgen05 scoring kernel writes into its
 * shared-memory operand tiles for the listed passages (1-based pids + pid_base), produced by the very device
 * functions the hot kernel runs (`decompress`, src/indexing/codecs/residual.jl:759-784, with `_unpackbits` /
 * `_unbinarize` / `bucket_weights[idx]` 698-721 and `_normalize_array!`, src/utils.jl:320-325, in packed fp16):
 *   out_norm uint16 (IEEE fp16 bits) [sum of doclens][dim]  normalised operand rows, passage after passage
 *   out_raw  uint16 (IEEE fp16 bits) [sum of doclens][dim]  fp16(centroid) + fp16(w[bucket]) before normalisation --
 *            exactly reproducible on the host, so it pins every unpacked bucket index of that code path.
 * capacity_rows = rows both buffers hold.  CB_ERR_UNSUPPORTED when the index shape (dim != 128, nbits not in
 * {1,2,4}) or a listed passage (> 480 tokens) is not taken by that kernel. */
int32_t cb_debug_tc_operand(cb_index* index, const int64_t* pids, int64_t n_pids, uint16_t* out_norm,
                            uint16_t* out_raw, int64_t capacity_rows);

/* Stage 5 across shards: merges `n_lists` per-shard result lists (host, each [nq][k], unfilled
 * slots pid 0 / -inf) into the global first-k by (score desc, pid asc) -- the order the
 * reference's stable `sortperm(scores, rev=true)` over ascending pids produces
 * (src/searching.jl:125-127).  Runs on `device`. */
int32_t cb_merge_topk(int32_t device, int32_t n_lists, int32_t nq, int32_t k,
                      const int64_t* pids, const float* scores, int64_t* out_pids,
                      float* out_scores);

/* Same on device buffers ([n_lists][nq][k] as produced by an all-gather), enqueued on stream. */
int32_t cb_merge_topk_device(int32_t device, int32_t n_lists, int32_t nq, int32_t k,
                             const int64_t* d_pids, const float* d_scores, int64_t* d_out_pids,
                             float* d_out_scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COLBERT_B200_H */

// stage34_tc.cu -- the hot kernel: residual decompression fused into MaxSim on the 5th-generation
// tensor cores (tcgen05 + TMEM), passage-major over a whole query batch.
//
// What it replaces, per (query q, candidate passage p) pair, with nothing materialised:
//   `_collect_compressed_embs_for_pids` (src/search/ranking.jl:46-67), `decompress`
//   (src/indexing/codecs/residual.jl:759-784), the `Q' * D` sgemm and the per-pid
//   `sum(maximum(..., dims = 2))` loop of `maxsim` (ranking.jl:76-84).
//
// Design (B200-first, not a translation):
//   * Passage-major.  A persistent CTA takes one passage at a time, decompresses it ONCE into a
//     shared-memory operand tile (fp16, UMMA K-major SWIZZLE_128B layout) and scores it against
//     every query of the batch that holds it as a candidate (its 1024-bit bitmap row).  The packed
//     index therefore streams from HBM once per batch instead of once per (query, passage) pair,
//     and decompression cost is amortised over ~17-113 queries.
//   * MMA orientation: A = 4 queries x 32 tokens (M = 128 rows; the batch's fp16 "row image" is
//     L2-resident and a query is one contiguous 8 KB block of it, fetched by 1-D bulk async
//     copies), B = passage tokens (N = doclen padded to 16, <= 240 per chunk), K = dim = 128.
//     The fp32 accumulator D[128 x N] lives in TMEM; each epilogue thread owns one (query, token)
//     row, so "max over document tokens" is an in-register max over its TMEM columns and "sum over
//     query tokens" is one warp reduction.  Padded token rows duplicate the last real token, so
//     no column masking is needed.  Passages of 241..480 tokens are two chunks (two accumulators,
//     one running maximum).
//   * What bounds it (measured): the SM's 128 B/clk shared-memory / L1 data pipe.  A 4-query group moves
//     ~95 KB through it (32 KB bulk-copy write + 32 KB operand read of the query tile, 8 x N x 32 B of passage tile, ~25 KB of
//     decompression loads / stores / table reads) = ~750 clocks; the kernel runs at ~900 clocks per group (83 % of the pipe),
//     with the SM clock at 1.63-1.67 GHz under the 1000 W power cap.  In isolation (tools/pipe_skeleton.cu) the same MMAs,
//     copies and accumulator reads take 416-550 clocks per group; getting from ~1040 (one MMA issuer, one epilogue set) to ~900
//     took doubling every serial role at once: each of them alone sat at ~800-1000 clocks of its own latency per group.
//   * Warp-specialised, mbarrier-pipelined, 22 warps: warp 0 = scheduler (candidate lists, ring allocation), warps 1 and 12 =
//     MMA issuers (one thread each, alternate groups, one accumulator each), warps 2-3 = query-tile loaders, warps 4-7 and 8-11
//     = two epilogue sets (one TMEM lane quarter per warp; set s reads issuer s's accumulator), warp 13 idle, warps 14-21 =
//     decompression (two teams on alternate passages).  Pipelines: passage entries (4 meta slots + a variable-size
//     shared-memory ring of operand tiles), query tiles (3 stages of 32 KB), TMEM accumulators (2 x 256 columns).
// Requires dim = 128, T = 32, nbits in {1, 2, 4}; longer passages and every other shape are scored
// by the generic kernel (stage34_generic.cu).
#include "common.cuh"
#include "ptx.cuh"

namespace {

#ifndef TC_EPI_SETS
#define TC_EPI_SETS 2                  // epilogue warp sets (each set = 4 warps); with two issuers set s reads issuer s's accumulator(s)
#endif
#ifndef TC_META_CODES
#define TC_META_CODES 128              // centroid codes of a passage's first tokens staged in its meta slot by the scheduler (0 = the decompression teams load all codes themselves)
#endif
#ifndef TC_L2_HINTS
#define TC_L2_HINTS 2                  // 1 = query-tile bulk copies carry an L2 evict_last policy (no effect); 2 = packed codes / residuals, read once per batch, are loaded with ld.global.cs (evict first: ~1 % at C, profiles/r02_ab_two_issuer_tuning.txt); 4 = the fp16 centroid rows are loaded with an L2 evict_last policy
#endif
#ifndef TC_BACKOFF_NS
#define TC_BACKOFF_NS 0                // nanosleep between polls of the waits that are usually long and have slack (loaders: free stage, scheduler: slot / ring)
#endif
#ifndef TC_BACKOFF_EPI_NS
#define TC_BACKOFF_EPI_NS 0            // ... of the epilogue's wait for an accumulator
#endif
#ifndef TC_BACKOFF_DEC_NS
#define TC_BACKOFF_DEC_NS 20           // ... of the decompression warps' wait for the next passage
#endif
#ifndef TC_NISSUE
#define TC_NISSUE 2                    // MMA issuer threads: 2 = two warps issue alternate groups, each into its own accumulator(s) (see tc_issuer_role)
#endif
#ifndef TC_EPI_MODE
#define TC_EPI_MODE ((TC_EPI_SETS > 1) ? 0 : 1)   // accumulator read-out: 0 = one 32-column load at a time, 1 = double-buffered, 2 = all loads of up to
#endif                                              // 80 columns issued at once, ONE tcgen05.wait::ld, accumulator released before the fold
// Register budget per warpgroup (setmaxnreg; 0 = leave the launch allocation).  The pool is threads x launch registers; the
// 704-thread default build (two issuers, two epilogue sets) runs every role inside the 80 registers of the launch allocation.
#ifndef TC_REG_CTRL
#if TC_EPI_SETS == 1
#define TC_REG_CTRL 96
#define TC_REG_EPI 176
#define TC_REG_DEC 120
#else
#define TC_REG_CTRL 0
#define TC_REG_EPI 0
#define TC_REG_DEC 0
#endif
#endif
constexpr int TC_NEPI_WARPS = 4 * TC_EPI_SETS;
constexpr int TC_DEC_WARP0 = 4 + TC_NEPI_WARPS;   // first warp after the epilogue sets
constexpr int TC_DIM = 128, TC_T = 32;
#ifndef TC_NACC
#define TC_NACC 2                      // TMEM accumulators: 2 x 256 columns (one per issuer), or 4 x 128 (two per issuer; passages over 128 tokens
#endif                                 // then take up to 4 chunks: measured no faster at C and slower at B)
constexpr int TC_MAX_BROWS = (TC_NACC == 4) ? 128 : 240;   // rows (tokens) per chunk; multiple of 16, <= accumulator columns
constexpr int TC_MAX_CHUNKS = (TC_NACC == 4) ? 4 : 2; // chunks per passage (accumulator passes per group)
constexpr int TC_MAX_ASTAGES = 6;
#ifndef TC_NSLOT_LOG2
#define TC_NSLOT_LOG2 2
#endif
constexpr int TC_NSLOT = 1 << TC_NSLOT_LOG2;   // passage entries in flight (meta slots == tile barriers)
constexpr int TC_A_BYTES = 128 * TC_DIM * 2;  // 32 KB: 4 queries x 32 tokens x 128 x fp16
constexpr int TC_Q_BYTES = TC_T * TC_DIM * 2; // 8 KB per query
#ifndef TC_ABLATE
#define TC_ABLATE 0                    // measurement-only builds (results are garbage): 1 = one 8 KB query copy per group instead of
#endif                                 // four, 2 = epilogue reads 16 accumulator columns, 4 = no decompression, 8 = 2 of 8 MMA K-steps, 16 = no pair append, 32 = no L2 prefetch of packed bytes
#ifndef TC_NLOAD
#define TC_NLOAD 2                     // query-tile loader warps (2: warps 2-3; 4: two more taken from the decompression pool)
#endif
constexpr int TC_NXISS = (TC_NISSUE == 2) ? 2 : 0;               // warps set aside for the second issuer (the first of them issues, the other idles)
#ifndef TC_NDEC_WARPS_
#define TC_NDEC_WARPS_ ((TC_NISSUE == 2) ? 8 : 8 - (TC_NLOAD - 2))   // two issuers: 22 warps in all (704 threads); else 16 (512 threads)
#endif
constexpr int TC_NDEC_WARPS = TC_NDEC_WARPS_;
constexpr int TC_XISS_WARP0 = TC_DEC_WARP0 + (TC_NLOAD - 2);     // second issuer's warp (loaders 2.. sit before it)
constexpr int TC_DEC_FIRST = TC_XISS_WARP0 + TC_NXISS;           // first decompression warp
#ifndef TC_NTEAMS_
#define TC_NTEAMS_ 2
#endif
constexpr int TC_NTEAMS = TC_NTEAMS_;   // decompression teams (round-robin over passages)
constexpr int TC_THREADS = 32 * (TC_DEC_FIRST + TC_NDEC_WARPS);
constexpr int TC_LAUNCH_REGS = (65536 / TC_THREADS) & ~7;   // what __launch_bounds__(TC_THREADS, 1) lets ptxas give every thread
template <int N> __device__ __forceinline__ void tc_reg_budget() {   // a role's setmaxnreg (whole warpgroup); 0 = keep the launch allocation
  if constexpr (N > TC_LAUNCH_REGS) ptx::reg_inc<N>();
  else if constexpr (N > 0 && N < TC_LAUNCH_REGS) ptx::reg_dec<N>();
}
// tensor memory: TC_NACC accumulators of TC_D_COLS fp32 columns
constexpr int TC_NACC_ = TC_NACC;
constexpr uint32_t TC_TMEM_COLS = 512, TC_D_COLS = 512 / TC_NACC_;

// Measurement-only build (-DTC_PROF=1): every mbarrier wait is timed with clock64 and charged to (warp, wait tag); slot 0 of a
// warp holds the clocks of its whole role loop.  Read back with cb_debug_tc_prof (tools/tc_wait_profile.py).
#ifndef TC_PROF
#define TC_PROF 0
#endif
#ifndef TC_TRACE_EVENTS
#define TC_TRACE_EVENTS 0xfff          // which events of the trace are compiled in (every probe costs its role ~100 clocks)
#endif
#if TC_PROF
// TC_PROF bit 0: wait / busy accounting of every warp of every CTA (TCW, TCP_*); bit 1: event trace of CTA 0 (TCT)
__device__ unsigned long long g_tc_prof[160 * 32 * 24];
__shared__ unsigned long long s_prof[32 * 24];
constexpr int TC_TRACE_G0 = 20000, TC_TRACE_N = 4096, TC_TRACE_E0 = 4700;   // per-passage events (TCE, rows 7-11): entries TC_TRACE_E0 .. +4095
  // steady state: groups TC_TRACE_G0 .. +4095: g_tc_trace[event][group] = clock64
__device__ long long g_tc_trace[16 * TC_TRACE_N];
#if TC_PROF & 1
#define TCW(bar_, par_, tag_, ...)                                                          \
  do {                                                                                      \
    const long long t0_ = clock64();                                                        \
    ptx::mbar_wait(bar_, par_, tag_, ##__VA_ARGS__);                                        \
    if (lane == 0 || warp == 1) s_prof[warp * 24 + (tag_)] += (unsigned long long)(clock64() - t0_);   \
  } while (0)
#define TCP_BEGIN() const long long tp0_ = clock64()
#define TCP_END(tag_) do { if (lane == 0 || warp == 1) s_prof[warp * 24 + (tag_)] += (unsigned long long)(clock64() - tp0_); } while (0)
#else
#define TCW(bar_, par_, tag_, ...) ptx::mbar_wait(bar_, par_, tag_, ##__VA_ARGS__)
#define TCP_BEGIN() do {} while (0)
#define TCP_END(tag_) do {} while (0)
#endif
#if TC_PROF & 2
#define TCT(ev_, g_) do { if (((TC_TRACE_EVENTS >> (ev_)) & 1) && blockIdx.x == 0 && (lane == 0 || warp == 1) && (g_) >= TC_TRACE_G0 && (g_) < TC_TRACE_G0 + TC_TRACE_N) g_tc_trace[(ev_) * TC_TRACE_N + (g_) - TC_TRACE_G0] = clock64(); } while (0)
#define TCE(ev_, e_) do { if (blockIdx.x == 0 && lane == 0 && (e_) >= TC_TRACE_E0 && (e_) < TC_TRACE_E0 + TC_TRACE_N) g_tc_trace[(ev_) * TC_TRACE_N + (e_) - TC_TRACE_E0] = clock64(); } while (0)
#else
#define TCT(ev_, g_) do {} while (0)
#define TCE(ev_, e_) do {} while (0)
#endif
#else
#define TCW(bar_, par_, tag_, ...) ptx::mbar_wait(bar_, par_, tag_, ##__VA_ARGS__)
#define TCP_BEGIN() do {} while (0)
#define TCP_END(tag_) do {} while (0)
#define TCT(ev_, g_) do {} while (0)
#define TCE(ev_, e_) do {} while (0)
#endif

struct Meta {            // one passage entry, written by the scheduler
  int ncand;             // candidate queries of the passage (< 0: end of stream)
  int L;                 // doclen
  int nchunk;            // 1 .. TC_MAX_CHUNKS
  int n0, n1;            // rows of chunks 0 .. nchunk-2 / of the last chunk (padded to 16; n1 = 0 for one chunk)
  int pid;               // local 0-based pid
  uint32_t b_off;        // byte offset of the operand tile(s) in the ring
  uint32_t pad_;
  long long e0;          // first embedding of the passage
  uint16_t q[CB_NQ_CHUNK];
  int32_t codes[TC_META_CODES > 0 ? TC_META_CODES : 1];   // codes of tokens 0 .. TC_META_CODES-1 (see tc_scheduler_role)
};

struct Barriers {
  uint64_t b_full[TC_NSLOT], b_empty[TC_NSLOT], meta_full[TC_NSLOT], meta_empty[TC_NSLOT];
  uint64_t a_full[TC_MAX_ASTAGES], a_empty[TC_MAX_ASTAGES];
  uint64_t d_full[TC_EPI_SETS][4], d_empty[4];
  uint64_t a_full2[2][TC_MAX_ASTAGES];   // TC_NISSUE == 2: "tile landed" barriers per (issuer, stage) (see tc_loader_role)
};

struct TcParams {
  const __half* centroids_h; const float* weights; const int32_t* codes; const uint8_t* residuals;
  const int64_t* offsets; int64_t Np; int R, W, nastages, ring_bytes, long_limit;
  const uint8_t* qprep;           // query row image: [nq][8 KB]
  const uint32_t* bitmap; const int64_t* list_off; int32_t* cursors; uint64_t* pairs;
  const int32_t* pid_list; int64_t n_list;   // optional: only these passages (sparse bitmaps, e.g. the PLAID rescoring pass)
  unsigned long long* stats;      // batch counters (cb_stats_dev): [6] groups, [7] operand rows over groups, [8] operand rows over passages
  const int* q_flag;              // != 0: the batch breaks the |query token| <= 255 precondition; the generic kernel scores it
};

// max over 32 / 16 TMEM columns held in registers, folded into 4 independent chains
__device__ __forceinline__ void fold32(const uint32_t (&r)[32], float& m0, float& m1, float& m2, float& m3) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    m0 = fmaxf(m0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
    m1 = fmaxf(m1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
    m2 = fmaxf(m2, fmaxf(__uint_as_float(r[i + 4]), __uint_as_float(r[i + 5])));
    m3 = fmaxf(m3, fmaxf(__uint_as_float(r[i + 6]), __uint_as_float(r[i + 7])));
  }
}
__device__ __forceinline__ void fold16(const uint32_t (&r)[16], float& m0, float& m1, float& m2, float& m3) {
  m0 = fmaxf(m0, fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])));
  m1 = fmaxf(m1, fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3])));
  m2 = fmaxf(m2, fmaxf(__uint_as_float(r[4]), __uint_as_float(r[5])));
  m3 = fmaxf(m3, fmaxf(__uint_as_float(r[6]), __uint_as_float(r[7])));
  m0 = fmaxf(m0, fmaxf(__uint_as_float(r[8]), __uint_as_float(r[9])));
  m1 = fmaxf(m1, fmaxf(__uint_as_float(r[10]), __uint_as_float(r[11])));
  m2 = fmaxf(m2, fmaxf(__uint_as_float(r[12]), __uint_as_float(r[13])));
  m3 = fmaxf(m3, fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15])));
}

// Decompression helpers.  EIGHT lanes expand one token (a warp expands 4 tokens at a time): lane l8
// owns dims 8*l8 .. 8*l8+7 and 64+8*l8 .. 64+8*l8+7, i.e. 16-byte chunk l8 of BOTH 64-element
// K-blocks of the fp16 operand row.  With that split (a) each of the lane's two 16-byte stores is,
// across the 8 lanes of a token, one whole 128-byte swizzled row of one K-block: the quarter-warp
// that shared memory serves per wavefront covers all 32 banks (the earlier "two adjacent chunks per
// lane" split put lanes l8 and l8+4 on the same banks: 2.5 G excess store wavefronts per batch,
// ncu), and (b) each of its two 16-byte centroid loads is, across the 8 lanes, one whole 128-byte
// line (half-used lines before: twice the L1 wavefronts per token).  Its 2 x 8*NBITS packed bits ...
template <int NBITS> struct Bits16 { uint32_t lo, hi; };   // NBITS bytes each: dims 8*l8.. and 64+8*l8..
template <int NBITS>
__device__ __forceinline__ Bits16<NBITS> load_bits16(const uint8_t* __restrict__ emb, int l8) {
  Bits16<NBITS> b;
  if constexpr (TC_L2_HINTS & 2) {     // read once per batch: evict first
    if constexpr (NBITS == 1) { b.lo = __ldcs(emb + l8); b.hi = __ldcs(emb + 8 + l8); }
    else if constexpr (NBITS == 2) { b.lo = __ldcs(reinterpret_cast<const uint16_t*>(emb) + l8); b.hi = __ldcs(reinterpret_cast<const uint16_t*>(emb) + 8 + l8); }
    else { b.lo = __ldcs(reinterpret_cast<const uint32_t*>(emb) + l8); b.hi = __ldcs(reinterpret_cast<const uint32_t*>(emb) + 8 + l8); }
    return b;
  }
  if constexpr (NBITS == 1) { b.lo = emb[l8]; b.hi = emb[8 + l8]; }
  else if constexpr (NBITS == 2) { b.lo = reinterpret_cast<const uint16_t*>(emb)[l8]; b.hi = reinterpret_cast<const uint16_t*>(emb)[8 + l8]; }
  else { b.lo = reinterpret_cast<const uint32_t*>(emb)[l8]; b.hi = reinterpret_cast<const uint32_t*>(emb)[8 + l8]; }
  return b;
}
// ... are expanded through a shared-memory table indexed by packed BYTE: entry = the 8/NBITS bucket
// weights (fp16) of the dims packed in that byte (`_unpackbits`/`_unbinarize` + `bucket_weights[idx]`,
// src/indexing/codecs/residual.jl:709-719: dim d's index sits at bit d*NBITS, LSB first).  Every
// entry is replicated across the 128-byte bank row (entry e of byte b for replica r at
// b*128 + r*entry_bytes), and a lane reads replica lane % replicas, so the lanes that share one
// shared-memory wavefront always hit different banks whatever bytes they hold: no bank conflicts.
#ifndef TC_LUT_ROW
#define TC_LUT_ROW 128                  // bytes per table row (entry replicated TC_LUT_ROW / entry size times): 128 = no bank conflicts at all
#endif
template <int NBITS> struct LutGeom {
  static constexpr int ENTRY_BYTES = 16 / NBITS;          // 8/NBITS fp16 weights
  static constexpr int REPLICAS = (TC_LUT_ROW / ENTRY_BYTES) > 0 ? (TC_LUT_ROW / ENTRY_BYTES) : 1;      // 8, 16, 32 at 128-byte rows
  static constexpr int ROW = REPLICAS * ENTRY_BYTES;
};
constexpr int TC_LUT_BYTES = 256 * TC_LUT_ROW;
template <int NBITS>
__device__ __forceinline__ void lookup_weights16(const uint8_t* __restrict__ lut_lane, const Bits16<NBITS>& b, __half2 (&w)[8]) {
  // lut_lane = table + (lane % replicas) * entry_bytes; w[0..3] = dims 8*l8.., w[4..7] = dims 64+8*l8..
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t bits = h ? b.hi : b.lo;
    if constexpr (NBITS == 1) {
      const uint4 x = *reinterpret_cast<const uint4*>(lut_lane + (bits & 255u) * (uint32_t)LutGeom<NBITS>::ROW);
      w[4 * h] = *reinterpret_cast<const __half2*>(&x.x); w[4 * h + 1] = *reinterpret_cast<const __half2*>(&x.y);
      w[4 * h + 2] = *reinterpret_cast<const __half2*>(&x.z); w[4 * h + 3] = *reinterpret_cast<const __half2*>(&x.w);
    } else if constexpr (NBITS == 2) {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const uint2 x = *reinterpret_cast<const uint2*>(lut_lane + ((bits >> (8 * j)) & 255u) * (uint32_t)LutGeom<NBITS>::ROW);
        w[4 * h + 2 * j] = *reinterpret_cast<const __half2*>(&x.x); w[4 * h + 2 * j + 1] = *reinterpret_cast<const __half2*>(&x.y);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t x = *reinterpret_cast<const uint32_t*>(lut_lane + ((bits >> (8 * j)) & 255u) * (uint32_t)LutGeom<NBITS>::ROW);
        w[4 * h + j] = *reinterpret_cast<const __half2*>(&x);
      }
    }
  }
}

// v = centroid + w[bucket]; v /= (|v| + eps) (`decompress` + `_normalize_array!`, residual.jl:776-781,
// utils.jl:320-325); this lane's two 16-byte chunks of operand row `row`.  The operand is fp16 and
// the centroid image already is, so the sum is formed in packed fp16 (one rounding of 2^-12 relative,
// random); the squared norm is accumulated over 4-term fp16 chains, summed in fp32 across the lane
// and the 8-lane group; the scale is applied as v*hi + v*lo with (hi, lo) the fp16 split of the
// fp32 1/(|v|+eps), so no systematic per-token scale error survives.  Measured score error against
// the fp32 oracle stays below 2e-4 relative (tolerance 1e-3, tests/test_gpu_parity.py).
// DUMP (parity hook only, cb_debug_tc_operand): additionally writes the un-normalised fp16 sums of this lane
// to raw_row[dims] -- fp16(centroid) + fp16(w[bucket]) is exactly reproducible on the host, which pins
// every unpacked bucket index of THIS code path bit for bit.
template <int NBITS, bool DUMP = false>
__device__ __forceinline__ void finish_token16(const uint8_t* __restrict__ lut_lane, const Bits16<NBITS>& bits, const uint4 (&craw)[2],
                                               int l8, uint8_t* tile, int kb_stride, int row, __half* raw_row = nullptr) {
  __half2 v[8];
  lookup_weights16<NBITS>(lut_lane, bits, v);
  const __half2* c0 = reinterpret_cast<const __half2*>(&craw[0]);
  const __half2* c1 = reinterpret_cast<const __half2*>(&craw[1]);
#pragma unroll
  for (int j = 0; j < 4; j++) { v[j] = __hadd2(c0[j], v[j]); v[4 + j] = __hadd2(c1[j], v[4 + j]); }
  if constexpr (DUMP) {
    if (raw_row != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        reinterpret_cast<__half2*>(raw_row + 8 * l8)[j] = v[j];
        reinterpret_cast<__half2*>(raw_row + 64 + 8 * l8)[j] = v[4 + j];
      }
    }
  }
  __half2 s0 = __hmul2(v[0], v[0]), s1 = __hmul2(v[1], v[1]), s2 = __hmul2(v[2], v[2]), s3 = __hmul2(v[3], v[3]);
  s0 = __hfma2(v[4], v[4], s0); s1 = __hfma2(v[5], v[5], s1); s2 = __hfma2(v[6], v[6], s2); s3 = __hfma2(v[7], v[7], s3);
  const float2 f0 = __half22float2(s0), f1 = __half22float2(s1), f2 = __half22float2(s2), f3 = __half22float2(s3);
  float ss = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);   // stays inside the 8-lane group
  const float inv = 1.0f / (sqrtf(ss) + 1.1920929e-07f);   // X ./ (norm + eps)
  const __half hi = __float2half_rn(inv);
  const __half lo = __float2half_rn(inv - __half2float(hi));
  const __half2 hi2 = __half2half2(hi), lo2 = __half2half2(lo);
  uint8_t* base = tile + (row >> 3) * 1024 + (row & 7) * 128 + ((l8 ^ (row & 7)) << 4);   // SWIZZLE_128B: chunk c of row r sits at c ^ (r & 7)
#pragma unroll
  for (int h = 0; h < 2; h++) {     // h = K-block
    uint4 o;
    __half2 t;
    t = __hfma2(v[4 * h + 0], lo2, __hmul2(v[4 * h + 0], hi2)); o.x = *reinterpret_cast<const uint32_t*>(&t);
    t = __hfma2(v[4 * h + 1], lo2, __hmul2(v[4 * h + 1], hi2)); o.y = *reinterpret_cast<const uint32_t*>(&t);
    t = __hfma2(v[4 * h + 2], lo2, __hmul2(v[4 * h + 2], hi2)); o.z = *reinterpret_cast<const uint32_t*>(&t);
    t = __hfma2(v[4 * h + 3], lo2, __hmul2(v[4 * h + 3], hi2)); o.w = *reinterpret_cast<const uint32_t*>(&t);
    *reinterpret_cast<uint4*>(base + h * kb_stride) = o;
  }
}

// Operand-tile geometry of a passage of L tokens: one chunk of n0 rows (L padded to 16), or nchunk <= TC_MAX_CHUNKS balanced
// chunks when L > TC_MAX_BROWS -- chunks 0 .. nchunk-2 have n0 rows, the last one n1 (one accumulator pass each, one running
// maximum).  n1 = 0 for a single chunk.
__device__ __forceinline__ void tc_tile_geometry(int L, int& nchunk, int& n0, int& n1) {
  nchunk = 1; n0 = (L + 15) & ~15; n1 = 0;
  if (L > TC_MAX_BROWS) {
    nchunk = (L + TC_MAX_BROWS - 1) / TC_MAX_BROWS;
    n0 = (((L + nchunk - 1) / nchunk) + 15) & ~15;
    n1 = (L - (nchunk - 1) * n0 + 15) & ~15;
  }
}
__device__ __forceinline__ int tc_total_rows(int nchunk, int n0, int n1) { return nchunk == 1 ? n0 : (nchunk - 1) * n0 + n1; }

// byte -> bucket weights table (fp16, every entry replicated across its 128-byte bank row)
template <int NBITS>
__device__ __forceinline__ void tc_fill_lut(uint8_t* s_lut, const float* __restrict__ weights, int tid, int nthreads) {
  constexpr int DPB = 8 / NBITS, REP = LutGeom<NBITS>::REPLICAS;
  for (int i = tid; i < 256 * REP * DPB; i += nthreads) {
    const int j = i % DPB, r = (i / DPB) % REP, byte = i / (DPB * REP);
    reinterpret_cast<__half*>(s_lut + byte * LutGeom<NBITS>::ROW + r * LutGeom<NBITS>::ENTRY_BYTES)[j] =
        __float2half_rn(weights[(byte >> (j * NBITS)) & ((1 << NBITS) - 1)]);
  }
}

#ifndef TC_DBATCH_
#define TC_DBATCH_ ((TC_NISSUE == 2) ? 4 : 5)   // (the 704-thread build has 80 registers per thread: 5 spills)
#endif
constexpr int TC_DBATCH = TC_DBATCH_;   // decompression rounds whose loads are all in flight at once
constexpr int TC_TEAM_WARPS = TC_NDEC_WARPS / TC_NTEAMS;

// One team warp's share of a passage: packed codes/residuals -> normalised fp16 operand tile(s) in shared
// memory (rows past the last token re-expand the last token).  Used by the scoring kernel's
// decompression role and, unchanged, by the parity hook kernel k_tc_dump (DUMP = true).
template <int NBITS, bool DUMP>
__device__ __forceinline__ void tc_decompress_passage(const TcParams& P, const uint8_t* __restrict__ lut_lane, int dw, int lane,
                                                      int L, int nchunk, int n0, int n1, int64_t e0, uint8_t* tile0, __half* raw_out,
                                                      const int32_t* s_codes = nullptr) {
  constexpr int TEAM_WARPS = TC_TEAM_WARPS;
  [[maybe_unused]] const uint64_t l2_keep = (TC_L2_HINTS & 4) ? ptx::l2_policy_evict_last() : 0ull;   // centroid rows: re-read ~2300 times per batch
  // code of token t: from the meta slot (staged by the scheduler) when it is there, else from global memory
  auto code_of = [&](int t) -> int32_t {
    if (TC_META_CODES > 0 && s_codes != nullptr && t < TC_META_CODES) return s_codes[t];
    return (TC_L2_HINTS & 2) ? __ldcs(P.codes + e0 + t) : P.codes[e0 + t];
  };
  const int l8 = lane & 7;
  const int nrows = tc_total_rows(nchunk, n0, n1);              // operand rows (multiple of 16)
  // operand row of this lane in round j: rr = 4 * (dw + TEAM_WARPS * j) + (lane >> 3)
  const int rr0 = 4 * dw + (lane >> 3);
  const int nround = (nrows - 4 * dw + 4 * TEAM_WARPS - 1) / (4 * TEAM_WARPS);   // warp-uniform
  const int nround_max = (nrows + 4 * TEAM_WARPS - 1) / (4 * TEAM_WARPS);
  const int nbatch = (nround_max + TC_DBATCH - 1) / TC_DBATCH;
  const int per = (nround_max + nbatch - 1) / nbatch;           // balanced batch length (<= TC_DBATCH)
  int32_t code_next[TC_DBATCH];
#pragma unroll
  for (int i = 0; i < TC_DBATCH; i++) {
    const int t = min(rr0 + 4 * TEAM_WARPS * i, L - 1);
    code_next[i] = (i < per && i < nround) ? code_of(t) : 0;
  }
  for (int j0 = 0; j0 < nround; j0 += per) {
    Bits16<NBITS> bits[TC_DBATCH];
    uint4 cr[TC_DBATCH][2];
#pragma unroll
    for (int i = 0; i < TC_DBATCH; i++) {
      if (i < per && j0 + i < nround) {
        const int t = min(rr0 + 4 * TEAM_WARPS * (j0 + i), L - 1);
        bits[i] = load_bits16<NBITS>(P.residuals + (e0 + t) * P.R, l8);
        const uint4* crow = reinterpret_cast<const uint4*>(P.centroids_h + (int64_t)code_next[i] * TC_DIM) + l8;
        if (TC_L2_HINTS & 4) { cr[i][0] = ptx::ld_global_v4_hint(crow, l2_keep); cr[i][1] = ptx::ld_global_v4_hint(crow + 8, l2_keep); }
        else { cr[i][0] = crow[0]; cr[i][1] = crow[8]; }
      }
    }
#pragma unroll
    for (int i = 0; i < TC_DBATCH; i++) {
      const int t = min(rr0 + 4 * TEAM_WARPS * (j0 + per + i), L - 1);
      code_next[i] = (i < per && j0 + per + i < nround) ? code_of(t) : 0;
    }
#pragma unroll
    for (int i = 0; i < TC_DBATCH; i++) {
      if (i < per && j0 + i < nround) {   // warp-uniform
        const int rr = rr0 + 4 * TEAM_WARPS * (j0 + i);
        const int c = (rr >= n0 ? 1 : 0) + (rr >= 2 * n0 ? 1 : 0) + (rr >= 3 * n0 ? 1 : 0);   // chunk of this row (nchunk <= 4)
        finish_token16<NBITS, DUMP>(lut_lane, bits[i], cr[i], l8, tile0 + c * n0 * 256, ((nchunk > 1 && c == nchunk - 1) ? n1 : n0) * 128, rr - c * n0,
                                    (DUMP && rr < L) ? raw_out + (size_t)rr * TC_DIM : nullptr);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Roles of the scoring kernel
// ---------------------------------------------------------------------------------------------
struct TcCtx {           // the CTA's shared-memory carve-up
  uint8_t* ring; uint8_t* a_tile0; Meta* meta; Barriers* bar; uint32_t* s_region; uint8_t* s_lut; int NA;
};

// ===== query-tile loaders: EVERY loader serves EVERY group, loader li fetching the group's queries
// j = li, li + NLOAD, ...  One warp sustains only ~31 B/clk of 8 KB bulk copies (a copy occupies
// its issuing warp for ~270 clocks: tools/l2_to_sm_ceiling.cu, profiles/r02_l2_to_sm_ceiling.txt),
// while the L2 -> SM path itself carries 76 B/clk/SM; spreading one group's four copies over
// several issuers shortens the fill time of a stage. =====
template <int NLOAD>
__device__ __forceinline__ void tc_loader_role(const TcParams& P, const TcCtx& S, const int li, const int warp, const int lane) {
  Barriers* const bar = S.bar; Meta* const meta = S.meta; uint8_t* const a_tile0 = S.a_tile0; const int NA = S.NA;
  (void)warp;
  uint32_t st = 0, a_par = 1;   // stage of the current group / parity of its next a_empty phase
  [[maybe_unused]] const uint64_t l2pol = (TC_L2_HINTS & 1) ? ptx::l2_policy_evict_last() : 0ull;
  [[maybe_unused]] uint32_t turn = 0;   // TC_NISSUE == 2: issuer of the current group
  [[maybe_unused]] int gc = 0;
  for (int e = 0;; e++) {
    const int slot = e & (TC_NSLOT - 1);
    TCW(&bar->meta_full[slot], (e >> TC_NSLOT_LOG2) & 1, 8);
    const Meta& m = meta[slot];
    const int ncand = m.ncand;
    if (ncand < 0) break;
    const int ngroups = (ncand + 3) >> 2;
    for (int g = 0; g < ngroups; g++) {
      const uint32_t st_g = st, par_g = a_par;
      if (++st == (uint32_t)NA) { st = 0; a_par ^= 1u; }
      TCW(&bar->a_empty[st_g], par_g, 9, TC_BACKOFF_NS);
      if (li == 0) TCT(6, gc);
      gc++;
      // With two issuers the "tile landed" barrier of a stage is the one of the group's ISSUER: a waiter must observe every phase
      // of a barrier it waits on, in order (an issuer asking for its group G while the stage's previous tile, the other issuer's
      // group G - NA, has not landed yet would take the phase of G - 2 NA -- same parity -- for its own).
#if TC_NISSUE == 2
      uint64_t* const full = &bar->a_full2[turn][st_g];
      turn ^= 1u;
#else
      uint64_t* const full = &bar->a_full[st_g];
#endif
      const int nqg = (TC_ABLATE & 1) ? 1 : min(4, ncand - g * 4);
      uint8_t* dst = a_tile0 + (size_t)st_g * TC_A_BYTES;
      const int nmine = (nqg - li + NLOAD - 1) / NLOAD;      // queries li, li + NLOAD, ... < nqg
      if (nmine <= 0) {
        if (ptx::elect_one()) ptx::mbar_arrive(full);
        continue;
      }
      const int qv = (lane < nmine) ? (int)m.q[g * 4 + li + lane * NLOAD] : 0;   // lane i holds this loader's i-th query
      if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(full, (uint32_t)nmine * TC_Q_BYTES);
      for (int i = 0; i < nmine; i++) {
        const int q = __shfl_sync(0xffffffffu, qv, i);
        if (ptx::elect_one()) {
          if (TC_L2_HINTS & 1) ptx::bulk_g2s_hint(dst + (li + i * NLOAD) * TC_Q_BYTES, P.qprep + (size_t)q * TC_Q_BYTES, TC_Q_BYTES, full, l2pol);
          else ptx::bulk_g2s(dst + (li + i * NLOAD) * TC_Q_BYTES, P.qprep + (size_t)q * TC_Q_BYTES, TC_Q_BYTES, full);
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&bar->meta_empty[slot]);
  }
}

// ===== MMA issue by TC_NISSUE = 2 threads (shared-memory A operand).  Why: tcgen05.mma issue is asynchronous only up to a queue
// of ~6 instructions, and ONE thread needs more time per 4-query group than the tensor pipe (8 MMAs = ~420 clocks at N = 80): two
// mbarrier probes of ~100 clocks each (try_wait + dependent branch, even when the phase completed long ago), ~50 uniform-datapath
// instructions of descriptor arithmetic, three commits.  tools/pipe_skeleton.cu isolates it: the same loop with loop-invariant
// descriptors and no waits runs at 416 clocks per group, with per-group descriptors and the two waits at 680-730, and with two
// issuers on alternate groups at ~510-550 again.  Issuer `me` takes the groups G = me (mod 2) of the CTA's running group count
// into its own accumulator(s) me, me + 2, ...; both walk every passage and keep the stage / group counters of all groups.  A
// passage's tile is released by BOTH (b_empty counts 2: a commit covers "all my earlier MMAs", also when none read this tile),
// and both are among the arrivals that release its meta slot (an issuer without a group in the passage could otherwise still
// be about to read a slot the scheduler has republished). =====
__device__ __forceinline__ void tc_issuer_role(const TcCtx& S, const uint32_t tmem_base, const int me, const int warp, const int lane) {
  Barriers* const bar = S.bar; Meta* const meta = S.meta; const int NA = S.NA;
  (void)warp; (void)lane;
  constexpr int NI = 2, ACC_PER = TC_NACC_ / NI;
  const uint32_t a_lo0 = ((ptx::smem_u32(S.a_tile0) & 0x3ffffu) >> 4) | (1u << 16);   // descriptor low words: start address >> 4, LBO field = 1
  const uint32_t ring_lo = ((ptx::smem_u32(S.ring) & 0x3ffffu) >> 4) | (1u << 16);
  constexpr uint32_t HI_A = (2048u >> 4) | (1u << 14) | (2u << 29);   // SBO 2048 | version 1 | SWIZZLE_128B
  constexpr uint32_t HI_B = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024
  const uint32_t period = (NA & 1) ? 2u * (uint32_t)NA : (uint32_t)NA;   // barrier (issuer, stage) serves every lcm(2, NA)-th group
  uint32_t st = 0, pq = 0, a_par = 0;   // query-tile stage of the running group G / G % period / parity of phase G / period of its a_full barrier
  uint32_t ka = 0, d_par = 1;        // own accumulator me + NI * ka / parity of the next d_empty phase of the own accumulators
  uint32_t turn = 0;                 // running group count mod NI
  [[maybe_unused]] int gc = 0;
  for (int e = 0;; e++) {
    const int slot = e & (TC_NSLOT - 1);
    const uint32_t ph = (e >> TC_NSLOT_LOG2) & 1;
    TCW(&bar->meta_full[slot], ph, 4);
    const int ncand = meta[slot].ncand;
    if (ncand < 0) break;
    const int nchunk = meta[slot].nchunk, n0 = meta[slot].n0, n1 = meta[slot].n1;
    const uint32_t b_lo0 = ring_lo + (meta[slot].b_off >> 4);
    const uint32_t kb0 = (uint32_t)n0 * 8u, kb1 = (uint32_t)n1 * 8u;   // K-block stride (rows * 128 B) >> 4
    const uint32_t idesc0 = ptx::idesc_f16(128, n0, 0), idesc1 = ptx::idesc_f16(128, n1 > 0 ? n1 : 16, 0);
    const int ngroups = (ncand + 3) >> 2;
    bool waited = false;
    for (int g = 0; g < ngroups; g++) {
      const bool mine = turn == (uint32_t)me;
      turn = (turn + 1u) & (uint32_t)(NI - 1);
      const uint32_t st_g = st, par_g = a_par;
      if (++st == (uint32_t)NA) st = 0;
      if (++pq == period) { pq = 0; a_par ^= 1u; }
      if (!mine) { gc++; continue; }
      if (!waited) { TCW(&bar->b_full[slot], ph, 5); waited = true; }
      TCW(&bar->a_full2[me][st_g], par_g, 6);
      TCT(0, gc);
      const uint32_t a_lo = a_lo0 + st_g * (uint32_t)(TC_A_BYTES >> 4);
      for (int c = 0; c < nchunk; c++) {
        const uint32_t ds = (uint32_t)me + (uint32_t)NI * ka;
        TCW(&bar->d_empty[ds], d_par, 7);
        ptx::tc_fence_after();
        TCT(1, gc);
        const uint32_t d_tmem = tmem_base + ds * TC_D_COLS;
        const bool lastc = c > 0 && c == nchunk - 1;                   // chunk c starts c * n0 * 256 bytes in; the last of several has n1 rows
        const uint32_t b_lo = b_lo0 + (uint32_t)c * (uint32_t)n0 * 16u, kb = lastc ? kb1 : kb0, idesc = lastc ? idesc1 : idesc0;
#pragma unroll
        for (int k = 0; k < ((TC_ABLATE & 8) ? 2 : 8); k++) {
          const uint64_t db = ((uint64_t)HI_B << 32) | (uint64_t)(b_lo + (uint32_t)(k >> 2) * kb + (uint32_t)(k & 3) * 2);
          const uint64_t da = ((uint64_t)HI_A << 32) | (uint64_t)(a_lo + (uint32_t)((k >> 2) * 64 + (k & 3) * 2));
          ptx::mma_f16_ss(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
        }
        if (c == nchunk - 1) ptx::tc_commit(&bar->a_empty[st_g]);
        ptx::tc_commit(&bar->d_full[TC_EPI_SETS == 2 ? me : 0][ds]);
        TCT(2, gc);
        if (++ka == (uint32_t)ACC_PER) { ka = 0; d_par ^= 1u; }
      }
      gc++;
    }
    ptx::tc_commit(&bar->b_empty[slot]);   // arrives after this thread's earlier MMAs have retired (none of them may have read this tile)
    ptx::mbar_arrive(&bar->meta_empty[slot]);
  }
}

// ===== scheduler: candidate list, ring allocation and meta of every passage with >= 1 candidate =====
__device__ __forceinline__ void tc_scheduler_role(const TcParams& P, const TcCtx& S, const int warp, const int lane) {
  Barriers* const bar = S.bar; Meta* const meta = S.meta; uint32_t* const s_region = S.s_region;
  (void)warp;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  uint32_t head = 0;             // next free byte of the ring
  int e = 0;                     // entry counter
  unsigned long long n_groups = 0, n_group_rows = 0, n_passage_rows = 0;   // what this CTA pushes through the tensor pipe (batch counters)
  int tail = 0;                  // entries < tail are known to have released their tile
  // Everything the scheduler reads from global memory for passage p -- its extent, its bitmap words
  // and the extent of the passage whose packed bytes it will prefetch -- is requested TWO
  // iterations ahead and carried in registers, so no global latency sits on the per-passage path
  // (one exposed L2/HBM round trip per passage made the scheduler the bottleneck of the kernel at
  // ~17 candidates per passage: nothing downstream ever saw a full pipeline).
  struct Hdr { int64_t o0, o1, f0, f1, p; uint32_t w; };
  // With a passage list (sparse bitmap) item i is passage pid_list[i]: the pid is requested one more
  // iteration ahead than the header that depends on it.
  const int64_t n_items = P.pid_list ? P.n_list : P.Np;
  auto pid_of = [&](int64_t i) -> int64_t { return (P.pid_list && i < n_items) ? (int64_t)P.pid_list[i] : i; };
  auto load_hdr = [&](int64_t i, int64_t p) {
    Hdr h; h.o0 = h.o1 = h.f0 = h.f1 = 0; h.w = 0u; h.p = p;
    if (i < n_items) {
      h.o0 = P.offsets[p]; h.o1 = P.offsets[p + 1];
      h.w = (lane < P.W) ? P.bitmap[p * P.W + lane] : 0u;
      const int64_t pf = p + 6 * stride;       // pull its packed bytes from HBM into L2 a few passages ahead
      if (!P.pid_list && pf < P.Np) { h.f0 = P.offsets[pf]; h.f1 = P.offsets[pf + 1]; }
    }
    return h;
  };
  Hdr h1 = load_hdr(first, pid_of(first)), h2 = load_hdr(first + stride, pid_of(first + stride));
  int64_t pq = pid_of(first + 2 * stride);
  // The centroid codes of a passage's first TC_META_CODES tokens ride along in its meta slot: code -> centroid row is a dependent
  // pair of global round trips for the decompression teams (~1.5k clocks each under load, two to three times per passage);
  // the scheduler requests the codes of the NEXT item one iteration ahead, so the teams start at the second round trip.
  constexpr int NCR = (TC_META_CODES + 31) / 32;
  [[maybe_unused]] int32_t c_next[NCR > 0 ? NCR : 1];
  auto load_codes = [&](const Hdr& hh, int32_t (&c)[NCR > 0 ? NCR : 1]) {
#pragma unroll
    for (int j = 0; j < NCR; j++) {
      const int64_t t = lane + 32 * j;
      c[j] = (t < hh.o1 - hh.o0) ? P.codes[hh.o0 + t] : 0;
    }
  };
  if (TC_META_CODES > 0) load_codes(h1, c_next);
  for (int64_t it = first; it < n_items; it += stride) {
    const Hdr h = h1;
    h1 = h2;
    [[maybe_unused]] int32_t c_cur[NCR > 0 ? NCR : 1];
    if (TC_META_CODES > 0) {
#pragma unroll
      for (int j = 0; j < NCR; j++) c_cur[j] = c_next[j];
      load_codes(h1, c_next);
    }
    const int64_t pq_next = pid_of(it + 3 * stride);
    h2 = load_hdr(it + 2 * stride, pq);
    pq = pq_next;
    const int64_t p = h.p;
    const int64_t e0 = h.o0;
    const int L = (int)(h.o1 - h.o0);
    uint32_t w = h.w;
    // packed bytes to pull from HBM into L2 ahead of the decompression teams: the passage 6 steps ahead (whole index), or -- with a
    // passage list, where a passage is only known one header ahead -- the NEXT item (its header was requested an iteration ago)
    const int64_t pf0 = P.pid_list ? h1.o0 : h.f0, pf1 = P.pid_list ? h1.o1 : h.f1;
    if (pf1 > pf0 && !(TC_ABLATE & 32)) {
      const char* r0 = reinterpret_cast<const char*>(P.residuals) + pf0 * P.R;
      const char* c0 = reinterpret_cast<const char*>(P.codes) + pf0 * 4;
      const int64_t rbytes = (pf1 - pf0) * P.R, cbytes = (pf1 - pf0) * 4;
      for (int64_t o = (int64_t)lane * 128; o < rbytes; o += 32 * 128) ptx::prefetch_l2(r0 + o);
      for (int64_t o = (int64_t)lane * 128; o < cbytes; o += 32 * 128) ptx::prefetch_l2(c0 + o);
    }
    if (L <= 0 || L > P.long_limit) w = 0u;          // empty, or too long: the generic kernel scores it
    if (!__any_sync(0xffffffffu, w != 0u)) continue;  // no query of the batch wants this passage
    // tile geometry
    int nchunk, n0, n1;
    tc_tile_geometry(L, nchunk, n0, n1);
    const uint32_t bytes = (uint32_t)tc_total_rows(nchunk, n0, n1) * 256u;

    // meta slot: wait until entry e-4 has been fully consumed (that also frees its tile)
    const int slot = e & (TC_NSLOT - 1);
    TCW(&bar->meta_empty[slot], ((e >> TC_NSLOT_LOG2) & 1) ^ 1, 1, TC_BACKOFF_NS);
    TCE(7, e);
    if (tail < e - (TC_NSLOT - 1)) tail = e - (TC_NSLOT - 1);
    // ring region: first fit at head, else wrap to 0; wait for the live entries it overlaps.  The
    // regions of the last TC_NSLOT entries sit in shared memory (dynamically indexed local arrays
    // would live in local memory, and with the whole carve-out given to shared memory there is no
    // L1 to hold them)
    uint32_t off = head;
    if (off + bytes > (uint32_t)P.ring_bytes) off = 0;
    int need = tail;             // entries < need must be free
#pragma unroll
    for (int j = 1; j < TC_NSLOT; j++) {
      const int ej = e - j;
      if (ej >= tail) {
        const int sj = ej & (TC_NSLOT - 1);
        if (off < s_region[2 * sj + 1] && s_region[2 * sj] < off + bytes && need < ej + 1) need = ej + 1;
      }
    }
    for (; tail < need; tail++) TCW(&bar->b_empty[tail & (TC_NSLOT - 1)], (tail >> TC_NSLOT_LOG2) & 1, 2, TC_BACKOFF_NS);
    __syncwarp();
    if (lane == 0) { s_region[2 * slot] = off; s_region[2 * slot + 1] = off + bytes; }
    __syncwarp();
    head = off + bytes;
    // candidate list
    Meta& m = meta[slot];
    const int c = __popc(w);
    int pre = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += v; }
    int base = pre - c;
    while (w) { const int b = __ffs(w) - 1; w &= w - 1; m.q[base++] = (uint16_t)(lane * 32 + b); }
    const int ncand = __shfl_sync(0xffffffffu, pre, 31);
    if (TC_META_CODES > 0) {
#pragma unroll
      for (int j = 0; j < NCR; j++) m.codes[lane + 32 * j] = c_cur[j];
    }
    if (lane == 0) {
      m.ncand = ncand; m.L = L; m.nchunk = nchunk; m.n0 = n0; m.n1 = n1; m.pid = (int)p; m.b_off = off; m.e0 = e0;
    }
    n_groups += (unsigned long long)((ncand + 3) >> 2);
    n_group_rows += (unsigned long long)((ncand + 3) >> 2) * (bytes >> 8);
    n_passage_rows += bytes >> 8;
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&bar->meta_full[slot]);
    TCE(8, e);
    e++;
  }
  // end of stream
  const int slot = e & (TC_NSLOT - 1);
  TCW(&bar->meta_empty[slot], ((e >> TC_NSLOT_LOG2) & 1) ^ 1, 3);
  if (lane == 0) { meta[slot].ncand = -1; ptx::mbar_arrive(&bar->meta_full[slot]); }
  if (lane == 0 && P.stats != nullptr) {
    atomicAdd(P.stats + 6, n_groups); atomicAdd(P.stats + 7, n_group_rows); atomicAdd(P.stats + 8, n_passage_rows);
  }
}

// ===== decompression: packed codes/residuals -> normalised fp16 operand tile(s) =====
// Eight lanes per token, four tokens per warp-round (fewer, wider instructions per token than a
// finer split).  What matters besides instruction count is memory-level parallelism (code ->
// centroid row is a dependent pair of loads, ~1-2k clocks under load), so (a) the eight warps
// form TC_NTEAMS teams that expand alternate passages concurrently, and (b) a team expands a
// passage in balanced batches of up to TC_DBATCH rounds with every load of a batch issued
// before any of it is consumed and the codes of the next batch already requested.  Operand
// rows past the last token (padding to 16) re-expand the last token: no column masking later.
template <int NBITS>
__device__ __forceinline__ void tc_decompress_role(const TcParams& P, const TcCtx& S, const int team, const int dw, const int warp, const int lane) {
  Barriers* const bar = S.bar; Meta* const meta = S.meta; uint8_t* const ring = S.ring; uint8_t* const s_lut = S.s_lut;
  (void)warp;
  const uint8_t* lut_lane = s_lut + (lane & (LutGeom<NBITS>::REPLICAS - 1)) * LutGeom<NBITS>::ENTRY_BYTES;
  for (int e = team;; e += TC_NTEAMS) {
    // An entry of another team between this team's previous entry and e may end the stream.  Every
    // decompression warp looks at EVERY entry exactly once and is one of the arrivals that release
    // its meta slot, so the scheduler cannot republish a slot (and flip the parity this probe waits
    // on) before the probe has happened.
    bool stop = false;
    for (int ee = (e >= TC_NTEAMS ? e - TC_NTEAMS + 1 : 0); ee <= e; ee++) {
      const int sl = ee & (TC_NSLOT - 1);
      TCW(&bar->meta_full[sl], (ee >> TC_NSLOT_LOG2) & 1, 12, TC_BACKOFF_DEC_NS);
      if (meta[sl].ncand < 0) { stop = true; break; }
      if (ee != e) { __syncwarp(); if (lane == 0) ptx::mbar_arrive(&bar->meta_empty[sl]); }
    }
    if (stop) break;
    const int slot = e & (TC_NSLOT - 1);
    const Meta& m = meta[slot];
    if (dw == 0) TCE(9, e);
    { TCP_BEGIN();
    if (!(TC_ABLATE & 4)) tc_decompress_passage<NBITS, false>(P, lut_lane, dw, lane, m.L, m.nchunk, m.n0, m.n1, m.e0, ring + m.b_off, nullptr, m.codes);
    TCP_END(20); }
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) { ptx::mbar_arrive(&bar->b_full[slot]); ptx::mbar_arrive(&bar->meta_empty[slot]); }
    if (dw == 0) TCE(10, e);
  }
}

template <int NBITS>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_maxsim_tc(TcParams P) {
  if (*P.q_flag != 0) return;   // whole grid, before any barrier / TMEM allocation
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment: align the dynamic window by hand
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NA = P.nastages;
  uint8_t* ring = smem;                                   // P.ring_bytes (multiple of 1024)
  uint8_t* a_tile0 = smem + P.ring_bytes;                 // NA stages of 32 KB
  Meta* meta = reinterpret_cast<Meta*>(a_tile0 + (size_t)NA * TC_A_BYTES);   // [TC_NSLOT]
  Barriers* bar = reinterpret_cast<Barriers*>(meta + TC_NSLOT);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
  uint32_t* s_region = s_tmem + 4;                        // [TC_NSLOT][2]: ring region (begin, end) of the entry in each slot
  uint8_t* s_lut = reinterpret_cast<uint8_t*>(s_region + 2 * TC_NSLOT) + 64;   // [256][128 B] (aligned below)
  s_lut += (128u - (ptx::smem_u32(s_lut) & 127u)) & 127u;

  // the warp index goes through a shuffle so that ptxas knows it is warp-uniform: every role branch
  // and everything derived from `warp` then stays on the uniform datapath (no WARPSYNC before the
  // TMEM loads, no divergence checks before the reductions: worth ~10 % of kernel time, measured)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
#if TC_PROF
  for (int i = tid; i < 32 * 24; i += TC_THREADS) s_prof[i] = 0;
  const long long prof_t0 = clock64();
#endif

  if (tid == 0) {
    for (int i = 0; i < TC_NSLOT; i++) {
      ptx::mbar_init(&bar->b_full[i], TC_NDEC_WARPS / TC_NTEAMS); ptx::mbar_init(&bar->b_empty[i], TC_NISSUE);
      ptx::mbar_init(&bar->meta_full[i], 1);          ptx::mbar_init(&bar->meta_empty[i], TC_NLOAD + TC_NEPI_WARPS + TC_NDEC_WARPS + (TC_NISSUE == 2 ? 2 : 0));
    }
    for (int i = 0; i < TC_NACC_; i++) {
      for (int s = 0; s < TC_EPI_SETS; s++) ptx::mbar_init(&bar->d_full[s][i], 1);
      ptx::mbar_init(&bar->d_empty[i], 4);
    }
    for (int i = 0; i < TC_MAX_ASTAGES; i++) { ptx::mbar_init(&bar->a_full[i], TC_NLOAD); ptx::mbar_init(&bar->a_empty[i], 1); }
    if (TC_NISSUE == 2) for (int i = 0; i < TC_MAX_ASTAGES; i++) { ptx::mbar_init(&bar->a_full2[0][i], TC_NLOAD); ptx::mbar_init(&bar->a_full2[1][i], TC_NLOAD); }
    ptx::fence_barrier_init();
  }
  tc_fill_lut<NBITS>(s_lut, P.weights, tid, TC_THREADS);
  if (warp == 1) ptx::tmem_alloc(s_tmem, TC_TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const TcCtx S{ring, a_tile0, meta, bar, s_region, s_lut, NA};

  // ===== query-tile loaders: EVERY loader serves EVERY group, loader li fetching the group's queries
  // j = li, li + TC_NLOAD, ...  One warp sustains only ~31 B/clk of 8 KB bulk copies (a copy occupies
  // its issuing warp for ~270 clocks: tools/l2_to_sm_ceiling.cu, profiles/r02_l2_to_sm_ceiling.txt),
  // while the L2 -> SM path itself carries 76 B/clk/SM; spreading one group's four copies over
  // several issuers shortens the fill time of a stage, which is what paces the 3-stage ring. =====

  // Register budget per warpgroup (setmaxnreg sits at the top of each role's branch so ptxas
  // allocates per role): the epilogue keeps two TMEM load batches in flight, the rest need little.
  if (warp < 4) {
  tc_reg_budget<TC_REG_CTRL>();
  if (warp == 0) {
    tc_scheduler_role(P, S, warp, lane);
  } else if (warp == 1) {
    // ===== MMA issuer: ONE elected thread runs the whole loop (no per-group elect / reconvergence);
    // stage and parity counters are carried incrementally and every descriptor is a precomputed low
    // word plus a constant, so a group costs a few dozen instructions.  The issuer is the serial
    // resource of the kernel: at N = 80 a group's 8 MMAs are only ~320 tensor clocks. =====
#if TC_NISSUE == 2
    if (ptx::elect_one()) tc_issuer_role(S, tmem_base, 0, warp, lane);
#else
    if (ptx::elect_one()) {
      [[maybe_unused]] const uint32_t a_lo0 = ((ptx::smem_u32(a_tile0) & 0x3ffffu) >> 4) | (1u << 16);   // descriptor low words: start
      const uint32_t ring_lo = ((ptx::smem_u32(ring) & 0x3ffffu) >> 4) | (1u << 16);    // address >> 4, LBO field = 1
      [[maybe_unused]] constexpr uint32_t HI_A = (2048u >> 4) | (1u << 14) | (2u << 29);   // SBO 2048 | version 1 | SWIZZLE_128B
      constexpr uint32_t HI_B = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024
      uint32_t st = 0, a_par = 0;        // query-tile stage / parity of its next a_full phase
      uint32_t ds = 0, d_par = 1;        // accumulator / parity of its next d_empty phase
      uint32_t eset = 0;                 // epilogue set that consumes the next group
      [[maybe_unused]] int gc = 0;       // running group count (TC_PROF trace)
      for (int e = 0;; e++) {
        const int slot = e & (TC_NSLOT - 1);
        const uint32_t ph = (e >> TC_NSLOT_LOG2) & 1;
        TCW(&bar->meta_full[slot], ph, 4);
        const int ncand = meta[slot].ncand;
        if (ncand < 0) break;
        const int nchunk = meta[slot].nchunk, n0 = meta[slot].n0, n1 = meta[slot].n1;
        const uint32_t b_lo0 = ring_lo + (meta[slot].b_off >> 4);
        const uint32_t kb0 = (uint32_t)n0 * 8u, kb1 = (uint32_t)n1 * 8u;   // K-block stride (rows * 128 B) >> 4
        const uint32_t idesc0 = ptx::idesc_f16(128, n0, 0), idesc1 = ptx::idesc_f16(128, n1 > 0 ? n1 : 16, 0);
        const int ngroups = (ncand + 3) >> 2;
        TCW(&bar->b_full[slot], ph, 5);
        for (int g = 0; g < ngroups; g++) {
          TCW(&bar->a_full[st], a_par, 6);
          const uint32_t a_lo = a_lo0 + st * (uint32_t)(TC_A_BYTES >> 4);
          TCT(0, gc);
          for (int c = 0; c < nchunk; c++) {
            TCW(&bar->d_empty[ds], d_par, 7);
            ptx::tc_fence_after();
            TCT(1, gc);
            const uint32_t d_tmem = tmem_base + ds * TC_D_COLS;
            const bool lastc = c > 0 && c == nchunk - 1;                   // chunk c starts c * n0 * 256 bytes in; the last of several has n1 rows
            const uint32_t b_lo = b_lo0 + (uint32_t)c * (uint32_t)n0 * 16u, kb = lastc ? kb1 : kb0, idesc = lastc ? idesc1 : idesc0;
            TCP_BEGIN();
#pragma unroll
            for (int k = 0; k < ((TC_ABLATE & 8) ? 2 : 8); k++) {
              const uint64_t db = ((uint64_t)HI_B << 32) | (uint64_t)(b_lo + (uint32_t)(k >> 2) * kb + (uint32_t)(k & 3) * 2);
              const uint64_t da = ((uint64_t)HI_A << 32) | (uint64_t)(a_lo + (uint32_t)((k >> 2) * 64 + (k & 3) * 2));
              ptx::mma_f16_ss(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
            }
            if (c == nchunk - 1) ptx::tc_commit(&bar->a_empty[st]);
            ptx::tc_commit(&bar->d_full[eset][ds]);
            TCP_END(16);
            TCT(2, gc);
            ds = (ds + 1u) & (uint32_t)(TC_NACC_ - 1);
            d_par ^= (ds == 0u) ? 1u : 0u;
          }
          if (++eset == (uint32_t)TC_EPI_SETS) eset = 0;
          if (++st == (uint32_t)NA) { st = 0; a_par ^= 1u; }
          gc++;
        }
        ptx::tc_commit(&bar->b_empty[slot]);   // arrives after the passage's last MMA retires
      }
    }
#endif
    __syncwarp();
  } else {
    tc_loader_role<TC_NLOAD>(P, S, warp - 2, warp, lane);
  }
  } else if (warp < TC_DEC_WARP0) {
    tc_reg_budget<TC_REG_EPI>();
    // ===== epilogue: TMEM -> max over tokens -> sum over query tokens -> pair list =====
    const int q4 = warp & 3;                // TMEM lane quarter == query slot inside the group
    const uint32_t myset = (uint32_t)(warp - 4) >> 2;   // this warp's set scores groups ug % TC_EPI_SETS == myset
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    uint32_t ud = 0, eset = 0;
    [[maybe_unused]] uint32_t turn = 0, udi0 = 0, udi1 = 0;   // TC_NISSUE == 2: issuer of the running group / accumulator passes of each issuer so far
    uint32_t fpar = 0;                      // bit ds = parity of this set's next phase of d_full[myset][ds]
    // Output batching: the (query, key) record of the r-th scored pair of this warp is parked in
    // lane r % 32; every 32 records the whole warp appends them to the per-query lists with 32
    // atomics in flight at once, and the stores that depend on the atomics' results are deferred
    // to the next flush (software pipelining), so no global latency sits on the per-group path.
    int cnt = 0, my_q = 0;
    [[maybe_unused]] int gc = 0;
    uint64_t my_key = 0;
    bool pend = false;
    uint64_t pend_key = 0;
    uint64_t* pend_ptr = nullptr;
    int pend_pos = 0;
    auto flush = [&]() {
      if (pend) pend_ptr[pend_pos] = pend_key;
      pend = lane < cnt;
      if (pend) {
        pend_ptr = P.pairs + P.list_off[my_q];
        pend_pos = atomicAdd(&P.cursors[my_q], 1);
        pend_key = my_key;
      }
      cnt = 0;
    };
    for (int e = 0;; e++) {
      const int slot = e & (TC_NSLOT - 1);
      TCW(&bar->meta_full[slot], (e >> TC_NSLOT_LOG2) & 1, 10);
      const Meta& m = meta[slot];
      const int ncand = m.ncand;
      if (ncand < 0) break;
      const int nchunk = m.nchunk, n0 = m.n0, n1 = m.n1;
      const uint32_t pid_inv = 0xffffffffu - (uint32_t)m.pid;
      const int ngroups = (ncand + 3) >> 2;
      for (int g = 0; g < ngroups; g++) {
        const bool mine = (eset == myset);
        if (++eset == (uint32_t)TC_EPI_SETS) eset = 0;
#if TC_NISSUE == 2
        const uint32_t iss = turn;
        turn ^= 1u;
        if (!mine) { if (iss) udi1 += nchunk; else udi0 += nchunk; }
#endif
        if (!mine) { ud += nchunk; gc++; continue; }
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // 4 chains: ILP for the ALU pipe
        for (int c = 0; c < nchunk; c++, ud++) {
#if TC_NISSUE == 2
          const int ds = (int)(iss + 2u * ((iss ? udi1 : udi0) & (uint32_t)(TC_NACC_ / 2 - 1)));   // issuer iss cycles over accumulators iss, iss + 2, ...
          if (iss) udi1++; else udi0++;
#else
          const int ds = ud & (TC_NACC_ - 1);
#endif
          const int ncol = (TC_ABLATE & 2) ? 16 : ((c > 0 && c == nchunk - 1) ? n1 : n0);
          const int nfull = ncol >> 5;           // full 32-column chunks (<= 7)
          TCW(&bar->d_full[myset][ds], (fpar >> ds) & 1u, 11, TC_BACKOFF_EPI_NS);
          fpar ^= 1u << ds;
          ptx::tc_fence_after();
          if (warp == 4 || warp == 8) TCT(3, gc);
          const uint32_t taddr = tmem_base + ds * TC_D_COLS + lane_off;
          // max over the chunk's tokens == max over this thread's ncol TMEM columns; the load of
          // the next 32 columns is in flight while the current 32 are folded
          uint32_t ra[32], rt[16];
          if constexpr (TC_EPI_MODE == 2) {
            // Every load of a step (two 32-column loads + the 16-column tail: 80 columns, the typical passage) is issued
            // before the ONE tcgen05.wait::ld that covers them (a wait costs > 100 clocks whatever it waits for); once the
            // last step's values are in registers the accumulator goes back to the MMA issuer, and only then are they folded.
            uint32_t rb[32];
            int col = 0;
            while (true) {
              const int rem = ncol - col;
              const int n32 = min(rem >> 5, 2);
              const bool t16 = (rem - n32 * 32) == 16;
              if (n32 > 0) ptx::tmem_ld_32x32b_x32(taddr + col, ra);
              if (n32 > 1) ptx::tmem_ld_32x32b_x32(taddr + col + 32, rb);
              if (t16) ptx::tmem_ld_32x32b_x16(taddr + col + n32 * 32, rt);
              ptx::tmem_ld_wait();
              col += n32 * 32 + (t16 ? 16 : 0);
              const bool last = col >= ncol;
              if (last) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bar->d_empty[ds]);
              }
              if (n32 > 0) fold32(ra, m0, m1, m2, m3);
              if (n32 > 1) fold32(rb, m0, m1, m2, m3);
              if (t16) fold16(rt, m0, m1, m2, m3);
              if (last) break;
            }
          } else {
          if constexpr (TC_EPI_MODE == 0) {
            // two sets share the work: single-buffered loads keep a set inside the 96 registers every
            // warp of the 640-thread CTA gets (no setmaxnreg: the pool is only what the launch allocated)
#pragma unroll 1
            for (int i = 0; i < nfull; i++) {
              ptx::tmem_ld_32x32b_x32(taddr + i * 32, ra);
              ptx::tmem_ld_wait();
              fold32(ra, m0, m1, m2, m3);
            }
            if (ncol & 16) { ptx::tmem_ld_32x32b_x16(taddr + nfull * 32, rt); ptx::tmem_ld_wait(); }
          } else {
          uint32_t rb[32];
          if (nfull > 0) ptx::tmem_ld_32x32b_x32(taddr, ra);
          else ptx::tmem_ld_32x32b_x16(taddr, rt);
          ptx::tmem_ld_wait();
          for (int i = 0; i < nfull; i += 2) {
            if (i + 1 < nfull) ptx::tmem_ld_32x32b_x32(taddr + (i + 1) * 32, rb);
            else if (ncol & 16) ptx::tmem_ld_32x32b_x16(taddr + nfull * 32, rt);
            fold32(ra, m0, m1, m2, m3);
            ptx::tmem_ld_wait();
            if (i + 1 >= nfull) break;
            if (i + 2 < nfull) ptx::tmem_ld_32x32b_x32(taddr + (i + 2) * 32, ra);
            else if (ncol & 16) ptx::tmem_ld_32x32b_x16(taddr + nfull * 32, rt);
            fold32(rb, m0, m1, m2, m3);
            ptx::tmem_ld_wait();
          }
          }
          // every column of this accumulator is in registers: hand it back to the MMA issuer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bar->d_empty[ds]);
          if (ncol & 16) fold16(rt, m0, m1, m2, m3);
          }
        }
        if (warp == 4 || warp == 8) TCT(4, gc);
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        // sum over the 32 query tokens: exact (order-free) integer warp reduction of 2^-18 fixed point.
        // Range: the operand rows are unit vectors, so |mx| <= |q token|; the 32-term sum cannot
        // overflow while every |q token| <= 255 (cb_tc_prep_rows flags a batch that breaks this and the
        // batch is scored by the generic fp32 kernel instead).  Resolution 3.8e-6 per token: <= 6.1e-5
        // absolute on a score, far inside the 1e-3 relative tolerance.
        const int isum = __reduce_add_sync(0xffffffffu, __float2int_rn(mx * 262144.0f));
        const int qi = g * 4 + q4;
        if (qi < ncand && !((TC_ABLATE & 16) && isum != 0x7fffffff)) {
          const float score = (float)isum * (1.0f / 262144.0f);
          if (lane == (cnt & 31)) {
            my_q = m.q[qi];
            my_key = ((uint64_t)cb_orderable(score) << 32) | pid_inv;
          }
          if (++cnt == 32) flush();
        }
        if (warp == 4 || warp == 8) TCT(5, gc);
        gc++;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->meta_empty[slot]);
    }
    flush();
    flush();
  } else {
    tc_reg_budget<TC_REG_DEC>();
    if (warp < TC_XISS_WARP0) {
      tc_loader_role<TC_NLOAD>(P, S, warp - TC_DEC_WARP0 + 2, warp, lane);
    } else if (warp < TC_DEC_FIRST) {
#if TC_NISSUE == 2
      if (warp == TC_XISS_WARP0) { if (ptx::elect_one()) tc_issuer_role(S, tmem_base, 1, warp, lane); __syncwarp(); }
#endif
    } else {
    // ===== decompression: packed codes/residuals -> normalised fp16 operand tile(s) =====
    // Eight lanes per token, four tokens per warp-round (fewer, wider instructions per token than a
    // finer split).  What matters besides instruction count is memory-level parallelism (code ->
    // centroid row is a dependent pair of loads, ~1-2k clocks under load), so (a) the eight warps
    // form TC_NTEAMS teams that expand alternate passages concurrently, and (b) a team expands a
    // passage in balanced batches of up to TC_DBATCH rounds with every load of a batch issued
    // before any of it is consumed and the codes of the next batch already requested.  Operand
    // rows past the last token (padding to 16) re-expand the last token: no column masking later.
    tc_decompress_role<NBITS>(P, S, (warp - TC_DEC_FIRST) / TC_TEAM_WARPS, (warp - TC_DEC_FIRST) % TC_TEAM_WARPS, warp, lane);
    }
  }

#if TC_PROF
  if (lane == 0) s_prof[warp * 24] = (unsigned long long)(clock64() - prof_t0);
#endif
  ptx::tc_fence_before();
  __syncthreads();
#if TC_PROF
  for (int i = tid; i < 32 * 24; i += TC_THREADS) g_tc_prof[(size_t)blockIdx.x * 32 * 24 + i] = s_prof[i];
#endif
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}


// Parity hook (cb_debug_tc_operand): the decompression of the scoring kernel -- the SAME device functions,
// one team of TC_TEAM_WARPS warps per listed passage -- with the operand tile copied out un-swizzled.
// out_norm / out_raw: fp16 [sum of doclens][128]; passages the tcgen05 kernel does not take (empty, or longer
// than long_limit) leave their rows untouched.
template <int NBITS>
__global__ void __launch_bounds__(32 * TC_TEAM_WARPS)
k_tc_dump(TcParams P, const int32_t* __restrict__ pids, const int64_t* __restrict__ out_off, __half* __restrict__ out_norm,
          __half* __restrict__ out_raw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tile0 = smem;                                   // up to TC_MAX_CHUNKS * (TC_MAX_BROWS + 16) rows
  uint8_t* s_lut = smem + TC_MAX_CHUNKS * (TC_MAX_BROWS + 16) * 256;
  const int tid = threadIdx.x, dw = tid >> 5, lane = tid & 31;
  tc_fill_lut<NBITS>(s_lut, P.weights, tid, 32 * TC_TEAM_WARPS);
  __syncthreads();
  const int64_t p = pids[blockIdx.x];
  const int64_t e0 = P.offsets[p];
  const int L = (int)(P.offsets[p + 1] - e0);
  if (L <= 0 || L > P.long_limit) return;
  int nchunk, n0, n1;
  tc_tile_geometry(L, nchunk, n0, n1);
  const uint8_t* lut_lane = s_lut + (lane & (LutGeom<NBITS>::REPLICAS - 1)) * LutGeom<NBITS>::ENTRY_BYTES;
  tc_decompress_passage<NBITS, true>(P, lut_lane, dw, lane, L, nchunk, n0, n1, e0, tile0, out_raw + out_off[blockIdx.x] * TC_DIM);
  __syncthreads();
  // row rr of chunk c, 16-byte chunk j (K-block j >> 3): the inverse of the address finish_token16 wrote
  for (int i = tid; i < L * 16; i += 32 * TC_TEAM_WARPS) {
    const int rr = i >> 4, j = i & 15;
    const int c = (rr >= n0 ? 1 : 0) + (rr >= 2 * n0 ? 1 : 0), row = rr - c * n0;
    const uint8_t* t = tile0 + c * n0 * 256 + (j >> 3) * (((nchunk > 1 && c == nchunk - 1) ? n1 : n0) * 128);
    const uint4 v = *reinterpret_cast<const uint4*>(t + (row >> 3) * 1024 + (row & 7) * 128 + (((j & 7) ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(out_norm + (out_off[blockIdx.x] + rr) * TC_DIM + j * 8) = v;
  }
}

size_t tc_fixed_smem_bytes(int nbits) {
  (void)nbits;
  return 1024 + (size_t)TC_NSLOT * sizeof(Meta) + sizeof(Barriers) + 16 + 8 * TC_NSLOT + 64 + 128 + TC_LUT_BYTES + 64;
}

}  // namespace

// Shape: dim 128, 32 query tokens, nbits 1 / 2 / 4.  Range: the decompression squares centroid + weight in packed
// fp16 (two-term chains), which overflows beyond ~180 per component: an index whose centroids or weights are not
// of order 1 (every ColBERT index is: centroids are k-means means of unit vectors) goes to the generic fp32 kernel.
bool cb_stage34_tc_supported(const cb_index* ix, int T) {
  return ix->dim == TC_DIM && T == TC_T && (ix->nbits == 1 || ix->nbits == 2 || ix->nbits == 4) &&
         (ix->centroid_norm_max + ix->weight_abs_max <= 100.f);
}

__global__ void k_collect_long_tc(const int64_t* __restrict__ offsets, int64_t Np, int64_t limit,
                                  int32_t* __restrict__ out /* [0] = count, [1..] pids */) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < Np && offsets[p + 1] - offsets[p] > limit) out[1 + atomicAdd(&out[0], 1)] = (int32_t)p;
}

int32_t cb_stage34_tc(cb_index* ix, const float* dQ, int nq, int T, int W, const uint32_t* d_bitmap,
                      const int64_t* d_list_off, int32_t* d_cursors, uint64_t* d_pairs, cudaStream_t st) {
  CB_REQUIRE(cb_stage34_tc_supported(ix, T), CB_ERR_UNSUPPORTED, "shape not supported by the tcgen05 scoring kernel");
  if (nq == 0 || ix->Np == 0) return CB_OK;
  // shared-memory split: query-tile stages vs. the operand-tile ring
  const size_t budget = 232448 - (TC_PROF ? 7168 : 0);  // 227 KB opt-in shared memory per CTA on sm_100 (minus the static counters of a TC_PROF build)
  int nast = ix->opt_tc_astages > 0 ? ix->opt_tc_astages : 3;
  if (nast < 2) nast = 2;
  if (nast > TC_MAX_ASTAGES) nast = TC_MAX_ASTAGES;
  const size_t fixed = tc_fixed_smem_bytes(ix->nbits);
  CB_REQUIRE(fixed + (size_t)nast * TC_A_BYTES + 16 * 256 <= budget, CB_ERR_UNSUPPORTED,
             "internal: tcgen05 kernel shared memory does not fit");
  const size_t ring = (budget - fixed - (size_t)nast * TC_A_BYTES) & ~(size_t)1023;
  // the ring must hold the largest passage the kernel takes; longer ones go to the generic kernel
  int64_t long_limit = TC_MAX_CHUNKS * TC_MAX_BROWS;
  {
    const int64_t ring_rows = (int64_t)(ring / 256) - 16 * TC_MAX_CHUNKS;    // chunk padding: up to 15 extra rows per chunk
    if (long_limit > ring_rows) long_limit = ring_rows;
    if (long_limit < 1) long_limit = 1;
  }
  const size_t smem = fixed + (size_t)nast * TC_A_BYTES + ring;

  // per-batch query row image (fp16, swizzled): stage 1 has usually built it already
  if (ix->q_prep_src != dQ || ix->q_prep_rows != (int64_t)nq * TC_T) {
    CB_TRY(ix->q_prep.ensure((size_t)nq * TC_Q_BYTES));
    CB_CUDA(cudaMemsetAsync(ix->q_flag.p, 0, sizeof(int), st));
    CB_TRY(cb_tc_prep_rows(dQ, (int64_t)nq * TC_T, (int64_t)nq * TC_T, ix->q_prep.as<uint8_t>(), st, ix->q_flag.as<int>()));
    ix->q_prep_src = dQ;
    ix->q_prep_rows = (int64_t)nq * TC_T;
  }
  TcParams P{};
  P.centroids_h = ix->centroids_h; P.weights = ix->weights; P.codes = ix->codes; P.residuals = ix->residuals;
  P.offsets = ix->offsets; P.Np = ix->Np; P.R = ix->R; P.W = W; P.nastages = nast;
  P.ring_bytes = (int)ring; P.long_limit = (int)long_limit;
  P.qprep = ix->q_prep.as<uint8_t>(); P.bitmap = d_bitmap; P.list_off = d_list_off; P.cursors = d_cursors; P.pairs = d_pairs;
  P.pid_list = ix->tc_active_list; P.n_list = ix->tc_active_n;
  P.q_flag = ix->q_flag.as<int>();
  P.stats = cb_stats_dev(ix);
  if (P.pid_list && P.n_list == 0) return CB_OK;
  int64_t grid = ix->sm_count;
  const int64_t n_items = P.pid_list ? P.n_list : ix->Np;
  if (grid > n_items) grid = n_items;
#define CB_TC_KERNEL k_maxsim_tc
#define CB_TC_NTHREADS TC_THREADS
#define CB_TC_LAUNCH(NB)                                                                                         \
  do {                                                                                                           \
    CB_CUDA(cudaFuncSetAttribute(CB_TC_KERNEL<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    CB_TC_KERNEL<NB><<<(unsigned)grid, CB_TC_NTHREADS, smem, st>>>(P);                                           \
  } while (0)
  if (ix->nbits == 1) CB_TC_LAUNCH(1);
  else if (ix->nbits == 2) CB_TC_LAUNCH(2);
  else CB_TC_LAUNCH(4);
#undef CB_TC_LAUNCH
  CB_LAUNCH_CHECK();

  // passages longer than the kernel takes: generic kernel on just those
  if (ix->max_doclen > long_limit) {
    if (ix->n_long < 0 || ix->long_limit != (int32_t)long_limit) {   // the list is a property of the index: build once
      CB_TRY(ix->long_list.ensure(sizeof(int32_t) * (size_t)(ix->Np + 1)));
      int32_t* d_long = ix->long_list.as<int32_t>();
      CB_CUDA(cudaMemsetAsync(d_long, 0, sizeof(int32_t), st));
      k_collect_long_tc<<<(unsigned)((ix->Np + 255) / 256), 256, 0, st>>>(ix->offsets, ix->Np, long_limit, d_long);
      CB_LAUNCH_CHECK();
      int32_t* h = reinterpret_cast<int32_t*>(ix->pinned_total + 5);
      CB_CUDA(cudaMemcpyAsync(h, d_long, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CB_CUDA(cudaStreamSynchronize(st));
      ix->n_long = *h;
      ix->long_limit = (int32_t)long_limit;
    }
    if (ix->n_long > 0)
      CB_TRY(cb_stage34_generic(ix, dQ, nq, T, W, d_bitmap, ix->long_list.as<int32_t>() + 1, ix->n_long, d_list_off,
                                d_cursors, d_pairs, st));
  }
  return CB_OK;
}

#if TC_PROF
extern "C" int32_t cb_debug_tc_prof(unsigned long long* out /* [160][32][24] */) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(out, g_tc_prof, sizeof(unsigned long long) * 160 * 32 * 24));
  return CB_OK;
}
extern "C" int32_t cb_debug_tc_trace(long long* out /* [16][4096] */) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 16 * TC_TRACE_N));
  return CB_OK;
}
#endif

// ---------------------------------------------------------------------------------------------
// parity hook: what the tcgen05 kernel's decompression writes into its operand tiles
// ---------------------------------------------------------------------------------------------
__global__ void k_tc_dump_lens(const int64_t* __restrict__ pids, int64_t n, int64_t pid_base, int64_t Np, const int64_t* __restrict__ offsets,
                               int32_t* __restrict__ local, int64_t* __restrict__ lens) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = pids[i] - 1 - pid_base;
  if (p < 0 || p >= Np) { local[i] = 0; lens[i] = -1; return; }
  local[i] = (int32_t)p;
  lens[i] = offsets[p + 1] - offsets[p];
}

extern "C" int32_t cb_debug_tc_operand(cb_index* ix, const int64_t* pids, int64_t n_pids, uint16_t* out_norm, uint16_t* out_raw,
                                       int64_t capacity_rows) {
  CB_REQUIRE(ix != nullptr, CB_ERR_BAD_ARG, "index handle is NULL");
  CB_REQUIRE(cb_stage34_tc_supported(ix, TC_T), CB_ERR_UNSUPPORTED, "the tcgen05 scoring kernel does not take this index (dim / nbits)");
  CB_REQUIRE(n_pids >= 0 && (n_pids == 0 || (pids && out_norm && out_raw)), CB_ERR_BAD_ARG, "bad argument");
  if (n_pids == 0) return CB_OK;
  CB_CUDA(cudaSetDevice(ix->device));
  int64_t *d_pids = nullptr;
  CB_CUDA(cudaMalloc((void**)&d_pids, (size_t)n_pids * (8 + 8 + 8 + 4) + 64));
  struct G { void* p; ~G() { cudaFree(p); } } g{d_pids};
  int64_t* d_lens = d_pids + n_pids;
  int64_t* d_off = d_lens + n_pids;
  int32_t* d_local = reinterpret_cast<int32_t*>(d_off + n_pids);
  CB_CUDA(cudaMemcpy(d_pids, pids, sizeof(int64_t) * n_pids, cudaMemcpyHostToDevice));
  k_tc_dump_lens<<<(unsigned)((n_pids + 255) / 256), 256>>>(d_pids, n_pids, ix->pid_base, ix->Np, ix->offsets, d_local, d_lens);
  CB_LAUNCH_CHECK();
  std::string lens_h((size_t)n_pids * 8, '\0'), off_h((size_t)n_pids * 8, '\0');
  int64_t* lens = reinterpret_cast<int64_t*>(&lens_h[0]);
  int64_t* off = reinterpret_cast<int64_t*>(&off_h[0]);
  CB_CUDA(cudaMemcpy(lens, d_lens, sizeof(int64_t) * n_pids, cudaMemcpyDeviceToHost));
  int64_t total = 0;
  const int64_t long_limit = TC_MAX_CHUNKS * TC_MAX_BROWS;
  for (int64_t i = 0; i < n_pids; i++) {
    CB_REQUIRE(lens[i] >= 0, CB_ERR_BOUNDS, "pid out of range 1:%lld (+ pid_base)", (long long)ix->Np);
    CB_REQUIRE(lens[i] <= long_limit, CB_ERR_UNSUPPORTED, "passage longer than the tcgen05 kernel takes (%lld tokens)", (long long)lens[i]);
    off[i] = total;
    total += lens[i];
  }
  CB_REQUIRE(total <= capacity_rows, CB_ERR_BAD_ARG, "output buffers hold %lld rows, %lld needed", (long long)capacity_rows, (long long)total);
  if (total == 0) return CB_OK;
  CB_CUDA(cudaMemcpy(d_off, off, sizeof(int64_t) * n_pids, cudaMemcpyHostToDevice));
  __half* d_out = nullptr;
  CB_CUDA(cudaMalloc((void**)&d_out, (size_t)total * TC_DIM * 2 * 2));
  G g2{d_out};
  CB_CUDA(cudaMemset(d_out, 0, (size_t)total * TC_DIM * 2 * 2));
  TcParams P{};
  P.centroids_h = ix->centroids_h; P.weights = ix->weights; P.codes = ix->codes; P.residuals = ix->residuals;
  P.offsets = ix->offsets; P.Np = ix->Np; P.R = ix->R; P.long_limit = (int)long_limit;
  const size_t smem = 1024 + (size_t)TC_MAX_CHUNKS * (TC_MAX_BROWS + 16) * 256 + TC_LUT_BYTES + 128;
#define CB_TC_DUMP(NB)                                                                                          \
  do {                                                                                                          \
    CB_CUDA(cudaFuncSetAttribute(k_tc_dump<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    k_tc_dump<NB><<<(unsigned)n_pids, 32 * TC_TEAM_WARPS, smem>>>(P, d_local, d_off, d_out, d_out + total * TC_DIM); \
  } while (0)
  if (ix->nbits == 1) CB_TC_DUMP(1);
  else if (ix->nbits == 2) CB_TC_DUMP(2);
  else CB_TC_DUMP(4);
#undef CB_TC_DUMP
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemcpy(out_norm, d_out, (size_t)total * TC_DIM * 2, cudaMemcpyDeviceToHost));
  CB_CUDA(cudaMemcpy(out_raw, d_out + total * TC_DIM, (size_t)total * TC_DIM * 2, cudaMemcpyDeviceToHost));
  return CB_OK;
}

// index_open.cu -- cb_index_open: the native reader of the on-disk index the reference's Indexer writes
// (src/savers.jl:16-29 `save_codec`, 52-84 `save_chunk`, src/indexing.jl:88-143 plan.json / ivf), i.e. the
// loading half of `Searcher(index_path)` (src/searching.jl:18-59) without the host staging:
//   load_config          (src/loaders.jl:66-74)   -> dim, nbits from config.json
//   load_codec           (src/loaders.jl:10-38)   -> centroids.jld2, bucket_weights.jld2 (cutoffs / avg_residual are not
//                                                    read at search time)
//   ivf.jld2, ivf_lengths.jld2 (src/searching.jl:50-51)
//   load_doclens         (src/loaders.jl:76-89)   -> doclens.<chunk>.jld2, checked against plan.json num_embeddings
//   load_compressed_embs (src/loaders.jl:91-113)  -> <chunk>.codes.jld2, <chunk>.residuals.jld2
// The JLD2 files are mapped, not parsed into Julia objects: arrays go from the page cache to the device.  With
// n_shards > 1 only the chunks overlapping this shard's passage range are touched (per-rank chunk selection) and
// the shard's IVF is rebuilt on the device from its own codes (`_build_ivf`, src/indexing/collection_indexer.jl:349-353).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <sstream>
#include <vector>

#include "common.cuh"
#include "jld2.h"

namespace {

bool read_text(const std::string& path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

// value of a top-level numeric field of a JSON object written by JSON.print (src/savers.jl:110-121, indexing.jl:91-96)
bool json_int(const std::string& js, const char* key, int64_t& v) {
  const std::string pat = std::string("\"") + key + "\"";
  size_t p = js.find(pat);
  if (p == std::string::npos) return false;
  p = js.find(':', p + pat.size());
  if (p == std::string::npos) return false;
  p++;
  while (p < js.size() && (js[p] == ' ' || js[p] == '\t' || js[p] == '\n' || js[p] == '\r')) p++;
  char* end = nullptr;
  const double d = strtod(js.c_str() + p, &end);
  if (end == js.c_str() + p) return false;
  v = (int64_t)d;
  return true;
}

struct Obj {   // one `single_stored_object`
  jld2::File file;
  jld2::Array a;
};

int32_t load_obj(const std::string& path, Obj& o, jld2::DType want, int want_ndims, const char* what) {
  std::string err;
  CB_REQUIRE(o.file.open(path, err), CB_ERR_BAD_ARG, "%s", err.c_str());
  CB_REQUIRE(o.file.read("single_stored_object", o.a, err), CB_ERR_BAD_ARG, "%s: %s", path.c_str(), err.c_str());
  CB_REQUIRE(o.a.dtype == want && o.a.ndims == want_ndims, CB_ERR_DOMAIN, "%s: expected %s (%s, %d-d), found %s, %d-d", path.c_str(), what,
             jld2::dtype_name(want), want_ndims, jld2::dtype_name(o.a.dtype), o.a.ndims);
  return CB_OK;
}

}  // namespace

extern "C" int32_t cb_jld2_read(const char* path, const char* name, int64_t info[11], void* out, int64_t capacity_bytes) {
  CB_REQUIRE(path && info, CB_ERR_BAD_ARG, "NULL argument");
  jld2::File f;
  jld2::Array a;
  std::string err;
  CB_REQUIRE(f.open(path, err), CB_ERR_BAD_ARG, "%s", err.c_str());
  CB_REQUIRE(f.read(name ? name : "single_stored_object", a, err), CB_ERR_BAD_ARG, "%s: %s", path, err.c_str());
  info[0] = a.dtype; info[1] = a.elem_size; info[2] = a.ndims;
  for (int i = 0; i < 8; i++) info[3 + i] = i < a.ndims ? a.dims[i] : 0;
  if (out != nullptr && capacity_bytes > 0) {
    const int64_t n = std::min<int64_t>(capacity_bytes, a.count * a.elem_size);
    if (n > 0) memcpy(out, a.data, (size_t)n);
  }
  return CB_OK;
}

extern "C" int32_t cb_index_open(cb_index** out, const char* index_path, int32_t device, int32_t shard, int32_t n_shards,
                                 int64_t* out_pid_base) {
  CB_REQUIRE(out != nullptr, CB_ERR_BAD_ARG, "out handle pointer is NULL");
  *out = nullptr;
  CB_REQUIRE(index_path != nullptr, CB_ERR_BAD_ARG, "index_path is NULL");
  CB_REQUIRE(n_shards >= 1 && shard >= 0 && shard < n_shards, CB_ERR_BAD_ARG, "shard %d of %d", shard, n_shards);
  const std::string dir = index_path;
  std::string cfg, plan;
  // searching.jl:19-21: "Index at ... does not exist!"; loaders.jl:77: "plan.json not found!"
  CB_REQUIRE(read_text(dir + "/config.json", cfg), CB_ERR_BAD_ARG, "Index at %s does not exist! (no config.json)", index_path);
  CB_REQUIRE(read_text(dir + "/plan.json", plan), CB_ERR_BAD_ARG, "plan.json not found!");
  int64_t dim = 0, nbits = 0, num_chunks = 0, num_embeddings = 0;
  CB_REQUIRE(json_int(cfg, "dim", dim) && json_int(cfg, "nbits", nbits), CB_ERR_BAD_ARG, "config.json lacks dim / nbits");
  CB_REQUIRE(json_int(plan, "num_chunks", num_chunks) && json_int(plan, "num_embeddings", num_embeddings), CB_ERR_BAD_ARG,
             "plan.json lacks num_chunks / num_embeddings");
  CB_REQUIRE(dim > 0 && dim % 8 == 0, CB_ERR_DOMAIN, "dim should be a multiple of 8!");   // loaders.jl:94
  CB_REQUIRE(nbits >= 1 && nbits <= CB_MAX_NBITS, CB_ERR_UNSUPPORTED, "nbits must be in 1..%d (got %lld)", CB_MAX_NBITS, (long long)nbits);
  const int64_t R = dim / 8 * nbits;

  // codec (loaders.jl:10-38; type asserts 27-30)
  Obj cen, bw;
  CB_TRY(load_obj(dir + "/centroids.jld2", cen, jld2::DT_F32, 2, "centroids::Matrix{Float32}"));
  CB_TRY(load_obj(dir + "/bucket_weights.jld2", bw, jld2::DT_F32, 1, "bucket_weights::Vector{Float32}"));
  CB_REQUIRE(cen.a.dims[1] == dim, CB_ERR_BAD_ARG, "centroids are %lld-dimensional, config.json says dim = %lld", (long long)cen.a.dims[1],
             (long long)dim);
  CB_REQUIRE(bw.a.dims[0] == ((int64_t)1 << nbits), CB_ERR_DOMAIN, "bucket_weights should have length 2^nbits!");
  const int64_t K = cen.a.dims[0];

  // doclens of every chunk (loaders.jl:76-89) -- small; they also give the chunk -> passage / embedding map
  std::vector<Obj> dl((size_t)num_chunks);
  std::vector<int64_t> doclens;
  std::vector<int64_t> chunk_e0((size_t)num_chunks + 1, 0);   // first embedding of every chunk
  for (int64_t c = 0; c < num_chunks; c++) {
    CB_TRY(load_obj(dir + "/doclens." + std::to_string(c + 1) + ".jld2", dl[c], jld2::DT_I64, 1, "doclens::Vector{Int}"));
    const int64_t* p = reinterpret_cast<const int64_t*>(dl[c].a.data);
    int64_t s = 0;
    for (int64_t i = 0; i < dl[c].a.count; i++) { CB_REQUIRE(p[i] >= 0, CB_ERR_DOMAIN, "doclens must be non-negative"); s += p[i]; }
    doclens.insert(doclens.end(), p, p + dl[c].a.count);
    chunk_e0[c + 1] = chunk_e0[c] + s;
  }
  CB_REQUIRE(chunk_e0[num_chunks] == num_embeddings, CB_ERR_BAD_ARG, "sum(doclens): %lld, num_embeddings: %lld",
             (long long)chunk_e0[num_chunks], (long long)num_embeddings);   // loaders.jl:86-88
  const int64_t Np_all = (int64_t)doclens.size();

  // passage range of this shard: cut r = first passage boundary whose embedding offset reaches r/n of all embeddings
  std::vector<int64_t> cs((size_t)Np_all + 1, 0);
  for (int64_t i = 0; i < Np_all; i++) cs[i + 1] = cs[i] + doclens[i];
  auto cut = [&](int r) -> int64_t {
    if (r <= 0) return 0;
    if (r >= n_shards) return Np_all;
    const int64_t target = (int64_t)((__int128)num_embeddings * r / n_shards);
    return (int64_t)(std::lower_bound(cs.begin(), cs.end(), target) - cs.begin());
  };
  int64_t p_lo = cut(shard), p_hi = cut(shard + 1);
  if (p_hi < p_lo) p_hi = p_lo;
  const int64_t e_lo = cs[p_lo], e_hi = cs[p_hi], Ne = e_hi - e_lo, Np = p_hi - p_lo;

  // compressed embeddings of the chunks that overlap [e_lo, e_hi) (loaders.jl:91-113)
  std::vector<uint32_t> codes((size_t)Ne);
  std::vector<uint8_t> residuals((size_t)Ne * R);
  for (int64_t c = 0; c < num_chunks; c++) {
    const int64_t a = std::max(e_lo, chunk_e0[c]), b = std::min(e_hi, chunk_e0[c + 1]);
    if (b <= a) continue;                                  // per-rank chunk selection: untouched files are never opened
    Obj co, re;
    CB_TRY(load_obj(dir + "/" + std::to_string(c + 1) + ".codes.jld2", co, jld2::DT_U32, 1, "codes::Vector{UInt32}"));
    CB_TRY(load_obj(dir + "/" + std::to_string(c + 1) + ".residuals.jld2", re, jld2::DT_U8, 2, "residuals::Matrix{UInt8}"));
    const int64_t n_c = chunk_e0[c + 1] - chunk_e0[c];
    CB_REQUIRE(co.a.count == n_c, CB_ERR_BAD_ARG, "chunk %lld: %lld codes, doclens sum to %lld", (long long)(c + 1), (long long)co.a.count,
               (long long)n_c);
    CB_REQUIRE(re.a.dims[0] == n_c && re.a.dims[1] == R, CB_ERR_DOMAIN,
               "chunk %lld: residuals are (%lld, %lld), expected (%lld, %lld)", (long long)(c + 1), (long long)re.a.dims[1],
               (long long)re.a.dims[0], (long long)R, (long long)n_c);
    memcpy(codes.data() + (a - e_lo), co.a.data + (a - chunk_e0[c]) * 4, (size_t)(b - a) * 4);
    memcpy(residuals.data() + (a - e_lo) * R, re.a.data + (a - chunk_e0[c]) * R, (size_t)(b - a) * R);
  }

  // the IVF: as stored for the whole index; a shard rebuilds its own from its codes on the device
  Obj ivf, ivl;
  const int64_t* ivf_p = nullptr;
  const int64_t* ivl_p = nullptr;
  if (n_shards == 1) {
    CB_TRY(load_obj(dir + "/ivf.jld2", ivf, jld2::DT_I64, 1, "ivf::Vector{Int}"));
    CB_TRY(load_obj(dir + "/ivf_lengths.jld2", ivl, jld2::DT_I64, 1, "ivf_lengths::Vector{Int}"));
    CB_REQUIRE(ivl.a.count == K, CB_ERR_BAD_ARG, "length(ivf_lengths) = %lld, %lld centroids", (long long)ivl.a.count, (long long)K);
    CB_REQUIRE(ivf.a.count == Ne, CB_ERR_BAD_ARG, "length(ivf) must be equal to sum(ivf_lengths)! (%lld vs %lld)", (long long)ivf.a.count,
               (long long)Ne);
    ivf_p = reinterpret_cast<const int64_t*>(ivf.a.data);
    ivl_p = reinterpret_cast<const int64_t*>(ivl.a.data);
  }
  if (out_pid_base) *out_pid_base = p_lo;
  return cb_index_create(out, device, (int32_t)dim, (int32_t)nbits, K, Np, Ne, reinterpret_cast<const float*>(cen.a.data),
                         reinterpret_cast<const float*>(bw.a.data), codes.data(), residuals.data(), doclens.data() + p_lo, ivf_p, ivl_p,
                         p_lo, 0);
}

// multi.cu -- passage-sharded search over several GPUs of one box behind the C ABI: ONE process, ONE host thread,
// one stream per device.  The reference is single-process (src/infra/config.jl:57-58: rank / nranks exist but the
// search path never uses them), so a drop-in must be too: `search(searcher, ...)` (src/searching.jl:93-128) maps to
// one cb_multi_search_batch call.
//
// Per batch (nothing below waits for the host until the final read-back):
//   Q           host -> device 0, then device 0 -> every peer over NVLink (one PCIe upload instead of N)
//   stage 1     split by QUERY: device d probes queries [d nq/N, (d+1) nq/N) (every shard holds the same centroids)
//               and pushes its cells into every peer's buffer (peer copies; 256 KB per batch in all)
//   stages 2-5  every device on its own passage range (cb_search_batch_cells_device), all devices concurrently
//   merge       the per-shard top-k lists (nq * k * 12 B each) are pushed to device 0 and merged there by
//               (score desc, pid asc) -- the order of the reference's stable sortperm (src/searching.jl:125-127)
// The exchanges are point-to-point peer copies ordered by CUDA events: two tiny, latency-bound messages per device;
// a NCCL communicator would add nothing here (the torchrun harness in sharding.py uses NCCL for the same exchange).
#include <vector>

#include "common.cuh"

struct cb_multi {
  int n = 0;
  std::vector<cb_index*> shard;
  std::vector<cudaStream_t> stream;
  std::vector<cudaEvent_t> ev_q, ev_cells, ev_lists;
  std::vector<DevBuf> dq, cells, lp, ls, lc;   // per device: queries, all cells, local lists
  DevBuf all_p, all_s, out_p, out_s;           // device 0: gathered lists, merged result
  std::vector<int32_t> h_counts;
  bool owns_shards = false;
};

static void multi_free(cb_multi* m) {
  if (!m) return;
  for (int d = 0; d < m->n; d++) {
    cudaSetDevice(m->shard[d]->device);
    m->dq[d].release(); m->cells[d].release(); m->lp[d].release(); m->ls[d].release(); m->lc[d].release();
    if (m->stream[d]) cudaStreamDestroy(m->stream[d]);
    if (m->ev_q[d]) cudaEventDestroy(m->ev_q[d]);
    if (m->ev_cells[d]) cudaEventDestroy(m->ev_cells[d]);
    if (m->ev_lists[d]) cudaEventDestroy(m->ev_lists[d]);
  }
  if (m->n > 0) {
    cudaSetDevice(m->shard[0]->device);
    m->all_p.release(); m->all_s.release(); m->out_p.release(); m->out_s.release();
  }
  if (m->owns_shards) for (cb_index* ix : m->shard) cb_index_destroy(ix);
  cudaGetLastError();
  delete m;
}

extern "C" int32_t cb_multi_create(cb_multi** out, int32_t n_shards, cb_index* const* shards) {
  CB_REQUIRE(out != nullptr, CB_ERR_BAD_ARG, "out handle pointer is NULL");
  *out = nullptr;
  CB_REQUIRE(n_shards >= 1 && n_shards <= 64 && shards != nullptr, CB_ERR_BAD_ARG, "bad shard list");
  for (int d = 0; d < n_shards; d++) {
    CB_REQUIRE(shards[d] != nullptr, CB_ERR_BAD_ARG, "shard %d is NULL", d);
    CB_REQUIRE(shards[d]->dim == shards[0]->dim && shards[d]->nbits == shards[0]->nbits && shards[d]->K == shards[0]->K, CB_ERR_BAD_ARG,
               "shard %d has another dim / nbits / number of centroids than shard 0", d);
  }
  cb_multi* m = new (std::nothrow) cb_multi();
  CB_REQUIRE(m != nullptr, CB_ERR_OOM, "host allocation failed");
  m->n = n_shards;
  m->shard.assign(shards, shards + n_shards);
  m->stream.assign(n_shards, nullptr); m->ev_q.assign(n_shards, nullptr); m->ev_cells.assign(n_shards, nullptr);
  m->ev_lists.assign(n_shards, nullptr);
  m->dq.resize(n_shards); m->cells.resize(n_shards); m->lp.resize(n_shards); m->ls.resize(n_shards); m->lc.resize(n_shards);
  for (int d = 0; d < n_shards; d++) {
    cudaError_t e = cudaSetDevice(shards[d]->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream[d], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_q[d], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_cells[d], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_lists[d], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      cb_set_error("CUDA error %s while setting up device %d", cudaGetErrorName(e), shards[d]->device);
      multi_free(m);
      return CB_ERR_CUDA;
    }
    for (int p = 0; p < n_shards; p++) {   // direct NVLink access between the devices of the group (already-enabled is fine)
      if (shards[p]->device == shards[d]->device) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, shards[d]->device, shards[p]->device) == cudaSuccess && can) cudaDeviceEnablePeerAccess(shards[p]->device, 0);
      cudaGetLastError();
    }
  }
  *out = m;
  return CB_OK;
}

extern "C" int32_t cb_multi_open(cb_multi** out, const char* index_path, int32_t n_gpus, const int32_t* device_ids) {
  CB_REQUIRE(out != nullptr, CB_ERR_BAD_ARG, "out handle pointer is NULL");
  *out = nullptr;
  CB_REQUIRE(n_gpus >= 1 && n_gpus <= 64, CB_ERR_BAD_ARG, "n_gpus = %d", n_gpus);
  std::vector<cb_index*> sh((size_t)n_gpus, nullptr);
  for (int d = 0; d < n_gpus; d++) {
    const int32_t s = cb_index_open(&sh[d], index_path, device_ids ? device_ids[d] : d, d, n_gpus, nullptr);
    if (s != CB_OK) {
      for (int j = 0; j < d; j++) cb_index_destroy(sh[j]);
      return s;
    }
  }
  const int32_t s = cb_multi_create(out, n_gpus, sh.data());
  if (s != CB_OK) { for (cb_index* ix : sh) cb_index_destroy(ix); return s; }
  (*out)->owns_shards = true;
  return CB_OK;
}

extern "C" int32_t cb_multi_destroy(cb_multi* m) {
  multi_free(m);
  return CB_OK;
}

extern "C" int32_t cb_multi_info(const cb_multi* m, int32_t* n_shards, cb_index** shards, int32_t capacity) {
  CB_REQUIRE(m && n_shards, CB_ERR_BAD_ARG, "NULL argument");
  *n_shards = m->n;
  for (int d = 0; d < m->n && d < capacity && shards; d++) shards[d] = m->shard[d];
  return CB_OK;
}

extern "C" int32_t cb_multi_search_batch(cb_multi* m, const float* Q, int32_t nq, int32_t T, int32_t nprobe, int32_t k,
                                         int64_t* out_pids, float* out_scores, int32_t* out_counts) {
  CB_REQUIRE(m != nullptr, CB_ERR_BAD_ARG, "handle is NULL");
  CB_REQUIRE(nq >= 0 && T >= 1, CB_ERR_BAD_ARG, "bad query shape (nq = %d, T = %d)", nq, T);
  CB_REQUIRE(nq == 0 || (Q && out_pids && out_scores && out_counts), CB_ERR_BAD_ARG, "NULL pointer");
  CB_REQUIRE(nprobe >= 1 && nprobe <= CB_MAX_NPROBE, CB_ERR_UNSUPPORTED, "nprobe must be in 1..%d (got %d)", CB_MAX_NPROBE, nprobe);
  CB_REQUIRE(k >= 1 && k <= CB_MAX_K && (int64_t)m->n * k <= 8192, CB_ERR_UNSUPPORTED, "k must be in 1..%d and n_shards * k <= 8192", CB_MAX_K);
  if (nq == 0) return CB_OK;
  const int n = m->n;
  const int dim = m->shard[0]->dim;
  const size_t qbytes = sizeof(float) * (size_t)nq * T * dim;
  const int per = (nq + n - 1) / n;                                   // queries whose stage 1 one device computes
  const size_t cell_row = sizeof(int32_t) * (size_t)T * nprobe;       // bytes of one query's cells
  const size_t nk = (size_t)nq * k;
  for (int d = 0; d < n; d++) {
    CB_CUDA(cudaSetDevice(m->shard[d]->device));
    CB_TRY(m->dq[d].ensure(qbytes));
    CB_TRY(m->cells[d].ensure(cell_row * (size_t)per * n));
    CB_TRY(m->lp[d].ensure(sizeof(int64_t) * nk));
    CB_TRY(m->ls[d].ensure(sizeof(float) * nk));
    CB_TRY(m->lc[d].ensure(sizeof(int32_t) * (size_t)nq));
  }
  const int dev0 = m->shard[0]->device;
  CB_CUDA(cudaSetDevice(dev0));
  CB_TRY(m->all_p.ensure(sizeof(int64_t) * nk * n));
  CB_TRY(m->all_s.ensure(sizeof(float) * nk * n));
  CB_TRY(m->out_p.ensure(sizeof(int64_t) * nk));
  CB_TRY(m->out_s.ensure(sizeof(float) * nk));

  // queries: host -> device 0 -> peers
  CB_CUDA(cudaMemcpyAsync(m->dq[0].p, Q, qbytes, cudaMemcpyHostToDevice, m->stream[0]));
  for (int d = 1; d < n; d++)
    CB_CUDA(cudaMemcpyPeerAsync(m->dq[d].p, m->shard[d]->device, m->dq[0].p, dev0, qbytes, m->stream[0]));
  CB_CUDA(cudaEventRecord(m->ev_q[0], m->stream[0]));
  // stage 1, split by query; every device pushes its slice of the cells to every peer
  for (int d = 0; d < n; d++) {
    CB_CUDA(cudaSetDevice(m->shard[d]->device));
    if (d > 0) CB_CUDA(cudaStreamWaitEvent(m->stream[d], m->ev_q[0], 0));
    const int lo = d * per < nq ? d * per : nq, hi = lo + per < nq ? lo + per : nq;
    if (hi > lo) {
      int32_t* mine = m->cells[d].as<int32_t>() + (size_t)lo * T * nprobe;
      CB_TRY(cb_probe_device(m->shard[d], m->dq[d].as<float>() + (size_t)lo * T * dim, hi - lo, T, nprobe, mine, m->stream[d]));
      for (int p = 0; p < n; p++)
        if (p != d)
          CB_CUDA(cudaMemcpyPeerAsync(m->cells[p].as<int32_t>() + (size_t)lo * T * nprobe, m->shard[p]->device, mine, m->shard[d]->device,
                                      cell_row * (size_t)(hi - lo), m->stream[d]));
    }
    CB_CUDA(cudaEventRecord(m->ev_cells[d], m->stream[d]));
  }
  // stages 2-5 on every shard, then the local lists go to device 0
  for (int d = 0; d < n; d++) {
    CB_CUDA(cudaSetDevice(m->shard[d]->device));
    for (int p = 0; p < n; p++)
      if (p != d) CB_CUDA(cudaStreamWaitEvent(m->stream[d], m->ev_cells[p], 0));
    CB_TRY(cb_search_batch_cells_device(m->shard[d], m->dq[d].as<float>(), m->cells[d].as<int32_t>(), nq, T, nprobe, k,
                                        m->lp[d].as<int64_t>(), m->ls[d].as<float>(), m->lc[d].as<int32_t>(), m->stream[d]));
    CB_CUDA(cudaMemcpyPeerAsync(m->all_p.as<int64_t>() + nk * d, dev0, m->lp[d].p, m->shard[d]->device, sizeof(int64_t) * nk, m->stream[d]));
    CB_CUDA(cudaMemcpyPeerAsync(m->all_s.as<float>() + nk * d, dev0, m->ls[d].p, m->shard[d]->device, sizeof(float) * nk, m->stream[d]));
    CB_CUDA(cudaEventRecord(m->ev_lists[d], m->stream[d]));
  }
  // merge on device 0 and read back
  CB_CUDA(cudaSetDevice(dev0));
  for (int d = 1; d < n; d++) CB_CUDA(cudaStreamWaitEvent(m->stream[0], m->ev_lists[d], 0));
  if (n > 1) {
    CB_TRY(cb_merge_topk_device(dev0, n, nq, k, m->all_p.as<int64_t>(), m->all_s.as<float>(), m->out_p.as<int64_t>(),
                                m->out_s.as<float>(), m->stream[0]));
    CB_CUDA(cudaMemcpyAsync(out_pids, m->out_p.p, sizeof(int64_t) * nk, cudaMemcpyDeviceToHost, m->stream[0]));
    CB_CUDA(cudaMemcpyAsync(out_scores, m->out_s.p, sizeof(float) * nk, cudaMemcpyDeviceToHost, m->stream[0]));
  } else {
    CB_CUDA(cudaMemcpyAsync(out_pids, m->lp[0].p, sizeof(int64_t) * nk, cudaMemcpyDeviceToHost, m->stream[0]));
    CB_CUDA(cudaMemcpyAsync(out_scores, m->ls[0].p, sizeof(float) * nk, cudaMemcpyDeviceToHost, m->stream[0]));
  }
  // candidate counts: the sum over the shards (every passage lives in exactly one)
  m->h_counts.assign((size_t)nq * n, 0);
  for (int d = 0; d < n; d++) {
    CB_CUDA(cudaSetDevice(m->shard[d]->device));
    CB_CUDA(cudaMemcpyAsync(m->h_counts.data() + (size_t)nq * d, m->lc[d].p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, m->stream[d]));
  }
  for (int d = 0; d < n; d++) {
    CB_CUDA(cudaSetDevice(m->shard[d]->device));
    CB_CUDA(cudaStreamSynchronize(m->stream[d]));
  }
  for (int q = 0; q < nq; q++) {
    int64_t c = 0;
    for (int d = 0; d < n; d++) c += m->h_counts[(size_t)nq * d + q];
    out_counts[q] = (int32_t)c;
  }
  return CB_OK;
}

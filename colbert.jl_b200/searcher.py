"""Host-side mirror of the reference's search interface (src/searching.jl, src/search/ranking.jl,
the decompress half of src/indexing/codecs/residual.jl), written in Python because Julia is not
installed in this image; julia/ColBERTB200.jl is the same shim as `ccall`s.

Same names, argument meaning and error behaviour as the reference; arrays use the reference's
*Julia shapes* (centroids (dim, K), residuals (R, N_e), Q (dim, T[, nq]), 1-based codes / eids /
pids), so a Julia array handed over unchanged is `np.asarray(x).T` of the C layout the ABI wants.
All compute happens in libcolbert_b200.so on the GPU; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from ._lib import BoundsError, DimensionMismatch, DomainError  # noqa: F401  (re-exported)


@dataclass
class ColBERTConfig:
    """The subset of `ColBERTConfig` (src/infra/config.jl:54-90) the search path reads, with the
    reference's defaults."""
    dim: int = 128
    nbits: int = 2
    nprobe: int = 2
    query_maxlen: int = 32
    doc_maxlen: int = 300
    ncandidates: int = 8192   # documented upstream but never applied (SURVEY 3.1 quirk i)
    index_path: str = ""


def _c(a, dtype):
    """C-contiguous array of `dtype` (copy only when needed)."""
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _julia_to_c(a, dtype):
    """Julia-shaped matrix (inner, n) -> C layout [n][inner] without a copy when `a` is the
    transposed view of a C-contiguous array (or F-contiguous)."""
    return _c(np.asarray(a).T, dtype)


def _build_emb2pid(doclens):
    """src/searching.jl:82-91.  The library itself keeps the equivalent passage offsets; this
    host helper exists so code written against the reference keeps working."""
    doclens = np.asarray(doclens, dtype=np.int64)
    return np.repeat(np.arange(1, len(doclens) + 1, dtype=np.int64), doclens)


class Searcher:
    """`struct Searcher` (src/searching.jl:1-16) without the encoder: owns one index (or one
    passage-range shard) resident on a GPU.  Construction uploads the arrays once
    (src/searching.jl:44-59) -- the hot path never touches host memory again."""

    def __init__(self, config: ColBERTConfig, centroids, bucket_cutoffs, bucket_weights, ivf, ivf_lengths,
                 doclens, codes, residuals, device: int = 0, pid_base: int = 0):
        lib = L.load()
        self.config = config
        self.device = device
        self.pid_base = int(pid_base)
        self.bucket_cutoffs = None if bucket_cutoffs is None else _c(bucket_cutoffs, np.float32)
        cen = _julia_to_c(centroids, np.float32)            # [K][dim]
        if cen.ndim != 2 or cen.shape[1] != config.dim:
            raise DimensionMismatch(f"centroids must be (dim={config.dim}, K), got {np.shape(centroids)}")
        w = _c(bucket_weights, np.float32)
        if w.shape != (1 << config.nbits,):
            raise DomainError("bucket_weights should have length 2^nbits!")
        codes_c = _c(codes, np.uint32)
        res = _julia_to_c(residuals, np.uint8)              # [N_e][R]
        R = config.dim // 8 * config.nbits
        if res.shape != (len(codes_c), R):
            raise DomainError("The number of codes should be equal to the number of residual embeddings "
                              f"(residuals must be ({R}, {len(codes_c)}), got {np.shape(residuals)})")
        dl = _c(doclens, np.int64)
        ivf_c = None if ivf is None else _c(ivf, np.int64)
        ivl_c = None if ivf_lengths is None else _c(ivf_lengths, np.int64)
        if ivl_c is not None and len(ivl_c) != cen.shape[0]:
            raise DimensionMismatch("length(ivf_lengths) must equal the number of centroids")
        if ivf_c is not None and len(ivf_c) != len(codes_c):
            raise DimensionMismatch("length(ivf) must be equal to sum(ivf_lengths)!")
        self.K, self.n_passages, self.n_embeddings = cen.shape[0], len(dl), len(codes_c)
        self._h = C.c_void_p()
        L.check(lib.cb_index_create(C.byref(self._h), device, config.dim, config.nbits, self.K, self.n_passages,
                                    self.n_embeddings, _ptr(cen), _ptr(w), _ptr(codes_c), _ptr(res), _ptr(dl),
                                    _ptr(ivf_c), _ptr(ivl_c), self.pid_base, 0))

    @classmethod
    def from_device(cls, config: ColBERTConfig, K, n_passages, n_embeddings, centroids_ptr, bucket_weights_ptr,
                    codes_ptr, residuals_ptr, doclens_ptr, ivf_ptr=None, ivf_lengths_ptr=None, device=0, pid_base=0,
                    borrow_residuals=False):
        """Index whose arrays already live in device memory (raw device pointers, C layouts of the
        header): the hand-off a device-side loader / generator uses.  `borrow_residuals`: the index reads the
        caller's residual array in place (no second copy of the bulk of the index); the caller keeps it alive until close()."""
        lib = L.load()
        self = cls.__new__(cls)
        self.config, self.device, self.pid_base = config, device, int(pid_base)
        self.bucket_cutoffs = None
        self.K, self.n_passages, self.n_embeddings = int(K), int(n_passages), int(n_embeddings)
        self._h = C.c_void_p()
        L.check(lib.cb_index_create(C.byref(self._h), device, config.dim, config.nbits, self.K, self.n_passages,
                                    self.n_embeddings, centroids_ptr, bucket_weights_ptr, codes_ptr, residuals_ptr,
                                    doclens_ptr, ivf_ptr, ivf_lengths_ptr, self.pid_base,
                                    L.CB_FLAG_DEVICE_POINTERS | (L.CB_FLAG_BORROW_RESIDUALS if borrow_residuals else 0)))
        return self

    @classmethod
    def open(cls, index_path: str, device: int = 0, shard: int = 0, n_shards: int = 1):
        """`Searcher(index_path)` (src/searching.jl:18-59) without the encoder: reads the directory the reference's
        Indexer wrote (JLD2 files + config.json / plan.json) natively and uploads it.  `shard` / `n_shards` open one
        passage range of a passage-sharded deployment (only the overlapping chunk files are read)."""
        import json
        import os
        lib = L.load()
        if not os.path.isdir(index_path):
            raise L.ColBERTB200Error(f"Index at {index_path} does not exist! Please build the index first and try again.")
        c = json.load(open(os.path.join(index_path, "config.json")))
        self = cls.__new__(cls)
        self.config = ColBERTConfig(dim=int(c["dim"]), nbits=int(c["nbits"]), nprobe=int(c.get("nprobe", 2)),
                                    query_maxlen=int(c.get("query_maxlen", 32)), doc_maxlen=int(c.get("doc_maxlen", 300)),
                                    ncandidates=int(c.get("ncandidates", 8192)), index_path=index_path)
        self.device, self.bucket_cutoffs = device, None
        self._h = C.c_void_p()
        base = C.c_int64()
        L.check(lib.cb_index_open(C.byref(self._h), index_path.encode(), device, shard, n_shards, C.byref(base)))
        self.pid_base = int(base.value)
        info = self.info()
        self.K, self.n_passages, self.n_embeddings = info["K"], info["n_passages"], info["n_embeddings"]
        return self

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            L.load().cb_index_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- knobs / counters ----------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        L.check(L.load().cb_set_option(self._h, key.encode(), int(value)))

    def stat(self, key: str) -> float:
        v = C.c_double()
        L.check(L.load().cb_get_stat(self._h, key.encode(), C.byref(v)))
        return v.value

    def info(self):
        a = (C.c_int64 * 8)()
        L.check(L.load().cb_index_info(self._h, a))
        return dict(zip(("dim", "nbits", "K", "n_passages", "n_embeddings", "device", "pid_base", "bytes"), list(a)))

    # -- the hot path --------------------------------------------------------------------------
    def _q_batch(self, Q):
        """(dim, T) or (dim, T, nq) Julia-shaped -> C [nq][T][dim] float32."""
        Q = np.asarray(Q)
        if Q.ndim == 2:
            Q = Q[:, :, None]
        if Q.ndim != 3 or Q.shape[0] != self.config.dim:
            raise DimensionMismatch(f"Q must be (dim={self.config.dim}, T[, nq]), got {Q.shape}")
        return _c(np.transpose(Q, (2, 1, 0)), np.float32)

    def search_batch(self, Q, k: int, nprobe: int | None = None):
        """Batched `search` minus the encoder: Q (dim, T, nq) -> (pids (nq, k) int64 1-based,
        scores (nq, k) float32, counts (nq,) int32 = number of candidates per query).  Slots
        beyond counts[q] hold pid 0 / -inf (the single-query `search` raises instead)."""
        Qc = self._q_batch(Q)
        nq, T, _ = Qc.shape
        nprobe = self.config.nprobe if nprobe is None else nprobe
        pids = np.zeros((nq, k), dtype=np.int64)
        scores = np.full((nq, k), -np.inf, dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        L.check(L.load().cb_search_batch(self._h, _ptr(Qc), nq, T, nprobe, k, _ptr(pids), _ptr(scores), _ptr(counts)))
        return pids, scores, counts

    def search_batch_device(self, q_ptr, nq, T, k, out_pids_ptr, out_scores_ptr, out_counts_ptr, stream=None,
                            nprobe=None):
        """Device-resident variant: raw device pointers (e.g. torch `.data_ptr()`), Q as C
        [nq][T][dim] float32; work is enqueued on `stream` (a cudaStream_t as int)."""
        nprobe = self.config.nprobe if nprobe is None else nprobe
        L.check(L.load().cb_search_batch_device(self._h, q_ptr, nq, T, nprobe, k, out_pids_ptr, out_scores_ptr,
                                                out_counts_ptr, stream))

    def probe_device(self, q_ptr, nq, T, out_cells_ptr, stream=None, nprobe=None):
        """Stage 1 alone on device buffers: int32 cells [nq][T][nprobe] (1-based, 0 = none)."""
        nprobe = self.config.nprobe if nprobe is None else nprobe
        L.check(L.load().cb_probe_device(self._h, q_ptr, nq, T, nprobe, out_cells_ptr, stream))

    def search_batch_cells_device(self, q_ptr, cells_ptr, nq, T, k, out_pids_ptr, out_scores_ptr, out_counts_ptr,
                                  stream=None, nprobe=None):
        """`search_batch_device` with the stage-1 cells supplied by the caller (see `probe_device`)."""
        nprobe = self.config.nprobe if nprobe is None else nprobe
        L.check(L.load().cb_search_batch_cells_device(self._h, q_ptr, cells_ptr, nq, T, nprobe, k, out_pids_ptr,
                                                      out_scores_ptr, out_counts_ptr, stream))

    def search_batch_plaid(self, Q, k: int, ncells: int = 4, centroid_score_threshold: float = 0.4, ndocs: int = 1000):
        """PLAID-style pruned search (BASELINE config 5; no reference counterpart, semantics in
        oracle.plaid_search): candidates from `ncells` cells per token, centroid-score threshold
        pruning, approximate centroid-only scores, the best `ndocs` re-scored exactly, top-k.
        Q (dim, T, nq) -> (pids (nq, k), scores (nq, k), counts (nq,) = exactly scored passages)."""
        Qc = self._q_batch(Q)
        nq, T, _ = Qc.shape
        pids = np.zeros((nq, k), dtype=np.int64)
        scores = np.full((nq, k), -np.inf, dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        L.check(L.load().cb_search_batch_plaid(self._h, _ptr(Qc), nq, T, ncells, float(centroid_score_threshold), ndocs, k,
                                               _ptr(pids), _ptr(scores), _ptr(counts)))
        return pids, scores, counts

    def search_batch_plaid_device(self, q_ptr, nq, T, k, out_pids_ptr, out_scores_ptr, out_counts_ptr, ncells=4,
                                  centroid_score_threshold=0.4, ndocs=1000, stream=None):
        """Device-resident variant of `search_batch_plaid` (raw device pointers, work on `stream`)."""
        L.check(L.load().cb_search_batch_plaid_device(self._h, q_ptr, nq, T, ncells, float(centroid_score_threshold), ndocs,
                                                      k, out_pids_ptr, out_scores_ptr, out_counts_ptr, stream))

    def probe(self, Q, nprobe=None):
        """Stage 1: `_topk(Q' * centroids, nprobe, dims = 2)`: (nq, T, nprobe) 1-based centroid ids
        (best first) and their fp32 scores."""
        Qc = self._q_batch(Q)
        nq, T, _ = Qc.shape
        nprobe = self.config.nprobe if nprobe is None else nprobe
        cells = np.zeros((nq, T, nprobe), dtype=np.int32)
        scores = np.zeros((nq, T, nprobe), dtype=np.float32)
        L.check(L.load().cb_probe(self._h, _ptr(Qc), nq, T, nprobe, _ptr(cells), _ptr(scores)))
        return cells, scores

    def retrieve(self, Q, nprobe=None):
        """`retrieve` (src/search/ranking.jl:23-44) on the resident index: sorted unique 1-based
        candidate pids of ONE query Q (dim, T)."""
        Qc = self._q_batch(Q)
        if Qc.shape[0] != 1:
            raise DimensionMismatch("retrieve takes one query")
        nprobe = self.config.nprobe if nprobe is None else nprobe
        lib = L.load()
        n = C.c_int64()
        L.check(lib.cb_retrieve(self._h, _ptr(Qc), Qc.shape[1], nprobe, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int64)
        if n.value:
            L.check(lib.cb_retrieve(self._h, _ptr(Qc), Qc.shape[1], nprobe, _ptr(out), n.value, C.byref(n)))
        return out

    def score_pids(self, Q, pids):
        """Stages 3+4 in place: MaxSim scores of `pids` (1-based) for ONE query -- what
        `_collect_compressed_embs_for_pids` -> `decompress` -> `maxsim` computes."""
        Qc = self._q_batch(Q)
        pids = _c(pids, np.int64)
        out = np.zeros(len(pids), dtype=np.float32)
        L.check(L.load().cb_score_pids(self._h, _ptr(Qc), Qc.shape[1], _ptr(pids), len(pids), _ptr(out)))
        return out

    def debug_tc_operand(self, pids, n_rows: int):
        """Parity hook: the fp16 operand rows the tcgen05 scoring kernel's decompression produces for
        `pids` (1-based): (normalised (n_rows, dim) float16, un-normalised centroid+weight (n_rows, dim)
        float16), passages concatenated in the order given.  n_rows = sum of their doclens."""
        pids = _c(pids, np.int64)
        norm = np.zeros((n_rows, self.config.dim), dtype=np.float16)
        raw = np.zeros((n_rows, self.config.dim), dtype=np.float16)
        L.check(L.load().cb_debug_tc_operand(self._h, _ptr(pids), len(pids), _ptr(norm), _ptr(raw), n_rows))
        return norm, raw


class MultiSearcher:
    """The passage-range shards of one index on several GPUs of one box, driven by ONE host thread through
    cb_multi_* (no torch.distributed, no process per GPU): what a single-process caller such as the reference's
    `search(searcher, query, k)` needs.  `searchers` = Searcher objects (one per shard, each with its pid_base, on
    its device); or use `MultiSearcher.open(index_path, n_gpus)` for an index directory."""

    def __init__(self, searchers):
        lib = L.load()
        self.searchers = list(searchers)
        self.config = self.searchers[0].config
        arr = (C.c_void_p * len(self.searchers))(*[s._h for s in self.searchers])
        self._m = C.c_void_p()
        L.check(lib.cb_multi_create(C.byref(self._m), len(self.searchers), arr))

    @classmethod
    def open(cls, index_path: str, n_gpus: int, device_ids=None):
        import json
        import os
        lib = L.load()
        self = cls.__new__(cls)
        c = json.load(open(os.path.join(index_path, "config.json")))
        self.config = ColBERTConfig(dim=int(c["dim"]), nbits=int(c["nbits"]), nprobe=int(c.get("nprobe", 2)),
                                    query_maxlen=int(c.get("query_maxlen", 32)), index_path=index_path)
        self.searchers = []
        ids = None if device_ids is None else (C.c_int32 * n_gpus)(*device_ids)
        self._m = C.c_void_p()
        L.check(lib.cb_multi_open(C.byref(self._m), index_path.encode(), n_gpus, ids))
        return self

    def search_batch(self, Q, k: int, nprobe: int | None = None):
        """Q (dim, T, nq) -> (pids (nq, k), scores (nq, k), counts (nq,)) over the whole index."""
        Q = np.asarray(Q)
        if Q.ndim == 2:
            Q = Q[:, :, None]
        Qc = _c(np.transpose(Q, (2, 1, 0)), np.float32)
        nq, T, _ = Qc.shape
        nprobe = self.config.nprobe if nprobe is None else nprobe
        pids = np.zeros((nq, k), dtype=np.int64)
        scores = np.full((nq, k), -np.inf, dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        L.check(L.load().cb_multi_search_batch(self._m, _ptr(Qc), nq, T, nprobe, k, _ptr(pids), _ptr(scores), _ptr(counts)))
        return pids, scores, counts

    def search_batch_ptr(self, q_host_ptr, nq, T, k, pids_ptr, scores_ptr, counts_ptr, nprobe=None):
        nprobe = self.config.nprobe if nprobe is None else nprobe
        L.check(L.load().cb_multi_search_batch(self._m, q_host_ptr, nq, T, nprobe, k, pids_ptr, scores_ptr, counts_ptr))

    def close(self):
        m, self._m = getattr(self, "_m", None), None
        if m:
            L.load().cb_multi_destroy(m)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def search(searcher: Searcher, Q, k: int):
    """`search(searcher, query, k)` (src/searching.jl:93-128) minus `encode_queries`: Q is the
    encoder's output for the query, (dim, query_maxlen) -- or (dim, query_maxlen, 1).
    Returns (pids[1:k], scores[1:k]); raises BoundsError when fewer than k candidates exist, as
    `pids[1:k]` does upstream (searching.jl:127).  Q (dim, T, nq) with nq > 1 is the batched
    extension and returns ((nq, k) pids, (nq, k) scores)."""
    Qa = np.asarray(Q)
    if Qa.ndim == 2 or (Qa.ndim == 3 and Qa.shape[2] == 1):
        if Qa.shape[1] != searcher.config.query_maxlen:
            raise AssertionError(f"size(Q): {Qa.shape}, query_maxlen: {searcher.config.query_maxlen}")
    pids, scores, counts = searcher.search_batch(Qa, k)
    short = np.nonzero(counts < k)[0]
    if len(short):
        q = int(short[0])
        raise BoundsError(f"attempt to access {int(counts[q])}-element Vector{{Int64}} at index [1:{k}]"
                          + (f" (query {q + 1})" if pids.shape[0] > 1 else ""))
    if Qa.ndim == 2 or Qa.shape[2] == 1:
        return pids[0], scores[0]
    return pids, scores


def retrieve(ivf, ivf_lengths, centroids, emb2pid, nprobe, Q, device=0):
    """Signature-compatible `retrieve` (src/search/ranking.jl:23-25) for callers that hold bare
    arrays: builds a throw-away resident index (dim zero-padded to a multiple of 8 -- dot products
    are unchanged) and runs stages 1+2.  `emb2pid` must be non-decreasing (it always is:
    `_build_emb2pid`)."""
    ivf = np.asarray(ivf, dtype=np.int64)
    ivf_lengths = np.asarray(ivf_lengths, dtype=np.int64)
    emb2pid = np.asarray(emb2pid, dtype=np.int64)
    centroids = np.asarray(centroids, dtype=np.float32)
    Q = np.asarray(Q, dtype=np.float32)
    if len(ivf) != int(ivf_lengths.sum()):
        raise DimensionMismatch("length(ivf) must be equal to sum(ivf_lengths)!")
    if len(emb2pid) and np.any(np.diff(emb2pid) < 0):
        raise L.Unsupported("emb2pid must be non-decreasing")
    n_e = len(emb2pid)
    n_p = int(emb2pid.max()) if n_e else 0
    doclens = np.bincount(emb2pid, minlength=n_p + 1)[1:].astype(np.int64)
    codes = np.ones(n_e, dtype=np.uint32)
    codes[ivf - 1] = np.repeat(np.arange(1, len(ivf_lengths) + 1), ivf_lengths).astype(np.uint32)
    dim = centroids.shape[0]
    pdim = (dim + 7) // 8 * 8
    cen = np.zeros((pdim, centroids.shape[1]), dtype=np.float32)
    cen[:dim] = centroids
    Qp = np.zeros((pdim, Q.shape[1]), dtype=np.float32)
    Qp[:dim] = Q
    cfg = ColBERTConfig(dim=pdim, nbits=1, nprobe=nprobe, query_maxlen=Q.shape[1])
    res = np.zeros((pdim // 8, n_e), dtype=np.uint8)
    with Searcher(cfg, cen, None, np.zeros(2, np.float32), ivf, ivf_lengths, doclens, codes, res, device=device) as s:
        return s.retrieve(Qp, nprobe)


def decompress(dim, nbits, centroids, bucket_weights, codes, residuals, bsize=10000, device=0,
               return_bucket_indices=False, return_unnormalized=False):
    """`decompress` (src/indexing/codecs/residual.jl:759-784): centroids (dim, K), codes 1-based
    UInt32, residuals (dim/8*nbits, n) -> Float32 (dim, n).  `bsize` is accepted for signature
    compatibility (the GPU kernel does not batch)."""
    cen = _julia_to_c(centroids, np.float32)
    codes_c = _c(codes, np.uint32)
    res = _julia_to_c(residuals, np.uint8)
    w = _c(bucket_weights, np.float32)
    if dim % 8 != 0:
        raise DomainError("dim should be a multiple of 8!")
    if len(codes_c) != res.shape[0]:
        raise DomainError("The number of codes should be equal to the number of residual embeddings!")
    if res.shape[1] != dim // 8 * nbits:
        raise DomainError("The dimension each residual in binary_residuals should be (dim / 8) * nbits!")
    if len(w) != (1 << nbits):
        raise DomainError("bucket_weights should have length 2^nbits!")
    n = len(codes_c)
    out = np.zeros((n, dim), dtype=np.float32)
    idx = np.zeros((n, dim), dtype=np.uint8) if return_bucket_indices else None
    raw = np.zeros((n, dim), dtype=np.float32) if return_unnormalized else None
    L.check(L.load().cb_decompress(device, dim, nbits, cen.shape[0], _ptr(cen), _ptr(w), _ptr(codes_c), _ptr(res), n,
                                   _ptr(out), _ptr(idx), _ptr(raw)))
    ret = [out.T]
    if return_bucket_indices:
        ret.append(idx.T)
    if return_unnormalized:
        ret.append(raw.T)
    return ret[0] if len(ret) == 1 else tuple(ret)


def compress(centroids, bucket_cutoffs, dim, nbits, embs, bsize=10000, device=0):
    """`compress` (src/indexing/codecs/residual.jl:586-604): centroids (dim, K), embs (dim, n) ->
    (codes UInt32 (n,) 1-based, residuals UInt8 (dim/8*nbits, n)).  `bsize` accepted for signature compatibility."""
    if dim % 8 != 0:
        raise DomainError("dims should be a multiple of 8!")
    cut = _c(bucket_cutoffs, np.float32)
    if len(cut) != (1 << nbits) - 1:
        raise DomainError("length(bucket_cutoffs) should be 2^nbits - 1!")
    cen = _julia_to_c(centroids, np.float32)
    e = _julia_to_c(embs, np.float32)
    if cen.shape[1] != dim or (e.size and e.shape[1] != dim):
        raise DimensionMismatch("centroids / embs must have `dim` rows")
    n = e.shape[0]
    codes = np.zeros(n, dtype=np.uint32)
    res = np.zeros((n, dim // 8 * nbits), dtype=np.uint8)
    L.check(L.load().cb_compress(device, dim, nbits, cen.shape[0], _ptr(cen), _ptr(cut), _ptr(e), n, _ptr(codes), _ptr(res)))
    return codes, res.T


def maxsim(Q, D, pids, doclens, device=0):
    """`maxsim(Q, D, pids, doclens)` (src/search/ranking.jl:69-86): Q (dim, T), D (dim, M)."""
    Qc = _julia_to_c(Q, np.float32)
    Dc = _julia_to_c(D, np.float32)
    if Qc.shape[1] != Dc.shape[1] and Dc.shape[0] > 0:
        raise DimensionMismatch("Q and D must share their first dimension")
    pids = _c(pids, np.int64)
    dl = _c(doclens, np.int64)
    out = np.zeros(len(pids), dtype=np.float32)
    L.check(L.load().cb_maxsim(device, Qc.shape[1], Qc.shape[0], _ptr(Qc), _ptr(Dc), Dc.shape[0], _ptr(pids),
                               len(pids), _ptr(dl), len(dl), _ptr(out)))
    return out


_JLD2_DTYPES = {1: np.float32, 2: np.float64, 3: np.int8, 4: np.uint8, 5: np.int16, 6: np.uint16, 7: np.int32, 8: np.uint32,
                9: np.int64, 10: np.uint64}


def load_object(path: str, name: str = "single_stored_object"):
    """`JLD2.load_object(path)` for the plain numeric arrays / scalars the index is made of, through the library's
    native reader (host only).  Returns the numpy array in C layout: a Julia Matrix{T}(a, b) comes back with shape (b, a)."""
    lib = L.load()
    info = (C.c_int64 * 11)()
    L.check(lib.cb_jld2_read(path.encode(), name.encode(), info, None, 0))
    shape = tuple(info[3 + i] for i in range(info[2]))
    out = np.zeros(shape, dtype=_JLD2_DTYPES[info[0]])
    if out.nbytes:
        L.check(lib.cb_jld2_read(path.encode(), name.encode(), info, _ptr(out) if out.ndim else out.ctypes.data_as(C.c_void_p), out.nbytes))
    return out


def merge_topk(pids, scores, device=0):
    """Cross-shard stage 5: pids / scores (n_lists, nq, k) -> (nq, k) first-k by (score desc,
    pid asc)."""
    pids = _c(pids, np.int64)
    scores = _c(scores, np.float32)
    n_lists, nq, k = pids.shape
    op = np.zeros((nq, k), dtype=np.int64)
    os_ = np.zeros((nq, k), dtype=np.float32)
    L.check(L.load().cb_merge_topk(device, n_lists, nq, k, _ptr(pids), _ptr(scores), _ptr(op), _ptr(os_)))
    return op, os_

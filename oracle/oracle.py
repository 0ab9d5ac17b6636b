"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy (fp32, OpenBLAS) restatement of the search-time scoring path of
JuliaGenAI/ColBERT.jl.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module;
nothing under `colbert.jl_b200/` does (the product path has no CPU fallback).

Parity status: PINNED against every golden vector the reference's own tests
hold for this path (tests/test_oracle_golden.py lists them with file:line);
`decompress` numerics (centroid add + L2 normalise) and end-to-end `search`
are NOT pinned by the reference's tests and follow the
source only.  The reference itself (Julia) cannot run in this image.

Conventions: every function keeps the reference's *Julia* conventions so the
golden vectors apply literally:
  * matrices use the Julia shape, e.g. centroids (dim, K), residuals (R, N_e),
    Q (dim, T);  a Julia `Matrix{T}(a, b)` is handed to the C ABI as C `T[b][a]`
    which is simply `.T` of these arrays;
  * codes, eids (`ivf`), pids and centroid ids are 1-BASED integers.
All `file:line` citations are relative to /root/reference/.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
EPS32 = np.finfo(np.float32).eps  # Julia eps(Float32)


class DimensionMismatch(Exception):
    """Mirror of Julia's DimensionMismatch."""


class DomainError(Exception):
    """Mirror of Julia's DomainError."""


class BoundsError(IndexError):
    """Mirror of Julia's BoundsError."""


# --------------------------------------------------------------------------- utils
def _head(v):
    """src/utils.jl:334-336 -- all but the last element."""
    v = list(v) if not isinstance(v, np.ndarray) else v
    return v[:-1] if len(v) > 0 else v[:0]


def _offsets_1based(lengths):
    """`cumsum([1; _head(lengths)])` (ranking.jl:13-14,50-51,77; searching.jl:85)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    out = np.ones(len(lengths), dtype=np.int64)
    if len(lengths) > 1:
        out[1:] += np.cumsum(lengths[:-1])
    return out


def _normalize_array(X, dims=1):
    """src/utils.jl:320-325 -- X ./= (sqrt.(sum(abs2, X, dims)) .+ eps(T)); note
    eps is ADDED to the norm (not a clamp).  `dims` is the 1-based Julia dim."""
    X = np.asarray(X)
    norms = np.sqrt(np.sum(X * X, axis=dims - 1, keepdims=True, dtype=X.dtype))
    X /= (norms + np.finfo(X.dtype).eps)
    return X


def _topk(data, k, dims=1):
    """src/utils.jl:327-332 -- mapslices(v -> partialsortperm(v, 1:k, rev=true)).
    Returns 1-based indices, best first; ties resolve to the LOWER index (the
    `Perm` ordering used by partialsortperm compares indices on equal keys)."""
    if dims not in (1, 2):
        raise DomainError("dims must be 1 or 2!")
    data = np.asarray(data)

    def one(v):
        # stable sort of -v: equal keys keep ascending index order
        return np.argsort(-v, kind="stable")[:k] + 1

    if dims == 1:
        return np.stack([one(data[:, j]) for j in range(data.shape[1])], axis=1)
    return np.stack([one(data[i, :]) for i in range(data.shape[0])], axis=0)


def _topk_rows_fast(data, k):
    """Same result as `_topk(data, k, dims=2)` but O(n) per row via argpartition;
    falls back to the exact stable path for rows where a tie straddles rank k."""
    data = np.asarray(data)
    n = data.shape[1]
    if k >= n:
        return _topk(data, k, dims=2)
    part = np.argpartition(-data, k - 1, axis=1)[:, :k]
    out = np.empty((data.shape[0], k), dtype=np.int64)
    for i in range(data.shape[0]):
        row = data[i]
        idx = part[i]
        kth = row[idx].min()
        if np.count_nonzero(row == kth) > np.count_nonzero(row[idx] == kth) or \
                len(np.unique(row[idx])) < k:
            out[i] = np.argsort(-row, kind="stable")[:k] + 1
        else:
            out[i] = idx[np.argsort(-row[idx], kind="stable")] + 1
    return out


# --------------------------------------------------------------------------- searching.jl
def _build_emb2pid(doclens):
    """src/searching.jl:82-91 -- pid repeated doclens[pid] times (1-based pids)."""
    doclens = np.asarray(doclens, dtype=np.int64)
    return np.repeat(np.arange(1, len(doclens) + 1, dtype=np.int64), doclens)


# --------------------------------------------------------------------------- ranking.jl
def _cids_to_eids(eids, centroid_ids, ivf, ivf_lengths):
    """src/search/ranking.jl:7-21 (`_cids_to_eids!`) -- fills `eids` in place."""
    centroid_ids = np.asarray(centroid_ids, dtype=np.int64)
    ivf = np.asarray(ivf, dtype=np.int64)
    ivf_lengths = np.asarray(ivf_lengths, dtype=np.int64)
    sel = ivf_lengths[centroid_ids - 1] if len(centroid_ids) else np.zeros(0, np.int64)
    if len(eids) != int(sel.sum()):
        raise DimensionMismatch("length(eids) must be equal to sum(ivf_lengths[centroid_ids])!")
    if len(ivf) != int(ivf_lengths.sum()):
        raise DimensionMismatch("length(ivf) must be equal to sum(ivf_lengths)!")
    centroid_ivf_offsets = _offsets_1based(ivf_lengths)
    eid_offsets = _offsets_1based(sel)
    for idx, cid in enumerate(centroid_ids):
        eo = eid_offsets[idx] - 1
        bl = ivf_lengths[cid - 1]
        io = centroid_ivf_offsets[cid - 1] - 1
        eids[eo:eo + bl] = ivf[io:io + bl]
    return eids


def centroid_scores(Q, centroids):
    """`cells = Q' * centroids` (ranking.jl:27) -- fp32 sgemm (OpenBLAS here, the
    BLAS family Julia bundles).  Q (dim, T), centroids (dim, K) -> (T, K)."""
    return np.asarray(Q, dtype=F32).T @ np.asarray(centroids, dtype=F32)


def fixed_order_dot(A, B):
    """Adjudication path: fp32 dot products in a FIXED order (k = 0..dim-1,
    product rounded, then added: no FMA), bit-identical to the CUDA exact
    re-score kernel (`__fadd_rn(acc, __fmul_rn(q, c))`).  A (n, dim), B (n, dim)
    -> (n,) row-wise dots."""
    A = np.asarray(A, dtype=F32)
    B = np.asarray(B, dtype=F32)
    acc = np.zeros(A.shape[0], dtype=F32)
    for k in range(A.shape[1]):
        acc = (acc + (A[:, k] * B[:, k]).astype(F32)).astype(F32)
    return acc


def probe_cells(Q, centroids, nprobe, fast=True):
    """ranking.jl:27-32 -- sorted unique 1-based centroid ids probed by the query."""
    cells = centroid_scores(Q, centroids)
    top = _topk_rows_fast(cells, nprobe) if fast else _topk(cells, nprobe, dims=2)
    return np.unique(top.reshape(-1)), top, cells


def retrieve(ivf, ivf_lengths, centroids, emb2pid, nprobe, Q):
    """src/search/ranking.jl:23-44 -- candidate pids (sorted ascending, unique,
    1-based): every passage owning >= 1 embedding whose code is in the union of
    the per-token top-`nprobe` centroids.  No candidate cap is applied."""
    ivf_lengths = np.asarray(ivf_lengths, dtype=np.int64)
    centroid_ids, _, _ = probe_cells(Q, centroids, nprobe)
    eids = np.empty(int(ivf_lengths[centroid_ids - 1].sum()), dtype=np.int64)
    _cids_to_eids(eids, centroid_ids, ivf, ivf_lengths)
    eids = np.unique(eids)                       # sort(unique(eids))        :39
    pids = np.unique(np.asarray(emb2pid)[eids - 1])  # sort(unique(emb2pid[eids])) :42
    return pids.astype(np.int64)


def _collect_compressed_embs_for_pids(doclens, codes, residuals, pids):
    """src/search/ranking.jl:46-67 -- gathers candidates' codes and residual columns
    contiguously in `pids` order (pids may repeat / be unsorted).
    residuals has the Julia shape (R, N_e)."""
    doclens = np.asarray(doclens, dtype=np.int64)
    pids = np.asarray(pids, dtype=np.int64)
    pid_offsets = _offsets_1based(doclens)
    sel = doclens[pids - 1] if len(pids) else np.zeros(0, np.int64)
    offsets = _offsets_1based(sel)
    num_embeddings = int(sel.sum())
    codes_packed = np.zeros(num_embeddings, dtype=np.uint32)
    residuals_packed = np.zeros((residuals.shape[0], num_embeddings), dtype=np.uint8, order="F")
    for idx, pid in enumerate(pids):
        o = offsets[idx] - 1
        po = pid_offsets[pid - 1] - 1
        n = doclens[pid - 1]
        codes_packed[o:o + n] = codes[po:po + n]
        residuals_packed[:, o:o + n] = residuals[:, po:po + n]
    return codes_packed, residuals_packed


def _collect_compressed_embs_for_pids_fast(doclens, codes, residuals, pids, pid_offsets=None):
    """Vectorised twin of `_collect_compressed_embs_for_pids` (same outputs; one fancy-index
    gather instead of a per-pid loop) used when timing the CPU baseline, so that the Python loop
    overhead -- which Julia does not pay -- is not billed to the reference."""
    doclens = np.asarray(doclens, dtype=np.int64)
    pids = np.asarray(pids, dtype=np.int64)
    if pid_offsets is None:
        pid_offsets = _offsets_1based(doclens)
    sel = doclens[pids - 1] if len(pids) else np.zeros(0, np.int64)
    total = int(sel.sum())
    starts = np.repeat(pid_offsets[pids - 1] - 1 - (_offsets_1based(sel) - 1), sel)
    idx = starts + np.arange(total, dtype=np.int64)
    return codes[idx], np.asfortranarray(residuals[:, idx])


def maxsim(Q, D, pids, doclens):
    """src/search/ranking.jl:69-86 -- score[p] = sum_t max_{e in p} Q[:,t].D[:,e].
    Q (dim, T), D (dim, M) with M = sum(doclens[pids])."""
    doclens = np.asarray(doclens, dtype=np.int64)
    pids = np.asarray(pids, dtype=np.int64)
    sel = doclens[pids - 1] if len(pids) else np.zeros(0, np.int64)
    if int(sel.sum()) != D.shape[1]:
        raise DimensionMismatch("The total number of embeddings for pids does not match "
                                "with the dimension of D!")
    scores = np.zeros(len(pids), dtype=F32)
    query_doc_scores = np.asarray(Q, dtype=F32).T @ np.asarray(D, dtype=F32)  # (T, M)
    offsets = _offsets_1based(sel)
    for idx in range(len(pids)):
        n = sel[idx]
        o = offsets[idx] - 1
        pid_scores = query_doc_scores[:, o:o + n]
        scores[idx] = np.sum(np.max(pid_scores, axis=1), dtype=F32)
    return scores


def maxsim_fast(Q, D, pids, doclens):
    """Vectorised twin of `maxsim` (np.maximum.reduceat) for large candidate sets;
    identical up to fp32 summation order of the final 32-term sum.  Requires every
    selected doclen >= 1."""
    doclens = np.asarray(doclens, dtype=np.int64)
    pids = np.asarray(pids, dtype=np.int64)
    sel = doclens[pids - 1]
    if int(sel.sum()) != D.shape[1]:
        raise DimensionMismatch("D does not match sum(doclens[pids])")
    if len(pids) == 0:
        return np.zeros(0, dtype=F32)
    assert sel.min() >= 1
    qd = np.asarray(Q, dtype=F32).T @ np.asarray(D, dtype=F32)
    starts = (_offsets_1based(sel) - 1)
    mx = np.maximum.reduceat(qd, starts, axis=1)  # (T, n_pids)
    return np.sum(mx, axis=0, dtype=F32)


# --------------------------------------------------------------------------- residual.jl (codec)
def _binarize(data, nbits):
    """src/indexing/codecs/residual.jl:197-208 -- (dim, b) ints -> Bool (nbits, dim, b),
    bit 0 = LSB."""
    data = np.asarray(data)
    if data.size and (data.min() < 0 or data.max() > (1 << nbits) - 1):
        raise DomainError("All values in the matrix should be in range [0, 2^nbits - 1]!")
    data = data.astype(np.int64)
    pos = np.arange(nbits, dtype=np.int64).reshape(nbits, 1, 1)
    return ((data[None, :, :] >> pos) & 1).astype(bool)


def _unbinarize(data):
    """src/indexing/codecs/residual.jl:233-240 -- Bool (nbits, dim, b) -> Int (dim, b)."""
    data = np.asarray(data, dtype=bool)
    nbits = data.shape[0]
    pos = (np.int64(1) << np.arange(nbits, dtype=np.int64)).reshape(nbits, 1, 1)
    return np.sum(data.astype(np.int64) * pos, axis=0)


def _bucket_indices(data, bucket_cutoffs):
    """src/indexing/codecs/residual.jl:348-351 -- searchsortedfirst(cutoffs, x) - 1
    == number of cutoffs strictly less than x == np.searchsorted(side='left')."""
    data = np.asarray(data)
    bucket_cutoffs = np.asarray(bucket_cutoffs)
    return np.searchsorted(bucket_cutoffs, data, side="left").astype(np.int64).reshape(data.shape)


def _packbits(bitsarray):
    """src/indexing/codecs/residual.jl:400-407 -- Bool (nbits, dim, b) -> UInt8
    (dim/8*nbits, b).  Flat (column-major) bit stream, LSB-first inside each byte
    (BitArray.chunks semantics)."""
    bitsarray = np.asarray(bitsarray, dtype=bool)
    nbits, dim, batch = bitsarray.shape
    if dim % 8 != 0:
        raise DomainError("dim should be a multiple of 8!")
    flat = bitsarray.flatten(order="F")
    packed = np.packbits(flat, bitorder="little")[: flat.size >> 3]
    return packed.reshape(((dim >> 3) * nbits, batch), order="F")


def _unpackbits(packedbits, nbits):
    """src/indexing/codecs/residual.jl:428-441 -- UInt8 (R, b) -> Bool (nbits, dim, b)."""
    packedbits = np.asarray(packedbits, dtype=np.uint8)
    if packedbits.shape[0] % nbits != 0:
        raise DomainError("The first dimension of packbits should be a multiple of nbits!")
    batch = packedbits.shape[1]
    dim = (packedbits.shape[0] // nbits) << 3
    bits = np.unpackbits(packedbits.flatten(order="F"), bitorder="little")
    return bits.astype(bool).reshape((nbits, dim, batch), order="F")


def binarize(dim, nbits, bucket_cutoffs, residuals):
    """src/indexing/codecs/residual.jl:518-536."""
    if dim % 8 != 0:
        raise DomainError("dims should be a multiple of 8!")
    if len(bucket_cutoffs) != (1 << nbits) - 1:
        raise DomainError("length(bucket_cutoffs) should be 2^nbits - 1!")
    idx = _bucket_indices(residuals, bucket_cutoffs)
    return _packbits(_binarize(idx, nbits))


def unpack_bucket_indices(dim, nbits, binary_residuals):
    """0-based bucket index per (dim, b): `_unbinarize(_unpackbits(...))`
    (residual.jl:709-710) without the +1."""
    return _unbinarize(_unpackbits(binary_residuals, nbits))


def unpack_bucket_indices_fast(dim, nbits, binary_residuals):
    """Closed form of `unpack_bucket_indices` for nbits in {1, 2, 4, 8} (dimensions never
    straddle bytes): idx(d) = (byte[(d*nbits) >> 3] >> ((d*nbits) & 7)) & (2^nbits - 1).
    Pinned equal to the generic bit-stream path by tests/test_oracle_golden.py.  Returns uint8."""
    assert nbits in (1, 2, 4, 8)
    per = 8 // nbits
    b = np.asarray(binary_residuals, dtype=np.uint8)
    out = np.empty((dim, b.shape[1]), dtype=np.uint8)
    mask = np.uint8((1 << nbits) - 1)
    for j in range(per):
        out[j::per, :] = (b >> np.uint8(j * nbits)) & mask
    return out


def decompress_residuals(dim, nbits, bucket_weights, binary_residuals):
    """src/indexing/codecs/residual.jl:698-721."""
    binary_residuals = np.asarray(binary_residuals, dtype=np.uint8)
    if dim % 8 != 0:
        raise DomainError("dim should be a multiple of 8!")
    if binary_residuals.shape[0] != (dim // 8) * nbits:
        raise DomainError("The dimension each residual in binary_residuals should be "
                          "(dim / 8) * nbits!")
    if len(bucket_weights) != (1 << nbits):
        raise DomainError("bucket_weights should have length 2^nbits!")
    idx = unpack_bucket_indices(dim, nbits, binary_residuals) + 1
    if idx.size and (idx.min() < 1 or idx.max() > len(bucket_weights)):
        raise BoundsError("unpacked indices out of range")
    return np.asarray(bucket_weights, dtype=F32)[idx - 1]


def decompress(dim, nbits, centroids, bucket_weights, codes, residuals, bsize=10000,
               return_unnormalized=False, fast=False):
    """src/indexing/codecs/residual.jl:759-784 -- v = centroids[:, code] + w[bucket];
    v ./= (||v||_2 + eps).  Returns Float32 (dim, M)."""
    codes = np.asarray(codes)
    if len(codes) != residuals.shape[1]:
        raise DomainError("The number of codes should be equal to the number of residual "
                          "embeddings!")
    K = centroids.shape[1]
    if len(codes) and (codes.min() < 1 or codes.max() > K):
        raise DomainError("All the codes must be in the valid range of centroid IDs!")
    embeddings = np.empty((dim, len(codes)), dtype=F32, order="F")
    raw = np.empty((dim, len(codes)), dtype=F32, order="F") if return_unnormalized else None
    centroids = np.asarray(centroids, dtype=F32)
    for o in range(0, len(codes), bsize):
        e = min(len(codes), o + bsize)
        bc = codes[o:e].astype(np.int64) - 1
        if fast and nbits in (1, 2, 4, 8):
            res = np.asarray(bucket_weights, dtype=F32)[unpack_bucket_indices_fast(dim, nbits, residuals[:, o:e])]
        else:
            res = decompress_residuals(dim, nbits, bucket_weights, residuals[:, o:e])
        batch = (centroids[:, bc] + res).astype(F32)
        if raw is not None:
            raw[:, o:e] = batch
        embeddings[:, o:e] = _normalize_array(batch, dims=1)
    if return_unnormalized:
        return embeddings, raw
    return embeddings


def compress_into_codes(codes, centroids, embs, bsize=1000):
    """src/indexing/codecs/residual.jl:67-81 (`compress_into_codes!`) -- fixture side."""
    n = embs.shape[1]
    if len(codes) != n:
        raise DimensionMismatch("length(codes) must be equal to the number of embeddings!")
    for o in range(0, n, bsize):
        e = min(n, o + bsize)
        dots = embs[:, o:e].T @ centroids
        codes[o:e] = np.argmax(dots, axis=1) + 1   # first max wins, like Julia argmax
    return codes


def compress(centroids, bucket_cutoffs, dim, nbits, embs, bsize=10000):
    """src/indexing/codecs/residual.jl:586-604 -- fixture side (defines the bit layout)."""
    n = embs.shape[1]
    codes = np.zeros(n, dtype=np.uint32)
    out = np.empty(((dim // 8) * nbits, n), dtype=np.uint8, order="F")
    for o in range(0, n, bsize):
        e = min(n, o + bsize)
        compress_into_codes(codes[o:e], centroids, embs[:, o:e])
        res = embs[:, o:e] - centroids[:, codes[o:e].astype(np.int64) - 1]
        out[:, o:e] = binarize(dim, nbits, bucket_cutoffs, res)
    return codes, out


# --------------------------------------------------------------------------- collection_indexer.jl
def _build_ivf(codes, num_partitions):
    """src/indexing/collection_indexer.jl:349-353 -- ivf = sortperm(codes) (stable,
    1-based eids grouped by centroid, ascending eid inside a cell)."""
    codes = np.asarray(codes)
    ivf = np.argsort(codes, kind="stable").astype(np.int64) + 1
    ivf_lengths = np.bincount(codes.astype(np.int64), minlength=num_partitions + 1)[1:num_partitions + 1]
    return ivf, ivf_lengths.astype(np.int64)


def _bucket_cutoffs_and_weights(nbits, heldout_avg_residual):
    """src/indexing/collection_indexer.jl:141-152 (Statistics.quantile, type-7 = numpy
    default 'linear')."""
    num_options = 1 << nbits
    quantiles = np.arange(num_options) / num_options
    cq, wq = quantiles[1:], quantiles + 0.5 / num_options
    flat = np.asarray(heldout_avg_residual).reshape(-1)
    return (np.quantile(flat, cq).astype(F32), np.quantile(flat, wq).astype(F32))


# --------------------------------------------------------------------------- search (minus the encoder)
class Index:
    """The host arrays `struct Searcher` holds (src/searching.jl:1-16), Julia shapes,
    1-based values."""

    def __init__(self, dim, nbits, centroids, bucket_weights, ivf, ivf_lengths, doclens, codes,
                 residuals, nprobe=2):
        self.dim, self.nbits, self.nprobe = dim, nbits, nprobe
        self.centroids = centroids            # (dim, K) f32
        self.bucket_weights = bucket_weights  # (2^nbits,) f32
        self.ivf = ivf                        # (N_e,) int64, 1-based eids
        self.ivf_lengths = ivf_lengths        # (K,) int64
        self.doclens = doclens                # (N_p,) int64
        self.codes = codes                    # (N_e,) uint32, 1-based
        self.residuals = residuals            # (R, N_e) uint8
        self.emb2pid = _build_emb2pid(doclens)


def search_all_scores(index: Index, Q, fast=True):
    """src/searching.jl:103-122 for one query matrix Q (dim, T): returns
    (pids ascending, scores in pid order)."""
    pids = retrieve(index.ivf, index.ivf_lengths, index.centroids, index.emb2pid, index.nprobe, Q)
    if fast:
        codes_packed, residuals_packed = _collect_compressed_embs_for_pids_fast(
            index.doclens, index.codes, index.residuals, pids)
    else:
        codes_packed, residuals_packed = _collect_compressed_embs_for_pids(
            index.doclens, index.codes, index.residuals, pids)
    D = decompress(index.dim, index.nbits, index.centroids, index.bucket_weights,
                   codes_packed, residuals_packed, fast=fast)
    if fast and len(pids) and index.doclens[pids - 1].min() >= 1:
        scores = maxsim_fast(Q, D, pids, index.doclens)
    else:
        scores = maxsim(Q, D, pids, index.doclens)
    return pids, scores


def search(index: Index, Q, k, fast=True):
    """src/searching.jl:93-128 minus `encode_queries`: stable descending sort of the
    scores (ties keep ascending-pid order), first k.  Raises BoundsError when fewer
    than k candidates exist (searching.jl:127)."""
    pids, scores = search_all_scores(index, Q, fast=fast)
    indices = np.argsort(-scores, kind="stable")   # sortperm(scores, rev=true)
    if k > len(pids):
        raise BoundsError(f"attempt to access {len(pids)}-element Vector at index [1:{k}]")
    return pids[indices][:k], scores[indices][:k]


# ---------------------------------------------------------------------------------------------
# PLAID-style pruned search (BASELINE.json config 5; SURVEY.md section 8c "PLAID knobs", 8f rank f3)
#
# The reference has NO implementation of this (README.md:187 lists it as roadmap), so parity is
# UNPINNED: the semantics below are the ones this oracle defines, following the PLAID paper
# (Santhanam et al. 2022, stages 1-3) on top of the reference's own retrieve / decompress / maxsim:
#   1. candidates   = retrieve(...) with nprobe = ncells                      (ranking.jl:23-44)
#   2. centroid pruning: centroid c survives for the query iff max_t S[t, c] >= threshold, with
#      S[t, c] the fixed-order fp32 dot of query token t and centroid c (`fixed_order_dot`)
#   3. approximate score of a candidate = sum over query tokens t of
#      max(0, max over the passage's tokens e whose code survives of S[t, code_e]);
#      the sum over the 32 (zero-padded) token maxima is the pairwise tree a warp butterfly forms
#      (`tree_sum32`), so the value is reproducible bit for bit
#   4. keep the first `ndocs` candidates under (approximate score desc, pid asc)
#   5. exact MaxSim (decompress + maxsim) on those, stable descending sort, first k.
# ---------------------------------------------------------------------------------------------
def fixed_order_scores(Q, centroids):
    """S (T, K): S[t, c] = fixed_order_dot(Q[:, t], centroids[:, c])."""
    Qt = np.ascontiguousarray(np.asarray(Q, dtype=F32).T)          # (T, dim)
    Ct = np.ascontiguousarray(np.asarray(centroids, dtype=F32).T)  # (K, dim)
    acc = np.zeros((Qt.shape[0], Ct.shape[0]), dtype=F32)
    for k in range(Qt.shape[1]):
        acc = (acc + (Qt[:, k][:, None] * Ct[:, k][None, :]).astype(F32)).astype(F32)
    return acc


def tree_sum32(v):
    """Sum of 32 fp32 values (last axis, zero-padded to 32) in the order of a warp xor-butterfly:
    a[i] += a[i ^ 16], then ^8, ^4, ^2, ^1; the result is a[0]."""
    v = np.asarray(v, dtype=F32)
    if v.shape[-1] < 32:
        pad = np.zeros(v.shape[:-1] + (32 - v.shape[-1],), dtype=F32)
        v = np.concatenate([v, pad], axis=-1)
    assert v.shape[-1] == 32
    idx = np.arange(32)
    a = v.copy()
    for o in (16, 8, 4, 2, 1):
        a = (a + a[..., idx ^ o]).astype(F32)
    return a[..., 0]


def plaid_approx_scores(index: "Index", Q, ncells, centroid_score_threshold, vectorized=True):
    """Steps 1-3: (candidate pids ascending, approximate scores in pid order, surviving centroid
    ids 1-based ascending)."""
    pids = retrieve(index.ivf, index.ivf_lengths, index.centroids, index.emb2pid, ncells, Q)
    S = fixed_order_scores(Q, index.centroids)                       # (T, K)
    keep = S.max(axis=0) >= F32(centroid_score_threshold)            # (K,)
    Sp = np.where(keep[None, :], np.maximum(S, F32(0)), F32(0)).astype(F32)
    doclens = np.asarray(index.doclens, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(doclens)])
    approx = np.zeros(len(pids), dtype=F32)
    codes0 = np.asarray(index.codes, dtype=np.int64) - 1
    if not vectorized:                                               # the definition, one candidate at a time
        for i, pid in enumerate(pids):
            c = codes0[off[pid - 1]:off[pid]]
            if len(c) == 0:
                continue
            m = Sp[:, c].max(axis=1)                                 # (T,) >= 0
            approx[i] = tree_sum32(m)
        return pids, approx, np.nonzero(keep)[0] + 1
    # same arithmetic, a few thousand candidates at a time (max is exact, the sum order is tree_sum32's)
    lens = doclens[pids - 1]
    SpT = np.ascontiguousarray(Sp.T)                                 # (K, T): one row per centroid
    for a in range(0, len(pids), 4096):
        sl = slice(a, min(a + 4096, len(pids)))
        ln = lens[sl]
        nz = ln > 0
        if not nz.any():
            continue
        starts = off[pids[sl] - 1]
        tok = np.concatenate([np.arange(s0, s0 + n) for s0, n in zip(starts[nz], ln[nz])])
        seg = np.concatenate([[0], np.cumsum(ln[nz])[:-1]])
        m = np.maximum.reduceat(SpT[codes0[tok]], seg, axis=0)       # (n_nonempty, T)
        out = np.zeros(len(ln), dtype=F32)
        out[nz] = tree_sum32(m)
        approx[sl] = out
    return pids, approx, np.nonzero(keep)[0] + 1


def plaid_search(index: "Index", Q, k, ncells, centroid_score_threshold, ndocs, return_selected=False):
    """Steps 1-5.  Returns (pids[:k'], scores[:k']) with k' = min(k, #selected) -- the batched
    entry point truncates instead of throwing (as `cb_search_batch` does)."""
    pids, approx, _ = plaid_approx_scores(index, Q, ncells, centroid_score_threshold)
    order = np.lexsort((pids, -approx.astype(np.float64)))           # approx desc, then pid asc
    sel = np.sort(pids[order[:ndocs]])
    codes_packed, residuals_packed = _collect_compressed_embs_for_pids_fast(index.doclens, index.codes, index.residuals, sel)
    D = decompress(index.dim, index.nbits, index.centroids, index.bucket_weights, codes_packed, residuals_packed, fast=True)
    scores = maxsim(Q, D, sel, index.doclens) if len(sel) else np.zeros(0, dtype=F32)
    indices = np.argsort(-scores, kind="stable")
    kk = min(k, len(sel))
    out = (sel[indices][:kk], scores[indices][:kk])
    return out + (sel, pids, approx) if return_selected else out


def merge_topk(all_pids, all_scores, k):
    """Checker for the cross-shard merge (no reference counterpart: the reference is single
    device).  all_pids / all_scores (n_lists, nq, k'), empty slots pid 0 / -inf.  Every passage
    lives in exactly one shard and shard pid ranges ascend, so the first k of the stable
    descending sort over the concatenation of ALL candidates (searching.jl:125-127) is the first
    k of the per-shard first-k lists under (score desc, pid asc)."""
    all_pids = np.asarray(all_pids, dtype=np.int64)
    all_scores = np.asarray(all_scores, dtype=F32)
    n, nq, _ = all_pids.shape
    out_p = np.zeros((nq, k), dtype=np.int64)
    out_s = np.full((nq, k), -np.inf, dtype=F32)
    for q in range(nq):
        p = all_pids[:, q, :].reshape(-1)
        s = all_scores[:, q, :].reshape(-1)
        keep = p > 0
        p, s = p[keep], s[keep]
        order = np.lexsort((p, -s.astype(np.float64)))   # score desc, then pid asc
        m = min(k, len(order))
        out_p[q, :m] = p[order[:m]]
        out_s[q, :m] = s[order[:m]]
    return out_p, out_s

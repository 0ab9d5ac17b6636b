"""Deterministic synthetic indexes and queries of the shapes BASELINE.json names.
Host (numpy) generator for tests and small runs; `device_index` builds the same distribution
directly in HBM with torch for the full-size benchmark (no multi-GB host staging).
Arrays come out in the C layouts of include/colbert_b200.h; `.T` gives the Julia shapes."""
from __future__ import annotations

import numpy as np

BUCKET_WEIGHTS = {
    1: np.array([-0.02, 0.02], dtype=np.float32),
    # nbits = 2: the weights shown in the reference's README (README.md:100)
    2: np.array([-0.041035336, -0.009812315, 0.008938393, 0.039779153], dtype=np.float32),
}


def bucket_weights(nbits: int) -> np.ndarray:
    if nbits in BUCKET_WEIGHTS:
        return BUCKET_WEIGHTS[nbits].copy()
    # mid-quantiles of N(0, 0.03^2) (what `_bucket_cutoffs_and_weights`,
    # src/indexing/collection_indexer.jl:141-152, yields on Gaussian residuals)
    from statistics import NormalDist
    n = 1 << nbits
    nd = NormalDist(0.0, 0.03)
    return np.array([nd.inv_cdf((i + 0.5) / n) for i in range(n)], dtype=np.float32)


def build_ivf(codes: np.ndarray, K: int):
    """Host mirror of `_build_ivf` (src/indexing/collection_indexer.jl:349-353): ivf =
    sortperm(codes) (stable; 1-based eids), ivf_lengths = counts per centroid."""
    ivf = np.argsort(codes, kind="stable").astype(np.int64) + 1
    lens = np.bincount(codes.astype(np.int64), minlength=K + 1)[1:K + 1].astype(np.int64)
    return ivf, lens


def make_index(n_passages: int, K: int, dim: int = 128, nbits: int = 2, doclen_mean: float = 120.0,
               doclen_std: float = 40.0, doclen_min: int = 8, doclen_max: int = 300, seed: int = 1000,
               profile: str = "uniform"):
    """Returns a dict of host arrays (C layouts): centroids [K][dim] f32 (rows L2-normalised),
    bucket_weights, doclens int64, codes uint32 (1-based), residuals uint8 [N_e][R], ivf, ivf_lengths."""
    g = lambda s: np.random.Generator(np.random.PCG64(seed + s))
    cen = g(1).standard_normal((K, dim), dtype=np.float32)
    cen /= np.linalg.norm(cen, axis=1, keepdims=True)
    doclens = np.clip(np.rint(g(2).normal(doclen_mean, doclen_std, n_passages)), doclen_min, doclen_max).astype(np.int64)
    n_e = int(doclens.sum())
    if profile == "uniform":
        codes = g(3).integers(1, K + 1, n_e, dtype=np.uint32)
    elif profile == "clustered":  # each passage draws 8 home centroids, 85 % of its tokens from them
        r = g(3)
        home = r.integers(1, K + 1, (n_passages, 8), dtype=np.uint32)
        pid_of = np.repeat(np.arange(n_passages), doclens)
        pick = home[pid_of, r.integers(0, 8, n_e)]
        rand = r.integers(1, K + 1, n_e, dtype=np.uint32)
        codes = np.where(r.random(n_e) < 0.85, pick, rand).astype(np.uint32)
    else:
        raise ValueError(profile)
    R = dim // 8 * nbits
    residuals = g(4).integers(0, 256, (n_e, R), dtype=np.uint8)
    ivf, ivf_lengths = build_ivf(codes, K)
    return dict(dim=dim, nbits=nbits, K=K, n_passages=n_passages, n_embeddings=n_e, centroids=cen,
                bucket_weights=bucket_weights(nbits), doclens=doclens, codes=codes, residuals=residuals,
                ivf=ivf, ivf_lengths=ivf_lengths)


def make_queries(centroids: np.ndarray, nq: int, T: int = 32, seed: int = 2001, nprobe: int = 2,
                 min_gap: float = 1e-4, noise: float = 0.5):
    """Queries [nq][T][dim]: token = normalise(centroid[c] + noise * g / sqrt(dim)), c uniform,
    g ~ N(0, I).  Rows whose nprobe-th / (nprobe+1)-th centroid-score gap is below `min_gap` are
    regenerated, so the probed cell set is well defined under any fp32 summation order."""
    K, dim = centroids.shape
    rng = np.random.Generator(np.random.PCG64(seed))
    Q = np.empty((nq * T, dim), dtype=np.float32)
    todo = np.arange(nq * T)
    for _ in range(20):
        if len(todo) == 0:
            break
        c = rng.integers(0, K, len(todo))
        v = centroids[c] + (noise / np.sqrt(dim)) * rng.standard_normal((len(todo), dim), dtype=np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        Q[todo] = v.astype(np.float32)
        if K <= nprobe:
            todo = todo[:0]
            break
        bad = []
        for s in range(0, len(todo), 2048):
            rows = todo[s:s + 2048]
            sc = Q[rows] @ centroids.T
            part = -np.partition(-sc, nprobe, axis=1)[:, :nprobe + 1]
            part.sort(axis=1)
            gap = part[:, 1] - part[:, 0]          # nprobe-th minus (nprobe+1)-th best
            bad.append(rows[gap < min_gap])
        todo = np.concatenate(bad) if bad else todo[:0]
    if len(todo):
        raise RuntimeError("could not generate well-separated queries")
    return Q.reshape(nq, T, dim)


def shard_ranges(doclens: np.ndarray, n_shards: int):
    """Contiguous passage ranges balanced by embedding count: returns n_shards + 1
    passage boundaries."""
    csum = np.concatenate([[0], np.cumsum(np.asarray(doclens, dtype=np.int64))])
    targets = csum[-1] * np.arange(1, n_shards) / n_shards
    cuts = np.searchsorted(csum, targets, side="left")
    return np.concatenate([[0], cuts, [len(doclens)]]).astype(np.int64)


def take_shard(index: dict, lo: int, hi: int):
    """The passage range [lo, hi) of a host index as a standalone index (rank-local IVF =
    `_build_ivf` of the local codes: ascending local eids per cell)."""
    doclens = index["doclens"]
    csum = np.concatenate([[0], np.cumsum(doclens)])
    e0, e1 = int(csum[lo]), int(csum[hi])
    codes = index["codes"][e0:e1]
    ivf, ivf_lengths = build_ivf(codes, index["K"])
    out = dict(index)
    out.update(n_passages=hi - lo, n_embeddings=e1 - e0, doclens=doclens[lo:hi], codes=codes,
               residuals=index["residuals"][e0:e1], ivf=ivf, ivf_lengths=ivf_lengths, pid_base=lo)
    return out

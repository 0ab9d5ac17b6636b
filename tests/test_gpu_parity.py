"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs (sizes the oracle finishes in seconds), against the reference's golden vectors,
and -- at larger sizes -- through size-independent properties.

Bars (BASELINE.json north_star): candidate pid sets and unpacked codes BIT-EXACT; MaxSim scores
within 1e-3 relative (fp32); top-k order identical except ties inside that tolerance.
"""
import numpy as np
import pytest

import colbert_jl_b200 as cb
from colbert_jl_b200 import synthetic as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-3   # north_star tolerance for MaxSim scores
rng = np.random.default_rng(7)


def make_searcher(ix, nprobe=2, T=32, **kw):
    cfg = cb.ColBERTConfig(dim=ix["dim"], nbits=ix["nbits"], nprobe=nprobe, query_maxlen=T)
    return cb.Searcher(cfg, ix["centroids"].T, None, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"],
                       ix["doclens"], ix["codes"], ix["residuals"].T, pid_base=ix.get("pid_base", 0), **kw)


def oracle_index(ix, nprobe=2):
    return O.Index(ix["dim"], ix["nbits"], ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"],
                   ix["doclens"], ix["codes"], ix["residuals"].T, nprobe=nprobe)


def check_topk(pids, scores, o_pids, o_scores, all_pids, all_scores, k):
    """top-k order identical except ties inside the tolerance."""
    np.testing.assert_allclose(scores, o_scores, rtol=SCORE_RTOL, atol=1e-5)
    if np.array_equal(pids, o_pids):
        return
    lookup = dict(zip(all_pids.tolist(), all_scores.tolist()))
    kth = o_scores[k - 1]
    for p, s, op in zip(pids, scores, o_pids):
        if p != op:  # a swap is legal only between near-tied scores
            assert p in lookup
            assert abs(lookup[int(p)] - lookup[int(op)]) <= SCORE_RTOL * max(1.0, abs(kth)), (p, op)


# --------------------------------------------------------------------------------------------
# golden vectors of the reference's own tests, through the ABI
# --------------------------------------------------------------------------------------------
def test_retrieve_golden():  # test/search/ranking.jl:71-83
    pids = cb.retrieve([3, 1, 4, 5, 6, 2], [2, 3, 1], np.array([[1.0, 0, 0], [0, 0, 1.0]], np.float32),
                       [10, 20, 30, 40, 50, 60], 2, np.array([[0.5, 0.5]], np.float32).T)
    assert pids.tolist() == [10, 20, 30]


def test_maxsim_golden():  # test/search/ranking.jl:137-152
    Q = np.array([[1.0, 0.5], [0.5, 1.0]], np.float32)
    D = np.array([[0.8, 0.3, 0.1], [0.2, 0.7, 0.4]], np.float32)
    assert cb.maxsim(Q, D, [1, 2], [1, 2]).tolist() == [1.5, 1.5]
    with pytest.raises(cb.DimensionMismatch):
        cb.maxsim(Q, D[:, :2], [1, 2], [1, 2])


def test_maxsim_shapes_vs_oracle():  # test/search/ranking.jl:154-161
    doclens = rng.integers(1, 11, 1000)
    Q = rng.random((128, 100), dtype=np.float32)
    D = rng.random((128, int(doclens.sum())), dtype=np.float32)
    pids = np.arange(1, 1001)
    s = cb.maxsim(Q, D, pids, doclens)
    assert s.dtype == np.float32 and s.shape == (1000,)
    np.testing.assert_allclose(s, O.maxsim(Q, D, pids, doclens), rtol=1e-5)


def test_build_emb2pid_golden():  # test/searching.jl:10-17 (host helper of the mirror)
    assert cb._build_emb2pid([3, 2, 4]).tolist() == [1, 1, 1, 2, 2, 3, 3, 3, 3]
    assert cb._build_emb2pid([0, 2, 0, 3]).tolist() == [2, 2, 4, 4, 4]


def test_unpackbits_golden_through_decompress():  # test/indexing/codecs/residual.jl:277-816
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "unpackbits_test1.json")))
    # 64 embeddings of dim 8, nbits 1: bucket index d of embedding e is bit (e*8 + d) of the stream
    packed = np.array(g["packed"], dtype=np.uint8).reshape(1, 64)
    w = np.array([0.0, 1.0], np.float32)
    cen = np.zeros((8, 1), np.float32)
    _, idx = cb.decompress(8, 1, cen, w, np.ones(64, np.uint32), packed, return_bucket_indices=True)
    assert idx.flatten(order="F").astype(int).tolist() == g["expected_bits"]


@pytest.mark.parametrize("nbits", [1, 2, 3, 4, 5, 8])
def test_decompress_residuals_inverts_binarize(nbits):  # test/indexing/codecs/residual.jl:975-991
    dim = 8 * int(rng.integers(1, 21))
    cut = np.sort(rng.random((1 << nbits) - 1, dtype=np.float32))
    w = np.sort(rng.random(1 << nbits, dtype=np.float32))
    res = rng.random((dim, int(rng.integers(1, 100))), dtype=np.float32)
    expected = w[np.searchsorted(cut, res, side="left")]
    packed = O.binarize(dim, nbits, cut, res)
    # zero centroid: the un-normalised output is exactly bucket_weights[idx]
    _, raw = cb.decompress(dim, nbits, np.zeros((dim, 1), np.float32), w, np.ones(res.shape[1], np.uint32), packed,
                           return_unnormalized=True)
    assert np.array_equal(raw, expected)


@pytest.mark.parametrize("nbits", [1, 2, 4, 6])
def test_decompress_vs_oracle(nbits):  # residual.jl:759-784 (numerics unpinned upstream: source-defined)
    dim, K, n = 128, 300, 5000
    cen = O._normalize_array(rng.standard_normal((dim, K)).astype(np.float32))
    w = S.bucket_weights(nbits)
    codes = rng.integers(1, K + 1, n).astype(np.uint32)
    res = rng.integers(0, 256, (dim // 8 * nbits, n), dtype=np.uint8)
    emb, idx, raw = cb.decompress(dim, nbits, cen, w, codes, res, return_bucket_indices=True, return_unnormalized=True)
    o_emb, o_raw = O.decompress(dim, nbits, cen, w, codes, res, return_unnormalized=True)
    assert np.array_equal(idx, O.unpack_bucket_indices(dim, nbits, res))      # unpacked codes: bit-exact
    assert np.array_equal(raw, o_raw)                                          # c + w[b]: bit-exact
    np.testing.assert_allclose(emb, o_emb, rtol=2e-6, atol=1e-8)
    assert emb.shape == (dim, n) and emb.dtype == np.float32


def test_decompress_errors():  # residual.jl:763-768
    dim, K = 16, 4
    cen = rng.random((dim, K), dtype=np.float32)
    w = np.zeros(4, np.float32)
    res = np.zeros((4, 3), np.uint8)
    with pytest.raises(cb.DomainError):
        cb.decompress(dim, 2, cen, w, np.array([1, 2], np.uint32), res)
    with pytest.raises(cb.DomainError):
        cb.decompress(dim, 2, cen, w, np.array([1, 5, 2], np.uint32), res)
    with pytest.raises(cb.DomainError):
        cb.decompress(dim, 2, cen, w, np.array([0, 1, 2], np.uint32), res)


def test_index_create_validation():
    ix = S.make_index(50, 16, dim=16, doclen_mean=5, doclen_std=2, doclen_min=1, doclen_max=9)
    bad = dict(ix)
    bad["codes"] = ix["codes"].copy()
    bad["codes"][3] = 17
    with pytest.raises(cb.DomainError):
        make_searcher(bad)
    bad = dict(ix)
    bad["doclens"] = ix["doclens"].copy()
    bad["doclens"][0] += 1
    with pytest.raises(cb.DimensionMismatch):
        make_searcher(bad)
    bad = dict(ix)
    bad["ivf_lengths"] = ix["ivf_lengths"].copy()
    bad["ivf_lengths"][0] += 1
    with pytest.raises(cb.DimensionMismatch):
        make_searcher(bad)


# --------------------------------------------------------------------------------------------
# stage-by-stage and end-to-end parity on seeded synthetic indexes
# --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tiny():
    ix = S.make_index(2000, 512, seed=11)
    Q = S.make_queries(ix["centroids"], 24, seed=12)
    return ix, Q, oracle_index(ix)


def test_probe_matches_oracle(tiny):
    ix, Q, oix = tiny
    with make_searcher(ix) as s:
        for nprobe in (1, 2, 4):
            Qn = S.make_queries(ix["centroids"], 8, seed=40 + nprobe, nprobe=nprobe)
            cells, scores = s.probe(np.transpose(Qn, (2, 1, 0)), nprobe=nprobe)
            for q in range(Qn.shape[0]):
                top = O._topk(O.centroid_scores(Qn[q].T, oix.centroids), nprobe, dims=2)
                assert np.array_equal(cells[q], top)            # same cells, same order
                fix = O.fixed_order_dot(np.repeat(Qn[q], nprobe, axis=0), ix["centroids"][top.reshape(-1) - 1])
                assert np.array_equal(scores[q].reshape(-1), fix)   # bit-exact fixed-order fp32


@pytest.mark.parametrize("K,nq", [(300, 5), (1000, 9), (4096, 24)])
def test_probe_tensor_core_equals_exact_fp32(K, nq):
    """Stage 1 on tcgen05 (fp16 operands) must pick the cells of the exact fp32 `_topk`
    (src/utils.jl:327-332): K not a multiple of the 128-centroid tile, row counts not a multiple of
    the 256-row unit, every nprobe; identical to the SIMT fp32 path bit for bit."""
    ix = S.make_index(50, K, seed=60 + nq, doclen_mean=10, doclen_std=3, doclen_min=1, doclen_max=20)
    with make_searcher(ix) as s:
        for nprobe in (1, 2, 4, 12):
            Qn = S.make_queries(ix["centroids"], nq, seed=70 + nprobe, nprobe=nprobe)
            Qj = np.transpose(Qn, (2, 1, 0))
            s.set_option("stage1_impl", 2)
            cells_tc, scores_tc = s.probe(Qj, nprobe=nprobe)
            assert s.stat("stage1_tc_rows") == nq * 32, "tcgen05 stage-1 kernel did not run"
            s.set_option("stage1_impl", 1)
            cells_fp, scores_fp = s.probe(Qj, nprobe=nprobe)
            assert s.stat("stage1_tc_rows") == 0
            assert np.array_equal(cells_tc, cells_fp) and np.array_equal(scores_tc, scores_fp)
            for q in range(nq):
                top = O._topk(O.centroid_scores(Qn[q].T, ix["centroids"].T), nprobe, dims=2)
                assert np.array_equal(cells_tc[q], top)


def test_probe_tensor_core_near_ties_fall_back_to_exact_scan():
    """More near-identical centroids than the shortlist holds, closer together than fp16 rounding
    can resolve: the guard must flag those rows and the exact scan decides (ties -> lower id, like
    `partialsortperm`)."""
    ix = S.make_index(50, 512, seed=81, doclen_mean=10, doclen_std=3, doclen_min=1, doclen_max=20)
    cen = ix["centroids"]
    for i in range(24):                      # 24 > CB_TOPR copies of centroid 7, last-bits differences
        cen[300 + i] = cen[7]
        cen[300 + i, i % 128] += np.float32(1e-7 * (i % 5))
    Q = np.stack([np.concatenate([cen[7:8].repeat(16, axis=0), cen[9:10].repeat(16, axis=0)])])
    with make_searcher(ix) as s:
        s.set_option("stage1_impl", 2)
        cells, scores = s.probe(np.transpose(Q, (2, 1, 0)), nprobe=2)
        assert s.stat("stage1_tc_rows") == 32 and s.stat("flagged_rows") >= 16
    fix = np.stack([O.fixed_order_dot(np.repeat(Q[0][t:t + 1], 512, axis=0), cen) for t in range(32)])
    want = np.stack([np.lexsort((np.arange(512), -fix[t]))[:2] + 1 for t in range(32)])
    assert np.array_equal(cells[0], want)
    assert np.array_equal(scores[0], np.take_along_axis(fix, want - 1, axis=1))


def test_probe_ties_resolve_to_lower_id():
    # duplicated centroids force exact score ties; `partialsortperm` keeps the lower index
    ix = S.make_index(200, 64, dim=32, seed=3, doclen_mean=6, doclen_std=2, doclen_min=1, doclen_max=10)
    ix["centroids"][10] = ix["centroids"][40]
    ix["centroids"][41] = ix["centroids"][40]
    Q = np.stack([ix["centroids"][40:41].repeat(4, axis=0)])  # one query, 4 tokens == centroid 40
    with make_searcher(ix, T=4) as s:
        cells, _ = s.probe(np.transpose(Q, (2, 1, 0)), nprobe=2)
    assert cells[0].tolist() == [[11, 41]] * 4
    top = O._topk(O.centroid_scores(Q[0].T, ix["centroids"].T), 2, dims=2)
    assert cells[0].tolist() == top.tolist()


def test_retrieve_candidate_sets_bit_exact(tiny):
    ix, Q, oix = tiny
    with make_searcher(ix) as s:
        for q in range(8):
            got = s.retrieve(Q[q].T)
            want = O.retrieve(oix.ivf, oix.ivf_lengths, oix.centroids, oix.emb2pid, 2, Q[q].T)
            assert np.array_equal(got, want)


def test_score_pids_matches_collect_decompress_maxsim(tiny):
    ix, Q, oix = tiny
    pids = np.array([5, 1, 1999, 2000, 5, 77, 1024], dtype=np.int64)   # unsorted, repeated (ranking.jl:124-134)
    with make_searcher(ix) as s:
        got = s.score_pids(Q[0].T, pids)
    cp, rp = O._collect_compressed_embs_for_pids(oix.doclens, oix.codes, oix.residuals, pids)
    D = O.decompress(oix.dim, oix.nbits, oix.centroids, oix.bucket_weights, cp, rp)
    want = O.maxsim(Q[0].T, D, pids, oix.doclens)
    np.testing.assert_allclose(got, want, rtol=1e-5)


@pytest.mark.parametrize("force_generic", [1, 0])
def test_search_batch_vs_oracle(tiny, force_generic):
    ix, Q, oix = tiny
    k = 10
    with make_searcher(ix) as s:
        s.set_option("force_generic", force_generic)
        pids, scores, counts = s.search_batch(np.transpose(Q, (2, 1, 0)), k)
        if not force_generic:
            assert s.stat("tc_pairs") > 0, "tensor-core scoring kernel did not run"
    for q in range(Q.shape[0]):
        allp, alls = O.search_all_scores(oix, Q[q].T)
        assert counts[q] == len(allp)
        op, osc = O.search(oix, Q[q].T, k)
        check_topk(pids[q], scores[q], op, osc, allp, alls, k)


def test_search_single_query_api_and_bounds_error(tiny):
    ix, Q, oix = tiny
    with make_searcher(ix) as s:
        pids, scores = cb.search(s, Q[3].T, 5)
        op, osc = O.search(oix, Q[3].T, 5)
        assert pids.shape == (5,) and scores.dtype == np.float32
        np.testing.assert_allclose(scores, osc, rtol=SCORE_RTOL)
        n_cand = len(O.search_all_scores(oix, Q[3].T)[0])
    tiny_ix = S.make_index(6, 64, dim=16, seed=5, doclen_mean=3, doclen_std=1, doclen_min=1, doclen_max=4)
    Qt = S.make_queries(tiny_ix["centroids"], 1, T=4, seed=6)
    with make_searcher(tiny_ix, T=4) as s:
        with pytest.raises(cb.BoundsError):      # searching.jl:127
            cb.search(s, Qt[0].T, 7)
    assert n_cand > 5


@pytest.mark.parametrize("nbits,dim,T", [(1, 128, 32), (4, 128, 32), (2, 64, 32), (2, 128, 8), (3, 96, 5)])
def test_search_other_shapes(nbits, dim, T):
    ix = S.make_index(600, 128, dim=dim, nbits=nbits, seed=20 + nbits, doclen_mean=30, doclen_std=20, doclen_min=1,
                      doclen_max=120)
    Q = S.make_queries(ix["centroids"], 6, T=T, seed=21)
    oix = oracle_index(ix)
    with make_searcher(ix, T=T) as s:
        pids, scores, counts = s.search_batch(np.transpose(Q, (2, 1, 0)), 7)
    for q in range(Q.shape[0]):
        allp, alls = O.search_all_scores(oix, Q[q].T)
        assert counts[q] == len(allp)
        if len(allp) >= 7:
            op, osc = O.search(oix, Q[q].T, 7)
            check_topk(pids[q], scores[q], op, osc, allp, alls, 7)


def test_edge_cases_empty_and_ragged():
    # zero-length passages, a 1-token passage, a passage longer than any tile, empty cells
    dim, K = 128, 64
    doclens = np.array([0, 1, 0, 700, 3, 0, 257, 256, 129, 0], dtype=np.int64)
    n_e = int(doclens.sum())
    r = np.random.default_rng(5)
    ix = dict(dim=dim, nbits=2, K=K, n_passages=len(doclens), n_embeddings=n_e,
              centroids=S.make_index(1, K, seed=9)["centroids"], bucket_weights=S.bucket_weights(2), doclens=doclens,
              codes=r.integers(1, K // 2, n_e).astype(np.uint32),      # upper half of the cells stays empty
              residuals=r.integers(0, 256, (n_e, 32), dtype=np.uint8))
    ix["ivf"], ix["ivf_lengths"] = S.build_ivf(ix["codes"], K)
    Q = S.make_queries(ix["centroids"], 5, seed=10)
    oix = oracle_index(ix)
    for fg in (1, 0):
        with make_searcher(ix) as s:
            s.set_option("force_generic", fg)
            pids, scores, counts = s.search_batch(np.transpose(Q, (2, 1, 0)), 4)
            for q in range(5):
                want = O.retrieve(oix.ivf, oix.ivf_lengths, oix.centroids, oix.emb2pid, 2, Q[q].T)
                assert np.array_equal(s.retrieve(Q[q].T), want)
                allp, alls = O.search_all_scores(oix, Q[q].T)
                assert counts[q] == len(allp)
                kk = min(4, len(allp))
                order = np.lexsort((allp, -alls))
                np.testing.assert_allclose(scores[q][:kk], alls[order][:kk], rtol=SCORE_RTOL)
                assert np.all(pids[q][kk:] == 0) and np.all(np.isneginf(scores[q][kk:]))
    # empty batch and empty index
    with make_searcher(ix) as s:
        p, sc, c = s.search_batch(np.zeros((dim, 32, 0), np.float32), 3)
        assert p.shape == (0, 3)
    empty = dict(ix, n_passages=0, n_embeddings=0, doclens=np.zeros(0, np.int64), codes=np.zeros(0, np.uint32),
                 residuals=np.zeros((0, 32), np.uint8), ivf=np.zeros(0, np.int64), ivf_lengths=np.zeros(K, np.int64))
    with make_searcher(empty) as s:
        p, sc, c = s.search_batch(np.transpose(Q, (2, 1, 0)), 3)
        assert c.tolist() == [0] * 5 and np.all(p == 0)


@pytest.mark.parametrize("n_passages,n_queries,doclen_mean,doclen_std,K", [
    (3000, 16, 40, 15, 256),     # every query a candidate of every passage: 4 full groups per passage, short MMAs (N = 48)
    (3000, 23, 30, 20, 256),     # ragged last groups (23 candidates = 5 full groups + 3)
    (1500, 9, 300, 60, 128),     # passages over 240 tokens: two accumulator passes per group
])
def test_two_issuer_pipeline_equals_generic_kernel(n_passages, n_queries, doclen_mean, doclen_std, K):
    """The tcgen05 scoring kernel issues alternate 4-query groups from two threads.  Its first version shared one
    "tile landed" barrier per query-tile stage between them; on exactly these shapes (many groups per passage, short
    passages, few passages per CTA) an issuer took an old phase of the barrier for its own and the kernel died with a
    launch failure.  The fused kernel must return what the generic fp32 kernel returns (final scores are exact fp32
    re-scores in both: bit-identical)."""
    ix = S.make_index(n_passages, K, seed=31, doclen_mean=doclen_mean, doclen_std=doclen_std)
    Q = S.make_queries(ix["centroids"], n_queries, seed=32)
    Qj = np.transpose(Q, (2, 1, 0))
    with make_searcher(ix) as s:
        for _ in range(3):                                           # timing-dependent: a few runs
            p1, s1, c1 = s.search_batch(Qj, 10)
        assert s.stat("tc_pairs") > 0
        s.set_option("force_generic", 1)
        p2, s2, c2 = s.search_batch(Qj, 10)
    assert np.array_equal(c1, c2)
    assert np.array_equal(p1, p2)
    np.testing.assert_array_equal(s1, s2)



def test_ivf_built_on_device_equals_given_ivf(tiny):
    ix, Q, oix = tiny
    no_ivf = dict(ix, ivf=None, ivf_lengths=None)
    with make_searcher(ix) as a, make_searcher(no_ivf) as b:
        for q in range(4):
            assert np.array_equal(a.retrieve(Q[q].T), b.retrieve(Q[q].T))


def test_topk_ties_ascending_pid():
    # identical passages score identically: the reference's stable sort keeps ascending pids
    ix = S.make_index(40, 32, dim=64, seed=8, doclen_mean=6, doclen_std=0, doclen_min=6, doclen_max=6)
    ix["codes"] = np.tile(ix["codes"][:6], 40)
    ix["residuals"] = np.tile(ix["residuals"][:6], (40, 1))
    ix["ivf"], ix["ivf_lengths"] = S.build_ivf(ix["codes"], 32)
    Q = ix["centroids"][ix["codes"][:4].astype(int) - 1][None]        # probes the passages' own cells
    with make_searcher(ix, T=4) as s:
        pids, scores, counts = s.search_batch(np.transpose(Q, (2, 1, 0)), 10)
    assert counts[0] == 40 and pids[0].tolist() == list(range(1, 11))
    assert np.all(scores[0] == scores[0][0])


def test_merge_topk_and_sharded_search_equals_unsharded():
    ix = S.make_index(3000, 256, seed=31, doclen_mean=40, doclen_std=15)
    Q = S.make_queries(ix["centroids"], 16, seed=32)
    Qj = np.transpose(Q, (2, 1, 0))
    k = 10
    with make_searcher(ix) as s:
        want_p, want_s, want_c = s.search_batch(Qj, k)
    bounds = S.shard_ranges(ix["doclens"], 3)
    parts = []
    for r in range(3):
        sh = S.take_shard(ix, int(bounds[r]), int(bounds[r + 1]))
        with make_searcher(sh) as s:
            parts.append(s.search_batch(Qj, k))
    mp, ms = cb.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    assert np.array_equal(sum(p[2] for p in parts), want_c)
    assert np.array_equal(mp, want_p)
    np.testing.assert_array_equal(ms, want_s)


# --------------------------------------------------------------------------------------------
# larger sizes: size-independent properties instead of the (slow) oracle
# --------------------------------------------------------------------------------------------
def test_properties_medium_index():
    ix = S.make_index(60000, 16384, seed=51)
    Q = S.make_queries(ix["centroids"], 64, seed=52)
    Qj = np.transpose(Q, (2, 1, 0))
    k = 10
    with make_searcher(ix) as s:
        p1, s1, c1 = s.search_batch(Qj, k)
        p2, s2, c2 = s.search_batch(Qj, k)                      # idempotence / determinism
        assert np.array_equal(p1, p2) and np.array_equal(s1, s2) and np.array_equal(c1, c2)
        assert np.all(np.diff(s1, axis=1) <= 0)                 # sortedness
        perm = rng.permutation(64)                              # batch-order invariance
        p3, s3, _ = s.search_batch(Qj[:, :, perm], k)
        assert np.array_equal(p3, p1[perm]) and np.array_equal(s3, s1[perm])
        pq, sq, _ = s.search_batch(Qj[:, :, 5:6], k)            # batch == single query
        assert np.array_equal(pq[0], p1[5]) and np.array_equal(sq[0], s1[5])
        # top-k scores agree with exact fp32 re-scoring of the same pids, and candidates are a set
        for q in (0, 17, 63):
            exact = s.score_pids(Q[q].T, p1[q])
            np.testing.assert_allclose(s1[q], exact, rtol=SCORE_RTOL)
            cand = s.retrieve(Q[q].T)
            assert len(cand) == c1[q] and np.all(np.diff(cand) > 0)
            assert np.all(np.isin(p1[q], cand))
        # every candidate truly owns a probed code (checksum of the candidate set vs a numpy recount)
        cells, _ = s.probe(Qj[:, :, :2])
        e2p = cb._build_emb2pid(ix["doclens"])
        for q in range(2):
            want = np.unique(e2p[np.isin(ix["codes"], np.unique(cells[q]))])
            assert np.array_equal(s.retrieve(Q[q].T), want)


# --------------------------------------------------------------------------------------------
# the decompression of the FUSED tcgen05 kernel itself (not the fp32 hook): bucket indices bit-exact
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbits", [1, 2, 4])
def test_fused_kernel_decompression_bit_exact(nbits):
    """`north_star`: decompressed codes bit-exact.  cb_debug_tc_operand runs the device functions of the
    hot kernel's decompression role (byte LUT, packed-fp16 add, normalisation, swizzled tile stores).
    (1) fp16(centroid) + fp16(w[bucket]) is exactly reproducible in numpy float16, so equality pins every
    unpacked bucket index of that code path (`_unpackbits`/`_unbinarize`, residual.jl:233-240, 428-441);
    (2) the normalised operand equals fp16(oracle `decompress`) within 1 fp16 ulp (+ the 2^-11 relative
    error the packed-fp16 norm carries).  Residual bytes cover all 256 values at every byte position."""
    dim, K, Np = 128, 64, 96
    r = np.random.default_rng(300 + nbits)
    doclens = r.integers(1, 40, Np).astype(np.int64)
    doclens[:6] = [1, 15, 16, 17, 240, 333]           # chunk-boundary cases: padding, 2-chunk passages
    Ne = int(doclens.sum())
    R = dim // 8 * nbits
    res = r.integers(0, 256, (Ne, R), dtype=np.uint8)
    res[:256, :] = np.arange(256, dtype=np.uint8)[:, None]          # every byte value at every byte position
    res[256:512, :] = (np.arange(256, dtype=np.uint8)[:, None] + np.arange(R, dtype=np.uint8)[None, :] * 37)
    codes = r.integers(1, K + 1, Ne).astype(np.uint32)
    cen = O._normalize_array(r.standard_normal((dim, K)).astype(np.float32))
    w = S.bucket_weights(nbits)
    ivf, ivf_lengths = O._build_ivf(codes.astype(np.int64), K)
    ix = dict(dim=dim, nbits=nbits, centroids=cen.T.copy(), bucket_weights=w, ivf=ivf, ivf_lengths=ivf_lengths,
              doclens=doclens, codes=codes, residuals=res)
    with make_searcher(ix) as s:
        norm, raw = s.debug_tc_operand(np.arange(1, Np + 1), Ne)
    idx = O.unpack_bucket_indices(dim, nbits, res.T).T                 # (Ne, dim) bucket indices, oracle
    c16 = cen.T.astype(np.float16)[codes.astype(np.int64) - 1]         # fp16(centroid row)
    w16 = w.astype(np.float16)[idx]
    expect_raw = (c16 + w16).astype(np.float16)                        # one correctly rounded fp16 add, like __hadd2
    assert np.array_equal(raw.view(np.uint16), expect_raw.view(np.uint16))
    # distinct weights -> the raw value identifies the bucket index for every (embedding, dim)
    assert len(np.unique(w.astype(np.float16))) == len(w)
    # (2a) normalisation: against the exact normalisation of those same fp16 sums (`_normalize_array!`: / (norm + eps))
    r64 = raw.astype(np.float64)
    exact = r64 / (np.sqrt((r64 * r64).sum(axis=1, keepdims=True)) + np.float64(np.finfo(np.float32).eps))
    ulp = np.spacing(np.abs(exact.astype(np.float16))).astype(np.float64)
    ulps_a = float((np.abs(norm.astype(np.float64) - exact) / ulp).max())
    # (2b) against the fp32 oracle `decompress` (adds the fp16 rounding of centroid, weight and their sum)
    o_emb = O.decompress(dim, nbits, cen, w, codes, res.T).T.astype(np.float64)
    # (absolute error against the largest component of the row: the centroid's fp16 rounding is relative to the
    # centroid component, so an "ulp of the result" is meaningless where centroid and weight nearly cancel)
    rel_b = float((np.abs(norm.astype(np.float64) - o_emb).max(axis=1) / np.abs(o_emb).max(axis=1)).max())
    print(f"nbits={nbits}: fused-kernel operand vs exact normalisation of its fp16 sums: {ulps_a:.2f} fp16 ulp; "
          f"vs fp32 oracle decompress: {rel_b:.2e} of the row maximum")
    assert ulps_a <= 2.5, ulps_a      # 1/2 ulp final rounding + the packed-fp16 norm (2-term fp16 chains, ~2^-10 relative)
    assert rel_b <= 2e-3, rel_b       # four half-ulp fp16 roundings (centroid, sum, two in the scaling) of a component near the row maximum


# --------------------------------------------------------------------------------------------
# f4: cb_compress <-> `compress` (src/indexing/codecs/residual.jl:586-604)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbits", [1, 2, 3, 4, 8])
def test_compress_vs_oracle(nbits):
    """codes == the oracle's argmax (fixture embeddings have a top-1 rank gap far above fp32 rounding, like the
    queries of the search fixtures); packed residual bytes bit-exact, dimensions straddling bytes included."""
    dim, K, n = 128, 700, 3000            # K not a multiple of the 128-centroid tile
    cen = O._normalize_array(rng.standard_normal((dim, K)).astype(np.float32))                    # (dim, K)
    embs = S.make_queries(cen.T.copy(), n // 32 + 1, seed=900 + nbits, nprobe=1).reshape(-1, dim)[:n].T.copy()   # (dim, n)
    cut = np.sort(rng.normal(0, 0.03, (1 << nbits) - 1).astype(np.float32))
    codes, res = cb.compress(cen, cut, dim, nbits, embs)
    o_codes, o_res = O.compress(cen, cut, dim, nbits, embs)
    assert codes.dtype == np.uint32 and res.dtype == np.uint8 and res.shape == (dim // 8 * nbits, n)
    assert np.array_equal(codes, o_codes)
    assert np.array_equal(res, o_res)


def test_compress_inverts_through_decompress_residuals():   # test/indexing/codecs/residual.jl:975-991
    """decompress_residuals(binarize(x)) == bucket_weights[searchsortedfirst(cutoffs, x)]: with one zero centroid the
    residual IS the embedding and cb_decompress's un-normalised output IS bucket_weights[idx]."""
    for nbits in (1, 2, 5):
        dim = 8 * int(rng.integers(1, 21))
        cut = np.sort(rng.random((1 << nbits) - 1, dtype=np.float32))
        w = np.sort(rng.random(1 << nbits, dtype=np.float32))
        x = rng.random((dim, int(rng.integers(1, 100))), dtype=np.float32)
        zero = np.zeros((dim, 1), np.float32)
        codes, packed = cb.compress(zero, cut, dim, nbits, x)
        assert np.all(codes == 1)
        assert np.array_equal(packed, O.binarize(dim, nbits, cut, x))
        _, raw = cb.decompress(dim, nbits, zero, w, codes, packed, return_unnormalized=True)
        assert np.array_equal(raw, w[np.searchsorted(cut, x, side="left")])


def test_compress_ties_and_errors():
    dim = 16
    cen = np.zeros((dim, 5), np.float32)
    cen[0, 1] = cen[0, 3] = 1.0                       # centroids 2 and 4 identical: argmax returns the first maximum
    x = np.zeros((dim, 2), np.float32)
    x[0, :] = 1.0
    codes, _ = cb.compress(cen, np.zeros(3, np.float32), dim, 2, x)
    assert codes.tolist() == [2, 2]
    with pytest.raises(cb.DomainError):               # residual.jl:525
        cb.compress(cen, np.zeros(2, np.float32), dim, 2, x)
    with pytest.raises(cb.DomainError):               # residual.jl:523
        cb.compress(np.zeros((12, 5), np.float32), np.zeros(3, np.float32), 12, 2, np.zeros((12, 2), np.float32))

"""Pins the CPU oracle (oracle/oracle.py) against every golden vector the reference's
own tests hold for the search-time scoring path.  Each test
cites the reference test file:line it restates.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(1234)


# ---- test/search/ranking.jl ------------------------------------------------------------
def test_cids_to_eids_golden():  # test/search/ranking.jl:5-11
    eids = np.empty(5, dtype=np.int64)
    O._cids_to_eids(eids, [2, 1], [1, 2, 3, 4, 5, 6], [3, 2, 1])
    assert eids.tolist() == [4, 5, 1, 2, 3]


def test_cids_to_eids_random_partition():  # test/search/ranking.jl:13-36
    for _ in range(5):
        n_e, n_part = int(rng.integers(1, 1000)), int(rng.integers(1, 20))
        assign = rng.integers(1, n_part + 1, n_e)
        mapping = [[] for _ in range(n_part)]
        for eid, a in enumerate(assign, start=1):
            mapping[a - 1].append(eid)
        ivf, lens = [], []
        for m in mapping:
            rng.shuffle(m)
            ivf += m
            lens.append(len(m))
        cids = (rng.permutation(n_part) + 1)[: int(rng.integers(1, n_part + 1))]
        eids = np.empty(sum(lens[c - 1] for c in cids), dtype=np.int64)
        O._cids_to_eids(eids, cids, ivf, lens)
        pos = 0
        for c in cids:
            assert eids[pos:pos + lens[c - 1]].tolist() == mapping[c - 1]
            pos += lens[c - 1]


def test_cids_to_eids_empty():  # test/search/ranking.jl:38-52
    eids = np.empty(0, dtype=np.int64)
    O._cids_to_eids(eids, [], [], [])
    assert eids.tolist() == []
    O._cids_to_eids(eids, [], [1, 2, 3, 4, 5, 6], [3, 2, 1])
    assert eids.tolist() == []


def test_cids_to_eids_mismatch():  # test/search/ranking.jl:54-68
    with pytest.raises(O.DimensionMismatch):
        O._cids_to_eids(np.empty(5, np.int64), [1, 2, 3], [1, 2, 3, 4, 5, 6], [2, 2, 2])
    with pytest.raises(O.DimensionMismatch):
        O._cids_to_eids(np.empty(6, np.int64), [1, 2, 3], [1, 2, 3, 4, 5], [2, 2, 2])


def test_retrieve_golden():  # test/search/ranking.jl:71-83
    ivf = [3, 1, 4, 5, 6, 2]
    ivf_lengths = [2, 3, 1]
    centroids = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    emb2pid = np.array([10, 20, 30, 40, 50, 60])
    Q = np.array([[0.5, 0.5]], dtype=np.float32).T
    assert O.retrieve(ivf, ivf_lengths, centroids, emb2pid, 2, Q).tolist() == [10, 20, 30]


def test_collect_compressed_embs_golden():  # test/search/ranking.jl:86-121
    doclens = [3, 2, 4]
    codes = np.arange(1, 10, dtype=np.uint32)
    residuals = np.array([[0x11, 0x12, 0x13, 0x14, 0x15, 0x16, 0x17, 0x18, 0x19],
                          [0x21, 0x22, 0x23, 0x24, 0x25, 0x26, 0x27, 0x28, 0x29]], dtype=np.uint8)
    c, r = O._collect_compressed_embs_for_pids(doclens, codes, residuals, [1, 3])
    assert c.tolist() == [1, 2, 3, 6, 7, 8, 9]
    assert r.tolist() == [[0x11, 0x12, 0x13, 0x16, 0x17, 0x18, 0x19],
                          [0x21, 0x22, 0x23, 0x26, 0x27, 0x28, 0x29]]
    c, r = O._collect_compressed_embs_for_pids(doclens, codes, residuals, [])
    assert c.shape == (0,) and r.shape == (2, 0)
    # zero doclens
    doclens = [3, 0, 4]
    codes = np.array([1, 2, 3, 6, 7, 8, 9], dtype=np.uint32)
    residuals = residuals[:, [0, 1, 2, 5, 6, 7, 8]]
    c, r = O._collect_compressed_embs_for_pids(doclens, codes, residuals, [1, 3])
    assert c.tolist() == [1, 2, 3, 6, 7, 8, 9]
    assert r.tolist() == residuals.tolist()


def test_collect_compressed_embs_shapes():  # test/search/ranking.jl:123-134
    n = int(rng.integers(1, 1000))
    doclens = rng.integers(1, 101, n)
    codes = rng.integers(0, 2**32, doclens.sum(), dtype=np.uint32)
    residuals = rng.integers(0, 256, (16, doclens.sum()), dtype=np.uint8)
    pids = rng.integers(1, n + 1, int(rng.integers(1, n + 1)))
    c, r = O._collect_compressed_embs_for_pids(doclens, codes, residuals, pids)
    assert len(c) == doclens[pids - 1].sum() and r.shape == (16, doclens[pids - 1].sum())
    assert c.dtype == np.uint32 and r.dtype == np.uint8
    cf, rf = O._collect_compressed_embs_for_pids_fast(doclens, codes, residuals, pids)
    assert np.array_equal(cf, c) and np.array_equal(rf, r)


def test_maxsim_golden():  # test/search/ranking.jl:137-152  (tested with == upstream)
    Q = np.array([[1.0, 0.5], [0.5, 1.0]], dtype=np.float32)
    D = np.array([[0.8, 0.3, 0.1], [0.2, 0.7, 0.4]], dtype=np.float32)
    s = O.maxsim(Q, D, [1, 2], [1, 2])
    assert s.dtype == np.float32 and s.tolist() == [1.5, 1.5]
    assert O.maxsim_fast(Q, D, [1, 2], [1, 2]).tolist() == [1.5, 1.5]
    with pytest.raises(O.DimensionMismatch):
        O.maxsim(Q, np.array([[0.8, 0.3]], dtype=np.float32), [1, 2], [1, 2])


def test_maxsim_shapes():  # test/search/ranking.jl:154-161
    doclens = rng.integers(1, 11, 1000)
    Q = rng.random((128, 100), dtype=np.float32)
    D = rng.random((128, doclens.sum()), dtype=np.float32)
    s = O.maxsim(Q, D, np.arange(1, 1001), doclens)
    assert s.shape == (1000,) and s.dtype == np.float32
    np.testing.assert_allclose(O.maxsim_fast(Q, D, np.arange(1, 1001), doclens), s, rtol=1e-6)


# ---- test/searching.jl -----------------------------------------------------------------
def test_build_emb2pid_golden():  # test/searching.jl:4-41
    n = int(rng.integers(1, 1000))
    assert O._build_emb2pid([n]).tolist() == [1] * n
    assert O._build_emb2pid([3, 2, 4]).tolist() == [1, 1, 1, 2, 2, 3, 3, 3, 3]
    assert O._build_emb2pid([0, 2, 0, 3]).tolist() == [2, 2, 4, 4, 4]
    assert O._build_emb2pid([]).tolist() == []
    doclens = rng.integers(0, 101, int(rng.integers(1, 500)))
    e2p = O._build_emb2pid(doclens)
    assert np.all(np.diff(e2p) >= 0) and len(e2p) == doclens.sum()
    for pid in np.nonzero(doclens)[0] + 1:
        assert np.count_nonzero(e2p == pid) == doclens[pid - 1]


# ---- test/utils.jl ---------------------------------------------------------------------
def test_normalize_array():  # test/utils.jl:147-161
    X = rng.random((int(rng.integers(1, 100)), int(rng.integers(1, 100))), dtype=np.float32)
    O._normalize_array(X, dims=1)
    np.testing.assert_allclose(np.linalg.norm(X, axis=0), 1, rtol=1e-5)
    X = rng.random((17, 9), dtype=np.float32)
    O._normalize_array(X, dims=2)
    np.testing.assert_allclose(np.linalg.norm(X, axis=1), 1, rtol=1e-5)


def test_topk_golden():  # test/utils.jl:163-181
    data = np.array([[3.0, 1.0, 4.0], [1.0, 5.0, 9.0], [2.0, 6.0, 5.0]])
    assert O._topk(data, 2, dims=1).tolist() == [[1, 3, 2], [3, 2, 3]]
    assert O._topk(data, 2, dims=2).tolist() == [[3, 1], [3, 2], [2, 3]]
    assert O._topk_rows_fast(data, 2).tolist() == [[3, 1], [3, 2], [2, 3]]
    with pytest.raises(O.DomainError):
        O._topk(data, 2, dims=3)


def test_topk_ties_lower_index_first():  # Perm ordering of partialsortperm
    data = np.array([[1.0, 2.0, 2.0, 2.0, 0.5]])
    assert O._topk(data, 2, dims=2).tolist() == [[2, 3]]
    assert O._topk_rows_fast(data, 2).tolist() == [[2, 3]]


def test_head_golden():  # test/utils.jl:183-213
    assert O._head([1, 2, 3, 4]) == [1, 2, 3]
    assert O._head([10]) == []
    assert O._head([]) == []
    assert O._head(["a", "b", "c"]) == ["a", "b"]
    assert O._head([1.5, 2.5, 3.5]) == [1.5, 2.5]


# ---- test/indexing/codecs/residual.jl --------------------------------------------------
def _jl(vals, *shape):
    """Julia `reshape(vals, shape...)` (column-major)."""
    return np.array(vals).reshape(shape, order="F")


def test_binarize_golden():  # test/indexing/codecs/residual.jl:59-100
    cases = [
        (np.array([[0, 1], [2, 3]]), 3, _jl([0, 0, 0, 0, 1, 0, 1, 0, 0, 1, 1, 0], 3, 2, 2)),
        (np.array([[0, 1], [2, 3]]), 2, _jl([0, 0, 0, 1, 1, 0, 1, 1], 2, 2, 2)),
        (np.array([[7]]), 3, _jl([1, 1, 1], 3, 1, 1)),
        (np.array([[0, 1], [0, 1]]), 1, _jl([0, 0, 1, 1], 1, 2, 2)),
    ]
    for data, nbits, expected in cases:
        assert np.array_equal(O._binarize(data, nbits), expected.astype(bool))
    with pytest.raises(O.DomainError):
        O._binarize(np.array([[0, 1], [4, 2]]), 2)


def test_unbinarize_golden():  # test/indexing/codecs/residual.jl:117-151
    nbits = int(rng.integers(1, 11))
    z = np.zeros((nbits, 7, 5), dtype=bool)
    assert np.array_equal(O._unbinarize(z), np.zeros((7, 5), dtype=np.int64))
    assert np.array_equal(O._unbinarize(~z), ((1 << nbits) - 1) * np.ones((7, 5), dtype=np.int64))
    assert O._unbinarize(_jl([1, 0, 0, 1, 1], 5, 1, 1).astype(bool)).tolist() == [[25]]
    data = _jl([1, 1, 1, 0, 1, 1, 1, 0, 0, 0, 1, 1,
                0, 0, 1, 0, 1, 0, 0, 0, 0, 1, 1, 0], 6, 2, 2).astype(bool)
    assert O._unbinarize(data).tolist() == [[55, 20], [49, 24]]


def test_unbinarize_inverts_binarize():  # test/indexing/codecs/residual.jl:153-161
    nbits = int(rng.integers(1, 21))
    data = rng.integers(0, 1 << nbits, (int(rng.integers(1, 21)), int(rng.integers(1, 21))))
    assert np.array_equal(O._unbinarize(O._binarize(data, nbits)), data)


def test_bucket_indices_golden():  # test/indexing/codecs/residual.jl:163-214
    assert O._bucket_indices(np.array([[1, 6], [3, 12]]), [0, 5, 10, 15]).tolist() == [[1, 2], [1, 3]]
    assert O._bucket_indices(np.array([[5, 15]]), np.zeros(0, np.float32)).tolist() == [[0, 0]]
    assert O._bucket_indices(np.array([[1.1, 2.5, 7.8]]), [0.0, 2.0, 5.0, 10.0]).tolist() == [[1, 2, 3]]


def test_packbits_golden():  # test/indexing/codecs/residual.jl:216-268
    bits = [1, 1, 0, 1, 1, 0, 1, 1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 1, 0, 1, 1, 1,
            0, 0, 0, 1, 0, 1, 1, 0, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0,
            1, 1, 0, 1, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 0]
    expected = [0b11011011, 0b01101010, 0b00111011, 0b11011010, 0b11100000,
                0b01011001, 0b11111011, 0b01010001]
    out = O._packbits(_jl(bits, 1, 64, 1).astype(bool))
    assert out.shape == (8, 1) and out[:, 0].tolist() == expected
    shape = (int(rng.integers(1, 21)), 8 * int(rng.integers(1, 21)), int(rng.integers(1, 21)))
    R = shape[0] * shape[1] // 8
    assert np.array_equal(O._packbits(np.zeros(shape, bool)), np.zeros((R, shape[2]), np.uint8))
    assert np.array_equal(O._packbits(np.ones(shape, bool)), np.full((R, shape[2]), 0xFF, np.uint8))
    alt = np.ones(int(np.prod(shape)), bool)
    alt[1::2] = False
    assert np.array_equal(O._packbits(alt.reshape(shape, order="F")), np.full((R, shape[2]), 0x55, np.uint8))
    with pytest.raises(O.DomainError):
        O._packbits(np.ones((3, 7, 5), bool))


def test_unpackbits_golden():  # test/indexing/codecs/residual.jl:270-832
    g = json.load(open(os.path.join(HERE, "golden", "unpackbits_test1.json")))
    packed = np.array(g["packed"], dtype=np.uint8).reshape(g["packed_shape_julia"], order="F")
    out = O._unpackbits(packed, g["nbits"])
    assert out.shape == (1, 8, 64)
    assert out.flatten(order="F").astype(int).tolist() == g["expected_bits"]
    nbits = int(rng.integers(1, 11))
    z = np.zeros((nbits * 3, 4), np.uint8)
    assert not O._unpackbits(z, nbits).any() and O._unpackbits(z, nbits).shape == (nbits, 24, 4)
    assert O._unpackbits(z + 0xFF, nbits).all()


def test_unpackbits_inverts_packbits():  # test/indexing/codecs/residual.jl:843-850
    nbits = int(rng.integers(1, 11))
    bits = rng.integers(0, 2, (nbits, 8 * int(rng.integers(1, 21)), int(rng.integers(1, 21)))).astype(bool)
    assert np.array_equal(O._unpackbits(O._packbits(bits), nbits), bits)


def test_binarize_shapes_and_errors():  # test/indexing/codecs/residual.jl:852-869
    dim, nbits = 8 * int(rng.integers(1, 21)), int(rng.integers(1, 9))
    cut = np.sort(rng.random((1 << nbits) - 1, dtype=np.float32))
    res = rng.random((dim, 13), dtype=np.float32)
    b = O.binarize(dim, nbits, cut, res)
    assert b.dtype == np.uint8 and b.shape == (dim // 8 * nbits, 13)
    with pytest.raises(O.DomainError):
        O.binarize(7, 7, np.sort(rng.random(127)), rng.random((7, 10)))
    with pytest.raises(O.DomainError):
        O.binarize(8, 8, np.sort(rng.random(254)), rng.random((8, 10)))


def test_decompress_residuals_errors():  # test/indexing/codecs/residual.jl:955-973
    with pytest.raises(O.DomainError):
        O.decompress_residuals(7, 7, np.zeros(128, np.float32), np.zeros((0, 3), np.uint8))
    with pytest.raises(O.DomainError):
        O.decompress_residuals(8, 8, np.zeros(255, np.float32), np.zeros((8, 3), np.uint8))
    with pytest.raises(O.DomainError):
        O.decompress_residuals(8, 8, np.zeros(256, np.float32), np.zeros((7, 64), np.uint8))


@pytest.mark.parametrize("nbits", [1, 2, 3, 4, 5, 8])
def test_decompress_residuals_inverts_binarize(nbits):  # residual.jl tests :975-991 (exact ==)
    dim = 8 * int(rng.integers(1, 21))
    cut = np.sort(rng.random((1 << nbits) - 1, dtype=np.float32))
    w = np.sort(rng.random(1 << nbits, dtype=np.float32))
    res = rng.random((dim, int(rng.integers(1, 100))), dtype=np.float32)
    expected = w[np.searchsorted(cut, res, side="left")]
    got = O.decompress_residuals(dim, nbits, w, O.binarize(dim, nbits, cut, res))
    assert np.array_equal(expected, got)


def test_nbits2_layout_closed_form():  # SURVEY 8a "Resulting bit layout"
    dim = 128
    packed = rng.integers(0, 256, (32, 50), dtype=np.uint8)
    idx = O.unpack_bucket_indices(dim, 2, packed)
    d = np.arange(dim)
    assert np.array_equal(idx, (packed[d >> 2, :] >> (2 * (d & 3))[:, None]) & 3)
    packed4 = rng.integers(0, 256, (64, 50), dtype=np.uint8)
    idx4 = O.unpack_bucket_indices(dim, 4, packed4)
    assert np.array_equal(idx4, (packed4[d >> 1, :] >> (4 * (d & 1))[:, None]) & 15)
    packed1 = rng.integers(0, 256, (16, 50), dtype=np.uint8)
    idx1 = O.unpack_bucket_indices(dim, 1, packed1)
    assert np.array_equal(idx1, (packed1[d >> 3, :] >> (d & 7)[:, None]) & 1)
    for nb, pk, want in ((1, packed1, idx1), (2, packed, idx), (4, packed4, idx4)):
        assert np.array_equal(O.unpack_bucket_indices_fast(dim, nb, pk), want)
    cen = O._normalize_array(rng.standard_normal((dim, 9)).astype(np.float32))
    codes = rng.integers(1, 10, 50).astype(np.uint32)
    w = np.linspace(-0.04, 0.04, 4).astype(np.float32)
    a, ra = O.decompress(dim, 2, cen, w, codes, packed, fast=True, return_unnormalized=True)
    b, rb = O.decompress(dim, 2, cen, w, codes, packed, fast=False, return_unnormalized=True)
    assert np.array_equal(ra, rb)                       # c + w[b]: identical
    np.testing.assert_allclose(a, b, rtol=1e-6)         # norm summation order may differ by an ulp


def test_decompress_shapes():  # test/indexing/codecs/residual.jl:993-1007
    dim, nbits, b = 8 * int(rng.integers(1, 21)), int(rng.integers(1, 9)), int(rng.integers(1, 100))
    w = np.sort(rng.random(1 << nbits, dtype=np.float32))
    cen = rng.random((dim, int(rng.integers(1, 100))), dtype=np.float32)
    codes = rng.integers(1, cen.shape[1] + 1, b).astype(np.uint32)
    res = rng.integers(0, 256, (dim // 8 * nbits, b), dtype=np.uint8)
    emb = O.decompress(dim, nbits, cen, w, codes, res, bsize=int(rng.integers(1, b + 6)))
    assert emb.dtype == np.float32 and emb.shape == (dim, b)
    np.testing.assert_allclose(np.linalg.norm(emb, axis=0), 1, rtol=1e-5)
    with pytest.raises(O.DomainError):
        O.decompress(dim, nbits, cen, w, codes[:-1], res)
    bad = codes.copy()
    bad[0] = cen.shape[1] + 1
    with pytest.raises(O.DomainError):
        O.decompress(dim, nbits, cen, w, bad, res)


def test_compress_roundtrip_properties():  # test/indexing/codecs/residual.jl:871-953 (cases 2, 3)
    dim, nbits = 8 * int(rng.integers(1, 21)), int(rng.integers(1, 9))
    cut = np.sort(rng.random((1 << nbits) - 1, dtype=np.float32))
    embs = O._normalize_array(rng.random((dim, int(rng.integers(2, 21))), dtype=np.float32))
    perm = rng.permutation(embs.shape[1])
    codes, res = O.compress(embs[:, perm], cut, dim, nbits, embs)
    assert codes.tolist() == (np.argsort(perm) + 1).tolist()
    assert not res.any()


# ---- test/indexing/collection_indexer.jl -----------------------------------------------
def test_build_ivf_golden():  # test/indexing/collection_indexer.jl:286-292
    codes = np.array([5, 3, 8, 2, 5, 5, 4, 2, 2, 1, 3], dtype=np.uint32)
    ivf, lens = O._build_ivf(codes, 10)
    assert ivf.tolist() == [10, 4, 8, 9, 2, 11, 7, 1, 5, 6, 3]
    assert lens.tolist() == [1, 3, 2, 1, 3, 0, 0, 1, 0, 0]


def test_bucket_cutoffs_and_weights_golden():  # test/indexing/collection_indexer.jl:85-93
    h = np.array([[0.0, 0.2], [0.4, 0.6], [0.8, 1.0]], dtype=np.float32)
    cut, w = O._bucket_cutoffs_and_weights(2, h)
    np.testing.assert_allclose(cut, [0.25, 0.5, 0.75], rtol=1e-6)
    np.testing.assert_allclose(w, [0.125, 0.375, 0.625, 0.875], rtol=1e-6)


# ---- search (unpinned upstream; consistency of the restatement with itself) ------------
def test_search_small_consistency():
    dim, nbits, K, n_p = 16, 2, 8, 30
    cen = O._normalize_array(rng.standard_normal((dim, K)).astype(np.float32))
    doclens = rng.integers(1, 6, n_p)
    n_e = int(doclens.sum())
    codes = rng.integers(1, K + 1, n_e).astype(np.uint32)
    res = rng.integers(0, 256, (dim // 8 * nbits, n_e), dtype=np.uint8)
    w = np.array([-0.04, -0.01, 0.01, 0.04], dtype=np.float32)
    ivf, lens = O._build_ivf(codes, K)
    idx = O.Index(dim, nbits, cen, w, ivf, lens, doclens, codes, res, nprobe=2)
    Q = O._normalize_array(rng.standard_normal((dim, 4)).astype(np.float32))
    pids, scores = O.search(idx, Q, 3)
    allp, alls = O.search_all_scores(idx, Q, fast=False)
    # brute force: every passage containing a probed code
    cells, _, _ = O.probe_cells(Q, cen, 2)
    e2p = O._build_emb2pid(doclens)
    brute = np.unique(e2p[np.isin(codes, cells)])
    assert allp.tolist() == brute.tolist()
    order = np.lexsort((allp, -alls))
    assert pids.tolist() == allp[order][:3].tolist()
    with pytest.raises(O.BoundsError):
        O.search(idx, Q, len(allp) + 1)

"""CPU tests of the synthetic index / query generator the parity tests and the bench fixtures use
(SURVEY.md section 8d): its IVF must be the reference's `_build_ivf` of its codes
(src/indexing/collection_indexer.jl:349-353), shapes must be the C layouts the ABI takes, and the
queries must have a well-defined probed-cell set (rank gap) so that candidate sets do not depend on
the summation order of the centroid scores."""
import numpy as np

from colbert_jl_b200 import synthetic as S
from oracle import oracle as O


def test_index_shapes_and_ivf():
    for nbits in (1, 2, 4):
        ix = S.make_index(300, 64, nbits=nbits, seed=5)
        K, dim = ix["centroids"].shape
        assert (K, dim) == (64, 128) and ix["dim"] == 128 and ix["nbits"] == nbits
        n_e = int(ix["doclens"].sum())
        assert ix["codes"].shape == (n_e,) and ix["codes"].min() >= 1 and ix["codes"].max() <= K
        assert ix["residuals"].shape == (n_e, dim // 8 * nbits) and ix["residuals"].dtype == np.uint8
        assert len(ix["bucket_weights"]) == 2 ** nbits
        np.testing.assert_allclose(np.linalg.norm(ix["centroids"], axis=1), 1.0, rtol=1e-5)
        ivf, ivl = O._build_ivf(ix["codes"], K)
        assert np.array_equal(ivf, ix["ivf"]) and np.array_equal(ivl, ix["ivf_lengths"])
        assert ivl.sum() == n_e


def test_index_is_deterministic_and_profiles_differ():
    a, b = S.make_index(200, 128, seed=9), S.make_index(200, 128, seed=9)
    assert all(np.array_equal(a[k], b[k]) for k in ("centroids", "codes", "residuals", "doclens"))
    c = S.make_index(200, 128, seed=9, profile="clustered")
    assert np.array_equal(a["doclens"], c["doclens"]) and not np.array_equal(a["codes"], c["codes"])
    # clustered: most tokens of a passage come from few centroids
    off = np.concatenate([[0], np.cumsum(c["doclens"])])
    distinct = np.mean([len(np.unique(c["codes"][off[i]:off[i + 1]])) / max(1, c["doclens"][i]) for i in range(200)])
    distinct_u = np.mean([len(np.unique(a["codes"][off[i]:off[i + 1]])) / max(1, a["doclens"][i]) for i in range(200)])
    assert distinct < distinct_u


def test_queries_have_a_rank_gap():
    ix = S.make_index(50, 512, seed=11)
    for nprobe in (2, 4):
        Q = S.make_queries(ix["centroids"], 4, seed=12, nprobe=nprobe, min_gap=1e-4)
        assert Q.shape == (4, 32, 128) and Q.dtype == np.float32
        np.testing.assert_allclose(np.linalg.norm(Q, axis=2), 1.0, rtol=1e-5)
        sc = Q.reshape(-1, 128) @ ix["centroids"].T
        srt = -np.sort(-sc, axis=1)
        assert np.all(srt[:, nprobe - 1] - srt[:, nprobe] > 1e-4 * 0.5)     # the gap the generator enforces (fp slack)

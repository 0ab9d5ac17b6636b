"""GPU parity tests of the PLAID-style pruned search (BASELINE.json config 5) through the C ABI against
`oracle.plaid_search`.  The reference has no implementation of this mode (README.md:187: roadmap), so
the oracle's semantics are the definition ("PLAID knobs"): parity UNPINNED by the
reference, pinned here against the oracle bit for bit where the arithmetic allows it:
  * the selected set (first ndocs candidates under approximate score desc, pid asc) is compared through
    `counts` and the final top-k, which the exact scores (tolerance 1e-3) then order;
  * with ndocs >= #candidates the result must equal the exhaustive search with nprobe = ncells.
"""
import numpy as np
import pytest

import colbert_jl_b200 as cb
from colbert_jl_b200 import synthetic as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu
SCORE_RTOL = 1e-3


def make_searcher(ix, nprobe=2, T=32):
    cfg = cb.ColBERTConfig(dim=ix["dim"], nbits=ix["nbits"], nprobe=nprobe, query_maxlen=T)
    return cb.Searcher(cfg, ix["centroids"].T, None, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"],
                       ix["doclens"], ix["codes"], ix["residuals"].T)


def oracle_index(ix, nprobe=2):
    return O.Index(ix["dim"], ix["nbits"], ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"],
                   ix["doclens"], ix["codes"], ix["residuals"].T, nprobe=nprobe)


def check_against_oracle(ix, Q, k, ncells, thr, ndocs):
    """Q (nq, T, dim) as make_queries returns it."""
    oix = oracle_index(ix)
    with make_searcher(ix, T=Q.shape[1]) as s:
        pids, scores, counts = s.search_batch_plaid(np.transpose(Q, (2, 1, 0)), k, ncells=ncells,
                                                    centroid_score_threshold=thr, ndocs=ndocs)
        rescored = s.stat("plaid_rescored")
    total = 0
    for q in range(Q.shape[0]):
        op, osc, sel, cand, approx = O.plaid_search(oix, Q[q].T, k, ncells, thr, ndocs, return_selected=True)
        assert counts[q] == len(sel) == min(ndocs, len(cand))
        total += len(sel)
        kk = len(op)
        np.testing.assert_allclose(scores[q, :kk], osc, rtol=SCORE_RTOL, atol=1e-5)
        assert np.all(pids[q, kk:] == 0)
        # every returned pid must be one the oracle selected (the selection itself is exact arithmetic) ...
        assert set(pids[q, :kk].tolist()) <= set(sel.tolist())
        # ... and the order may differ from the oracle's only between scores tied inside the tolerance
        if not np.array_equal(pids[q, :kk], op):
            codes_packed, res_packed = O._collect_compressed_embs_for_pids_fast(oix.doclens, oix.codes, oix.residuals, sel)
            D = O.decompress(oix.dim, oix.nbits, oix.centroids, oix.bucket_weights, codes_packed, res_packed, fast=True)
            exact = dict(zip(sel.tolist(), O.maxsim(Q[q].T, D, sel, oix.doclens).tolist()))
            for p, o in zip(pids[q, :kk], op):
                assert abs(exact[int(p)] - exact[int(o)]) <= SCORE_RTOL * max(1.0, abs(osc[-1]))
    assert rescored == total


@pytest.mark.parametrize("thr,ndocs", [(0.4, 64), (0.3, 200), (0.6, 32)])
def test_plaid_matches_oracle(thr, ndocs):
    ix = S.make_index(6000, 4096, seed=21)
    Q = S.make_queries(ix["centroids"], 12, seed=22, nprobe=4)
    check_against_oracle(ix, Q, 10, 4, thr, ndocs)


def test_plaid_zero_score_fill():
    """A threshold nothing reaches: every approximate score is 0, the selection is the first ndocs
    candidates in ascending pid order."""
    ix = S.make_index(3000, 2048, seed=23)
    Q = S.make_queries(ix["centroids"], 5, seed=24, nprobe=4)
    check_against_oracle(ix, Q, 10, 4, 1.5, 40)
    check_against_oracle(ix, Q, 10, 2, 0.97, 100)      # a few positives, the rest filled


def test_plaid_ndocs_covers_all_candidates_equals_exhaustive():
    ix = S.make_index(900, 2048, seed=25)
    Q = S.make_queries(ix["centroids"], 6, seed=26, nprobe=4)
    with make_searcher(ix, nprobe=4) as s:
        p1, s1, c1 = s.search_batch_plaid(np.transpose(Q, (2, 1, 0)), 10, ncells=4, centroid_score_threshold=0.4, ndocs=1024)
        p0, s0, c0 = s.search_batch(np.transpose(Q, (2, 1, 0)), 10, nprobe=4)
    assert np.all(c0 <= 1024), "test index too large for this property"
    assert np.array_equal(c1, c0)
    assert np.array_equal(p1, p0)
    np.testing.assert_array_equal(s1, s0)


def test_plaid_short_queries_and_other_nbits():
    ix = S.make_index(2000, 1024, seed=27, nbits=4)
    Q = S.make_queries(ix["centroids"], 4, seed=28, nprobe=3)[:, :8, :]      # T = 8: generic scoring kernel
    check_against_oracle(ix, np.ascontiguousarray(Q), 5, 3, 0.4, 50)


def test_plaid_argument_errors():
    ix = S.make_index(500, 256, seed=29)
    Q = S.make_queries(ix["centroids"], 2, seed=30)
    with make_searcher(ix) as s:
        with pytest.raises(cb.Unsupported):
            s.search_batch_plaid(np.transpose(Q, (2, 1, 0)), 10, ncells=4, ndocs=5000)
        with pytest.raises(cb.Unsupported):
            s.search_batch_plaid(np.transpose(Q, (2, 1, 0)), 10, ncells=40)
        # a threshold so low that every centroid survives for every query: more hits per passage than the build holds
        with pytest.raises(cb.Unsupported):
            s.search_batch_plaid(np.transpose(Q, (2, 1, 0)), 10, ncells=4, centroid_score_threshold=-1.0, ndocs=100)

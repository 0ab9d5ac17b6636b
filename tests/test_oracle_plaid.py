"""CPU tests of the oracle's PLAID-style pruned search (oracle.plaid_search).  The reference has no
implementation of it (README.md:187 roadmap; SURVEY.md section 8c: parity unpinned), so these pin the
oracle's own invariants: the pieces it is built from are the reference's pinned functions, and it
degenerates to the reference's exhaustive `search` when nothing is pruned."""
import numpy as np

from colbert_jl_b200 import synthetic as S
from oracle import oracle as O


def oracle_index(ix, nprobe=2):
    return O.Index(ix["dim"], ix["nbits"], ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"],
                   ix["doclens"], ix["codes"], ix["residuals"].T, nprobe=nprobe)


def test_tree_sum32_order():
    v = np.arange(32, dtype=np.float32)
    assert O.tree_sum32(v) == 496.0
    # the butterfly order, spelled out for 4 values padded with zeros: ((a0 + 0) + (a2 + 0)) + ((a1 + 0) + (a3 + 0)) ...
    big = np.zeros(32, dtype=np.float32)
    big[0], big[16], big[8] = 1e8, 1.0, -1e8
    # step 16: a[0] = 1e8 + 1 = 1e8 (fp32), a[8] = -1e8; step 8: a[0] = 0
    assert O.tree_sum32(big) == 0.0
    assert O.tree_sum32(np.ones(8, dtype=np.float32)) == 8.0      # zero padding


def test_fixed_order_scores_matches_fixed_order_dot():
    rng = np.random.default_rng(3)
    Q = rng.standard_normal((16, 5)).astype(np.float32)
    C = rng.standard_normal((16, 9)).astype(np.float32)
    Sc = O.fixed_order_scores(Q, C)
    for t in range(5):
        for c in range(9):
            assert Sc[t, c] == O.fixed_order_dot(Q[:, t][None, :], C[:, c][None, :])[0]


def test_plaid_without_pruning_is_exhaustive_search():
    ix = S.make_index(600, 512, seed=41)
    Q = S.make_queries(ix["centroids"], 3, seed=42, nprobe=4)
    oix = oracle_index(ix, nprobe=4)
    for q in range(3):
        p, s, sel, cand, approx = O.plaid_search(oracle_index(ix), Q[q].T, 10, 4, 0.4, 10 ** 6, return_selected=True)
        pe, se = O.search(oix, Q[q].T, 10)
        assert np.array_equal(sel, cand)
        assert np.array_equal(p, pe)
        np.testing.assert_allclose(s, se, rtol=1e-6)      # maxsim vs maxsim_fast: last-ulp summation differences


def test_plaid_selection_rule():
    ix = S.make_index(1500, 2048, seed=43)
    Q = S.make_queries(ix["centroids"], 2, seed=44, nprobe=4)
    oix = oracle_index(ix)
    for q in range(2):
        cand, approx, surv = O.plaid_approx_scores(oix, Q[q].T, 4, 0.4)
        assert np.all(np.diff(cand) > 0) and np.all(approx >= 0)
        # the home centroid of every query token scores ~0.9: it survives
        assert len(surv) >= 1
        _, _, sel, _, _ = O.plaid_search(oix, Q[q].T, 5, 4, 0.4, 30, return_selected=True)
        assert len(sel) == min(30, len(cand))
        worst_in = approx[np.isin(cand, sel)].min()
        best_out = approx[~np.isin(cand, sel)].max() if len(sel) < len(cand) else -1.0
        assert worst_in >= best_out
        # an unreachable threshold: all approximate scores are 0, the selection is the lowest pids
        _, _, sel0, cand0, approx0 = O.plaid_search(oix, Q[q].T, 5, 4, 2.0, 30, return_selected=True)
        assert np.all(approx0 == 0) and np.array_equal(sel0, cand0[:30])


def test_plaid_vectorized_equals_definition():
    ix = S.make_index(1200, 1024, seed=45, doclen_min=0, doclen_mean=20.0, doclen_std=15.0)     # includes empty passages
    Q = S.make_queries(ix["centroids"], 2, seed=46, nprobe=4)
    oix = oracle_index(ix)
    for q in range(2):
        for thr in (0.3, 0.4):
            p1, a1, s1 = O.plaid_approx_scores(oix, Q[q].T, 4, thr, vectorized=True)
            p0, a0, s0 = O.plaid_approx_scores(oix, Q[q].T, 4, thr, vectorized=False)
            assert np.array_equal(p1, p0) and np.array_equal(s1, s0)
            assert np.array_equal(a1, a0)          # bit for bit

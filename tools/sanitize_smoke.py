"""A small run of every kernel family of the library on a tiny index, meant to be executed under compute-sanitizer
(race detection / sanitizers).  tools/run_sanitizer.sh runs it under memcheck, racecheck and synccheck."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colbert_jl_b200 as cb  # noqa: E402
from colbert_jl_b200 import synthetic as S  # noqa: E402

ix = S.make_index(400, 2048, seed=5, doclen_mean=60, doclen_std=40, doclen_min=1, doclen_max=300)
Q = S.make_queries(ix["centroids"], 6, seed=6)
Qj = np.transpose(Q, (2, 1, 0))
cfg = cb.ColBERTConfig(dim=128, nbits=2, nprobe=2, query_maxlen=32)
with cb.Searcher(cfg, ix["centroids"].T, None, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"], ix["codes"],
                 ix["residuals"].T) as s:
    p, sc, c = s.search_batch(Qj, 5)                       # stage 1 (tcgen05), 2, 3+4 (tcgen05), 5 + exact re-score
    assert s.stat("tc_pairs") > 0
    s.retrieve(Q[0].T)
    s.probe(Qj)
    s.score_pids(Q[0].T, p[0])
    s.search_batch_plaid(Qj, 5, ncells=2, centroid_score_threshold=0.6, ndocs=50)
    s.set_option("force_generic", 1)
    s.set_option("stage1_impl", 1)
    p2, sc2, _ = s.search_batch(Qj, 5)                     # SIMT stage 1 + generic scoring kernel
    assert np.array_equal(p, p2)
    norm, raw = s.debug_tc_operand(np.arange(1, 21), int(ix["doclens"][:20].sum()))
cb.decompress(128, 2, ix["centroids"].T, ix["bucket_weights"], ix["codes"][:100], ix["residuals"][:100].T)
cb.compress(ix["centroids"].T, np.array([-0.02, 0.0, 0.02], np.float32), 128, 2, Q.reshape(-1, 128)[:64].T)
print("sanitize_smoke: ok")

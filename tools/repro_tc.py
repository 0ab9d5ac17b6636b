"""Small tcgen05-scoring-kernel run against the generic fp32 kernel (same library): for debugging A/B builds, also under compute-sanitizer.
Usage: repro_tc.py [n_passages] [n_queries] [doclen_mean] [n_centroids] [seed] [doclen_std]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colbert_jl_b200 as cb  # noqa: E402
from colbert_jl_b200 import synthetic as S  # noqa: E402

npass = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dl = float(sys.argv[3]) if len(sys.argv) > 3 else 120.0
K = int(sys.argv[4]) if len(sys.argv) > 4 else 512
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 11
kw = {"doclen_std": float(sys.argv[6])} if len(sys.argv) > 6 else {}
ix = S.make_index(npass, K, seed=seed, doclen_mean=dl, **kw)
Q = S.make_queries(ix["centroids"], nq, seed=seed + 1)
Qj = np.transpose(Q, (2, 1, 0))
cfg = cb.ColBERTConfig(dim=128, nbits=2, nprobe=2, query_maxlen=32)
with cb.Searcher(cfg, ix["centroids"].T, None, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"], ix["codes"],
                 ix["residuals"].T) as s:
    p, sc, c = s.search_batch(Qj, 10)
    print("tc_pairs", s.stat("tc_pairs"), "pairs", s.stat("pairs"), flush=True)
    s.set_option("force_generic", 1)
    p2, sc2, c2 = s.search_batch(Qj, 10)
    print("pids equal:", np.array_equal(p, p2), "counts equal:", np.array_equal(c, c2), "max score diff", float(np.abs(sc - sc2).max()))

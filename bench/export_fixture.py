"""Writes one synthetic index + queries of tests/bench shape as raw little-endian arrays (Julia shapes,
column-major = the C layouts this repo uses) so that bench/julia_baseline.jl can run the REAL
ColBERT.jl functions on exactly the inputs the CUDA path and the oracle see.
usage: python bench/export_fixture.py OUTDIR [--passages 20000] [--centroids 4096] [--queries 32] [--nbits 2]
Also writes the oracle's top-k for every query (expected.json) for a parity check on the Julia side."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from colbert_jl_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--passages", type=int, default=20000)
    ap.add_argument("--centroids", type=int, default=4096)
    ap.add_argument("--queries", type=int, default=32)
    ap.add_argument("--nbits", type=int, default=2)
    ap.add_argument("--nprobe", type=int, default=2)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)
    ix = S.make_index(a.passages, a.centroids, nbits=a.nbits, seed=1000)
    Q = S.make_queries(ix["centroids"], a.queries, seed=2001, nprobe=a.nprobe)
    arrays = {"centroids.f32": ix["centroids"],              # C [K][dim]      = Julia Matrix{Float32}(dim, K)
              "bucket_weights.f32": ix["bucket_weights"],
              "codes.u32": ix["codes"].astype(np.uint32),    # 1-based
              "residuals.u8": ix["residuals"],               # C [N_e][R]      = Julia Matrix{UInt8}(R, N_e)
              "doclens.i64": ix["doclens"].astype(np.int64),
              "ivf.i64": ix["ivf"].astype(np.int64),         # 1-based eids
              "ivf_lengths.i64": ix["ivf_lengths"].astype(np.int64),
              "queries.f32": Q}                              # C [nq][T][dim]  = Julia Array{Float32,3}(dim, T, nq)
    for name, arr in arrays.items():
        np.ascontiguousarray(arr).tofile(os.path.join(a.outdir, name))
    oix = O.Index(ix["dim"], a.nbits, ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"],
                  ix["codes"], ix["residuals"].T, nprobe=a.nprobe)
    expected = []
    for q in range(a.queries):
        p, s = O.search(oix, Q[q].T, a.k)
        expected.append({"pids": p.tolist(), "scores": [float(x) for x in s]})
    meta = {"dim": int(ix["dim"]), "nbits": a.nbits, "K": a.centroids, "N_p": a.passages, "N_e": int(ix["doclens"].sum()),
            "R": int(ix["residuals"].shape[1]), "nq": a.queries, "T": int(Q.shape[1]), "nprobe": a.nprobe, "k": a.k}
    json.dump(meta, open(os.path.join(a.outdir, "meta.json"), "w"), indent=1)
    json.dump(expected, open(os.path.join(a.outdir, "expected.json"), "w"))
    print("wrote", a.outdir, meta)


if __name__ == "__main__":
    main()

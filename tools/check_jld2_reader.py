"""Pins the native JLD2 reader against an index written by the REAL JLD2.jl (bench/julia_write_index.jl, needs Julia):
every .jld2 file of the directory must read back equal to its raw twin.  usage: python tools/check_jld2_reader.py <dir>"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colbert_jl_b200 as cb  # noqa: E402

d = sys.argv[1]
sh = json.load(open(os.path.join(d, "raw", "shapes.json")))
raw = lambda n, t: np.fromfile(os.path.join(d, "raw", n + ".bin"), dtype=t)
R = sh["dim"] // 8 * sh["nbits"]
assert np.array_equal(cb.load_object(os.path.join(d, "centroids.jld2")), raw("centroids", np.float32).reshape(sh["K"], sh["dim"]))
assert np.array_equal(cb.load_object(os.path.join(d, "bucket_weights.jld2")), raw("bucket_weights", np.float32))
assert np.array_equal(cb.load_object(os.path.join(d, "ivf.jld2")), raw("ivf", np.int64))
assert np.array_equal(cb.load_object(os.path.join(d, "ivf_lengths.jld2")), raw("ivf_lengths", np.int64))
codes = np.concatenate([cb.load_object(os.path.join(d, f"{c + 1}.codes.jld2")) for c in range(sh["n_chunks"])])
res = np.concatenate([cb.load_object(os.path.join(d, f"{c + 1}.residuals.jld2")) for c in range(sh["n_chunks"])])
dl = np.concatenate([cb.load_object(os.path.join(d, f"doclens.{c + 1}.jld2")) for c in range(sh["n_chunks"])])
assert np.array_equal(codes, raw("codes", np.uint32)) and np.array_equal(res, raw("residuals", np.uint8).reshape(-1, R))
assert np.array_equal(dl, raw("doclens", np.int64))
print("native reader == real JLD2 files: ok")
if cb.load().cb_device_count() > 0:
    with cb.Searcher.open(d) as s:
        print("cb_index_open:", s.info())

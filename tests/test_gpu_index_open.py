"""GPU: cb_index_open -- an index DIRECTORY in the reference's on-disk layout (src/savers.jl; written
here by tests/jld2_writer.py) opened natively gives the same searcher as the arrays handed over in memory, unsharded
and as passage-range shards that read only their own chunk files."""
import os

import numpy as np
import pytest

import colbert_jl_b200 as cb
from colbert_jl_b200 import sharding as SH
from colbert_jl_b200 import synthetic as S
from tests import jld2_writer as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def on_disk(tmp_path_factory):
    ix = S.make_index(2500, 512, seed=601)
    path = str(tmp_path_factory.mktemp("index"))
    W.write_index(path, ix, n_chunks=4)
    Q = S.make_queries(ix["centroids"], 12, seed=602)
    return ix, path, Q


def _from_arrays(ix, **kw):
    cfg = cb.ColBERTConfig(dim=ix["dim"], nbits=ix["nbits"], nprobe=2, query_maxlen=32)
    return cb.Searcher(cfg, ix["centroids"].T, None, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"],
                       ix["codes"], ix["residuals"].T, **kw)


def test_open_equals_in_memory_index(on_disk):
    ix, path, Q = on_disk
    Qj = np.transpose(Q, (2, 1, 0))
    with _from_arrays(ix) as a, cb.Searcher.open(path) as b:
        assert b.config.dim == 128 and b.config.nbits == ix["nbits"] and b.pid_base == 0
        ia, ib = a.info(), b.info()
        assert {k: ia[k] for k in ("dim", "nbits", "K", "n_passages", "n_embeddings")} == \
               {k: ib[k] for k in ("dim", "nbits", "K", "n_passages", "n_embeddings")}
        ra, rb = a.search_batch(Qj, 10), b.search_batch(Qj, 10)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        assert np.array_equal(a.retrieve(Q[0].T), b.retrieve(Q[0].T))


def test_open_one_shard_reads_only_its_chunks(on_disk):
    ix, path, Q = on_disk
    Qj = np.transpose(Q, (2, 1, 0))
    n = 3
    lists_p, lists_s = [], []
    with _from_arrays(ix) as whole:
        ref_p, ref_s, _ = whole.search_batch(Qj, 10)
    slices = SH.shard_slices(ix["doclens"], n)
    for r in range(n):
        with cb.Searcher.open(path, shard=r, n_shards=n) as s:
            lo, hi, e_lo, e_hi = slices[r]
            assert s.pid_base == lo and s.n_passages == hi - lo and s.n_embeddings == e_hi - e_lo   # same cuts as sharding.py
            p, sc, _ = s.search_batch(Qj, 10)
            lists_p.append(p)
            lists_s.append(sc)
    mp, ms = cb.merge_topk(np.stack(lists_p), np.stack(lists_s))
    assert np.array_equal(mp, ref_p) and np.array_equal(ms, ref_s)
    # a shard must not need the chunk files of other shards (per-rank chunk selection)
    os.rename(os.path.join(path, "4.codes.jld2"), os.path.join(path, "4.codes.jld2.away"))
    try:
        with cb.Searcher.open(path, shard=0, n_shards=n) as s:
            assert s.n_passages == slices[0][1]
        with pytest.raises(cb.DimensionMismatch):
            cb.Searcher.open(path)
    finally:
        os.rename(os.path.join(path, "4.codes.jld2.away"), os.path.join(path, "4.codes.jld2"))


def test_open_validates_like_the_reference_loaders(on_disk, tmp_path):
    ix, path, _ = on_disk
    import json
    import shutil
    bad = str(tmp_path / "bad")
    shutil.copytree(path, bad)
    plan = json.load(open(os.path.join(bad, "plan.json")))
    plan["num_embeddings"] += 1                                   # loaders.jl:86-88
    json.dump(plan, open(os.path.join(bad, "plan.json"), "w"))
    with pytest.raises(cb.DimensionMismatch, match="sum\\(doclens\\)"):
        cb.Searcher.open(bad)
    shutil.rmtree(bad)
    shutil.copytree(path, bad)
    W.save_object(os.path.join(bad, "centroids.jld2"), ix["centroids"].astype(np.float64))   # loaders.jl:27 `isa Matrix{Float32}`
    with pytest.raises(cb.DomainError, match="Matrix\\{Float32\\}"):
        cb.Searcher.open(bad)
    with pytest.raises(cb.ColBERTB200Error, match="does not exist"):
        cb.Searcher.open(str(tmp_path / "nowhere"))

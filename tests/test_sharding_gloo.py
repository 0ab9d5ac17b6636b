"""Host-side logic of the N > 1 path on CPU: shard boundaries, and the per-shard top-k exchange
over `torch.distributed` (gloo, world_size 2).  The scoring itself needs a GPU (tests -m gpu);
here every rank's "local search" is the oracle restricted to its passage range, which is exactly
what a shard computes, and the merge is the oracle's merge (the product merge kernel is covered
by test_gpu_parity.py::test_merge_topk_and_sharded_search_equals_unsharded)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import colbert_jl_b200 as cb  # noqa: E402
from colbert_jl_b200 import sharding as SH  # noqa: E402
from colbert_jl_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def test_shard_bounds_balanced_by_embeddings():
    rng = np.random.default_rng(5)
    dl = rng.integers(1, 300, size=10_000)
    cs = np.concatenate([[0], np.cumsum(dl)])
    for n in (1, 2, 3, 4, 8):
        b = SH.shard_bounds(cs, n)
        assert b[0] == 0 and b[-1] == len(dl) and len(b) == n + 1
        assert all(b[i] <= b[i + 1] for i in range(n))
        embs = [cs[b[i + 1]] - cs[b[i]] for i in range(n)]
        assert sum(embs) == cs[-1]
        assert max(embs) - min(embs) <= 2 * 300          # within one passage of perfect balance
    sl = SH.shard_slices(dl, 4)
    assert sl[0][0] == 0 and sl[-1][1] == len(dl) and sl[-1][3] == cs[-1]
    assert all(sl[i][1] == sl[i + 1][0] and sl[i][3] == sl[i + 1][2] for i in range(3))


def test_shard_bounds_degenerate():
    assert SH.shard_bounds([0], 4) == [0, 0, 0, 0, 0]                      # empty index
    assert SH.shard_bounds([0, 5], 4) == [0, 0, 1, 1, 1] or SH.shard_bounds([0, 5], 4)[-1] == 1
    b = SH.shard_bounds([0, 0, 0, 7, 7], 2)                                 # empty passages
    assert b[0] == 0 and b[-1] == 4 and b[1] in (0, 1, 2, 3, 4)
    with pytest.raises(ValueError):
        SH.shard_bounds([1, 2], 2)
    with pytest.raises(ValueError):
        SH.shard_bounds([0, 2], 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_oracle_topk(ix, lo, hi, e_lo, e_hi, Q, k):
    """What one shard returns: the oracle's search over passages [lo, hi) with pid_base = lo."""
    codes = ix["codes"][e_lo:e_hi]
    ivf, ivl = O._build_ivf(codes, ix["centroids"].shape[0])
    oix = O.Index(128, 2, ix["centroids"].T, ix["bucket_weights"], ivf, ivl, ix["doclens"][lo:hi], codes,
                  ix["residuals"][e_lo:e_hi].T, nprobe=2)
    nq = Q.shape[0]
    P = np.zeros((nq, k), dtype=np.int64)
    Sc = np.full((nq, k), -np.inf, dtype=np.float32)
    for q in range(nq):
        pids, scores = O.search_all_scores(oix, Q[q].T)
        order = np.argsort(-scores, kind="stable")[:k]
        P[q, :len(order)] = pids[order] + lo
        Sc[q, :len(order)] = scores[order]
    return P, Sc


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ix = S.make_index(600, 256, seed=31)
        Q = S.make_queries(ix["centroids"], 6, seed=32)
        k = 5
        lo, hi, e_lo, e_hi = SH.shard_slices(ix["doclens"], world)[rank]
        P, Sc = _shard_oracle_topk(ix, lo, hi, e_lo, e_hi, Q, k)

        class LocalShard:                      # stands in for the GPU Searcher of this rank
            device = -1

            def search_batch_device(self, q_ptr, nq, T, k_, p_ptr, s_ptr, c_ptr, stream=None):
                raise AssertionError("not used: the gloo test feeds the local lists directly")

        def merge(all_p, all_s, out_p, out_s):
            mp, ms = O.merge_topk(all_p.numpy(), all_s.numpy(), k)
            out_p.copy_(torch.from_numpy(mp))
            out_s.copy_(torch.from_numpy(ms))

        sh = SH.ShardedSearcher(LocalShard(), merge=merge)
        assert sh.world == world
        all_p, all_s = SH.gather_topk(torch.from_numpy(P), torch.from_numpy(Sc))
        assert tuple(all_p.shape) == (world, Q.shape[0], k)
        assert torch.equal(all_p[rank], torch.from_numpy(P))          # own slot holds own list
        out_p = torch.zeros((Q.shape[0], k), dtype=torch.int64)
        out_s = torch.zeros((Q.shape[0], k), dtype=torch.float32)
        sh.merge(all_p, all_s, out_p, out_s)
        # every rank ends with the same global first-k, equal to the unsharded oracle
        oix = O.Index(128, 2, ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"],
                      ix["codes"], ix["residuals"].T, nprobe=2)
        for q in range(Q.shape[0]):
            op, osc = O.search(oix, Q[q].T, k)
            assert np.array_equal(out_p[q].numpy(), op), (rank, q, out_p[q], op)
            np.testing.assert_allclose(out_s[q].numpy(), osc, rtol=1e-6)
        gathered = [torch.zeros_like(out_p) for _ in range(world)]
        dist.all_gather(gathered, out_p)
        assert all(torch.equal(g, out_p) for g in gathered)
        # stage 1 split by query: every rank probes its query slice, the cells are all-gathered
        nq, T, nprobe = 7, 4, 2                                    # nq not a multiple of the world size
        lo_q, hi_q, per = SH.query_slice(nq, world, rank)
        assert (lo_q, hi_q, per) == ((0, 4, 4) if rank == 0 else (4, 7, 4))
        truth = torch.arange(world * per * T * nprobe, dtype=torch.int32).view(world * per, T, nprobe)
        cells = torch.zeros_like(truth)
        cells[rank * per:(rank + 1) * per] = truth[rank * per:(rank + 1) * per]
        SH.gather_cells(cells, per, rank)
        assert torch.equal(cells, truth)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_topk_exchange_world2_gloo():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_sharded_searcher_forwards_plaid_options():
    """world = 1 (no process group): the PLAID knobs reach the shard's `search_batch_plaid_device` unchanged."""
    import torch
    calls = []

    class Shard:
        device = -1

        def search_batch_device(self, *a, **kw):
            calls.append(("exhaustive", kw))

        def search_batch_plaid_device(self, q_ptr, nq, T, k, p_ptr, s_ptr, c_ptr, **kw):
            calls.append(("plaid", nq, T, k, kw))

    sh = SH.ShardedSearcher(Shard())
    Q = torch.zeros((3, 32, 128))
    out_p, out_s, out_c = torch.zeros((3, 7), dtype=torch.int64), torch.zeros((3, 7)), torch.zeros(3, dtype=torch.int32)
    sh.search_batch_device(Q, 7, out_p, out_s, out_c)
    sh.search_batch_device(Q, 7, out_p, out_s, out_c, plaid=dict(ncells=4, centroid_score_threshold=0.4, ndocs=1000))
    assert calls[0][0] == "exhaustive"
    assert calls[1] == ("plaid", 3, 32, 7, {"stream": None, "ncells": 4, "centroid_score_threshold": 0.4, "ndocs": 1000})

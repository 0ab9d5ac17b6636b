"""GPU tests of the device-resident entry points (the encoder hand-off) and of the
sync-free batch driver: cb_search_batch_device on a caller's stream, CB_FLAG_DEVICE_POINTERS,
cb_probe_device + cb_search_batch_cells_device (the stage-1 split a sharded deployment uses), more than
CB_NQ_CHUNK queries, both pair-list sizing paths, the range gate of the tensor-core kernel, and the exact
fp32 final ranking.  torch is only the owner of device memory and streams here."""
import numpy as np
import pytest

import colbert_jl_b200 as cb
from colbert_jl_b200 import synthetic as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _searcher(ix, **kw):
    cfg = cb.ColBERTConfig(dim=ix["dim"], nbits=ix["nbits"], nprobe=2, query_maxlen=32)
    return cb.Searcher(cfg, ix["centroids"].T, None, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"],
                       ix["codes"], ix["residuals"].T, **kw)


@pytest.fixture(scope="module")
def small():
    ix = S.make_index(3000, 512, seed=501)
    Q = S.make_queries(ix["centroids"], 40, seed=502)          # [nq][T][dim]
    return ix, Q


def _device_search(s, Qc, k, stream=None, cells=None):
    dev = torch.device("cuda", 0)
    nq, T, _ = Qc.shape
    Qd = torch.from_numpy(np.ascontiguousarray(Qc)).to(dev)
    p = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    sc = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    c = torch.zeros((nq,), dtype=torch.int32, device=dev)
    h = None if stream is None else stream.cuda_stream
    if cells is None:
        s.search_batch_device(Qd.data_ptr(), nq, T, k, p.data_ptr(), sc.data_ptr(), c.data_ptr(), stream=h)
    else:
        s.search_batch_cells_device(Qd.data_ptr(), cells.data_ptr(), nq, T, k, p.data_ptr(), sc.data_ptr(), c.data_ptr(), stream=h)
    (stream.synchronize() if stream is not None else torch.cuda.synchronize())
    return p.cpu().numpy(), sc.cpu().numpy(), c.cpu().numpy()


def test_device_entry_point_equals_host_entry_point_on_a_side_stream(small):
    ix, Q = small
    with _searcher(ix) as s:
        hp, hs, hc = s.search_batch(np.transpose(Q, (2, 1, 0)), 10)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            dp, ds, dc = _device_search(s, Q, 10, stream=side)
        assert np.array_equal(hp, dp) and np.array_equal(hs, ds) and np.array_equal(hc, dc)
        assert s.stat("pairs") == float(hc.sum())               # lazy counters of the asynchronous call


def test_index_from_device_pointers(small):
    ix, Q = small
    dev = torch.device("cuda", 0)
    cen = torch.from_numpy(ix["centroids"]).to(dev)
    w = torch.from_numpy(ix["bucket_weights"]).to(dev)
    codes = torch.from_numpy(ix["codes"].astype(np.int32)).to(dev)      # UInt32 bit pattern
    res = torch.from_numpy(ix["residuals"]).to(dev)
    dl = torch.from_numpy(ix["doclens"].astype(np.int64)).to(dev)
    cfg = cb.ColBERTConfig(dim=128, nbits=ix["nbits"], nprobe=2, query_maxlen=32)
    with _searcher(ix) as ref:
        hp, hs, hc = ref.search_batch(np.transpose(Q, (2, 1, 0)), 7)
    s = cb.Searcher.from_device(cfg, cen.shape[0], dl.numel(), codes.numel(), cen.data_ptr(), w.data_ptr(), codes.data_ptr(),
                                res.data_ptr(), dl.data_ptr(), None, None, device=0)    # IVF built on the device
    try:
        dp, ds, dc = _device_search(s, Q, 7)
    finally:
        s.close()
    assert np.array_equal(hp, dp) and np.array_equal(hs, ds) and np.array_equal(hc, dc)
    # CB_FLAG_BORROW_RESIDUALS: the caller's residual array is read in place (and survives the index)
    s = cb.Searcher.from_device(cfg, cen.shape[0], dl.numel(), codes.numel(), cen.data_ptr(), w.data_ptr(), codes.data_ptr(),
                                res.data_ptr(), dl.data_ptr(), None, None, device=0, borrow_residuals=True)
    try:
        bp, bs, bc = _device_search(s, Q, 7)
    finally:
        s.close()
    assert np.array_equal(hp, bp) and np.array_equal(hs, bs) and np.array_equal(hc, bc)
    assert torch.equal(res.cpu(), torch.from_numpy(ix["residuals"]))      # untouched and still allocated


def test_more_queries_than_one_chunk(small):
    """nq = 2500 > CB_NQ_CHUNK = 1024: three chunks; every query must equal its own single-chunk result."""
    ix, _ = small
    Q = S.make_queries(ix["centroids"], 2500, seed=503)
    with _searcher(ix) as s:
        p, sc, c = _device_search(s, Q, 5)
        for lo in (0, 1000, 2040, 2400):
            p1, s1, c1 = _device_search(s, Q[lo:lo + 60], 5)
            assert np.array_equal(p[lo:lo + 60], p1) and np.array_equal(sc[lo:lo + 60], s1) and np.array_equal(c[lo:lo + 60], c1)
    oix = O.Index(128, ix["nbits"], ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"],
                  ix["codes"], ix["residuals"].T, nprobe=2)
    for q in (0, 1023, 1024, 2047, 2048, 2499):
        op, osc = O.search(oix, Q[q].T, 5)
        np.testing.assert_allclose(sc[q], osc, rtol=1e-3)
        assert np.array_equal(p[q], op)


def test_cells_supplied_by_the_caller(small):
    """cb_probe_device on two query halves (what two ranks would do) + cb_search_batch_cells_device
    == cb_search_batch_device, bit for bit."""
    ix, Q = small
    dev = torch.device("cuda", 0)
    nq, T, _ = Q.shape
    with _searcher(ix) as s:
        ref = _device_search(s, Q, 10)
        Qd = torch.from_numpy(np.ascontiguousarray(Q)).to(dev)
        cells = torch.zeros((nq, T, 2), dtype=torch.int32, device=dev)
        half = nq // 2
        s.probe_device(Qd.data_ptr(), half, T, cells.data_ptr())
        s.probe_device(Qd[half:].contiguous().data_ptr(), nq - half, T, cells[half:].data_ptr())
        torch.cuda.synchronize()
        hc, _ = s.probe(np.transpose(Q, (2, 1, 0)))
        assert np.array_equal(cells.cpu().numpy(), hc)          # 1-based, same as the host hook
        got = _device_search(s, Q, 10, cells=cells)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)
        assert s.stat("bad_cells") == 0
        cells[0, 0, 0] = 10 ** 6                                # out of range: ignored and counted, never a fault
        _device_search(s, Q, 10, cells=cells)
        assert s.stat("bad_cells") == 1


def test_pair_list_sizing_paths_agree(small):
    ix, Q = small
    with _searcher(ix) as s:
        a = _device_search(s, Q, 10)
        s.set_option("sync_pairs", 1)
        b = _device_search(s, Q, 10)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert s.stat("max_cell_len") == float(np.max(ix["ivf_lengths"]))


def test_final_ranking_is_on_exact_fp32_scores(small):
    """The returned scores are the exact fp32 ones (== cb_score_pids bit for bit, == oracle within fp32
    noise), not the fp16-operand tensor-core scores; with the re-score switched off they are the latter."""
    ix, Q = small
    oix = O.Index(128, ix["nbits"], ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"],
                  ix["codes"], ix["residuals"].T, nprobe=2)
    with _searcher(ix) as s:
        p, sc, c = _device_search(s, Q, 10)
        assert s.stat("tc_pairs") > 0 and s.stat("rescore_unsafe") == 0
        for q in range(0, Q.shape[0], 7):
            assert np.array_equal(s.score_pids(Q[q].T, p[q]), sc[q])
            op, osc = O.search(oix, Q[q].T, 10)
            assert np.array_equal(p[q], op)
            np.testing.assert_allclose(sc[q], osc, rtol=2e-5)
        s.set_option("exact_rescore", 0)
        p0, sc0, _ = _device_search(s, Q, 10)
        np.testing.assert_allclose(sc0, sc, rtol=1e-3)
        assert not np.array_equal(sc0, sc)


def test_out_of_range_queries_are_routed_to_the_fp32_kernel(small):
    """the tensor-core kernel's fixed-point token sum needs |query token| <= 255.  A batch that
    breaks it is scored by the generic kernel (device-side gate), with the scores the oracle gets."""
    ix, Q = small
    Qs = (Q[:6] * np.float32(300.0)).astype(np.float32)        # same directions -> same cells and candidates
    oix = O.Index(128, ix["nbits"], ix["centroids"].T, ix["bucket_weights"], ix["ivf"], ix["ivf_lengths"], ix["doclens"],
                  ix["codes"], ix["residuals"].T, nprobe=2)
    with _searcher(ix) as s:
        p, sc, c = _device_search(s, Qs, 10)
        assert s.stat("tc_pairs") == 0 and s.stat("generic_pairs") == float(c.sum())
        for q in range(6):
            op, osc = O.search(oix, Qs[q].T, 10)
            assert np.array_equal(p[q], op)
            np.testing.assert_allclose(sc[q], osc, rtol=1e-4)
        _device_search(s, Q[:6], 10)                           # and the gate re-opens for the next batch
        assert s.stat("tc_pairs") > 0


# --------------------------------------------------------------------------------------------
# several shards behind the C ABI, one process, one host thread (cb_multi_*)
# --------------------------------------------------------------------------------------------
def _multi_equals_unsharded(ix, Q, devices):
    from colbert_jl_b200 import sharding as SH
    Qj = np.transpose(Q, (2, 1, 0))
    with _searcher(ix) as whole:
        ref = whole.search_batch(Qj, 10)
    n = len(devices)
    shards = []
    for r, (lo, hi, e_lo, e_hi) in enumerate(SH.shard_slices(ix["doclens"], n)):
        part = S.take_shard(ix, lo, hi)
        cfg = cb.ColBERTConfig(dim=128, nbits=ix["nbits"], nprobe=2, query_maxlen=32)
        shards.append(cb.Searcher(cfg, part["centroids"].T, None, part["bucket_weights"], None, None, part["doclens"], part["codes"],
                                  part["residuals"].T, device=devices[r], pid_base=lo))
    try:
        with cb.MultiSearcher(shards) as m:
            for _ in range(2):                                  # twice: workspaces are reused across batches
                got = m.search_batch(Qj, 10)
                for a, b in zip(ref, got):
                    assert np.array_equal(a, b)
            got5 = m.search_batch(Qj[:, :, :5], 10)             # nq not a multiple of the shard count
            assert np.array_equal(got5[0], ref[0][:5]) and np.array_equal(got5[1], ref[1][:5])
    finally:
        for s in shards:
            s.close()


def test_multi_three_shards_on_one_device_equal_unsharded(small):
    ix, Q = small
    _multi_equals_unsharded(ix, Q, [0, 0, 0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multi_two_gpus_equal_unsharded(small):
    ix, Q = small
    _multi_equals_unsharded(ix, Q, [0, 1])


def test_multi_open_index_directory(small, tmp_path):
    from tests import jld2_writer as W
    ix, Q = small
    path = str(tmp_path / "ix")
    W.write_index(path, ix, n_chunks=5)
    Qj = np.transpose(Q, (2, 1, 0))
    with _searcher(ix) as whole:
        ref = whole.search_batch(Qj, 10)
    ndev = min(2, torch.cuda.device_count())
    with cb.MultiSearcher.open(path, 2, device_ids=[0, ndev - 1]) as m:
        got = m.search_batch(Qj, 10)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)

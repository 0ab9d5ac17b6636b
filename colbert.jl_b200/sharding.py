"""Passage-range sharding of one index over the GPUs of a box (one process per GPU).

The reference is single-process, single-device (src/infra/config.jl:57-58: `rank`/`nranks` exist
but are not used at search time).  Passages are independent in every stage after stage 1, so the
index shards by contiguous passage range, balanced by embedding count; each rank scores its own
shard for the same queries.  Stage 1 (query tokens x centroids) depends on the queries only and every
shard holds the same centroids, so it is split by QUERY instead: rank r probes queries
[r nq/N, (r+1) nq/N) and the ranks all-gather the cells (nq * T * nprobe int32 = 256 KB per batch);
replicated on every rank it was a third of the 8-GPU step (round 1).  The other exchange is the per-shard
top-k lists -- nq * k * 12 bytes per rank -- all-gathered and merged by (score desc, pid asc), the order of
the reference's stable `sortperm` over ascending pids (src/searching.jl:125-127).

This module is host-side plumbing only: the scoring and the merge run in libcolbert_b200.so; the
exchange is `torch.distributed` (NCCL over NVLink on GPUs; the same code runs on the `gloo` backend
with CPU tensors in the world_size-2 tests, with the merge injected by the test).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(doclens_cumsum, n_shards: int):
    """Passage boundaries [b_0 = 0, b_1, ..., b_n = N_p]: shard r owns passages b_r .. b_{r+1}-1.
    `doclens_cumsum` is the INCLUSIVE prefix sum of doclens with a leading 0 (length N_p + 1).
    Cut r is the first passage boundary whose embedding offset reaches r/n of all embeddings, so
    shards are balanced by embedding count (= by bytes and by scoring work), not by passages."""
    cs = np.asarray(doclens_cumsum, dtype=np.int64)
    if cs.ndim != 1 or len(cs) < 1 or cs[0] != 0:
        raise ValueError("doclens_cumsum must be 1-D, start with 0 and have length N_p + 1")
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    total = int(cs[-1])
    cuts = [int(np.searchsorted(cs, total * r // n_shards, side="left")) for r in range(1, n_shards)]
    b = [0] + cuts + [len(cs) - 1]
    for i in range(1, len(b)):            # monotone even for degenerate inputs (empty passages, n > N_p)
        b[i] = max(b[i], b[i - 1])
    return b


def shard_slices(doclens, n_shards: int):
    """[(p_lo, p_hi, e_lo, e_hi)] per shard: passage range and embedding range (0-based, half open).
    `p_lo` is the `pid_base` of the shard (cb_index_create)."""
    dl = np.asarray(doclens, dtype=np.int64)
    cs = np.zeros(len(dl) + 1, dtype=np.int64)
    np.cumsum(dl, out=cs[1:])
    b = shard_bounds(cs, n_shards)
    return [(b[r], b[r + 1], int(cs[b[r]]), int(cs[b[r + 1]])) for r in range(n_shards)]


def gather_topk(pids, scores, group=None):
    """All-gathers the per-shard result lists: pids int64 [nq][k], scores float32 [nq][k] (torch
    tensors on this rank's device) -> ([world][nq][k], [world][nq][k]).  The only data-path
    collective of the search."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nq, k = pids.shape
    all_p = torch.empty((world * nq, k), dtype=pids.dtype, device=pids.device)      # rank-major concatenation
    all_s = torch.empty((world * nq, k), dtype=scores.dtype, device=scores.device)
    dist.all_gather_into_tensor(all_p, pids.contiguous(), group=group)
    dist.all_gather_into_tensor(all_s, scores.contiguous(), group=group)
    return all_p.view(world, nq, k), all_s.view(world, nq, k)


def query_slice(nq: int, world: int, rank: int):
    """Queries [lo, hi) whose stage 1 rank `rank` computes, and the common padded slice length."""
    per = (nq + world - 1) // world
    lo = min(nq, rank * per)
    return lo, min(nq, lo + per), per


def gather_cells(cells_all, per: int, rank: int, group=None):
    """All-gathers the stage-1 cells: `cells_all` int32 [world * per][T][nprobe] whose rows
    [rank * per, (rank + 1) * per) this rank has filled; on return every rank holds all rows."""
    import torch.distributed as dist
    mine = cells_all[rank * per:(rank + 1) * per]
    dist.all_gather_into_tensor(cells_all.view(-1), mine.reshape(-1).clone() if cells_all.device.type == "cpu" else mine.view(-1), group=group)
    return cells_all


class ShardedSearcher:
    """One rank's view of a passage-sharded index: `searcher` holds this rank's shard (created with
    pid_base = first passage of the shard, so its pids are already global)."""

    def __init__(self, searcher, group=None, merge=None, shard_stage1=True):
        import torch.distributed as dist
        self.searcher = searcher
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._merge = merge   # injected by the CPU (gloo) tests; None = cb_merge_topk_device
        self.shard_stage1 = shard_stage1
        self._cells = None

    def merge(self, all_p, all_s, out_p, out_s, stream=None):
        if self._merge is not None:
            return self._merge(all_p, all_s, out_p, out_s)
        from . import _lib as L
        n, nq, k = all_p.shape
        L.check(L.load().cb_merge_topk_device(self.searcher.device, n, nq, k, all_p.data_ptr(), all_s.data_ptr(),
                                              out_p.data_ptr(), out_s.data_ptr(), stream))

    def probe_and_gather(self, Qd, stream=None):
        """Stage 1 split by query: this rank probes queries [rank * per, (rank + 1) * per) and the ranks all-gather the
        cells (int32 [world * per][T][nprobe], 1-based); every rank returns all rows."""
        import torch
        nq, T, _ = Qd.shape
        nprobe = self.searcher.config.nprobe
        lo, hi, per = query_slice(nq, self.world, self.rank)
        if self._cells is None or tuple(self._cells.shape) != (self.world * per, T, nprobe):
            self._cells = torch.zeros((self.world * per, T, nprobe), dtype=torch.int32, device=Qd.device)
        if hi > lo:
            self.searcher.probe_device(Qd[lo:hi].data_ptr(), hi - lo, T, self._cells[lo:hi].data_ptr(), stream=stream)
        gather_cells(self._cells, per, self.rank, self.group)
        return self._cells

    def search_batch_device(self, Qd, k, out_p, out_s, out_c, stream=None, local_p=None, local_s=None, plaid=None):
        """Qd float32 [nq][T][dim] on this rank's GPU (identical on every rank); out_p / out_s
        [nq][k] receive the GLOBAL first-k on every rank; out_c [nq] the local candidate counts.

        `plaid` = dict(ncells=, centroid_score_threshold=, ndocs=) runs the PLAID-style pruned search
        (`cb_search_batch_plaid_device`) on every shard instead.  Each shard then keeps ITS first
        `ndocs` candidates, so the union re-scored exactly is a superset of what one unsharded index
        would select (never a worse top-k); `out_c` holds the local re-scored counts."""
        import torch
        nq, T, _ = Qd.shape

        def local(p, s_):
            if plaid is None and self.world > 1 and self.shard_stage1:
                # stage 1 split by query: probe my slice, all-gather the cells, search with the cells given
                cells = self.probe_and_gather(Qd, stream)
                self.searcher.search_batch_cells_device(Qd.data_ptr(), cells.data_ptr(), nq, T, k, p.data_ptr(),
                                                        s_.data_ptr(), out_c.data_ptr(), stream=stream)
            elif plaid is None:
                self.searcher.search_batch_device(Qd.data_ptr(), nq, T, k, p.data_ptr(), s_.data_ptr(), out_c.data_ptr(),
                                                  stream=stream)
            else:
                self.searcher.search_batch_plaid_device(Qd.data_ptr(), nq, T, k, p.data_ptr(), s_.data_ptr(), out_c.data_ptr(),
                                                        stream=stream, **plaid)

        if self.world == 1:
            local(out_p, out_s)
            return
        lp = local_p if local_p is not None else torch.empty_like(out_p)
        ls = local_s if local_s is not None else torch.empty_like(out_s)
        local(lp, ls)
        all_p, all_s = gather_topk(lp, ls, self.group)
        self.merge(all_p, all_s, out_p, out_s, stream)

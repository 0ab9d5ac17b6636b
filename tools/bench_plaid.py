"""BASELINE.json config 5 on one B200: PLAID-style deep probe (ncells = 4, centroid_score_threshold = 0.4,
top-1000 candidates fully MaxSim-reranked per query, k = 100) over the synthetic workloads of bench.py.
Prints one JSON line: throughput (CUDA events), the exhaustive search with nprobe = ncells next to it,
pruning statistics, and a parity gate against oracle.plaid_search on a few queries.
usage: python tools/bench_plaid.py [--workload C|B|S] [--steps 3] [--warmup 2] [--parity-queries 2]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C", choices=list(B.WORKLOADS))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--ncells", type=int, default=4)
    ap.add_argument("--threshold", type=float, default=0.4)
    ap.add_argument("--ndocs", type=int, default=1000)
    ap.add_argument("--nbits", type=int, default=2)
    ap.add_argument("--parity-queries", type=int, default=2)
    ap.add_argument("--skip-exhaustive", action="store_true")
    args = ap.parse_args()
    import torch
    import colbert_jl_b200 as cb
    from oracle import oracle as O
    wl = B.WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    T, dim, nq, k = 32, 128, args.nq, args.k
    cen, doclens, csum = B.gen_global(torch, wl, dev)
    codes, res = B.gen_shard(torch, wl, dev, csum, 0, wl["passages"], args.nbits)
    Qd = B.gen_queries(torch, cen, nq, T, args.ncells, dev)
    w = torch.from_numpy(B.bucket_weights(args.nbits)).to(dev)
    cfg = cb.ColBERTConfig(dim=dim, nbits=args.nbits, nprobe=args.ncells, query_maxlen=T)
    s = cb.Searcher.from_device(cfg, wl["K"], wl["passages"], codes.numel(), cen.data_ptr(), w.data_ptr(), codes.data_ptr(),
                                res.data_ptr(), doclens.contiguous().data_ptr(), None, None, device=0, pid_base=0)
    out_p = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    out_s = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    out_c = torch.zeros((nq,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    def plaid():
        s.search_batch_plaid_device(Qd.data_ptr(), nq, T, k, out_p.data_ptr(), out_s.data_ptr(), out_c.data_ptr(),
                                    ncells=args.ncells, centroid_score_threshold=args.threshold, ndocs=args.ndocs, stream=stream)

    def exhaustive():
        s.search_batch_device(Qd.data_ptr(), nq, T, k, out_p.data_ptr(), out_s.data_ptr(), out_c.data_ptr(), stream=stream,
                              nprobe=args.ncells)

    ms_ex = None if args.skip_exhaustive else timed(exhaustive)
    pairs_ex = None if args.skip_exhaustive else s.stat("pairs")
    ms_pl = timed(plaid)
    stats = {"candidate_pairs": s.stat("plaid_candidates"), "surviving_query_centroid_pairs": s.stat("plaid_survivors"),
             "exactly_rescored_pairs": s.stat("plaid_rescored"), "launches_per_step": int(s.stat("launches"))}
    gp, gs, gc = out_p.cpu().numpy(), out_s.cpu().numpy(), out_c.cpu().numpy()
    parity = None
    if args.parity_queries > 0:
        oix = B.host_oracle_index(torch, wl, args.nbits, cen, doclens, codes, res, args.ncells)
        Qh = Qd[:args.parity_queries].cpu().numpy()
        t0 = time.perf_counter()
        ok, ok_set, ok_sel, ties_only, max_rel, nswap = True, True, True, True, 0.0, 0
        for q in range(args.parity_queries):
            op, osc, sel, cand, approx = O.plaid_search(oix, Qh[q].T, k, args.ncells, args.threshold, args.ndocs, return_selected=True)
            kk = len(op)
            ok &= bool(np.array_equal(gp[q, :kk], op))
            ok_set &= set(gp[q, :kk].tolist()) == set(op.tolist())
            ok_sel &= bool(gc[q] == len(sel)) and set(gp[q, :kk].tolist()) <= set(sel.tolist())
            max_rel = max(max_rel, float(np.max(np.abs(gs[q, :kk] - osc) / np.abs(osc))))
            # positions where the order differs must be ties inside the tolerance (north_star)
            oscore = dict(zip(op.tolist(), osc.tolist()))
            for i in np.nonzero(gp[q, :kk] != op)[0]:
                nswap += 1
                a = oscore.get(int(gp[q, i]))
                ties_only &= a is not None and abs(a - float(osc[i])) <= 1e-3 * abs(float(osc[i]))
        parity = {"queries_checked": args.parity_queries, "topk_pids_identical": ok, "topk_sets_identical": ok_set,
                  "within_oracle_selection": ok_sel, "order_differences": nswap, "differences_are_ties_within_tolerance": ties_only,
                  "max_rel_score_err": max_rel, "tolerance": 1e-3,
                  "oracle_s_per_query": (time.perf_counter() - t0) / args.parity_queries}
    print(json.dumps({"metric": "queries/sec (PLAID-style deep probe, BASELINE config 5)", "value": nq / (ms_pl * 1e-3),
                      "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_pl,
                      "config": {"workload": wl["name"], "ncells": args.ncells, "centroid_score_threshold": args.threshold,
                                 "ndocs": args.ndocs, "k": k, "queries_per_step": nq, "nbits": args.nbits},
                      "exhaustive_same_ncells": None if ms_ex is None else
                      {"ms_per_step": ms_ex, "value": nq / (ms_ex * 1e-3), "pairs_per_step": pairs_ex},
                      "plaid": stats, "parity": parity}))


if __name__ == "__main__":
    main()

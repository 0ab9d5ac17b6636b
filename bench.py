#!/usr/bin/env python
"""bench.py -- queries/sec of the search-time scoring path on a synthetic index of the shape
BASELINE.json names, on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C|B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one 1024-query batch (32 tokens x dim 128, k = 10, nprobe = 2) through stages 1-5
against the WHOLE index.  For N > 1 the index is passage-sharded (contiguous ranges balanced by
embedding count); every rank scores its shard for the same queries and only the per-shard top-k
lists are exchanged (NCCL all-gather) and merged -- total work is fixed, so scaling is "strong".
Rank 0 prints ONE JSON line.  `--impl reference` times the CPU restatement of ColBERT.jl's own
search (oracle/oracle.py; Julia is not installed in this image) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (fits one B200: ~26 GB)
    "C": dict(passages=8_800_000, K=1 << 18, mean=68.0, std=25.0,
              name="synthetic MS MARCO-scale 8.8M-passage index (~600M embeddings, 2^18 centroids)"),
    # BASELINE.json configs[1]
    "B": dict(passages=1_000_000, K=1 << 16, mean=120.0, std=40.0,
              name="synthetic 1M-passage index (~120 tok/passage, 2^16 centroids)"),
    # small smoke size for development
    "S": dict(passages=100_000, K=1 << 14, mean=120.0, std=40.0, name="synthetic 100k-passage dev index"),
}
BLOCK = 32768  # passages per generation block: the global index is identical for every N


EXTRA_CONFIGS = ["B", "C_nbits1", "C_nbits4", "C_clustered", "C_plaid"]   # the other BASELINE.json configs, N = 1 only


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C", choices=sorted(WORKLOADS))
    ap.add_argument("--nbits", type=int, default=2)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nprobe", type=int, default=2)
    ap.add_argument("--cpu-queries", type=int, default=3, help="queries timed for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-generic", action="store_true", help="score with the generic SIMT kernel only")
    ap.add_argument("--opt", action="append", default=[], help="library tuning knob key=value (cb_set_option), repeatable")
    ap.add_argument("--profile", default="uniform", choices=["uniform", "clustered"], help="code profile of the synthetic index (SURVEY 8d)")
    ap.add_argument("--no-extra", action="store_true", help="skip the legs for the other BASELINE configs (N = 1)")
    ap.add_argument("--extra", default=",".join(EXTRA_CONFIGS), help="comma-separated subset of " + ",".join(EXTRA_CONFIGS))
    ap.add_argument("--extra-steps", type=int, default=5)
    ap.add_argument("--extra-workload", default=None, choices=sorted(WORKLOADS), help="(debug) run the extra legs on this workload size")
    ap.add_argument("--no-gate", action="store_true", help="skip the all-query parity gate")
    ap.add_argument("--ablation", action="store_true", help="the library is a measurement-only ablation build: results are garbage, skip the checks")
    ap.add_argument("--single-process", action="store_true",
                    help="N > 1 without torchrun: ONE process / ONE host thread drives all GPUs through cb_multi_search_batch "
                         "(the drop-in shape: the reference is single-process)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic index, generated directly in HBM (SURVEY.md 8d distribution, torch Philox streams)
# ------------------------------------------------------------------------------------------------
def gen_global(torch, wl, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(1001)
    cen = torch.randn((wl["K"], 128), generator=g, device=dev, dtype=torch.float32)
    cen /= cen.norm(dim=1, keepdim=True)
    g.manual_seed(1002)
    dl = torch.randn((wl["passages"],), generator=g, device=dev, dtype=torch.float32) * wl["std"] + wl["mean"]
    doclens = dl.round().clamp_(8, 300).to(torch.int64)
    csum = torch.zeros(wl["passages"] + 1, dtype=torch.int64, device=dev)
    csum[1:] = torch.cumsum(doclens, 0)
    return cen, doclens, csum


def shard_bounds(torch, csum, n):
    """Contiguous passage ranges balanced by embedding count (colbert.jl_b200/sharding.py)."""
    from colbert_jl_b200 import sharding as SH
    return SH.shard_bounds(csum.cpu().numpy(), n)


def gen_shard(torch, wl, dev, csum, lo, hi, nbits, profile="uniform"):
    """codes (int32 holding 1-based ids) and residual bytes of passages [lo, hi).
    profile "clustered" (secondary): each passage draws 8 home centroids and 85 % of its tokens
    come from them (mimics the fan-out of a real collection); "uniform" is the worst-case headline."""
    R = 128 // 8 * nbits
    e_lo, e_hi = int(csum[lo]), int(csum[hi])
    codes = torch.empty(e_hi - e_lo, dtype=torch.int32, device=dev)
    res = torch.empty((e_hi - e_lo, R), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    for b in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
        p0, p1 = b * BLOCK, min((b + 1) * BLOCK, csum.numel() - 1)
        b0, b1 = int(csum[p0]), int(csum[p1])
        g.manual_seed(1003_000_000 + b)
        cb_ = torch.randint(1, wl["K"] + 1, (b1 - b0,), generator=g, device=dev, dtype=torch.int32)
        if profile == "clustered":
            npass = p1 - p0
            home = torch.randint(1, wl["K"] + 1, (npass, 8), generator=g, device=dev, dtype=torch.int32)
            pid_of = torch.repeat_interleave(torch.arange(npass, device=dev), csum[p0 + 1:p1 + 1] - csum[p0:p1])
            pick = home[pid_of, torch.randint(0, 8, (b1 - b0,), generator=g, device=dev)]
            cb_ = torch.where(torch.rand((b1 - b0,), generator=g, device=dev) < 0.85, pick, cb_)
            del home, pid_of, pick
        g.manual_seed(1004_000_000 + b)
        rb = torch.randint(0, 256, (b1 - b0, R), generator=g, device=dev, dtype=torch.uint8)
        s0, s1 = max(b0, e_lo), min(b1, e_hi)
        if s1 > s0:
            codes[s0 - e_lo:s1 - e_lo] = cb_[s0 - b0:s1 - b0]
            res[s0 - e_lo:s1 - e_lo] = rb[s0 - b0:s1 - b0]
        del cb_, rb
    return codes, res


def gen_queries(torch, cen, nq, T, nprobe, dev, seed=2001, min_gap=1e-4):
    """token = normalise(centroid[c] + 0.5 g / sqrt(dim)); rows whose nprobe-th / (nprobe+1)-th
    centroid-score gap is below min_gap are regenerated (probed cell set well defined)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    K, dim = cen.shape
    Q = torch.empty((nq * T, dim), device=dev, dtype=torch.float32)
    todo = torch.arange(nq * T, device=dev)
    for _ in range(20):
        if todo.numel() == 0:
            break
        c = torch.randint(0, K, (todo.numel(),), generator=g, device=dev)
        v = cen[c] + (0.5 / dim ** 0.5) * torch.randn((todo.numel(), dim), generator=g, device=dev)
        Q[todo] = v / v.norm(dim=1, keepdim=True)
        bad = []
        for s in range(0, todo.numel(), 1024):
            rows = todo[s:s + 1024]
            top = torch.topk(Q[rows] @ cen.T, nprobe + 1, dim=1).values
            bad.append(rows[(top[:, nprobe - 1] - top[:, nprobe]) < min_gap])
        todo = torch.cat(bad)
    assert todo.numel() == 0, "could not generate well-separated queries"
    return Q.reshape(nq, T, dim).contiguous()


def bucket_weights(nbits):
    from colbert_jl_b200 import synthetic as S
    return S.bucket_weights(nbits)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
def host_oracle_index(torch, wl, nbits, cen, doclens, codes, res, nprobe):
    """Pulls one (unsharded) index to the host in the Julia shapes the oracle takes."""
    from oracle import oracle as O
    codes_d = codes.to(torch.int64)
    order = torch.sort(codes_d, stable=True).indices       # `_build_ivf`: sortperm(codes), stable
    ivf = (order + 1).cpu().numpy()
    ivf_lengths = torch.bincount(codes_d, minlength=wl["K"] + 1)[1:].cpu().numpy()
    del codes_d, order
    return O.Index(128, nbits, cen.cpu().numpy().T, bucket_weights(nbits), ivf, ivf_lengths, doclens.cpu().numpy(),
                   codes.cpu().numpy().view(np.uint32), res.cpu().numpy().T, nprobe=nprobe)


def time_oracle(oix, Qh, k, n_queries):
    from oracle import oracle as O
    results, t0 = [], time.perf_counter()
    for q in range(n_queries):
        results.append(O.search(oix, Qh[q].T, k))
    return (time.perf_counter() - t0) / n_queries, results


_JSON_OUT = None


def emit(line):
    """The ONE JSON line of the run, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def digest(*tensors):
    """sha1 over the raw bytes of the result tensors: equal digests <=> bit-identical results."""
    import hashlib
    h = hashlib.sha1()
    for t in tensors:
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def all_query_gate(torch, s, Qd, out_p, out_s, out_c, k, nprobe, lo, hi, full=True):
    """Parity gate over ALL queries of the batch (seconds, not the 15 s/query of the CPU oracle):
      * every returned (pid, score) whose passage lives on this shard is re-scored by cb_score_pids -- the exact
        fp32 generic kernel, itself oracle-checked in tests/ -- and must match bit for bit;
      * order: scores non-increasing, equal scores in ascending pid (the reference's stable sortperm)."""
    P, Sc = out_p.cpu().numpy(), out_s.cpu().numpy()
    Qh = Qd.cpu().numpy()
    nq = P.shape[0]
    bad_score = bad_order = checked = 0
    step = 1 if full else 16
    for q in range(0, nq, step):
        mine = (P[q] > lo) & (P[q] <= hi)
        if mine.any():
            ex = s.score_pids(Qh[q].T, P[q][mine])
            checked += int(mine.sum())
            bad_score += int(np.sum(ex != Sc[q][mine]))
        valid = P[q] > 0
        sc, pp = Sc[q][valid], P[q][valid]
        if len(sc) > 1:
            d = np.diff(sc)
            bad_order += int(np.sum(d > 0) + np.sum((d == 0) & (np.diff(pp) < 0)))
    return {"queries": len(range(0, nq, step)), "pairs_rescored_exact_fp32": checked, "score_mismatches": bad_score,
            "order_violations": bad_order}


def retrieve_count_gate(s, Qd, out_c, nprobe, every):
    """cb_retrieve (stages 1+2 for one query, bit-exact vs the oracle in tests/) must give the batch's candidate counts."""
    import ctypes as C
    import colbert_jl_b200 as cb
    lib = cb.load()
    Qh, C_ = Qd.cpu().numpy(), out_c.cpu().numpy()
    bad = n = 0
    cnt = C.c_int64()
    for q in range(0, Qh.shape[0], every):
        qc = np.ascontiguousarray(Qh[q])
        cb._lib.check(lib.cb_retrieve(s._h, qc.ctypes.data_as(C.c_void_p), Qh.shape[1], nprobe, None, 0, C.byref(cnt)))
        bad += int(cnt.value != C_[q])
        n += 1
    return {"queries": n, "count_mismatches": bad}


class Leg:
    """One workload on this rank's shard: index generation, searcher, timed steps, per-stage profile, gates."""

    def __init__(self, torch, cb, args, wl_key, nbits, profile, nprobe, k, rank, world, local, dist):
        self.torch, self.cb, self.args = torch, cb, args
        self.wl, self.nbits, self.profile, self.nprobe, self.k = WORKLOADS[wl_key], nbits, profile, nprobe, k
        self.rank, self.world, self.dist = rank, world, dist
        self.dev = torch.device("cuda", local)
        self.local = local
        self.sm_mhz = None       # SM clock sampled during the timed region (denominator of the shared-memory data-pipe roofline)
        wl = self.wl
        self.cen, self.doclens, self.csum = gen_global(torch, wl, self.dev)
        bounds = shard_bounds(torch, self.csum, world)
        self.lo, self.hi = bounds[rank], bounds[rank + 1]
        self.codes, self.res = gen_shard(torch, wl, self.dev, self.csum, self.lo, self.hi, nbits, profile)
        self.Qd = gen_queries(torch, self.cen, args.nq, 32, nprobe, self.dev)
        self.w = torch.from_numpy(bucket_weights(nbits)).to(self.dev)
        dl = self.doclens[self.lo:self.hi].contiguous()
        self.ne_local = self.codes.numel()
        cfg = cb.ColBERTConfig(dim=128, nbits=nbits, nprobe=nprobe, query_maxlen=32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.s = cb.Searcher.from_device(cfg, wl["K"], self.hi - self.lo, self.codes.numel(), self.cen.data_ptr(), self.w.data_ptr(),
                                         self.codes.data_ptr(), self.res.data_ptr(), dl.data_ptr(), None, None, device=local,
                                         pid_base=self.lo)
        self.t_build = time.perf_counter() - t0
        if args.force_generic:
            self.s.set_option("force_generic", 1)
        for kv in args.opt:
            key, val = kv.split("=")
            self.s.set_option(key, int(val))
        nq = args.nq
        self.out_p = torch.zeros((nq, k), dtype=torch.int64, device=self.dev)
        self.out_s = torch.zeros((nq, k), dtype=torch.float32, device=self.dev)
        self.out_c = torch.zeros((nq,), dtype=torch.int32, device=self.dev)
        from colbert_jl_b200 import sharding as SH
        self.sharded = SH.ShardedSearcher(self.s)
        self.loc_p, self.loc_s = torch.zeros_like(self.out_p), torch.zeros_like(self.out_s)
        self.stream = torch.cuda.current_stream().cuda_stream

    def drop_host_copies(self):
        self.codes = self.res = None
        self.torch.cuda.empty_cache()

    def close(self):
        self.s.close()
        for a in ("cen", "doclens", "csum", "codes", "res", "Qd", "out_p", "out_s", "out_c", "loc_p", "loc_s", "sharded", "s"):
            setattr(self, a, None)
        self.torch.cuda.empty_cache()

    def step(self, plaid=None, Q=None):
        self.sharded.search_batch_device(self.Qd if Q is None else Q, self.k, self.out_p, self.out_s, self.out_c, stream=self.stream,
                                         local_p=self.loc_p, local_s=self.loc_s, plaid=plaid)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, steps, warmup, plaid=None, sample_clocks=False):
        torch = self.torch
        for _ in range(warmup):
            self.step(plaid)
        self.barrier()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(self.local)
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step(plaid)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        clocks = sampler.stop() if sampler else None
        if clocks and clocks.get("sm_mhz"):
            self.sm_mhz = float(clocks["sm_mhz"])
        return float(ms.item()) / steps, clocks

    def stage_profile(self, nprof=3):
        """per-stage CUDA events on the launching stream (cb_set_option("profile")), this rank's shard"""
        s, torch = self.s, self.torch
        s.set_option("profile", 1)
        prof = {"ms_stage1": 0.0, "ms_stage2": 0.0, "ms_stage34": 0.0, "ms_stage5": 0.0, "ms_total": 0.0}
        for it in range(nprof + 1):      # (the first call is a warm-up)
            if self.world > 1 and self.sharded.shard_stage1:
                # the step as the timed region runs it: stage 1 on this rank's query slice + the all-gather of the cells
                # (torch events on the launching stream), then stages 2-5 with the cells given
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                cells = self.sharded.probe_and_gather(self.Qd, self.stream)
                e1.record()
                s.search_batch_cells_device(self.Qd.data_ptr(), cells.data_ptr(), self.args.nq, 32, self.k, self.loc_p.data_ptr(),
                                            self.loc_s.data_ptr(), self.out_c.data_ptr(), stream=self.stream)
                torch.cuda.synchronize()
                ms1 = e0.elapsed_time(e1)
            else:
                s.search_batch_device(self.Qd.data_ptr(), self.args.nq, 32, self.k, self.loc_p.data_ptr(), self.loc_s.data_ptr(),
                                      self.out_c.data_ptr(), stream=self.stream)
                torch.cuda.synchronize()
                ms1 = None
            if it == 0:
                continue
            for key in prof:
                v = s.stat(key)
                if ms1 is not None and key == "ms_stage1":
                    v = ms1
                elif ms1 is not None and key == "ms_total":
                    v = v + ms1
                prof[key] += v / nprof
        s.set_option("profile", 0)
        return prof

    def roofline(self, prof, pairs, pair_embs, workload_key):
        """HBM-equivalent roofline of the dominant kernel (algorithmic bytes / kernel time / measured copy
        bandwidth) first; the tensor and L2 -> SM views of the same launch beside it."""
        args = self.args
        pk = peaks()
        R = 128 // 8 * self.nbits
        T, dim, nq, k = 32, 128, args.nq, self.k
        flops = 2.0 * T * dim * pair_embs                      # 8192 flop per (query, candidate embedding)
        alg_bytes = pair_embs * (4 + R) + pairs * 16 + nq * k * 12
        t34 = prof["ms_stage34"] * 1e-3
        gbs = alg_bytes / t34 / 1e9 if t34 > 0 else 0.0
        tf = flops / t34 / 1e12 if t34 > 0 else 0.0
        traffic = None    # dram__bytes_read + write of one launch, from the committed ncu --set full capture of this workload
        for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if traffic is None and os.path.exists(tpath) and self.world == 1 and not args.force_generic and self.profile == "uniform":
                traffic = json.load(open(tpath))["dram_bytes_per_launch"].get(f"{workload_key}:nbits={self.nbits}")
        l2_bytes = pairs * 8192.0 + self.ne_local * (256.0 + 4 + R)
        # Shared-memory data pipe (128 B/clk/SM): what the launch moves through it, from the kernel's own counters of the last
        # step -- per 4-query group the 32 KB query tile is written once (bulk copy) and read once (the MMA's A operand), the
        # passage tile (N rows x 256 B) is read once per group (B operand), and every decompressed row costs a 256 B store and
        # 256 B of bucket-weight table reads.  tools/mma_operand_bench.cu: a shared-memory-operand MMA streams 128 B/clk.
        try:
            groups, group_rows, passage_rows = self.s.stat("tc_groups"), self.s.stat("tc_group_rows"), self.s.stat("tc_passage_rows")
        except Exception:
            groups = group_rows = passage_rows = 0.0
        smem_bytes = groups * 65536.0 + group_rows * 256.0 + passage_rows * 512.0
        sm_hz = (self.sm_mhz or 0.0) * 1e6
        smem_peak = 128.0 * 148 * sm_hz / 1e9
        smem = {"bytes_per_launch": smem_bytes, "achieved": smem_bytes / t34 / 1e9 if t34 > 0 else 0.0, "unit": "GB/s",
                "peak": smem_peak, "frac": (smem_bytes / t34 / 1e9 / smem_peak) if (t34 > 0 and smem_peak > 0) else None,
                "groups_per_launch": groups, "clocks_per_group_per_sm": (t34 * sm_hz * 148 / groups) if groups else None,
                "peak_source": "128 B/clk/SM x 148 SMs x the SM clock sampled during the timed region",
                "note": "the resource that bounds k_maxsim_tc (DESIGN.md section 4): bulk-copy writes + tcgen05 operand reads + "
                        "decompression stores / table reads share the SM's 128 B/clk shared-memory data pipe; measured in isolation "
                        "(profiles/r02_mma_operand_bench.txt) an M=128, N=80 shared-memory-operand MMA takes 52 clocks = 128 B/clk"}
        return {"kernel": "k_maxsim_tc (fused decompress + MaxSim, tcgen05)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"],
                "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": traffic,
                "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": pk["src"] + ", STREAM-style copy",
                "note": "algorithmic bytes = (4 + R) B x pair embeddings + 16 B x pairs (SURVEY 8d); the packed index itself streams "
                        "from DRAM once per batch (passage-major), so real DRAM traffic is far below this",
                "tensor": {"achieved": tf, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": tf / pk["tf_sus"],
                           "peak_source": "sustained bf16 cuBLAS (fp16 and bf16 share the tcgen05 rate)"},
                "l2_to_sm": {"bytes_per_launch": l2_bytes, "achieved": l2_bytes / t34 / 1e9 if t34 > 0 else 0.0, "unit": "GB/s",
                             "measured_chip_ceiling": 22200.0, "one_issuing_warp_8KB_copies": 8890.0,
                             "note": "8 KB query tile per pair + 292 B per indexed embedding; ceilings measured by "
                                     "tools/l2_to_sm_ceiling.cu (profiles/r02_l2_to_sm_ceiling.txt); the kernel is bound by the "
                                     "shared-memory / L1 data pipe (TMA writes + UMMA operand reads + LSU), DESIGN.md section 4"},
                "smem_data_pipe": smem,
                "stage_ms": prof, "pairs_per_step": pairs, "pair_embeddings_per_step": pair_embs}


def plaid_oracle_gate(torch, leg, knobs, nqc):
    """BASELINE config 5 has no reference implementation (README.md:187): the semantics are the oracle's own
    (oracle.plaid_search), so this gate is labelled parity-UNPINNED."""
    from oracle import oracle as O
    oix = host_oracle_index(torch, leg.wl, leg.nbits, leg.cen, leg.doclens, leg.codes, leg.res, knobs["ncells"])
    gp, gs, gc = leg.out_p.cpu().numpy(), leg.out_s.cpu().numpy(), leg.out_c.cpu().numpy()
    Qh = leg.Qd[:nqc].cpu().numpy()
    ok_set = ok_sel = True
    max_rel, nswap, nswap_beyond_noise = 0.0, 0, 0
    for q in range(nqc):
        op, osc, sel, cand, approx = O.plaid_search(oix, Qh[q].T, leg.k, knobs["ncells"], knobs["centroid_score_threshold"],
                                                    knobs["ndocs"], return_selected=True)
        kk = len(op)
        ok_set &= set(gp[q, :kk].tolist()) == set(op.tolist())
        ok_sel &= bool(gc[q] == len(sel)) and set(gp[q, :kk].tolist()) <= set(sel.tolist())
        max_rel = max(max_rel, float(np.max(np.abs(gs[q, :kk] - osc) / np.abs(osc))))
        oscore = dict(zip(op.tolist(), osc.tolist()))
        for i in np.nonzero(gp[q, :kk] != op)[0]:
            nswap += 1
            a = oscore.get(int(gp[q, i]))
            # float noise between two fp32 summation orders: 32 x 128 terms ~ 1e-5 relative
            nswap_beyond_noise += int(a is None or abs(a - float(osc[i])) > 1e-5 * abs(float(osc[i])))
    return {"pinned": False, "why": "no reference implementation of PLAID pruning exists (README.md:187); semantics = oracle.plaid_search",
            "queries_checked": nqc, "topk_sets_identical": bool(ok_set), "selection_counts_identical": bool(ok_sel),
            "order_differences": nswap, "order_differences_beyond_fp32_noise": nswap_beyond_noise,
            "max_rel_score_err": max_rel, "tolerance": 1e-3}


def main_single_process(args, torch, cfg_out):
    """--gpus N --single-process: the N shards live in this process, one per device, and every step is ONE
    cb_multi_search_batch call with host buffers (H2D of the queries and D2H of the results inside the timed region)."""
    import colbert_jl_b200 as cb
    n, nq, k, T, dim = args.gpus, args.nq, args.k, 32, 128
    assert torch.cuda.device_count() >= n, f"--gpus {n} but only {torch.cuda.device_count()} devices are visible"
    legs = []
    for r in range(n):
        torch.cuda.set_device(r)
        lg = Leg(torch, cb, args, args.workload, args.nbits, args.profile, args.nprobe, k, r, n, r, None)
        lg.drop_host_copies()
        legs.append(lg)
    torch.cuda.set_device(0)
    multi = cb.MultiSearcher([lg.s for lg in legs])
    Qh = torch.empty((nq, T, dim), dtype=torch.float32, pin_memory=True)
    Qh.copy_(legs[0].Qd)
    hp = torch.zeros((nq, k), dtype=torch.int64, pin_memory=True)
    hs = torch.zeros((nq, k), dtype=torch.float32, pin_memory=True)
    hc = torch.zeros((nq,), dtype=torch.int32, pin_memory=True)

    def step():
        multi.search_batch_ptr(Qh.data_ptr(), nq, T, k, hp.data_ptr(), hs.data_ptr(), hc.data_ptr(), nprobe=args.nprobe)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()                                       # synchronous: results are on the host when it returns
    sec = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    launches = sum(int(lg.s.stat("launches")) for lg in legs) * 2 + 1     # probe + search per shard, + the merge
    pairs = sum(lg.s.stat("pairs") for lg in legs)
    result_digest = digest(hp, hs)
    parity = {"digest": result_digest, "rescore_unsafe_queries": sum(lg.s.stat("rescore_unsafe") for lg in legs)}
    if not args.no_gate:
        tot = [0, 0, 0]
        for r, lg in enumerate(legs):
            torch.cuda.set_device(r)
            g = all_query_gate(torch, lg.s, legs[0].Qd, hp, hs, hc, k, args.nprobe, lg.lo, lg.hi)
            tot = [tot[0] + g["pairs_rescored_exact_fp32"], tot[1] + g["score_mismatches"], max(tot[2], g["order_violations"])]
        parity["all_query_gate"] = {"queries": nq, "returned_pairs_rescored_exact_fp32": tot[0], "score_mismatches": tot[1],
                                    "order_violations": tot[2]}
    dpath = os.path.join(ROOT, "profiles", "r02_result_digests.json")
    dkey = f"{args.workload}:nbits={args.nbits}:profile={args.profile}:nq={nq}:k={k}:nprobe={args.nprobe}"
    if os.path.exists(dpath):
        ref_d = json.load(open(dpath)).get(dkey)
        parity["digest_n1_committed"], parity["equals_n1"] = ref_d, (None if ref_d is None else bool(ref_d == result_digest))
    val = nq / sec
    emit(json.dumps({"metric": "queries/sec", "value": val, "unit": "queries/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "f16 operands / f32 accumulate (tcgen05); f32 exact decisions and final scores", "data": "synthetic",
                      "config": dict(cfg_out, sharding=f"passage-range x{n}", launcher="one process, one host thread: cb_multi_search_batch"),
                      "clocks": clocks, "timing": "host wall clock around the synchronous C-ABI call (host buffers in and out)",
                      "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": nq * T * dim * 4, "d2h_bytes_per_step": nq * k * 12 + nq * 4 * n},
                      "gpu_launches": launches * args.steps, "pairs_per_step": pairs, "parity": parity}))
    multi.close()
    for lg in legs:
        lg.close()


def main():
    # Libraries write to file descriptor 1 behind Python's back (NCCL prints its version banner there when
    # NCCL_DEBUG is set): keep the original stdout for the JSON line only and point fd 1 at stderr.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    import torch
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    T, dim, nq, k = 32, 128, args.nq, args.k
    cfg_out = {"workload": wl["name"] + f", nbits={args.nbits}", "passages": wl["passages"], "centroids": wl["K"],
               "nbits": args.nbits, "queries_per_step": nq, "query_tokens": T, "dim": dim, "k": k, "nprobe": args.nprobe,
               "code_profile": args.profile, "sharding": f"passage-range x{world}",
               "l2_policy": "inputs (packed index, GBs) are far larger than the 126 MB L2; no flush needed",
               "generator": "torch Philox streams on the device (seeds 1001-1004 / 2001, SURVEY 8d distributions); the unit tests use "
                            "the numpy PCG64 twin in colbert.jl_b200/synthetic.py: same distributions, different streams"}

    if args.impl == "reference":
        if rank != 0:
            return
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        cen, doclens, csum = gen_global(torch, wl, dev)
        codes, res = gen_shard(torch, wl, dev, csum, 0, wl["passages"], args.nbits, args.profile)
        Qd = gen_queries(torch, cen, max(8, args.steps + args.warmup), T, args.nprobe, dev)
        oix = host_oracle_index(torch, wl, args.nbits, cen, doclens, codes, res, args.nprobe)
        Qh = Qd.cpu().numpy()
        del codes, res
        per = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            from oracle import oracle as O
            O.search(oix, Qh[s % Qh.shape[0]].T, k)
            if s >= args.warmup:
                per.append(time.perf_counter() - t0)
        sec = float(np.mean(per))
        val = 1.0 / sec
        sample = f"1 query per step ({args.steps} timed) over the full index, numpy/OpenBLAS restatement of ColBERT.jl search"
        emit(json.dumps({"impl": "reference", "metric": "queries/sec", "value": val, "unit": "queries/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_out,
                          "cpu_baseline": {"value": val, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # -------------------------------------------------------------------------------------------- ours
    import colbert_jl_b200 as cb
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product path has no CPU fallback"
    if args.single_process and args.gpus > 1 and world == 1:
        return main_single_process(args, torch, cfg_out)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    leg = Leg(torch, cb, args, args.workload, args.nbits, args.profile, args.nprobe, k, rank, world, local, dist)
    s, Qd, out_p, out_s, out_c = leg.s, leg.Qd, leg.out_p, leg.out_s, leg.out_c
    keep_for_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline)
    if not keep_for_cpu:
        leg.drop_host_copies()
    lib = cb.load()
    t_build = leg.t_build

    ms_per_step, clocks = leg.timed(args.steps, args.warmup, sample_clocks=True)
    value = nq / (ms_per_step * 1e-3)
    if os.environ.get("CB_TC_PROF_OUT") and hasattr(lib, "cb_debug_tc_prof"):   # -DTC_PROF=1 measurement builds only (tools/tc_wait_profile.py)
        import ctypes as C
        import numpy as _np
        _buf = _np.zeros((160, 32, 24), dtype=_np.uint64)
        lib.cb_debug_tc_prof(_buf.ctypes.data_as(C.c_void_p))
        _np.save(os.environ["CB_TC_PROF_OUT"], _buf)
        _tr = _np.zeros((16, 4096), dtype=_np.int64)
        lib.cb_debug_tc_trace(_tr.ctypes.data_as(C.c_void_p))
        _np.save(os.environ["CB_TC_PROF_OUT"].replace(".npy", "_trace.npy"), _tr)
    # kernels launched by this rank's library per step (+ the all-gathers / merge of the sharded path)
    launches = int(s.stat("launches")) + (3 if world > 1 else 0)
    pairs, pair_embs = s.stat("pairs"), s.stat("pair_embeddings")
    flagged, unsafe = s.stat("flagged_rows"), s.stat("rescore_unsafe")
    result_digest = digest(out_p, out_s)

    # ---- end to end: pinned host buffers in, host results out, every step
    Qh = torch.empty((nq, T, dim), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    hp = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
    hs = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
    hc = torch.empty((nq,), dtype=torch.int32, pin_memory=True)
    Qd2 = torch.empty_like(Qd)

    def e2e_step():
        if world == 1:  # the reference-facing C-ABI call with HOST buffers (copies inside)
            cb._lib.check(lib.cb_search_batch(s._h, Qh.data_ptr(), nq, T, args.nprobe, k, hp.data_ptr(), hs.data_ptr(), hc.data_ptr()))
        else:
            Qd2.copy_(Qh, non_blocking=True)
            leg.step(Q=Qd2)
            hp.copy_(out_p, non_blocking=True)
            hs.copy_(out_s, non_blocking=True)
            hc.copy_(out_c, non_blocking=True)
            torch.cuda.synchronize()

    e2e_step()
    leg.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    leg.barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = nq * args.steps / float(t_e2e.item())
    host_equals_device = bool(torch.equal(hp.to(dev), out_p) and torch.equal(hs.to(dev), out_s))
    if not args.ablation:   # (ablation builds compute garbage on purpose)
        assert host_equals_device, "host and device entry points disagree"

    # ---- roofline of the dominant kernel (fused decompress + MaxSim), CUDA events on its stream
    prof = leg.stage_profile()
    if world > 1:   # per-rank shard numbers: report the slowest rank's stage times, the job's pair totals
        tt = torch.tensor([prof[kx] for kx in sorted(prof)], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        prof = dict(zip(sorted(prof), [float(x) for x in tt.tolist()]))
    roofline = leg.roofline(prof, pairs, pair_embs, args.workload)
    roofline["scope"] = ("this rank's shard (stage times = max over ranks; ms_stage1 = stage 1 on the rank's query slice + the all-gather of the cells, "
                         "which also absorbs the skew between the ranks' previous steps)") if world > 1 else "the whole index"

    # ---- parity at every N: all-query gate on every rank, result digest against the committed N = 1 digest
    parity = {"digest": result_digest, "tolerance": 1e-3, "stage1_rows_redone_by_exact_scan": flagged, "rescore_unsafe_queries": unsafe,
              "host_entry_equals_device_entry": host_equals_device}
    if not args.no_gate:
        g = all_query_gate(torch, s, Qd, out_p, out_s, out_c, k, args.nprobe, leg.lo, leg.hi)
        rc = retrieve_count_gate(s, Qd, out_c, args.nprobe, every=4)
        tt = torch.tensor([g["pairs_rescored_exact_fp32"], g["score_mismatches"], g["order_violations"], rc["queries"],
                           rc["count_mismatches"]], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            tt[2] //= world   # every rank checks the same (global) lists for order
        parity["all_query_gate"] = {"queries": nq, "returned_pairs_rescored_exact_fp32": int(tt[0]), "score_mismatches": int(tt[1]),
                                    "order_violations": int(tt[2]), "retrieve_count_queries_x_shards": int(tt[3]),
                                    "retrieve_count_mismatches": int(tt[4]),
                                    "how": "cb_score_pids (exact fp32, oracle-checked in tests/) on every returned pair, bit for bit; "
                                           "cb_retrieve candidate counts (bit-exact vs the oracle in tests/) on every 4th query per shard"}
    dpath = os.path.join(ROOT, "profiles", "r02_result_digests.json")
    dkey = f"{args.workload}:nbits={args.nbits}:profile={args.profile}:nq={nq}:k={k}:nprobe={args.nprobe}"
    if os.path.exists(dpath):
        ref_d = json.load(open(dpath)).get(dkey)
        parity["digest_n1_committed"] = ref_d
        parity["equals_n1"] = (None if ref_d is None else bool(ref_d == result_digest))
    parity["digest_key"] = dkey

    # ---- CPU baseline (N = 1, rank 0): the oracle on a bounded sample of the same workload, + oracle parity on that sample
    cpu = None
    if keep_for_cpu:
        oix = host_oracle_index(torch, wl, args.nbits, leg.cen, leg.doclens, leg.codes, leg.res, args.nprobe)
        leg.drop_host_copies()
        nqc = args.cpu_queries
        sec, results = time_oracle(oix, Qd[:nqc].cpu().numpy(), k, nqc)
        del oix
        cpu = {"value": 1.0 / sec, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{nqc} of the {nq} queries over the full index; numpy/OpenBLAS restatement of ColBERT.jl search "
                         f"(Julia absent), BLAS threads = {os.cpu_count()}"}
        gp, gs = out_p[:nqc].cpu().numpy(), out_s[:nqc].cpu().numpy()
        ok_p = all(np.array_equal(gp[q], results[q][0]) for q in range(nqc))
        max_rel = max(float(np.max(np.abs(gs[q] - results[q][1]) / np.abs(results[q][1]))) for q in range(nqc))
        parity["oracle"] = {"queries_checked": nqc, "topk_pids_identical": bool(ok_p), "max_rel_score_err": max_rel}
        assert max_rel <= 1e-3, "scores differ from the oracle by more than the 1e-3 tolerance"

    # ---- the other BASELINE.json configs (N = 1): same harness, short runs, same gates
    configs = None
    if world == 1 and not args.no_extra:
        leg.close()
        del leg, s, Qd, out_p, out_s, out_c
        torch.cuda.empty_cache()
        configs = {}
        plaid_knobs = dict(ncells=4, centroid_score_threshold=0.4, ndocs=1000)
        spec = {"B": ("B", 2, "uniform", 2, 10, None), "C_nbits1": ("C", 1, "uniform", 2, 10, None),
                "C_nbits4": ("C", 4, "uniform", 2, 10, None), "C_clustered": ("C", 2, "clustered", 2, 10, None),
                "C_plaid": ("C", 2, "uniform", 4, 100, plaid_knobs)}
        for name in [x for x in args.extra.split(",") if x]:
            wk, nb, profile, nprobe, kk, plaid = spec[name]
            wk = args.extra_workload or wk
            t_leg = time.perf_counter()
            try:
                lg = Leg(torch, cb, args, wk, nb, profile, nprobe, kk, 0, 1, local, None)
                ms, _ = lg.timed(args.extra_steps, 3, plaid=plaid, sample_clocks=True)
                out = {"workload": WORKLOADS[wk]["name"], "nbits": nb, "code_profile": profile, "nprobe": nprobe, "k": kk,
                       "ms_per_step": ms, "value": nq / (ms * 1e-3), "unit": "queries/s", "steps": args.extra_steps, "warmup": 3,
                       "digest": digest(lg.out_p, lg.out_s)}
                if plaid is None:
                    prs, pes = lg.s.stat("pairs"), lg.s.stat("pair_embeddings")
                    out["rescore_unsafe_queries"] = lg.s.stat("rescore_unsafe")
                    rf = lg.roofline(lg.stage_profile(2), prs, pes, wk)
                    out["hbm_equiv"] = {"achieved": rf["achieved"], "peak": rf["peak"], "unit": "GB/s", "frac": rf["frac"]}
                    out["tensor_frac"] = rf["tensor"]["frac"]
                    out["smem_data_pipe_frac"] = rf["smem_data_pipe"]["frac"]
                    out["clocks_per_group_per_sm"] = rf["smem_data_pipe"]["clocks_per_group_per_sm"]
                    out["stage_ms"] = rf["stage_ms"]
                    out["pairs_per_step"], out["pair_embeddings_per_step"] = prs, pes
                    if not args.no_gate:
                        out["parity"] = all_query_gate(torch, lg.s, lg.Qd, lg.out_p, lg.out_s, lg.out_c, kk, nprobe, lg.lo, lg.hi, full=False)
                        out["parity"].update(retrieve_count_gate(lg.s, lg.Qd, lg.out_c, nprobe, every=64))
                else:
                    out["plaid"] = dict(plaid_knobs, candidate_pairs=lg.s.stat("plaid_candidates"),
                                        surviving_query_centroid_pairs=lg.s.stat("plaid_survivors"),
                                        exactly_rescored_pairs=lg.s.stat("plaid_rescored"))
                    if not args.no_gate:
                        out["parity"] = all_query_gate(torch, lg.s, lg.Qd, lg.out_p, lg.out_s, lg.out_c, kk, nprobe, lg.lo, lg.hi, full=False)
                        if not args.no_cpu_baseline:
                            out["parity"]["oracle"] = plaid_oracle_gate(torch, lg, plaid_knobs, 1)
                lg.close()
                del lg
            except Exception as ex:  # a leg must never take the headline line down with it
                out = {"error": f"{type(ex).__name__}: {ex}"}
            out["leg_wall_s"] = time.perf_counter() - t_leg
            configs[name] = out
            torch.cuda.empty_cache()

    if rank == 0:
        emit(json.dumps({"metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (tcgen05); f32 exact decisions and final scores",
                          "data": "synthetic", "config": cfg_out, "clocks": clocks,
                          "e2e": {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": nq * T * dim * 4,
                                  "d2h_bytes_per_step": nq * k * 12 + nq * 4},
                          "gpu_launches": launches * args.steps, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                          "configs": configs, "index_build_s": t_build}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Timeline of CTA 0 of a -DTC_PROF=1 build (event trace written by bench.py next to CB_TC_PROF_OUT): per-group clocks between the
pipeline events of the scoring kernel.  Usage: tc_trace.py prof_trace.npy"""
import sys
import numpy as np

EV = ["mma:a_full", "mma:d_empty", "mma:issued", "epi:d_full", "epi:read_done", "epi:end", "load:a_empty", "conv:a_full", "conv:at_empty", "conv:at_full"]
t = np.load(sys.argv[1]).astype(np.int64)[:10]
if (t[0] == 0).all():      # light trace: only mma:issued (2), epi:d_full (3), epi:end (5)
    ok = (t[2] > 0) & (t[3] > 0) & (t[5] > 0)
    t = t[:, ok]
    d = lambda x: x.astype(float)
    print(f"light trace, {int(ok.sum())} groups")
    for name, v in (("period: issued(g) -> issued(g+1)", np.diff(t[2])), ("issued -> epilogue sees d_full", t[3] - t[2]), ("epilogue: d_full -> end", t[5] - t[3]),
                    ("epilogue idle: end(g) -> d_full(g+1)", t[3][1:] - t[5][:-1]), ("issuer lead: issued(g+1) - d_full(g)", t[2][1:] - t[3][:-1]),
                    ("issued(g+2) - end(g)   [accumulator reuse]", t[2][2:] - t[5][:-2])):
        v = d(v); print(f"  {name:46s} median {np.median(v):7.0f}  mean {v.mean():7.0f}  p10 {np.percentile(v, 10):7.0f}  p90 {np.percentile(v, 90):7.0f}")
    b = t[2][100]
    for g in range(100, 110):
        print(f"  g{g}: issued={t[2][g]-b} d_full={t[3][g]-b} end={t[5][g]-b}")
    sys.exit(0)
ok = (t[0] > 0) & (t[2] > 0) & (t[3] > 0) & (t[5] > 0)
n = int(ok.sum())
t = t[:, ok]
print(f"{n} groups traced; period (mma:issued to next) median {np.median(np.diff(t[2])):.0f}, mean {np.diff(t[2]).mean():.0f}")
def stat(name, d):
    d = d[np.isfinite(d)]
    print(f"  {name:42s} median {np.median(d):7.0f}  mean {d.mean():7.0f}  p10 {np.percentile(d, 10):7.0f}  p90 {np.percentile(d, 90):7.0f}")
stat("mma: a_full -> d_empty granted", (t[1] - t[0]).astype(float))
stat("mma: d_empty -> issued (8 MMA + commits)", (t[2] - t[1]).astype(float))
stat("mma issued -> epi sees d_full (TC exec)", (t[3] - t[2]).astype(float))
stat("epi: d_full -> all columns in registers", (t[4] - t[3]).astype(float))
stat("epi: read done -> end of group", (t[5] - t[4]).astype(float))
stat("epi: end(g) -> d_full(g+1)", (t[3][1:] - t[5][:-1]).astype(float))
stat("mma: issued(g) -> a_full seen (g+1)", (t[0][1:] - t[2][:-1]).astype(float))
if (t[6] > 0).all():
    stat("load: copy issue(g) -> mma a_full(g)", (t[0] - t[6]).astype(float))
if (t[7] > 0).all() and len(sys.argv) > 2 and sys.argv[2] == "lanes":      # two-lane kernel: lane 0, events 7-9 = conversion of own group g
    stat("work: convert start -> rows in registers", (t[8] - t[7]).astype(float))
    stat("work: rows in registers -> A tile stored", (t[9] - t[8]).astype(float))
    stat("work: A tile stored(g) -> issuer sees it", (t[0] - t[9]).astype(float))
    stat("work: A tile stored(g) -> d_full(g)", (t[3] - t[9]).astype(float))
    stat("work: epilogue end(g) -> convert start(g+2)", (t[7][2:] - t[5][:-2]).astype(float))
    stat("work: convert done(g+2) -> d_full(g+1)", (t[3][1:-1] - t[9][2:]).astype(float))
elif (t[7] > 0).all():
    stat("conv: a_full -> at_empty granted", (t[8] - t[7]).astype(float))
    stat("conv: at_empty -> at_full arrive", (t[9] - t[8]).astype(float))
    stat("conv at_full -> mma a_full seen", (t[0] - t[9]).astype(float))
print("sample (clocks relative to the first event; one row per group):")
b = t[0][100]
for g in range(100, 112):
    print("  g%-4d " % g + "  ".join(f"{EV[i]}={t[i][g] - b}" for i in range(10) if t[i][g] > 0))

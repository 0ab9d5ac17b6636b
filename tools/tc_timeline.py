"""Timeline of CTA 0 of a -DTC_PROF=2 -DTC_TRACE_EVENTS=0x3f build: per-passage events (scheduler / decompression / issuer) next to
the per-group events of its groups.  usage: tc_timeline.py trace.npy [first_entry_offset] [n_entries]"""
import sys
import numpy as np
t = np.load(sys.argv[1]).astype(np.int64)
N = t.shape[1]
G0, E0 = 20000, 4700
o0 = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ne = int(sys.argv[3]) if len(sys.argv) > 3 else 14
g, p = t[:7], np.concatenate([t[12:13], t[7:12]])          # g[ev][group - G0] (6 = loader issue); p[0] = first group of entry, p[1..5] = events 7..11
first = p[0]
base = p[1][o0]
def r(x): return int(x - base) if x > 0 else None
print("clocks relative to the scheduler getting the slot of the first entry shown")
for i in range(o0, o0 + ne):
    fg = int(first[i]); ng = int(first[i + 1] - fg) if first[i + 1] > 0 else 0
    print(f"entry +{i}: sched slot {r(p[1][i])} publish {r(p[2][i])} | dec start {r(p[3][i])} done {r(p[4][i])} | issuer b_full {r(p[5][i])}  groups {fg}..{fg + ng - 1}")
    for gg in range(fg, fg + ng):
        k = gg - G0
        if 0 <= k < N:
            print(f"      g{gg}: copies issued {r(g[6][k])} a_full {r(g[0][k])} d_empty {r(g[1][k])} issued {r(g[2][k])} | epi d_full {r(g[3][k])} read {r(g[4][k])} end {r(g[5][k])}")
# summary statistics over the traced window
ok = (p[1] > 0) & (p[2] > 0) & (p[3] > 0) & (p[4] > 0) & (p[5] > 0)
idx = np.nonzero(ok)[0]
def st(n, v):
    v = v.astype(float); print(f"  {n:52s} median {np.median(v):7.0f} mean {v.mean():7.0f} p10 {np.percentile(v, 10):7.0f} p90 {np.percentile(v, 90):7.0f}")
print(f"{len(idx)} entries traced")
st("passage period (publish to publish)", np.diff(p[2][idx]))
st("sched: slot granted -> published", p[2][idx] - p[1][idx])
st("published -> decompression starts", p[3][idx] - p[2][idx])
st("decompression: start -> done", p[4][idx] - p[3][idx])
st("decompression done -> issuer opens the passage", p[5][idx] - p[4][idx])
st("published -> issuer opens the passage", p[5][idx] - p[2][idx])
gk = (g[6] > 0) & (g[0] > 0) & (g[2] > 0)
st("group: copies issued -> a_full seen by issuer", (g[0] - g[6])[gk])
gi = np.nonzero(gk)[0]; gi = gi[gi >= 3]; gi = gi[gk[gi - 3]]
st("group: issued(g-3) -> copies issued(g)  [stage reuse]", (g[6][gi] - g[2][gi - 3]))
st("group: issued(g-3) -> a_full(g)", (g[0][gi] - g[2][gi - 3]))
i2 = idx[(idx + 1 < N)]
i2 = i2[ok[i2 + 1]]
st("issuer: open(e) -> open(e+1)", p[5][i2 + 1] - p[5][i2])

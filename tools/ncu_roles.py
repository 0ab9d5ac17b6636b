"""Per-role summary of an ncu source page of k_maxsim_tc: finds every mbarrier first-try wait site,
maps it to the `tag` of ptx::mbar_wait, and sums stall samples between role boundaries.
usage: python tools/ncu_roles.py report.ncu-rep"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
A = [(int(r[ci["Address"]][-5:], 16), r[ci["Source"]].strip(), int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]), r) for r in data]
tot = sum(a[2] for a in A)
print("kernel", rows[0][1][:60], "samples", tot, "warp-instr", sum(a[3] for a in A))
TAGS = {1: "sched:meta_empty", 2: "sched:b_empty(ring)", 3: "sched:end", 4: "mma:meta_full", 5: "mma:b_full", 6: "mma:a_full",
        7: "mma:d_empty", 8: "load:meta_full", 9: "load:a_empty", 10: "epi:meta_full", 11: "epi:d_full", 12: "dec:meta_full"}
sites = []
for i, (ad, src, s, n, r) in enumerate(A):
    if "TRYWAIT" in src:
        tag = None
        for j in range(i + 1, min(i + 45, len(A))):
            m = re.search(r"(IMAD\.MOV\.U32|MOV) R\d+, (RZ, RZ, )?(0x[0-9a-f]+)", A[j][1])
            if m and 0 < int(m.group(3), 16) < 32: tag = int(m.group(3), 16); break
            if "BPT.TRAP" in A[j][1]: break
        # the whole wait routine: first try, spin loop (second TRYWAIT, clock reads, compare), up to the trap
        end = i + 3
        for j in range(i + 1, min(i + 45, len(A))):
            if "BPT.TRAP" in A[j][1]: end = j + 1; break
        if i > 0 and any("TRYWAIT" in A[j][1] for j in range(max(0, i - 12), i)): continue   # the spin-loop TRYWAIT of the previous site
        wait = sum(A[j][2] for j in range(i, min(end, len(A))))
        sites.append((ad, tag, n, wait))
agg = collections.OrderedDict()
for ad, tag, n, w in sites:
    k = TAGS.get(tag, f"tag {tag}")
    a = agg.setdefault(k, [0, 0]); a[0] = max(a[0], n); a[1] += w
for k, (n, w) in agg.items():
    print(f"  wait {k:22s} first-try exec {n:>12d}  samples {w:>8d} ({100*w/tot:.1f}% of all)")
# role boundaries: first wait site of each role
firsts = {}
for ad, tag, n, w in sites:
    role = TAGS.get(tag, "?").split(":")[0]
    firsts.setdefault(role, ad)
order = sorted(firsts.items(), key=lambda x: x[1])
for i, (role, lo) in enumerate(order):
    hi = order[i + 1][1] if i + 1 < len(order) else 1 << 30
    t = sum(a[2] for a in A if lo <= a[0] < hi); ins = sum(a[3] for a in A if lo <= a[0] < hi)
    st = collections.Counter()
    for a in A:
        if lo <= a[0] < hi:
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h: st[h[6:]] += int(a[4][ci[h]])
    print(f"role {role:6s} [{lo:#x},{hi:#x}) samples {t:>8d} ({100*t/tot:.1f}%) warp-instr {ins:>12d}", [(k, v) for k, v in st.most_common(5)])

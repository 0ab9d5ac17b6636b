"""Summarises an ncu report's source page: top stalled SASS instructions and (optionally) per
address-range totals.  usage: python tools/ncu_src.py report.ncu-rep [top_n] [lo:hi:name ...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in data)
print("kernel:", rows[0][1], " total samples", tot)
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:topn]:
    st = sorted([(s[6:], int(r[ci[s]])) for s in stalls if int(r[ci[s]]) > 0], key=lambda x: -x[1])[:2]
    print(r[ci["Address"]][-5:], r[ci["# Samples"]].rjust(8), r[ci["Instructions Executed"]].rjust(11),
          r[ci["Source"]].strip()[:64].ljust(64), st)
for spec in sys.argv[3:]:
    lo, hi, name = spec.split(":")
    lo, hi = int(lo, 16), int(hi, 16)
    t, ins, st = 0, 0, {s: 0 for s in stalls}
    for r in data:
        ad = int(r[ci["Address"]][-5:], 16)
        if lo <= ad < hi:
            t += int(r[ci["# Samples"]]); ins += int(r[ci["Instructions Executed"]])
            for s in stalls: st[s] += int(r[ci[s]])
    print(name, "samples", t, "warp-instr", ins, sorted([(k[6:], v) for k, v in st.items() if v > t * 0.03], key=lambda x: -x[1]))

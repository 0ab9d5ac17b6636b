"""SASS lines of an ncu source page between two addresses with samples / executed counts / top stall.
usage: python tools/ncu_src_range.py report.ncu-rep 0xLO 0xHI [min_samples]"""
import csv, io, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    ad = int(r[ci["Address"]][-5:], 16)
    if lo <= ad < hi:
        s = int(r[ci["# Samples"]])
        if s < mins: continue
        top = sorted(((int(r[ci[h]]), h[6:]) for h in st), reverse=True)[:2]
        print(f"{ad:#07x} {s:>7d} {int(r[ci['Instructions Executed']]):>11d}  {r[ci['Source']].strip()[:90]:90s} {top}")

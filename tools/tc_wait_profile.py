"""Prints the per-role wait / busy split of a -DTC_PROF=1 build of the scoring kernel (numpy file written by bench.py when
CB_TC_PROF_OUT is set): clocks per 4-query group, averaged over the CTAs.  Usage: tc_wait_profile.py prof.npy groups_per_step"""
import sys
import numpy as np

TAGS = {1: "sched:meta_empty", 2: "sched:b_empty", 3: "sched:end", 4: "mma:meta_full", 5: "mma:b_full", 6: "mma:a_full", 7: "mma:d_empty",
        8: "load:meta_full", 9: "load:a_empty", 10: "epi:meta_full", 11: "epi:d_full", 12: "dec:meta_full", 16: "mma:issue(busy)",
        20: "dec:passage(busy)"}
p = np.load(sys.argv[1]).astype(np.float64)          # [cta][warp][tag]
groups = float(sys.argv[2]) if len(sys.argv) > 2 else None
ncta = int((p[:, 0, 0] > 0).sum())
p = p[:ncta]
tot = p[:, :, 0].mean(axis=0)                         # role-loop clocks per warp
per_cta_groups = groups / ncta if groups else None
print(f"CTAs {ncta}, kernel clocks per CTA {tot.max():.3e}" + (f", clocks per group {tot.max() / per_cta_groups:.0f}" if groups else ""))
for w in range(32):
    if tot[w] == 0:
        continue
    parts = []
    for t in range(1, 24):
        v = p[:, w, t].mean()
        if v > 0:
            parts.append(f"{TAGS.get(t, t)} {100 * v / tot[w]:.1f}%" + (f" ({v / per_cta_groups:.0f})" if groups else ""))
    print(f"warp {w:2d}: " + ", ".join(parts))

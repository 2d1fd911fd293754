"""Warp-state samples of one kernel grouped by ROLE, from an .ncu-rep captured with --import-source on (no GPU needed):
    python tools/ncu_source_roles.py profiles/r2_deform_bf16.ncu-rep [kernel-name regex] [min share of samples to list a line]
Instructions that execute equally often belong to the same loop / role (e.g. the gather loop of the tcgen05 deform_conv2d kernel:
16 warps x 1024 CTAs x 36 steps), so grouping the SASS lines by their execution count gives the share of warp time and of
instructions per role; the hottest single instructions follow.  Used for profiles/deform_conv2d_r2.md and roi_align_r2.md."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kern] if kern else [])
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
hdr = [r for r in rows if "Address" in r][0]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows if len(r) == len(hdr) and r[isamp].isdigit()]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
tote = sum(int(r[iex]) for r in data)
print(f"{len(data)} SASS lines, {tot} warp samples, {tote / 1e6:.1f} M warp instructions")
roles = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for r in data:
    g = roles[int(r[iex])]
    g[0] += int(r[isamp]); g[1] += int(r[iex]); g[2] += 1
    for h in stalls:
        g[3][h] += int(r[hdr.index(h)])
for k, g in sorted(roles.items(), key=lambda kv: -kv[1][0])[:10]:
    top = ", ".join(f"{h[6:]} {100 * n / max(g[0], 1):.0f}%" for h, n in g[3].most_common(4))
    print(f"executed {k:>9d} x: {g[2]:4d} lines, {100 * g[0] / tot:5.1f}% of samples, {100 * g[1] / tote:5.1f}% of instructions | {top}")
for i, r in enumerate(data):
    s = int(r[isamp])
    if s > tot * thr:
        top = sorted(((int(r[hdr.index(h)]), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{i:5d} {r[isrc].strip()[:58]:58s} {100 * s / tot:4.1f}%  x{r[iex]}  {top}")

"""Per-source-line samples for one stall reason (cuda,sass view)."""
import csv, subprocess, sys, collections
rep, reason = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; fname = ""; agg = collections.defaultdict(lambda: [0, ""]); tot = 0
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if hdr is None:
        if "Instructions Executed" in r: hdr = r; ic = r.index(reason)
        continue
    if len(r) <= ic: continue
    if r[0] != "": cur = fname + ":" + r[0]; agg[cur][1] = r[1][:100]; continue
    try: n = int(r[ic])
    except ValueError: continue
    if cur: agg[cur][0] += n; tot += n
print(reason, "total samples", tot)
for l, (n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/max(tot,1):5.1f}%  {l}: {src}")

"""Aggregate executed warp-instructions / stall samples per CUDA source line (cuda,sass view)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; fname = ""
agg = collections.defaultdict(lambda: [0, 0, ""])
tot = stot = 0
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if hdr is None:
        if "Instructions Executed" in r: hdr = r; iN = r.index("Instructions Executed"); iS = r.index("# Samples")
        continue
    if len(r) <= iN: continue
    if r[0] != "":
        cur = (fname, r[0]); agg[cur][2] = r[1][:95]
        continue
    try: n = int(r[iN]); s = int(r[iS])
    except ValueError: continue
    if cur is None: continue
    agg[cur][0] += n; agg[cur][1] += s; tot += n; stot += s
print("total warp-inst", tot, "samples", stot)
for (f, l), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/max(tot,1):5.1f}% inst {100*s/max(stot,1):5.1f}% samp  {f}:{l}: {src}")

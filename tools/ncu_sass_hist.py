"""Opcode histogram weighted by executed warp-instructions + top stall locations from an .ncu-rep."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
hist = collections.Counter(); tot = 0; samp = collections.Counter(); stot = 0
stall_by_op = collections.defaultdict(collections.Counter)
lines = []
for r in rows[2:]:
    if len(r) <= iN: continue
    try: n = int(r[iN]); s = int(r[iSamp])
    except ValueError: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    op = m.group(2) if m else "?"
    hist[op] += n; tot += n; samp[op] += s; stot += s
    lines.append((s, n, r[iS][:90]))
print(f"total warp-instructions {tot}, samples {stot}")
for op, n in hist.most_common(28):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%   samples {100*samp[op]/max(stot,1):5.1f}%")
print("--- top sampled instructions")
for s, n, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{100*s/max(stot,1):5.2f}%  exec {n:9d}  {src}")

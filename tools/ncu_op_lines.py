"""Executed warp-instructions of the given opcodes per CUDA source line (cuda,sass view).
usage: ncu_op_lines.py report.ncu-rep IMAD,LOP3[,...|*] [top]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; ops = sys.argv[2].split(','); top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; fname = ""
agg = collections.defaultdict(lambda: [0, ""]); sub = collections.Counter(); tot = 0; alltot = 0
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if hdr is None:
        if "Instructions Executed" in r: hdr = r; iN = r.index("Instructions Executed")
        continue
    if len(r) <= iN: continue
    if r[0] != "":
        cur = (fname, r[0]); agg[cur][1] = r[1][:100]; continue
    try: n = int(r[iN])
    except ValueError: continue
    alltot += n
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)(\S*)\s*(.*)", r[3])
    if not m: continue
    if m.group(2) in ops or ops == ['*']:
        agg[cur][0] += n; tot += n
        key = m.group(2) + m.group(3)
        if m.group(2) == "IMAD" and "RZ, RZ" in m.group(4): key += " (move)"
        sub[key] += n
print("ops", ops, "total", tot, "of", alltot, "%.1f%%" % (100 * tot / max(alltot, 1)))
for k, n in sub.most_common(14): print(f"   {100*n/alltot:5.2f}%  {k}")
for (f, l), (n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if n: print(f"{100*n/alltot:5.2f}%  {f}:{l}: {src}")

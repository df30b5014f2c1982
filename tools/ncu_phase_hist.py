"""Executed warp-instructions and stall samples of a kernel grouped by phase.
usage: ncu_phase_hist.py file.ncu-rep file.cubin kernel-substring phases.txt
phases.txt: lines "first_line last_line name" for the kernel-body file (kernels_tensor.cuh).
The SASS stream (address order = nvdisasm order) is walked and every instruction is attributed to
the phase of the most recent kernel-body line, so inlined physics.cuh / vmap3.cuh code inherits
the phase of its call site; out-of-line functions are reported under their own label."""
import csv, re, subprocess, sys, collections
rep, cubin, pat, phases_file = sys.argv[1:5]
body_file = "kernels_tensor.cuh"
phases = [(int(a), int(b), n) for a, b, n in (l.split(None, 2) for l in open(phases_file) if l.strip())]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for k, r in enumerate(rows):
    if "Instructions Executed" in r:
        hdr = r; rows = rows[k + 1:]; break
iN, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
ins = []; src = None; grab = False; label = "main"
for l in txt.splitlines():
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        if grab and ins: break
        grab = pat in m.group(1); continue
    if not grab: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: src = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*(\$\S+):", l)
    if m:
        nm = m.group(1)
        mm = re.search(r"\$_ZN3sse(\d+)([A-Za-z_0-9]+)", nm)
        label = mm.group(2)[:int(mm.group(1))] if mm else nm[-30:]
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+((?:@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]*)", l)
    if m: ins.append((m.group(1).split()[-1], src, label))
assert len(ins) == len(rows), (len(ins), len(rows))
agg = collections.defaultdict(lambda: [0, 0, 0]); cur = "prologue"; tot = stot = 0
for (op, s, label), r in zip(ins, rows):
    n, sm = int(r[iN]), int(r[iS])
    if label != "main":
        key = "ool:" + label
    else:
        if s and s[0] == body_file:
            for a, b, nm in phases:
                if a <= s[1] <= b: cur = nm; break
        key = cur
    agg[key][0] += n; agg[key][1] += sm
    if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"): agg[key][2] += n
    tot += n; stot += sm
print(f"total warp-inst {tot} samples {stot}")
for k, (n, s, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*s/stot:5.1f}% time(samples) {100*n/tot:5.1f}% inst  fp64 share {100*f/max(n,1):4.0f}%  {k}")

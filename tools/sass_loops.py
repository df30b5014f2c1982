"""Static view of the loops of a kernel in a cubin: for every backward branch, the body size in
SASS instructions, its opcode mix and the source line of the loop head (needs -lineinfo).
usage: sass_loops.py file.cubin kernel-name-substring"""
import collections, re, subprocess, sys
cubin, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines = txt.splitlines()
# split into functions
fn = None; body = []
funcs = {}
for l in lines:
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        fn = m.group(1); funcs[fn] = []; continue
    if fn: funcs[fn].append(l)
for name, b in funcs.items():
    if pat not in name: continue
    ins = []  # (label or None, opcode, text, srcline)
    labels = {}
    src = None
    for l in b:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: src = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*(\.L_x_\d+):", l)
        if m: labels[m.group(1)] = len(ins); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)\s*(.*);", l)
        if m: ins.append((m.group(1), m.group(2), src))
    print(name[:70], len(ins), "instructions")
    for idx, (op, args, s) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"(\.L_x_\d+)", args)
            if m and m.group(1) in labels and labels[m.group(1)] <= idx:
                a = labels[m.group(1)]
                c = collections.Counter(o.split(".")[0] for o, _, _ in ins[a:idx + 1])
                f64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD"))
                print(f"  loop [{a}:{idx}] {idx + 1 - a} instr, fp64 {f64}, head {ins[a][2]}, tail {s}: "
                      + ", ".join(f"{k} {v}" for k, v in c.most_common(10)))

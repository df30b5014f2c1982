#!/bin/bash
# Round 2, session H: full GPU suite on the current build; Hex diagonal-E fast path A/B vs the generic
# kernels; L2 prefetch on/off on the headline kernels
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_h.log 2>&1; el "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests_h.log
python - <<'PY'
import sys, os, time
sys.path.insert(0, '.')
from sse_b200 import problems
for generic in (0, 1):
    if generic: os.environ["SSE_B200_GENERIC"] = "1"
    solver, u0 = problems.euler_hex_case(p=4, M=32, lazy=False)
    h = solver.handle
    h.set_state(u0); h.time_residual(3)
    ms, ta, tb = h.time_residual(10, split=True)
    print("hex p4 M=32 generic=%d: %.4f ms/residual (A %.4f, B %.4f) -> %.3f G DOF/s" % (generic, ms/10, ta/10, tb/10, u0.size/(ms/10*1e-3)/1e9), flush=True)
    solver.close()
PY
B="python bench.py --M 20 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
for pf in 1 0; do
  SSE_B200_PREFETCH=$pf timeout 200 $B > gpurun_out/h_pf$pf.json 2> gpurun_out/h_pf$pf.err
  python -c "import json; d=json.load(open('gpurun_out/h_pf$pf.json')); print('prefetch=$pf M20', round(d['ms_per_step'],4), round(d['kernel_ms']['loop_a_ms'],4), round(d['kernel_ms']['loop_b_ms'],4))"
done
el done

#!/bin/bash
# Round 2, session Y: config 3 with the fused A (W/J) A^T stage of the scalar projection (scaling in registers)
mkdir -p gpurun_out
for lib in stablespectralelements.jl_b200/libsse_b200.so build/variants/*.so stablespectralelements.jl_b200/libsse_b200.so; do
  echo "== $(basename $lib .so)"
  SSE_B200_LIB=$PWD/$lib CFG3_M=32 timeout 200 python tools/bench_configs.py 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'], 'residual %.4f ms  A %.4f  B %.4f' % (d['ms_per_residual'], d['loop_a_ms'], d['loop_b_ms']))"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "adv3d" 2>&1 | tail -1

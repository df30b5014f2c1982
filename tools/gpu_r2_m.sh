#!/bin/bash
# Round 2, session M: k_physical streaming (default) vs staged with bulk copies, config-5 sweep
mkdir -p gpurun_out
for st in 0 1; do
  echo "== SSE_B200_PHYS_STAGED=$st"
  for c in "2 2 128" "2 3 128" "2 4 128" "2 4 256" "2 5 128" "2 6 64" "2 7 64" "2 8 64" "2 8 128" "1 4 65536"; do
    SSE_B200_PHYS_STAGED=$st python tools/cfg5_case.py $c
  done
  SSE_B200_PHYS_STAGED=$st timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "physical or advdiff or viscous" 2>&1 | tail -1
done

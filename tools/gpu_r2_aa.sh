#!/bin/bash
# Round 2, session AA: (1) loop A compiled for the entropy-projection path only (variant proj2) vs default at M=20;
# (2) host-buffer pipeline: stream priorities of loop A's stream vs the main stream, chunk counts
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
bash tools/gpu_ab.sh 20; el ab
for v in "0 32" "1 32" "-1 32" "0 24" "0 16"; do
  set -- $v
  SSE_B200_HOST_APRIO=$1 SSE_B200_HOST_CHUNKS=$2 timeout 400 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-secondary --no-check > gpurun_out/aa_$1_$2.json 2> gpurun_out/aa_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/aa_$1_$2.json'))
print('aprio=$1 chunks=$2: device %.3f ms, host-path %.3f ms (floor %.2f)' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pcie_floor_ms']))" || tail -3 gpurun_out/aa_$1_$2.err
  el "e2e $v"
done

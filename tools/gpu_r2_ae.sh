#!/bin/bash
# Round 2, session AE: loop B in pieces alternating between two streams (projection of one piece next to the
# flux kernel of the next): SSE_B200_B_PIECES at M=44, and on config 3
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
for np in 1 2 4 8 16 32 1; do
  SSE_B200_B_PIECES=$np timeout 300 python bench.py --steps 15 --warmup 4 --no-e2e --no-cpu-baseline --no-secondary --no-check > gpurun_out/ae_p$np.json 2> gpurun_out/ae_p$np.err
  python -c "
import json
d=json.load(open('gpurun_out/ae_p$np.json')); print('pieces=$np M=44', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))" || tail -3 gpurun_out/ae_p$np.err
done; el m44
for np in 1 4 16; do
  SSE_B200_B_PIECES=$np timeout 300 python bench.py --config 3 --M 40 --steps 15 --warmup 4 --no-e2e --no-cpu-baseline --no-secondary --no-check > gpurun_out/ae_c3_p$np.json 2> gpurun_out/ae_c3_p$np.err
  python -c "
import json
d=json.load(open('gpurun_out/ae_c3_p$np.json')); print('pieces=$np cfg3 M=40', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))" || tail -3 gpurun_out/ae_c3_p$np.err
done; el c3

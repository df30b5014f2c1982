#!/bin/bash
# Round 2, session AL: thread-strided loops of the specialised kernels with a literal stride (SSE_LOOP_T) vs
# blockDim.x (prev_ak): headline kernels at M=20, config 3 at M=40, config 2
mkdir -p gpurun_out
for lib in stablespectralelements.jl_b200/libsse_b200.so build/variants/prev_ak.so stablespectralelements.jl_b200/libsse_b200.so build/variants/prev_ak.so; do
  name=$(basename $lib .so)
  SSE_B200_LIB=$PWD/$lib timeout 300 python bench.py --M 20 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-check > gpurun_out/al_$name.json 2> gpurun_out/al_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/al_$name.json')); print('$name', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))
for x in d['secondary']['single_gpu'][:2]: print('   ', x['config'][:40], 'A %.4f B %.4f' % (x['loop_a_ms'], x['loop_b_ms']))
c=d['secondary']['cfg3']; print('    cfg3', c['ms_per_step'], c['kernel_ms']['loop_a_ms'], c['kernel_ms']['loop_b_ms'])" || tail -3 gpurun_out/al_$name.err
done

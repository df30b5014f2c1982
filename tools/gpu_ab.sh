#!/bin/bash
# A/B of the default library against every build/variants/*.so on the headline kernels (M given as $1)
M=${1:-20}
mkdir -p gpurun_out
for lib in stablespectralelements.jl_b200/libsse_b200.so build/variants/*.so stablespectralelements.jl_b200/libsse_b200.so; do
  name=$(basename $lib .so)
  SSE_B200_LIB=$PWD/$lib timeout 300 python bench.py --M $M --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary --no-check \
    > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/ab_$name.json')); print('$name', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))" || tail -3 gpurun_out/ab_$name.err
done

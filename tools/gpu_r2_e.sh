#!/bin/bash
# Round 2, session E: loop-B occupancy variants after the projection split; k_project_tet E=4 with M over X
mkdir -p gpurun_out
B="python bench.py --M 20 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
run() { # name lib engine
  SSE_B200_LIB=$2 SSE_B200_TET_ENGINE=$3 timeout 200 $B > gpurun_out/e_$1.json 2> gpurun_out/e_$1.err
  python -c "import json; d=json.load(open('gpurun_out/e_$1.json')); print('$1 M20', round(d['ms_per_step'],4), round(d['kernel_ms']['loop_a_ms'],4), round(d['kernel_ms']['loop_b_ms'],4))" || tail -3 gpurun_out/e_$1.err
}
D=$PWD/stablespectralelements.jl_b200/libsse_b200.so
run fused $D 0
run split $D 2
run fd5 $PWD/build/variants/fd5.so 2
run fd6 $PWD/build/variants/fd6.so 2
run fd5_fused $PWD/build/variants/fd5.so 0
SSE_B200_TET_ENGINE=2 CFG3_M=32 timeout 200 python tools/bench_configs.py 3 2> gpurun_out/cfg3_e.err | cut -c1-330

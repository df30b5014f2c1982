#!/bin/bash
# Round 2, session C: batched engine with asynchronous prologue + padded Z, E = 4 / 5 / 8
mkdir -p gpurun_out
B="python bench.py --M 20 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
run() { # name lib engine
  SSE_B200_LIB=$2 SSE_B200_TET_ENGINE=$3 timeout 200 $B > gpurun_out/c_$1.json 2> gpurun_out/c_$1.err
  python -c "import json; d=json.load(open('gpurun_out/c_$1.json')); print('$1 M20', round(d['ms_per_step'],4), round(d['kernel_ms']['loop_a_ms'],4), round(d['kernel_ms']['loop_b_ms'],4))" || tail -3 gpurun_out/c_$1.err
}
D=$PWD/stablespectralelements.jl_b200/libsse_b200.so
run old $D 0
run e5 $D 3
run e4 $PWD/build/variants/e4.so 3
run e8 $PWD/build/variants/e8.so 3
for eng in 0 3; do SSE_B200_TET_ENGINE=$eng CFG3_M=32 timeout 200 python tools/bench_configs.py 3 2> gpurun_out/cfg3_$eng.err | sed "s/^/cfg3 engine=$eng /" | cut -c1-330; done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "adv3d or euler3d_tet" 2>&1 | tail -2

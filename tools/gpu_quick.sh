#!/bin/bash
# quick A/B on the B200: headline kernels at M=20 (+ DMMA / DFMA peaks), optional config 3
mkdir -p gpurun_out
B="python bench.py --M ${M:-20} --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
timeout 300 $B > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python -c "import json; d=json.load(open('gpurun_out/quick.json')); print('M${M:-20}', round(d['ms_per_step'],4), d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], 'fp64 frac', round(d['roofline']['frac'],4))"
python -c "
import sys; sys.path.insert(0,'.')
from sse_b200 import device as dev
print('DFMA peak TFLOP/s', dev.measure_fp64_peak(0), 'DMMA peak TFLOP/s', dev.measure_dmma_peak(0))"
if [ "${CFG3:-0}" = "1" ]; then CFG3_M=32 timeout 200 python tools/bench_configs.py 3 2> gpurun_out/cfg3.err; fi

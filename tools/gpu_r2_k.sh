#!/bin/bash
# Round 2, session K: custom log/exp (flog/fexp) and the fused A-scale-A^T stage: parity tests,
# then A/B against -D variants (old = libdevice log/exp + unfused, nofuse, libdev) at M=20
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 900 python -m pytest tests -m gpu -x -q -k "elementary or parity or golden" > gpurun_out/gpu_tests_k.log 2>&1; el "tests rc=$?"; tail -3 gpurun_out/gpu_tests_k.log
for lib in stablespectralelements.jl_b200/libsse_b200.so build/variants/old.so build/variants/nofuse.so build/variants/libdev.so stablespectralelements.jl_b200/libsse_b200.so; do
  name=$(basename $lib .so)
  SSE_B200_LIB=$PWD/$lib timeout 300 python bench.py --M 20 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary --no-check \
    > gpurun_out/k_$name.json 2> gpurun_out/k_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/k_$name.json')); print('$name', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))" || tail -3 gpurun_out/k_$name.err
done
el done

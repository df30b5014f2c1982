#!/bin/bash
# Round 2, session W: warp-private projection kernel (k_project_tet_w) vs the CTA-level batched one
mkdir -p gpurun_out
run() {
  label=$1; lib=$2; shift 2
  env "$@" SSE_B200_LIB=$PWD/$lib timeout 300 python bench.py --M 20 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary --no-check > gpurun_out/w_$label.json 2> gpurun_out/w_$label.err
  python -c "
import json; d=json.load(open('gpurun_out/w_$label.json')); print('$label', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))" || tail -3 gpurun_out/w_$label.err
}
D=stablespectralelements.jl_b200/libsse_b200.so
run cta $D SSE_B200_PROJ_WARP=0
run warp_minb5 $D SSE_B200_PROJ_WARP=1
run warp_ew2_m3 build/variants/pw_ew2_m3.so SSE_B200_PROJ_WARP=1
run warp_ew2_m2 build/variants/pw_ew2_m2.so SSE_B200_PROJ_WARP=1
run cta $D SSE_B200_PROJ_WARP=0
SSE_B200_LIB=$PWD/build/variants/pw_ew2_m3.so SSE_B200_PROJ_WARP=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "euler3d_tet" 2>&1 | tail -1

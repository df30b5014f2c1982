#!/bin/bash
# First GPU call proposed for round 2 (everything below was CPU-verified on the emulator only):
#   1. the GPU tests that are the first hardware run of the opt-in paths and of sse_erk_step;
#   2. A/B of the split loop B (runtime switch) and of the -D tuning variants on the headline
#      configuration (M=20) and on config 3.
# Before the call, on the CPU:   python tools/build_variants.py $(cat tools/round2_variants.txt)
# and take build/variants/ out of .gpurunignore so that the libraries travel to the box.
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
SSE_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_time_integration.py tests/test_gpu_variants.py -m gpu -q > gpurun_out/t_new.log 2>&1; el "new GPU tests rc=$?"; tail -4 gpurun_out/t_new.log
B="python bench.py --M 20 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
show() { python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'ms/step %.3f' % d['ms_per_step'], d['kernel_ms'])" "$1" "$2" 2>/dev/null || echo "$2 FAILED"; }
timeout 200 $B > gpurun_out/ab_fused.json 2> gpurun_out/ab_fused.err; show gpurun_out/ab_fused.json fused
SSE_B200_SPLIT_B=1 timeout 200 $B > gpurun_out/ab_split.json 2> gpurun_out/ab_split.err; show gpurun_out/ab_split.json split_volminb6
CFG3_M=32 timeout 200 python tools/bench_configs.py 3 2> gpurun_out/cfg3_default.err | sed "s/^/cfg3_default /"
SSE_B200_STD_PIPE=1 CFG3_M=32 timeout 200 python tools/bench_configs.py 3 2> gpurun_out/cfg3_pipe.err | sed "s/^/cfg3_pipe_nb2 /"
SSE_B200_STD_PIPE=2 CFG3_M=32 timeout 200 python tools/bench_configs.py 3 2> gpurun_out/cfg3_pipe_bulk.err | sed "s/^/cfg3_pipe_bulk_nb2 /"
SKIP_TESTS=1 CFG3=1 bash tools/gpu_variants.sh 20
for lib in build/variants/vol_minb*.so; do   # occupancy target of the split volume kernel
  [ -e "$lib" ] || continue
  n=$(basename $lib .so)
  SSE_B200_SPLIT_B=1 SSE_B200_LIB=$PWD/$lib timeout 200 $B > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err; show gpurun_out/ab_$n.json split_$n
done
el done

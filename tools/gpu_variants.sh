#!/bin/bash
# Run on the GPU box: parity tests with the default build, then the per-kernel bench split for the
# default library and every tuning variant under build/variants/ (M given as $1, default 20).
M=${1:-20}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" > gpurun_out/gpu_tests.log 2>&1
  tail -5 gpurun_out/gpu_tests.log
fi
for lib in stablespectralelements.jl_b200/libsse_b200.so build/variants/*.so; do
  name=$(basename $lib .so)
  SSE_B200_LIB=$PWD/$lib timeout 300 python bench.py --M $M --steps 10 --warmup 3 --no-e2e --no-cpu-baseline \
    > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_$name.json"))
    print("$name", "ms/step %.3f" % d["ms_per_step"], d["kernel_ms"], "fp64 frac %.3f" % d["roofline_fp64"]["frac"])
except Exception as e:
    print("$name", "FAILED", e)
PY
  # config 3 (scalar standard-form kernels: SSE_STD_NB / SSE_NODAL_NB variants)
  if [ "${CFG3:-0}" = "1" ]; then
    SSE_B200_LIB=$PWD/$lib CFG3_M=${CFG3_M:-32} timeout 200 python tools/bench_configs.py 3 2> gpurun_out/cfg3_$name.err | sed "s/^/$name /"
  fi
done

#!/bin/bash
# Round 2, session AF: loop A (Tet p=4) with two elements per 256-thread CTA (SSE_B200_NODAL_NT=256) vs 128
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 600 python -m pytest tests/test_gpu_variants.py -x -q -m gpu -k loop_a_projection 2>&1 | tail -2; el test
for nt in 128 256 128 256; do
  SSE_B200_NODAL_NT=$nt timeout 300 python bench.py --M 20 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary --no-check > gpurun_out/af_nt$nt.json 2> gpurun_out/af_nt$nt.err
  python -c "
import json
d=json.load(open('gpurun_out/af_nt$nt.json')); print('NT=$nt M=20', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))" || tail -3 gpurun_out/af_nt$nt.err
done; el m20
SSE_B200_NODAL_NT=256 timeout 300 python bench.py --steps 15 --warmup 4 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/af_m44_nt256.json 2> gpurun_out/af_m44_nt256.err
python -c "
import json
d=json.load(open('gpurun_out/af_m44_nt256.json')); print('NT=256 M=44', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']), d['check']['dudt_digest_u64'])" || tail -3 gpurun_out/af_m44_nt256.err
el m44

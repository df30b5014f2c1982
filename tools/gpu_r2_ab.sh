#!/bin/bash
# Round 2, session AB: loop A with the projection mode compiled in (vs SSE_B200_NODAL_RT_PROJ=1) at M=20 and
# on config 2; parity of both; host-buffer pipeline with 32 / 48 / 64 chunks
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/ab_parity.log 2>&1; el "parity rc=$?"; tail -2 gpurun_out/ab_parity.log
timeout 600 python -m pytest tests/test_gpu_variants.py -x -q -m gpu -k loop_a_projection > gpurun_out/ab_parity_rt.log 2>&1; el "loop A instantiation test rc=$?"; tail -2 gpurun_out/ab_parity_rt.log
for rt in 0 1 0; do
  SSE_B200_NODAL_RT_PROJ=$rt timeout 300 python bench.py --M 20 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-check > gpurun_out/ab_rt$rt.json 2> gpurun_out/ab_rt$rt.err
  python -c "
import json
d=json.load(open('gpurun_out/ab_rt$rt.json')); print('rt_proj=$rt M=20', 'ms/step %.4f' % d['ms_per_step'], 'A %.4f B %.4f' % (d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms']))
for x in d['secondary']['single_gpu'][:1]: print('   ', x['config'], 'A %.4f B %.4f' % (x['loop_a_ms'], x['loop_b_ms']))" || tail -3 gpurun_out/ab_rt$rt.err
done; el rt
for c in 32 48 64; do
  SSE_B200_HOST_CHUNKS=$c timeout 400 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-secondary --no-check > gpurun_out/ab_c$c.json 2> gpurun_out/ab_c$c.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_c$c.json'))
print('chunks=$c: device %.3f ms (A %.3f B %.3f), host-path %.3f ms (floor %.2f)' % (d['ms_per_step'], d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], d['e2e']['ms_per_step'], d['e2e']['pcie_floor_ms']))" || tail -3 gpurun_out/ab_c$c.err
  el "e2e $c"
done

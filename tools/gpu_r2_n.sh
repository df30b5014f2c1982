#!/bin/bash
# Round 2, session N: host-buffer pipeline of sse_residual(where=HOST) at the headline size: tapered vs equal chunks, 16 vs 32 chunks
mkdir -p gpurun_out
for v in "1 32" "0 32" "1 16" "1 24"; do
  set -- $v
  SSE_B200_HOST_TAPER=$1 SSE_B200_HOST_CHUNKS=$2 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-check > gpurun_out/n_t$1_c$2.json 2> gpurun_out/n_t$1_c$2.err
  python -c "
import json; d=json.load(open('gpurun_out/n_t$1_c$2.json'))
print('taper=$1 chunks=$2: device %.3f ms (A %.3f B %.3f), e2e %.3f ms, pcie floor %.3f' % (d['ms_per_step'], d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], d['e2e']['ms_per_step'], d['e2e']['pcie_floor_ms']))" || tail -3 gpurun_out/n_t$1_c$2.err
done

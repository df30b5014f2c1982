#!/bin/bash
# Round 2, session T: same shared-memory carveout for loop A and loop B (co-residency), with and without caps
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" SSE_B200_HOST_NOCOPY=1 timeout 400 python bench.py --M ${M:-32} --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-check > gpurun_out/t_$label.json 2> gpurun_out/t_$label.err
  python -c "
import json; d=json.load(open('gpurun_out/t_$label.json'))
print('$label: monolithic %.3f ms (A %.3f B %.3f), chunked schedule %.3f ms' % (d['ms_per_step'], d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], d['e2e']['ms_per_step']))" || tail -3 gpurun_out/t_$label.err
}
run cv_base_c8 SSE_B200_CARVEOUT_MAX=1 SSE_B200_HOST_CHUNKS=8
run cv_a3_b3 SSE_B200_CARVEOUT_MAX=1 SSE_B200_HOST_CHUNKS=8 SSE_B200_NODAL_SMEM_PAD=58000 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000
run cv_a3_b3_c16 SSE_B200_CARVEOUT_MAX=1 SSE_B200_HOST_CHUNKS=16 SSE_B200_NODAL_SMEM_PAD=58000 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000
run cv_b3 SSE_B200_CARVEOUT_MAX=1 SSE_B200_HOST_CHUNKS=8 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000
run cv_b3_c16 SSE_B200_CARVEOUT_MAX=1 SSE_B200_HOST_CHUNKS=16 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000
run cv_a4_b3 SSE_B200_CARVEOUT_MAX=1 SSE_B200_HOST_CHUNKS=8 SSE_B200_NODAL_SMEM_PAD=42000 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000

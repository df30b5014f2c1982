#!/bin/bash
# Round 2, session S: static SM partition between loop A and loop B (both capped by shared-memory padding)
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" SSE_B200_HOST_NOCOPY=1 timeout 400 python bench.py --M ${M:-32} --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-check > gpurun_out/s_$label.json 2> gpurun_out/s_$label.err
  python -c "
import json; d=json.load(open('gpurun_out/s_$label.json'))
print('$label: monolithic %.3f ms (A %.3f B %.3f), chunked schedule %.3f ms' % (d['ms_per_step'], d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], d['e2e']['ms_per_step']))" || tail -3 gpurun_out/s_$label.err
}
# A alone at 3 / 4 / 6 CTAs per SM (13 KB + pad; 227 KB / n)
run a3 SSE_B200_HOST_CHUNKS=8 SSE_B200_NODAL_SMEM_PAD=58000
run a4 SSE_B200_HOST_CHUNKS=8 SSE_B200_NODAL_SMEM_PAD=42000
run a6 SSE_B200_HOST_CHUNKS=8 SSE_B200_NODAL_SMEM_PAD=24000
# static partitions: B at 3 CTAs (fd pad 14000, proj pad 24000 -> 3 CTAs) + A at 3
run a3_b3 SSE_B200_HOST_CHUNKS=8 SSE_B200_NODAL_SMEM_PAD=58000 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000
run a3_b3_c4 SSE_B200_HOST_CHUNKS=4 SSE_B200_NODAL_SMEM_PAD=58000 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000
run a3_b3fd_only SSE_B200_HOST_CHUNKS=8 SSE_B200_NODAL_SMEM_PAD=58000 SSE_B200_FD_SMEM_PAD=14000

#!/bin/bash
# Round 2, session R: co-residency of loop A (next chunk) with loop B in the chunked schedule (copies skipped):
# loop B's kernels capped at 3 / 4 CTAs per SM by padding their shared-memory request
mkdir -p gpurun_out
run() {  # label env...
  label=$1; shift
  env "$@" SSE_B200_HOST_NOCOPY=1 timeout 400 python bench.py --M ${M:-32} --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-check > gpurun_out/r_$label.json 2> gpurun_out/r_$label.err
  python -c "
import json; d=json.load(open('gpurun_out/r_$label.json'))
print('$label: monolithic %.3f ms (A %.3f B %.3f), chunked schedule %.3f ms' % (d['ms_per_step'], d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], d['e2e']['ms_per_step']))" || tail -3 gpurun_out/r_$label.err
}
run base_c8 SSE_B200_HOST_CHUNKS=8
run base_c16 SSE_B200_HOST_CHUNKS=16
run fd3_c8 SSE_B200_HOST_CHUNKS=8 SSE_B200_FD_SMEM_PAD=14000
run fd3_p4_c8 SSE_B200_HOST_CHUNKS=8 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=9000
run fd3_p4_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=9000
run fd3_p3_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_FD_SMEM_PAD=14000 SSE_B200_PROJ_SMEM_PAD=24000
run p4_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_PROJ_SMEM_PAD=9000

#!/bin/bash
# Round 2, session U: co-scheduling loop A and loop B on the same SMs with capped persistent grids (chunked schedule, copies skipped)
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" SSE_B200_HOST_NOCOPY=1 timeout 400 python bench.py --M ${M:-32} --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-check > gpurun_out/u_$label.json 2> gpurun_out/u_$label.err
  python -c "
import json; d=json.load(open('gpurun_out/u_$label.json'))
print('$label: monolithic %.3f ms (A %.3f B %.3f), chunked schedule %.3f ms' % (d['ms_per_step'], d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], d['e2e']['ms_per_step']))" || tail -3 gpurun_out/u_$label.err
}
run base_c16 SSE_B200_HOST_CHUNKS=16
run persist_full_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_CAP_A=12 SSE_B200_CAP_FD=4 SSE_B200_CAP_PROJ=5
run a3_fd3_p3_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_CAP_A=3 SSE_B200_CAP_FD=3 SSE_B200_CAP_PROJ=3
run a3_fd3_p3_c32 SSE_B200_HOST_CHUNKS=32 SSE_B200_CAP_A=3 SSE_B200_CAP_FD=3 SSE_B200_CAP_PROJ=3
run a3_fd3_p4_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_CAP_A=3 SSE_B200_CAP_FD=3 SSE_B200_CAP_PROJ=4
run a2_fd3_p4_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_CAP_A=2 SSE_B200_CAP_FD=3 SSE_B200_CAP_PROJ=4
run a4_fd3_p3_c16 SSE_B200_HOST_CHUNKS=16 SSE_B200_CAP_A=4 SSE_B200_CAP_FD=3 SSE_B200_CAP_PROJ=3
run a3_fd3_p3_c16_cv SSE_B200_CARVEOUT_MAX=1 SSE_B200_HOST_CHUNKS=16 SSE_B200_CAP_A=3 SSE_B200_CAP_FD=3 SSE_B200_CAP_PROJ=3

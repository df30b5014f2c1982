#!/bin/bash
# Round 2, session F (2 GPUs): sharded residual through the C ABI (sse_shard_*) vs the oracle, both
# backends; then bench.py at N = 2 with both backends (device-resident + host-buffer flows).
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tools/dist_check.py > gpurun_out/r2_dist_check_2gpu.log 2>&1; el "dist_check rc=$?"; grep -v "^W\|^\[W\|NCCL\|warn" gpurun_out/r2_dist_check_2gpu.log | tail -12
for be in library python; do
  SSE_B200_SHARD_BACKEND=$be timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary > gpurun_out/f_n2_$be.json 2> gpurun_out/f_n2_$be.err; el "bench N=2 $be rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/f_n2_$be.json'))
print('$be N=2 ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['e2e']['flow'][:40], d['check']['sharded_vs_single_gpu'].get('match'), d['check']['dudt_digest_u64'], d['check'].get('e2e_digest_matches'))" || tail -5 gpurun_out/f_n2_$be.err
done
SSE_B200_SHARD_BACKEND=python SSE_B200_SHARD_PIPELINE=1 timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary --no-check > gpurun_out/f_n2_pipe.json 2> gpurun_out/f_n2_pipe.err; el "bench N=2 python interleaved rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/f_n2_pipe.json'))
print('python+interleaved N=2 ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],3))"
el done

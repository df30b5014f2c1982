#!/bin/bash
# Round 2, session AI (4 GPUs): sharded flow with the boundary ranges on two streams (default) -- parity incl.
# the sharded CK54 step, driver-style bench at N = 4
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tools/dist_check.py > gpurun_out/r2_dist_check_4gpu_streams.log 2>&1; el "dist_check rc=$?"; grep -v "^W\|^\[W\|NCCL\|warn\|\*\*\*" gpurun_out/r2_dist_check_4gpu_streams.log | tail -16
timeout 600 $T bench.py --gpus 4 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/ai_n4.json 2> gpurun_out/ai_n4.err; el "bench N=4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/ai_n4.json')); c=d.get('check',{})
print('N=4 ms/step', round(d['ms_per_step'],4), 'A', round(d['kernel_ms']['loop_a_ms'],4), 'B', round(d['kernel_ms']['loop_b_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],3), c.get('sharded_vs_single_gpu',{}).get('match'), c.get('dudt_digest_u64'), c.get('e2e_digest_matches'))" || tail -5 gpurun_out/ai_n4.err

#!/bin/bash
# Round 2, session AD (2 GPUs): sharded device flow, SSE_B200_SHARD_STREAMS = 3 (flux kernel by ranges on
# streams, ONE projection launch over the shard) vs 0 / 1 / 2; parity (dist_check incl. config 3, digests)
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tools/dist_check.py > gpurun_out/ad_dist_check_2gpu.log 2>&1; el "dist_check rc=$?"; grep -v "^W\|^\[W\|NCCL\|warn" gpurun_out/ad_dist_check_2gpu.log | tail -11
run() {  # name, mode, extra args
  SSE_B200_SHARD_STREAMS=$2 timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-e2e ${@:3} > gpurun_out/ad_$1.json 2> gpurun_out/ad_$1.err; el "bench $1 rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/ad_$1.json')); c=d.get('check',{})
print('$1 streams=$2 N=2 ms/step', round(d['ms_per_step'],4), 'A', round(d['kernel_ms']['loop_a_ms'],4), 'B', round(d['kernel_ms']['loop_b_ms'],4), c.get('sharded_vs_single_gpu',{}).get('match'), c.get('dudt_digest_u64'))" || tail -5 gpurun_out/ad_$1.err
}
run m28_s3 3 --M 28
run m28_s1 1 --M 28 --no-check
run m28_s0 0 --M 28 --no-check
run m28_s2 2 --M 28 --no-check
run m28_s3b 3 --M 28 --no-check
run m44_s3 3
run m44_s1 1 --no-check
run c3_s3 3 --config 3 --M 28
run c3_s0 0 --config 3 --M 28 --no-check
run c3_s1 1 --config 3 --M 28 --no-check

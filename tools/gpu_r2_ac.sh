#!/bin/bash
# Round 2, session AC (2 GPUs): boundary ranges of the sharded device flow on their own streams
# (SSE_B200_SHARD_STREAMS = 0 old flow / 1 two boundary streams / 2 boundary chain behind the exchange only):
# parity (dist_check, in-run digests) and time at M=44 and at an N=8-sized shard (M=28 on 2 GPUs)
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tools/dist_check.py > gpurun_out/ac_dist_check_2gpu.log 2>&1; el "dist_check rc=$?"; grep -v "^W\|^\[W\|NCCL\|warn" gpurun_out/ac_dist_check_2gpu.log | tail -9
run() {  # name, mode, extra args
  SSE_B200_SHARD_STREAMS=$2 timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-e2e ${@:3} > gpurun_out/ac_$1.json 2> gpurun_out/ac_$1.err; el "bench $1 rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/ac_$1.json')); c=d.get('check',{})
print('$1 streams=$2 N=2 ms/step', round(d['ms_per_step'],4), 'A', round(d['kernel_ms']['loop_a_ms'],4), 'B', round(d['kernel_ms']['loop_b_ms'],4), c.get('sharded_vs_single_gpu',{}).get('match'), c.get('dudt_digest_u64'))" || tail -5 gpurun_out/ac_$1.err
}
run m44_s2 2
run m44_s0 0 --no-check
run m44_s1 1 --no-check
run m44_s2b 2 --no-check
run m28_s0 0 --M 28 --no-check
run m28_s1 1 --M 28 --no-check
run m28_s2 2 --M 28
run m28_s0b 0 --M 28 --no-check
timeout 300 python bench.py --M 28 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary --no-check > gpurun_out/ac_m28_n1.json 2> gpurun_out/ac_m28_n1.err
python -c "
import json; d=json.load(open('gpurun_out/ac_m28_n1.json')); print('M=28 N=1 ms/step', round(d['ms_per_step'],4), 'A', round(d['kernel_ms']['loop_a_ms'],4), 'B', round(d['kernel_ms']['loop_b_ms'],4))"
bash tools/gpu_ab.sh 20; el ab

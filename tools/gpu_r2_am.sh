#!/bin/bash
# Round 2, session AM (8 GPUs): driver-style bench at N = 8 on the final build
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 66 $T bench.py --gpus 8 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/am_n8.json 2> gpurun_out/am_n8.err; echo "bench N=8 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/am_n8.json')); c=d.get('check',{})
print('N=8 ms/step', round(d['ms_per_step'],4), 'value', d['value'], 'A', round(d['kernel_ms']['loop_a_ms'],4), 'B', round(d['kernel_ms']['loop_b_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e'].get('pcie_floor_ms'), c.get('sharded_vs_single_gpu',{}).get('match'), c.get('dudt_digest_u64'), c.get('e2e_digest_matches'))" || tail -5 gpurun_out/am_n8.err

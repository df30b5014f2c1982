#!/bin/bash
# Round 2, session J (N GPUs, default 4): sharded residual through the C ABI with two distinct peers
# per rank (periodic wrap), then the driver's bench command at that N
N=${N:-4}
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $T tools/dist_check.py > gpurun_out/r2_dist_check_${N}gpu.log 2>&1; el "dist_check rc=$?"; grep "world=" gpurun_out/r2_dist_check_${N}gpu.log | tail -8
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/j_n$N.json 2> gpurun_out/j_n$N.err; el "bench N=$N rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/j_n$N.json'))
print('N=$N ms/step', round(d['ms_per_step'],4), 'value', d['value'], 'e2e', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='flow'})
print('check', d['check']['dudt_digest_u64'], d['check']['sharded_vs_single_gpu'], d['check'].get('e2e_digest_matches'))
s=d['secondary']['cfg3']; print('cfg3', round(s['ms_per_step'],4), round(s['hbm_frac'],3), s['check']['dudt_digest_u64'] if s.get('check') else None)
" || tail -8 gpurun_out/j_n$N.err
el done

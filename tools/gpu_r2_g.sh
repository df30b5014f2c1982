#!/bin/bash
# Round 2, session G: the driver's two commands at N = 1, as the driver runs them
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/g_ref.json 2> gpurun_out/g_ref.err; el "reference arm rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/g_ref.json')); print('REF', d['value'], d['cpu_baseline']['cores'], d['ms_per_step'], d['run'])" || tail -5 gpurun_out/g_ref.err
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/g_n1.json 2> gpurun_out/g_n1.err; el "bench N=1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/g_n1.json'))
print('N1', d['value'], d['ms_per_step'], d['kernel_ms'], 'frac', d['roofline']['frac'])
print('e2e', d['e2e']); print('check', d['check']); print('cpu', d['cpu_baseline'])
s=d['secondary']; print('cfg3', {k:v for k,v in s['cfg3'].items() if k not in ('workload','check')})
for r in s['single_gpu']: print(r['config'][:40], round(r['ms_per_residual'],4), round(r['algorithmic_GBps']), round(r['hbm_frac'],3), r['setup_s'])
" || tail -8 gpurun_out/g_n1.err
el done

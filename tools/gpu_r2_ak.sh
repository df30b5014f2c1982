#!/bin/bash
# Round 2, session AK: final validation -- full GPU suite, smoke(), launch list, driver-style bench (both arms)
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_ak.log 2>&1; el "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_ak.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; el smoke
P="python bench.py --M 16 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_final2.csv $P > gpurun_out/ncu_lak.log 2>&1; el "launch list rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ak_ref.json 2> gpurun_out/ak_ref.err; el "reference arm rc=$?"
timeout 900 python bench.py > gpurun_out/ak_n1.json 2> gpurun_out/ak_n1.err; el "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/ak_n1.json'))
print('N=1 ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'A/B', round(d['kernel_ms']['loop_a_ms'],3), round(d['kernel_ms']['loop_b_ms'],3), 'fp64 frac', round(d['roofline']['frac'],4))
print('e2e', round(d['e2e']['ms_per_step'],3), 'floor', round(d['e2e']['pcie_floor_ms'],3), 'digest ok', d['check'].get('e2e_digest_matches'))
print('cfg3', round(d['secondary']['cfg3']['ms_per_step'],3), round(d['secondary']['cfg3']['hbm_frac'],3))
for x in d['secondary']['single_gpu']: print('  ', x['config'], round(x['ms_per_residual'],4), round(x['hbm_frac'],3))
r=json.load(open('gpurun_out/ak_ref.json')); print('reference arm', r['value'], r['cpu_baseline']['cores'])
" || tail -5 gpurun_out/ak_n1.err
el done

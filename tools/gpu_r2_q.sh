#!/bin/bash
# Round 2, session Q: final state -- GPU test suite, launch list, full captures of the three headline kernels,
# driver-style bench (ours + reference arm)
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_q.log 2>&1; el "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_q.log
P="python bench.py --M 16 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_final.csv $P > gpurun_out/ncu_lq.log 2>&1; el "launch list rc=$?"
for k in k_nodal_tensor k_fluxdiff_nodal k_project_tet; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/q_$k $P > gpurun_out/ncu_q_$k.log 2>&1; el "ncu $k rc=$?"
  bash tools/make_profile_md.sh /tmp/q_$k.ncu-rep "round 2 final, $k (Tet p=4 Euler), M=16 (24 576 elements); ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 $P" > gpurun_out/r2_final_$k.md
  [ -f /tmp/q_$k.ncu-rep ] && [ $(stat -c %s /tmp/q_$k.ncu-rep) -lt 12000000 ] && cp /tmp/q_$k.ncu-rep gpurun_out/
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/q_ref.json 2> gpurun_out/q_ref.err; el "reference arm rc=$?"
timeout 900 python bench.py > gpurun_out/q_n1.json 2> gpurun_out/q_n1.err; el "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/q_n1.json'))
print('N=1 ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'A/B', d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'], 'fp64 frac', round(d['roofline']['frac'],4))
print('e2e', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='flow'})
print('numa', d['run'].get('numa_rank0'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('cfg3', round(d['secondary']['cfg3']['ms_per_step'],3), round(d['secondary']['cfg3']['hbm_frac'],3))
for x in d['secondary']['single_gpu']: print('  ', x['config'], round(x['ms_per_residual'],4), round(x['hbm_frac'],3))
r=json.load(open('gpurun_out/q_ref.json')); print('reference arm', r['value'], r['cpu_baseline']['cores'])
" || tail -5 gpurun_out/q_n1.err
el done

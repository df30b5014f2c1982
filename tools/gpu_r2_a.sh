#!/bin/bash
# Round 2, GPU session A: full GPU suite on the multi-TU build, A/B of the loop-B flux changes,
# launch list in split mode (volume term alone) and --set full captures of loops A and B.
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_a.log 2>&1; el "gpu tests rc=$?"; tail -5 gpurun_out/gpu_tests_a.log
B="python bench.py --M 20 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
timeout 200 $B > gpurun_out/a_m20.json 2> gpurun_out/a_m20.err; el "bench M=20 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/a_m20.json')); print('M20', d['ms_per_step'], d['kernel_ms'], d['roofline']['frac'])"
SSE_B200_SPLIT_B=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_split_a.csv python bench.py --M 16 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check > gpurun_out/ncu_split.log 2>&1; el "split launch list rc=$?"
P="python bench.py --M 16 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fluxdiff_tensor -s 2 -c 1 -f -o /tmp/fluxdiff_a $P > gpurun_out/ncu_b.log 2>&1; el "ncu loop B rc=$?"
bash tools/make_profile_md.sh /tmp/fluxdiff_a.ncu-rep "round 2 session A, k_fluxdiff_tensor<3,5,Euler,collapsed,8> (loop B, half-velocity EC flux), M=16; ncu --set full --clock-control none --import-source on -k regex:k_fluxdiff_tensor -s 2 -c 1 $P" > gpurun_out/r2_fluxdiff_tensor_a.md; el "md loop B"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_nodal_tensor -s 2 -c 1 -f -o /tmp/nodal_a $P > gpurun_out/ncu_a.log 2>&1; el "ncu loop A rc=$?"
bash tools/make_profile_md.sh /tmp/nodal_a.ncu-rep "round 2 session A, k_nodal_tensor<3,5,Euler> (loop A), M=16; same command with -k regex:k_nodal_tensor" > gpurun_out/r2_nodal_tensor_a.md; el "md loop A"
python tools/ncu_line_hist.py /tmp/fluxdiff_a.ncu-rep 120 > gpurun_out/r2_fluxdiff_lines_a.txt 2>/dev/null
python tools/ncu_line_hist.py /tmp/nodal_a.ncu-rep 120 > gpurun_out/r2_nodal_lines_a.txt 2>/dev/null
ls -la /tmp/*.ncu-rep
for f in /tmp/fluxdiff_a.ncu-rep /tmp/nodal_a.ncu-rep; do [ $(stat -c %s $f) -lt 25000000 ] && cp $f gpurun_out/; done
el done

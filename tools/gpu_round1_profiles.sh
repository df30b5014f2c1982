#!/bin/bash
# Second GPU call of the round's end: launch list + --set full captures, summarised ON THE BOX
# (tools/make_profile_md.sh) so that only small text files travel back (the three .ncu-rep files
# of the first call exceeded gpurun's 64 MiB return limit and were dropped).
mkdir -p gpurun_out
S=$(date +%s)
el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
B="python bench.py --M 16 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v12.csv python bench.py --M 16 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; el "launch list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_fluxdiff_tensor -s 2 -c 1 -f -o /tmp/fluxdiff_v12 $B > gpurun_out/ncu_b.log 2>&1; el "ncu loop B rc=$?"
bash tools/make_profile_md.sh /tmp/fluxdiff_v12.ncu-rep "round 1 final (v12), k_fluxdiff_tensor<3,5,Euler,collapsed,8> (loop B), M=16 (24 576 elements); ncu --set full --clock-control none --import-source on -k regex:k_fluxdiff_tensor -s 2 -c 1 $B" > gpurun_out/r1_fluxdiff_tensor_v12.md; el "md loop B"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_nodal_tensor -s 2 -c 1 -f -o /tmp/nodal_v12 $B > gpurun_out/ncu_a.log 2>&1; el "ncu loop A rc=$?"
bash tools/make_profile_md.sh /tmp/nodal_v12.ncu-rep "round 1 final (v12), k_nodal_tensor<3,5,Euler> (loop A), M=16 (24 576 elements); same command with -k regex:k_nodal_tensor" > gpurun_out/r1_nodal_tensor_v12.md; el "md loop A"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_standard_tensor -s 2 -c 1 -f -o /tmp/standard_v12 env CFG3_M=16 python tools/bench_configs.py 3 > gpurun_out/ncu_s.log 2>&1; el "ncu cfg3 rc=$?"
bash tools/make_profile_md.sh /tmp/standard_v12.ncu-rep "round 1 final (v12), k_standard_tensor<3,5,adv,8,NB=4> (config 3 loop B), M=16 (24 576 elements); ncu --set full ... -k regex:k_standard_tensor env CFG3_M=16 python tools/bench_configs.py 3" > gpurun_out/r1_standard_tensor_v12.md; el "md cfg3"
ls -la /tmp/*.ncu-rep
CFG3_M=40 timeout 200 python tools/bench_configs.py 3 > gpurun_out/r1_cfg3_v12.json 2> gpurun_out/cfg3.err; el "cfg3 M=40 rc=$?"; cat gpurun_out/r1_cfg3_v12.json
du -sh gpurun_out
el done

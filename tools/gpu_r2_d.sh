#!/bin/bash
# Round 2, session D: profiles of the config-3 kernels and of loop B after the projection split
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2_launches_cfg3_d.csv env CFG3_M=24 python tools/bench_configs.py 3 > gpurun_out/ncu_l3.log 2>&1; el "cfg3 launch list rc=$?"
grep -v "^==" gpurun_out/r2_launches_cfg3_d.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | sort | uniq -c | sort -rn | head -6
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_standard_tensor -s 2 -c 1 -f -o /tmp/std_d env CFG3_M=24 python tools/bench_configs.py 3 > gpurun_out/ncu_s.log 2>&1; el "ncu cfg3 main rc=$?"
bash tools/make_profile_md.sh /tmp/std_d.ncu-rep "round 2 session D, k_standard_tensor<3,5,adv,8,NB=2,PROJ=false> (config 3 loop B without the projection tail), M=24 (82 944 elements)" > gpurun_out/r2_standard_tensor_d.md
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_project_tet -s 2 -c 1 -f -o /tmp/proj3_d env CFG3_M=24 python tools/bench_configs.py 3 > gpurun_out/ncu_p3.log 2>&1; el "ncu cfg3 project rc=$?"
bash tools/make_profile_md.sh /tmp/proj3_d.ncu-rep "round 2 session D, k_project_tet<5,1,5,5> (config 3 projection, 25 elements per CTA), M=24" > gpurun_out/r2_project_cfg3_d.md
P="python bench.py --M 16 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fluxdiff_nodal -s 2 -c 1 -f -o /tmp/fdn_d $P > gpurun_out/ncu_b.log 2>&1; el "ncu loop B rc=$?"
bash tools/make_profile_md.sh /tmp/fdn_d.ncu-rep "round 2 session D, k_fluxdiff_nodal<3,5,Euler,collapsed,8> (loop B up to the nodal residual), M=16" > gpurun_out/r2_fluxdiff_nodal_d.md
for f in /tmp/std_d.ncu-rep /tmp/proj3_d.ncu-rep /tmp/fdn_d.ncu-rep; do [ $(stat -c %s $f) -lt 20000000 ] && cp $f gpurun_out/; done
el done

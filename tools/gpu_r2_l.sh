#!/bin/bash
# Round 2, session L: where does k_physical (config 5) lose its time?  launch list + full capture, Tri p=4
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
python tools/cfg5_case.py 2 4 128; python tools/cfg5_case.py 2 4 256
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_cfg5_l.csv python tools/cfg5_case.py 2 4 256 2 > gpurun_out/ncu_l5.log 2>&1; el "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_physical -s 5 -c 1 -f -o /tmp/phys_l python tools/cfg5_case.py 2 4 256 2 > gpurun_out/ncu_p5.log 2>&1; el "ncu k_physical rc=$?"
bash tools/make_profile_md.sh /tmp/phys_l.ncu-rep "round 2 session L, k_physical<2,adv> stage 1 (config 5, Tri p=4, 131 072 elements; operators staged with cp.async)" > gpurun_out/r2_physical_l.md
cp /tmp/phys_l.ncu-rep gpurun_out/
el done

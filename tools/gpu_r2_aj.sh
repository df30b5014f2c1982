#!/bin/bash
# Round 2, session AJ: full ncu capture of the specialised loop-A kernel (report brought back for the source page)
mkdir -p gpurun_out
P="python bench.py --M 16 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
k=k_nodal_tensor
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/aj_$k $P > gpurun_out/ncu_aj_$k.log 2>&1; echo "ncu rc=$?"
bash tools/make_profile_md.sh /tmp/aj_$k.ncu-rep "round 2 session AJ, $k<3,5,Euler,PROJ_CT=2> (Tet p=4 Euler, entropy-projection instantiation), M=16 (24 576 elements); ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 $P" > gpurun_out/r2_aj_$k.md
for r in stall_no_inst stall_no_instruction stall_branch_resolving stall_dispatch stall_not_selected; do python tools/ncu_stall_lines.py /tmp/aj_$k.ncu-rep $r 8 2>/dev/null | cut -c1-150; done >> gpurun_out/r2_aj_$k.md
ls -la /tmp/aj_$k.ncu-rep; [ $(stat -c %s /tmp/aj_$k.ncu-rep) -lt 30000000 ] && cp /tmp/aj_$k.ncu-rep gpurun_out/
head -34 gpurun_out/r2_aj_$k.md

#!/bin/bash
# Round 2, session X: what bounds the warp-private projection kernel once the block barriers are gone?
mkdir -p gpurun_out
P="python bench.py --M 16 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
SSE_B200_PROJ_WARP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_project_tet_w -s 2 -c 1 -f -o /tmp/x_pw $P > gpurun_out/ncu_x_pw.log 2>&1; echo "rc=$?"
bash tools/make_profile_md.sh /tmp/x_pw.ncu-rep "round 2 session X, k_project_tet_w<5,5> (one element per warp, no block barriers), M=16" > gpurun_out/r2_project_tet_w_x.md
cp /tmp/x_pw.ncu-rep gpurun_out/
SSE_B200_LIB=$PWD/build/variants/pw_ew2_m3.so SSE_B200_PROJ_WARP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_project_tet_w -s 2 -c 1 -f -o /tmp/x_pw2 $P > gpurun_out/ncu_x_pw2.log 2>&1; echo "rc=$?"
bash tools/make_profile_md.sh /tmp/x_pw2.ncu-rep "round 2 session X, k_project_tet_w<5,5> two elements per warp, 3 CTAs/SM, M=16" > gpurun_out/r2_project_tet_w2_x.md

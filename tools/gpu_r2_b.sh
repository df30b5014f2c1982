#!/bin/bash
# Round 2, session B: A/B of the batched tensor-apply engine (loop A: k_nodal_tet; loop B split into
# k_fluxdiff_nodal + k_project_tet) against the one-element kernels, parity tests first.
mkdir -p gpurun_out
S=$(date +%s); el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_variants.py -m gpu -x -q -k "euler3d or full or split or entropy or fused or chunked" > gpurun_out/gpu_tests_b.log 2>&1; el "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests_b.log
B="python bench.py --M 20 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
for eng in 0 1 2 3; do
  SSE_B200_TET_ENGINE=$eng timeout 200 $B > gpurun_out/b_eng$eng.json 2> gpurun_out/b_eng$eng.err
  python -c "import json; d=json.load(open('gpurun_out/b_eng$eng.json')); print('engine=$eng M20', round(d['ms_per_step'],4), d['kernel_ms']['loop_a_ms'], d['kernel_ms']['loop_b_ms'])"
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_b.csv python bench.py --M 16 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check > gpurun_out/ncu_lb.log 2>&1; el "launch list rc=$?"
grep -v "^==" gpurun_out/r2_launches_b.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-260 | sort | uniq -c | sort -rn | head -8
P="python bench.py --M 16 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-check"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_nodal_tet -s 2 -c 1 -f -o /tmp/nodal_b $P > gpurun_out/ncu_a.log 2>&1; el "ncu loop A rc=$?"
bash tools/make_profile_md.sh /tmp/nodal_b.ncu-rep "round 2 session B, k_nodal_tet<5,5> (loop A on the batched engine), M=16" > gpurun_out/r2_nodal_tet_b.md
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_project_tet -s 2 -c 1 -f -o /tmp/project_b $P > gpurun_out/ncu_p.log 2>&1; el "ncu project rc=$?"
bash tools/make_profile_md.sh /tmp/project_b.ncu-rep "round 2 session B, k_project_tet<5,5,5>, M=16" > gpurun_out/r2_project_tet_b.md
for f in /tmp/nodal_b.ncu-rep /tmp/project_b.ncu-rep; do [ $(stat -c %s $f) -lt 25000000 ] && cp $f gpurun_out/; done
el done

#!/bin/bash
# One-shot GPU session for the end of round 1 (run under gpurun): new GPU tests, the headline
# bench line, the ncu launch list and --set full captures of both north-star kernels, config 3
# with/without the L2 prefetch, then the rest of the GPU suite.  Most important first: the call
# may be cut short by the remaining GPU budget.
mkdir -p gpurun_out
S=$(date +%s)
el() { echo "[t+$(( $(date +%s) - S ))s] $*"; }
el start; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 300 python -m pytest tests/test_gpu_sharded_emulation.py tests/test_golden_fixtures.py -m gpu -q > gpurun_out/t1.log 2>&1; el "t1 (sharded emulation + fixtures) rc=$?"; tail -4 gpurun_out/t1.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tet or hex or advdiff" > gpurun_out/t2.log 2>&1; el "t2 (tet/hex/advdiff parity) rc=$?"; tail -4 gpurun_out/t2.log
timeout 420 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; el "bench rc=$?"; cat gpurun_out/bench_v12.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v12.csv python bench.py --M 16 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; el "launch list rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_fluxdiff_tensor -s 2 -c 1 -f -o gpurun_out/fluxdiff_v12 python bench.py --M 16 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; el "ncu loop B rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_nodal_tensor -s 2 -c 1 -f -o gpurun_out/nodal_v12 python bench.py --M 16 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1; el "ncu loop A rc=$?"
CFG3_M=32 timeout 200 python tools/bench_configs.py 3 > gpurun_out/cfg3_v12.json 2> gpurun_out/cfg3.err; el "cfg3 rc=$?"; cat gpurun_out/cfg3_v12.json
CFG3_M=32 SSE_B200_PREFETCH=0 timeout 200 python tools/bench_configs.py 3 > gpurun_out/cfg3_v12_nopf.json 2>> gpurun_out/cfg3.err; el "cfg3 no prefetch rc=$?"; cat gpurun_out/cfg3_v12_nopf.json
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_standard_tensor -s 2 -c 1 -f -o gpurun_out/standard_v12 env CFG3_M=16 python tools/bench_configs.py 3 > gpurun_out/ncu_s.log 2>&1; el "ncu cfg3 rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_sharded_emulation.py > gpurun_out/t_all.log 2>&1; el "full gpu suite rc=$?"; tail -5 gpurun_out/t_all.log
el done

#!/bin/bash
# Round 2, session AH: host-buffer pipeline with chunks aligned to the mesh's x-layers (44 layers of 11 616
# elements at M=44: loop B of a chunk then needs only the two adjacent chunks)
mkdir -p gpurun_out
for v in "44 0" "22 0" "32 1" "44 1"; do
  set -- $v
  SSE_B200_HOST_TRACE=1 SSE_B200_HOST_CHUNKS=$1 SSE_B200_HOST_TAPER=$2 timeout 400 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-secondary --no-check > gpurun_out/ah_$1_$2.json 2> gpurun_out/ah_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/ah_$1_$2.json'))
print('chunks=$1 taper=$2: device %.3f ms, host-path %.3f ms (floor %.2f)' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pcie_floor_ms']))" || tail -3 gpurun_out/ah_$1_$2.err
  grep "sse host trace" gpurun_out/ah_$1_$2.err | tail -2 | cut -c1-420
done

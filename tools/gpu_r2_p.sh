#!/bin/bash
# Round 2, session P: is the chunked schedule itself slower than the monolithic kernels, or do the copies slow it down?
mkdir -p gpurun_out
for v in "1 32" "1 16" "1 8" "0 32"; do
  set -- $v
  SSE_B200_HOST_NOCOPY=$1 SSE_B200_HOST_CHUNKS=$2 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-check > gpurun_out/p_$1_$2.json 2> gpurun_out/p_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/p_$1_$2.json'))
print('nocopy=$1 chunks=$2: device %.3f ms, host-path %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']))" || tail -3 gpurun_out/p_$1_$2.err
done

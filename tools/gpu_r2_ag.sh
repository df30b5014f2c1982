#!/bin/bash
# Round 2, session AG: which copies cost the host-buffer pipeline its 1.8 ms?  SSE_B200_HOST_NOCOPY = 0 (all
# copies) / 1 (none) / 2 (no downloads) / 3 (no uploads), timeline of the default
mkdir -p gpurun_out
for m in 0 1 2 3; do
  SSE_B200_HOST_TRACE=$([ $m = 0 ] && echo 1 || echo 0) SSE_B200_HOST_NOCOPY=$m timeout 400 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-secondary --no-check > gpurun_out/ag_$m.json 2> gpurun_out/ag_$m.err
  python -c "
import json; d=json.load(open('gpurun_out/ag_$m.json'))
print('nocopy=$m: device %.3f ms, host-path %.3f ms (floor %.2f)' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pcie_floor_ms']))" || tail -3 gpurun_out/ag_$m.err
done
grep "sse host trace" gpurun_out/ag_0.err | tail -2 | cut -c1-1500

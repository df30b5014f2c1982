#!/bin/bash
# Round 2, session O: host-buffer pipeline with loop A on its own stream (on/off), with the e2e digest check
mkdir -p gpurun_out
for v in 1 0; do
  SSE_B200_HOST_TRACE=1 SSE_B200_HOST_ASTREAM=$v timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/o_a$v.json 2> gpurun_out/o_a$v.err
  python -c "
import json; d=json.load(open('gpurun_out/o_a$v.json'))
print('a_stream=$v: device %.3f ms, e2e %.3f ms, pcie floor %.3f, digest match %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pcie_floor_ms'], d['check'].get('e2e_digest_matches')))" || tail -3 gpurun_out/o_a$v.err
  grep "host trace" gpurun_out/o_a$v.err | tail -2 | cut -c1-900
done
timeout 600 python -m pytest tests -m gpu -x -q -k "host or parity" 2>&1 | tail -2

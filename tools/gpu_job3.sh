#!/bin/bash
# r02 session 2, GPU job 3: geometry staging change (all loads in flight) -- parity + timing; graph test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "geometry or graph or full_size or oil" > gpurun_out/j3_pytest.log 2>&1
tail -3 gpurun_out/j3_pytest.log
for g in 2 0; do
  ZEDO_GEOM=$g timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --oil-steps 200 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('geom=$g', d['value'], r['avg_launch_ms'], r['other_kernels_ms'], d['clocks'])"
done 2>&1 | tee gpurun_out/j3_geom_ab.log

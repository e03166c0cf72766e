#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "geometry or full_size or division" > gpurun_out/j4_pytest.log 2>&1
tail -2 gpurun_out/j4_pytest.log
for i in 1 2; do
  timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --oil-steps 200 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('prefetch', d['value'], r['avg_launch_ms'], r['other_kernels_ms'], d['clocks'])"
done 2>&1 | tee gpurun_out/j4_geom.log

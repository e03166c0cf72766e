#!/bin/bash
# r02c experiment matrix of the fp8lo hidden layer (ZEDO_CONVERT_HI8 bits: 1 = A hi8 on chip, 2 = W hi8 on chip,
# 4 = polling waits, 8 = half-block stages)
mkdir -p gpurun_out
for c in 8 11 9 15; do
  echo "== CONVERT_HI8=$c parity"
  ZEDO_CONVERT_HI8=$c timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "full_size or score_forward_vs_oracle or rows_are_independent" 2>&1 | tail -2
done 2>&1 | tee gpurun_out/jB_parity.log
lb() { ZEDO_CONVERT_HI8=$1 timeout 200 python tools/layer_bench.py 262144 30 fp8lo 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('conv$1', d['fp8lo/exp0'])"; }
for i in 1 2; do
  for c in 0 8 9 11 12 15 3; do lb $c; done
done 2>&1 | tee gpurun_out/jB_layer.log
for c in 0 8 11; do
  ZEDO_CONVERT_HI8=$c timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --oil-steps 200 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('conv=$c', d['value'], r['avg_launch_ms'], r['other_kernels_ms'], d['clocks'])"
done 2>&1 | tee gpurun_out/jB_loop.log

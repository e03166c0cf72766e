#!/bin/bash
mkdir -p gpurun_out
K32=$PWD/zedo_release_b200/libzedo_b200_k32.so
ZEDO_B200_LIB=$K32 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -8 > gpurun_out/jD_pytest_k32.log; cat gpurun_out/jD_pytest_k32.log
for i in 1 2; do
  timeout 200 python tools/layer_bench.py 262144 40 fp8lo,split3 0 2>&1 | tail -1
  ZEDO_B200_LIB=$K32 timeout 200 python tools/layer_bench.py 262144 40 fp8lo,split3 0 2>&1 | tail -1
done | tee gpurun_out/jD_layer.log
for lib in "" $K32; do
  ZEDO_B200_LIB=$lib timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --oil-steps 200 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('lib=$lib', d['value'], r['avg_launch_ms'], r['other_kernels_ms'], d['clocks'])"
done 2>&1 | tee gpurun_out/jD_loop.log
ZEDO_B200_LIB=$K32 timeout 200 python tools/small_batch.py 2>&1 | tail -3 | tee gpurun_out/jD_small_k32.log
timeout 200 python tools/small_batch.py 2>&1 | tail -3 | tee gpurun_out/jD_small_k64.log

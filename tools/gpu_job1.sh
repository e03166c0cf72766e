#!/bin/bash
# r02 session 2, GPU job 1: rays geometry parity + A/B timings (base = HEAD~ library, new = working tree)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/j1_pytest.log 2>&1
tail -5 gpurun_out/j1_pytest.log
BASE=$PWD/zedo_release_b200/libzedo_b200_base.so
lb() { python tools/layer_bench.py 262144 30 fp8lo 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['fp8lo/exp0'])"; }
for i in 1 2; do
  ZEDO_B200_LIB=$BASE lb base
  lb new_ew8
  ZEDO_LEAN_EW=16 lb new_ew16
  ZEDO_LEAN_EW=9 lb new_ew8lean
done 2>&1 | tee gpurun_out/j1_layer_ab.log
# geometry: old block kernel vs rays, inside a 200-step loop
for g in 2 0; do
  ZEDO_GEOM=$g timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --oil-steps 200 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('geom=$g', d['value'], r['avg_launch_ms'], r['other_kernels_ms'], d['clocks'])"
done 2>&1 | tee gpurun_out/j1_geom_ab.log
ZEDO_LEAN_EW=16 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --oil-steps 200 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ew16', d['value'], r['avg_launch_ms'], r['other_kernels_ms'], d['clocks'])" | tee -a gpurun_out/j1_geom_ab.log

#!/bin/bash
mkdir -p gpurun_out
python tools/small_batch.py 2>&1 | tail -1 | tee gpurun_out/j5_small_default.log
ZEDO_GEOM=3 python tools/small_batch.py 2>&1 | tail -1 | tee gpurun_out/j5_small_rays.log
ZEDO_GEOM=2 python tools/small_batch.py 2>&1 | tail -1 | tee gpurun_out/j5_small_block.log

#!/bin/bash
# r02 session 2, GPU job 2: full GPU test suite, ncu captures of the rays geometry kernel and the 16-warp first layer,
# per-kernel times at 1,024 poses
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/j2_pytest.log 2>&1
tail -4 gpurun_out/j2_pytest.log
ncu --set full --clock-control none --import-source on -k regex:oil_geom -c 4 -f -o gpurun_out/prof_oilgeom_r02b \
    python bench.py --steps 1 --warmup 0 --oil-steps 10 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:layer_tc_kernel -c 2 -f -o gpurun_out/prof_first_r02b \
    python bench.py --steps 1 --warmup 0 --oil-steps 4 --no-cpu > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/layer_bench.py 1024 50 fp8lo 0 2>&1 | tail -1 > gpurun_out/j2_layer_b1024.json
cat gpurun_out/j2_layer_b1024.json
python tools/small_batch.py 2>&1 | tail -1 | tee gpurun_out/j2_small_batch.log

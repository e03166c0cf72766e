#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02c_pytest_gpu.log; cat gpurun_out/r02c_pytest_gpu.log
bits="0,$((74<<8)),$((72<<8)),$((66<<8)),$((60<<8)),$((52<<8)),$((44<<8)),$((37<<8))"
ZEDO_B200_LIB=zedo_release_b200/libzedo_b200_exp.so timeout 300 python tools/layer_bench.py 262144 60 fp8lo $bits > gpurun_out/r02c_pairs_sweep.json 2>&1; cat gpurun_out/r02c_pairs_sweep.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

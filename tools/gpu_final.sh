#!/bin/bash
# end-of-round evidence on one GPU: full GPU test suite, evidence.sh (bench lines, ncu launch list and captures), sanitizer
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 1500 bash tools/evidence.sh $TAG 2>&1 | tail -12
timeout 900 bash tools/sanitize.sh 2>&1 | tail -8

#!/bin/bash
# N-GPU check of the product path: the multi-GPU tests, then the default bench line and C3 under torchrun
N=${1:-2}; TAG=${2:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_${N}gpu.log; cat gpurun_out/${TAG}_pytest_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus $N --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
python - gpurun_out/${TAG}_bench_n${N}.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "n_gpus", d["n_gpus"], "shard_ok", d["sharded_equals_unsharded_slice"], d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
tail -3 gpurun_out/${TAG}_bench_n${N}.err

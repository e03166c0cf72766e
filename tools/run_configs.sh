#!/bin/bash
# BASELINE.json configs through bench.py on N GPUs of this box: tools/run_configs.sh N "c3 c4 c5 c5h50" [steps] [warmup]
# One JSON line per config under gpurun_out/r2_configs/.  Evidence runs of the non-default configs (the driver's own
# BENCH / SCALE runs use the default c2 line).
N=${1:-8}; CONFIGS=${2:-"c3 c4 c5"}; STEPS=${3:-1}; WARM=${4:-1}
mkdir -p gpurun_out/r2_configs
port=29600
for c in $CONFIGS; do
  port=$((port+1))
  extra=""
  name=$c
  if [ "$c" = "c5h50" ]; then name=c5; extra="--hypo 50"; fi
  if [ "$c" = "c5mini" ]; then name=c5; extra="--dataset mini --joints 17 --net score"; fi
  out=gpurun_out/r2_configs/${c}_n${N}.json
  if [ "$N" = "1" ]; then
    python bench.py --config $name $extra --steps $STEPS --warmup $WARM --no-cpu > $out 2> ${out%.json}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --config $name $extra --steps $STEPS --warmup $WARM > $out 2> ${out%.json}.err
  fi
  python - "$out" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1),
          "gather_bytes", d["e2e"]["nccl_gather_bytes_per_step"], "shard_ok", d["sharded_equals_unsharded_slice"], "finite", d["results_finite"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done

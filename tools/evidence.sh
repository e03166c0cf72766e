#!/bin/bash
# One GPU, end of round: the numbers and captures DESIGN.md / profiles/ quote.  tools/evidence.sh r02b
TAG=${1:-r02b}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 2 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
python bench.py --mode split3 --steps 1 --warmup 2 --no-cpu > gpurun_out/${TAG}_bench_n1_split3.json 2>/dev/null
python bench.py --config c1 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c1_n1.json 2>/dev/null
python bench.py --config c5 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_bench_c5_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 0 --oil-steps 4 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:layer_tc -s 7 -c 6 -f -o gpurun_out/prof_layer_${TAG} \
    python bench.py --steps 1 --warmup 0 --oil-steps 4 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:oil_geom -s 2 -c 3 -f -o gpurun_out/prof_geom_${TAG} \
    python bench.py --steps 1 --warmup 0 --oil-steps 10 --no-cpu > /dev/null 2>&1
python tools/driver_loop_timing.py 1024 1000 > gpurun_out/${TAG}_dropin_driver_timing.json 2>/dev/null
python tools/graph_probe.py > gpurun_out/${TAG}_graph_probe.json 2>/dev/null
for f in gpurun_out/${TAG}_bench_n1.json gpurun_out/${TAG}_bench_reference_arm.json gpurun_out/${TAG}_bench_n1_split3.json gpurun_out/${TAG}_bench_c1_n1.json gpurun_out/${TAG}_bench_c5_n1.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], round(d["value"], 1), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
ls -la gpurun_out/*${TAG}*

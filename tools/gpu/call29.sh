#!/bin/bash
# the full default bench line on the final sources
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 280 python bench.py --steps 20 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "rc=$?" >> gpurun_out/r2y_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2y_bench.json")); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["frac"], r["traffic"], d["e2e"]["value"], d["c4"]["value"], d["c4"]["e2e"]["value"], d["c5"]["value"], d["c5"]["e2e"]["value"], d["gpu_launches"], d["cpu_baseline"]["value"])
PY
tail -n 2 gpurun_out/r2y_bench.err

#!/bin/bash
# bench line after moving nvidia-smi's start-up into the warm-up
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-composites > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "rc=$?" >> gpurun_out/r2w_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2w_bench.json")); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["fwd_ms"], r["inv_ms"], r["frac"], r["traffic"], d["clocks"], d["e2e"]["value"], d["e2e"]["clocks"])
PY
tail -n 2 gpurun_out/r2w_bench.err

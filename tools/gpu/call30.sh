#!/bin/bash
# bench line (transforms + CPU baseline only) after the CPU-sample change
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 100 python bench.py --steps 20 --warmup 3 --no-composites --no-eltwise --no-e2e > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "rc=$?" >> gpurun_out/r2z_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2z_bench.json')); print(d['value'], d['roofline']['traffic'], d['cpu_baseline'])"
tail -n 2 gpurun_out/r2z_bench.err

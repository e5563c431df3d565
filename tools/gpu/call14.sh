#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_north_star.py -m gpu -x -q -k "sharded" > $O/r2l_pytest.txt 2>&1; echo "rc=$?" >> $O/r2l_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-eltwise --no-cpu > $O/r2l_bench_1proc.json 2> $O/r2l_bench_1proc.err
tail -n 3 $O/r2l_pytest.txt; python - <<PY
import json
d=json.load(open("$O/r2l_bench_1proc.json")); print("1proc c5 latency:", d["c5"].get("latency_one_switch_host_buffers"), d["c5"]["ms_per_key_switch"])
PY
tail -n 3 $O/r2l_bench_1proc.err

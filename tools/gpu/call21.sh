#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "key_switch or sharded or resident or multi or rns" > $O/r2q_pytest.txt 2>&1; echo "rc=$?" >> $O/r2q_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-eltwise > $O/r2q_bench.json 2> $O/r2q_bench.err
timeout 150 python tools/stress.py 90 17 > $O/r2q_stress.log 2>&1
tail -n 3 $O/r2q_pytest.txt; python - <<PY
import json
d=json.load(open("$O/r2q_bench.json")); print("c5:", d["c5"]["ms_per_key_switch"], d["c5"]["value"], d["c5"]["e2e"], d["c5"].get("parity")); print("c4", d["c4"]["value"])
PY
tail -n 2 $O/r2q_stress.log

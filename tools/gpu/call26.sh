#!/bin/bash
# last pass of the round: full GPU suite, ncu re-capture summarised ON the box so that the bench line that follows
# carries the DRAM traffic of exactly these sources, both bench arms, smoke
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2v_pytest.txt 2>&1; echo "rc=$?" >> $O/r2v_pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file $O/r2v_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-composites --no-eltwise > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 8 -c 4 -o /tmp/prof_r2v -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-composites --no-eltwise > $O/r2v_ncu.log 2>&1
ncu -i /tmp/prof_r2v.ncu-rep --page raw --csv > $O/r2v_ncu_raw.csv 2>/dev/null
python tools/summarize_ncu.py r2v $O/r2v_launches.csv $O/r2v_ncu_raw.csv --batch 8192 > $O/r2v_summarize.log 2>&1
cp profiles/traffic.json $O/r2v_traffic.json; cp profiles/r2v_launches.md profiles/r2v_ncu.md $O/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2v_bench.json 2> $O/r2v_bench.err; echo "rc=$?" >> $O/r2v_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2v_bench_ref.json 2> $O/r2v_bench_ref.err
python __graft_entry__.py smoke > $O/r2v_smoke.log 2>&1
tail -n 3 $O/r2v_pytest.txt; tail -n 3 $O/r2v_summarize.log; python - <<PY
import json
d=json.load(open("$O/r2v_bench.json")); print(d["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["c4"]["value"], d["c5"]["value"], d["gpu_launches"])
PY
tail -n 2 $O/r2v_bench.err; head -c 200 $O/r2v_bench_ref.json; echo; tail -n 1 $O/r2v_smoke.log

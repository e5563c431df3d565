#!/bin/bash
# product multiplied on load: parity, timing against the unfused chain, bench + ncu re-capture on the new sources
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2t_pytest.txt 2>&1; echo "rc=$?" >> $O/r2t_pytest.txt
timeout 300 python tools/c4_fusion.py > $O/r2t_c4_fusion.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2t_bench.json 2> $O/r2t_bench.err; echo "rc=$?" >> $O/r2t_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file $O/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-composites --no-eltwise > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 8 -c 4 -o /tmp/prof_r2t -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-composites --no-eltwise > $O/r2t_ncu.log 2>&1
ncu -i /tmp/prof_r2t.ncu-rep --page raw --csv > $O/r2t_ncu_raw.csv 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/r2t_san_memcheck.log 2>&1
timeout 200 python tools/stress.py 90 29 > $O/r2t_stress.log 2>&1
python __graft_entry__.py smoke > $O/r2t_smoke.log 2>&1
tail -n 3 $O/r2t_pytest.txt; cat $O/r2t_c4_fusion.txt; head -c 300 $O/r2t_bench.json; echo; tail -n 2 $O/r2t_bench.err; tail -n 2 $O/r2t_san_memcheck.log; tail -n 1 $O/r2t_stress.log; tail -n 1 $O/r2t_smoke.log

#!/bin/bash
# round 2: full verification pass -- tests, bench (both arms), launch list, full ncu capture (CSV pages), sanitizer, stress
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2g_pytest.txt 2>&1; echo "rc=$?" >> $O/r2g_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2g_bench.json 2> $O/r2g_bench.err; echo "rc=$?" >> $O/r2g_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2g_bench_ref.json 2> $O/r2g_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-composites --no-eltwise > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 8 -c 4 -o /tmp/prof_r2g -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-composites --no-eltwise > $O/r2g_ncu.log 2>&1
ncu -i /tmp/prof_r2g.ncu-rep --page raw --csv > $O/r2g_ncu_raw.csv 2>/dev/null
for tool in memcheck racecheck; do
  HEXL_B200_PIPE_MIN_BATCH=1 HEXL_B200_PIPE=1 timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > $O/r2g_san_${tool}_pipe.log 2>&1
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > $O/r2g_san_${tool}.log 2>&1
done
timeout 400 python tools/stress.py 240 7 > $O/r2g_stress.log 2>&1
tail -n 3 $O/r2g_pytest.txt; cat $O/r2g_bench.json | head -c 6000; echo; tail -n 3 $O/r2g_bench.err; cat $O/r2g_bench_ref.json | head -c 1500; echo; tail -n 4 $O/r2g_san_*.log; tail -n 5 $O/r2g_stress.log

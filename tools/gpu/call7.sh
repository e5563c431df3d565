#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
M=gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r2f_bfly_ncu.csv ./tools/bin/bfly_bench > $O/r2f_bfly.txt 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:OpWideHiUse\|OpImad\|OpWN\|OpMixNow\|OpWNA --csv --log-file $O/r2f_pipe_ncu.csv ./tools/bin/pipe_bench > /dev/null 2>&1
./tools/bin/latency > $O/r2f_latency.txt 2>&1
./tools/bin/bfly_bench 2>&1 | grep -E "v16|v14|v5 " > $O/r2f_bfly_plain.txt
timeout 600 python -m pytest tests/test_montgomery.py -m gpu -x -q > $O/r2f_pytest.txt 2>&1; echo "rc=$?" >> $O/r2f_pytest.txt
cat $O/r2f_latency.txt $O/r2f_bfly_plain.txt; tail -n 5 $O/r2f_pytest.txt

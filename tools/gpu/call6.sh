#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
export HEXL_B200_PIPE_MIN_BATCH=1
: > $O/r2e_tune.txt
for la in 48 64 128; do HEXL_B200_PIPE=1 HEXL_B200_PIPE_LOOKAHEAD=$la python tools/tune_split.py 15 16 17 55 29 60 >> $O/r2e_tune.txt 2>&1; done
# DRAM traffic and time, split vs pipelined (one forward + one inverse launch each)
for pipe in 0 1; do
  HEXL_B200_PIPE=$pipe HEXL_B200_PIPE_LOOKAHEAD=64 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:ntt_ -s 8 -c 4 --csv --log-file $O/r2e_traffic_pipe$pipe.csv python tools/tune_split.py 16 > /dev/null 2>&1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or montgomery" > $O/r2e_pytest.txt 2>&1; echo "rc=$?" >> $O/r2e_pytest.txt
./tools/bin/latency > $O/r2e_latency.txt 2>&1
cat $O/r2e_tune.txt; cat $O/r2e_traffic_pipe0.csv $O/r2e_traffic_pipe1.csv | grep -v "^==" | cut -c1-250; tail -n 4 $O/r2e_pytest.txt; cat $O/r2e_latency.txt

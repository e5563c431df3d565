#!/bin/bash
# round 2, GPU call 2: micro-benchmarks (lost in call 1), latency probe, ncu of default vs fp64q3 row kernels (CSV pages only)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
./tools/bin/pipe_bench > $O/r2a_pipe_bench.txt 2>&1
./tools/bin/bfly_bench > $O/r2a_bfly_bench.txt 2>&1
./tools/bin/latency > $O/r2a_latency.txt 2>&1
for v in default fp64q3; do
  if [ $v = default ]; then unset HEXL_B200_LIB; else export HEXL_B200_LIB=$PWD/hexl_b200/lib/libhexl_b200_$v.so; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_row_fwd -s 2 -c 1 -o /tmp/prof_$v -f python tools/tune_split.py 16 > $O/r2a_ncu_$v.log 2>&1
  ncu -i /tmp/prof_$v.ncu-rep --page raw --csv > $O/r2a_ncu_${v}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$v.ncu-rep --page source --csv > $O/r2a_ncu_${v}_source.csv 2>/dev/null
done
cat $O/r2a_pipe_bench.txt; cat $O/r2a_bfly_bench.txt; cat $O/r2a_latency.txt

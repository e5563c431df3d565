#!/bin/bash
# round 2, GPU call 1: pipe/butterfly micro-benchmarks, FP64-assisted quotient variants on the real kernels, parity tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt
./tools/bin/pipe_bench > $O/r2a_pipe_bench.txt 2>&1
./tools/bin/bfly_bench > $O/r2a_bfly_bench.txt 2>&1
python tools/tune_split.py 12 14 16 17 55 > $O/r2a_tune_default.txt 2>&1
for v in 1 2 3; do
  HEXL_B200_LIB=$PWD/hexl_b200/lib/libhexl_b200_fp64q$v.so python tools/tune_split.py 12 14 16 17 55 > $O/r2a_tune_fp64q$v.txt 2>&1
  HEXL_B200_LIB=$PWD/hexl_b200/lib/libhexl_b200_fp64q$v.so timeout 600 python tests/variant_check.py > $O/r2a_check_fp64q$v.txt 2>&1
  echo "variant_check fp64q$v rc=$?" >> $O/r2a_check_fp64q$v.txt
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2a_pytest.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_row_fwd -s 2 -c 1 -o $O/prof_r2a_default -f python tools/tune_split.py 16 > $O/r2a_ncu_default.log 2>&1
HEXL_B200_LIB=$PWD/hexl_b200/lib/libhexl_b200_fp64q3.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_row_fwd -s 2 -c 1 -o $O/prof_r2a_fp64q3 -f python tools/tune_split.py 16 > $O/r2a_ncu_fp64q3.log 2>&1
tail -3 $O/r2a_pytest.txt; cat $O/r2a_tune_default.txt $O/r2a_tune_fp64q*.txt; tail -2 $O/r2a_check_fp64q*.txt

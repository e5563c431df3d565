#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
python tools/tune_split.py 12 16 17 55 > $O/r2c_tune.txt 2>&1
HEXL_B200_LIB=$PWD/hexl_b200/lib/libhexl_b200_mb4.so python tools/tune_split.py 12 16 17 55 >> $O/r2c_tune.txt 2>&1
HEXL_B200_SPLIT_ROW_LOG=11 python tools/tune_split.py 16 17 55 >> $O/r2c_tune.txt 2>&1
HEXL_B200_SPLIT_ROW_LOG=13 python tools/tune_split.py 16 17 55 >> $O/r2c_tune.txt 2>&1
HEXL_B200_LIB=$PWD/hexl_b200/lib/libhexl_b200_mb4.so HEXL_B200_SPLIT_ROW_LOG=11 python tools/tune_split.py 16 17 55 >> $O/r2c_tune.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_north_star.py -m gpu -x -q > $O/r2c_pytest.txt 2>&1; echo "rc=$?" >> $O/r2c_pytest.txt
cat $O/r2c_tune.txt; tail -n 25 $O/r2c_pytest.txt

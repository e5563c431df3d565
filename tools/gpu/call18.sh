#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
python tools/tune_split.py 12 16 17 55 > $O/r2n_tune.txt 2>&1
python tools/tune_split.py 15 16 17 60 >> $O/r2n_tune.txt 2>&1
python tools/tune_split.py 16 62 >> $O/r2n_tune.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2n_pytest.txt 2>&1; echo "rc=$?" >> $O/r2n_pytest.txt
timeout 200 python tools/stress.py 100 13 > $O/r2n_stress.log 2>&1
cat $O/r2n_tune.txt; tail -n 3 $O/r2n_pytest.txt; tail -n 2 $O/r2n_stress.log

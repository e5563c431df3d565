#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
python tools/tune_split.py 10 12 13 14 16 17 55 > $O/r2m_tune.txt 2>&1
python tools/tune_split.py 16 60 29 >> $O/r2m_tune.txt 2>&1
HEXL_B200_PIPE=1 python tools/tune_split.py 16 17 55 >> $O/r2m_tune.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2m_pytest.txt 2>&1; echo "rc=$?" >> $O/r2m_pytest.txt
timeout 300 python tools/stress.py 150 11 > $O/r2m_stress.log 2>&1
cat $O/r2m_tune.txt; tail -n 3 $O/r2m_pytest.txt; tail -n 2 $O/r2m_stress.log

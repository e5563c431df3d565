#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
python tools/tune_split.py 12 13 14 16 17 55 > $O/r2b_tune_default.txt 2>&1
python tools/tune_split.py 16 60 50 >> $O/r2b_tune_default.txt 2>&1
timeout 600 python tests/variant_check.py > $O/r2b_check.txt 2>&1; echo "rc=$?" >> $O/r2b_check.txt
./tools/bin/bfly_bench 2>&1 | grep -E "v5|v12|v13|v14|v15" > $O/r2b_bfly.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.txt 2>&1; echo "rc=$?" >> $O/r2b_pytest.txt
cat $O/r2b_tune_default.txt $O/r2b_bfly.txt; tail -3 $O/r2b_check.txt $O/r2b_pytest.txt

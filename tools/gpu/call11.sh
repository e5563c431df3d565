#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_north_star.py -m gpu -x -q -k "sharded or resident" > $O/r2j_pytest.txt 2>&1; echo "rc=$?" >> $O/r2j_pytest.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/r2j_san.log 2>&1
tail -n 30 $O/r2j_pytest.txt; tail -n 5 $O/r2j_san.log

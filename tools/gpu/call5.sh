#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
export HEXL_B200_PIPE_MIN_BATCH=1
HEXL_B200_PIPE=1 timeout 600 python tests/variant_check.py > $O/r2d_check_pipe.txt 2>&1; echo "rc=$?" >> $O/r2d_check_pipe.txt
HEXL_B200_PIPE=1 HEXL_B200_PIPE_LOOKAHEAD=1 HEXL_B200_PIPE_CTAS=1 timeout 600 python tests/variant_check.py > $O/r2d_check_pipe1.txt 2>&1; echo "rc=$?" >> $O/r2d_check_pipe1.txt
: > $O/r2d_tune.txt
HEXL_B200_PIPE=0 python tools/tune_split.py 14 15 16 17 55 29 >> $O/r2d_tune.txt 2>&1
HEXL_B200_PIPE=1 python tools/tune_split.py 14 15 16 17 55 29 >> $O/r2d_tune.txt 2>&1
for la in 4 8 32; do HEXL_B200_PIPE=1 HEXL_B200_PIPE_LOOKAHEAD=$la python tools/tune_split.py 16 55 29 >> $O/r2d_tune.txt 2>&1; done
for c in 2 4; do HEXL_B200_PIPE=1 HEXL_B200_PIPE_CTAS=$c python tools/tune_split.py 16 55 29 >> $O/r2d_tune.txt 2>&1; done
cat $O/r2d_tune.txt; tail -n 3 $O/r2d_check_pipe.txt $O/r2d_check_pipe1.txt

#!/bin/bash
# 2 GPUs: the cross-GPU host split tests and the 2-rank bench (both arms)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2h_gpus.txt; nvidia-smi topo -m >> $O/r2h_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_north_star.py -m gpu -x -q > $O/r2h_pytest.txt 2>&1; echo "rc=$?" >> $O/r2h_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2h_bench_2gpu.json 2> $O/r2h_bench_2gpu.err; echo "rc=$?" >> $O/r2h_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r2h_bench_ref_2gpu.json 2> /dev/null
tail -n 4 $O/r2h_pytest.txt; head -c 5000 $O/r2h_bench_2gpu.json; echo; tail -n 5 $O/r2h_bench_2gpu.err; head -c 600 $O/r2h_bench_ref_2gpu.json

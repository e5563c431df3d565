#!/bin/bash
# multi-GPU box: sharded key switch across GPUs, single-process latency leg, N-rank bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
G=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $O/r2s_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_north_star.py -m gpu -x -q -k "shard or devices or split" > $O/r2s_pytest.txt 2>&1; echo "rc=$?" >> $O/r2s_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-eltwise > $O/r2s_bench_1proc.json 2> $O/r2s_bench_1proc.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $G --steps 10 --warmup 3 > $O/r2s_bench_${G}gpu.json 2> $O/r2s_bench_${G}gpu.err; echo "rc=$?" >> $O/r2s_bench_${G}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $G --steps 3 --warmup 1 > $O/r2s_bench_ref_${G}gpu.json 2> /dev/null
tail -n 4 $O/r2s_pytest.txt; python - <<PY
import json
d=json.load(open("$O/r2s_bench_1proc.json")); print("1proc c5:", json.dumps(d["c5"])[:1500])
d=json.load(open("$O/r2s_bench_${G}gpu.json")); print("N=$G value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("link_GBps_each_way_per_gpu"), "c4", d["c4"]["value"], d["c4"]["e2e"], "c5", d["c5"]["value"], d["c5"]["e2e"])
PY
head -c 700 $O/r2s_bench_ref_${G}gpu.json; tail -n 3 $O/r2s_bench_${G}gpu.err

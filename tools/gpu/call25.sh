#!/bin/bash
# full GPU suite on the final sources
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2u_pytest.txt
tail -n 6 gpurun_out/r2u_pytest.txt

#!/bin/bash
# racecheck on the final sources (fused product kernels included)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r2x_san_racecheck.log 2>&1
tail -n 3 gpurun_out/r2x_san_racecheck.log

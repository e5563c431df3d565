#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
bash tools/ablate.sh > gpurun_out/r2i_ablate.txt 2>&1
cat gpurun_out/r2i_ablate.txt

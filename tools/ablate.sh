#!/bin/bash
# Timing-only ablation of the 64-bit row kernels (N = 2^16, 55-bit, 4096 polynomials): which part of the kernel
# costs what.  Needs the variant libraries:  bash tools/build_variant.sh "_ab1=-DHEXL_B200_ABLATE=1" ...
cd "$(dirname "$0")/.."
python tools/tune_split.py 16 55
for v in 1 3 7 15 31 4 8 16; do
  echo "ablate=$v (1 no loads, 2 no stores, 4 last-pass twiddles not from L2, 8 no smem exchanges, 16 no smem twiddle tables)"
  HEXL_B200_LIB=$PWD/hexl_b200/lib/libhexl_b200_ab$v.so python tools/tune_split.py 16 55
done

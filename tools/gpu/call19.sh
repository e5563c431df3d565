#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or multi or rns" > $O/r2o_pytest.txt 2>&1; echo "rc=$?" >> $O/r2o_pytest.txt
python tools/tune_split.py 16 17 60 > $O/r2o_tune.txt 2>&1
HEXL_B200_PIPE=0 python tools/tune_split.py 17 60 55 >> $O/r2o_tune.txt 2>&1
# C4 shape through the multi-modulus call: pipelined forward (default rule) vs split
python - >> $O/r2o_tune.txt 2>&1 <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import hexl_b200 as hb
n, nmod, group = 1 << 17, 16, 32
mods = hb.GeneratePrimes(nmod, 60, True, n)
ntts = [hb.NTT(n, q) for q in mods]
sz = n * group
a = torch.empty(nmod * sz, dtype=torch.int64, device="cuda"); b = torch.empty_like(a); r = torch.empty_like(a)
for i, q in enumerate(mods):
    a[i*sz:(i+1)*sz].random_(0, q); b[i*sz:(i+1)*sz].random_(0, q)
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("PIPE env", os.environ.get("HEXL_B200_PIPE"), "c4 poly multiply ms", t(lambda: hb.PolyMultiplyMulti(ntts, r, a, b, group)),
      "fwd multi ms", t(lambda: hb.ComputeForwardMulti(ntts, r, a, 1, 4, batch_per_modulus=group)),
      "inv multi ms", t(lambda: hb.ComputeInverseMulti(ntts, r, a, 1, 1, batch_per_modulus=group)))
PY
HEXL_B200_PIPE=0 python - >> $O/r2o_tune.txt 2>&1 <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import hexl_b200 as hb
n, nmod, group = 1 << 17, 16, 32
mods = hb.GeneratePrimes(nmod, 60, True, n)
ntts = [hb.NTT(n, q) for q in mods]
sz = n * group
a = torch.empty(nmod * sz, dtype=torch.int64, device="cuda"); b = torch.empty_like(a); r = torch.empty_like(a)
for i, q in enumerate(mods):
    a[i*sz:(i+1)*sz].random_(0, q); b[i*sz:(i+1)*sz].random_(0, q)
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("PIPE env", os.environ.get("HEXL_B200_PIPE"), "c4 poly multiply ms", t(lambda: hb.PolyMultiplyMulti(ntts, r, a, b, group)),
      "fwd multi ms", t(lambda: hb.ComputeForwardMulti(ntts, r, a, 1, 4, batch_per_modulus=group)),
      "inv multi ms", t(lambda: hb.ComputeInverseMulti(ntts, r, a, 1, 1, batch_per_modulus=group)))
PY
cat $O/r2o_tune.txt; tail -n 4 $O/r2o_pytest.txt

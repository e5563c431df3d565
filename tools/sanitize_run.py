#!/usr/bin/env python
"""Small workload touching every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hexl_b200 as hb  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for logn, bits, batch in ((3, 20, 3), (6, 29, 5), (10, 29, 5), (10, 55, 5), (11, 60, 3), (12, 55, 2), (12, 29, 2),
                          (13, 50, 1), (14, 29, 1), (14, 61, 1), (15, 29, 1), (16, 55, 1), (16, 29, 1), (17, 29, 1),
                          (17, 50, 1)):
    n = 1 << logn
    q = hb.GeneratePrimes(1, bits, True, n)[0]
    t = hb.NTT(n, q)
    x = torch.randint(0, q, (batch * n,), dtype=torch.int64, device="cuda", generator=g)
    y = torch.empty_like(x)
    t.ComputeForward(y, x, 1, 1)
    t.ComputeInverse(y, y, 1, 1)
    assert bool((y == x).all()), (logn, bits)
n = 1 << 12
mods = hb.GeneratePrimes(3, 50, True, n)
ntts = [hb.NTT(n, m) for m in mods]
a = torch.cat([torch.randint(0, m, (2 * n,), dtype=torch.int64, device="cuda", generator=g) for m in mods])
b = torch.cat([torch.randint(0, m, (2 * n,), dtype=torch.int64, device="cuda", generator=g) for m in mods])
o = torch.empty_like(a)
hb.PolyMultiplyMulti(ntts, o, a, b, 2)
hb.EltwiseFMAMod(o, a, 5, b, a.numel(), mods[0], 1)
hb.EltwiseReduceMod(o, a, a.numel(), mods[0], mods[0], 1)
out = torch.zeros(3 * n * 3, dtype=torch.int64, device="cuda")
hb.DyadicMultiply(out, a, b, n, mods)
decomp, kcc = 2, 2
kms = rns = decomp + 1
keys = [torch.cat([torch.randint(0, mods[i], (n,), dtype=torch.int64, device="cuda", generator=g)
                   for _ in range(kcc) for i in range(kms)]) for _ in range(decomp)]
tt = torch.cat([torch.randint(0, mods[j], (n,), dtype=torch.int64, device="cuda", generator=g) for j in range(decomp)])
res = torch.cat([torch.randint(0, mods[i], (n,), dtype=torch.int64, device="cuda", generator=g)
                 for _ in range(kcc) for i in range(decomp)])
hb.KeySwitch(res, tt, n, decomp, kms, rns, kcc, mods, keys, [hb.InverseMod(mods[-1] % mods[i], mods[i]) for i in range(decomp)])
h = np.arange(n, dtype=np.uint64) % np.uint64(mods[0])
ntts[0].ComputeForward(h, h, 1, 1)  # host-pointer staging path
# round 2: resident keys (device and host buffers, two ciphertexts), composite host paths, Montgomery helpers
modswitch = [hb.InverseMod(mods[-1] % mods[i], mods[i]) for i in range(decomp)]
kh = hb.KeySwitchKeys(keys, n, decomp, kms, kcc)
hb.KeySwitchResident(torch.cat([res, res]), torch.cat([tt, tt]), n, decomp, kms, rns, kcc, mods, kh, modswitch, 2)
hres = np.concatenate([res.cpu().numpy().view(np.uint64)] * 2)
hb.KeySwitchResident(hres, np.concatenate([tt.cpu().numpy().view(np.uint64)] * 2), n, decomp, kms, rns, kcc, mods, kh, modswitch, 2)
hb.set_host_devices([0, 0])
khs = hb.KeySwitchKeys(keys, n, decomp, kms, kcc, sharded_by_modulus=True)   # two shards on one GPU: peer copies + events
hb.set_host_devices([])
hres2 = np.concatenate([res.cpu().numpy().view(np.uint64)] * 2)
hb.KeySwitchResident(hres2, np.concatenate([tt.cpu().numpy().view(np.uint64)] * 2), n, decomp, kms, rns, kcc, mods, khs, modswitch, 2)
assert (hres2 == hres).all()
ha, hbb = a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64)
ho = np.zeros_like(ha)
hb.PolyMultiplyMulti(ntts, ho, ha, hbb, 2)
hb.ComputeForwardMulti(ntts, ho, ha, 1, 4, batch_per_modulus=2)
hb.EltwiseMultModMulti(ho, ha, hbb, 2 * n, mods)
hout = np.zeros(3 * n * 3, dtype=np.uint64)
hb.DyadicMultiply(hout, ha, hbb, n, mods)
qm, r = mods[0], 61
inv = hb.HenselLemma2adicRoot(r, qm)
hb.EltwiseMontgomeryFormIn(o, a, (1 << r) * (1 << r) % qm, 2 * n, qm, r, inv)
hb.EltwiseMontReduceMod(o, o, b, 2 * n, qm, r, inv)
hb.EltwiseMontgomeryFormOut(o, o, 2 * n, qm, r, inv)
torch.cuda.synchronize()
print("sanitize workload done")

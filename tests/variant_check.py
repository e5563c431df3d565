"""Parity of the NTT kernels under the HEXL_B200_* launch knobs of the calling environment
(run by tests/test_gpu_parity.py::test_ntt_kernel_variants, one process per setting)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import hexl_b200 as hb  # noqa: E402
import oracle  # noqa: E402
from util import uniform_below  # noqa: E402

checker = oracle.best_checker()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


for logn in (12, 14, 15, 16, 17):
    n = 1 << logn
    for bits in (29, 33, 50, 55, 60, 61):
        q = hb.GeneratePrimes(1, bits, True, n)[0]
        t = hb.NTT(n, q)
        batch = 5
        x = uniform_below(logn + bits, n * batch, q)
        o = dev(np.zeros_like(x))
        t.ComputeForward(o, dev(x), 1, 1)
        assert (host(o) == checker.ntt_forward(x, n, q)).all(), ("fwd", logn, bits)
        t.ComputeForward(o, dev(x), 1, 4)
        g = host(o)
        assert (g % np.uint64(q) == checker.ntt_forward(x, n, q)).all() and (g < np.uint64(4 * q)).all()
        t.ComputeInverse(o, dev(x), 1, 1)
        assert (host(o) == checker.ntt_inverse(x, n, q)).all(), ("inv", logn, bits)
        d = dev(x)
        t.ComputeForward(d, d, 1, 1)
        t.ComputeInverse(d, d, 1, 1)
        assert (host(d) == x).all(), ("round trip in place", logn, bits)
# extreme inputs: every coefficient at the top of its allowed range (largest lazy growth inside the kernels)
for logn, bits in ((12, 55), (16, 55), (16, 33), (17, 50), (20, 55)):
    n = 1 << logn
    q = hb.GeneratePrimes(1, bits, True, n)[0]
    t = hb.NTT(n, q)
    for in_mf in (1, 4):
        x = np.full(n * 2, q * in_mf - 1, dtype=np.uint64)
        x[n:] = uniform_below(logn, n, q * in_mf)
        x[n] = 0
        o = dev(np.zeros_like(x))
        t.ComputeForward(o, dev(x), in_mf, 1)
        assert (host(o) == checker.ntt_forward(x, n, q, in_mf, 1)).all(), ("fwd extreme", logn, bits, in_mf)
    for in_mf in (1, 2):
        x = np.full(n * 2, q * in_mf - 1, dtype=np.uint64)
        x[n:] = uniform_below(logn + 1, n, q * in_mf)
        o = dev(np.zeros_like(x))
        t.ComputeInverse(o, dev(x), in_mf, 1)
        assert (host(o) == checker.ntt_inverse(x, n, q, in_mf, 1)).all(), ("inv extreme", logn, bits, in_mf)
# RNS batches (multi-modulus launches, incl. the pipelined multi-modulus kernel when HEXL_B200_PIPE=1) and the product pipeline
for logn, bit_list in ((14, (50, 55, 45)), (15, (60, 59, 61)), (16, (55, 50)), (17, (60, 60 + 0, 55)), (17, (29, 28))):
    n = 1 << logn
    mods = []
    for b in bit_list:
        for cand in hb.GeneratePrimes(4, b, True, n):
            if cand not in mods:
                mods.append(cand)
                break
    ntts = [hb.NTT(n, q) for q in mods]
    group = 3
    sz = n * group
    a = np.concatenate([uniform_below(5 * i + logn, sz, q) for i, q in enumerate(mods)])
    b = np.concatenate([uniform_below(5 * i + logn + 1, sz, q) for i, q in enumerate(mods)])
    exp_f = np.concatenate([checker.ntt_forward(a[i * sz:(i + 1) * sz], n, q) for i, q in enumerate(mods)])
    exp_i = np.concatenate([checker.ntt_inverse(a[i * sz:(i + 1) * sz], n, q) for i, q in enumerate(mods)])
    o = dev(np.zeros_like(a))
    hb.ComputeForwardMulti(ntts, o, dev(a), 1, 1, batch_per_modulus=group)
    assert (host(o) == exp_f).all(), ("multi fwd", logn, bit_list)
    hb.ComputeInverseMulti(ntts, o, dev(a), 1, 1, batch_per_modulus=group)
    assert (host(o) == exp_i).all(), ("multi inv", logn, bit_list)
    d = dev(a)
    hb.ComputeForwardMulti(ntts, d, d, 1, 4, batch_per_modulus=group)
    g = host(d)
    qs = np.concatenate([np.full(sz, q, dtype=np.uint64) for q in mods])
    assert (g % qs == exp_f).all() and (g < qs * np.uint64(4)).all(), ("multi fwd lazy in place", logn)
    if max(mods) < (1 << 61):
        conv = np.concatenate([checker.ntt_inverse(checker.mult_mod(checker.ntt_forward(a[i * sz:(i + 1) * sz], n, q),
                                                                    checker.ntt_forward(b[i * sz:(i + 1) * sz], n, q), q), n, q)
                               for i, q in enumerate(mods)])
        hb.PolyMultiplyMulti(ntts, o, dev(a), dev(b), group)
        assert (host(o) == conv).all(), ("poly multiply", logn, bit_list)
print("variant ok", {k: v for k, v in os.environ.items() if k.startswith("HEXL_B200_")})

#!/usr/bin/env python
"""Randomised parity stress: many random (N, q, batch, mod factors, aliasing, pointer kind)
combinations of every entry point against the checker.  Not part of pytest (minutes long):
    python tools/stress.py [seconds] [seed]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb  # noqa: E402
import oracle  # noqa: E402

chk = oracle.best_checker()


def prime(bits, small, n):
    """a prime = 1 mod 2n of about `bits` bits (tiny ranges may hold none: widen)"""
    while True:
        try:
            return hb.GeneratePrimes(1, bits, small, n)[0]
        except hb.HexlB200Error:
            bits += 1

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


def rand_below(n, bound):
    return (rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)) % np.uint64(bound)


counts = {}
t_end = time.time() + budget
it = 0
while time.time() < t_end:
    it += 1
    kind = rng.choice(["ntt", "ntt", "ntt", "multi", "elt", "elt", "mont", "ks", "hostcomp"])
    if kind == "mont":  # Montgomery-form helpers: random odd q < R = 2^r <= 2^62, device / host / unaligned
        r = int(rng.integers(3, 63))
        q = int(rng.integers(1 << (r - 2), 1 << r)) | 1
        if q < 3:
            q = 3
        inv = hb.HenselLemma2adicRoot(r, q)
        n = int(rng.choice([1, 2, 5, 63, 1000, 4097, 1 << 16]))
        a, b = rand_below(n, q), rand_below(n, q)
        r2 = (1 << r) * (1 << r) % q
        on_dev = bool(rng.integers(0, 2))
        A, B = (dev(a), dev(b)) if on_dev else (a.copy(), b.copy())
        O = dev(np.zeros(n, dtype=np.uint64)) if on_dev else np.zeros(n, dtype=np.uint64)
        get = host if on_dev else (lambda t: t)
        assert (get(hb.EltwiseMontReduceMod(O, A, B, n, q, r, inv)) == chk.mont_reduce_mod(a, b, q, r, inv)).all(), ("mont mul", r, q, n)
        assert (get(hb.EltwiseMontgomeryFormIn(O, A, r2, n, q, r, inv)) == chk.montgomery_form_in(a, r2, q, r, inv)).all(), ("mont in", r, q, n)
        assert (get(hb.EltwiseMontgomeryFormOut(O, O, n, q, r, inv)) == a).all(), ("mont out", r, q, n)
        counts[kind] = counts.get(kind, 0) + 1
        continue
    if kind == "ks":  # key switch with resident keys: random shape, batch, host or device buffers, sharded or not
        logn = int(rng.integers(4, 12))
        n = 1 << logn
        decomp = int(rng.integers(1, 6))
        kms = rns = decomp + 1
        kcc = int(rng.integers(1, 4))
        bits = int(rng.choice([33, 45, 50, 58, 60]))
        mods = hb.GeneratePrimes(kms, max(bits, logn + 3), True, n)
        keys = [np.concatenate([rand_below(n, mods[i]) for _ in range(kcc) for i in range(kms)]) for _ in range(decomp)]
        batch = int(rng.integers(1, 4))
        t = np.concatenate([np.concatenate([rand_below(n, mods[j]) for j in range(decomp)]) for _ in range(batch)])
        res = np.concatenate([np.concatenate([rand_below(n, mods[i]) for _ in range(kcc) for i in range(decomp)]) for _ in range(batch)])
        ms = [hb.InverseMod(mods[-1] % mods[i], mods[i]) for i in range(decomp)]
        rs, ts = kcc * decomp * n, decomp * n
        exp = np.concatenate([chk.key_switch(res[c * rs:(c + 1) * rs].copy(), t[c * ts:(c + 1) * ts], n, decomp, kms, rns, kcc,
                                             mods, keys, ms) for c in range(batch)])
        mode = int(rng.integers(0, 3))
        if mode == 2:
            hb.set_host_devices([0] * int(rng.integers(1, 5)))
            h = hb.KeySwitchKeys(keys, n, decomp, kms, kcc, sharded_by_modulus=True)
            hb.set_host_devices([])
        else:
            h = hb.KeySwitchKeys(keys, n, decomp, kms, kcc)
        if mode == 0:
            d = dev(res)
            hb.KeySwitchResident(d, dev(t), n, decomp, kms, rns, kcc, mods, h, ms, batch)
            got = host(d)
        else:
            got = res.copy()
            hb.KeySwitchResident(got, t, n, decomp, kms, rns, kcc, mods, h, ms, batch)
        assert (got == exp).all(), ("ks", logn, decomp, kcc, bits, batch, mode)
        del h
        counts[kind] = counts.get(kind, 0) + 1
        continue
    if kind == "hostcomp":  # RNS composites on host buffers (chunked staging path)
        logn = int(rng.integers(2, 14))
        n = 1 << logn
        nm = int(rng.integers(1, 5))
        group = int(rng.integers(1, 4))
        mods = [prime(int(rng.choice([29, 40, 50, 55, 60])), True, n) for _ in range(nm)]
        mods = list(dict.fromkeys(mods))
        nm = len(mods)
        ntts = [hb.NTT(n, q) for q in mods]
        sz = n * group
        a = np.concatenate([rand_below(sz, q) for q in mods])
        b = np.concatenate([rand_below(sz, q) for q in mods])
        conv = np.concatenate([chk.ntt_inverse(chk.mult_mod(chk.ntt_forward(a[i * sz:(i + 1) * sz], n, q),
                                                             chk.ntt_forward(b[i * sz:(i + 1) * sz], n, q), q), n, q)
                               for i, q in enumerate(mods)])
        o = np.zeros_like(a)
        hb.PolyMultiplyMulti(ntts, o, a, b, group)
        assert (o == conv).all(), ("host polymul", logn, mods, group)
        fwd = np.concatenate([chk.ntt_forward(a[i * sz:(i + 1) * sz], n, q) for i, q in enumerate(mods)])
        hb.ComputeForwardMulti(ntts, o, a, 1, 1, batch_per_modulus=group)
        assert (o == fwd).all(), ("host fwd multi", logn, mods, group)
        counts[kind] = counts.get(kind, 0) + 1
        continue
    if kind in ("ntt", "multi"):
        logn = int(rng.integers(1, 18))
        n = 1 << logn
        lo = logn + 2
        bits = int(rng.choice([rng.integers(lo, 62), 29, 30, 31, 32, 55, 56, 60, 61])) if logn < 27 else 40
        bits = max(bits, lo)
        q = prime(bits, bool(rng.integers(0, 2)), n)
        fwd = bool(rng.integers(0, 2))
        in_mf = int(rng.choice([1, 2, 4] if fwd else [1, 2]))
        out_mf = int(rng.choice([1, 4] if fwd else [1, 2]))
        batch = int(rng.integers(1, max(2, min(40, (1 << 18) // n))))
        if kind == "ntt":
            t = hb.NTT(n, q)
            x = rand_below(n * batch, q * in_mf)
            exp = (chk.ntt_forward if fwd else chk.ntt_inverse)(x, n, q, in_mf, 1)
            where = rng.choice(["dev", "dev_inplace", "host", "host_inplace"])
            if where.startswith("dev"):
                d = dev(x)
                o = d if where.endswith("inplace") else torch.empty_like(d)
                (t.ComputeForward if fwd else t.ComputeInverse)(o, d, in_mf, out_mf)
                got = host(o)
            else:
                src = x.copy()
                got = src if where.endswith("inplace") else np.zeros_like(x)
                (t.ComputeForward if fwd else t.ComputeInverse)(got, src, in_mf, out_mf)
            tag = ("ntt", logn, bits, fwd, in_mf, out_mf, batch, where)
        else:
            L = int(rng.integers(2, 7))
            mods = [prime(int(rng.integers(lo, 61)), True, n) for _ in range(L)]
            ntts = [hb.NTT(n, m) for m in mods]
            x = np.concatenate([rand_below(n * batch, m * in_mf) for m in mods])
            exp = np.concatenate([(chk.ntt_forward if fwd else chk.ntt_inverse)(x[i * n * batch:(i + 1) * n * batch], n, m, in_mf, 1)
                                  for i, m in enumerate(mods)])
            d = dev(x)
            (hb.ComputeForwardMulti if fwd else hb.ComputeInverseMulti)(ntts, d, d, in_mf, out_mf, batch)
            got = host(d)
            q = None
            tag = ("multi", logn, L, fwd, in_mf, out_mf, batch)
        if out_mf == 1:
            ok = (got == exp).all()
        else:
            if q is None:
                qs = np.concatenate([np.full(n * batch, m, dtype=np.uint64) for m in mods])
            else:
                qs = np.uint64(q)
            ok = (got % qs == exp).all() and (got < qs * np.uint64(out_mf)).all()
    else:
        n = int(rng.choice([1, 2, 3, 7, 64, 1000, 4097, 100003, 1 << 20]))
        bits = int(rng.integers(4, 62))
        q = int(rng.integers(1 << (bits - 1), 1 << bits)) | 1
        a, b = rand_below(n, q), rand_below(n, q)
        op = rng.choice(["add", "sub", "mult", "fma", "reduce", "cmpadd", "cmpsub"])
        on_dev = bool(rng.integers(0, 2))
        A, B = (dev(a), dev(b)) if on_dev else (a.copy(), b.copy())
        R = (torch.empty_like(A) if on_dev else np.zeros_like(a)) if rng.integers(0, 2) else A
        if op == "add":
            hb.EltwiseAddMod(R, A, B, n, q); exp = chk.add_mod(a, b, q)
        elif op == "sub":
            hb.EltwiseSubMod(R, A, B, n, q); exp = chk.sub_mod(a, b, q)
        elif op == "mult":
            mf = int(rng.choice([1, 2, 4]))
            if q * mf >= (1 << 63) or q >= (1 << 62):
                continue
            a2, b2 = rand_below(n, q * mf), rand_below(n, q * mf)
            A, B = (dev(a2), dev(b2)) if on_dev else (a2.copy(), b2.copy())
            R = torch.empty_like(A) if on_dev else np.zeros_like(a2)
            hb.EltwiseMultMod(R, A, B, n, q, mf); exp = chk.mult_mod(a2, b2, q, mf)
        elif op == "fma":
            if q >= (1 << 61):
                continue
            s = int(rng.integers(0, q))
            use_c = bool(rng.integers(0, 2))
            hb.EltwiseFMAMod(R, A, s, B if use_c else None, n, q, 1); exp = chk.fma_mod(a, s, b if use_c else None, q, 1)
        elif op == "reduce":
            if q >= (1 << 62) or q < 2:
                continue
            x = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
            A = dev(x) if on_dev else x.copy()
            R = torch.empty_like(A) if on_dev else np.zeros_like(x)
            hb.EltwiseReduceMod(R, A, n, q, q, 1); exp = x % np.uint64(q)
        elif op == "cmpadd":
            cmp_, bound, diff = int(rng.integers(0, 8)), int(rng.integers(0, q)), int(rng.integers(1, q))
            hb.EltwiseCmpAdd(R, A, n, cmp_, bound, diff); exp = chk.cmp_add(a, cmp_, bound, diff)
        else:
            cmp_, bound, diff = int(rng.integers(0, 8)), int(rng.integers(0, q)), int(rng.integers(1, q))
            hb.EltwiseCmpSubMod(R, A, n, q, cmp_, bound, diff); exp = chk.cmp_sub_mod(a, q, cmp_, bound, diff)
        got = host(R) if on_dev else R
        ok = (got == exp).all()
        tag = ("elt", op, n, bits, on_dev)
    counts[tag[0]] = counts.get(tag[0], 0) + 1
    if not ok:
        print("MISMATCH", tag, "q =", q)
        sys.exit(1)
print(f"stress ok: {it} iterations in {budget:.0f} s, {counts}, checker = {chk.kind}")

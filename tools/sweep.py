#!/usr/bin/env python
"""Measured sweep over the BASELINE configs that are not the bench headline
(configs[2], [3], [4]): eltwise kernels and NTTs across N = 2^10..2^17 and modulus
sizes, the FwdNTT -> EltwiseMultMod -> InvNTT product pipeline at N = 2^17 / 60-bit /
L = 16, and the CKKS KeySwitch composite at N = 2^15 / L = 30 -- each next to the
compiled reference (oracle/_ref) timed on the host cores.  One GPU; writes a
markdown table to stdout (kept under profiles/).

    python tools/sweep.py > gpurun_out/sweep.md
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import hexl_b200 as hb  # noqa: E402
import oracle  # noqa: E402

PEAK = 6547.8
try:
    import json
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, KeyError, ValueError):
    pass


def gpu_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


def cpu_time(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def main():
    torch.cuda.set_device(0)
    ref = oracle.best_checker()
    threads = bench.cpu_threads()
    g = torch.Generator(device="cuda").manual_seed(1)
    print(f"# Sweep on 1 x B200 (HBM copy peak {PEAK:.0f} GB/s measured); CPU = {ref.kind}, {threads} threads\n")

    # ---------------------------------------------------------------- NTT sweep
    print("## Batched NTT (device-resident, out of place), total 2^28 coefficients per buffer\n")
    print("| N | q bits | mode | fwd M NTT/s | inv M NTT/s | fwd GB/s (16N B/NTT) | frac of HBM peak | CPU fwd k NTT/s | GPU/CPU |")
    print("|---|---|---|---|---|---|---|---|---|")
    for logn in (10, 11, 12, 13, 14, 15, 16, 17):
        n = 1 << logn
        batch = (1 << 28) // n
        for bits in (29, 50, 55, 60):
            if bits <= logn + 1:
                continue
            q = hb.GeneratePrimes(1, bits, True, n)[0]
            ntt = hb.NTT(n, q)
            x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda", generator=g)
            y = torch.empty_like(x)
            tf = gpu_time(lambda: ntt.ComputeForward(y, x, 1, 1))
            ti = gpu_time(lambda: ntt.ComputeInverse(x, y, 1, 1))
            mode = "SMALL" if q < (1 << 30) else ("FAST" if (1 << 32) <= q < (1 << 56) else "GENERIC")
            cb = max(threads * 8, 64) if logn >= 14 else max(threads * 64, 1024)
            hx = np.random.default_rng(0).integers(0, q, size=n * cb, dtype=np.uint64)
            if ref.kind == "reference":
                hy = np.empty_like(hx)
                tc = cpu_time(lambda: ref.ntt_forward(hx, n, q, 1, 1, threads=threads, out=hy))
            else:
                tc = cpu_time(lambda: ref.ntt_forward(hx, n, q, 1, 1, threads=threads))
            gbs = 16.0 * n * batch / tf / 1e9
            print(f"| 2^{logn} | {bits} | {mode} | {batch / tf / 1e6:.2f} | {batch / ti / 1e6:.2f} | {gbs:.0f} | "
                  f"{gbs / PEAK:.2f} | {cb / tc / 1e3:.1f} | {batch / tf / (cb / tc):.0f}x |")
            del x, y

    # ------------------------------------------------------------ eltwise sweep
    print("\n## Eltwise kernels, batch 4096 polynomials (BASELINE configs[2])\n")
    print("| op | N | q bits | GB/s | frac of HBM peak |")
    print("|---|---|---|---|---|")
    for logn in (10, 12, 14, 16, 17):
        en = 4096 << logn
        for bits in (40, 50, 60):
            q = hb.GeneratePrimes(1, bits, True, 1 << logn)[0]
            a = torch.randint(0, q, (en,), dtype=torch.int64, device="cuda", generator=g)
            b = torch.randint(0, q, (en,), dtype=torch.int64, device="cuda", generator=g)
            r = torch.empty_like(a)
            for name, bpe, fn in (("MultMod", 24, lambda: hb.EltwiseMultMod(r, a, b, en, q, 1)),
                                  ("FMAMod", 24, lambda: hb.EltwiseFMAMod(r, a, 12345, b, en, q, 1)),
                                  ("ReduceMod", 16, lambda: hb.EltwiseReduceMod(r, a, en, q, q, 1))):
                t = gpu_time(fn, reps=10)
                print(f"| {name} | 2^{logn} | {bits} | {bpe * en / t / 1e9:.0f} | {bpe * en / t / 1e9 / PEAK:.2f} |")
            del a, b, r

    # -------------------------------------------- polynomial product pipeline (config 3)
    print("\n## FwdNTT -> EltwiseMultMod -> InvNTT, N = 2^17, 60-bit primes, L = 16 RNS moduli (BASELINE configs[3], one GPU's share shown for all 16)\n")
    n, L, pb = 1 << 17, 16, 64
    mods = hb.GeneratePrimes(L, 60, True, n)
    ntts = [hb.NTT(n, q) for q in mods]
    A = [torch.randint(0, q, (pb, n), dtype=torch.int64, device="cuda", generator=g) for q in mods]
    B = [torch.randint(0, q, (pb, n), dtype=torch.int64, device="cuda", generator=g) for q in mods]

    def polymul():
        for i, q in enumerate(mods):
            ntts[i].ComputeForward(A[i], A[i], 1, 4)
            ntts[i].ComputeForward(B[i], B[i], 1, 4)
            hb.EltwiseMultMod(A[i], A[i], B[i], pb * n, q, 4)
            ntts[i].ComputeInverse(A[i], A[i], 1, 1)

    t = gpu_time(polymul, reps=3)
    units = L * pb
    print(f"* {units} residue products (batch {pb} x {L} moduli) in {t * 1e3:.2f} ms = **{units / t / 1e3:.1f} k products/s**, "
          f"{72.0 * n * units / t / 1e9:.0f} GB/s algorithmic (72N B per product, unfused)")
    # the same work as ONE call: hexl_b200_poly_multiply_multi (multi-modulus launches, lazy transforms)
    Aall, Ball = torch.cat([x.reshape(-1) for x in A]), torch.cat([x.reshape(-1) for x in B])
    out = torch.empty_like(Aall)
    tm = gpu_time(lambda: hb.PolyMultiplyMulti(ntts, out, Aall, Ball, pb), reps=3)
    print(f"* as one `hexl_b200_poly_multiply_multi` call (6 launches): {tm * 1e3:.2f} ms = **{units / tm / 1e3:.1f} k products/s**")
    del Aall, Ball, out
    for pb_small in (1, 2):
        As = torch.cat([x[:pb_small].reshape(-1) for x in A]); Bs = torch.cat([x[:pb_small].reshape(-1) for x in B])
        outs = torch.empty_like(As)
        t1 = gpu_time(lambda: hb.PolyMultiplyMulti(ntts, outs, As, Bs, pb_small), reps=20)

        def per_call():
            for i, q in enumerate(mods):
                sl = slice(i * pb_small * n, (i + 1) * pb_small * n)
                ntts[i].ComputeForward(As[sl], As[sl], 1, 4)
                ntts[i].ComputeForward(Bs[sl], Bs[sl], 1, 4)
                hb.EltwiseMultMod(As[sl], As[sl], Bs[sl], pb_small * n, q, 4)
                ntts[i].ComputeInverse(As[sl], As[sl], 1, 1)
        t2 = gpu_time(per_call, reps=20)
        print(f"* {pb_small} polynomial(s) x {L} moduli (one ciphertext component): one call {t1 * 1e6:.0f} us vs {4 * L} per-modulus calls {t2 * 1e6:.0f} us ({t2 / t1:.1f}x)")
    hxa = np.random.default_rng(1).integers(0, mods[0], size=n * 32, dtype=np.uint64)
    hxb = np.random.default_rng(2).integers(0, mods[0], size=n * 32, dtype=np.uint64)
    if ref.kind == "reference":
        def cpu_polymul():
            fa = ref.ntt_forward(hxa, n, mods[0], 1, 4, threads=threads)
            fb = ref.ntt_forward(hxb, n, mods[0], 1, 4, threads=threads)
            pr = ref.mult_mod(fa, fb, mods[0], 4, rows=32, threads=threads)
            ref.ntt_inverse(pr, n, mods[0], 1, 1, threads=threads)
        tc = cpu_time(cpu_polymul)
        print(f"* CPU reference ({threads} threads): {32 / tc / 1e3:.2f} k products/s -> GPU/CPU = {units / t / (32 / tc):.0f}x")
    del A, B

    # ------------------------------------------------- multi-modulus launches vs per-modulus calls
    print("\n## RNS batch: one ciphertext-sized unit per modulus (2 polynomials x L moduli), forward NTT, N = 2^15\n")
    print("| L moduli | one multi-modulus launch, us | L per-modulus calls, us | ratio |")
    print("|---|---|---|---|")
    n = 1 << 15
    for L in (4, 16, 30, 60):
        mods = hb.GeneratePrimes(L, 50, True, n)
        ntts = [hb.NTT(n, q) for q in mods]
        x = torch.cat([torch.randint(0, q, (2 * n,), dtype=torch.int64, device="cuda", generator=g) for q in mods])
        y = torch.empty_like(x)
        tm = gpu_time(lambda: hb.ComputeForwardMulti(ntts, y, x, 1, 1, 2), reps=20)

        def per_modulus():
            for i, t in enumerate(ntts):
                t.ComputeForward(y[i * 2 * n:(i + 1) * 2 * n], x[i * 2 * n:(i + 1) * 2 * n], 1, 1)
        tp = gpu_time(per_modulus, reps=20)
        print(f"| {L} | {tm * 1e6:.1f} | {tp * 1e6:.1f} | {tp / tm:.1f}x |")

    # ------------------------------------------------------------------ DyadicMultiply
    print("\n## DyadicMultiply (ciphertext tensor product), N = 2^15, batch of 64 ciphertext pairs as one call\n")
    print("| L moduli | GB/s (56 B per coefficient slot) | frac of HBM peak |")
    print("|---|---|---|")
    for L in (8, 30):
        big_n = n * 64  # DyadicMultiply is position-independent inside a modulus: a batch is a longer polynomial
        mods = hb.GeneratePrimes(L, 50, True, n)
        a = torch.cat([torch.randint(0, q, (big_n,), dtype=torch.int64, device="cuda", generator=g) for _ in range(2) for q in mods])
        b = torch.cat([torch.randint(0, q, (big_n,), dtype=torch.int64, device="cuda", generator=g) for _ in range(2) for q in mods])
        out = torch.empty(3 * big_n * L, dtype=torch.int64, device="cuda")
        t = gpu_time(lambda: hb.DyadicMultiply(out, a, b, big_n, mods), reps=10)
        print(f"| {L} | {56.0 * big_n * L / t / 1e9:.0f} | {56.0 * big_n * L / t / 1e9 / PEAK:.2f} |")
        del a, b, out

    # ------------------------------------------------------- KeySwitch (config 4 shape)
    print("\n## CKKS KeySwitch composite, N = 2^15 (BASELINE configs[4] shape)\n")
    print("| decomp moduli | GPU ms / key switch | CPU reference ms | GPU/CPU |")
    print("|---|---|---|---|")
    n = 1 << 15
    for decomp in (6, 14, 29):
        kms = rns = decomp + 1
        kcc = 2
        mods = hb.GeneratePrimes(kms, 50, True, n)
        tt = torch.cat([torch.randint(0, mods[j], (n,), dtype=torch.int64, device="cuda", generator=g) for j in range(decomp)])
        keys = [torch.cat([torch.randint(0, mods[i], (n,), dtype=torch.int64, device="cuda", generator=g)
                           for _ in range(kcc) for i in range(kms)]) for _ in range(decomp)]
        res = torch.cat([torch.randint(0, mods[i], (n,), dtype=torch.int64, device="cuda", generator=g)
                         for _ in range(kcc) for i in range(decomp)])
        ms = [hb.InverseMod(mods[-1] % mods[i], mods[i]) for i in range(decomp)]
        tg = gpu_time(lambda: hb.KeySwitch(res, tt, n, decomp, kms, rns, kcc, mods, keys, ms), reps=3, warm=1)
        if ref.kind == "reference" and getattr(ref, "has_seal", False):
            h_res = res.cpu().numpy().view(np.uint64).copy()
            h_tt = tt.cpu().numpy().view(np.uint64)
            h_keys = [k.cpu().numpy().view(np.uint64) for k in keys]
            tc = cpu_time(lambda: ref.key_switch(h_res, h_tt, n, decomp, kms, rns, kcc, mods, h_keys, ms), reps=1)
            print(f"| {decomp} | {tg * 1e3:.2f} | {tc * 1e3:.1f} | {tc / tg:.0f}x |")
        else:
            print(f"| {decomp} | {tg * 1e3:.2f} | n/a | n/a |")
        del tt, keys, res


if __name__ == "__main__":
    main()

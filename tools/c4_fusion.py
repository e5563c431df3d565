#!/usr/bin/env python
"""C4 product pipeline (N = 2^17 x 16 x 60-bit moduli and N = 2^16 x 8 x 55-bit) with the point-wise product folded into
the inverse transform (default) and as the unfused chain (HEXL_B200_NO_PRODUCT_FUSION=1, read once per process, hence
one subprocess per arm).  Prints ms per call and residue products/s; CUDA-event timed after warm-up."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def leg():
    import torch
    import hexl_b200 as hb
    gen = torch.Generator(device="cuda").manual_seed(7)
    for logn, nmod, bits, group in ((17, 16, 60, 32), (16, 8, 55, 64), (13, 8, 50, 256)):
        n = 1 << logn
        mods = hb.GeneratePrimes(nmod, bits, True, n)
        ntts = [hb.NTT(n, q) for q in mods]
        sz = n * group
        a = torch.empty(nmod * sz, dtype=torch.int64, device="cuda")
        b = torch.empty_like(a)
        for i, q in enumerate(mods):
            a[i * sz:(i + 1) * sz].random_(0, q, generator=gen)
            b[i * sz:(i + 1) * sz].random_(0, q, generator=gen)
        r = torch.empty_like(a)
        for _ in range(3):
            hb.PolyMultiplyMulti(ntts, r, a, b, group)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            hb.PolyMultiplyMulti(ntts, r, a, b, group)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  N=2^{logn} {nmod} x {bits}-bit x {group}: {ms:.3f} ms per call, {nmod * group / ms * 1e3:,.0f} residue products/s,"
              f" checksum {int(r.sum().item()) & 0xffffffff:08x}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "leg":
        leg()
    else:
        for name, val in (("multiplied on load (default)", "0"), ("unfused chain", "1")):
            print(name, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "leg"],
                           env={**os.environ, "HEXL_B200_NO_PRODUCT_FUSION": val}, check=True)

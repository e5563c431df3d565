#!/usr/bin/env python
"""Exact-rational model of the FP64-assisted quotient estimate of hexl_b200/csrc/ntt_kernels.cuh (TwH /
mul_tw_h, compiled only with -DHEXL_B200_FP64Q): two round-toward-minus-infinity binary64 fused multiply-adds
on {2^52 + word} operands reproduce the cross terms of floor(x * w' / 2^64) to within [-2, 0], so that the
quotient is low by 0, 1 or 2 exactly like the three-product integer estimate it replaces, and the constant
that rides along in the mantissa bits cancels through Mod::bias.  Runs on the CPU, no GPU needed:
    python tools/fp64_quot_model.py
"""
import random
import struct
from fractions import Fraction

KBIAS = 0x4330000000000002
M64 = (1 << 64) - 1


def round_down(x: Fraction) -> Fraction:
    """x > 0 rounded toward -inf to binary64 (normal range)"""
    e = x.numerator.bit_length() - x.denominator.bit_length()
    if Fraction(2) ** e > x:
        e -= 1
    if Fraction(2) ** (e + 1) <= x:
        e += 1
    ulp = Fraction(2) ** (e - 52)
    return (x // ulp) * ulp


def bits(x: Fraction) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def check(q: int, trials: int, rng: random.Random) -> int:
    worst = 0
    n = (1 << 64) - q
    bias = (KBIAS * q) & M64
    for _ in range(trials):
        w = rng.randrange(q)
        wp = (w << 64) // q
        kind = rng.random()
        x = rng.choice([0, 1, (1 << 32) - 1, 1 << 32, (1 << 63) - 1, (1 << 63) - (1 << 32)]) if kind < 0.1 \
            else rng.randrange(1 << rng.choice([10, 33, 58, 63]))
        if kind > 0.9:
            wp = rng.choice([0, M64, (1 << 32) - 1, M64 - (1 << 32) + 1, ((1 << 32) - 1) << 32, 1])
        x0, x1, b0, b1 = x & 0xFFFFFFFF, x >> 32, wp & 0xFFFFFFFF, wp >> 32
        K = Fraction((1 << 52) + 2 - ((b0 + b1) << 20))
        assert float(K) == K
        u = round_down(Fraction((1 << 52) + x0) * Fraction(b1, 1 << 32) + K)
        R = round_down(Fraction((1 << 52) + x1) * Fraction(b0, 1 << 32) + u)
        assert (1 << 52) <= R < (1 << 53)
        Qc = (x1 * b1 + bits(R)) & M64
        Q = (Qc - KBIAS) & M64
        d = ((x * wp) >> 64) - Q
        assert 0 <= d <= 2, (d, x, wp)
        worst = max(worst, d)
        if kind <= 0.9:
            T = (x * w + Qc * n + bias) & M64
            assert T == x * w - Q * q and T < 4 * q and T % q == (x * w) % q
    return worst


if __name__ == "__main__":
    rng = random.Random(1)
    for q in ((1 << 55) + 1234567, (1 << 56) - 5, (1 << 32) + 15, (1 << 61) - 1, (1 << 40) + 123):
        print(f"q ~ 2^{q.bit_length() - 1}: quotient low by at most {check(q, 20000, rng)} (allowed 2)")

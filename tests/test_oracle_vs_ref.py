"""The oracle restatement against the COMPILED reference (oracle/_ref) on random
inputs -- the reference never tests N > 2^13 itself, so large-N
parity rests on this.  Lazy outputs are compared bit for bit with the
reference's scalar tier and modulo q with its AVX-512 tiers (which the
reference's own tests do too: test/test-ntt-avx512.cpp:194-204).  CPU only."""
import numpy as np
import pytest

from util import uniform_below

NTT_CASES = [(2, 48), (4, 20), (8, 22), (16, 29), (64, 31), (1024, 30), (2048, 49), (4096, 50),
             (8192, 60), (16384, 58), (32768, 50), (65536, 55), (131072, 60)]


@pytest.mark.parametrize("n,bits", NTT_CASES)
def test_ntt_port_matches_reference(port, ref, n, bits):
    q = ref.generate_primes(1, bits, True, n)[0]
    assert port.generate_primes(1, bits, True, n)[0] == q
    assert port.minimal_primitive_root(2 * n, q) == ref.root(n, q)
    for a, b in zip(port.tables(n, q)[1:], ref.tables(n, q)):
        assert (a == b).all()
    batch = 3 if n <= 8192 else 1
    qq = np.uint64(q)
    for in_mf, out_mf in [(1, 1), (2, 1), (4, 1), (1, 4), (2, 4), (4, 4)]:
        x = uniform_below(n + in_mf, n * batch, q * in_mf)
        a = port.ntt_forward(x, n, q, in_mf, out_mf)
        assert (a == ref.ntt_forward(x, n, q, in_mf, out_mf, native=True)).all()
        d = ref.ntt_forward(x, n, q, in_mf, out_mf)
        assert (a % qq == d % qq).all() and (d < np.uint64(out_mf * q)).all()
        if out_mf == 1:
            assert (a == d).all()
    for in_mf, out_mf in [(1, 1), (2, 1), (1, 2), (2, 2)]:
        x = uniform_below(7 * n + in_mf, n * batch, q * in_mf)
        a = port.ntt_inverse(x, n, q, in_mf, out_mf)
        assert (a == ref.ntt_inverse(x, n, q, in_mf, out_mf, native=True)).all()
        d = ref.ntt_inverse(x, n, q, in_mf, out_mf)
        assert (a % qq == d % qq).all() and (d < np.uint64(out_mf * q)).all()
        if out_mf == 1:
            assert (a == d).all()
    x = uniform_below(99, n, q)
    assert (port.ntt_inverse(port.ntt_forward(x, n, q), n, q) == x).all()
    if n <= 4096:
        assert (port.ntt_forward_textbook(x, n, q) == ref.ntt_forward_textbook(x, n, q)).all()
        assert (ref.ntt_forward_radix4(x, n, q) == port.ntt_forward(x, n, q)).all()


@pytest.mark.parametrize("bits", [20, 30, 31, 32, 33, 40, 48, 50, 51, 52, 55, 58, 59, 60])
def test_eltwise_port_matches_reference(port, ref, bits):
    n = 1024 + 7  # the reference's own odd size (test-eltwise-reduce-mod.cpp:103)
    q = ref.generate_primes(1, bits, True, 1)[0]
    qq = np.uint64(q)
    a, b = uniform_below(1, n, q), uniform_below(2, n, q)
    for native in (True, False):
        assert (port.add_mod(a, b, q) == ref.add_mod(a, b, q, native=native)).all()
        assert (port.add_mod(a, int(b[0]), q) == ref.add_mod(a, int(b[0]), q, native=native)).all()
        assert (port.sub_mod(a, b, q) == ref.sub_mod(a, b, q, native=native)).all()
        assert (port.sub_mod(a, int(b[0]), q) == ref.sub_mod(a, int(b[0]), q, native=native)).all()
        for mf in (1, 2, 4):
            x, y = uniform_below(3, n, mf * q), uniform_below(4, n, mf * q)
            assert (port.mult_mod(x, y, q, mf) == ref.mult_mod(x, y, q, mf, native=native)).all()
        for mf in (1, 2, 4, 8):
            x, c = uniform_below(5, n, mf * q), uniform_below(6, n, mf * q)
            s = int(uniform_below(7, 1, mf * q)[0])
            assert (port.fma_mod(x, s, c, q, mf) == ref.fma_mod(x, s, c, q, mf, native=native)).all()
            assert (port.fma_mod(x, s, None, q, mf) == ref.fma_mod(x, s, None, q, mf, native=native)).all()
        wide = uniform_below(8, n, 1 << 63)
        assert (port.reduce_mod(wide, q, q, 1) == ref.reduce_mod(wide, q, q, 1, native=native)).all()
        lazy = ref.reduce_mod(wide, q, q, 2, native=native)
        assert (port.reduce_mod(wide, q, q, 2) % qq == lazy % qq).all() and (lazy < np.uint64(2 * q)).all()
        x4 = uniform_below(9, n, 4 * q)
        assert (port.reduce_mod(x4, q, 4, 1) == ref.reduce_mod(x4, q, 4, 1, native=native)).all()
        assert (port.reduce_mod(x4, q, 4, 2) == ref.reduce_mod(x4, q, 4, 2, native=native)).all()
        x2 = uniform_below(10, n, 2 * q)
        assert (port.reduce_mod(x2, q, 2, 1) == ref.reduce_mod(x2, q, 2, 1, native=native)).all()
        for cmp in range(8):
            bound, diff = int(a[5]), int(b[6]) or 1
            assert (port.cmp_add(a, cmp, bound, diff) == ref.cmp_add(a, cmp, bound, diff, native=native)).all()
            w = uniform_below(11, n, 1 << 64)
            bound = int(w[3])
            assert (port.cmp_sub_mod(w, q, cmp, bound, diff)
                    == ref.cmp_sub_mod(w, q, cmp, bound, diff, native=native)).all()


@pytest.mark.parametrize("logn,decomp,bits", [(4, 2, 59), (10, 3, 50), (13, 5, 58)])
def test_key_switch_port_matches_reference(port, ref, logn, decomp, bits):
    if not ref.has_seal:
        pytest.skip("oracle/_ref was built without the experimental/seal sources")
    n = 1 << logn
    kms = rns = decomp + 1
    kcc = 2
    mods = ref.generate_primes(kms, bits, True, n)
    t_target = np.concatenate([uniform_below(30 + j, n, mods[j]) for j in range(decomp)])
    keys = [np.concatenate([uniform_below(100 * j + 7 * k + i, n, mods[i]) for k in range(kcc) for i in range(kms)])
            for j in range(decomp)]
    result = np.concatenate([uniform_below(500 + 10 * k + i, n, mods[i]) for k in range(kcc) for i in range(decomp)])
    modswitch = [ref.inverse_mod(mods[-1] % mods[i], mods[i]) for i in range(decomp)]
    a = port.key_switch(result.copy(), t_target, n, decomp, kms, rns, kcc, mods, keys, modswitch)
    b = ref.key_switch(result.copy(), t_target, n, decomp, kms, rns, kcc, mods, keys, modswitch)
    assert (a == b).all()
    x = np.concatenate([uniform_below(1 + i, n, q) for _ in range(2) for i, q in enumerate(mods)])
    y = np.concatenate([uniform_below(9 + i, n, q) for _ in range(2) for i, q in enumerate(mods)])
    assert (port.dyadic_multiply(x, y, n, mods) == ref.dyadic_multiply(x, y, n, mods)).all()

"""CPU-side tests of the C-ABI library: it loads, exports every symbol the header
declares, its host number theory / table construction agree with the
reference's known answers and with the oracle, argument validation mirrors the
reference's HEXL_CHECKs, and -- with no GPU -- compute calls fail loudly instead
of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import kat_values

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "hexl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hexl_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(hb):
    lib = ctypes.CDLL(hb.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/hexl_b200.h but not exported"


def test_cpp_drop_in_headers_compile(hb, tmp_path):
    """A reference-style caller (the shape of example/example.cpp) compiles and
    links against include/hexl/hexl.hpp + libhexl_b200.so unchanged."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "hexl", "hexl.hpp")
    if not os.path.exists(hdr) or not shutil.which("g++"):
        pytest.skip("C++ headers or g++ not present")
    src = os.path.join(ROOT, "tests", "cpp", "example_caller.cpp")
    exe = tmp_path / "example_caller"
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", str(exe),
           "-L", os.path.dirname(hb.LIB_PATH), "-lhexl_b200", f"-Wl,-rpath,{os.path.dirname(hb.LIB_PATH)}"]
    subprocess.run(cmd, check=True)


def test_number_theory_kats(hb, kats):
    for degree, q, root in kats["minimal_primitive_root"]["cases"]:
        assert hb.MinimalPrimitiveRoot(degree, q) == root
        assert hb.IsPrimitiveRoot(hb.GeneratePrimitiveRoot(degree, q), degree, q)
    for root, degree, q, exp in kats["is_primitive_root"]["cases"]:
        assert hb.IsPrimitiveRoot(root, degree, q) == exp
    for x, y, q, exp in kats["multiply_mod"]["cases"]:
        assert hb.MultiplyMod(x, y, q) == exp
    for b, e, q, exp in kats["pow_mod"]["cases"]:
        assert hb.PowMod(b, e, q) == exp
    for x, q, exp in kats["inverse_mod"]["cases"]:
        assert hb.InverseMod(x, q) == exp
    for x, w, exp in kats["reverse_bits"]["cases"]:
        assert hb.ReverseBits(x, w) == exp
    for p in kats["is_prime"]["prime"]:
        assert hb.IsPrime(p)
    for p in kats["is_prime"]["composite"]:
        assert not hb.IsPrime(p)
    assert hb.AddUIntMod(7, 8, 10) == 5 and hb.SubUIntMod(3, 8, 10) == 5


def test_generate_primes_kat(hb, port, kats):
    g = kats["generate_primes"]
    for bits in range(g["bits"][0], g["bits"][1] + 1):
        for small in (True, False):
            ps = hb.GeneratePrimes(g["count"], bits, small, g["ntt_size"])
            assert ps == port.generate_primes(g["count"], bits, small, g["ntt_size"])
            for p in ps:
                assert p % (2 * g["ntt_size"]) == 1 and hb.IsPrime(p)


def test_tables_match_kat_and_oracle(hb, port, kats):
    for c in kats["ntt_powers"]["cases"]:
        assert [int(v) for v in hb.NTT(c["n"], c["q"]).GetRootOfUnityPowers()] == c["powers"]
    for n, bits in [(2, 48), (16, 29), (1024, 30), (4096, 50), (65536, 55), (131072, 60)]:
        q = hb.GeneratePrimes(1, bits, True, n)[0]
        t = hb.NTT(n, q)
        root, w, wp, iw, iwp = port.tables(n, q)
        assert t.GetMinimalRootOfUnity() == root and t.GetDegree() == n and t.GetModulus() == q
        assert (t.GetRootOfUnityPowers() == w).all()
        assert (t.GetPrecon64RootOfUnityPowers() == wp).all()
        assert (t.GetInvRootOfUnityPowers() == iw).all()
        assert (t.GetPrecon64InvRootOfUnityPowers() == iwp).all()
    # user-supplied root (ntt.hpp:75): any primitive 2N-th root is accepted
    q = hb.GeneratePrimes(1, 30, True, 64)[0]
    r = hb.MinimalPrimitiveRoot(128, q)
    other = hb.PowMod(r, 3, q)
    t = hb.NTT(64, q, other)
    assert t.GetMinimalRootOfUnity() == other
    assert (t.GetRootOfUnityPowers() == port.tables(64, q, other)[1]).all()


def test_argument_validation_mirrors_hexl_checks(hb):
    E = hb.HexlB200Error
    # NTT::CheckArguments, hexl/ntt/ntt-internal.cpp:171-186; test/test-ntt.cpp:21-94
    assert hb.NTT.CheckArguments(8, 769)
    for degree, q in [(7, 769), (8, 770), (8, 771), (1 << 21, 769), (8, (1 << 62) + 81)]:
        assert not hb.NTT.CheckArguments(degree, q)
        with pytest.raises(E):
            hb.NTT(degree, q)
    with pytest.raises(E):
        hb.NTT(8, 769, 2)  # not a primitive 16th root
    t = hb.NTT(8, 769)
    x = np.arange(8, dtype=np.uint64)
    for args in [(3, 1), (1, 2), (8, 1)]:  # bad forward mod factors (ntt-internal.cpp:193-197)
        with pytest.raises(E) as ei:
            t.ComputeForward(x, x, *args)
        assert ei.value.code == -1
    for args in [(4, 1), (1, 4)]:  # bad inverse mod factors (:257-260)
        with pytest.raises(E) as ei:
            t.ComputeInverse(x, x, *args)
        assert ei.value.code == -1
    # eltwise checks: n == 0, modulus bounds, mod factors, scalar bounds
    y = np.zeros(8, dtype=np.uint64)
    bad = [
        lambda: hb.EltwiseAddMod(y, x, x, 0, 769),
        lambda: hb.EltwiseAddMod(y, x, x, 8, 1),
        lambda: hb.EltwiseAddMod(y, x, x, 8, 1 << 63),
        lambda: hb.EltwiseAddMod(y, x, 769, 8, 769),
        lambda: hb.EltwiseSubMod(y, x, 800, 8, 769),
        lambda: hb.EltwiseMultMod(y, x, x, 8, 769, 3),
        lambda: hb.EltwiseMultMod(y, x, x, 8, 1 << 62, 1),
        lambda: hb.EltwiseFMAMod(y, x, 1, x, 8, 1 << 61, 1),
        lambda: hb.EltwiseFMAMod(y, x, 1, x, 8, 769, 3),
        lambda: hb.EltwiseFMAMod(y, x, 769, x, 8, 769, 1),
        lambda: hb.EltwiseReduceMod(y, x, 8, 769, 3, 1),
        lambda: hb.EltwiseReduceMod(y, x, 8, 769, 2, 4),
        lambda: hb.EltwiseCmpAdd(y, x, 8, hb.CMPINT.EQ, 1, 0),
        lambda: hb.EltwiseCmpSubMod(y, x, 8, 769, hb.CMPINT.EQ, 1, 0),
        lambda: hb.EltwiseCmpSubMod(y, x, 8, 769, hb.CMPINT.EQ, 1, 769),
    ]
    for f in bad:
        with pytest.raises(E) as ei:
            f()
        assert ei.value.code == -1


def test_argument_validation_of_the_batched_and_composite_entry_points(hb):
    """cheap checks of the RNS-batched calls and the SEAL composites fire before any device work"""
    E = hb.HexlB200Error
    a8, b8, c16 = hb.NTT(8, 769), hb.NTT(8, 17), hb.NTT(16, 769)
    x = np.arange(16, dtype=np.uint64)
    y = np.zeros(16, dtype=np.uint64)
    big = hb.NTT(8, hb.GeneratePrimes(1, 61, True, 8)[0])  # >= 2^61: lazy products would overflow
    bad = [
        lambda: hb.ComputeForwardMulti([a8, c16], y, x, 1, 1, 1),          # mixed degrees
        lambda: hb.ComputeForwardMulti([a8, b8], y, x, 3, 1, 1),           # bad input factor
        lambda: hb.ComputeInverseMulti([a8, b8], y, x, 1, 4, 1),           # bad output factor
        lambda: hb.EltwiseMultModMulti(y, x, x, 8, [769, 17], 3),          # bad input factor
        lambda: hb.EltwiseMultModMulti(y, x, x, 8, [769, 1], 1),           # modulus <= 1
        lambda: hb.EltwiseAddModMulti(y, x, x, 0, [769, 17]),              # n == 0
        lambda: hb.PolyMultiplyMulti([a8, c16], y, x, x, 1),               # mixed degrees
        lambda: hb.PolyMultiplyMulti([a8, big], y, x, x, 1),               # modulus too large for the lazy pipeline
        lambda: hb.DyadicMultiply(y, x, x, 0, [769]),                      # n == 0
        lambda: hb.DyadicMultiply(np.zeros(24, dtype=np.uint64), x, x, 8, [1 << 62]),
        lambda: hb.KeySwitch(y, x, 8, 1, 1, 2, 2, [769, 17], [x], [1]),    # key_modulus_size < rns_modulus_size
        lambda: hb.KeySwitch(y, x, 6, 1, 2, 2, 2, [769, 17], [x], [1]),    # n not a power of two
    ]
    for i, f in enumerate(bad):
        with pytest.raises(E) as ei:
            f()
        assert ei.value.code == -1, i


def test_no_cpu_fallback_without_gpu(hb):
    """The product path must fail loudly when no CUDA device is usable."""
    if hb.device_count() > 0:
        pytest.skip("a GPU is present")
    t = hb.NTT(8, 769)
    x = np.arange(8, dtype=np.uint64)
    y = np.zeros_like(x)
    for f in (lambda: t.ComputeForward(y, x, 1, 1), lambda: t.ComputeInverse(y, x, 1, 1),
              lambda: hb.EltwiseAddMod(y, x, x, 8, 769), lambda: hb.EltwiseMultMod(y, x, x, 8, 769, 1),
              lambda: hb.EltwiseReduceMod(y, x, 8, 769, 769, 1),
              lambda: hb.ComputeForwardMulti([t, hb.NTT(8, 17)], np.zeros(16, dtype=np.uint64), np.arange(16, dtype=np.uint64) % 17),
              lambda: hb.PolyMultiplyMulti([t], y, x, x, 1),
              lambda: hb.EltwiseMultModMulti(y, x, x, 8, [769]),
              lambda: hb.DyadicMultiply(np.zeros(12, dtype=np.uint64), np.arange(8, dtype=np.uint64) % 5,
                                        np.arange(8, dtype=np.uint64) % 5, 4, [769])):
        with pytest.raises(hb.HexlB200Error) as ei:
            f()
        assert ei.value.code == -2  # HEXL_B200_ERR_NO_DEVICE
    assert (y == 0).all()
    assert hb.launch_count() == 0


def test_product_never_touches_the_oracle():
    """Nothing under hexl_b200/ or include/ may reference oracle/."""
    for top in ("hexl_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            if "_obj" in dp or "__pycache__" in dp:
                continue
            for f in files:
                if f.endswith((".so", ".o", ".log", ".pyc")):
                    continue
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dp, f)

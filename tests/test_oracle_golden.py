"""The oracle (oracle/hexl_oracle.c) against the reference's own known-answer
vectors (tests/golden/reference_kats.json, restated from the reference's tests).
This is what pins the oracle; it runs on CPU."""
import numpy as np

from util import kat_modulus, kat_values


def test_ntt_forward_kats(port, kats):
    # test/test-ntt.cpp:231-355 pushes each tuple through these same variants
    for c in kats["ntt_forward"]["cases"]:
        n, q = c["n"], c["q"]
        x = np.array(c["input"], dtype=np.uint64)
        exp = np.array(c["output"], dtype=np.uint64)
        assert (port.ntt_forward(x, n, q, 1, 1) == exp).all(), c
        assert (port.ntt_forward(x, n, q, 2, 4) % np.uint64(q) == exp).all(), c
        assert (port.ntt_forward_textbook(x, n, q) == exp).all(), c
        assert (port.ntt_inverse_textbook(exp, n, q) == x).all(), c
        assert (port.ntt_inverse(exp, n, q, 1, 1) == x).all(), c
        assert (port.ntt_inverse(exp, n, q, 1, 2) % np.uint64(q) == x).all(), c


def test_ntt_powers_kat(port, kats):
    for c in kats["ntt_powers"]["cases"]:
        _, w, _, _, _ = port.tables(c["n"], c["q"])
        assert [int(v) for v in w] == c["powers"]


def test_ntt_roundtrip_kat(port, kats):
    for c in kats["ntt_roundtrip"]["cases"]:
        x = np.array(c["input"], dtype=np.uint64)
        y = port.ntt_forward(x, c["n"], c["q"])
        assert (port.ntt_inverse(y, c["n"], c["q"]) == x).all()


def test_number_theory_kats(port, kats):
    for degree, q, root in kats["minimal_primitive_root"]["cases"]:
        assert port.minimal_primitive_root(degree, q) == root
    for root, degree, q, exp in kats["is_primitive_root"]["cases"]:
        assert bool(port.L.orc_is_primitive_root(root, degree, q)) == exp
    for x, y, q, exp in kats["multiply_mod"]["cases"]:
        assert port.multiply_mod(x, y, q) == exp
    for b, e, q, exp in kats["pow_mod"]["cases"]:
        assert port.pow_mod(b, e, q) == exp
    for x, q, exp in kats["inverse_mod"]["cases"]:
        assert port.inverse_mod(x, q) == exp
    for x, w, exp in kats["reverse_bits"]["cases"]:
        assert port.reverse_bits(x, w) == exp
    for p in kats["is_prime"]["prime"]:
        assert port.is_prime(p)
    for p in kats["is_prime"]["composite"]:
        assert not port.is_prime(p)


def test_generate_primes_kat(port, kats):
    g = kats["generate_primes"]
    for bits in range(g["bits"][0], g["bits"][1] + 1):
        for small in (True, False):
            ps = port.generate_primes(g["count"], bits, small, g["ntt_size"])
            assert len(ps) == g["count"]
            for p in ps:
                assert p % (2 * g["ntt_size"]) == 1 and port.is_prime(p)
                assert (1 << bits) <= p <= (1 << (bits + 1))


def test_eltwise_kats(port, kats):
    gp = port.generate_primes
    for c in kats["eltwise_mult_mod"]["cases"]:
        q = kat_modulus(c["q"], gp)
        out = port.mult_mod(kat_values(c["op1"], q), kat_values(c["op2"], q), q, c["in_mf"])
        assert (out == kat_values(c["out"], q)).all(), c
    for c in kats["eltwise_fma_mod"]["cases"]:
        q = c["q"]
        a3 = None if c["arg3"] is None else kat_values(c["arg3"], q)
        out = port.fma_mod(kat_values(c["arg1"], q), c["arg2"], a3, q, c["in_mf"])
        assert (out == kat_values(c["out"], q)).all(), c
    s = kats["eltwise_fma_mod"]["in_mf_sweep"]
    for mf in s["in_mfs"]:
        q = s["q"]
        a1 = kat_values(s["arg1_base"], q) + np.uint64((mf - 1) * q)
        out = port.fma_mod(a1, s["arg2"], kat_values(s["arg3"], q), q, mf)
        assert (out == kat_values(s["out"], q)).all(), mf
    for c in kats["eltwise_reduce_mod"]["cases"]:
        q = c["q"]
        in_mf = q if c["in_mf"] == "q" else c["in_mf"]
        out = port.reduce_mod(kat_values(c["op"], q), q, in_mf, c["out_mf"])
        assert (out == kat_values(c["out"], q)).all(), c
    for name, fn in (("eltwise_add_mod", port.add_mod), ("eltwise_sub_mod", port.sub_mod)):
        for c in kats[name]["cases"]:
            q = kat_modulus(c["q"], gp)
            out = fn(kat_values(c["op1"], q), kat_values(c["op2"], q), q)
            assert (out == kat_values(c["out"], q)).all(), c
    for c in kats["eltwise_cmp_add"]["cases"]:
        out = port.cmp_add(kat_values(c["op1"], 0), c["cmp"], c["bound"], c["diff"])
        assert (out == kat_values(c["out"], 0)).all()
    for c in kats["eltwise_cmp_sub_mod"]["cases"]:
        out = port.cmp_sub_mod(kat_values(c["op1"], 0), c["q"], c["cmp"], c["bound"], c["diff"])
        assert (out == kat_values(c["out"], 0)).all()


def _seal_kats():
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return json.load(open(os.path.join(root, "tests", "golden", "seal_kats.json")))


def test_seal_composite_kats(port):
    """KeySwitch (test/experimental/seal/test-key-switch.cpp:16-186) and DyadicMultiply
    (test-dyadic-multiply.cpp:16-155) known answers pin the oracle's composites."""
    d = _seal_kats()
    k = d["key_switch"]
    res = np.array(k["input"], dtype=np.uint64)
    port.key_switch(res, k["t_target_iter_ptr"], k["coeff_count"], k["decomp_modulus_size"], k["key_modulus_size"],
                    k["rns_modulus_size"], k["key_component_count"], k["moduli"], k["k_switch_keys"],
                    k["modswitch_factors"])
    assert (res == np.array(k["expected_output"], dtype=np.uint64)).all()
    for c in d["dyadic_multiply"]["cases"]:
        n, nm = c["coeff_count"], len(c["moduli"])
        op2 = c["op2"] if c["op2"] is not None else c["op1"]
        out = port.dyadic_multiply(c["op1"][:2 * n * nm], op2[:2 * n * nm], n, c["moduli"])
        assert (out == np.array(c["exp_out"], dtype=np.uint64)).all(), c["name"]

"""TEST INFRASTRUCTURE ONLY -- loaders for the CPU checkers.

* ``Port``  : oracle/hexl_oracle.c, the plain-C restatement (oracle/_build/liboracle.so)
* ``Ref``   : the unmodified reference compiled from /root/reference by
              oracle/Makefile (oracle/_ref/libhexl_ref.so, or the scalar-only
              libhexl_ref_scalar.so when the host lacks AVX-512)

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  Nothing under hexl_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhexl_ref.so")
REF_SCALAR_SO = os.path.join(HERE, "_ref", "libhexl_ref_scalar.so")
REFERENCE_TREE = os.environ.get("HEXL_REFERENCE_TREE", "/root/reference")

u64 = C.c_uint64
vp = C.c_void_p


def build(port: bool = True, ref: bool = True, quiet: bool = True) -> None:
    """Compile the checkers.  The reference build needs /root/reference (this
    container only); on the GPU box the prebuilt oracle/_ref/*.so is used."""
    targets = []
    if port:
        targets.append("port")
    if ref and os.path.isdir(os.path.join(REFERENCE_TREE, "hexl")):
        targets.append("ref")
    if not targets:
        return
    cmd = ["make", "-C", HERE, "-j8", f"REF={REFERENCE_TREE}"] + targets
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data


def _host_has_avx512() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    need = ("avx512f", "avx512dq", "avx512vl", "avx512bw", "avx512cd", "avx512ifma", "avx512_vbmi2")
    return all(f in flags for f in need)


class Port:
    """ctypes view of oracle/hexl_oracle.c (kind = "port")."""

    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(port=True, ref=False)
        L = self.L = C.CDLL(PORT_SO)
        for name in ("multiply_mod", "add_mod", "sub_mod", "pow_mod"):
            f = getattr(L, "orc_" + name)
            f.restype, f.argtypes = u64, [u64, u64, u64]
        L.orc_inverse_mod.restype, L.orc_inverse_mod.argtypes = u64, [u64, u64]
        L.orc_reverse_bits.restype, L.orc_reverse_bits.argtypes = u64, [u64, u64]
        L.orc_is_prime.restype, L.orc_is_prime.argtypes = C.c_int, [u64]
        L.orc_is_primitive_root.restype, L.orc_is_primitive_root.argtypes = C.c_int, [u64, u64, u64]
        L.orc_minimal_primitive_root.restype = u64
        L.orc_minimal_primitive_root.argtypes = [u64, u64]
        L.orc_generate_primes.restype = C.c_int
        L.orc_generate_primes.argtypes = [vp, u64, u64, C.c_int, u64]
        L.orc_multiply_factor.restype = u64
        L.orc_multiply_factor.argtypes = [u64, C.c_uint, u64]
        L.orc_ntt_tables.argtypes = [u64, u64, u64, vp, vp, vp, vp]
        L.orc_ntt_forward.argtypes = [vp, vp, u64, u64, vp, vp, u64, u64, u64, C.c_int]
        L.orc_ntt_inverse.argtypes = [vp, vp, u64, u64, vp, vp, u64, u64, u64, C.c_int]
        L.orc_ntt_forward_textbook.argtypes = [vp, u64, u64, vp]
        L.orc_ntt_inverse_textbook.argtypes = [vp, u64, u64, vp]
        L.orc_eltwise_add_mod.argtypes = [vp, vp, vp, u64, u64]
        L.orc_eltwise_add_mod_scalar.argtypes = [vp, vp, u64, u64, u64]
        L.orc_eltwise_sub_mod.argtypes = [vp, vp, vp, u64, u64]
        L.orc_eltwise_sub_mod_scalar.argtypes = [vp, vp, u64, u64, u64]
        L.orc_eltwise_mult_mod.argtypes = [vp, vp, vp, u64, u64, u64]
        L.orc_eltwise_fma_mod.argtypes = [vp, vp, u64, vp, u64, u64, u64]
        L.orc_eltwise_reduce_mod.argtypes = [vp, vp, u64, u64, u64, u64]
        L.orc_eltwise_cmp_add.argtypes = [vp, vp, u64, C.c_int, u64, u64]
        L.orc_eltwise_cmp_sub_mod.argtypes = [vp, vp, u64, u64, C.c_int, u64, u64]
        L.orc_hensel_lemma_2adic_root.restype, L.orc_hensel_lemma_2adic_root.argtypes = u64, [C.c_uint32, u64]
        L.orc_montgomery_reduce.restype, L.orc_montgomery_reduce.argtypes = u64, [u64, u64, u64, C.c_int, u64]
        L.orc_eltwise_mont_reduce_mod.argtypes = [vp, vp, vp, u64, u64, C.c_int, u64]
        L.orc_eltwise_montgomery_form_in.argtypes = [vp, vp, u64, u64, u64, C.c_int, u64]
        L.orc_eltwise_montgomery_form_out.argtypes = [vp, vp, u64, u64, C.c_int, u64]
        L.orc_dyadic_multiply.argtypes = [vp, vp, vp, u64, vp, u64]
        L.orc_key_switch.argtypes = [vp, vp, u64, u64, u64, u64, u64, vp, vp, vp]
        self._tables = {}

    # -- number theory
    def multiply_mod(self, x, y, q): return self.L.orc_multiply_mod(x, y, q)
    def pow_mod(self, b, e, q): return self.L.orc_pow_mod(b, e, q)
    def inverse_mod(self, x, q): return self.L.orc_inverse_mod(x, q)
    def reverse_bits(self, x, w): return self.L.orc_reverse_bits(x, w)
    def is_prime(self, n): return bool(self.L.orc_is_prime(n))
    def minimal_primitive_root(self, degree, q): return self.L.orc_minimal_primitive_root(degree, q)
    def multiply_factor(self, x, shift, q): return self.L.orc_multiply_factor(x, shift, q)

    def generate_primes(self, num, bits, prefer_small=True, ntt_size=1):
        out = np.zeros(num, dtype=np.uint64)
        got = self.L.orc_generate_primes(_ptr(out), num, bits, int(prefer_small), ntt_size)
        assert got == num, "not enough primes"
        return [int(v) for v in out]

    # -- Montgomery-form helpers
    def hensel_lemma_2adic_root(self, r, q): return self.L.orc_hensel_lemma_2adic_root(r, q)
    def montgomery_reduce(self, t_hi, t_lo, q, r, inv_mod): return self.L.orc_montgomery_reduce(t_hi, t_lo, q, r, inv_mod)

    def mont_reduce_mod(self, a, b, q, r, inv_mod):
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
        res = np.empty_like(a)
        self.L.orc_eltwise_mont_reduce_mod(_ptr(res), _ptr(a), _ptr(b), a.size, q, r, inv_mod)
        return res

    def montgomery_form_in(self, a, r2_mod_q, q, r, inv_mod):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        res = np.empty_like(a)
        self.L.orc_eltwise_montgomery_form_in(_ptr(res), _ptr(a), r2_mod_q, a.size, q, r, inv_mod)
        return res

    def montgomery_form_out(self, a, q, r, inv_mod):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        res = np.empty_like(a)
        self.L.orc_eltwise_montgomery_form_out(_ptr(res), _ptr(a), a.size, q, r, inv_mod)
        return res

    # -- tables / transforms
    def tables(self, n, q, root=None):
        key = (n, q, root)
        if key not in self._tables:
            r = root if root is not None else self.minimal_primitive_root(2 * n, q)
            t = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
            self.L.orc_ntt_tables(n, q, r, *[_ptr(a) for a in t])
            self._tables[key] = (r, *t)
        return self._tables[key]

    def ntt_forward(self, x, n, q, in_mf=1, out_mf=1, root=None, threads=1):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        _, w, wp, _, _ = self.tables(n, q, root)
        out = np.empty_like(x)
        self.L.orc_ntt_forward(_ptr(out), _ptr(x), n, q, _ptr(w), _ptr(wp), in_mf, out_mf,
                               x.size // n, threads)
        return out

    def ntt_inverse(self, x, n, q, in_mf=1, out_mf=1, root=None, threads=1):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        _, _, _, iw, iwp = self.tables(n, q, root)
        out = np.empty_like(x)
        self.L.orc_ntt_inverse(_ptr(out), _ptr(x), n, q, _ptr(iw), _ptr(iwp), in_mf, out_mf,
                               x.size // n, threads)
        return out

    def ntt_forward_textbook(self, x, n, q, root=None):
        out = np.array(x, dtype=np.uint64)
        self.L.orc_ntt_forward_textbook(_ptr(out), n, q, _ptr(self.tables(n, q, root)[1]))
        return out

    def ntt_inverse_textbook(self, x, n, q, root=None):
        out = np.array(x, dtype=np.uint64)
        self.L.orc_ntt_inverse_textbook(_ptr(out), n, q, _ptr(self.tables(n, q, root)[3]))
        return out

    # -- eltwise (numpy in, numpy out)
    def _out(self, a): return np.empty_like(np.ascontiguousarray(a, dtype=np.uint64))

    def add_mod(self, a, b, q):
        a = np.ascontiguousarray(a, dtype=np.uint64); r = self._out(a)
        if np.isscalar(b) or isinstance(b, int):
            self.L.orc_eltwise_add_mod_scalar(_ptr(r), _ptr(a), int(b), a.size, q)
        else:
            b = np.ascontiguousarray(b, dtype=np.uint64)
            self.L.orc_eltwise_add_mod(_ptr(r), _ptr(a), _ptr(b), a.size, q)
        return r

    def sub_mod(self, a, b, q):
        a = np.ascontiguousarray(a, dtype=np.uint64); r = self._out(a)
        if np.isscalar(b) or isinstance(b, int):
            self.L.orc_eltwise_sub_mod_scalar(_ptr(r), _ptr(a), int(b), a.size, q)
        else:
            b = np.ascontiguousarray(b, dtype=np.uint64)
            self.L.orc_eltwise_sub_mod(_ptr(r), _ptr(a), _ptr(b), a.size, q)
        return r

    def mult_mod(self, a, b, q, in_mf=1):
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
        r = self._out(a)
        self.L.orc_eltwise_mult_mod(_ptr(r), _ptr(a), _ptr(b), a.size, q, in_mf)
        return r

    def fma_mod(self, a, b, c, q, in_mf=1):
        a = np.ascontiguousarray(a, dtype=np.uint64); r = self._out(a)
        c = None if c is None else np.ascontiguousarray(c, dtype=np.uint64)
        self.L.orc_eltwise_fma_mod(_ptr(r), _ptr(a), int(b), _ptr(c), a.size, q, in_mf)
        return r

    def reduce_mod(self, a, q, in_mf, out_mf):
        a = np.ascontiguousarray(a, dtype=np.uint64); r = self._out(a)
        self.L.orc_eltwise_reduce_mod(_ptr(r), _ptr(a), a.size, q, in_mf, out_mf)
        return r

    def cmp_add(self, a, cmp, bound, diff):
        a = np.ascontiguousarray(a, dtype=np.uint64); r = self._out(a)
        self.L.orc_eltwise_cmp_add(_ptr(r), _ptr(a), a.size, int(cmp), bound, diff)
        return r

    def cmp_sub_mod(self, a, q, cmp, bound, diff):
        a = np.ascontiguousarray(a, dtype=np.uint64); r = self._out(a)
        self.L.orc_eltwise_cmp_sub_mod(_ptr(r), _ptr(a), a.size, q, int(cmp), bound, diff)
        return r

    # -- SEAL-shaped composites
    def dyadic_multiply(self, op1, op2, n, moduli):
        return _dyadic_call(self.L.orc_dyadic_multiply, op1, op2, n, moduli)

    def key_switch(self, result, t_target, n, decomp, key_mod, rns, kcc, moduli, keys, modswitch):
        return _key_switch_call(self.L.orc_key_switch, result, t_target, n, decomp, key_mod, rns, kcc,
                                moduli, keys, modswitch)


def _key_switch_call(fn, result, t_target, n, decomp, key_mod, rns, kcc, moduli, keys, modswitch):
    """shared marshalling for the two checkers: numpy in, result updated in place"""
    keys = [np.ascontiguousarray(k, dtype=np.uint64) for k in keys]
    kp = (vp * len(keys))(*[k.ctypes.data for k in keys])
    moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
    modswitch = np.ascontiguousarray(modswitch, dtype=np.uint64)
    t_target = np.ascontiguousarray(t_target, dtype=np.uint64)
    fn(_ptr(result), _ptr(t_target), n, decomp, key_mod, rns, kcc, _ptr(moduli), kp, _ptr(modswitch))
    return result


def _dyadic_call(fn, op1, op2, n, moduli):
    op1 = np.ascontiguousarray(op1, dtype=np.uint64)
    op2 = np.ascontiguousarray(op2, dtype=np.uint64)
    moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
    out = np.zeros(3 * n * len(moduli), dtype=np.uint64)
    fn(_ptr(out), _ptr(op1), _ptr(op2), n, _ptr(moduli), len(moduli))
    return out


class Ref:
    """ctypes view of the compiled reference (kind = "reference").

    ``native=True`` on a call selects the reference's scalar C++ tier directly;
    otherwise the reference's own run-time dispatch decides (AVX-512 where the
    host has it)."""

    kind = "reference"

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO) or os.path.exists(REF_SCALAR_SO)

    def __init__(self, force_scalar_lib: bool = False):
        use_avx = (not force_scalar_lib) and os.path.exists(REF_SO) and _host_has_avx512()
        path = REF_SO if use_avx else REF_SCALAR_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run oracle.build() where /root/reference exists")
        self.path, self.avx512 = path, use_avx
        L = self.L = C.CDLL(path)
        L.ref_minimal_primitive_root.restype = u64
        L.ref_minimal_primitive_root.argtypes = [u64, u64]
        L.ref_is_prime.restype, L.ref_is_prime.argtypes = C.c_int, [u64]
        L.ref_generate_primes.restype = C.c_int
        L.ref_generate_primes.argtypes = [vp, u64, u64, C.c_int, u64]
        L.ref_inverse_mod.restype, L.ref_inverse_mod.argtypes = u64, [u64, u64]
        L.ref_pow_mod.restype, L.ref_pow_mod.argtypes = u64, [u64, u64, u64]
        L.ref_multiply_mod.restype, L.ref_multiply_mod.argtypes = u64, [u64, u64, u64]
        L.ref_reverse_bits.restype, L.ref_reverse_bits.argtypes = u64, [u64, u64]
        L.ref_ntt_create.restype, L.ref_ntt_create.argtypes = vp, [u64, u64]
        L.ref_ntt_create_root.restype, L.ref_ntt_create_root.argtypes = vp, [u64, u64, u64]
        L.ref_ntt_destroy.argtypes = [vp]
        L.ref_ntt_root.restype, L.ref_ntt_root.argtypes = u64, [vp]
        L.ref_ntt_tables.argtypes = [vp, vp, vp, vp, vp]
        for name in ("forward", "inverse", "forward_native", "inverse_native"):
            getattr(L, "ref_ntt_" + name).argtypes = [vp, vp, vp, u64, u64, u64, C.c_int]
        L.ref_ntt_forward_textbook.argtypes = [vp, vp]
        L.ref_ntt_inverse_textbook.argtypes = [vp, vp]
        L.ref_ntt_forward_radix4.argtypes = [vp, vp, vp, u64, u64]
        L.ref_ntt_inverse_radix4.argtypes = [vp, vp, vp, u64, u64]
        bt = [u64, C.c_int]
        L.ref_eltwise_add_mod.argtypes = [vp, vp, vp, u64, u64] + bt
        L.ref_eltwise_add_mod_scalar.argtypes = [vp, vp, u64, u64, u64] + bt
        L.ref_eltwise_sub_mod.argtypes = [vp, vp, vp, u64, u64] + bt
        L.ref_eltwise_sub_mod_scalar.argtypes = [vp, vp, u64, u64, u64] + bt
        L.ref_eltwise_mult_mod.argtypes = [vp, vp, vp, u64, u64, u64] + bt
        L.ref_eltwise_fma_mod.argtypes = [vp, vp, u64, vp, u64, u64, u64] + bt
        L.ref_eltwise_reduce_mod.argtypes = [vp, vp, u64, u64, u64, u64] + bt
        L.ref_eltwise_cmp_add.argtypes = [vp, vp, u64, C.c_int, u64, u64] + bt
        L.ref_eltwise_cmp_sub_mod.argtypes = [vp, vp, u64, u64, C.c_int, u64, u64] + bt
        L.ref_eltwise_add_mod_native.argtypes = [vp, vp, vp, u64, u64]
        L.ref_eltwise_add_mod_scalar_native.argtypes = [vp, vp, u64, u64, u64]
        L.ref_eltwise_sub_mod_native.argtypes = [vp, vp, vp, u64, u64]
        L.ref_eltwise_sub_mod_scalar_native.argtypes = [vp, vp, u64, u64, u64]
        L.ref_eltwise_mult_mod_native.argtypes = [vp, vp, vp, u64, u64, u64]
        L.ref_eltwise_fma_mod_native.argtypes = [vp, vp, u64, vp, u64, u64, u64]
        L.ref_eltwise_reduce_mod_native.argtypes = [vp, vp, u64, u64, u64, u64]
        L.ref_eltwise_cmp_add_native.argtypes = [vp, vp, u64, C.c_int, u64, u64]
        L.ref_eltwise_cmp_sub_mod_native.argtypes = [vp, vp, u64, u64, C.c_int, u64, u64]
        self.has_mont = hasattr(L, "ref_eltwise_montgomery")
        if self.has_mont:
            L.ref_hensel_lemma_2adic_root.restype, L.ref_hensel_lemma_2adic_root.argtypes = u64, [C.c_uint32, u64]
            L.ref_montgomery_reduce.restype, L.ref_montgomery_reduce.argtypes = u64, [u64, u64, u64, C.c_int, u64]
            L.ref_eltwise_montgomery.restype = C.c_int
            L.ref_eltwise_montgomery.argtypes = [C.c_int, vp, vp, vp, u64, u64, C.c_int, u64]
        self.has_seal = hasattr(L, "ref_key_switch")
        if self.has_seal:
            L.ref_dyadic_multiply.argtypes = [vp, vp, vp, u64, vp, u64]
            L.ref_key_switch.argtypes = [vp, vp, u64, u64, u64, u64, u64, vp, vp, vp]
        self._ntt = {}

    def tier(self, q: int) -> str:
        """Which NTT tier the reference dispatches for modulus q on this host
        (hexl/ntt/ntt-internal.cpp:202-240)."""
        if not self.avx512 or not self.L.ref_has_avx512dq():
            return "native-radix2"
        if self.L.ref_has_avx512ifma() and q < (1 << 50):
            return "avx512-ifma52"
        return "avx512-dq32" if q < (1 << 30) else "avx512-dq64"

    # -- number theory
    def multiply_mod(self, x, y, q): return self.L.ref_multiply_mod(x, y, q)
    def pow_mod(self, b, e, q): return self.L.ref_pow_mod(b, e, q)
    def inverse_mod(self, x, q): return self.L.ref_inverse_mod(x, q)
    def reverse_bits(self, x, w): return self.L.ref_reverse_bits(x, w)
    def is_prime(self, n): return bool(self.L.ref_is_prime(n))
    def minimal_primitive_root(self, degree, q): return self.L.ref_minimal_primitive_root(degree, q)

    def generate_primes(self, num, bits, prefer_small=True, ntt_size=1):
        out = np.zeros(num, dtype=np.uint64)
        got = self.L.ref_generate_primes(_ptr(out), num, bits, int(prefer_small), ntt_size)
        assert got == num
        return [int(v) for v in out]

    # -- NTT
    def _h(self, n, q, root=None):
        key = (n, q, root)
        if key not in self._ntt:
            self._ntt[key] = (self.L.ref_ntt_create(n, q) if root is None
                              else self.L.ref_ntt_create_root(n, q, root))
        return self._ntt[key]

    def root(self, n, q): return self.L.ref_ntt_root(self._h(n, q))

    def tables(self, n, q, root=None):
        t = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
        self.L.ref_ntt_tables(self._h(n, q, root), *[_ptr(a) for a in t])
        return t

    def _ntt_call(self, fn, x, n, q, in_mf, out_mf, root, threads, out=None):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        if out is None:
            out = np.empty_like(x)
        fn(self._h(n, q, root), _ptr(out), _ptr(x), in_mf, out_mf, x.size // n, threads)
        return out

    def ntt_forward(self, x, n, q, in_mf=1, out_mf=1, root=None, threads=1, native=False, out=None):
        fn = self.L.ref_ntt_forward_native if native else self.L.ref_ntt_forward
        return self._ntt_call(fn, x, n, q, in_mf, out_mf, root, threads, out)

    def ntt_inverse(self, x, n, q, in_mf=1, out_mf=1, root=None, threads=1, native=False, out=None):
        fn = self.L.ref_ntt_inverse_native if native else self.L.ref_ntt_inverse
        return self._ntt_call(fn, x, n, q, in_mf, out_mf, root, threads, out)

    def ntt_forward_textbook(self, x, n, q, root=None):
        out = np.array(x, dtype=np.uint64)
        self.L.ref_ntt_forward_textbook(self._h(n, q, root), _ptr(out))
        return out

    def ntt_inverse_textbook(self, x, n, q, root=None):
        out = np.array(x, dtype=np.uint64)
        self.L.ref_ntt_inverse_textbook(self._h(n, q, root), _ptr(out))
        return out

    def ntt_forward_radix4(self, x, n, q, in_mf=1, out_mf=1):
        x = np.ascontiguousarray(x, dtype=np.uint64); out = np.empty_like(x)
        self.L.ref_ntt_forward_radix4(self._h(n, q), _ptr(out), _ptr(x), in_mf, out_mf)
        return out

    def ntt_inverse_radix4(self, x, n, q, in_mf=1, out_mf=1):
        x = np.ascontiguousarray(x, dtype=np.uint64); out = np.empty_like(x)
        self.L.ref_ntt_inverse_radix4(self._h(n, q), _ptr(out), _ptr(x), in_mf, out_mf)
        return out

    # -- eltwise.  `rows`: the flat input is `rows` independent calls of n/rows elements
    @staticmethod
    def _prep(a, rows):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return a, np.empty_like(a), a.size // rows

    def add_mod(self, a, b, q, native=False, rows=1, threads=1):
        a, r, n = self._prep(a, rows)
        scalar = np.isscalar(b) or isinstance(b, int)
        if not scalar:
            b = np.ascontiguousarray(b, dtype=np.uint64)
        if native:
            (self.L.ref_eltwise_add_mod_scalar_native(_ptr(r), _ptr(a), int(b), a.size, q) if scalar
             else self.L.ref_eltwise_add_mod_native(_ptr(r), _ptr(a), _ptr(b), a.size, q))
        elif scalar:
            self.L.ref_eltwise_add_mod_scalar(_ptr(r), _ptr(a), int(b), n, q, rows, threads)
        else:
            self.L.ref_eltwise_add_mod(_ptr(r), _ptr(a), _ptr(b), n, q, rows, threads)
        return r

    def sub_mod(self, a, b, q, native=False, rows=1, threads=1):
        a, r, n = self._prep(a, rows)
        scalar = np.isscalar(b) or isinstance(b, int)
        if not scalar:
            b = np.ascontiguousarray(b, dtype=np.uint64)
        if native:
            (self.L.ref_eltwise_sub_mod_scalar_native(_ptr(r), _ptr(a), int(b), a.size, q) if scalar
             else self.L.ref_eltwise_sub_mod_native(_ptr(r), _ptr(a), _ptr(b), a.size, q))
        elif scalar:
            self.L.ref_eltwise_sub_mod_scalar(_ptr(r), _ptr(a), int(b), n, q, rows, threads)
        else:
            self.L.ref_eltwise_sub_mod(_ptr(r), _ptr(a), _ptr(b), n, q, rows, threads)
        return r

    def mult_mod(self, a, b, q, in_mf=1, native=False, rows=1, threads=1):
        a, r, n = self._prep(a, rows)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        if native:
            self.L.ref_eltwise_mult_mod_native(_ptr(r), _ptr(a), _ptr(b), a.size, q, in_mf)
        else:
            self.L.ref_eltwise_mult_mod(_ptr(r), _ptr(a), _ptr(b), n, q, in_mf, rows, threads)
        return r

    def fma_mod(self, a, b, c, q, in_mf=1, native=False, rows=1, threads=1):
        a, r, n = self._prep(a, rows)
        c = None if c is None else np.ascontiguousarray(c, dtype=np.uint64)
        if native:
            self.L.ref_eltwise_fma_mod_native(_ptr(r), _ptr(a), int(b), _ptr(c), a.size, q, in_mf)
        else:
            self.L.ref_eltwise_fma_mod(_ptr(r), _ptr(a), int(b), _ptr(c), n, q, in_mf, rows, threads)
        return r

    def reduce_mod(self, a, q, in_mf, out_mf, native=False, rows=1, threads=1):
        a, r, n = self._prep(a, rows)
        if native:
            self.L.ref_eltwise_reduce_mod_native(_ptr(r), _ptr(a), a.size, q, in_mf, out_mf)
        else:
            self.L.ref_eltwise_reduce_mod(_ptr(r), _ptr(a), n, q, in_mf, out_mf, rows, threads)
        return r

    def cmp_add(self, a, cmp, bound, diff, native=False, rows=1, threads=1):
        a, r, n = self._prep(a, rows)
        if native:
            self.L.ref_eltwise_cmp_add_native(_ptr(r), _ptr(a), a.size, int(cmp), bound, diff)
        else:
            self.L.ref_eltwise_cmp_add(_ptr(r), _ptr(a), n, int(cmp), bound, diff, rows, threads)
        return r

    def cmp_sub_mod(self, a, q, cmp, bound, diff, native=False, rows=1, threads=1):
        a, r, n = self._prep(a, rows)
        if native:
            self.L.ref_eltwise_cmp_sub_mod_native(_ptr(r), _ptr(a), a.size, q, int(cmp), bound, diff)
        else:
            self.L.ref_eltwise_cmp_sub_mod(_ptr(r), _ptr(a), n, q, int(cmp), bound, diff, rows, threads)
        return r


def _ref_mont(self, kind, a, b, q, r, inv_mod):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    res = np.empty_like(a)
    self.last_mont_was_avx512 = bool(self.L.ref_eltwise_montgomery(kind, _ptr(res), _ptr(a), _ptr(b), a.size, q, r, inv_mod))
    return res


Ref.hensel_lemma_2adic_root = lambda self, r, q: self.L.ref_hensel_lemma_2adic_root(r, q)
Ref.montgomery_reduce = lambda self, t_hi, t_lo, q, r, inv_mod: self.L.ref_montgomery_reduce(t_hi, t_lo, q, r, inv_mod)
Ref.mont_reduce_mod = lambda self, a, b, q, r, inv_mod: _ref_mont(self, 0, a, b, q, r, inv_mod)
Ref.montgomery_form_in = lambda self, a, r2_mod_q, q, r, inv_mod: _ref_mont(self, 1, a, np.array([r2_mod_q], dtype=np.uint64), q, r, inv_mod)
Ref.montgomery_form_out = lambda self, a, q, r, inv_mod: _ref_mont(self, 2, a, np.zeros(1, dtype=np.uint64), q, r, inv_mod)
Ref.dyadic_multiply = lambda self, op1, op2, n, moduli: _dyadic_call(self.L.ref_dyadic_multiply, op1, op2, n, moduli)
Ref.key_switch = lambda self, result, t_target, n, decomp, key_mod, rns, kcc, moduli, keys, modswitch: _key_switch_call(
    self.L.ref_key_switch, result, t_target, n, decomp, key_mod, rns, kcc, moduli, keys, modswitch)


def best_checker():
    """The strongest checker available on this host: the compiled reference if
    oracle/_ref travelled here, else the restatement."""
    if Ref.available():
        try:
            return Ref()
        except OSError:
            pass
    return Port()

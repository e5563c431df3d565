"""Host-logic test of the CUDA kernels' index arithmetic (no GPU): the numpy
model in tests/kernel_model.py uses the same register/shared-memory/twiddle-node
formulas as hexl_b200/csrc/ntt.cu and must reproduce the oracle's transform for
every row length and column-pass split, with permutation-valid and
bank-conflict-free shared-memory maps."""
import numpy as np
import pytest

from kernel_model import Model, plan_col_passes
from util import uniform_below


def build_model(port, n, q):
    _, w, _, iw, _ = port.tables(n, q)
    inv = [0] * n
    inv[0] = int(iw[0])
    pos, m = 1, n >> 1
    while m > 0:  # undo the reference's stage-sequential re-ordering
        for i in range(m):
            inv[m + i] = int(iw[pos])
            pos += 1
        m >>= 1
    inv_n = pow(n, -1, q)
    return Model(n, q, [int(v) for v in w], inv, inv_n, (inv_n * inv[1]) % q)


@pytest.mark.parametrize("logn,logc", [(4, 4), (5, 5), (6, 6), (7, 7), (8, 8), (9, 9), (10, 10), (11, 11),
                                       (12, 12), (13, 13), (8, 4), (10, 6), (12, 7), (13, 12), (13, 9),
                                       (14, 12), (14, 4), (15, 12)])
def test_model_matches_oracle(port, logn, logc):
    n = 1 << logn
    q = port.generate_primes(1, 28, True, n)[0]
    m = build_model(port, n, q)
    x = uniform_below(logn * 100 + logc, n, q)
    y = port.ntt_forward(x, n, q)
    assert [int(v) for v in m.forward(x, logc)] == [int(v) for v in y]
    rows = n >> logc
    assert m.cta_exchanges == rows * sum(1 for lb in _fwd_exchange_lbs(logc) if lb > 5)
    assert [int(v) for v in m.inverse(y, logc)] == [int(v) for v in x]


def _fwd_exchange_lbs(logc):
    """max(LB) of every exchange of the forward row kernel (CTA barrier iff > 5)"""
    lbs, prev, p = [], logc - 4, 1
    while logc - 4 * p - 1 >= 0:
        lb = max(logc - 4 * p - 4, 0)
        lbs.append(max(prev, lb))
        prev, p = lb, p + 1
    if logc > 4:
        lbs.append(max(0, min(logc - 4, 4)))
    return lbs


def test_one_cta_barrier_per_4096_row():
    assert [lb for lb in _fwd_exchange_lbs(12) if lb > 5] == [8]
    assert [lb for lb in _fwd_exchange_lbs(13) if lb > 5] == [9]
    assert [lb for lb in _fwd_exchange_lbs(10) if lb > 5] == [6]


def test_col_pass_plan():
    assert plan_col_passes(0) == []
    assert plan_col_passes(4) == [4]
    assert plan_col_passes(5) == [5]
    assert plan_col_passes(6) == [3, 3]
    assert plan_col_passes(8) == [4, 4]
    for top in range(1, 17):
        p = plan_col_passes(top)
        assert sum(p) == top and max(p) <= 5


@pytest.mark.parametrize("elem_bytes", [8, 4])
@pytest.mark.parametrize("logc", range(8, 15))
def test_shared_memory_maps_conflict_free(logc, elem_bytes):
    u = np.arange((1 << logc) // 16)
    lbs = {logc - 4, 0, min(logc - 4, 4)}
    p = 1
    while logc - 4 * p - 1 >= 0:
        lbs.add(max(logc - 4 * p - 4, 0))
        p += 1
    for lb in lbs:
        for e in range(16):
            assert Model.bank_conflict_degree(u, e, lb, elem_bytes) == 1


def test_swizzle_and_padding_maps():
    from kernel_model import pad_slot, reg_index, row_elems, swz
    j = np.arange(1 << 14)
    s = swz(j, 4)     # 32-bit rows: XOR swizzle, a bijection inside aligned blocks of 32
    assert (np.sort(s) == j).all() and ((s // 32) == (j // 32)).all()
    p = swz(j, 8)     # 64-bit rows: padded, strictly increasing, inside the padded row
    assert (np.diff(p) >= 1).all() and p[-1] < row_elems(1 << 14)
    u = np.arange(1 << 12)
    for lb in range(0, 11):   # the affine form the kernel uses
        for e in range(16):
            assert (swz(reg_index(u, e, lb), 8) == swz(reg_index(u, 0, lb), 8) + pad_slot(e, lb)).all()


@pytest.mark.parametrize("logr", [2, 3, 4, 5])
def test_cluster_row_ownership_is_consistent(logr):
    """ntt_dsmem_fwd/inv: the column phase stores row e into CTA e % K, slot e // K; the row
    phase of CTA `rank` transforms rows rank + lr*K from slot lr.  Both views must describe the
    same bijection between the R rows and the K x (R/K) shared-memory slots."""
    r = 1 << logr
    k = min(r, 8)
    rpc = r // k
    scatter = {(e % k, e // k): e for e in range(r)}
    assert len(scatter) == r and set(scatter) == {(c, s) for c in range(k) for s in range(rpc)}
    for rank in range(k):
        for lr in range(rpc):
            assert scatter[(rank, lr)] == rank + lr * k

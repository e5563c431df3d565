"""Numpy model of the index arithmetic of hexl_b200/csrc/ntt.cu.

It executes the same decomposition (column passes over tree nodes, row kernel
with 16 coefficients per thread, XOR-swizzled shared-memory transposes) with the
same formulas for register->coefficient maps, swizzle and twiddle-node indices,
vectorised over "threads".  It exists so the kernel's index logic can be tested
on a machine without a GPU (tests/test_kernel_model.py); arithmetic is done
fully reduced (Python-int-safe small moduli), since laziness is not what is
being modelled.
"""
from __future__ import annotations

import numpy as np


def swz(j, elem_bytes=8):
    """ntt_kernels.cuh: 64-bit rows are padded (element j at j + (j >> 4): one 8-byte pad per 16 elements, affine in
    the register slot); the 32-bit rows of SMALL mode XOR a 5-bit field (swz<unsigned>)"""
    return j + (j >> 4) if elem_bytes == 8 else j ^ ((j >> 4) & 31)


def pad_slot(e, lb):
    """the compile-time part of a padded address: pad(U | e << lb) == pad(U) + pad_slot(e, lb)"""
    return (e << lb) + ((e << lb) >> 4)


def row_elems(c, elem_bytes=8):
    return c + (c >> 4 if elem_bytes == 8 else 0)


def reg_index(u, e, lb):
    return ((u >> lb) << (lb + 4)) | (e << lb) | (u & ((1 << lb) - 1))


def plan_col_passes(top_stages):
    if top_stages <= 0:
        return []
    passes = (top_stages + 4) // 5
    out, left = [], top_stages
    for p in range(passes):
        r = (left + (passes - p) - 1) // (passes - p)
        out.append(r)
        left -= r
    return out


def mulmod(a, b, q):
    return (a.astype(object) * b.astype(object)) % q


class Model:
    def __init__(self, n, q, fwd_tree, inv_tree, inv_n, inv_n_w):
        self.n, self.q = n, q
        self.log_n = n.bit_length() - 1
        self.fwd = np.array(fwd_tree, dtype=object)
        self.inv = np.array(inv_tree, dtype=object)
        self.inv_n, self.inv_n_w = inv_n, inv_n_w
        self.cta_exchanges = 0

    # ---- butterflies on object arrays (fully reduced)
    def _fwd_bfly(self, X, Y, w):
        T = (Y * w) % self.q
        return (X + T) % self.q, (X - T) % self.q

    def _inv_bfly(self, X, Y, w):
        return (X + Y) % self.q, ((X - Y) * w) % self.q

    def _inv_last(self, X, Y):
        return ((X + Y) * self.inv_n) % self.q, ((X - Y) * self.inv_n_w) % self.q

    # ---- reg_stages of ntt.cu
    def reg_stages(self, v, u, base, logc, lb, hb, lob, fwd, fold):
        tw = self.fwd if fwd else self.inv
        for step in range(hb - lob + 1):
            beta = hb - step if fwd else lob + step
            eb = beta - lb
            sp = logc - 1 - beta
            node0 = (base << sp) + ((u >> lb) << (lb + 3 - beta))
            if (not fwd) and sp == 0 and fold:
                for l in range(1 << eb):
                    v[l], v[l | (1 << eb)] = self._inv_last(v[l], v[l | (1 << eb)])
            else:
                for g in range(8 >> eb):
                    w = tw[node0 + g]
                    for l in range(1 << eb):
                        e = (g << (eb + 1)) | l
                        f = self._fwd_bfly if fwd else self._inv_bfly
                        v[e], v[e | (1 << eb)] = f(v[e], v[e | (1 << eb)], w)

    def exchange(self, v, u, lb_from, lb_to, c):
        size = row_elems(c)
        smem = np.empty(size, dtype=object)
        writer = np.full(size, -1, dtype=int)
        written = np.zeros(size, dtype=int)
        for e in range(16):
            # the kernel's address form: one per-thread base plus a compile-time slot offset
            base = swz(reg_index(u, 0, lb_from))
            idx = base + pad_slot(e, lb_from)
            assert (idx == swz(reg_index(u, e, lb_from))).all()
            smem[idx] = v[e]
            writer[idx] = u
            np.add.at(written, idx, 1)
        assert written.max() == 1 and written.sum() == c, "smem write map is not injective"
        read = np.zeros(size, dtype=int)
        warp_local = True
        for e in range(16):
            idx = swz(reg_index(u, 0, lb_to)) + pad_slot(e, lb_to)
            assert (written[idx] == 1).all(), "read of a slot nobody wrote"
            v[e] = smem[idx]
            np.add.at(read, idx, 1)
            warp_local &= bool(((writer[idx] >> 5) == (u >> 5)).all())
        assert read.max() == 1 and read.sum() == c, "smem read map is not injective"
        # the kernel uses __syncwarp() instead of __syncthreads() exactly when max(lb) <= 5
        if max(lb_from, lb_to) <= 5:
            assert warp_local, (lb_from, lb_to)
        self.cta_exchanges += 0 if max(lb_from, lb_to) <= 5 else 1

    @staticmethod
    def bank_conflict_degree(u, e, lb, elem_bytes=8):
        """worst multiplicity of a 128-byte-wide bank slot inside one shared-memory wavefront:
        a half-warp of 8-byte accesses, or a full warp of 4-byte accesses"""
        idx = swz(reg_index(u, e, lb), elem_bytes)
        lanes = 16 if elem_bytes == 8 else 32
        worst = 1
        for h in range(0, len(u), lanes):
            banks = idx[h:h + lanes] & (lanes - 1)
            worst = max(worst, int(np.bincount(banks, minlength=lanes).max()))
        return worst

    def row(self, data, logc, base, fwd, fold):
        """data: one row of C coefficients (object array); returns transformed row"""
        c = 1 << logc
        t = c // 16
        u = np.arange(t)
        lb0 = logc - 4
        lb_io = min(lb0, 4)  # 128-byte-line load/store layout of the inverse input / forward output
        passes = (logc + 3) // 4
        if fwd:
            v = [data[reg_index(u, e, lb0)].copy() for e in range(16)]
            self.reg_stages(v, u, base, logc, lb0, logc - 1, lb0, True, False)
            prev_lb = lb0
            for p in range(1, passes + 1):
                hb = logc - 4 * p - 1
                if hb < 0:
                    break
                lb = max(hb - 3, 0)
                self.exchange(v, u, prev_lb, lb, c)
                self.reg_stages(v, u, base, logc, lb, hb, lb, True, False)
                prev_lb = lb
            if logc > 4:
                self.exchange(v, u, 0, lb_io, c)
            out_lb = lb_io
        else:
            v = [data[reg_index(u, e, lb_io)].copy() for e in range(16)]
            if logc > 4:
                self.exchange(v, u, lb_io, 0, c)
            for p in range(passes - 1, -1, -1):
                hb = logc - 4 * p - 1
                lb = max(hb - 3, 0)
                self.reg_stages(v, u, base, logc, lb, hb, lb, False, fold)
                if p > 0:
                    nhb = logc - 4 * (p - 1) - 1
                    nlb = max(nhb - 3, 0)
                    self.exchange(v, u, lb, nlb, c)
            out_lb = lb0
        out = np.empty(c, dtype=object)
        for e in range(16):
            out[reg_index(u, e, out_lb)] = v[e]
        return out

    def col_pass(self, data, logr, log_s, fwd, fold):
        """one column pass over a whole polynomial (object array of N)"""
        n, log_n = self.n, self.log_n
        r = 1 << logr
        log_cols = log_s - logr
        tw = self.fwd if fwd else self.inv
        out = data.copy()
        for blk in range(n >> log_s):
            base = (n >> log_s) + blk
            stw = {}
            for l in range(1, r):
                s = l.bit_length() - 1
                stw[l] = tw[(base << s) + (l - (1 << s))]
            c = np.arange(1 << log_cols)
            off = (blk << log_s) + c
            v = [data[off + (e << log_cols)].copy() for e in range(r)]
            for step in range(logr):
                s = step if fwd else logr - 1 - step
                eb = logr - 1 - s
                if (not fwd) and fold and s == 0 and log_s == log_n:
                    for l in range(1 << eb):
                        v[l], v[l | (1 << eb)] = self._inv_last(v[l], v[l | (1 << eb)])
                else:
                    for gi in range(1 << s):
                        w = stw[(1 << s) + gi]
                        for l in range(1 << eb):
                            e = (gi << (eb + 1)) | l
                            f = self._fwd_bfly if fwd else self._inv_bfly
                            v[e], v[e | (1 << eb)] = f(v[e], v[e | (1 << eb)], w)
            for e in range(r):
                out[off + (e << log_cols)] = v[e]
        return out

    def forward(self, x, log_c):
        data = np.array([int(t) for t in x], dtype=object)
        radices = plan_col_passes(self.log_n - log_c)
        log_s = self.log_n
        for r in radices:
            data = self.col_pass(data, r, log_s, True, False)
            log_s -= r
        c = 1 << log_c
        rows = self.n // c
        out = np.empty(self.n, dtype=object)
        for row in range(rows):
            out[row * c:(row + 1) * c] = self.row(data[row * c:(row + 1) * c], log_c, rows + row, True, False)
        return out

    def inverse(self, x, log_c):
        data = np.array([int(t) for t in x], dtype=object)
        radices = plan_col_passes(self.log_n - log_c)
        c = 1 << log_c
        rows = self.n // c
        out = np.empty(self.n, dtype=object)
        for row in range(rows):
            out[row * c:(row + 1) * c] = self.row(data[row * c:(row + 1) * c], log_c, rows + row, False,
                                                  len(radices) == 0)
        log_s = log_c
        for p in range(len(radices) - 1, -1, -1):
            log_s += radices[p]
            out = self.col_pass(out, radices[p], log_s, False, p == 0)
        return out

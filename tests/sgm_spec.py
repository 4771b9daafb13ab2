"""Numpy statement of the DECOMPOSED aggregation the CUDA kernels implement (sister_b200/csrc/sgm.cu), used by the CPU
tests to prove that the decomposition is bit-identical to the reference recurrence (sgm.cpp:26-455, restated in
oracle/sister_oracle.c so_sgm) before any GPU is involved. Test infrastructure only.

Decomposition:
  * every one of the 8 paths (2 passes x r0..r3) is a set of independent chains; a chain keeps the NORMALISED, CLAMPED
    state  a(d) = min(L(d) - min_d L, P2)  and emits the penalty term  Q(d) = L(d) - C(d) = min(a(d), a(d-1)+P1, a(d+1)+P1)
    in [0, P2] as one byte per cell;
  * diagonal chains wrap around the side borders (the state is reset to P2 on the wrap, which is the reference's
    "predecessor column off the image" rule, sgm.cpp:57-81), so every r1/r2/r3 chain has exactly Hp steps;
  * the first line of a pass contributes only its r0 (int32 arithmetic + 8-bit truncation, sgm.cpp:141-190, types.h:28);
    its "Q" byte is that truncated value itself and r1..r3 write 0 there;
  * S = nC * C + sum of the 8 byte volumes, nC = 8 (4 on the first line of either pass, i.e. rows 0 and Hp-1).
"""
import numpy as np

P1, P2 = 7, 100
INF = 1 << 14


def _step(a, c):
    """a: [n, D] clamped normalised state; c: [n, D] cost. Returns (Q, new a)."""
    up = np.full_like(a, INF); up[:, 1:] = a[:, :-1]
    dn = np.full_like(a, INF); dn[:, :-1] = a[:, 1:]
    q = np.minimum(a, np.minimum(up, dn) + P1)
    L = c + q
    m = L.min(axis=1, keepdims=True)
    return q, np.minimum(L - m, P2)


def _first(c):
    m = c.min(axis=1, keepdims=True)
    return np.minimum(c - m, P2)


def path_volumes(C):
    """C: [h, w, D] ints <= 252. Returns Q [8, h, w, D] uint8 (order: pass 0 r0..r3, pass 1 r0..r3)."""
    h, w, D = C.shape
    C = C.astype(np.int64)
    Q = np.zeros((8, h, w, D), np.int64)
    for p in range(2):
        di = dj = 1 if p == 0 else -1
        i1, j1, jl = (0, 0, w - 1) if p == 0 else (h - 1, w - 1, 0)
        rows = list(range(i1, h if p == 0 else -1, di))
        cols = list(range(j1, w if p == 0 else -1, dj))
        # ---- r0, first line: sgm.cpp:141-190 ----
        last = None
        for n, j in enumerate(cols):
            c = C[i1, j]
            if n == 0:
                nw = c.copy()
            else:
                up = np.full(D, 65535); up[1:] = last[:-1]
                dn = np.full(D, 65535); dn[:-1] = last[1:]
                mp = np.minimum(np.minimum(last, np.minimum(up, dn) + P1), mlast + P2) - mlast
                nw = (c + mp) & 0xFF
            Q[4 * p + 0, i1, j] = nw
            last, mlast = nw, nw.min()
        # ---- r0, other lines: all rows in lockstep along the columns ----
        rr = np.array(rows[1:])
        a = np.zeros((len(rr), D), np.int64)
        for j in cols:
            q, a = _step(a, C[rr, j])
            Q[4 * p + 0, rr, j] = q
        # ---- r2: all columns in lockstep down the rows ----
        a = _first(C[i1])
        for i in rows[1:]:
            q, a = _step(a, C[i])
            Q[4 * p + 2, i] = q
        # ---- r1 (moves +dj per row) and r3 (moves -dj per row): wrapped chains, one per first-line column ----
        for path, sj, enter in ((1, dj, j1), (3, -dj, jl)):
            pos = np.arange(w)
            a = _first(C[i1])
            for i in rows[1:]:
                pos = pos + sj
                wrapped = (pos < 0) | (pos >= w)
                pos = np.where(wrapped, enter, pos)
                a[wrapped] = P2
                q, a = _step(a, C[i, pos])
                Q[4 * p + path, i, pos] = q
    assert Q.max() <= 255
    return Q.astype(np.uint8)


def combine(C, Q):
    h = C.shape[0]
    nC = np.full((h, 1, 1), 8, np.int64)
    nC[0] = 4
    nC[h - 1] = 4
    return (nC * C.astype(np.int64) + Q.astype(np.int64).sum(axis=0)).astype(np.uint16)


def sgm_decomposed(C):
    return combine(C, path_volumes(C))

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


# ---------------------------------------------------------------------------------------------------------------------
# Paired sweeps (round 2; sister_b200/csrc/sgm.cu k_sgm_sweeps): the 8 paths are aggregated by FOUR sweeps of two paths
# each, and every sweep writes ONE byte volume holding the sum of its two penalty terms (<= 2 * P2 off the first lines):
#     sweep 0 (pass 0)  carrier r0 (a row, stepping along the columns)    + rider r1 (predecessor (i-1, j-1))
#     sweep 1 (pass 0)  carrier r2 (a column, stepping down the rows)     + rider r3 (predecessor (i-1, j+1))
#     sweep 2 (pass 1)  carrier r0 (a row, stepping right to left)        + rider r1 (predecessor (i+1, j+1))
#     sweep 3 (pass 1)  carrier r2 (a column, stepping up the rows)       + rider r3 (predecessor (i+1, j-1))
# A sweep is a set of chains n = 0 .. N-1 that all take step t at the same time. The carrier's state stays with its chain;
# the rider's state is handed from chain n-1 to chain n between steps: the rider of chain n at step t continues from what
# chain n-1 left after step t-1 -- which is exactly the diagonal predecessor, because chain n-1 is the neighbouring row
# (column) and step t-1 the neighbouring column (row). No chain ever waits for a chain on its other side, so the chains
# of a sweep form a one-directional pipeline (blocks of chains on the GPU), not a two-sided wavefront.
#     row sweeps (0, 2):    chain n = the n-th row from the pass's first line, step t = the t-th column from its first column
#     column sweeps (1, 3): chain n = the n-th column from the border the diagonal enters at, step t = the t-th row
# Border rules (sgm.cpp:57-81,103-138):
#     first-line cell (row sweeps: n == 0; column sweeps: t == 0): both paths restart from the zero state with the invalid
#         cost 255 read as 0 -- a step from a = 0 gives Q = 0 and a' = min(C - min C, P2) -- except the carrier r0 of a
#         row sweep, whose first line is the literal int32 + 8-bit-truncation arithmetic (sgm.cpp:141-190, types.h:28);
#     rider predecessor off the image (row sweeps: t == 0; column sweeps: n == 0), not on a first line: a = P2;
#     carrier r0 at the start of a row (t == 0): a = 0 (sgm.cpp:215-216).


def _zero_invalid(c):
    return np.where(c == 255, 0, c)


class Sweep:
    """Geometry of one sweep on an h x w frame restricted to the region of interest roi = (r0, r1, c0, c1) and the row band
    [b0, b1) (defaults: everything). N chains, steps [t0, t1); `carrier[n]`: chain n runs its carrier (rows / columns of the
    region); a cell's byte is wanted iff carrier[n] and ts0 <= t < ts1."""

    def __init__(self, s, h, w, roi=None, band=None):
        r0, r1, c0, c1 = roi if roi is not None else (0, h, 0, w)
        b0, b1 = band if band is not None else (0, h)
        self.s, self.h, self.w = s, h, w
        self.row = s in (0, 2)
        self.p = s // 2
        if s == 0:      # rows 0 .. r1-1, columns 0 .. c1-1
            lo, hi = max(b0, 0), min(b1, r1)
            self.n0, self.n1, self.t0, self.t1 = lo, hi, 0, c1
            self.cell = lambda n, t: (n, t)
            self.car = lambda n: r0 <= n < r1
            self.ts0, self.ts1 = c0, c1
        elif s == 2:    # rows h-1 .. r0, columns w-1 .. c0
            lo, hi = max(b0, r0), min(b1, h)
            self.n0, self.n1, self.t0, self.t1 = h - hi, h - lo, 0, w - c0
            self.cell = lambda n, t: (h - 1 - n, w - 1 - t)
            self.car = lambda n: r0 <= h - 1 - n < r1
            self.ts0, self.ts1 = w - c1, w - c0
        elif s == 1:    # columns w-1 .. c0, rows 0 .. r1-1
            lo, hi = max(b0, 0), min(b1, r1)
            self.n0, self.n1, self.t0, self.t1 = 0, w - c0, lo, hi
            self.cell = lambda n, t: (t, w - 1 - n)
            self.car = lambda n: c0 <= w - 1 - n < c1
            self.ts0, self.ts1 = r0, r1
        else:           # columns 0 .. c1-1, rows h-1 .. r0
            lo, hi = max(b0, r0), min(b1, h)
            self.n0, self.n1, self.t0, self.t1 = 0, c1, h - hi, h - lo
            self.cell = lambda n, t: (h - 1 - t, n)
            self.car = lambda n: c0 <= n < c1
            self.ts0, self.ts1 = h - r1, h - r0
        self.empty = self.n1 <= self.n0 or self.t1 <= self.t0


def run_sweep(C, sw, V, state_in=None):
    """Run sweep `sw` on C [h, w, D]; adds nothing, WRITES V[i, j] (int64 [h, w, D]) where wanted. Band hand-over:
    row sweeps take `state_in` = the rider states the last chain of the previous band left, one per step, [t1 - t0, D]
    (None: the band starts at the pass's first line) and return the same for the next band; column sweeps take / return
    (carrier states, rider states), each [N, D], as they stand between two steps."""
    h, w, D = C.shape
    C = C.astype(np.int64)
    if sw.empty:
        return None
    ns = np.arange(sw.n0, sw.n1)
    N = len(ns)
    car = np.array([sw.car(n) for n in ns])
    a_car = np.zeros((N, D), np.int64)                 # carrier states (row sweeps start every row from a = 0)
    rid_prev = np.zeros((N, D), np.int64)              # rider states after the previous step
    fl_last = fl_min = None                            # first-line r0: the truncated vector and its minimum
    exported = []
    if not sw.row and state_in is not None:
        a_car, rid_prev = state_in[0].astype(np.int64).copy(), state_in[1].astype(np.int64).copy()
    for t in range(sw.t0, sw.t1):
        ij = [sw.cell(n, t) for n in ns]
        c = np.stack([C[i, j] for i, j in ij])
        first_line = (ns == 0) if sw.row else np.full(N, t == 0)
        off_image = np.full(N, t == 0) if sw.row else (ns == 0)
        c_eff = np.where(first_line[:, None], _zero_invalid(c), c)
        # ---- rider: continue from chain n-1's state after step t-1
        rin = np.empty((N, D), np.int64)
        rin[1:] = rid_prev[:-1]
        if sw.row:
            rin[0] = state_in[t - sw.t0 - 1] if (state_in is not None and t > sw.t0) else 0  # previous band's last chain
        else:
            rin[0] = 0
        rin[off_image] = P2
        rin[first_line] = 0
        q_rid, rid_prev = _step(rin, c_eff)
        exported.append(rid_prev[-1].copy())
        # ---- carrier
        q_car, a_new = _step(a_car, c_eff)
        if sw.row and ns[0] == 0:  # chain 0 is the pass's first line: literal arithmetic of sgm.cpp:141-190
            c0v = c_eff[0]
            if t == 0:
                nw = c0v.copy()
            else:
                up = np.full(D, 65535); up[1:] = fl_last[:-1]
                dn = np.full(D, 65535); dn[:-1] = fl_last[1:]
                mp = np.minimum(np.minimum(fl_last, np.minimum(up, dn) + P1), fl_min + P2) - fl_min
                nw = (c0v + mp) & 0xFF
            fl_last, fl_min = nw, nw.min()
            q_car[0] = nw
        a_car = a_new
        if sw.ts0 <= t < sw.ts1:
            for k in range(N):
                if car[k]:
                    V[ij[k][0], ij[k][1]] = q_car[k] + q_rid[k]
    if sw.row:
        return np.stack(exported)
    return a_car, rid_prev


def sweep_volumes(C, roi=None):
    """C: [h, w, D] ints (<= 252, or raw costs with the 255 marker). Returns V [4, h, w, D] int64: the four pair volumes
    (zero where not wanted)."""
    h, w, D = C.shape
    V = np.zeros((4, h, w, D), np.int64)
    for s in range(4):
        run_sweep(C, Sweep(s, h, w, roi), V[s])
    assert V.max() <= 255
    return V


def combine4(C, V):
    h = C.shape[0]
    nC = np.full((h, 1, 1), 8, np.int64)
    nC[0] = 4
    nC[h - 1] = 4
    return (nC * C.astype(np.int64) + V.sum(axis=0)).astype(np.uint16)


def sgm_paired(C):
    return combine4(C, sweep_volumes(C))

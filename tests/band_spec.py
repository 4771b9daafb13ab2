"""Numpy stand-in for one row band of the aggregation (sister_b200/bands.py worker interface), built from the pieces of
tests/sgm_spec.py. Test infrastructure only: it lets the CPU tests run the band schedule -- in one process and over
world_size-2/3 gloo -- and prove that chains cut at band borders and continued from the neighbour's state give exactly
the single-band result. The state a band hands over is the chain state AFTER its last row; the border-crossing rule is
applied by the band that takes the next step (the CUDA kernel folds it into the state instead; each worker only has to
agree with itself)."""
import numpy as np
import torch

from sgm_spec import P2, _step


class NumpyBandWorker:
    def __init__(self, C, D, H, W, row0, row1):
        self.C = C.astype(np.int64)            # [Hp, Wp, D] fused cost of the whole frame (only the band's rows are read)
        self.D, self.H, self.W = D, H, W
        self.Hp, self.Wp = C.shape[:2]
        self.row0, self.row1 = row0, row1
        self.Q = np.zeros((8, self.Hp, self.Wp, D), np.int64)
        self.touched = np.zeros((self.Hp,), bool)

    def new_state(self):
        return torch.zeros(3 * self.Wp * self.D, dtype=torch.uint8)

    def _crop_rows(self):
        return max(self.row0, self.D), min(self.row1, self.D + self.H)

    def submit(self):
        lo, hi = self._crop_rows()
        self.touched[self.row0:self.row1] = True
        if hi <= lo:
            return
        rr = np.arange(lo, hi)
        for p, cols in ((0, range(0, self.Wp)), (1, range(self.Wp - 1, -1, -1))):
            a = np.zeros((len(rr), self.D), np.int64)
            for j in cols:
                q, a = _step(a, self.C[rr, j])
                self.Q[4 * p, rr, j] = q

    def vertical(self, p, state_in, want_out):
        dj = 1 if p == 0 else -1
        i1, j1, jl = (0, 0, self.Wp - 1) if p == 0 else (self.Hp - 1, self.Wp - 1, 0)
        rows = range(self.row0, self.row1) if p == 0 else range(self.row1 - 1, self.row0 - 1, -1)
        done = self.row0 if p == 0 else self.Hp - self.row1  # steps behind the chains when they enter the band
        st = None if state_in is None else state_in.numpy().reshape(3, self.Wp, self.D).astype(np.int64)
        out = np.zeros((3, self.Wp, self.D), np.int64)
        for t, (path, sj, enter) in enumerate(((1, dj, j1), (2, 0, 0), (3, -dj, jl))):
            pos = (np.arange(self.Wp) + sj * (done - 1 if done else 0)) % self.Wp  # position on the row before the band
            a = np.zeros((self.Wp, self.D), np.int64) if st is None else st[t]
            for i in rows:
                if i == i1:
                    pos = np.arange(self.Wp)   # first line of the pass: L = C, nothing added (a step from a = 0)
                else:
                    pos = pos + sj
                    wrapped = (pos < 0) | (pos >= self.Wp)
                    pos = np.where(wrapped, enter, pos)
                    a[wrapped] = P2
                q, a = _step(a, self.C[i, pos])
                self.Q[4 * p + path, i, pos] = q
            out[t] = a
        return torch.from_numpy(out.astype(np.uint8).reshape(-1)) if want_out else None

    def finish(self):
        lo, hi = self._crop_rows()
        if hi <= lo:
            return torch.zeros((0, self.W), dtype=torch.int16)
        S = 8 * self.C[lo:hi] + self.Q[:, lo:hi].sum(axis=0)
        disp = np.zeros((hi - lo, self.W), np.int64)
        for jj in range(self.W):
            j = jj + self.D
            disp[:, jj] = S[:, j, : min(j, self.D - 1) + 1].argmin(axis=1)
        enc = np.minimum(disp * 255, 65535).astype(np.uint16)
        return torch.from_numpy(enc.view(np.int16))

"""Numpy stand-in for one row band of the aggregation (sister_b200/bands.py worker interface), built from the paired
sweeps of tests/sgm_spec.py (the formulation sister_b200/csrc/sgm.cu implements). Test infrastructure only: it lets the
CPU tests run the band schedule -- in one process and over world_size-2/3 gloo -- and prove that sweeps cut at band
borders and continued from the neighbour's state give exactly the single-band result.

State a band hands to the next band of a pass (3 * Wp * D bytes, include/sister_b200.h):
    [0]  row sweep: the rider (diagonal path) states its LAST row left, one per step (column), in step order
    [1]  column sweep: the carrier (vertical path) states of all its chains after the band's last row
    [2]  column sweep: the rider (other diagonal) states of all its chains after the band's last row
all clamped normalised vectors a(d) <= P2, one byte per disparity."""
import numpy as np
import torch

from sgm_spec import Sweep, run_sweep


class NumpyBandWorker:
    def __init__(self, C, D, H, W, row0, row1):
        self.C = C.astype(np.int64)            # [Hp, Wp, D] fused cost of the whole frame (only the band's rows are read)
        self.D, self.H, self.W = D, H, W
        self.Hp, self.Wp = C.shape[:2]
        self.row0, self.row1 = row0, row1
        self.roi = (D, D + H, D, D + W)        # Rect(D, D, W, H), hpp:116-118
        self.V = np.zeros((4, self.Hp, self.Wp, D), np.int64)

    def new_state(self):
        return torch.zeros(3 * self.Wp * self.D, dtype=torch.uint8)

    def _crop_rows(self):
        return max(self.row0, self.D), min(self.row1, self.D + self.H)

    def submit(self):
        pass  # every sweep carries a diagonal path: nothing of the aggregation is local to a band any more

    def vertical(self, p, state_in, want_out):
        st = None if state_in is None else state_in.numpy().reshape(3, self.Wp, self.D).astype(np.int64)
        out = np.zeros((3, self.Wp, self.D), np.int64)
        band = (self.row0, self.row1)
        srow, scol = Sweep(2 * p, self.Hp, self.Wp, self.roi, band), Sweep(2 * p + 1, self.Hp, self.Wp, self.roi, band)
        if not srow.empty:
            ex = run_sweep(self.C, srow, self.V[2 * p], None if st is None else st[0][: srow.t1 - srow.t0])
            out[0][: len(ex)] = ex
        if not scol.empty:
            n = scol.n1 - scol.n0
            a, r = run_sweep(self.C, scol, self.V[2 * p + 1], None if st is None else (st[1][:n], st[2][:n]))
            out[1][:n], out[2][:n] = a, r
        return torch.from_numpy(out.astype(np.uint8).reshape(-1)) if want_out else None

    def finish(self):
        lo, hi = self._crop_rows()
        if hi <= lo:
            return torch.zeros((0, self.W), dtype=torch.int16)
        S = 8 * self.C[lo:hi] + self.V[:, lo:hi].sum(axis=0)
        disp = np.zeros((hi - lo, self.W), np.int64)
        for jj in range(self.W):
            j = jj + self.D
            disp[:, jj] = S[:, j, : min(j, self.D - 1) + 1].argmin(axis=1)
        enc = np.minimum(disp * 255, 65535).astype(np.uint16)
        return torch.from_numpy(enc.view(np.int16))


class StreamedNumpyBandWorker(NumpyBandWorker):
    """The same band with the scalable schedule's interface (sister_b200/bands.py: share_match, stream_rows): the WTA share is
    a token that is all-gathered, the row sweeps take their states from the neighbouring band through `mailboxes` -- a dict shared
    by the workers of one process, or torch.distributed send / recv between ranks -- and only the column sweeps are left for
    the two wavefronts. (numpy cannot run the bands' row sweeps at the same time; what is checked is the split of the state
    and the order of the calls.)"""
    share_match = True
    stream_rows = True

    def __init__(self, C, D, H, W, row0, row1, rank, world, mailboxes=None):
        super().__init__(C, D, H, W, row0, row1)
        self.rank, self.world, self.mailboxes = rank, world, mailboxes
        self.calls = []

    def submit_share(self, share, n_shares):
        self.calls.append("share")
        return torch.full((4,), share, dtype=torch.uint8)

    def submit_rest(self, gathered, n_shares):
        assert gathered.tolist() == [r for r in range(n_shares) for _ in range(4)]  # every rank's share, in share order
        self.calls.append("rest")

    def _row_sweep(self, p, st):
        s = Sweep(2 * p, self.Hp, self.Wp, self.roi, (self.row0, self.row1))
        out = np.zeros((self.Wp, self.D), np.int64)
        if not s.empty:
            ex = run_sweep(self.C, s, self.V[2 * p], None if st is None else st[: s.t1 - s.t0])
            out[: len(ex)] = ex
        return out

    def rows(self, passes=3):
        import torch.distributed as dist

        self.calls.append(f"rows{passes}")
        for p in (0, 1):
            if not (passes >> p) & 1:
                continue
            src, dst = (self.rank - 1, self.rank + 1) if p == 0 else (self.rank + 1, self.rank - 1)
            st = None
            if 0 <= src < self.world:
                if self.mailboxes is not None:
                    st = self.mailboxes[(p, src)]
                else:
                    buf = torch.zeros(self.Wp * self.D, dtype=torch.uint8)
                    dist.recv(buf, src=src)
                    st = buf.numpy().reshape(self.Wp, self.D).astype(np.int64)
            out = self._row_sweep(p, st)
            if 0 <= dst < self.world:
                if self.mailboxes is not None:
                    self.mailboxes[(p, self.rank)] = out
                else:
                    dist.send(torch.from_numpy(out.astype(np.uint8).reshape(-1)), dst=dst)

    def columns(self, p, state_in, want_out):
        self.calls.append(f"columns{p}")
        st = None if state_in is None else state_in.numpy().reshape(3, self.Wp, self.D).astype(np.int64)
        out = np.zeros((3, self.Wp, self.D), np.int64)
        s = Sweep(2 * p + 1, self.Hp, self.Wp, self.roi, (self.row0, self.row1))
        if not s.empty:
            n = s.n1 - s.n0
            a, r = run_sweep(self.C, s, self.V[2 * p + 1], None if st is None else (st[1][:n], st[2][:n]))
            out[1][:n], out[2][:n] = a, r
        return torch.from_numpy(out.astype(np.uint8).reshape(-1)) if want_out else None

    def vertical(self, p, state_in, want_out):
        raise AssertionError("the streamed schedule must not fall back to sister_band_vertical")

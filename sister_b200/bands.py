"""Row-band sharding of ONE large frame over the GPUs of one box (BASELINE.json configs[3]).

Each rank owns a contiguous band of rows of the PADDED frame. The volumes (fused cost, four SGM pair volumes, final sum)
are filled for the band only -- and, in an Engine made with max_band_rows, allocated for it only. Staging, census and the
masks are recomputed for the whole frame by every rank; the raw-cost WTA, the largest of the whole-frame stages, is shared:
every rank matches 1/G of each view's rows and the packed int16 maps are all-gathered (include/sister_b200.h, "Row bands";
a worker without share_match -- the numpy stand-in of the CPU tests, or EngineBandWorker(share_match=False) -- computes the
whole frame itself). Every sweep of the
aggregation carries a diagonal path, so all of it crosses the bands: pass 0 flows from rank 0 down to rank G-1, pass 1
from rank G-1 up to rank 0, and each rank continues from the state its neighbour left (band_state_bytes per pass, exact --
no approximate overlapping halos). While rank t works on pass 0, rank G-1-t works on pass 1, so the two wavefronts
overlap; everything else is fully parallel.

`band_program` is the per-rank list of operations; the order of the two messages a pair of neighbours exchanges is the
same on both sides (by the time slot the message is produced in), so blocking sends and receives cannot deadlock,
neither over gloo (CPU tests) nor over NCCL (one ordered stream per peer).
torch.distributed is the plumbing; the worker object is the Engine on the GPU box and a numpy stand-in in the tests.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def band_rows(hp: int, world: int, rank: int) -> Tuple[int, int]:
    """[row0, row1) of the padded frame rank `rank` owns: contiguous, the first hp % world ranks get one extra row."""
    if world <= 0 or not (0 <= rank < world) or hp < world:
        raise ValueError("bad band request")
    base, extra = divmod(hp, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def band_program(world: int, rank: int) -> List[Tuple[str, int, int]]:
    """Operations of one rank, in order: ("recv", pass, peer) | ("compute", pass, -1) | ("send", pass, peer).

    Pass 0 is computed by rank r in time slot r, pass 1 in slot world-1-r. The state of pass 0 goes r -> r+1 after slot r,
    that of pass 1 r -> r-1 after slot world-1-r. A compute gets the key (its slot, 1, pass), a message -- on BOTH of its
    ends -- the key (slot it is produced in, 2, pass); the program is the operations sorted by key. So a rank computes
    before it communicates within a slot (the two wavefronts overlap), a state arrives before the compute that needs it,
    and the messages between a pair of neighbours come in the same order on both ranks, which is what makes blocking
    transfers safe."""
    ops = []
    for p in (0, 1):
        slot = rank if p == 0 else world - 1 - rank
        src = rank - 1 if p == 0 else rank + 1
        dst = rank + 1 if p == 0 else rank - 1
        if 0 <= src < world:
            ops.append(((slot - 1, 2, p), ("recv", p, src)))       # produced by the neighbour in slot - 1: same key as its send
        ops.append(((slot, 1, p), ("compute", p, -1)))
        if 0 <= dst < world:
            ops.append(((slot, 2, p), ("send", p, dst)))
    ops.sort(key=lambda t: t[0])
    prog = [o for _, o in ops]
    # a receive must precede the compute that needs it (it does: slot - 1 < slot) -- checked, not assumed
    for p in (0, 1):
        names = [o[0] for o in prog if o[1] == p]
        assert names == [n for n in ("recv", "compute", "send") if n in names]
    return prog


def simulate_programs(world: int) -> int:
    """Run all ranks' programs with rendezvous (fully synchronous) transfers; returns the number of time steps, raises on
    deadlock. Used by the CPU tests for every world size."""
    progs = [band_program(world, r) for r in range(world)]
    pc = [0] * world
    have = [set() for _ in range(world)]  # passes whose input state has arrived
    steps = 0
    while any(pc[r] < len(progs[r]) for r in range(world)):
        progress = False
        matched = set()
        for r in range(world):
            if pc[r] >= len(progs[r]) or r in matched:
                continue
            kind, p, peer = progs[r][pc[r]]
            if kind == "compute":
                pc[r] += 1
                progress = True
            elif kind == "send":
                if peer not in matched and pc[peer] < len(progs[peer]) and progs[peer][pc[peer]] == ("recv", p, r):
                    pc[r] += 1
                    pc[peer] += 1
                    have[peer].add(p)
                    matched.update((r, peer))
                    progress = True
        if not progress:
            raise RuntimeError(f"deadlock at {[(progs[r][pc[r]] if pc[r] < len(progs[r]) else None) for r in range(world)]}")
        steps += 1
    return steps


def compute_banded(worker, world: int, rank: int, group=None, trace=None):
    """Run this rank's band. `worker` provides
         submit()                              -- whole-frame stages and the fused cost of the band
         submit_share(rank, world) / submit_rest(gathered, world)   -- the same with the raw-cost WTA shared between the
                                                  ranks (used when worker.share_match is set)
         vertical(pass, state_in, want_out)    -- the four paths of one pass inside the band; state_in / return value are torch
                                                  uint8 tensors on the worker's device (or None)
         new_state()                           -- an empty state tensor to receive into
         finish()                              -- final WTA; returns this band's rows of the H x W map (torch int16 tensor
                                                  holding the uint16 bits, possibly 0 rows)
    trace: a list that receives (label, seconds since the call) after each phase has COMPLETED on the device (the worker is
    drained after every phase, so a traced run is a little slower than an untraced one).
    Returns the band's rows."""
    import time

    import torch.distributed as dist

    t_start = time.perf_counter()

    def mark(label):
        if trace is not None:
            worker.drain()
            trace.append((label, time.perf_counter() - t_start))

    if world > 1 and getattr(worker, "share_match", False):
        import torch

        share = worker.submit_share(rank, world)          # this rank's rows of every view's WTA maps, packed
        mark("stage+census+match share")
        gathered = torch.empty(world * share.numel(), dtype=share.dtype, device=share.device)
        dist.all_gather_into_tensor(gathered, share, group=group)
        mark("all-gather WTA maps")
        worker.submit_rest(gathered, world)
        mark("masks+fuse band")
    else:
        worker.submit()
        mark("stage+census+match+masks+fuse band")
    # the row sweeps of all bands run at the same time, their states streamed through mailboxes in the neighbour's memory
    # (worker.stream_rows: the mailboxes were connected by connect_row_mailboxes); then only the column sweeps are left
    # for the two wavefronts
    streamed = world > 1 and getattr(worker, "stream_rows", False)
    if streamed:
        worker.rows()
        mark("row sweeps (streamed between the bands)")
    inbox = {}
    outbox = {}
    for kind, p, peer in band_program(world, rank):
        if kind == "recv":
            buf = worker.new_state()
            dist.recv(buf, src=peer, group=group)
            inbox[p] = buf
            mark(f"recv state pass {p}")
        elif kind == "compute":
            src = rank - 1 if p == 0 else rank + 1
            dst = rank + 1 if p == 0 else rank - 1
            step = worker.columns if streamed else worker.vertical
            outbox[p] = step(p, inbox.get(p) if 0 <= src < world else None, 0 <= dst < world)
            mark(f"{'column sweep' if streamed else 'sweeps'} pass {p}")
        else:
            dist.send(outbox[p], dst=peer, group=group)
            mark(f"send state pass {p}")
    rows = worker.finish()
    mark("final sum+WTA")
    return rows


def run_bands_in_process(workers):
    """All bands in ONE process (workers[r] is rank r's worker): the same programs, messages handed over directly.
    This is how the tests run G bands on one GPU (one slot per band; with an Engine(max_band_rows=...) the G slots together
    hold one frame's worth of volumes). Returns the list of the bands' rows."""
    world = len(workers)
    if world > 1 and all(getattr(w, "share_match", False) for w in workers):
        import torch

        gathered = torch.cat([w.submit_share(r, world) for r, w in enumerate(workers)])
        for w in workers:
            w.submit_rest(gathered, world)
    else:
        for w in workers:
            w.submit()
    streamed = world > 1 and all(getattr(w, "stream_rows", False) for w in workers)
    if streamed:
        # one GPU runs all the bands: a kernel whose producer is queued BEHIND it would wait for nothing, so the producers
        # go first -- pass 0 top to bottom, then pass 1 bottom to top (on G GPUs every band has its own and one launch does both)
        for w in workers:
            w.rows(passes=1)
        for w in reversed(workers):
            w.rows(passes=2)
    progs = [band_program(world, r) for r in range(world)]
    pc = [0] * world
    inbox = [dict() for _ in range(world)]
    outbox = [dict() for _ in range(world)]
    while any(pc[r] < len(progs[r]) for r in range(world)):
        progress = False
        for r in range(world):
            if pc[r] >= len(progs[r]):
                continue
            kind, p, peer = progs[r][pc[r]]
            if kind == "compute":
                src = r - 1 if p == 0 else r + 1
                dst = r + 1 if p == 0 else r - 1
                step = workers[r].columns if streamed else workers[r].vertical
                outbox[r][p] = step(p, inbox[r].get(p) if 0 <= src < world else None, 0 <= dst < world)
            elif kind == "send":
                if not (pc[peer] < len(progs[peer]) and progs[peer][pc[peer]] == ("recv", p, r)):
                    continue
                inbox[peer][p] = outbox[r][p]
                pc[peer] += 1
            else:
                continue  # a receive completes together with the matching send
            pc[r] += 1
            progress = True
        if not progress:
            raise RuntimeError("band programs deadlocked")
    return [w.finish() for w in workers]


def crop_rows_of_band(D: int, H: int, row0: int, row1: int) -> Tuple[int, int]:
    """Rows [a, b) of the H x W output map that the padded-frame band [row0, row1) holds (possibly empty)."""
    a, b = max(row0, D) - D, min(row1, D + H) - D
    return (a, b) if b > a else (0, 0)


def gather_band_rows(local_rows, D: int, H: int, hp: int, dst: int = 0, group=None):
    """Gather the ranks' rows (torch int16 [n_rows_local, W]) into the full [H, W] map on rank `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return local_rows
    spans = [crop_rows_of_band(D, H, *band_rows(hp, world, r)) for r in range(world)]
    cap = max(b - a for a, b in spans)
    w = local_rows.shape[1]
    send = torch.zeros((cap, w), dtype=local_rows.dtype, device=local_rows.device)
    send[: local_rows.shape[0]] = local_rows
    send = send.contiguous().view(torch.uint8)  # the collective moves bytes
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([recv[r].view(local_rows.dtype)[: spans[r][1] - spans[r][0]] for r in range(world)], dim=0)


class EngineBandWorker:
    """The band worker on the GPU box: drives sister_band_* of one Engine (one context on this rank's GPU)."""

    def __init__(self, engine, views, disp_count: int, rank: int, world: int, mode: int = 0, slot: int = 0, share_match: bool = True):
        import torch

        self.torch = torch
        self.eng = engine
        v0 = views[0]
        self.h, self.w = v0.shape[:2]
        self.ch = 3 if v0.ndim == 3 else 1
        self.D = disp_count
        self.mode = mode
        self.slot = slot
        self.share_match = share_match
        self.hp = self.h + 2 * disp_count
        self.row0, self.row1 = band_rows(self.hp, world, rank)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.rig = engine.upload_rig(views)
        self.state_bytes = engine.band_state_bytes(self.w, self.h, disp_count)
        self.out = torch.zeros((self.h, self.w), dtype=torch.int16, device=self.device)

    def new_state(self):
        return self.torch.empty(self.state_bytes, dtype=self.torch.uint8, device=self.device)

    def submit(self):
        self.eng.band_submit(self.slot, self.rig, self.w, self.h, self.ch, self.D, self.mode, self.row0, self.row1)

    def submit_share(self, share: int, n_shares: int):
        """Whole-frame staging and census, this share of the WTA rows; returns the packed maps (uint8 tensor, complete)."""
        nbytes = self.eng.band_share_bytes(self.w, self.h, self.D, n_shares)
        out = self.torch.zeros(nbytes, dtype=self.torch.uint8, device=self.device)
        self.torch.cuda.current_stream().synchronize()  # (the zero fill runs on torch's stream, the library on its own)
        self.eng.band_submit_share(self.slot, self.rig, self.w, self.h, self.ch, self.D, self.mode, self.row0, self.row1, share, n_shares,
                                   out.data_ptr())
        self.eng.sync(self.slot)
        return out

    def submit_rest(self, gathered, n_shares: int):
        self.torch.cuda.current_stream().synchronize()  # the gathered shares were written on torch's stream
        self.eng.band_submit_rest(self.slot, gathered.data_ptr(), n_shares)

    # ---- row sweeps streamed between the bands
    def make_row_mailboxes(self):
        """This band's two mailboxes (pass 0: written by the band above, pass 1: by the band below), zero-filled device memory
        from the library's allocator (so that it can be exported to the neighbours' processes)."""
        self.mbox = [self.eng.dev_alloc(self.state_bytes) for _ in range(2)]
        for m in self.mbox:
            self.eng.dev_memset(m, 0, self.state_bytes)
        self.peer_out = [0, 0]   # pass 0: the mailbox of the band below, pass 1: of the band above
        self.stream_rows = True
        self.frame = 0
        return self.mbox

    def rows(self, passes: int = 3):
        if passes != 2:            # a new frame starts with pass 0 (or with the launch that does both)
            self.frame += 1
        tag = (self.frame - 1) % 15 + 1
        self.eng.band_rows(self.slot, passes, self.mbox[0], self.peer_out[0], self.mbox[1], self.peer_out[1], tag)

    def columns(self, p: int, state_in, want_out: bool):
        out = self.new_state() if want_out else None
        if state_in is not None:
            self.torch.cuda.current_stream().synchronize()
        self.eng.band_columns(self.slot, p, state_in.data_ptr() if state_in is not None else 0, out.data_ptr() if want_out else 0)
        if want_out:
            self.eng.band_columns_wait(self.slot)  # the state must be complete before it is sent (the row sweeps, on the slot's
        return out                                 # main stream, are not waited for)

    def vertical(self, p: int, state_in, want_out: bool):
        out = self.new_state() if want_out else None
        if state_in is not None:
            self.torch.cuda.current_stream().synchronize()  # the received state was written on torch's stream
        self.eng.band_vertical(self.slot, p, state_in.data_ptr() if state_in is not None else 0, out.data_ptr() if want_out else 0)
        if want_out:
            self.eng.sync(self.slot)  # the library runs on its own stream: the state must be complete before it is sent
        return out

    def drain(self):
        self.eng.sync(self.slot)
        self.eng.band_columns_wait(self.slot)
        self.torch.cuda.synchronize()

    def finish(self):
        self.eng.band_finish(self.slot, self.out.data_ptr())
        self.eng.sync(self.slot)
        a, b = crop_rows_of_band(self.D, self.h, self.row0, self.row1)
        return self.out[a:b]


def connect_row_mailboxes(worker, world: int, rank: int, group=None):
    """One process per GPU: every worker makes its two mailboxes, the 64-byte IPC handles go round (all_gather_object), and each
    worker opens the neighbours' -- pass 0 writes into the band below's, pass 1 into the band above's (peer access over NVLink).
    Returns whether the mailboxes are connected (the same answer on every rank)."""
    import torch.distributed as dist

    mbox = worker.make_row_mailboxes()
    handles = [None] * world
    dist.all_gather_object(handles, [worker.eng.ipc_export(m) for m in mbox], group=group)
    ok = True
    try:
        if rank + 1 < world:
            worker.peer_out[0] = worker.eng.ipc_open(handles[rank + 1][0])
        if rank > 0:
            worker.peer_out[1] = worker.eng.ipc_open(handles[rank - 1][1])
    except Exception:  # no peer access between two of the GPUs: every rank falls back to the hand-over at the end of a band
        ok = False
    oks = [None] * world
    dist.all_gather_object(oks, ok, group=group)
    if not all(oks):
        disconnect_row_mailboxes(worker, group=group)
        return False
    return True


def disconnect_row_mailboxes(worker, group=None):
    """Close the neighbours' mailboxes, then (after everybody has) free this worker's own."""
    import torch.distributed as dist

    worker.eng.sync(worker.slot)
    for p in worker.peer_out:
        if p:
            worker.eng.ipc_close(p)
    worker.peer_out = [0, 0]
    worker.stream_rows = False
    if dist.is_initialized():
        dist.barrier(group=group)
    for m in worker.mbox:
        worker.eng.dev_free(m)
    worker.mbox = []


def connect_row_mailboxes_in_process(workers):
    """All bands in one process (one GPU): the neighbours' mailboxes are plain pointers."""
    for w in workers:
        w.make_row_mailboxes()
    for r, w in enumerate(workers):
        if r + 1 < len(workers):
            w.peer_out[0] = workers[r + 1].mbox[0]
        if r > 0:
            w.peer_out[1] = workers[r - 1].mbox[1]


def as_uint16(t) -> np.ndarray:
    return t.cpu().numpy().view(np.uint16)

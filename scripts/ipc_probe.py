#!/usr/bin/env python
"""Feasibility probe (2 ranks under torchrun): rank 0 exports a device buffer, rank 1 opens it and writes into it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sister_b200  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = sister_b200.Engine(96, 64, 32, n_slots=1, device=local)
n = 1 << 20
mine = eng.dev_alloc(n)
eng.dev_memset(mine, 0, n)
handles = [None] * world
dist.all_gather_object(handles, eng.ipc_export(mine))
peer = eng.ipc_open(handles[(rank + 1) % world])
eng.dev_upload(peer, np.full(n, rank + 1, np.uint8))  # write into the NEXT rank's buffer
torch.cuda.synchronize()
dist.barrier()
got = np.zeros(n, np.uint8)
eng.dev_download(mine, got)
print(rank, "sees", int(got[0]), int(got[-1]), "expected", (rank - 1) % world + 1, flush=True)
dist.barrier()
eng.ipc_close(peer)
dist.barrier()
eng.dev_free(mine)
eng.close()
dist.destroy_process_group()

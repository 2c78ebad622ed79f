"""configs[2]: batched perft(4) over every position of the Chess324 stand-in book, sharded over the ranks of a
torchrun launch (position i on rank i mod world), counts gathered with one all-reduce and checked per position
against the oracle on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/book_multi.py
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import workloads  # noqa: E402

import pabi_b200 as pb  # noqa: E402
from pabi_b200 import distributed as D  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
device = torch.device("cuda", local)
engine = pb.Engine(local)
book = workloads.chess324_synth_book()
DEPTH = 4
for _ in range(3):  # warm-up
    counts = D.perft_batch_distributed(engine, book, DEPTH, device)
steps, dev_ms = 5, 0.0
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    counts = D.perft_batch_distributed(engine, book, DEPTH, device)
    dev_ms += engine.last_timing.device_ms
torch.cuda.synchronize()
wall_ms = (time.perf_counter() - t0) * 1e3
dev_ms, wall_ms = D.all_reduce_max([dev_ms, wall_ms], device)
if rank == 0:
    import oracle
    want = [oracle.perft(oracle.Position.from_buffer_copy(book[i].tobytes()), DEPTH, 8) for i in range(len(book))]
    assert counts.tolist() == want, "per-position counts differ from the oracle"
    nodes = int(counts.sum())
    print(json.dumps({"config": f"perft({DEPTH}) x {len(book)} chess324-synth positions, sharded by position over {world} GPU(s)",
                      "nodes": nodes, "n_gpus": world, "device_ms_per_pass": round(dev_ms / steps, 4), "e2e_ms_per_pass": round(wall_ms / steps, 4),
                      "gnps_device": round(nodes / (dev_ms / steps) / 1e6, 1), "gnps_e2e": round(nodes / (wall_ms / steps) / 1e6, 1),
                      "checked": "every position against the oracle"}), flush=True)
engine.close()
if world > 1:
    dist.destroy_process_group()

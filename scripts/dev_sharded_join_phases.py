"""Development aid: where the time of the sharded 1 M-row join goes, per rank (run under torchrun).
    torchrun --nproc-per-node 4 scripts/dev_sharded_join_phases.py [rows_total]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_join_data  # noqa: E402
from video_fingerprint_b200 import sharding  # noqa: E402
from video_fingerprint_b200.fingerprint import threshold_join_device  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
E = make_join_data(n_total, dev, seed=11)
lo, hi = sharding.row_block(n_total, world, rank)
local = E[lo:hi].contiguous()
del E


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for it in range(int(os.environ.get("ITERS", "3"))):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    marks = [("start", ev())]
    gather = sharding._RowGather(local)
    counts = gather.counts
    marks.append(("gather issued", ev()))
    threshold_join_device(local, 0.95, q=local, q_row0=lo)
    marks.append(("diagonal block", ev()))
    full, _ = gather.wait()
    marks.append(("gather wait", ev()))
    for q_lo, q_hi, c_lo, c_hi in sharding.symmetric_block_plan(world, rank, counts)[1:]:
        q = local if (q_lo, q_hi) == (0, local.shape[0]) else local[q_lo:q_hi]
        threshold_join_device(full[c_lo:c_hi], 0.95, q=q, q_row0=lo + q_lo)
        marks.append((f"block {q_hi - q_lo} x {c_hi - c_lo}", ev()))
    torch.cuda.synchronize()
    line = ", ".join(f"{name} {marks[k][1].elapsed_time(e):.2f}" for k, (name, e) in enumerate(marks[1:]))
    print(f"iter {it} rank {rank}: total {marks[0][1].elapsed_time(marks[-1][1]):.2f} ms: {line}", flush=True)
if world > 1:
    dist.destroy_process_group()

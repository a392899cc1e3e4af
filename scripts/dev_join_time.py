"""Dev helper: join throughput of one row block against databases of growing size, by L2 prefetch distance (tuning key 5)."""
import sys
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

lib = _native.load()
n_q = 262144
dists = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 16]
g = torch.Generator(device="cuda").manual_seed(1)
for n_db in (262144, 1048576, 2097152):
    db = torch.randn((n_db, 256), generator=g, device="cuda")
    db = db / db.norm(dim=1, keepdim=True)
    q = db[:n_q].contiguous()
    for d in dists:
        lib.vfp_set_tuning(5, d)
        i, j, s = vfp.threshold_join_device(db, 0.95, q=q, q_row0=0)
        cap = int(i.numel()) + 4096
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            i, j, s = vfp.threshold_join_device(db, 0.95, q=q, q_row0=0, capacity=cap)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"n_db {n_db:8d} prefetch {d:3d}: {ms:7.2f} ms  {n_q * n_db / ms / 1e6:7.0f} Gpairs/s  {n_q * n_db * 512 / ms / 1e9:6.0f} TFLOP/s  pairs {i.numel()}", flush=True)
    del db, q

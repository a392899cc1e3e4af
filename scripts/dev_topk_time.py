"""Dev helper: top-k search throughput by L2 prefetch distance (tuning key 6)."""
import sys
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

lib = _native.load()
n_q, n_db = 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 4194304
dists = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 8]
g = torch.Generator(device="cuda").manual_seed(1)
db = torch.randn((n_db, 256), generator=g, device="cuda"); db = db / db.norm(dim=1, keepdim=True)
q = torch.randn((n_q, 256), generator=g, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
ref = None
for d in dists:
    lib.vfp_set_tuning(6, d)
    S, I = vfp.topk_inner_product_device(q, db, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S, I = vfp.topk_inner_product_device(q, db, 10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if ref is None: ref = I.clone()
    print(f"top-10 {n_q} x {n_db} prefetch {d:3d}: {ms:8.2f} ms {n_q * n_db / ms / 1e6:7.0f} Gpairs/s  same indices as first: {bool(torch.equal(I, ref))}", flush=True)

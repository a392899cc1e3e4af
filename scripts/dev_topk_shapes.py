"""Development: top-k search time on one GPU for the query counts a rank sees when cfg 5 (100 000 queries x 10 M rows) is
sharded over 1, 2, 4, 8 GPUs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_fingerprint_b200 as vfp
n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
g = torch.Generator(device="cuda").manual_seed(21)
db = torch.randn((n_db, 256), generator=g, device="cuda"); db /= db.norm(dim=1, keepdim=True)
for n_q in (12_500, 25_000, 50_000, 100_000):
    Q = torch.randn((n_q, 256), generator=g, device="cuda"); Q /= Q.norm(dim=1, keepdim=True)
    vfp.topk_inner_product_device(Q, db, 10)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        S, I = vfp.topk_inner_product_device(Q, db, 10)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    print(f"n_q {n_q}: {ms:.1f} ms, {n_q * n_db * 512 / ms / 1e9:.0f} TFLOP/s", flush=True)

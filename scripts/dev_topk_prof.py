import torch, sys
sys.path.insert(0, ".")
import video_fingerprint_b200 as vfp
g = torch.Generator(device="cuda").manual_seed(1)
db = torch.randn((1048576, 256), generator=g, device="cuda"); db = db / db.norm(dim=1, keepdim=True)
q = torch.randn((16384, 256), generator=g, device="cuda"); q = q / q.norm(dim=1, keepdim=True)
for _ in range(2):
    S, I = vfp.topk_inner_product_device(q, db, 10)
torch.cuda.synchronize()

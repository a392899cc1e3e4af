"""Development: a few forward passes over N clips x 64 frames (for ncu captures). usage: dev_forward_small.py [clips] [key=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

lib = _native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    assert lib.vfp_set_tuning(int(k), int(v)) == 0
torch.manual_seed(0)
m = vfp.create_model("attention").eval()
frames = (torch.rand((n * 64, 3, 64, 64), device="cuda") ).to(torch.bfloat16)
for _ in range(3):
    e = m.fingerprint_packed(frames, [64] * n)
torch.cuda.synchronize()
print("ok", float(e.norm(dim=1).mean()), lib.vfp_device_error_word())

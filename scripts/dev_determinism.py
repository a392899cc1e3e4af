"""Dev helper: run the default forward several times on the same packed input and compare bit patterns."""
import sys
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 600
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = _native.load()
lib.vfp_set_tuning(1, mode)
torch.manual_seed(0)
m = vfp.create_model("attention").eval()
x = torch.rand(n_clips * 24, 3, 64, 64)
u8 = torch.round(x * 255).to(torch.uint8)
for name, fr in (("u8", u8), ("u8_hwc", u8.permute(0, 2, 3, 1).contiguous()), ("bf16", x.to(torch.bfloat16))):
    fr = fr.cuda()
    outs = [m.fingerprint_packed(fr, [24] * n_clips).cpu() for _ in range(4)]
    bad = [int((outs[0] != o).any(dim=1).sum()) for o in outs[1:]]
    print(name, "clips differing from run 0:", bad, "max abs diff", [float((outs[0] - o).abs().max()) for o in outs[1:]])

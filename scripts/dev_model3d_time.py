"""Dev helper: throughput of the 3-D CNN model forward (bf16 frames resident in HBM)."""
import sys, time
import torch
import video_fingerprint_b200 as vfp

n, t, fs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 64, 16
torch.manual_seed(0)
m = vfp.create_model("3d", frame_stride=fs).eval()
m.clips_per_pass = int(sys.argv[2]) if len(sys.argv) > 2 else 256
x = torch.rand(n, t, 3, 64, 64, device="cuda").to(torch.bfloat16)
for _ in range(2):
    m(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"3-D model, {n} clips x {t} frames, frame_stride {fs}, {m.clips_per_pass} clips per pass: {ms:.2f} ms -> {n / ms * 1e3:.0f} videos/s", flush=True)

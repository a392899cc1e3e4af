"""Development: phase trace of the fused feed-forward kernel (block 0) from a -DVFP_FFN_TRACE build (VFP_B200_LIB=build/libvfp_trace.so)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native
lib = _native.load()
torch.manual_seed(0)
m = vfp.create_model("attention").eval()
n = 2048
frames = torch.rand((n * 64, 3, 64, 64), device="cuda").to(torch.bfloat16)
m.fingerprint_packed(frames, [64] * n)
buf = (C.c_longlong * 8192)()
lib.vfp_debug_ffn_trace.restype = C.c_int
lib.vfp_debug_ffn_trace(buf, 4096)      # drop the warm-up
m.fingerprint_packed(frames, [64] * n)
k = lib.vfp_debug_ffn_trace(buf, 4096)
ev = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(k) if buf[2 * i + 1] != 0)
t0 = ev[0][0]
names = {100: "I g1 wait d1_empty", 101: "I g1 go", 102: "I g1 issued", 110: "I g2 wait h_full", 111: "I g2 go", 112: "I g2 issued",
         200: "A wait d1_full", 201: "A d1_full", 202: "A ld half0 done", 203: "A ld half1 done", 204: "A gelu done", 205: "A output done", 206: "A h_empty ok", 207: "A h written"}
for t, tag in ev[40:200]:
    print(f"{t - t0:9d}  {names.get(tag, tag)}")

"""Dev helper: time of the stem kernel, optionally with parts of its pipeline compiled out.

The knock-out mask is a COMPILE-TIME switch of stem_ts_kernel.cuh (-DVFP_STEM_KNOCKOUT=mask, bits documented there; results
are garbage, only the time matters). Build the variants next to the real library and point the loader at one of them:

    for m in 1 2 4 8 16 32 64 127 1023; do
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC \
           -DVFP_STEM_KNOCKOUT=$m -shared -o build/libvfp_ko$m.so video_fingerprint_b200/csrc/vfp_b200.cu; done
    VFP_B200_LIB=build/libvfp_ko127.so PYTHONPATH=. python scripts/dev_knockout.py 1024
"""
import ctypes as C
import sys

import torch

import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = _native.load()
lib.vfp_set_tuning(1, mode)
torch.manual_seed(0)
m = vfp.create_model("attention").eval()
fr = torch.rand(n_clips * 64, 3, 64, 64, device="cuda").to(torch.bfloat16)
lengths = [64] * n_clips
stage_ms = (C.c_double * 32)()
launches = C.c_uint64(0)
names = [lib.vfp_profile_stage_name(i).decode() for i in range(lib.vfp_profile_num_stages())]
for _ in range(2):
    m.fingerprint_packed(fr, lengths)
torch.cuda.synchronize()
lib.vfp_profile_enable(1)
for _ in range(3):
    m.fingerprint_packed(fr, lengths)
torch.cuda.synchronize()
lib.vfp_profile_read(stage_ms, 32, C.byref(launches), 1)
lib.vfp_profile_enable(0)
t = stage_ms[names.index("stem_fused")] / 3
print(f"{_native.LIB_PATH} mode {mode}: stem {t:.3f} ms  ({t * 1e6 / (n_clips * 64) * 148:.0f} ns per frame per SM)  err {lib.vfp_device_error_word():#x}", flush=True)

"""Dev helper: stem_ts_kernel stage time with parts of the pipeline knocked out (vfp_set_tuning key 4 mask)."""
import ctypes as C
import sys

import torch

import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
masks = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 4, 8, 16, 32, 64]
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib = _native.load()
lib.vfp_set_tuning(1, mode)
torch.manual_seed(0)
m = vfp.create_model("attention").eval()
fr = torch.rand(n_clips * 64, 3, 64, 64, device="cuda").to(torch.bfloat16)
lengths = [64] * n_clips
stage_ms = (C.c_double * 32)()
launches = C.c_uint64(0)
names = [lib.vfp_profile_stage_name(i).decode() for i in range(lib.vfp_profile_num_stages())]
for mask in masks:
    lib.vfp_set_tuning(4, mask)
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
    if mask & 2048:
        t64 = (C.c_longlong * 32)()
        lib.vfp_debug_stem_timers(t64, 32, 1)
        n_fr = (n_clips * 64 + 147) // 148 * 5   # frames of block 0 over the 5 forwards
        sites = {0: "issuer a_full", 1: "issuer d_empty", 2: "issuer acc_empty", 3: "issuer c1_full", 5: "gen tile_full", 6: "gen a_empty",
                 8: "epi1 d_full", 9: "epi1 c1_empty", 10: "epi1 halo c1_empty", 11: "epi2 acc_full", 13: "xpose tile_empty", 14: "xpose raw_full",
                 16: "copy raw_empty", 20: "TOTAL copy", 21: "TOTAL issuer", 22: "TOTAL xpose", 23: "TOTAL epi2", 24: "TOTAL gen", 25: "TOTAL epi1"}
        print("   cycles per frame (block 0):", {v: round(t64[k] / n_fr) for k, v in sites.items()})
    print(f"mask {mask:3d}: stem {t:.3f} ms  ({t * 1e6 / (n_clips * 64) * 148:.0f} ns per frame per SM)  err {lib.vfp_device_error_word():#x}", flush=True)

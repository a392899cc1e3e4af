"""Dev helper: A/B a tuning key on the full forward: python scripts/dev_ab.py <clips> <key> <v0,v1,..>; prints stage times and
the cosine of each variant's embeddings against the first one."""
import ctypes as C
import sys
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

n_clips, key = int(sys.argv[1]), int(sys.argv[2])
values = [int(v) for v in sys.argv[3].split(",")]
lib = _native.load()
torch.manual_seed(0)
m = vfp.create_model("attention").eval()
fr = torch.rand(n_clips * 64, 3, 64, 64, device="cuda").to(torch.bfloat16)
lengths = [64] * n_clips
stage_ms = (C.c_double * 32)()
launches = C.c_uint64(0)
names = [lib.vfp_profile_stage_name(i).decode() for i in range(lib.vfp_profile_num_stages())]
ref = None
for v in values:
    lib.vfp_set_tuning(key, v)
    for _ in range(2):
        out = m.fingerprint_packed(fr, lengths)
    torch.cuda.synchronize()
    lib.vfp_profile_enable(1)
    for _ in range(3):
        out = m.fingerprint_packed(fr, lengths)
    torch.cuda.synchronize()
    lib.vfp_profile_read(stage_ms, 32, C.byref(launches), 1)
    lib.vfp_profile_enable(0)
    out = out.cpu()
    if ref is None:
        ref = out
    cos = torch.nn.functional.cosine_similarity(out.double(), ref.double(), dim=1).min()
    st = {n: round(stage_ms[i] / 3, 3) for i, n in enumerate(names) if stage_ms[i] > 0}
    print(f"key {key} = {v}: total {sum(st.values()):.3f} ms, min cos vs first {float(cos):.7f}, err {lib.vfp_device_error_word():#x}\n   {st}", flush=True)

"""Development check: embeddings under a vfp_set_tuning setting against a base setting (same inputs) and the oracle.
usage: dev_tuning_parity.py [base key=value ...] -- key=value [key=value ...]   (no "--": the base is the default setting)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native
from oracle.forward_oracle import fingerprint_clips
from oracle.weights import make_clips, make_state_dict

lib = _native.load()
sd = make_state_dict(2, "stress")
m = vfp.create_model("attention").eval()
m.load_state_dict(sd)
g = torch.Generator().manual_seed(5)
lengths = [int(t) for t in torch.randint(10, 200, (700,), generator=g)]
frames = torch.randint(0, 256, (sum(lengths), 3, 64, 64), dtype=torch.uint8, generator=g).cuda()
args = sys.argv[1:]
base_t, new_t = (args[: args.index("--")], args[args.index("--") + 1 :]) if "--" in args else ([], args)
for kv in base_t:
    k, v = kv.split("=")
    assert lib.vfp_set_tuning(int(k), int(v)) == 0, kv
base = m.fingerprint_packed(frames, lengths)
for kv in new_t:
    k, v = kv.split("=")
    assert lib.vfp_set_tuning(int(k), int(v)) == 0, kv
new = m.fingerprint_packed(frames, lengths)
torch.cuda.synchronize()
err = lib.vfp_device_error_word()
cos = torch.nn.functional.cosine_similarity(base.double(), new.double(), dim=1)
print(f"tuning {sys.argv[1:]}: device error {err:#x}, max |diff| {float((base - new).abs().max()):.3e}, min cosine vs default {float(cos.min()):.8f}")
clips = make_clips(41, [37, 64, 10, 150], "colour")
want = torch.stack(fingerprint_clips(sd, clips))
got = m.fingerprint_clips(clips).cpu()
c2 = torch.nn.functional.cosine_similarity(want.double(), got.double(), dim=1)
print(f"min cosine vs oracle {float(c2.min()):.8f}")
assert err == 0 and float(cos.min()) > 0.99999 and float(c2.min()) >= 0.9999

"""Dev helper: accuracy of the three stem paths against the fp32 oracle on the stress weights."""
import torch
from oracle.forward_oracle import forward_oracle
from oracle.weights import make_clips, make_state_dict
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native

lib = _native.load()
sd = make_state_dict(2, "stress")
m = vfp.create_model("attention").eval()
m.load_state_dict(sd)
clips = make_clips(9, [24] * 24, "colour")
x = torch.stack(clips)
want = forward_oracle(sd, x).double()
u8 = torch.round(x * 255).to(torch.uint8).cuda()
def cos(a, b):
    a = a.double(); b = b.double()
    return ((a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1)))
outs = {}
for mode in (0, 1, 2):
    lib.vfp_set_tuning(1, mode)
    outs[mode] = m(u8).cpu()
    c = cos(outs[mode], want)
    print(f"mode {mode}: cos vs oracle min {c.min():.7f} mean {c.mean():.7f}")
e_f32 = m(x.cuda()).cpu()
print("fp32-frame path vs oracle", float(cos(e_f32, want).min()))
for mode in (0, 1, 2):
    print(f"mode {mode} vs fp32-frame path: min {cos(outs[mode], e_f32).min():.7f};  vs mode 0: {cos(outs[mode], outs[0]).min():.7f}")
